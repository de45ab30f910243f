"""Drop-in for the reference's vectree/utils.py: its own functions (read_ply_data, write_ply_data, dec2bin, bin2dec) are re-exported from
the reference's file unchanged; `load_vqgaussian` is lightgaussian_b200.vectree.load_vqgaussian (same files, same result, indices unpacked and
codebook rows gathered on the GPU)."""
import importlib.util
import os

from lightgaussian_b200.vectree import load_vqgaussian  # noqa: F401

_ref = None
for _d in __import__("vectree").__path__:
    _f = os.path.join(_d, "utils.py")
    if os.path.exists(_f) and os.path.abspath(_f) != os.path.abspath(__file__):
        _spec = importlib.util.spec_from_file_location("_reference_vectree_utils", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        break
if _ref is not None:
    read_ply_data, write_ply_data, dec2bin, bin2dec = _ref.read_ply_data, _ref.write_ply_data, _ref.dec2bin, _ref.bin2dec
