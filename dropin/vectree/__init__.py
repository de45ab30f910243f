"""Overlay of the reference's `vectree` directory as imported from the repository root (scene/gaussian_model.py:24 does
`from vectree.utils import load_vqgaussian, write_ply_data`): `vectree.utils` is served from here (GPU index unpacking and gather),
every other submodule (`vectree.vq`, `vectree.vectree`) still resolves to the reference's own file."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _cand = os.path.join(_p or ".", "vectree")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.exists(os.path.join(_cand, "vq.py")):
        __path__.append(os.path.abspath(_cand))
        break
