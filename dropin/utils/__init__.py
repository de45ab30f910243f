"""Overlay of the reference's `utils` package (a namespace package without __init__.py): this regular package is found first
(dropin/ precedes the reference checkout on sys.path), serves `utils.loss_utils` from here (fused image loss) and extends its
__path__ with the reference's `utils/` directory so that every other submodule (`utils.general_utils`, `utils.sh_utils`, ...)
still resolves to the reference's own, unmodified file."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _cand = os.path.join(_p or ".", "utils")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.exists(os.path.join(_cand, "general_utils.py")):
        __path__.append(os.path.abspath(_cand))
        break
