"""Drop-in for the reference's `gaussian_renderer` package (render, count_render, network_gui, GaussianModel)."""
from lightgaussian_b200.renderer import render, count_render  # noqa: F401
from . import network_gui  # noqa: F401

import os as _os

try:  # render.py:22 / render_video.py:23 do `from gaussian_renderer import GaussianModel`
    from scene.gaussian_model import GaussianModel  # noqa: F401
except Exception:  # reference checkout not on the path: the name is simply absent
    GaussianModel = None

if GaussianModel is not None and _os.environ.get("LGR_FUSED_OPTIM", "1") != "0":
    # row N3: the AdamW built by GaussianModel.training_setup becomes FusedAdamW (bit-identical updates, one launch per step)
    # and prune_points uses the fused compaction; LGR_FUSED_OPTIM=0 keeps torch.optim.AdamW and the reference's surgery.
    from lightgaussian_b200 import optim as _optim
    _optim.install(GaussianModel)
if GaussianModel is None:
    del GaussianModel
