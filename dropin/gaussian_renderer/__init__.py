"""Drop-in for the reference's `gaussian_renderer` package (render, count_render, network_gui, GaussianModel)."""
from lightgaussian_b200.renderer import render, count_render  # noqa: F401
from . import network_gui  # noqa: F401

try:  # render.py:22 / render_video.py:23 do `from gaussian_renderer import GaussianModel`
    from scene.gaussian_model import GaussianModel  # noqa: F401
except Exception:  # reference checkout not on the path: the name is simply absent
    pass
