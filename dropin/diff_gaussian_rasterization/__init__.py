"""Drop-in for the reference's `diff_gaussian_rasterization` package: put <repo>/dropin and <repo> on
PYTHONPATH ahead of the reference checkout and its scripts import this instead of the pybind extension."""
from lightgaussian_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    _RasterizeGaussians,
    _C,
)
