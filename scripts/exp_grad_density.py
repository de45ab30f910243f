"""GPU experiment: what fraction of the Gaussians receives a non-zero gradient from ONE view of the bench scene?
(visible ones whose pixels saturate before reaching them get exact zeros) -- decides whether a sparse gradient exchange could pay."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightgaussian_b200.model import GaussianParams, TorchCamera, pipeline_params  # noqa: E402
from lightgaussian_b200.renderer import render  # noqa: E402
from lightgaussian_b200.synth import make_scene, make_cameras  # noqa: E402

P, W, H = 3_000_000, 1920, 1080
scene = make_scene(P, sh_degree=3, seed=0)
pc = GaussianParams(scene["raw"], 3, "cuda")
cams = [TorchCamera(c, "cuda") for c in make_cameras(16, W, H)]
pipe, bg = pipeline_params(), torch.zeros(3, device="cuda")
tgt = torch.rand(3, H, W, device="cuda")
for i in (0, 5, 11):
    for p in pc.parameters():
        p.grad = None
    pkg = render(cams[i], pc, pipe, bg)
    (pkg["render"] - tgt).abs().mean().backward()
    vis = (pkg["radii"] > 0).float().mean().item()
    nz_dc = (pc._features_dc.grad.reshape(P, -1).abs().sum(1) > 0).float().mean().item()
    nz_xyz = (pc._xyz.grad.abs().sum(1) > 0).float().mean().item()
    nz_op = (pc._opacity.grad.abs().sum(1) > 0).float().mean().item()
    print(f"view {i}: visible {vis:.3f}  nonzero dRGB rows {nz_dc:.3f}  nonzero xyz rows {nz_xyz:.3f}  nonzero opacity rows {nz_op:.3f}")
