"""Sparse gradient exchange kernels (csrc/lgr_sparse.cuh) on ONE GPU: with LGR_SPARSE_SINGLE=1 the fused backward runs
flag -> scan -> index -> per-Gaussian backward on the compacted list -> accumulate (world = 1, the rank's own buffer as the only "peer")
instead of the dense K7+K8.  Same gradients as the dense kernel up to the run-to-run rounding of the blend-backward atomics, exact zeros in
the same rows; the 2-GPU run over NVLink symmetric memory is in tests/test_gpu_multi.py."""
import os

import numpy as np
import pytest
import torch

from lightgaussian_b200.model import GaussianParams, TorchCamera, pipeline_params
from lightgaussian_b200.renderer import render
from lightgaussian_b200.synth import make_scene, make_cameras

pytestmark = pytest.mark.gpu
NAMES = ["_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity"]


def _grads(pc, cam, pipe, bg, target, sparse):
    os.environ["LGR_SPARSE_SINGLE"] = "1" if sparse else "0"
    try:
        for p in pc.parameters():
            p.grad = None
        screen = None
        pkg = render(cam, pc, pipe, bg)
        screen = pkg["viewspace_points"]
        screen.retain_grad()
        (pkg["render"] - target).abs().mean().backward()
        return {n: getattr(pc, n).grad.clone() for n in NAMES}, screen.grad.clone()
    finally:
        os.environ["LGR_SPARSE_SINGLE"] = "0"


@pytest.mark.parametrize("P,W,H,deg,scale", [(20000, 320, 240, 3, 1.5), (200000, 640, 480, 3, 1.0), (5000, 160, 120, 2, 2.0), (777, 100, 75, 1, 3.0)])
def test_sparse_path_equals_dense_path(P, W, H, deg, scale):
    scene = make_scene(P, sh_degree=deg, seed=P, scale_mult=scale)
    pc = GaussianParams(scene["raw"], deg, "cuda")
    cams = [TorchCamera(c, "cuda") for c in make_cameras(3, W, H)]
    pipe, bg = pipeline_params(), torch.zeros(3, device="cuda")
    target = torch.rand(3, H, W, device="cuda")
    for cam in cams[:2]:
        dense, g2d_d = _grads(pc, cam, pipe, bg, target, sparse=False)
        sparse, g2d_s = _grads(pc, cam, pipe, bg, target, sparse=True)
        nz_dense = dense["_opacity"].reshape(-1) != 0
        for n in NAMES:
            d, s = dense[n].reshape(P, -1), sparse[n].reshape(P, -1)
            assert torch.isfinite(s).all()
            assert float((d - s).abs().max()) <= 2e-4 * float(d.abs().max()) + 1e-12, n     # atomics reorder between two runs of the blend backward
            assert bool((s[~nz_dense] == 0).all()), n                                         # rows without a gradient are exact zeros in both
        assert float((g2d_d - g2d_s).abs().max()) <= 2e-4 * float(g2d_d.abs().max()) + 1e-12
        frac = float(nz_dense.float().mean())
        assert 0.0 < frac < 1.0


def test_sparse_path_with_no_visible_gaussian():
    scene = make_scene(1000, sh_degree=3, seed=1)
    scene["raw"]["xyz"][:] += 100.0                          # everything behind / outside the frustum
    pc = GaussianParams(scene["raw"], 3, "cuda")
    cam = TorchCamera(make_cameras(1, 64, 48)[0], "cuda")
    g, g2d = _grads(pc, cam, pipeline_params(), torch.zeros(3, device="cuda"), torch.rand(3, 48, 64, device="cuda"), sparse=True)
    for n in NAMES:
        assert float(g[n].abs().max()) == 0.0
    assert float(g2d.abs().max()) == 0.0
