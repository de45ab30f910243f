"""render()'s fused-activation path (raw leaves + TMA-staged SH rows) against the reference-compatible path
(torch activations + cat -> rasterizer) of the SAME library: images / radii / counts must be bit-identical,
leaf gradients equal up to the atomics' summation order."""
import os

import numpy as np
import pytest
import torch

from lightgaussian_b200.model import GaussianParams, TorchCamera, pipeline_params
from lightgaussian_b200.synth import make_scene, make_cameras, inside_camera
from tests.util import rel_inf, rel_l2

pytestmark = pytest.mark.gpu


def _render_both(P, W, H, deg_active, M, cam_np, seed, scale_mult=1.5, count=False):
    from lightgaussian_b200 import renderer
    scene = make_scene(P, sh_degree=3, seed=seed, scale_mult=scale_mult)
    raw = dict(scene["raw"])
    raw["features_rest"] = np.ascontiguousarray(raw["features_rest"][:, :M - 1])
    cam = TorchCamera(cam_np, "cuda")
    bg = torch.tensor([0.2, 0.4, 0.6], device="cuda")
    pipe = pipeline_params()
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    out = {}
    for mode in ("fused", "plain"):
        os.environ["LGR_FUSED"] = "1" if mode == "fused" else "0"
        pc = GaussianParams(raw, 3, "cuda")
        pc.active_sh_degree = deg_active
        fn = renderer.count_render if count else renderer.render
        pkg = fn(cam, pc, pipe, bg)
        res = {k: v.detach().cpu().numpy() for k, v in pkg.items() if k != "viewspace_points"}
        if not count:
            ((pkg["render"] - tgt) ** 2).mean().backward()
            res["grads"] = [p.grad.detach().cpu().numpy() for p in pc.parameters()]
            res["g2d"] = pkg["viewspace_points"].grad.detach().cpu().numpy()
        out[mode] = res
    os.environ["LGR_FUSED"] = "1"
    return out


def test_inkernel_activations_match_this_torch_build():
    from lightgaussian_b200.rasterizer import fused_activations_match_torch
    assert fused_activations_match_torch(torch.device("cuda", 0)), \
        "exp / sigmoid / F.normalize of this torch build no longer match lgr_raw.cuh's operation order"


@pytest.mark.parametrize("P,W,H,deg,M", [(4096, 160, 120, 3, 16), (1000 + 13, 96, 80, 3, 16), (2048 + 7, 128, 96, 2, 16),
                                         (1500, 96, 64, 1, 16), (1200, 64, 64, 0, 16), (3000 + 5, 128, 96, 2, 9), (777, 64, 48, 1, 4)])
def test_fused_equals_plain(P, W, H, deg, M):
    cam = make_cameras(5, W, H)[2]
    o = _render_both(P, W, H, deg, M, cam, seed=100 + P)
    f, p = o["fused"], o["plain"]
    assert (p["radii"] > 0).sum() > 20
    np.testing.assert_array_equal(f["radii"], p["radii"])
    np.testing.assert_array_equal(f["render"], p["render"])          # bit-identical image
    for a, b, name in zip(f["grads"], p["grads"], ("xyz", "dc", "rest", "scaling", "rotation", "opacity")):
        assert a.shape == b.shape and np.isfinite(a).all()
        assert rel_inf(a, b) <= 1e-3, f"{name}: {rel_inf(a, b)}"
        assert rel_l2(a, b) <= 1e-4, f"{name}: {rel_l2(a, b)}"
    assert rel_inf(f["g2d"], p["g2d"]) <= 1e-3


def test_fused_equals_plain_with_heavy_culling():
    W, H = 160, 112
    o = _render_both(5000, W, H, 3, 16, inside_camera(W, H), seed=9, scale_mult=1.0)
    f, p = o["fused"], o["plain"]
    assert 0 < (p["radii"] > 0).sum() < 2500
    np.testing.assert_array_equal(f["render"], p["render"])
    for a, b in zip(f["grads"], p["grads"]):
        assert rel_inf(a, b) <= 1e-3
        assert np.all(a[p["radii"] <= 0] == 0)


def test_fused_count_render_equals_plain():
    W, H = 128, 96
    o = _render_both(3000 + 11, W, H, 3, 16, make_cameras(5, W, H)[1], seed=4, count=True)
    f, p = o["fused"], o["plain"]
    np.testing.assert_array_equal(f["render"], p["render"])
    np.testing.assert_array_equal(f["gaussians_count"], p["gaussians_count"])
    np.testing.assert_array_equal(f["important_score"], p["important_score"])
    assert f["gaussians_count"].sum() > 0


def test_compact_sh_gradient_rebuild_matches_dense():
    """lgr_backward_raw(compact) + lgr_sh_grad_from_views == dense SH gradients, for one view bit for bit and for the
    sum of two views up to float summation order (the exchange used by view-parallel training)."""
    from lightgaussian_b200 import rasterizer as R
    W, H, P = 160, 120, 3000 + 17
    scene = make_scene(P, sh_degree=3, seed=31, scale_mult=1.5)
    pc = GaussianParams(scene["raw"], 3, "cuda", requires_grad=False)
    leaves = (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity)
    bg = torch.zeros(3, device="cuda")
    dense, rgbs, cams = [], [], []
    for ci in (1, 3):
        cam = TorchCamera(make_cameras(5, W, H)[ci], "cuda")
        rs = R.GaussianRasterizationSettings(H, W, float(np.tan(cam.FoVx / 2)), float(np.tan(cam.FoVy / 2)), bg, 1.0, cam.world_view_transform,
                                             cam.full_proj_transform, 3, cam.camera_center, False, False, False)
        _, _, Rn, color, radii, geom, binning, img, lv = R._forward_raw_native(False, rs, *leaves)
        dpix = torch.randn(3, H, W, generator=torch.Generator().manual_seed(ci)).cuda()
        gd, _, _, _ = R.backward_raw_native(rs, Rn, dpix, *leaves, radii, geom, binning, img, compact=False)
        gc, _, d_rgb, flat = R.backward_raw_native(rs, Rn, dpix, *leaves, radii, geom, binning, img, compact=True)
        for a, b in ((gd[0], gc[0]), (gd[3], gc[3]), (gd[4], gc[4]), (gd[5], gc[5])):   # the small leaves are unaffected by the mode
            assert rel_inf(b.cpu().numpy(), a.cpu().numpy()) <= 1e-3
        d_dc1, d_rest1 = R.sh_grad_from_views(pc._xyz, cam.camera_center.reshape(1, 3), d_rgb.reshape(1, P, 3), pc._features_dc,
                                              pc._features_rest, 3)
        # same products, nothing to sum: identical up to the atomics' order inside dL/dRGB itself
        assert rel_inf(d_rest1.cpu().numpy(), gd[2].cpu().numpy()) <= 1e-3
        assert rel_inf(d_dc1.cpu().numpy(), gd[1].cpu().numpy()) <= 1e-3
        dense.append(gd)
        rgbs.append(d_rgb)
        cams.append(cam.camera_center)
    d_dc, d_rest = R.sh_grad_from_views(pc._xyz, torch.stack(cams), torch.stack(rgbs), pc._features_dc, pc._features_rest, 3)
    ref_rest = (dense[0][2] + dense[1][2]).cpu().numpy()
    ref_dc = (dense[0][1] + dense[1][1]).cpu().numpy()
    assert np.abs(ref_rest).max() > 0
    assert rel_inf(d_rest.cpu().numpy(), ref_rest) <= 1e-3 and rel_l2(d_rest.cpu().numpy(), ref_rest) <= 1e-4
    assert rel_inf(d_dc.cpu().numpy(), ref_dc) <= 1e-3


def _student_like(raw, device="cuda"):
    """GaussianParams whose _features_rest is exactly what GaussianModel.onedownSHdegree() leaves behind for 3 -> 2
    (scene/gaussian_model.py:129-136): a NON-contiguous [P,8,3] view (strides 45,3,1) of a detached [P,15,3] clone, requires_grad."""
    pc = GaussianParams(raw, 3, device)
    full = pc._features_rest.clone().detach()
    pc._features_rest = full[:, :8, :]
    pc._features_rest.requires_grad = True
    pc.max_sh_degree, pc.active_sh_degree = 2, 2
    return pc


@pytest.mark.parametrize("P", [3000 + 5, 4096, 33])
def test_strided_student_features_rest_takes_the_fused_path(P):
    """The distillation student's leaf is read in place through its row stride: same image bit for bit and same gradients as a
    contiguous copy of the same values; the gradient comes back dense [P,8,3]."""
    from lightgaussian_b200 import renderer, trace
    W, H = 128, 96
    scene = make_scene(P, sh_degree=3, seed=200 + P, scale_mult=1.5)
    cam = TorchCamera(make_cameras(5, W, H)[2], "cuda")
    bg = torch.tensor([0.1, 0.3, 0.5], device="cuda")
    pipe = pipeline_params()
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).cuda()
    stu = _student_like(scene["raw"])
    assert not stu._features_rest.is_contiguous() and stu._features_rest.stride() == (45, 3, 1)
    assert renderer._can_fuse(stu, pipe, None)
    n0 = trace.counters.get("render_fused_strided_rest", 0)
    pkg = renderer.render(cam, stu, pipe, bg)
    assert trace.counters.get("render_fused_strided_rest", 0) == n0 + 1
    ((pkg["render"] - tgt) ** 2).mean().backward()
    raw2 = dict(scene["raw"])
    raw2["features_rest"] = np.ascontiguousarray(raw2["features_rest"][:, :8])
    con = GaussianParams(raw2, 2, "cuda")
    pkg2 = renderer.render(cam, con, pipe, bg)
    ((pkg2["render"] - tgt) ** 2).mean().backward()
    assert torch.equal(pkg["render"], pkg2["render"]) and torch.equal(pkg["radii"], pkg2["radii"])
    g = stu._features_rest.grad
    assert g.shape == (P, 8, 3) and g.is_contiguous()
    for a, b, name in zip(stu.parameters(), con.parameters(), ("xyz", "dc", "rest", "scaling", "rotation", "opacity")):
        assert rel_inf(a.grad.cpu().numpy(), b.grad.cpu().numpy()) <= 1e-3, name
    # count mode on the strided leaf
    c1 = renderer.count_render(cam, stu, pipe, bg)
    c2 = renderer.count_render(cam, con, pipe, bg)
    assert torch.equal(c1["gaussians_count"], c2["gaussians_count"]) and torch.equal(c1["render"], c2["render"])


@pytest.mark.parametrize("mode", [1, 2])
def test_other_kback_modes_still_agree(mode):
    """lgr_set_kback_mode(1): the one-warp-per-32-Gaussians K7+K8 kernel (used for ranged / compact launches of the exchange paths);
    (2): zero-fill in its own kernel instead of inside the blend backward -- against the plain path, like the default above"""
    from lightgaussian_b200 import capi
    capi.set_kback_mode(mode)
    try:
        o = _render_both(3000 + 5, 128, 96, 3, 16, make_cameras(5, 128, 96)[2], seed=21)
    finally:
        capi.set_kback_mode(0)
    f, p = o["fused"], o["plain"]
    np.testing.assert_array_equal(f["render"], p["render"])
    for a, b in zip(f["grads"], p["grads"]):
        assert rel_inf(a, b) <= 1e-3 and rel_l2(a, b) <= 1e-4
