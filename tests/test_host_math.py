"""The PRODUCT's per-Gaussian math (lightgaussian_b200/csrc/lgr_math.cuh) compiled for the host and compared with
the oracle -- catches arithmetic slips in the kernels' math without a GPU.  (The kernels themselves are covered
by the -m gpu tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.lgo import Oracle
from tests.util import make_config, view_from_camera, rel_inf
from lightgaussian_b200.synth import make_scene, make_cameras, inside_camera

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "math_host.cpp")
LIB = os.path.join(HERE, "native", "_build", "libmath_host.so")
HDR = os.path.join(os.path.dirname(HERE), "lightgaussian_b200", "csrc", "lgr_math.cuh")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17", SRC, "-o", LIB])
    return C.CDLL(LIB)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _host_preprocess(lib, v, a, M=16):
    P = a["means3D"].shape[0]
    out = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
               cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
               clamped=np.zeros(P, np.uint8), tiles_touched=np.zeros(P, np.uint32))
    lib.mh_preprocess.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2 + [C.c_float] + [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 8
    lib.mh_preprocess(P, v.sh_degree, M, _p(a["means3D"]), _p(a["scales"]), v.scale_modifier, _p(a["rotations"]),
                      _p(np.ascontiguousarray(a["opacities"].reshape(-1))), _p(a["shs"]), None, None, _p(v.viewmatrix), _p(v.projmatrix),
                      _p(v.campos), v.W, v.H, v.tanfovx, v.tanfovy,
                      *[_p(out[k]) for k in ("radii", "means2D", "depths", "cov3D", "rgb", "conic_opacity", "clamped", "tiles_touched")])
    return out


@pytest.mark.parametrize("case", ["sphere400", "inside", "hd_deg1"])
def test_forward_math_is_bit_identical_to_oracle(lib, case):
    if case == "sphere400":
        cam, deg, mod, P = make_cameras(4, 400, 400)[1], 3, 1.0, 20000
    elif case == "inside":
        cam, deg, mod, P = inside_camera(403, 277), 2, 1.7, 20000
    else:
        cam, deg, mod, P = make_cameras(7, 1920, 1080)[3], 1, 1.0, 5000
    a = make_scene(P, seed=3)["act"]
    v = view_from_camera(cam, (0, 0, 0), deg, mod)
    g = Oracle().preprocess(v, a["means3D"], a["opacities"], shs=a["shs"], scales=a["scales"], rotations=a["rotations"])
    h = _host_preprocess(lib, v, a)
    vis = g["radii"] > 0
    assert vis.sum() > 100
    for k in ("radii", "means2D", "depths", "rgb", "conic_opacity", "tiles_touched"):
        np.testing.assert_array_equal(h[k], g[k], err_msg=k)
    np.testing.assert_array_equal(h["cov3D"][vis], g["cov3D"][vis])
    bits = g["clamped"][:, 0] | (g["clamped"][:, 1] << 1) | (g["clamped"][:, 2] << 2)
    np.testing.assert_array_equal(h["clamped"], bits)


def test_backward_math_matches_oracle(lib):
    act, view, dpix = make_config("inside")
    o = Oracle()
    f = o.forward(view, act["means3D"], act["opacities"], shs=act["shs"], scales=act["scales"], rotations=act["rotations"])
    b = o.backward(view, f, dpix, act["means3D"], shs=act["shs"], scales=act["scales"], rotations=act["rotations"])
    P, M = act["means3D"].shape[0], 16
    geo = f["geom"]
    bits = (geo["clamped"][:, 0] | (geo["clamped"][:, 1] << 1) | (geo["clamped"][:, 2] << 2)).astype(np.uint8)
    out = dict(dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32), dL_dsh=np.zeros((P, M, 3), np.float32),
               dL_dscales=np.zeros((P, 3), np.float32), dL_drotations=np.zeros((P, 4), np.float32))
    lib.mh_preprocess_backward.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_float] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 8
    dm2 = np.ascontiguousarray(b["dL_dmeans2D"][:, :2])
    lib.mh_preprocess_backward(P, view.sh_degree, M, _p(act["means3D"]), _p(geo["radii"]), _p(act["shs"]), _p(bits), _p(act["scales"]),
                               _p(act["rotations"]), view.scale_modifier, _p(geo["cov3D"]), _p(view.viewmatrix), _p(view.projmatrix),
                               _p(view.campos), view.W, view.H, view.tanfovx, view.tanfovy, _p(dm2), _p(b["dL_dconic"]), _p(b["dL_dcolors"]),
                               _p(out["dL_dmeans3D"]), _p(out["dL_dcov3D"]), _p(out["dL_dsh"]), _p(out["dL_dscales"]), _p(out["dL_drotations"]))
    for k in out:
        assert rel_inf(out[k], b[k]) < 2e-5, (k, rel_inf(out[k], b[k]))


def test_exchange_ranges_cover_all_gaussians_on_block_boundaries(monkeypatch):
    """host logic of the view-parallel exchange: ranges handed to lgr_backward_raw_end_range start at multiples of 256 and tile [0, P)"""
    from lightgaussian_b200 import rasterizer
    for chunks in ("1", "2", "4", "7"):
        monkeypatch.setenv("LGR_EXCHANGE_CHUNKS", chunks)
        for P in (1, 255, 256, 257, 1000, 65536, 3_000_000):
            r = rasterizer._exchange_chunks(P)
            assert r[0][0] == 0 and sum(n for _, n in r) == P and len(r) <= int(chunks)
            assert all(f % 256 == 0 and n > 0 for f, n in r)
            assert all(r[k][0] + r[k][1] == r[k + 1][0] for k in range(len(r) - 1))
    assert rasterizer.exchange_info(1) == ("none", None)
