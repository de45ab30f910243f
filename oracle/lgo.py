"""numpy front-end of the CPU oracle (oracle/lgo.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product package (lightgaussian_b200/) never does.

The call structure mirrors the reference's CudaRasterizer::Rasterizer::{forward, forwardCount,
backward} (RAST/cuda_rasterizer/rasterizer_impl.cu:198-337, 441-584, 341-435) but every stage is
exposed separately so tests can check the CUDA path stage by stage.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force: bool = False) -> None:
    """Compile liblgo.so / liblgo_d.so (gcc, a second or two)."""
    out = os.path.join(_HERE, "_build", "liblgo.so")
    src = os.path.join(_HERE, "lgo.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def _lib(double: bool = False):
    key = "d" if double else "f"
    if key not in _LIBS:
        build()
        name = "liblgo_d.so" if double else "liblgo.so"
        _LIBS[key] = C.CDLL(os.path.join(_HERE, "_build", name))
    return _LIBS[key]


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class View:
    """Per-view constants == the numeric fields of GaussianRasterizationSettings
    (RAST/diff_gaussian_rasterization/__init__.py:248-261)."""
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] row-major, transposed convention (cameras.py:70-72)
    projmatrix: np.ndarray   # [4,4]
    campos: np.ndarray       # [3]
    bg: np.ndarray           # [3]
    sh_degree: int = 3
    scale_modifier: float = 1.0


class Oracle:
    """dtype float32 (bit-faithful restatement) or float64 (for finite-difference checks)."""

    def __init__(self, double: bool = False):
        self.double = double
        self.lib = _lib(double)
        self.dt = np.float64 if double else np.float32
        self.creal = C.c_double if double else C.c_float
        self.pfx = "lgo_d_" if double else "lgo_"

    def _f(self, name):
        fn = getattr(self.lib, self.pfx + name)
        return fn

    def _a(self, x, shape=None):
        if x is None:
            return None
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dt))
        if shape is not None:
            a = a.reshape(shape)
        return a

    # -- stages ---------------------------------------------------------------------------
    def preprocess(self, v: View, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                   cov3D_precomp=None):
        means3D = self._a(means3D)
        P = means3D.shape[0]
        shs = self._a(shs)
        M = 0 if shs is None else shs.shape[1]
        opacities = self._a(opacities).reshape(-1)
        scales, rotations = self._a(scales), self._a(rotations)
        cov3D_precomp, colors_precomp = self._a(cov3D_precomp), self._a(colors_precomp)
        out = dict(
            radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), self.dt), depths=np.zeros(P, self.dt),
            cov3D=np.zeros((P, 6), self.dt), rgb=np.zeros((P, 3), self.dt), conic_opacity=np.zeros((P, 4), self.dt),
            clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
        fn = self._f("preprocess")
        fn.restype = None
        r = self.creal
        fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, r, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, r, r] + [C.c_void_p] * 8
        fn(P, int(v.sh_degree), M, _p(means3D), _p(scales), v.scale_modifier, _p(rotations), _p(opacities), _p(shs),
           _p(cov3D_precomp), _p(colors_precomp), _p(self._a(v.viewmatrix)), _p(self._a(v.projmatrix)), _p(self._a(v.campos)),
           v.W, v.H, v.tanfovx, v.tanfovy, _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]), _p(out["cov3D"]),
           _p(out["rgb"]), _p(out["conic_opacity"]), _p(out["clamped"]), _p(out["tiles_touched"]))
        return out

    def mark_visible(self, v: View, means3D):
        means3D = self._a(means3D)
        P = means3D.shape[0]
        present = np.zeros(P, np.uint8)
        fn = self._f("mark_visible")
        fn.restype = None
        fn.argtypes = [C.c_int] + [C.c_void_p] * 4
        fn(P, _p(means3D), _p(self._a(v.viewmatrix)), _p(self._a(v.projmatrix)), _p(present))
        return present.astype(bool)

    def bin(self, v: View, means2D, depths, radii, tiles_touched=None):
        means2D, depths = self._a(means2D), self._a(depths)
        radii = np.ascontiguousarray(radii, dtype=np.int32)
        P = radii.shape[0]
        gx, gy = (v.W + 15) // 16, (v.H + 15) // 16
        if tiles_touched is None:
            # recompute the rectangle areas the same way the C side does, via a dry run
            R = self._count_instances(v, means2D, depths, radii)
        else:
            R = int(np.asarray(tiles_touched, dtype=np.int64).sum())
        point_list = np.zeros(max(R, 1), np.uint32)
        ranges = np.zeros((gx * gy, 2), np.uint32)
        fn = self._f("bin")
        fn.restype = C.c_int64
        fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        R2 = fn(P, _p(means2D), _p(depths), _p(radii), v.W, v.H, _p(point_list), _p(ranges))
        assert R2 == R, (R2, R)
        return point_list[:R], ranges

    def _count_instances(self, v, means2D, depths, radii):
        # sum of rectangle areas; python restatement of auxiliary.h:46-56 in the array dtype
        gx, gy = (v.W + 15) // 16, (v.H + 15) // 16
        vis = radii > 0
        rf = radii.astype(self.dt)
        px, py = means2D[:, 0], means2D[:, 1]
        f = self.dt
        x0 = np.clip(((px - rf) * f(0.0625)).astype(np.int64), 0, gx)
        y0 = np.clip(((py - rf) * f(0.0625)).astype(np.int64), 0, gy)
        x1 = np.clip(((((px + rf) + f(16)) - f(1)) * f(0.0625)).astype(np.int64), 0, gx)
        y1 = np.clip(((((py + rf) + f(16)) - f(1)) * f(0.0625)).astype(np.int64), 0, gy)
        return int((((x1 - x0) * (y1 - y0))[vis]).sum())

    def blend_forward(self, v: View, ranges, point_list, means2D, colors, conic_opacity, count=None, want_fragile=False):
        N = v.W * v.H
        out_color = np.zeros((3, v.H, v.W), self.dt)
        final_T = np.zeros(N, self.dt)
        n_contrib = np.zeros(N, np.uint32)
        fragile = np.zeros(N, np.uint8) if want_fragile else None
        fn = self._f("blend_forward")
        fn.restype = None
        fn.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 11
        pl = np.ascontiguousarray(point_list, dtype=np.uint32)
        if pl.size == 0:
            pl = np.zeros(1, np.uint32)
        fn(v.W, v.H, _p(np.ascontiguousarray(ranges, dtype=np.uint32)), _p(pl), _p(self._a(means2D)), _p(self._a(colors)),
           _p(self._a(conic_opacity)), _p(self._a(v.bg)), _p(out_color), _p(final_T), _p(n_contrib), _p(count), _p(fragile))
        return dict(color=out_color, final_T=final_T, n_contrib=n_contrib,
                    fragile=None if fragile is None else fragile.reshape(v.H, v.W).astype(bool))

    def blend_backward(self, v: View, P, ranges, point_list, means2D, conic_opacity, colors, final_T, n_contrib, dL_dpix):
        out = dict(dL_dmean2D=np.zeros((P, 2), self.dt), dL_dconic=np.zeros((P, 3), self.dt),
                   dL_dopacity=np.zeros(P, self.dt), dL_dcolor=np.zeros((P, 3), self.dt))
        fn = self._f("blend_backward")
        fn.restype = None
        fn.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 13
        pl = np.ascontiguousarray(point_list, dtype=np.uint32)
        if pl.size == 0:
            pl = np.zeros(1, np.uint32)
        fn(P, v.W, v.H, _p(np.ascontiguousarray(ranges, dtype=np.uint32)), _p(pl), _p(self._a(means2D)),
           _p(self._a(conic_opacity)), _p(self._a(colors)), _p(self._a(v.bg)), _p(self._a(final_T)),
           _p(np.ascontiguousarray(n_contrib, dtype=np.uint32)), _p(self._a(dL_dpix)), _p(out["dL_dmean2D"]),
           _p(out["dL_dconic"]), _p(out["dL_dopacity"]), _p(out["dL_dcolor"]))
        return out

    def preprocess_backward(self, v: View, means3D, radii, clamped, cov3D, dL_dmean2D, dL_dconic, dL_dcolor, shs=None,
                            scales=None, rotations=None):
        means3D = self._a(means3D)
        P = means3D.shape[0]
        shs = self._a(shs)
        M = 0 if shs is None else shs.shape[1]
        out = dict(dL_dmeans3D=np.zeros((P, 3), self.dt), dL_dcov3D=np.zeros((P, 6), self.dt),
                   dL_dsh=np.zeros((P, M, 3), self.dt), dL_dscales=np.zeros((P, 3), self.dt),
                   dL_drotations=np.zeros((P, 4), self.dt))
        fn = self._f("preprocess_backward")
        fn.restype = None
        r = self.creal
        fn.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6 + [r] + [C.c_void_p] * 4 + [C.c_int, C.c_int, r, r] + \
                      [C.c_void_p] * 8
        fn(P, int(v.sh_degree), M, _p(means3D), _p(np.ascontiguousarray(radii, dtype=np.int32)), _p(shs),
           _p(np.ascontiguousarray(clamped, dtype=np.uint8)), _p(self._a(scales)), _p(self._a(rotations)), v.scale_modifier,
           _p(self._a(cov3D)), _p(self._a(v.viewmatrix)), _p(self._a(v.projmatrix)), _p(self._a(v.campos)), v.W, v.H,
           v.tanfovx, v.tanfovy, _p(self._a(dL_dmean2D)), _p(self._a(dL_dconic)), _p(self._a(dL_dcolor)),
           _p(out["dL_dmeans3D"]), _p(out["dL_dcov3D"]), _p(out["dL_dsh"]), _p(out["dL_dscales"]), _p(out["dL_drotations"]))
        return out

    # -- whole passes (same composition as Rasterizer::forward / forwardCount / backward) ---
    def forward(self, v: View, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, count=False, want_fragile=False):
        geom = self.preprocess(v, means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp)
        point_list, ranges = self.bin(v, geom["means2D"], geom["depths"], geom["radii"], geom["tiles_touched"])
        colors = self._a(colors_precomp) if colors_precomp is not None else geom["rgb"]
        P = geom["radii"].shape[0]
        cnt = np.zeros(P, np.int64) if count else None
        img = self.blend_forward(v, ranges, point_list, geom["means2D"], colors, geom["conic_opacity"], cnt, want_fragile)
        res = dict(geom=geom, point_list=point_list, ranges=ranges, colors=colors, num_rendered=int(point_list.shape[0]), **img)
        if count:
            res["gaussians_count"] = cnt
            # exact-sum definition of the significance score: opacity * count
            res["important_score"] = (cnt.astype(self.dt) * np.asarray(opacities, dtype=self.dt).reshape(-1)).astype(self.dt)
        return res

    def backward(self, v: View, fwd, dL_dpix, means3D, shs=None, scales=None, rotations=None, cov3D_precomp=None):
        geom = fwd["geom"]
        P = geom["radii"].shape[0]
        g2 = self.blend_backward(v, P, fwd["ranges"], fwd["point_list"], geom["means2D"], geom["conic_opacity"], fwd["colors"],
                                 fwd["final_T"], fwd["n_contrib"], dL_dpix)
        cov3D = self._a(cov3D_precomp) if cov3D_precomp is not None else geom["cov3D"]
        g3 = self.preprocess_backward(v, means3D, geom["radii"], geom["clamped"], cov3D, g2["dL_dmean2D"], g2["dL_dconic"],
                                      g2["dL_dcolor"], shs, scales, rotations)
        means2D_grad = np.zeros((P, 3), self.dt)
        means2D_grad[:, :2] = g2["dL_dmean2D"]
        return dict(dL_dmeans2D=means2D_grad, dL_dcolors=g2["dL_dcolor"], dL_dopacity=g2["dL_dopacity"].reshape(P, 1),
                    dL_dconic=g2["dL_dconic"], **g3)
