"""Deterministic synthetic scenes and cameras for tests and bench.py.

There is no dataset or checkpoint on the GPU box, so every measurement uses a seeded random
Gaussian cloud and cameras on a Fibonacci sphere.  Camera matrices follow the reference's
convention exactly (scene/cameras.py:70-85, utils/graphics_utils.py:38-77): `world_view_transform`
and `full_proj_transform` are stored TRANSPOSED (row-vector convention), `camera_center` is row 3
of the inverse view transform.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class SynthCamera:
    """The attributes gaussian_renderer.render() reads from a viewpoint camera
    (reference gaussian_renderer/__init__.py:49-62), as numpy arrays."""
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: np.ndarray  # [4,4] float32, transposed W2C
    full_proj_transform: np.ndarray   # [4,4] float32, transposed (Proj @ W2C)
    camera_center: np.ndarray         # [3] float32
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """Same frustum as utils/graphics_utils.py:51-77 (float32 arithmetic like torch.zeros(4,4))."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    bottom, left = -top, -right
    Pm = np.zeros((4, 4), np.float32)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def camera_from_pose(R_w2c: np.ndarray, t: np.ndarray, W: int, H: int, fovx: float, znear=0.01, zfar=100.0) -> SynthCamera:
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
    Rt = np.zeros((4, 4), np.float64)
    Rt[:3, :3] = R_w2c
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    wvt = np.float32(Rt).T.copy()                       # transposed storage
    proj_t = projection_matrix(znear, zfar, fovx, fovy).T.copy()
    full = (wvt @ proj_t).astype(np.float32)            # == (Proj @ W2C)^T
    center = np.linalg.inv(wvt.astype(np.float64))[3, :3].astype(np.float32)
    return SynthCamera(W, H, fovx, fovy, np.ascontiguousarray(wvt), np.ascontiguousarray(full), np.ascontiguousarray(center),
                       znear, zfar)


def look_at(cam_pos, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """W2C rotation / translation for a camera with x right, y down, z forward (COLMAP convention)."""
    c = np.asarray(cam_pos, np.float64)
    f = np.asarray(target, np.float64) - c
    f /= np.linalg.norm(f)
    u = np.asarray(up, np.float64)
    if abs(float(f @ u)) > 0.99:
        u = np.array([0.0, 1.0, 0.0])
    x = np.cross(f, u)
    x /= np.linalg.norm(x)
    y = np.cross(f, x)
    R = np.stack([x, y, f])
    return R, -R @ c


def make_cameras(n: int, W: int, H: int, fovx_deg: float = 60.0, radius: float = 3.0) -> List[SynthCamera]:
    """n poses on a Fibonacci sphere of the given radius, looking at the origin, up = +z."""
    cams = []
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n):
        z = 1.0 - 2.0 * (i + 0.5) / n
        r = math.sqrt(max(0.0, 1.0 - z * z))
        th = golden * i
        pos = radius * np.array([r * math.cos(th), r * math.sin(th), z])
        R, t = look_at(pos)
        cams.append(camera_from_pose(R, t, W, H, math.radians(fovx_deg)))
    return cams


def inside_camera(W: int, H: int, fovx_deg: float = 60.0) -> SynthCamera:
    """Correctness-only camera INSIDE the cloud, looking +x: exercises the z<=0.2 cull, the 1.3*tanfov
    clamp and very large splats."""
    R, t = look_at((0.3, 0.2, 0.1), target=(1.3, 0.2, 0.1))
    return camera_from_pose(R, t, W, H, math.radians(fovx_deg))


def make_scene(P: int, sh_degree: int = 3, seed: int = 0, scale_mult: float = 1.0) -> dict:
    """Raw (pre-activation) parameters with GaussianModel's layout (scene/gaussian_model.py:98-118) and the
    activated tensors render() feeds the rasterizer.  All float32, C-contiguous numpy."""
    rng = np.random.default_rng(seed)
    M = (sh_degree + 1) ** 2
    s0 = 0.6 * P ** (-1.0 / 3.0) * scale_mult
    raw = dict(
        xyz=rng.uniform(-1.0, 1.0, (P, 3)).astype(np.float32),
        scaling=(math.log(s0) + 0.5 * rng.standard_normal((P, 3))).astype(np.float32),
        rotation=rng.standard_normal((P, 4)).astype(np.float32),
        opacity=(2.0 * rng.standard_normal((P, 1))).astype(np.float32),
        features_dc=rng.standard_normal((P, 1, 3)).astype(np.float32),
        features_rest=(0.2 * rng.standard_normal((P, M - 1, 3))).astype(np.float32),
    )
    rot = raw["rotation"]
    act = dict(
        means3D=raw["xyz"],
        scales=np.exp(raw["scaling"]).astype(np.float32),
        rotations=(rot / np.sqrt((rot * rot).sum(1, keepdims=True))).astype(np.float32),
        opacities=(1.0 / (1.0 + np.exp(-raw["opacity"]))).astype(np.float32),
        shs=np.ascontiguousarray(np.concatenate([raw["features_dc"], raw["features_rest"]], axis=1)),
    )
    return dict(raw=raw, act=act, sh_degree=sh_degree, P=P, M=M)
