"""render() / count_render() with the reference's signatures and return dictionaries
(reference gaussian_renderer/__init__.py:22-124 and :127-229).  The two differ only in f_count and the two
extra outputs, so both go through one helper here.

`pc` is anything with GaussianModel's getters (scene/gaussian_model.py:98-118): get_xyz, get_opacity,
get_scaling, get_rotation, get_features, active_sh_degree, max_sh_degree and, for
pipe.compute_cov3D_python, get_covariance(scaling_modifier).
"""
from __future__ import annotations

import math

import torch

import os

import torch.nn.functional as F

from . import trace

from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_raw_leaves, fused_activations_match_torch,
                         rest_row_stride)

_LEAVES = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity")


def _can_fuse(pc, pipe, override_color) -> bool:
    """The fused path needs GaussianModel-style raw leaves with the standard activations
    (scene/gaussian_model.py:35-43: exp / sigmoid / normalize) and the default pipeline flags."""
    if os.environ.get("LGR_FUSED", "1") == "0" or override_color is not None:
        return False
    if pipe.convert_SHs_python or pipe.compute_cov3D_python:
        return False
    if not all(hasattr(pc, n) for n in _LEAVES):
        return False
    if getattr(pc, "scaling_activation", torch.exp) is not torch.exp:
        return False
    if getattr(pc, "opacity_activation", torch.sigmoid) is not torch.sigmoid:
        return False
    if getattr(pc, "rotation_activation", F.normalize) is not F.normalize:
        return False
    t = pc._xyz
    # every leaf dense, except that _features_rest may be a row-strided view: the distillation student's
    # `_features_rest[:, :8, :]` (scene/gaussian_model.py:129-136) is read in place through its row stride
    if not (t.is_cuda and all(getattr(pc, n).dtype == torch.float32 and (getattr(pc, n).is_contiguous() or n == "_features_rest")
                              for n in _LEAVES)):
        return False
    if pc._features_rest.dim() != 3 or pc._features_dc.shape[1] != 1 or pc._features_rest.shape[1] < 1:
        return False
    if not pc._features_rest.is_contiguous() and rest_row_stride(pc._features_rest) == 0:
        return False
    if (pc.active_sh_degree + 1) ** 2 > 1 + pc._features_rest.shape[1]:
        return False
    return fused_activations_match_torch(t.device)

_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def eval_sh_torch(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """PyTorch SH evaluation for pipe.convert_SHs_python (the role of utils/sh_utils.py:57-120).
    sh: [..., C, (max_deg+1)^2], dirs: [..., 3] unit vectors -> [..., C]."""
    out = _SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        out = out - _SH_C1 * y * sh[..., 1] + _SH_C1 * z * sh[..., 2] - _SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            out = (out + _SH_C2[0] * xy * sh[..., 4] + _SH_C2[1] * yz * sh[..., 5] + _SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                   + _SH_C2[3] * xz * sh[..., 7] + _SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                out = (out + _SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + _SH_C3[1] * xy * z * sh[..., 10]
                       + _SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                       + _SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + _SH_C3[5] * z * (xx - yy) * sh[..., 14]
                       + _SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return out


def _render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, f_count):
    xyz = pc.get_xyz
    # grad placeholder for the screen-space means, as the reference builds it (:37-46)
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=pipe.debug,
        f_count=f_count,
    )
    if _can_fuse(pc, pipe, override_color):
        trace.bump("render_fused")
        if not pc._features_rest.is_contiguous():
            trace.bump("render_fused_strided_rest")
        outputs = rasterize_raw_leaves(pc._xyz, screenspace_points, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation,
                                       pc._opacity, settings)
        return _package(outputs, screenspace_points, f_count)

    trace.bump("render_unfused")
    rasterizer = GaussianRasterizer(raster_settings=settings)

    geometry = dict(scales=None, rotations=None, cov3D_precomp=None)
    if pipe.compute_cov3D_python:
        geometry["cov3D_precomp"] = pc.get_covariance(scaling_modifier)
    else:
        geometry["scales"], geometry["rotations"] = pc.get_scaling, pc.get_rotation

    appearance = dict(shs=None, colors_precomp=None)
    if override_color is not None:
        appearance["colors_precomp"] = override_color
    elif pipe.convert_SHs_python:
        feats = pc.get_features
        shs_view = feats.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
        dirs = xyz - viewpoint_camera.camera_center.repeat(feats.shape[0], 1)
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        appearance["colors_precomp"] = torch.clamp_min(eval_sh_torch(pc.active_sh_degree, shs_view, dirs) + 0.5, 0.0)
    else:
        appearance["shs"] = pc.get_features

    outputs = rasterizer(means3D=xyz, means2D=screenspace_points, opacities=pc.get_opacity, **appearance, **geometry)
    return _package(outputs, screenspace_points, f_count)


def _package(outputs, screenspace_points, f_count):
    if f_count:
        gaussians_count, important_score, rendered_image, radii = outputs
    else:
        rendered_image, radii = outputs
    result = {
        "render": rendered_image,
        "viewspace_points": screenspace_points,
        "visibility_filter": radii > 0,
        "radii": radii,
    }
    if f_count:
        result["gaussians_count"] = gaussians_count
        result["important_score"] = important_score
    return result


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """Render the scene.  Background tensor (bg_color) must be on the GPU."""
    return _render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, False)


def count_render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """render() plus per-Gaussian `gaussians_count` and `important_score` for this view (prune.py:133-157)."""
    return _render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, True)
