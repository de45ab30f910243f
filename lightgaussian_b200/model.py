"""Minimal stand-ins for the host-side objects render() consumes, for tests and bench.py on a box without
the reference checkout: a parameter container exposing GaussianModel's getters
(scene/gaussian_model.py:98-118) and PipelineParams' three flags (arguments/__init__.py:62-68).
The reference's own GaussianModel / Camera objects work with render() unchanged."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F


def pipeline_params(convert_SHs_python=False, compute_cov3D_python=False, debug=False):
    return SimpleNamespace(convert_SHs_python=convert_SHs_python, compute_cov3D_python=compute_cov3D_python, debug=debug)


class GaussianParams:
    """Raw (pre-activation) leaves with the reference's names and activations."""

    def __init__(self, raw: dict, sh_degree: int, device="cuda", requires_grad=True):
        def leaf(a):
            t = torch.as_tensor(a, dtype=torch.float32).to(device).contiguous()
            return t.requires_grad_(requires_grad)
        self._xyz = leaf(raw["xyz"])
        self._features_dc = leaf(raw["features_dc"])
        self._features_rest = leaf(raw["features_rest"])
        self._scaling = leaf(raw["scaling"])
        self._rotation = leaf(raw["rotation"])
        self._opacity = leaf(raw["opacity"])
        self.max_sh_degree = sh_degree
        self.active_sh_degree = sh_degree

    def parameters(self):
        return [self._xyz, self._features_dc, self._features_rest, self._scaling, self._rotation, self._opacity]

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return F.normalize(self._rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    def get_covariance(self, scaling_modifier=1.0):
        """Sigma = R S S^T R^T, upper triangle (scene/gaussian_model.py:29-33, utils/general_utils.py:68-119)."""
        s = scaling_modifier * self.get_scaling
        q = self.get_rotation
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        L = R * s[:, None, :]
        S = L @ L.transpose(1, 2)
        return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1)


class TorchCamera:
    """A SynthCamera moved to a device: the attributes render() reads (gaussian_renderer/__init__.py:49-62)."""

    def __init__(self, cam, device="cuda"):
        self.image_width, self.image_height = cam.image_width, cam.image_height
        self.FoVx, self.FoVy = cam.FoVx, cam.FoVy
        self.world_view_transform = torch.from_numpy(cam.world_view_transform).to(device)
        self.full_proj_transform = torch.from_numpy(cam.full_proj_transform).to(device)
        self.camera_center = torch.from_numpy(cam.camera_center).to(device)
