"""Run with a stack's PYTHONPATH (tests/scripts_harness.py): build the reference's own GaussianModel from raw leaves and save the
`(capture(), iteration)` tuple its scripts load with --start_checkpoint / --teacher_model (scene/gaussian_model.py:62-76)."""
import sys
from argparse import ArgumentParser

import numpy as np
import torch
from torch import nn

from arguments import OptimizationParams
from scene.gaussian_model import GaussianModel

raw_path, out_path, iteration = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = np.load(raw_path)
parser = ArgumentParser()
op = OptimizationParams(parser)
opt = op.extract(parser.parse_args([]))
g = GaussianModel(3)
leaf = lambda a: nn.Parameter(torch.from_numpy(np.ascontiguousarray(a)).float().cuda().requires_grad_(True))  # noqa: E731
g._xyz, g._features_dc, g._features_rest = leaf(raw["xyz"]), leaf(raw["features_dc"]), leaf(raw["features_rest"])
g._scaling, g._rotation, g._opacity = leaf(raw["scaling"]), leaf(raw["rotation"]), leaf(raw["opacity"])
g.active_sh_degree = 3
g.max_radii2D = torch.zeros((g._xyz.shape[0]), device="cuda")
g.spatial_lr_scale = float(raw["spatial_lr_scale"])
g.training_setup(opt)
torch.save((g.capture(), iteration), out_path)
print("checkpoint", out_path, tuple(g._xyz.shape), type(g.optimizer).__name__)
