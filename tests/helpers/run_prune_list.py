"""Run with a stack's PYTHONPATH: the reference's own `prune.prune_list` (prune.py:133-157) on its own Scene + restored GaussianModel,
exactly as prune_finetune.py:71-76,225 sets them up; saves the summed counts / scores."""
import sys
from argparse import ArgumentParser

import numpy as np
import torch

from arguments import ModelParams, OptimizationParams, PipelineParams
from prune import prune_list, calculate_v_imp_score
from scene import Scene, GaussianModel
from utils.general_utils import safe_state

parser = ArgumentParser()
lp, op, pp = ModelParams(parser), OptimizationParams(parser), PipelineParams(parser)
parser.add_argument("--start_checkpoint", type=str)
parser.add_argument("--out", type=str)
args = parser.parse_args(sys.argv[1:])
safe_state(True)
dataset, opt, pipe = lp.extract(args), op.extract(args), pp.extract(args)
import os
os.makedirs(dataset.model_path, exist_ok=True)       # prepare_output_and_logger() does this in the scripts (utils/logger_utils.py:33)
gaussians = GaussianModel(dataset.sh_degree)
scene = Scene(dataset, gaussians)
gaussians.training_setup(opt)
model_params, _ = torch.load(args.start_checkpoint)
gaussians.restore(model_params, opt)
background = torch.tensor([0, 0, 0], dtype=torch.float32, device="cuda")
with torch.no_grad():
    gaussian_list, imp_list = prune_list(gaussians, scene, pipe, background)
    v_list = calculate_v_imp_score(gaussians, imp_list, 0.1)
np.savez(args.out, count=gaussian_list.cpu().numpy(), score=imp_list.cpu().numpy(), v=v_list.cpu().numpy())
