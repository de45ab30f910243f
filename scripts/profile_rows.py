"""Workload for the ncu capture of the N2/N3 kernels (run under `ncu -k regex:image_loss|adamw_multi|compact_gather`):
fused loss forward+backward on one 3x1080x1920 image, one FusedAdamW step over the six groups of a 3M-Gaussian model, one prune
compaction (66 %) of parameters + both moments."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lightgaussian_b200 import loss as fused_loss
from lightgaussian_b200.optim import FusedAdamW, compact_rows

dev = "cuda"
H, W, P = 1080, 1920, 3_000_000
x = torch.rand(3, H, W, device=dev, requires_grad=True)
y = torch.rand(3, H, W, device=dev)
shapes = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
params = {k: torch.nn.Parameter(torch.randn((P,) + s, device=dev)) for k, s in shapes.items()}
opt = FusedAdamW([{"params": [p], "lr": 1e-3, "name": k} for k, p in params.items()], lr=0.0, eps=1e-15)
for it in range(3):
    x.grad = None
    fused_loss.l1_ssim_loss(x, y, 0.2).backward()
    for p in params.values():
        p.grad = torch.randn_like(p) * 1e-3
    opt.step()
keep = torch.rand(P, device=dev) > 0.66
tensors = []
for p in params.values():
    tensors += [p.detach(), opt.state[p]["exp_avg"], opt.state[p]["exp_avg_sq"]]
for it in range(2):
    outs = compact_rows(tensors, keep)
torch.cuda.synchronize()
print("rows kept", outs[0].shape[0])
