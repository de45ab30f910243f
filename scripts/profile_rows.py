"""Workload for the ncu capture of the N2/N3 kernels (run under `ncu -k regex:image_loss|adamw_multi|compact_gather`):
fused loss forward+backward on one 3x1080x1920 image, one FusedAdamW step over the six groups of a 3M-Gaussian model, one prune
compaction (66 %) of parameters + both moments."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lightgaussian_b200 import loss as fused_loss
from lightgaussian_b200.optim import FusedAdamW, compact_rows
from lightgaussian_b200 import vectree

dev = "cuda"
H, W, P = 1080, 1920, 3_000_000
x = torch.rand(3, H, W, device=dev, requires_grad=True)
y = torch.rand(3, H, W, device=dev)
shapes = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
params = {k: torch.nn.Parameter(torch.randn((P,) + s, device=dev)) for k, s in shapes.items()}
opt = FusedAdamW([{"params": [p], "lr": 1e-3, "name": k} for k, p in params.items()], lr=0.0, eps=1e-15)
for it in range(2):
    x.grad = None
    fused_loss.l1_ssim_loss(x, y, 0.2).backward()
    for p in params.values():
        p.grad = torch.randn_like(p) * 1e-3
        p.grad[::3] = 0.0                                  # culled Gaussians: exact zeros, as rendered gradients have
    opt.step()
xv = torch.randn(80000, 27, device=dev) * 0.5
wv = torch.rand(80000, device=dev) ** 2
vq = vectree.VectorQuantize(dim=27, codebook_size=8192).to(dev).train()
for it in range(2):
    vq(xv.unsqueeze(0), weight=wv.reshape(1, -1, 1))
for it in range(2):
    x.grad = None
    fused_loss.l1_loss(x, y).backward()
keep = torch.rand(P, device=dev) > 0.66
tensors = []
for p in params.values():
    tensors += [p.detach(), opt.state[p]["exp_avg"], opt.state[p]["exp_avg_sq"]]
for it in range(2):
    outs = compact_rows(tensors, keep)
torch.cuda.synchronize()
print("rows kept", outs[0].shape[0])
