"""GPU micro-benchmark of the optimizer row: one AdamW step over the six groups of a 3M-Gaussian model (177 M parameters), fused
(lgr_adamw_step) against torch.optim.AdamW, for well-scaled gradients and for faint ones (zeros / denormal second moments)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightgaussian_b200.optim import FusedAdamW  # noqa: E402

dev, P = "cuda", 3_000_000
shapes = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}


def run(cls, scale, zero_frac):
    torch.manual_seed(0)
    params = {k: torch.nn.Parameter(torch.randn((P,) + s, device=dev)) for k, s in shapes.items()}
    opt = cls([{"params": [p], "lr": 1e-3, "name": k} for k, p in params.items()], lr=0.0, eps=1e-15)
    grads = {k: torch.randn_like(p) * scale for k, p in params.items()}
    if zero_frac:
        dead = torch.rand(P, device=dev) < zero_frac
        for g in grads.values():
            g[dead] = 0.0
    for k, p in params.items():
        p.grad = grads[k]
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


n_el = sum(P * int(torch.tensor(s).prod()) for s in shapes.values())
for name, scale, zf in [("well-scaled", 1e-3, 0.0), ("30% exact zeros", 1e-3, 0.3), ("faint (1e-20): denormal moments", 1e-20, 0.3)]:
    tf, tt = run(FusedAdamW, scale, zf), run(torch.optim.AdamW, scale, zf)
    print(f"{name:34s} fused {tf:.3f} ms ({28 * n_el / tf / 1e6:.0f} GB/s)   torch {tt:.3f} ms   x{tt / tf:.2f}")
