"""Generates tests/golden/pytorch_adamw.npz: ten steps of torch.optim.AdamW ITSELF (CPU), configured as the reference does
(scene/gaussian_model.py:184-217: six groups with their own lr, lr=0.0 default, eps=1e-15), on seeded inputs.
Run:  python tests/golden/make_adamw_golden.py"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(5)
shapes = {"xyz": (257, 3), "f_dc": (257, 1, 3), "f_rest": (257, 15, 3), "opacity": (257, 1), "scaling": (257, 3), "rotation": (257, 4)}
lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
params = {k: torch.nn.Parameter(torch.from_numpy(rng.standard_normal(s).astype(np.float32))) for k, s in shapes.items()}
opt = torch.optim.AdamW([{"params": [params[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
out = {f"p0_{k}": v.detach().numpy().copy() for k, v in params.items()}
out["lrs"] = np.array([lrs[k] for k in shapes])
STEPS = 10
for it in range(STEPS):
    for k, p in params.items():
        g = rng.standard_normal(shapes[k]).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 1))
        g[rng.random(shapes[k]) < 0.3] = 0.0                      # Gaussians outside the view get exact zeros
        out[f"g{it}_{k}"] = g
        p.grad = torch.from_numpy(g.copy())
    if it == 5:
        opt.param_groups[0]["lr"] = 1.0e-4                        # update_learning_rate() changes the xyz group in flight
    opt.step()
for k, p in params.items():
    out[f"p_{k}"] = p.detach().numpy().copy()
    out[f"m_{k}"] = opt.state[p]["exp_avg"].numpy().copy()
    out[f"v_{k}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
np.savez_compressed(os.path.join(HERE, "pytorch_adamw.npz"), **out)
print("ok", STEPS)
