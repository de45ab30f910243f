"""CPU: the AdamW oracle (oracle/adamw_oracle.py) against torch.optim.AdamW's own CPU results (golden), and the host logic of
FusedAdamW that needs no GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import adamw_oracle as ao

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pytorch_adamw.npz")
NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]


def replay(g, steps=10):
    lrs = dict(zip(NAMES, g["lrs"]))
    res = {}
    for k in NAMES:
        p = g[f"p0_{k}"].copy()
        m, v = np.zeros_like(p), np.zeros_like(p)
        for it in range(steps):
            lr = 1.0e-4 if (k == "xyz" and it >= 5) else float(lrs[k])
            p, m, v = ao.adamw_step(p, g[f"g{it}_{k}"], m, v, it + 1, lr)
        res[k] = (p, m, v)
    return res


def test_oracle_reproduces_torch_adamw():
    g = np.load(GOLD)
    res = replay(g)
    for k in NAMES:
        p, m, v = res[k]
        # torch's CPU kernels may or may not contract to FMA: allow a few ulp after 10 steps, nothing more
        for ours, name in ((p, "p"), (m, "m"), (v, "v")):
            ref = g[f"{name}_{k}"]
            assert np.abs(ours - ref).max() <= 4e-7 * np.abs(ref).max(), (k, name)


def test_fused_adamw_keeps_the_reference_facing_surface():
    from lightgaussian_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(4, 3))
    opt = FusedAdamW([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    grp = opt.param_groups[0]
    assert grp["name"] == "xyz" and grp["lr"] == 0.1 and grp["eps"] == 1e-15 and grp["weight_decay"] == 0.01 and grp["betas"] == (0.9, 0.999)
    assert isinstance(opt, torch.optim.AdamW)
    sd = opt.state_dict()
    assert sd["param_groups"][0]["name"] == "xyz"
    opt.step()                                   # no gradients: nothing to do, no library needed
    p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError):            # CPU parameters are refused: there is no CPU path
        opt.step()
    with pytest.raises(NotImplementedError):
        FusedAdamW([p], amsgrad=True)


def test_install_swaps_optimizer_and_prune_on_a_gaussian_model_class():
    from lightgaussian_b200 import optim

    class FakeModel:                                     # the two methods of scene/gaussian_model.py that install() touches
        def training_setup(self, training_args):
            self._xyz = torch.nn.Parameter(torch.zeros(5, 3))
            self._opacity = torch.nn.Parameter(torch.zeros(5, 1))
            self.optimizer = torch.optim.AdamW([{"params": [self._xyz], "lr": 1e-4 * training_args, "name": "xyz"},
                                                {"params": [self._opacity], "lr": 0.05, "name": "opacity"}], lr=0.0, eps=1e-15)
            return "ret"

        def prune_points(self, mask):
            raise AssertionError("replaced")

    optim.install(FakeModel)
    optim.install(FakeModel)                             # idempotent
    m = FakeModel()
    assert m.training_setup(2.0) == "ret"
    assert isinstance(m.optimizer, optim.FusedAdamW)
    assert [g["name"] for g in m.optimizer.param_groups] == ["xyz", "opacity"]
    assert m.optimizer.param_groups[0]["lr"] == 2e-4 and m.optimizer.param_groups[0]["eps"] == 1e-15
    assert m.optimizer.param_groups[0]["params"][0] is m._xyz
    assert FakeModel.prune_points is optim.prune_points
    with pytest.raises(TypeError):
        optim.to_fused(torch.optim.SGD([m._xyz], lr=0.1))
