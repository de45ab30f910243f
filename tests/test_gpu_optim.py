"""GPU parity of the optimizer row (N3): FusedAdamW (lgr_adamw_step) against torch.optim.AdamW on the same device -- BIT-EXACT,
parameters and both moments, over many steps with the reference's six-group configuration -- against the numpy oracle and
the torch-CPU golden; fused prune compaction (lgr_compact_plan / lgr_compact_rows) against boolean indexing, and
`optim.prune_points` against the reference's GaussianModel.prune_points surgery restated with torch ops."""
import copy
import os

import numpy as np
import pytest
import torch

from lightgaussian_b200.optim import FusedAdamW, compact_rows, prune_points
from oracle import adamw_oracle as ao

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pytorch_adamw.npz")
NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
LRS = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}


def _groups(P, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn((P,) + SHAPES[k], generator=g) for k in NAMES}


def _make(cls, init):
    params = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in init.items()}
    opt = cls([{"params": [params[k]], "lr": LRS[k], "name": k} for k in NAMES], lr=0.0, eps=1e-15)
    return params, opt


def _grads(P, it, seed):
    g = torch.Generator().manual_seed(1000 * seed + it)
    out = {}
    for k in NAMES:
        x = torch.randn((P,) + SHAPES[k], generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
        x[torch.rand(P, generator=g) < 0.3] = 0.0
        out[k] = x.cuda()
    return out


@pytest.mark.parametrize("P", [1, 1000, 4096 * 3 // 3 + 5, 70001])
def test_bit_exact_against_torch_adamw(P):
    init = _groups(P, 1)
    ours_p, ours = _make(FusedAdamW, init)
    ref_p, ref = _make(torch.optim.AdamW, init)
    for it in range(25):
        gr = _grads(P, it, 7)
        for k in NAMES:
            ours_p[k].grad = gr[k].clone()
            ref_p[k].grad = gr[k].clone()
        if it == 10:
            ours.param_groups[0]["lr"] = ref.param_groups[0]["lr"] = 1.0e-4       # update_learning_rate()
        ours.step()
        ref.step()
    for k in NAMES:
        assert torch.equal(ours_p[k], ref_p[k]), k
        assert torch.equal(ours.state[ours_p[k]]["exp_avg"], ref.state[ref_p[k]]["exp_avg"]), k
        assert torch.equal(ours.state[ours_p[k]]["exp_avg_sq"], ref.state[ref_p[k]]["exp_avg_sq"]), k
        assert float(ours.state[ours_p[k]]["step"]) == float(ref.state[ref_p[k]]["step"]) == 25.0


@pytest.mark.parametrize("scale", [1e-18, 1e-21, 1e30])
def test_bit_exact_on_zeros_denormals_and_extremes(scale):
    """faint gradients: g*g underflows to denormals / zero, exp_avg_sq becomes denormal, the quotient m/d spans the whole exponent
    range -- the operands that take the fp64 route inside the kernel (and the IEEE slow path inside torch's)."""
    P = 30011
    init = _groups(P, 5)
    ours_p, ours = _make(FusedAdamW, init)
    ref_p, ref = _make(torch.optim.AdamW, init)
    for it in range(6):
        gr = _grads(P, it, 21)
        for k in NAMES:
            g = gr[k] * scale if it != 3 else torch.zeros_like(gr[k])        # one step with exact zeros everywhere
            ours_p[k].grad = g.clone()
            ref_p[k].grad = g.clone()
        ours.step()
        ref.step()
    for k in NAMES:
        assert torch.equal(ours_p[k], ref_p[k]), k
        assert torch.equal(ours.state[ours_p[k]]["exp_avg"], ref.state[ref_p[k]]["exp_avg"]), k
        assert torch.equal(ours.state[ours_p[k]]["exp_avg_sq"], ref.state[ref_p[k]]["exp_avg_sq"]), k


def test_bit_exact_when_the_quotient_is_subnormal():
    """first moments that have decayed for hundreds of steps against second moments that have not: m/d lands in the subnormal
    range (the lanes that take the exact binary64 route), with parameters small enough for those bits to show."""
    P = 20000
    g = torch.Generator().manual_seed(3)
    init = {k: torch.randn((P,) + SHAPES[k], generator=g) * 1e-40 for k in NAMES}
    ours_p, ours = _make(FusedAdamW, init)
    ref_p, ref = _make(torch.optim.AdamW, init)
    zero = {k: torch.zeros((P,) + SHAPES[k]).cuda() for k in NAMES}
    for opt, prm in ((ours, ours_p), (ref, ref_p)):
        for k in NAMES:
            prm[k].grad = zero[k].clone()
        opt.step()                                            # creates the state
    for k in NAMES:
        expo = torch.rand((P,) + SHAPES[k], generator=g) * 20 - 46                      # |m| from 1e-46 (flushes to 0) to 1e-26
        m = (torch.randn((P,) + SHAPES[k], generator=g).sign() * 10.0 ** expo).float().cuda()
        v = (10.0 ** (torch.rand((P,) + SHAPES[k], generator=g) * 8 - 6)).float().cuda()   # 1e-6 .. 1e2
        for opt, prm in ((ours, ours_p), (ref, ref_p)):
            opt.state[prm[k]]["exp_avg"].copy_(m)
            opt.state[prm[k]]["exp_avg_sq"].copy_(v)
    for it in range(3):
        for opt, prm in ((ours, ours_p), (ref, ref_p)):
            for k in NAMES:
                prm[k].grad = zero[k].clone()
            opt.step()
    for k in NAMES:
        assert torch.equal(ours_p[k], ref_p[k]), k
        assert torch.equal(ours.state[ours_p[k]]["exp_avg"], ref.state[ref_p[k]]["exp_avg"]), k
        assert float(ours_p[k].detach().abs().max()) > 0


def test_against_oracle_and_torch_cpu_golden():
    g = np.load(GOLD)
    params = {k: torch.nn.Parameter(torch.from_numpy(g[f"p0_{k}"]).cuda()) for k in NAMES}
    lrs = dict(zip(NAMES, g["lrs"]))
    opt = FusedAdamW([{"params": [params[k]], "lr": float(lrs[k]), "name": k} for k in NAMES], lr=0.0, eps=1e-15)
    orc = {k: (g[f"p0_{k}"].copy(), np.zeros_like(g[f"p0_{k}"]), np.zeros_like(g[f"p0_{k}"])) for k in NAMES}
    for it in range(10):
        for k in NAMES:
            params[k].grad = torch.from_numpy(g[f"g{it}_{k}"]).cuda()
            lr = 1.0e-4 if (k == "xyz" and it >= 5) else float(lrs[k])
            orc[k] = ao.adamw_step(*orc[k][:1], g[f"g{it}_{k}"], *orc[k][1:], it + 1, lr)
        if it == 5:
            opt.param_groups[0]["lr"] = 1.0e-4
        opt.step()
    for k in NAMES:
        got = (params[k].detach().cpu().numpy(), opt.state[params[k]]["exp_avg"].cpu().numpy(), opt.state[params[k]]["exp_avg_sq"].cpu().numpy())
        for a, o, name in zip(got, orc[k], "pmv"):
            assert np.abs(a - o).max() <= 2.5e-7 * np.abs(o).max(), (k, name)          # oracle FMA emulation: double-rounding ties only
            ref = g[f"{name}_{k}"]
            assert np.abs(a - ref).max() <= 4e-7 * np.abs(ref).max(), (k, name)        # torch CPU (no FMA contraction there)


def test_state_dict_round_trip_and_unaligned_views():
    init = _groups(333, 3)
    p1, o1 = _make(FusedAdamW, init)
    for it in range(3):
        for k, gr in _grads(333, it, 9).items():
            p1[k].grad = gr
        o1.step()
    p2, o2 = _make(FusedAdamW, {k: v.detach().cpu() for k, v in p1.items()})
    # GaussianModel.restore (scene/gaussian_model.py:86-96); deep copies: load_state_dict aliases tensors that need no cast
    o2.load_state_dict(copy.deepcopy(o1.state_dict()))
    pr, orf = _make(torch.optim.AdamW, {k: v.detach().cpu() for k, v in p1.items()})
    orf.load_state_dict(copy.deepcopy(o1.state_dict()))
    gr = _grads(333, 99, 9)
    for k in NAMES:
        p2[k].grad = gr[k].clone()
        pr[k].grad = gr[k].clone()
    o2.step()
    orf.step()
    for k in NAMES:
        assert torch.equal(p2[k], pr[k]), k


@pytest.mark.parametrize("P,frac", [(1, 1.0), (1, 0.0), (5000, 0.34), (200003, 0.66), (4097, 0.0), (4097, 1.0)])
def test_compact_rows_equals_boolean_indexing(P, frac):
    g = torch.Generator().manual_seed(P)
    keep = (torch.rand(P, generator=g) < frac).cuda()
    tensors = [torch.randn((P,) + s, generator=g).cuda() for s in [(3,), (1, 3), (15, 3), (1,), (4,)]]
    tensors.append(torch.randint(0, 1000, (P,), generator=g, dtype=torch.int32).cuda())
    outs = compact_rows(tensors, keep)
    for t, o in zip(tensors, outs):
        assert o.dtype == t.dtype and torch.equal(o, t[keep])


class _Model:
    """the attributes GaussianModel.prune_points touches (scene/gaussian_model.py:587-600)"""


def _reference_prune(model, mask):
    """GaussianModel._prune_optimizer + prune_points restated with the reference's torch ops (scene/gaussian_model.py:564-600)"""
    valid = ~mask
    attr = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}
    for group in model.optimizer.param_groups:
        st = model.optimizer.state.get(group["params"][0], None)
        new = torch.nn.Parameter(group["params"][0][valid].requires_grad_(True))
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = st["exp_avg"][valid], st["exp_avg_sq"][valid]
            del model.optimizer.state[group["params"][0]]
            model.optimizer.state[new] = st
        group["params"][0] = new
        setattr(model, attr[group["name"]], new)
    model.xyz_gradient_accum = model.xyz_gradient_accum[valid]
    model.denom = model.denom[valid]
    model.max_radii2D = model.max_radii2D[valid]


def _model(P, cls, with_state):
    params, opt = _make(cls, _groups(P, 11))
    m = _Model()
    m.optimizer = opt
    for k, a in zip(NAMES, ["_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"]):
        setattr(m, a, params[k])
    gen = torch.Generator().manual_seed(17)
    m.xyz_gradient_accum = torch.rand(P, 1, generator=gen).cuda()
    m.denom = torch.rand(P, 1, generator=gen).cuda()
    m.max_radii2D = torch.rand(P, generator=gen).cuda()
    if with_state:
        for it in range(2):
            for k, gr in _grads(P, it, 13).items():
                params[k].grad = gr
            opt.step()
    return m


@pytest.mark.parametrize("with_state", [True, False])
def test_prune_points_matches_reference_surgery_and_training_continues(with_state):
    P = 20011
    a, b = _model(P, FusedAdamW, with_state), _model(P, torch.optim.AdamW, with_state)
    mask = (torch.rand(P, generator=torch.Generator().manual_seed(2)) < 0.66).cuda()      # prune_ratio 0.66
    prune_points(a, mask)
    _reference_prune(b, mask)
    attrs = ["_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"]
    for n in attrs + ["xyz_gradient_accum", "denom", "max_radii2D"]:
        assert torch.equal(getattr(a, n), getattr(b, n)), n
    n_keep = int((~mask).sum())
    for ga, gb, n in zip(a.optimizer.param_groups, b.optimizer.param_groups, attrs):
        assert ga["params"][0] is getattr(a, n) and ga["params"][0].requires_grad and ga["params"][0].shape[0] == n_keep
        sa, sb = a.optimizer.state.get(ga["params"][0]), b.optimizer.state.get(gb["params"][0])
        assert (sa is None) == (sb is None) == (not with_state)
        if sa is not None:
            assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])
    for k, n in zip(NAMES, attrs):                       # one more optimizer step on the pruned model
        gr = torch.randn_like(getattr(a, n))
        getattr(a, n).grad = gr.clone()
        getattr(b, n).grad = gr.clone()
    a.optimizer.step()
    b.optimizer.step()
    for n in attrs:
        assert torch.equal(getattr(a, n), getattr(b, n)), n


def test_full_size_3m_step_matches_torch():
    P = 3_000_000
    g = torch.Generator(device="cuda").manual_seed(0)
    p0 = torch.randn(P, 15, 3, device="cuda", generator=g)
    gr = torch.randn(P, 15, 3, device="cuda", generator=g) * 1e-3
    pa, pb = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    oa = FusedAdamW([{"params": [pa], "lr": 1.25e-4, "name": "f_rest"}], lr=0.0, eps=1e-15)
    ob = torch.optim.AdamW([{"params": [pb], "lr": 1.25e-4, "name": "f_rest"}], lr=0.0, eps=1e-15)
    for _ in range(3):
        pa.grad, pb.grad = gr, gr
        oa.step()
        ob.step()
    assert torch.equal(pa, pb)
    assert torch.equal(oa.state[pa]["exp_avg_sq"], ob.state[pb]["exp_avg_sq"])


def test_row_strided_parameter_is_updated_in_place_bit_exactly():
    """distill_train.py:79-80 registers the student's NON-contiguous _features_rest[:, :8, :] (scene/gaussian_model.py:129-136) in the
    optimizer.  FusedAdamW must update that view in place (the caller's tensor stays the key of optimizer.state) with the same bits as
    torch.optim.AdamW on a contiguous copy, leaving the columns outside the view untouched."""
    import torch
    from lightgaussian_b200.optim import FusedAdamW
    P = 5000 + 3
    g = torch.Generator().manual_seed(0)
    full0 = torch.randn(P, 15, 3, generator=g).cuda()
    full = full0.clone()
    view = full[:, :8, :]
    view.requires_grad = True
    assert not view.is_contiguous()
    ref = view.detach().clone().contiguous().requires_grad_(True)
    lr = 2.5e-3 / 20
    opt = FusedAdamW([{"params": [view], "lr": lr, "name": "f_rest"}], lr=0.0, eps=1e-15)
    opt_ref = torch.optim.AdamW([{"params": [ref], "lr": lr, "name": "f_rest"}], lr=0.0, eps=1e-15)
    for step in range(12):
        grad = (torch.randn(P, 8, 3, generator=g) * 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))).cuda()
        grad[::7] = 0.0
        view.grad, ref.grad = grad.clone(), grad.clone()
        opt.step()
        opt_ref.step()
    assert torch.equal(view.detach(), ref.detach())
    st, st_ref = opt.state[view], opt_ref.state[ref]
    assert st["exp_avg"].is_contiguous() and torch.equal(st["exp_avg"], st_ref["exp_avg"]) and torch.equal(st["exp_avg_sq"], st_ref["exp_avg_sq"])
    assert torch.equal(full[:, 8:, :], full0[:, 8:, :])                 # coefficients outside the view are untouched
    assert view.data_ptr() == full.data_ptr()                           # updated in place, no re-materialisation
