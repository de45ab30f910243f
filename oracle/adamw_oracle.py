"""CPU oracle of the optimizer step (TEST INFRASTRUCTURE ONLY): numpy restatement of torch.optim.AdamW as the reference
configures it -- `torch.optim.AdamW(l, lr=0.0, eps=1e-15)`, scene/gaussian_model.py:217; betas (0.9, 0.999), weight_decay 0.01,
amsgrad off -- following torch/optim/adam.py (_single_tensor_adam / _multi_tensor_adam, decoupled weight decay) op for op in
float32, host scalars in double.  The fused multiply-adds of the CUDA functors are emulated through float64 (exact product,
one rounding to double, one to float: differs from a true FMA only on rare double-rounding ties).
Pinned by tests/golden/pytorch_adamw.npz: ten steps of torch.optim.AdamW itself on CPU (tests/golden/make_adamw_golden.py)."""
import numpy as np

f32, f64 = np.float32, np.float64


def fma32(a, b, c):
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15, weight_decay=0.01):
    """one update, `step` = count after increment; returns new (p, m, v) as float32 arrays"""
    p, g, m, v = (np.asarray(a, f32) for a in (p, g, m, v))
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    neg_step = f32((lr / bc1) * -1)
    bc2_sqrt = f32(bc2 ** 0.5)
    p1 = p * f32(1 - lr * weight_decay)
    m1 = fma32(f32(1 - beta1), g - m, m)
    v1 = fma32(f32(1 - beta2), g * g, v * f32(beta2))
    d = np.sqrt(v1) / bc2_sqrt + f32(eps)
    p2 = fma32(neg_step, m1 / d, p1)
    return p2, m1, v1


def compact_rows(tensors, keep):
    """GaussianModel._prune_optimizer's `t[mask]` (scene/gaussian_model.py:564-585)"""
    keep = np.asarray(keep, bool)
    return [np.asarray(t)[keep] for t in tensors]
