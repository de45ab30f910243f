"""nearest-code search at the reference's size (80 000 x 8192 x 27): tensor-core path vs the FP32 kernel, device time per call"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightgaussian_b200 import capi, vectree as vt

g = torch.Generator(device="cuda").manual_seed(5)
x = torch.randn(80000, 27, device="cuda", generator=g) * 0.5
e = torch.randn(8192, 27, device="cuda", generator=g) * 0.7
for mode in (0, 1):
    capi.set_vq_mode(mode)
    for _ in range(3):
        vt.vq_assign(x, e)
    torch.cuda.synchronize()
    capi.profile_collect()
    capi.profile_enable(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        vt.vq_assign(x, e)
    b.record()
    torch.cuda.synchronize()
    prof = capi.profile_collect()
    capi.profile_enable(False)
    ms, n = prof["vq_assign_kernel"]
    print(f"mode {mode}: {a.elapsed_time(b) / 20:.3f} ms per vq_assign call, assign stage {ms / max(n, 1):.3f} ms")
capi.set_vq_mode(0)
