"""GPU experiment: find the explicit operation order that reproduces torch's activations bit for bit."""
import ctypes as C, os, torch
import torch.nn.functional as F
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libexp_act.so"))
for f in (lib.run_exp, lib.run_sigmoid, lib.run_normalize):
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
g = torch.Generator().manual_seed(0)
n = 3_000_000
x = (torch.randn(n, generator=g) * 2 - 5).cuda()
ref = torch.exp(x)
for v in range(2):
    y = torch.empty_like(x); lib.run_exp(x.data_ptr(), y.data_ptr(), n, v); torch.cuda.synchronize()
    print("exp variant", v, "mismatches", int((y != ref).sum()))
x = (torch.randn(n, generator=g) * 2).cuda()
ref = torch.sigmoid(x)
for v in range(3):
    y = torch.empty_like(x); lib.run_sigmoid(x.data_ptr(), y.data_ptr(), n, v); torch.cuda.synchronize()
    print("sigmoid variant", v, "mismatches", int((y != ref).sum()))
q = torch.randn(n, 4, generator=g).cuda()
ref = F.normalize(q)
for v in range(9):
    y = torch.empty_like(q); lib.run_normalize(q.data_ptr(), y.data_ptr(), n, v); torch.cuda.synchronize()
    print("normalize variant", v, "mismatching rows", int((y != ref).any(dim=1).sum()))
# norm alone
nrm = q.norm(2, dim=1)
