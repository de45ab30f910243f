"""GPU experiment: which rounding does each torch foreach functor on the AdamW path use?  Compares torch._foreach_* outputs with
float64-emulated candidates (fused vs unfused) and prints the fraction of bitwise matches.  Diagnostic only."""
import torch

torch.manual_seed(0)
n = 1 << 20
dev = "cuda"
m, g, v, p = (torch.randn(n, device=dev) for _ in range(4))
v = v.abs()
f64 = torch.float64


def frac(a, b):
    return float((a == b).float().mean())


def fma(a, b, c):
    return (a.to(f64) * b.to(f64) + c.to(f64)).float()


w1 = 1 - 0.9
out = torch._foreach_lerp([m], [g], w1)[0]
w = torch.tensor(w1, dtype=torch.float32, device=dev)
print("lerp   fma(w, g-m, m):", frac(out, fma(w, g - m, m)), " unfused:", frac(out, m + w * (g - m)))
val = 1 - 0.999
out = torch._foreach_addcmul([v], [g], [g], val)[0]
a = torch.tensor(val, dtype=torch.float32, device=dev)
print("addcmul fma(a, g*g, v):", frac(out, fma(a, g * g, v)), " fma(a*g, g, v):", frac(out, fma(a * g, g, v)), " unfused a*(g*g):", frac(out, v + a * (g * g)),
      " unfused (a*g)*g:", frac(out, v + (a * g) * g))
s = 0.0316227766 ** 1.0
out = torch._foreach_div([v], [s])[0]
sf = torch.tensor(s, dtype=torch.float32, device=dev)
print("div    true:", frac(out, (v.to(f64) / sf.to(f64)).float()), " mul-by-reciprocal:", frac(out, v * (1.0 / sf)))
st = -1.6e-3
out = torch._foreach_addcdiv([p], [m], [v + 1e-3], [st])[0]
sc = torch.tensor(st, dtype=torch.float32, device=dev)
q = (m.to(f64) / (v + 1e-3).to(f64)).float()
print("addcdiv fma(s, m/d, p):", frac(out, fma(sc, q, p)), " unfused:", frac(out, p + sc * q), " (s*m)/d fused:", frac(out, fma(sc * m, 1.0 / (v + 1e-3), p)))
out = torch._foreach_sqrt([v])[0]
print("sqrt   ieee:", frac(out, v.to(f64).sqrt().float()))
