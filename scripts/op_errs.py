import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crfconv_b200 import ops, _lib
L = _lib.lib()
def lrelu(v, s): return torch.where(v > 0, v, v * s)
def errs(a, b):
    d = (a.double() - b)
    return f"max {float(d.abs().max() / b.abs().max()):.1e} l2 {float(d.norm() / b.norm()):.1e}"
for M in (9000, 100000):
  for (C1, C2, Cout) in ((64, 64, 64), (16, 0, 64), (128, 0, 16), (16, 0, 16), (64, 0, 16)):
    for fast in (0, 1):
        g = torch.Generator(device="cuda").manual_seed(1)
        rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
        X1, X2 = rn(M, C1), (rn(M, C2) if C2 else None)
        W = rn(Cout, C1 + C2) / (C1 + C2) ** 0.5
        sc1, sh1 = 1 + 0.2 * rn(C1), 0.2 * rn(C1)
        gamma, beta = 1 + 0.2 * rn(Cout), 0.2 * rn(Cout)
        dY = rn(M, Cout)
        L.crfconv_set_fast_path(fast)
        bn = ops.BN(Cout, X1.device)
        H = ops.linear_fwd(X1, W, scale1=sc1, shift1=sh1, slope1=0.1, X2=X2, stats=bn.stats)
        A = lrelu(X1.double() * sc1.double() + sh1.double(), 0.1)
        if C2: A = torch.cat([A, X2.double()], 1)
        Href = A @ W.double().t()
        e_f = errs(H, Href)
        ops.bn_finalize_fwd(bn, M, gamma, beta, 1e-5, 0.1, True, None, None)
        Href = H.double(); mu, var = Href.mean(0), Href.var(0, unbiased=False)
        Hh = (Href - mu) * (var + 1e-5).rsqrt(); V = Hh * gamma.double() + beta.double()
        dV = torch.where(V > 0, dY.double(), dY.double() * 0.1)
        dgam, dbet = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
        ops.bn_backward_prepare(dY, H, bn, 0.1, dgam, dbet)
        dH = gamma.double() * (var + 1e-5).rsqrt() * (dV - dV.mean(0) - Hh * (dV * Hh).mean(0))
        dX1, dX2 = torch.empty_like(X1), (torch.zeros_like(X2) if C2 else None)
        dW = torch.zeros_like(W)
        ops.linear_bwd(dY, H, bn, 0.1, X1, W, scale1=sc1, shift1=sh1, slope1=0.1, X2=X2, dX1=dX1, dX2=dX2, dW=dW)
        dA = dH @ W.double()
        print(f"M={M} {C1}+{C2}->{Cout} fast={fast}: fwd {e_f} | dX1 {errs(dX1, dA[:, :C1])} | dW {errs(dW, dH.t() @ A)}")
