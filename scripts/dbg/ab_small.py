import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from crfconv_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda")
out = {}
for (M, C1, C2, Co, pro) in [(320, 512, 0, 64, False), (320, 64, 0, 64, True), (1280, 256, 0, 64, False), (1280, 64, 0, 64, True), (1280, 64, 0, 256, False), (1280, 256, 256, 256, True)]:
    g = torch.Generator(device="cuda").manual_seed(M + C1 + Co)
    X1 = torch.randn(M, C1, generator=g, device=dev) + 0.5
    X2 = torch.randn(M, C2, generator=g, device=dev) if C2 else None
    W = torch.randn(Co, C1 + C2, generator=g, device=dev) / (C1 + C2) ** 0.5
    sc, sh = (1 + 0.2 * torch.randn(C1, generator=g, device=dev), 0.2 * torch.randn(C1, generator=g, device=dev)) if pro else (None, None)
    bn = ops.BN(Co, dev)
    H = ops.linear_fwd(X1, W, scale1=sc, shift1=sh, slope1=0.1, X2=X2, stats=bn.stats)
    A = X1.double()
    if pro:
        A = torch.nn.functional.leaky_relu(A * sc.double() + sh.double(), 0.1)
    if C2:
        A = torch.cat([A, X2.double()], 1)
    Hr = A @ W.double().t()
    st = bn.stats.view(-1, 2 * Co).double().sum(0)
    e_h = float((H.double() - Hr).abs().max() / Hr.abs().max())
    e_s = float(((st[:Co] - Hr.sum(0)).abs() / Hr.abs().sum(0)).max())
    e_q = float(((st[Co:] - (Hr ** 2).sum(0)).abs() / (Hr ** 2).sum(0)).max())
    print(f"M={M} {C1}+{C2}->{Co} pro={pro}: H {e_h:.2e}  sum {e_s:.2e}  sq {e_q:.2e}  slots used {int((bn.stats.view(-1, 2*Co).abs().sum(1) > 0).sum())}")
