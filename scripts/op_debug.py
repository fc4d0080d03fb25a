import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crfconv_b200 import ops, _lib
L = _lib.lib()
M, C1, Cout = int(sys.argv[1]), 16, 64
g = torch.Generator(device="cuda").manual_seed(1)
rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
X1 = rn(M, C1); W = rn(Cout, C1) / 4; gamma, beta = 1 + 0.2 * rn(Cout), 0.2 * rn(Cout); dY = rn(M, Cout)
res = {}
for fast in (0, 1):
    L.crfconv_set_fast_path(fast)
    bn = ops.BN(Cout, X1.device)
    H = ops.linear_fwd(X1, W, stats=bn.stats)
    ops.bn_finalize_fwd(bn, M, gamma, beta, 1e-5, 0.1, True, None, None)
    dgam, dbet = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    ops.bn_backward_prepare(dY, H, bn, 0.1, dgam, dbet)
    dX1 = torch.empty_like(X1); dW = torch.zeros_like(W)
    ops.linear_bwd(dY, H, bn, 0.1, X1, W, dX1=dX1, dW=dW)
    # plain variants
    dXp = torch.empty_like(X1); dWp = torch.zeros_like(W)
    ops.linear_bwd(dY, None, None, 1.0, X1, W, dX1=dXp, dW=dWp)
    torch.cuda.synchronize()
    res[fast] = (H.clone(), dX1, dW, dXp, dWp, bn.k1.clone(), bn.k2.clone())
for i, name in enumerate(["H", "dX1", "dW", "dX_plain", "dW_plain", "k1", "k2"]):
    a, b = res[0][i], res[1][i]
    d = (a - b).abs()
    thr = 1e-3 * a.abs().max()
    bad = (d > thr).nonzero()
    print(name, "max diff", float(d.max() / a.abs().max()), "n_bad", bad.shape[0])
    if bad.shape[0] and bad.shape[1] == 2:
        rows = bad[:, 0]
        print("   rows%128:", sorted(set((rows % 128).tolist()))[:40], " tiles:", sorted(set((rows // 128).tolist()))[:20], " cols:", sorted(set(bad[:, 1].tolist()))[:20])
