"""Device timing of the Linear kernels on the hot-path shapes (development aid, not the bench).
Each shape is run on NSETS rotating input sets so that consecutive launches never find their inputs in the 126 MB L2;
20 back-to-back launches per measurement, CUDA events, GB/s = algorithmic bytes (rows × (Cin + Cout) × 4) / time.
GENERIC=1 times the generic kernels instead (crfconv_set_fast_path(0))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from crfconv_b200 import ops

if os.environ.get("GENERIC") == "1":
    ops._lib.lib().crfconv_set_fast_path(0)

M = int(os.environ.get("M", 6 * 40960))
NSETS = 4
dev = "cuda"
torch.manual_seed(0)


def timeit(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


shapes = [(64, 64, 64, True), (16, 0, 64, False), (64, 0, 16, False), (128, 0, 16, False), (16, 0, 16, True)]   # C1, C2, Cout, prologue
for C1, C2, Co, pro in shapes:
    rows = M if C1 != 128 else M // 4
    X1 = [torch.randn(rows, C1, device=dev) for _ in range(NSETS)]
    X2 = [torch.randn(rows, C2, device=dev) for _ in range(NSETS)] if C2 else [None] * NSETS
    Y = [torch.empty(rows, Co, device=dev) for _ in range(NSETS)]
    W = torch.randn(Co, C1 + C2, device=dev) * 0.1
    sc = torch.rand(C1, device=dev) + 0.5 if pro else None
    sh = torch.randn(C1, device=dev) * 0.1 if pro else None
    stats = torch.zeros(ops.STAT_SLOTS * 2 * Co, device=dev)

    def run(i):
        j = i % NSETS
        ops.linear_fwd(X1[j], W, scale1=sc, shift1=sh, slope1=0.1 if pro else 1.0, X2=X2[j], stats=stats, out=Y[j])

    t = timeit(run)
    nb = rows * (C1 + C2 + Co) * 4
    print(f"linear_fwd [{C1}+{C2} -> {Co}] rows={rows}: {t:7.1f} us   {nb / t / 1e3:7.0f} GB/s   ({nb / 1e6:.0f} MB)")

# ---- weight gradients (plain Linear form: dH = dY; the BN-backward transform adds one more row-sized read of H)
print("--- wgrad only (dW = dY^T X), rows x (Cout + Ktot) x 4 bytes")
for C1, C2, Co in [(64, 64, 64), (16, 0, 64), (64, 0, 16), (128, 0, 16)]:
    rows = M if C1 != 128 else M // 4
    X1 = [torch.randn(rows, C1, device=dev) for _ in range(NSETS)]
    X2 = [torch.randn(rows, C2, device=dev) for _ in range(NSETS)] if C2 else [None] * NSETS
    dY = [torch.randn(rows, Co, device=dev) for _ in range(NSETS)]
    W = torch.randn(Co, C1 + C2, device=dev) * 0.1
    dW = torch.zeros_like(W)
    scratch = torch.zeros(ops.GRAD_SLOTS * W.numel(), device=dev)

    def runw(i):
        j = i % NSETS
        ops.linear_bwd(dY[j], None, None, 1.0, X1[j], W, X2=X2[j], dW=dW, scratch=scratch, scratch_stride=W.numel())

    t = timeit(runw)
    nb = rows * (C1 + C2 + Co) * 4
    print(f"wgrad [{Co} <- {C1}+{C2}] rows={rows}: {t:7.1f} us   {nb / t / 1e3:7.0f} GB/s   ({nb / 1e6:.0f} MB)")

# ---- input gradients (plain Linear form: dH = dY)
print("--- dgrad only (dX = dY W), rows x (Cout + Ktot) x 4 bytes")
for C1, C2, Co in [(64, 64, 64), (16, 0, 64), (64, 0, 16), (128, 0, 16)]:
    rows = M if C1 != 128 else M // 4
    dY = [torch.randn(rows, Co, device=dev) for _ in range(NSETS)]
    X1 = torch.empty(rows, C1, device=dev)
    W = torch.randn(Co, C1 + C2, device=dev) * 0.1
    dX1 = [torch.empty(rows, C1, device=dev) for _ in range(NSETS)]
    dX2 = [torch.empty(rows, C2, device=dev) for _ in range(NSETS)] if C2 else [None] * NSETS
    X2 = torch.empty(rows, C2, device=dev) if C2 else None

    def rund(i):
        j = i % NSETS
        ops.linear_bwd(dY[j], None, None, 1.0, X1, W, X2=X2, dX1=dX1[j], dX2=dX2[j])

    t = timeit(rund)
    nb = rows * (C1 + C2 + Co) * 4
    print(f"dgrad [{Co} <- {C1}+{C2}] rows={rows}: {t:7.1f} us   {nb / t / 1e3:7.0f} GB/s   ({nb / 1e6:.0f} MB)")
