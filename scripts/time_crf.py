"""Quick device timing of the CRF layer (S1 shape) and kNN — development aid, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crfconv_b200 import nearest_neighbors as nn_
from crfconv_b200.continuous_crf_conv_big import ContinuousGaussianCRFConv

B, N, K = int(os.environ.get("B", 6)), 40960, 16
dev = "cuda"
torch.manual_seed(0)
pos = torch.rand(B, N, 3, device=dev) * torch.tensor([8.0, 6.0, 3.0], device=dev)

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t = timeit(lambda: nn_.knn_batch(pos, pos, K))
print(f"knn_batch B={B} N={N} K={K}: {t:.3f} ms  -> {B*N/t/1e3:.1f} M queries/s")
nbr = nn_.knn_batch(pos, pos, K)
choice = torch.randperm(N, device=dev)[: N // 4]
sub = pos[:, choice].contiguous()
t = timeit(lambda: nn_.knn_batch(sub, pos, 1))
print(f"knn_batch up_idx (K=1, support N/4): {t:.3f} ms")
up = nn_.knn_batch(sub, pos, 1)
layer = ContinuousGaussianCRFConv(128, 64, 64, steps=1).to(dev).train()
unary = torch.randn(B, N // 4, 128, device=dev, requires_grad=True)
pair = torch.randn(B, N, 64, device=dev, requires_grad=True)
def fwd():
    return layer(unary, pair, up, nbr)
def fwdbwd():
    out = layer(unary, pair, up, nbr)
    out.backward(torch.ones_like(out))
tf = timeit(fwd)
tfb = timeit(fwdbwd)
A = 79298560 * B
print(f"CRF fwd {tf:.3f} ms, fwd+bwd {tfb:.3f} ms -> {B*N/tfb/1e3:.2f} M points/s, algorithmic {A/tfb/1e6:.1f} GB/s ({A/tfb/1e6/6534*100:.1f}% of measured HBM peak)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fwdbwd(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
