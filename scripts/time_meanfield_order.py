"""How much do the mean-field kernels gain when the points of a cloud arrive in a spatially coherent order?  (development aid)
Same clouds, same kernels; 'random' = points in generation order, 'morton' = points sorted by a 10-bit/axis Morton key before
the neighbour search (so neighbor_idx / up_idx refer to the sorted order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crfconv_b200 import nearest_neighbors as nn_, ops

B, N, K, F = 6, 40960, 16, 16
dev = "cuda"
torch.manual_seed(0)


def part1by2(x):
    x = x & 0x3FF
    x = (x | (x << 16)) & 0x30000FF
    x = (x | (x << 8)) & 0x300F00F
    x = (x | (x << 4)) & 0x30C30C3
    x = (x | (x << 2)) & 0x9249249
    return x


def morton(pos):
    lo, hi = pos.amin(dim=1, keepdim=True), pos.amax(dim=1, keepdim=True)
    q = ((pos - lo) / (hi - lo + 1e-9) * 1023).long()
    return part1by2(q[..., 0]) | (part1by2(q[..., 1]) << 1) | (part1by2(q[..., 2]) << 2)


def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


# surface-like clouds: points on the faces of a box (what indoor scans look like), random order
u = torch.rand(B, N, 2, device=dev)
face = torch.randint(0, 3, (B, N), device=dev)
pos = torch.zeros(B, N, 3, device=dev)
for f in range(3):
    m = face == f
    a, b = [(1, 2), (0, 2), (0, 1)][f]
    p = torch.zeros(B, N, 3, device=dev)
    p[..., a] = u[..., 0]; p[..., b] = u[..., 1]; p[..., f] = (torch.rand(B, N, device=dev) > 0.5).float()
    pos[m] = p[m]
pos = pos * torch.tensor([8.0, 6.0, 3.0], device=dev)
big = torch.empty(48 * 1024 * 1024, device=dev)

for mode in ("random", "morton"):
    p = pos
    if mode == "morton":
        order = morton(pos).argsort(dim=1)
        p = torch.gather(pos, 1, order[..., None].expand(-1, -1, 3)).contiguous()
    nbr = nn_.knn_batch(p, p, K)
    sub = p[:, ::4].contiguous()                          # coarse level: every 4th point (keeps the order's coherence)
    up = nn_.knn_batch(sub, p, 1).reshape(B, N)
    Hy = torch.randn(B * N, F, device=dev); z = torch.randn(B * N, F, device=dev)
    sc = torch.rand(F, device=dev) * 0.2 + 0.2
    c = torch.eye(F, device=dev) + 0.1 * torch.randn(F, F, device=dev)
    Cm, Minv = ops.crf_compat_fwd(c)
    g = torch.randn(B * N, F, device=dev)
    Gz, m_out, v_out, h_out = (torch.empty(B * N, F, device=dev) for _ in range(4))
    gprev, Gy = torch.zeros(B * N, F, device=dev), torch.zeros(B * N, F, device=dev)
    tf = timeit(lambda: (big.zero_(), ops.crf_step_fwd(Hy, sc, z, z, nbr, Cm, Minv, B, N, K))) - timeit(lambda: big.zero_())
    tb = timeit(lambda: (big.zero_(), ops.crf_step_bwd(Hy, sc, z, z, nbr, Cm, Minv, g, Gz, gprev, Gy, m_out, v_out, h_out, False, B, N, K))) - timeit(lambda: big.zero_())
    print(f"{mode:7s}: crf_step_fwd {tf:6.1f} us   crf_step_bwd {tb:6.1f} us   (L2 flushed between launches)")
