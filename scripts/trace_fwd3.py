"""Timeline of one CTA of the tcgen05 forward kernel (needs a library built with CRFCONV_NVCC_EXTRA=-DCRF_FWD3_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from crfconv_b200 import ops, _lib

L = _lib.lib()
fn = L.crfconv_debug_fwd3_trace
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
M, C1, C2, Co = 6 * 40960, int(os.environ.get("C1", 64)), int(os.environ.get("C2", 64)), int(os.environ.get("CO", 64))
dev = "cuda"
X1 = torch.randn(M, C1, device=dev); X2 = torch.randn(M, C2, device=dev) if C2 else None
W = torch.randn(Co, C1 + C2, device=dev) * 0.1
sc = torch.rand(C1, device=dev) + 0.5; sh = torch.randn(C1, device=dev) * 0.1
stats = torch.zeros(ops.STAT_SLOTS * 2 * Co, device=dev)
Y = torch.empty(M, Co, device=dev)
big = torch.empty(64 * 1024 * 1024, device=dev)
for it in range(3):
    big.zero_()                                   # flush L2
    torch.cuda.synchronize()
    fn(None, 0, 1)
    ops.linear_fwd(X1, W, scale1=sc, shift1=sh, slope1=0.1, X2=X2, stats=stats, out=Y)
    torch.cuda.synchronize()
buf = np.zeros(3 * 8192, dtype=np.int64)
n = fn(buf.ctypes.data, 8192, 0)
rec = buf[: 3 * n].reshape(n, 3)
rec = rec[rec[:, 0] > 0]
n = len(rec)
t0 = rec[:, 2].min()
names = {1: "P loads issued", 2: "P slot acquired", 3: "P converted", 4: "P published", 5: "I saw slab", 6: "I mma issued", 7: "E acc complete",
         8: "E tmem read", 9: "E tile stored", 10: "I loop top", 11: "I slab visible", 12: "I desc", 13: "I mmas issued"}
rec = rec[np.argsort(rec[:, 2], kind="stable")]
print(f"{n} records; total span {(rec[:, 2].max() - t0) / 1.9e3:.1f} us (at 1.9 GHz)")
lim = int(os.environ.get("LIMIT", 0))
for ev, idx, t in rec[:lim]:
    print(f"{(t - t0) / 1.9e3:8.2f} us  {names[int(ev)]:16s} {int(idx)}")
# per-slab summary
import collections
per = collections.defaultdict(dict)
for ev, idx, t in rec:
    if int(ev) <= 9: per[(int(ev) <= 6, int(idx))][int(ev)] = (t - t0) / 1.9e3
print("slab: issued -> slot -> converted -> published | issuer saw -> issued")
for q in range(0, 52):
    d = per.get((True, q), {})
    print(q, " ".join(f"{d.get(k, float('nan')):7.2f}" for k in (1, 2, 3, 4, 5, 6)))
print("issuer per slab: loop top -> slab visible -> elected(5) -> desc ready -> mmas issued -> committed(6)")
fine = collections.defaultdict(dict)
for ev, idx, t in rec:
    if int(ev) >= 10 or int(ev) in (5, 6):
        fine[int(idx)][int(ev)] = (t - t0) / 1.9e3
for q in range(8, 40):
    d = fine.get(q, {})
    print(q, " ".join(f"{d.get(k, float('nan')):7.2f}" for k in (10, 11, 5, 12, 13, 6)))
print("tile: acc complete -> tmem read -> stored")
for ti in range(13):
    d = per.get((False, ti), {})
    print(ti, " ".join(f"{d.get(k, float('nan')):7.2f}" for k in (7, 8, 9)))
