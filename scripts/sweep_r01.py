"""Round-1 measurement sweep for BASELINE configs[1] (kNN + grid subsampling, N = 10k … 1M) and configs[2] (full network,
6 x 40,960 points): product on the GPU vs the reference's CPU path on the box's host cores.  Writes gpurun_out/sweep_r01.json."""
import json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crfconv_b200 import nearest_neighbors as nn_, grid_subsampling as gs, multiscale
from crfconv_b200.point_conv_big import PointConvResNet
from oracle import native as on, synthetic, layers as ol

def gpu_time(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

def cpu_time(fn, n=1):
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n

res = {"cores": os.cpu_count(), "knn": [], "subsample": [], "net": {}}
for N in (10000, 40960, 100000, 400000, 1000000):
    pts = synthetic.room_cloud(1, N, seed=N)
    d = torch.from_numpy(pts).cuda()
    for K in (16, 32):
        tg = gpu_time(lambda: nn_.knn_batch(d, d, K))
        th = cpu_time(lambda: nn_.knn(pts[0], pts[0], K), 2)                      # host-pointer C ABI: H2D + search + D2H
        tr = cpu_time(lambda: on.ref_knn(pts[0], pts[0], K, omp=True)) if (on.have_ref_knn() and (K == 16 or N <= 100000)) else None
        res["knn"].append({"N": N, "K": K, "gpu_qps": N / tg, "host_api_qps": N / th, "ref_cpu_qps": (N / tr) if tr else None})
        print(res["knn"][-1], flush=True)
rng = np.random.default_rng(0)
for N, dl in ((10000, 0.04), (100000, 0.04), (1000000, 0.06)):
    pts = synthetic.room_cloud(1, N, seed=N)[0]
    f = rng.integers(0, 256, (N, 3)).astype(np.float32); c = rng.integers(0, 13, (N,)).astype(np.int32)
    dp, df, dc = torch.from_numpy(pts).cuda(), torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()
    tg = gpu_time(lambda: gs.compute(dp, features=df, classes=dc, sampleDl=dl, order="key"), n=3, warm=1)
    th = cpu_time(lambda: gs.compute(pts, features=f, classes=c, sampleDl=dl), 2)   # host API, reference row order
    tr = cpu_time(lambda: on.ref_grid_subsample(pts, f, c, dl)) if on.have_ref_subsample() else None
    res["subsample"].append({"N": N, "dl": dl, "gpu_key_order_pps": N / tg, "host_api_ref_order_pps": N / th, "ref_cpu_pps": (N / tr) if tr else None})
    print(res["subsample"][-1], flush=True)
# ---- full network (config 3): B = 6, N = 40,960, 13 classes
B, N = 6, 40960
pos = torch.from_numpy(synthetic.room_cloud(B, N, seed=1)).cuda()
t_ms = gpu_time(lambda: multiscale.build_multiscale(pos, generator=torch.Generator().manual_seed(0)), n=3, warm=1)
ms = multiscale.build_multiscale(pos, generator=torch.Generator().manual_seed(0))
torch.manual_seed(0)
net = PointConvResNet(6, 13).cuda().train()
x = torch.cat([pos, torch.rand(B, N, 3, device="cuda")], -1)
y = torch.randint(0, 13, (B * N,), device="cuda")
data = types.SimpleNamespace(x=x, multiscale=ms)
def fb():
    net.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(net(data), y).backward()
t_net = gpu_time(fb, n=3, warm=2)
res["net"] = {"B": B, "N": N, "multiscale_build_s": t_ms, "fwd_bwd_s": t_net, "points_per_s": B * N / t_net, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
print(res["net"], flush=True)
# CPU port of the same network on 2 clouds (bounded sample)
Bc = 2
msc = [types.SimpleNamespace(pos=l.pos[:Bc].cpu(), neighbor_idx=l.neighbor_idx[:Bc].cpu(), sub_idx=l.sub_idx[:Bc].cpu(), up_idx=l.up_idx[:Bc].cpu()) for l in ms]
onet = ol.PointConvResNet(6, 13).train(); onet.load_state_dict(net.state_dict())
dc_ = types.SimpleNamespace(x=x[:Bc].cpu(), multiscale=msc)
def fbc():
    onet.zero_grad(); torch.nn.functional.cross_entropy(onet(dc_), y[:Bc * N].cpu()).backward()
fbc()
tc = cpu_time(fbc, 2)
res["net"]["cpu_port_points_per_s"] = Bc * N / tc
res["net"]["cpu_knn_pyramid_s_per_cloud"] = None
print(res["net"], flush=True)
json.dump(res, open("gpurun_out/sweep_r01.json", "w"), indent=1)
