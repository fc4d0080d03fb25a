"""Full PointConvResNet fwd+bwd, B = 6 x 40,960 points: eager launches vs CUDA-graph replay (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from crfconv_b200 import train_dp
from crfconv_b200.distributed import FlatGradients
from crfconv_b200.graphs import GraphedStep
from crfconv_b200.point_conv_big import PointConvResNet

B, N = int(os.environ.get("B", 6)), 40960
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = PointConvResNet(6, 13).to(dev).train()
grads = FlatGradients(model)
pos, feats, labels, gen = train_dp.synthetic_shard(B, N, 13, dev, seed=1)
data = train_dp.make_batch(pos, feats, labels, generator=gen)


def fwd_bwd():
    grads.zero()
    loss = F.cross_entropy(model(data), data.y.reshape(-1) - 1)
    loss.backward()
    return loss.detach()


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


te = timeit(fwd_bwd)
step = GraphedStep(fwd_bwd)
tg = timeit(step.replay)
tm = timeit(lambda: train_dp.make_batch(pos, feats, labels, generator=gen))
print(f"PointConvResNet fwd+bwd B={B} N={N}: eager {te:.2f} ms ({B*N/te/1e3:.1f} M points/s), graph replay {tg:.2f} ms ({B*N/tg/1e3:.1f} M points/s); "
      f"5-level pyramid build {tm:.2f} ms; peak memory {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")

if os.environ.get("PROFILE", "0") == "1":
    from crfconv_b200 import ops
    prof = ops.profile_calls(fwd_bwd, repeats=2)
    tot = sum(v["ms"] for v in prof.values())
    print(f"per-call sum {tot:.2f} ms over {sum(v['calls'] for v in prof.values()):.0f} calls")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:28]:
        print(f"  {k:34s} {v['ms']:7.3f} ms  x{v['calls']:<5.0f} {v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.0f} GB/s")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        fwd_bwd(); torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
