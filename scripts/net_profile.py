"""Per-call CUDA-event profile of one eager PointConvResNet step (config C3 by default): which C-ABI calls (with their shapes) carry
the time.   python scripts/net_profile.py [B] [N] [classes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as Fn

from crfconv_b200 import ops, train_dp
from crfconv_b200.distributed import FlatGradients
from crfconv_b200.point_conv_big import PointConvResNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40960
ncls = int(sys.argv[3]) if len(sys.argv) > 3 else 13
dev = torch.device("cuda")
torch.manual_seed(1234)
net = PointConvResNet(6, ncls).to(dev).train()
grads = FlatGradients(net, direct=True)
pos, feat, lab, gen = train_dp.synthetic_shard(B, N, ncls, dev, seed=77)
data = train_dp.make_batch(pos, feat, lab, generator=gen)
target = (lab.reshape(-1) - 1).contiguous()


def step():
    grads.zero()
    Fn.cross_entropy(net(data), target).backward()


prof = ops.profile_calls(step, repeats=2)
tot = sum(v["ms"] for v in prof.values())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"eager step {e0.elapsed_time(e1):.2f} ms; sum of crfconv_b200 calls {tot:.2f} ms over {sum(v['calls'] for v in prof.values()):.0f} calls")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:45]:
    print(f"{k:44s} {v['ms'] * 1e3:9.1f} us  calls {v['calls']:5.1f}  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f} GB/s")
