"""A few eager fwd+loss+bwd steps of the full PointConvResNet (config C3: B x 40,960 points, 13 classes) for the ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/net_launches.csv python scripts/net_once.py 2 [B] [N] [classes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as Fn

from crfconv_b200 import losses, train_dp
from crfconv_b200.distributed import FlatGradients
from crfconv_b200.point_conv_big import PointConvResNet

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40960
ncls = int(sys.argv[4]) if len(sys.argv) > 4 else 13
dev = torch.device("cuda")
torch.manual_seed(1234)
net = PointConvResNet(6, ncls).to(dev).train()
grads = FlatGradients(net, direct=True)
pos, feat, lab, gen = train_dp.synthetic_shard(B, N, ncls, dev, seed=77)
data = train_dp.make_batch(pos, feat, lab, generator=gen)
target = (lab.reshape(-1) - 1).contiguous()
torch.cuda.synchronize()
for _ in range(steps):
    grads.zero()
    loss = losses.cross_entropy(net(data), target)
    loss.backward()
    torch.cuda.synchronize()
print("done", float(loss))
