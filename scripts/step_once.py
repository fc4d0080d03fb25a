"""A few eager fwd+bwd steps of the benchmark layer (S1: 6 clouds x 40,960 points) and nothing else — the target of the ncu captures:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/step_once.py 3
  ncu --set full --clock-control none --import-source on -o gpurun_out/full python scripts/step_once.py 2

(the first step also pays for lazy initialisation; read the LAST step of a capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from crfconv_b200.continuous_crf_conv_big import ContinuousGaussianCRFConv

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda")
torch.manual_seed(1234)
layer = ContinuousGaussianCRFConv(bench.CU, bench.CP, bench.CP, steps=1).to(dev).train()
with torch.no_grad():
    layer.c.add_(0.1 * torch.randn_like(layer.c))
s = bench.make_inputs(torch, B, dev, seed=0)
s["unary"].requires_grad_(True)
s["pairwise"].requires_grad_(True)
cot = torch.ones(B, bench.N_POINTS, bench.CP, device=dev)
for _ in range(steps):
    layer.zero_grad(set_to_none=True)
    out = layer(s["unary"], s["pairwise"], s["up_idx"], s["neighbor_idx"])
    out.backward(cot)
    s["unary"].grad = None
    s["pairwise"].grad = None
    torch.cuda.synchronize()
print("done")
