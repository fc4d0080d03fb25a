import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crfconv_b200 import nearest_neighbors as nn_
from oracle import synthetic
for B, N, K in ((6, 40960, 16), (1, 1000000, 16), (6, 40960, 32)):
    pos = torch.from_numpy(synthetic.room_cloud(B, N, 1)).cuda()
    q2 = pos.clone()
    for name, q in (("self", pos), ("other", q2)):
        for _ in range(3): nn_.knn_batch(pos, q, K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): nn_.knn_batch(pos, q, K)
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} N={N} K={K} {name}: {10*B*N/(e0.elapsed_time(e1)*1e-3)/1e6:.0f} M queries/s")
