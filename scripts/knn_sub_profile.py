"""kNN (6 x 40,960 points, K = 16 and the K = 1 up-index query) and grid subsampling (100 k / 1 M points) for the ncu captures and a
per-kernel timing of the subsampling pipeline:
   python scripts/knn_sub_profile.py time            per-phase CUDA-event timing (subsampling at 100 k and 1 M points)
   ncu --set full --csv --page raw --log-file x.csv python scripts/knn_sub_profile.py once"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from crfconv_b200 import grid_subsampling as gs, nearest_neighbors as nn_
from oracle import synthetic

mode = sys.argv[1] if len(sys.argv) > 1 else "time"
B, N = 6, 40960
pos = torch.from_numpy(synthetic.room_cloud(B, N, seed=1)).cuda()
sub = pos[:, ::4].contiguous()
rng = np.random.default_rng(0)


def sub_inputs(n):
    p = torch.from_numpy(synthetic.room_cloud(1, n, seed=n)[0]).cuda()
    f = torch.from_numpy(rng.integers(0, 256, (n, 3)).astype(np.float32)).cuda()
    c = torch.from_numpy(rng.integers(0, 13, (n,)).astype(np.int32)).cuda()
    return p, f, c


if mode == "once":
    nn_.knn_batch(pos, pos, 16)
    nn_.knn_batch(sub, pos, 1)
    for n, dl in ((100000, 0.04), (1000000, 0.06)):
        p, f, c = sub_inputs(n)
        gs.compute(p, features=f, classes=c, sampleDl=dl, order="key")
    torch.cuda.synchronize()
    print("done")
else:
    def t(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    us = t(lambda: nn_.knn_batch(pos, pos, 16))
    print(f"knn_batch self K=16: {us:.0f} us  {B * N / us:.1f} M queries/s")
    us = t(lambda: nn_.knn_batch(sub, pos, 1))
    print(f"knn_batch up-index K=1 (support N/4): {us:.0f} us  {B * N / us:.1f} M queries/s")
    for K in (32, 64, 128):
        us = t(lambda: nn_.knn_batch(pos, pos, K), reps=3)
        print(f"knn_batch self K={K}: {us:.0f} us  {B * N / us:.1f} M queries/s")
    for n, dl in ((10000, 0.04), (100000, 0.04), (1000000, 0.06), (1000000, 0.04)):
        p, f, c = sub_inputs(n)
        us = t(lambda: gs.compute(p, features=f, classes=c, sampleDl=dl, order="key"), reps=5)
        m = gs.compute(p, features=f, classes=c, sampleDl=dl, order="key")[0].shape[0]
        print(f"grid_subsample N={n} dl={dl}: {us:.0f} us  {n / us:.1f} M points/s  ({m} voxels; 13 RANDOM labels: most multi-point voxels tie)")
        c2 = ((p[:, 0] * 2).floor().to(torch.int32) + (p[:, 1] * 2).floor().to(torch.int32) * 3) % 13      # spatially coherent labels (0.5 m blocks)
        us = t(lambda: gs.compute(p, features=f, classes=c2, sampleDl=dl, order="key"), reps=5)
        print(f"grid_subsample N={n} dl={dl}: {us:.0f} us  {n / us:.1f} M points/s  (spatially coherent labels: ties are rare)")
        us = t(lambda: gs.compute(p, features=f, sampleDl=dl, order="key"), reps=5)
        print(f"grid_subsample N={n} dl={dl}: {us:.0f} us  {n / us:.1f} M points/s  (no labels)")
