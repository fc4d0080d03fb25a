"""On-GPU multiscale pyramid builder — the direct caller of the kNN kernel in the reference's data pipeline
(datasets/s3dis_dataset.py:416-449, semantic3d_dataset.py:501-534: ``_multiscale_compute_fn`` inside the DataLoader collate).
Per level: kNN(pos, pos, K) → random subsample (one ``randperm`` shared by the batch, like the reference) → sub_idx rows of
the kNN table → up_idx = kNN(sub_pos, pos, 1).  Everything stays on the device: no index tensor crosses PCIe.

Returns plain namespaces with the attribute names the reference's ``Data`` objects carry (pos, neighbor_idx, sub_idx,
up_idx), which is all ``PointConvResNet.forward`` reads (models/point_conv_big.py:142-167).
"""
from __future__ import annotations

import types

import torch

from . import nearest_neighbors


def build_multiscale(pos, num_scales=5, kernel_size=(16, 16, 16, 16, 16), ratio=(4, 4, 4, 4, 2), sample_method="random", generator=None):
    """pos: [B, N, 3] CUDA float tensor.  `generator` (CPU torch.Generator) makes the random choices reproducible."""
    if not pos.is_cuda:
        raise RuntimeError("build_multiscale runs on CUDA tensors only (no CPU fallback)")
    method = sample_method.lower()
    if method not in ("random", "fps"):
        raise NotImplementedError("Only `random` or `fps` sampling method is implemented!")       # the reference's message (:438)
    out = []
    pos = pos.float().contiguous()
    for i in range(num_scales):
        N = pos.shape[1]
        neighbor_idx = nearest_neighbors.knn_batch_cuda(pos, pos, kernel_size[i])          # [B, N, K]
        sample_num = N // ratio[i]
        if method == "random":
            choice = torch.randperm(N, generator=generator)[:sample_num].to(pos.device)     # shared by the whole batch (:424)
            sub_pos = pos[:, choice, :].contiguous()
            sub_idx = neighbor_idx[:, choice, :].contiguous()
        else:                                                                               # farthest point sampling per cloud (:434-437)
            from . import graph_ops
            choice = graph_ops.furthest_point_sampling(pos, sample_num)                     # [B, S]
            sub_pos = pos.gather(1, choice.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
            sub_idx = neighbor_idx.gather(1, choice.unsqueeze(-1).expand(-1, -1, neighbor_idx.shape[-1])).contiguous()
        up_idx = nearest_neighbors.knn_batch_cuda(sub_pos, pos, 1)                          # [B, N, 1]
        out.append(types.SimpleNamespace(pos=pos, neighbor_idx=neighbor_idx, sub_idx=sub_idx, up_idx=up_idx))
        pos = sub_pos
    return out
