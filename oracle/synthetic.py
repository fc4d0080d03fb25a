"""TEST INFRASTRUCTURE ONLY (oracle).  Seeded synthetic inputs of the shapes named in SURVEY.md §8(d), and a CPU
restatement of the reference's multiscale builder (datasets/s3dis_dataset.py:416-449) parameterised by the kNN
function (oracle brute force, compiled reference, or — in GPU tests — the product, to cross-check it)."""
from __future__ import annotations

import types

import numpy as np
import torch


def room_cloud(B, N, seed=0, box=(8.0, 6.0, 3.0)):
    """Uniform points in an S3DIS-room-like box.  [B,N,3] f32."""
    rng = np.random.default_rng(seed)
    return (rng.random((B, N, 3)) * np.asarray(box)).astype(np.float32)


def lattice_cloud(n=12, step=0.25):
    """n³ lattice: every query has massive distance ties."""
    g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
    return (g * step).astype(np.float32)


def duplicated_cloud(N, seed=0, box=(8.0, 6.0, 3.0)):
    """Half of the points are exact duplicates of the other half (S3DIS pads small rooms with duplicates,
    s3dis_dataset.py:376-377)."""
    rng = np.random.default_rng(seed)
    half = (rng.random((N // 2, 3)) * np.asarray(box)).astype(np.float32)
    pts = np.concatenate([half, half[: N - N // 2]], 0)
    return pts[rng.permutation(N)]


def build_multiscale(pos, knn_batch_fn, num_scales=5, K=16, ratios=(4, 4, 4, 4, 2), seed=0):
    """pos [B,N,3] f32 numpy.  Mirrors _multiscale_compute_fn with sample_method='random': one `choice` shared by the
    whole batch per level.  Returns a list of namespaces (pos, neighbor_idx, sub_idx, up_idx) of torch CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    out = []
    pos = torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float32))
    for i in range(num_scales):
        nbr = torch.from_numpy(np.asarray(knn_batch_fn(pos.numpy(), pos.numpy(), K)))
        n_sub = pos.shape[1] // ratios[i]
        choice = torch.randperm(pos.shape[1], generator=g)[:n_sub]
        sub_pos = pos[:, choice, :].contiguous()
        sub_idx = nbr[:, choice, :].contiguous()
        up_idx = torch.from_numpy(np.asarray(knn_batch_fn(sub_pos.numpy(), pos.numpy(), 1)))
        out.append(types.SimpleNamespace(pos=pos, neighbor_idx=nbr, sub_idx=sub_idx, up_idx=up_idx))
        pos = sub_pos
    return out


def crf_layer_inputs(B, N, K=16, Cu=128, Cp=64, ratio=4, seed=0, knn_batch_fn=None):
    """Config C1/S1 of SURVEY.md §8(d): unary [B,N/ratio,Cu], pairwise [B,N,Cp], up_idx [B,N,1], neighbor_idx [B,N,K]."""
    pos = room_cloud(B, N, seed)
    g = torch.Generator().manual_seed(seed)
    nbr = torch.from_numpy(np.asarray(knn_batch_fn(pos, pos, K)))
    choice = torch.randperm(N, generator=g)[: N // ratio]
    sub_pos = np.ascontiguousarray(pos[:, choice.numpy(), :])
    up_idx = torch.from_numpy(np.asarray(knn_batch_fn(sub_pos, pos, 1)))
    unary = torch.randn(B, N // ratio, Cu, generator=g)
    pairwise = torch.randn(B, N, Cp, generator=g)
    return types.SimpleNamespace(pos=torch.from_numpy(pos), unary=unary, pairwise=pairwise, up_idx=up_idx, neighbor_idx=nbr)
