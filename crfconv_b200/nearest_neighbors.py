"""Drop-in for the reference's ``nearest_neighbors`` Cython module (utils/nearest_neighbors/knn.pyx:33-109), backed by
the sm_100a grid-hash kNN kernels.  Same names, arguments and return conventions:

    knn(pts[N,3], queries[Q,3], K, omp=False)              -> int64 [Q,K]
    knn_batch(pts[B,N,3], queries[B,Q,3], K, omp=False)    -> int64 [B,Q,K]
    knn_batch_distance_pick(pts[B,N,3], nqueries, K, omp=False, seed=None) -> (int64 [B,Q,K], float32 [B,Q,3])   (knn.pyx:111-149)

* numpy arrays / CPU tensors / anything ``np.ascontiguousarray(·, float32)`` accepts (knn.pyx:54-55,96-97) go through
  the host-pointer C-ABI call (H2D, search, D2H) and a fresh ``np.int64`` array is returned, like the reference.
* CUDA tensors stay on the device: a CUDA int64 tensor is returned, stream-ordered on the current stream (fast path,
  used by the on-GPU multiscale builder) — nothing crosses PCIe.
``omp`` is accepted for signature compatibility and ignored (the GPU path has no serial variant).
Results are the reference's: ascending squared-L2 distance with nanoflann's f32 operation order; equal distances are
ordered by index (nanoflann's own tie order is kd-tree-traversal dependent).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _workspace(nbytes, device):
    """Scratch for one kNN / grid-subsampling call.  Allocated per call from torch's caching allocator, which is stream-ordered:
    a block is only handed out again to work queued behind its previous use on the same stream, calls on different streams
    (a prefetch stream building the next pyramid, a CUDA-graph capture with its private pool) never share cell tables, and the
    allocator recycles the block without a cudaMalloc."""
    import torch
    return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)


def knn_batch_cuda(pts, queries, K):
    """pts [B,N,3], queries [B,Q,3] CUDA tensors → int64 CUDA tensor [B,Q,K]."""
    import torch
    L = _lib.lib()
    pts = pts.detach().to(torch.float32).contiguous()
    queries = queries.detach().to(torch.float32).contiguous()
    assert pts.dim() == 3 and queries.dim() == 3 and pts.shape[2] == 3 and queries.shape[2] == 3
    B, N, _ = pts.shape
    Q = queries.shape[1]
    out = torch.zeros((B, Q, K), dtype=torch.int64, device=pts.device)
    with torch.cuda.device(pts.device):
        nbytes = L.crfconv_knn_workspace_bytes(B, N, Q, K)
        ws = _workspace(nbytes, pts.device)
        rc = L.crfconv_knn_batch(pts.data_ptr(), B, N, queries.data_ptr(), Q, K, out.data_ptr(), ws.data_ptr(),
                                 ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "knn_batch")
    return out


def knn_batch(pts, queries, K, omp=False):
    if _is_cuda_tensor(pts):
        return knn_batch_cuda(pts, queries, K)
    L = _lib.lib()
    pts_c = np.ascontiguousarray(pts, dtype=np.float32)
    queries_c = np.ascontiguousarray(queries, dtype=np.float32)
    B, N, dim = pts_c.shape
    Q = queries_c.shape[1]
    indices = np.zeros((B, Q, K), dtype=np.int64)
    rc = L.crfconv_cpp_knn_batch(pts_c.ctypes.data, B, N, dim, queries_c.ctypes.data, Q, K, indices.ctypes.data)
    _lib.check(rc, "knn_batch")
    return indices


def knn(pts, queries, K, omp=False):
    if _is_cuda_tensor(pts):
        return knn_batch_cuda(pts[None], queries[None], K)[0]
    pts_c = np.ascontiguousarray(pts, dtype=np.float32)
    queries_c = np.ascontiguousarray(queries, dtype=np.float32)
    return knn_batch(pts_c[None], queries_c[None], K, omp)[0]


def knn_batch_distance_pick(pts, nqueries, K, omp=False, seed=None):
    """Coverage sampler of the reference (knn.pyx:111-149 → knn_.cxx:138-271): returns (indices [B,nqueries,K], queries [B,nqueries,3]).
    `seed` seeds the std::mt19937 stream (the reference uses time(0), which is what seed=None does); numpy / CPU input → numpy output
    through the host-pointer C ABI, CUDA tensors stay on the device."""
    import time
    L = _lib.lib()
    if seed is None:
        seed = int(time.time())
    seed = int(seed) & 0xFFFFFFFF
    if _is_cuda_tensor(pts):
        import torch
        p = pts.detach().to(torch.float32).contiguous()
        B, N, _ = p.shape
        idx = torch.zeros((B, nqueries, K), dtype=torch.int64, device=p.device)
        q = torch.zeros((B, nqueries, 3), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            ws = _workspace(L.crfconv_knn_distance_pick_workspace_bytes(B, N), p.device)
            rc = L.crfconv_knn_batch_distance_pick(p.data_ptr(), B, N, int(nqueries), int(K), seed, idx.data_ptr(), q.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "knn_batch_distance_pick")
        return idx, q
    pts_c = np.ascontiguousarray(pts, dtype=np.float32)
    B, N, dim = pts_c.shape
    indices = np.zeros((B, nqueries, K), dtype=np.int64)
    queries = np.zeros((B, nqueries, dim), dtype=np.float32)
    rc = L.crfconv_cpp_knn_batch_distance_pick(pts_c.ctypes.data, B, N, dim, queries.ctypes.data, int(nqueries), int(K), indices.ctypes.data, seed)
    _lib.check(rc, "knn_batch_distance_pick")
    return indices, queries


def radius_batch_cuda(pts, queries, r, K):
    """pts [B,N,3], queries [B,Q,3] CUDA → int64 [B,Q,K]: the up-to-K smallest support indices within distance r per query, -1 padded."""
    import torch
    L = _lib.lib()
    pts = pts.detach().to(torch.float32).contiguous()
    queries = queries.detach().to(torch.float32).contiguous()
    B, N, _ = pts.shape
    Q = queries.shape[1]
    out = torch.empty((B, Q, K), dtype=torch.int64, device=pts.device)
    with torch.cuda.device(pts.device):
        ws = _workspace(L.crfconv_knn_workspace_bytes(B, N, Q, K), pts.device)
        rc = L.crfconv_radius_batch(pts.data_ptr(), B, N, queries.data_ptr(), Q, float(r), int(K), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _lib.stream_ptr())
    _lib.check(rc, "radius_batch")
    return out
