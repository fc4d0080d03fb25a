"""Graph builders of the reference's PyG-API family on the sm_100a kernels — the functions ``models/point_conv.py:140-195,267-280``,
``models/continuous_crf_conv.py:52`` and ``models/discrete_crf_conv.py:44`` import from torch_geometric / torch_cluster /
torch_points_kernels (third-party packages that are not part of the reference tree; their semantics are restated here and in the
CPU checker used by the tests; parity with those packages is unpinned):

    furthest_point_sampling(pos[B,N,3], nsamples)                      -> int64 [B, nsamples]      (datasets/s3dis_dataset.py:435)
    fps(pos[N,3], batch, ratio, random_start=False)                    -> int64 [sum ceil(ratio·n_b)]
    knn(x, y, k, batch_x, batch_y) / knn_graph(pos, k, batch, loop)    -> int64 [2, E]
    radius(x, y, r, batch_x, batch_y, max_num_neighbors) / radius_graph(pos, r, batch, loop, max_num_neighbors) -> int64 [2, E]
    knn_interpolate(x, pos_x, pos_y, batch_x, batch_y, k=3)            -> [N_y, C]   (differentiable wrt x)

Conventions follow torch_cluster: ``knn`` / ``radius`` return ``[row = index into y (query), col = index into x]`` grouped by query;
the ``*_graph`` forms return ``[source j, target i]`` (flow = source_to_target).  ``radius`` keeps, per query, the first
``max_num_neighbors`` support points in ascending index order (torch_cluster's CUDA kernel); distances use the kNN arithmetic
(f32, no FMA).  ``batch`` vectors must be sorted (PyG convention).  CUDA tensors only — there is no CPU fallback.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from . import nearest_neighbors as nn_


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("crfconv_b200 graph builders run on CUDA tensors only (no CPU fallback)")


def _ptr(batch, n, device):
    """CSR offsets [B+1] of a sorted batch vector (None ⇒ one cloud)."""
    if batch is None:
        return torch.tensor([0, n], dtype=torch.int64, device=device)
    counts = torch.bincount(batch.to(torch.int64))
    return torch.cat([counts.new_zeros(1), counts.cumsum(0)])


def _clouds(ptr):
    p = ptr.tolist()
    return [(p[i], p[i + 1]) for i in range(len(p) - 1)]


def fps_ptr(pos, ptr, nsample, start=None):
    """pos [N,3]; ptr [B+1] (tensor); nsample: int or per-cloud list → flat int64 global indices, cloud after cloud."""
    _need_cuda(pos)
    L = _lib.lib()
    pos = pos.detach().to(torch.float32).contiguous()
    B = ptr.numel() - 1
    ns = torch.as_tensor(nsample, dtype=torch.int64).expand(B).contiguous().to(pos.device) if not torch.is_tensor(nsample) else nsample.to(pos.device, torch.int64)
    st = torch.zeros(B, dtype=torch.int64, device=pos.device) if start is None else start.to(pos.device, torch.int64)
    optr = torch.cat([ns.new_zeros(1), ns.cumsum(0)])
    out = torch.empty(int(optr[-1].item()), dtype=torch.int64, device=pos.device)
    dist = torch.empty(pos.shape[0], dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        rc = L.crfconv_fps(pos.data_ptr(), ptr.data_ptr(), B, ns.data_ptr(), st.data_ptr(), out.data_ptr(), optr.data_ptr(), dist.data_ptr(),
                           _lib.stream_ptr())
    _lib.check(rc, "fps")
    return out


def furthest_point_sampling(pos, nsamples):
    """torch_points_kernels.furthest_point_sampling: pos [B,N,3] → LOCAL indices [B, nsamples], starting from point 0 of each cloud."""
    B, N, _ = pos.shape
    ptr = torch.arange(B + 1, dtype=torch.int64, device=pos.device) * N
    flat = fps_ptr(pos.reshape(B * N, 3), ptr, int(nsamples))
    return flat.view(B, nsamples) - ptr[:-1, None]


def fps(pos, batch=None, ratio=0.5, random_start=False):
    """torch_cluster.fps: ceil(ratio·n_b) samples per cloud, flat global indices.  random_start=False starts at the first point of
    each cloud (torch_cluster's default is a random start, which has no reproducible reference value)."""
    ptr = _ptr(batch, pos.shape[0], pos.device)
    n = ptr[1:] - ptr[:-1]
    ns = torch.ceil(n.double() * ratio).to(torch.int64)
    start = (torch.rand(n.numel(), device=pos.device) * n).to(torch.int64) if random_start else None
    return fps_ptr(pos, ptr, ns, start)


def _search(x, y, ptr_x, ptr_y, K, r=None):
    """Per cloud: [Q_b, K] neighbour table (global x indices; -1 padding in radius mode)."""
    out = []
    cx, cy = _clouds(ptr_x), _clouds(ptr_y)
    sizes = {(b[1] - b[0], q[1] - q[0]) for b, q in zip(cx, cy)}
    if len(sizes) == 1 and len(cx) > 1:                       # equal-sized clouds: one batched call
        (n, q), = sizes
        B = len(cx)
        xs, ys = x.reshape(B, n, 3), y.reshape(B, q, 3)
        tab = nn_.radius_batch_cuda(xs, ys, r, K) if r is not None else nn_.knn_batch_cuda(xs, ys, K)
        off = ptr_x[:-1].view(B, 1, 1)
        return torch.where(tab >= 0, tab + off, tab).reshape(B * q, K)
    for (x0, x1), (y0, y1) in zip(cx, cy):
        xs, ys = x[x0:x1][None], y[y0:y1][None]
        tab = (nn_.radius_batch_cuda(xs, ys, r, K) if r is not None else nn_.knn_batch_cuda(xs, ys, K))[0]
        out.append(torch.where(tab >= 0, tab + x0, tab))
    return torch.cat(out, 0)


def knn(x, y, k, batch_x=None, batch_y=None):
    """torch_cluster.knn: for every y_i its k nearest x_j of the same cloud → [2, N_y·k] = [row (y index), col (x index)]."""
    _need_cuda(x)
    x, y = x.detach().float().contiguous(), y.detach().float().contiguous()
    tab = _search(x, y, _ptr(batch_x, x.shape[0], x.device), _ptr(batch_y, y.shape[0], y.device), k)
    row = torch.arange(y.shape[0], device=x.device).repeat_interleave(k)
    return torch.stack([row, tab.reshape(-1)])


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """torch_cluster.radius: [row (y index), col (x index)], per query the first max_num_neighbors x_j within r in index order."""
    _need_cuda(x)
    x, y = x.detach().float().contiguous(), y.detach().float().contiguous()
    tab = _search(x, y, _ptr(batch_x, x.shape[0], x.device), _ptr(batch_y, y.shape[0], y.device), max_num_neighbors, r=r)
    mask = tab >= 0
    row = torch.arange(y.shape[0], device=x.device)[:, None].expand_as(tab)[mask]
    return torch.stack([row, tab[mask]])


def _to_graph(edge, loop, flow):
    row, col = edge[0], edge[1]                               # row = target (query), col = source
    if not loop:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([col, row]) if flow == "source_to_target" else torch.stack([row, col])


def knn_graph(pos, k, batch=None, loop=False, flow="source_to_target"):
    """torch_cluster.knn_graph: edge_index [2, E] = [source j, target i]."""
    return _to_graph(knn(pos, pos, k if loop else k + 1, batch, batch), loop, flow)


def radius_graph(pos, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target"):
    """torch_cluster.radius_graph: edge_index [2, E] = [source j, target i] (searches max_num_neighbors + 1 when loop=False, then
    drops the self edges — like torch_cluster)."""
    return _to_graph(radius(pos, pos, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1), loop, flow)


def csr_by_target(target, source, num_nodes):
    """Edge list grouped by target → (eptr [N+1], col [E], perm) with perm the stable edge permutation that groups the edges."""
    perm = torch.argsort(target, stable=True)
    counts = torch.bincount(target, minlength=num_nodes)
    eptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).contiguous()
    return eptr, source[perm].contiguous(), perm


# ------------------------------------------------------------------------------------------------ differentiable edge ops
def _p(t):
    return None if t is None else t.data_ptr()


class EdgeSoftmax(torch.autograd.Function):
    """s[e] = softmax over the edges of each target node of −‖y_target − y_source‖² (continuous_crf_conv.py:55-56,115-116)."""

    @staticmethod
    def forward(ctx, y, eptr, col):
        _need_cuda(y)
        yc = y.detach().float().contiguous()
        s = torch.empty(col.numel(), dtype=torch.float32, device=y.device)
        _lib.check(_lib.lib().crfconv_edge_softmax_fwd(_p(yc), _p(eptr), _p(col), _p(s), yc.shape[0], yc.shape[1], _lib.stream_ptr()), "edge_softmax_fwd")
        ctx.save_for_backward(yc, eptr, col, s)
        return s

    @staticmethod
    def backward(ctx, ds):
        yc, eptr, col, s = ctx.saved_tensors
        dy = torch.zeros_like(yc)
        _lib.check(_lib.lib().crfconv_edge_softmax_bwd(_p(yc), _p(eptr), _p(col), _p(s), _p(ds.contiguous().float()), _p(dy), yc.shape[0], yc.shape[1],
                                                      _lib.stream_ptr()), "edge_softmax_bwd")
        return dy, None, None


class SpMM(torch.autograd.Function):
    """out[i] = Σ_{edges e of target i} w[e]·x[col[e]]  — scatter_add(w ⊙ x[col], row) (continuous_crf_conv.py:64-65)."""

    @staticmethod
    def forward(ctx, w, x, eptr, col):
        _need_cuda(x)
        wc, xc = w.detach().float().contiguous(), x.detach().float().contiguous()
        N = eptr.numel() - 1
        out = torch.empty((N, xc.shape[1]), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().crfconv_spmm_fwd(_p(xc), _p(eptr), _p(col), _p(wc), _p(out), N, xc.shape[1], _lib.stream_ptr()), "spmm_fwd")
        ctx.save_for_backward(wc, xc, eptr, col)
        return out

    @staticmethod
    def backward(ctx, g):
        wc, xc, eptr, col = ctx.saved_tensors
        dw = torch.empty_like(wc) if ctx.needs_input_grad[0] else None
        dx = torch.zeros_like(xc) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.lib().crfconv_spmm_bwd(_p(xc), _p(eptr), _p(col), _p(wc), _p(g.contiguous().float()), _p(dw), _p(dx), eptr.numel() - 1,
                                              xc.shape[1], _lib.stream_ptr()), "spmm_bwd")
        return dw, dx, None, None


class EdgeGauss(torch.autograd.Function):
    """w[e] = Σ_k Wk[k]·exp(−‖f_k[source] − f_k[target]‖²), f [N, Kk, H] (discrete_crf_conv.py:49-56)."""

    @staticmethod
    def forward(ctx, f, Wk, eptr, col):
        _need_cuda(f)
        fc, wk = f.detach().float().contiguous(), Wk.detach().float().contiguous().view(-1)
        N, Kk, H = fc.shape
        E = col.numel()
        w = torch.empty(E, dtype=torch.float32, device=f.device)
        g = torch.empty((E, Kk), dtype=torch.float32, device=f.device)
        _lib.check(_lib.lib().crfconv_edge_gauss_fwd(_p(fc), _p(eptr), _p(col), _p(wk), _p(w), _p(g), N, Kk, H, _lib.stream_ptr()), "edge_gauss_fwd")
        ctx.save_for_backward(fc, wk, eptr, col, g)
        ctx.wshape = Wk.shape
        return w

    @staticmethod
    def backward(ctx, dw):
        fc, wk, eptr, col, g = ctx.saved_tensors
        N, Kk, H = fc.shape
        df, dWk = torch.zeros_like(fc), torch.zeros_like(wk)
        _lib.check(_lib.lib().crfconv_edge_gauss_bwd(_p(fc), _p(eptr), _p(col), _p(wk), _p(g), _p(dw.contiguous().float()), _p(df), _p(dWk), N, Kk, H,
                                                    _lib.stream_ptr()), "edge_gauss_bwd")
        return df, dWk.view(ctx.wshape), None, None


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3):
    """torch_geometric.nn.knn_interpolate (models/point_conv.py:267-280): inverse-squared-distance weighted mean of the k nearest x."""
    with torch.no_grad():
        row, col = knn(pos_x, pos_y, k, batch_x, batch_y)
        diff = pos_x[col] - pos_y[row]
        w = 1.0 / torch.clamp((diff * diff).sum(-1), min=1e-16)
        eptr = torch.arange(pos_y.shape[0] + 1, dtype=torch.int64, device=x.device) * k
        den = SpMM.apply(w, torch.ones((pos_x.shape[0], 1), device=x.device), eptr, col.contiguous())
    return SpMM.apply(w, x, eptr, col.contiguous()) / den
