"""Edge-list form of the continuous-CRF layer — drop-in for ``ContinuousGaussianCRFConv`` of the reference's PyG family
(models/continuous_crf_conv.py:72-133): same constructor, same ``forward(x, y, pos, edge_index)``, same sub-module names and
``state_dict`` keys (``unary_net.{0,1}``, ``pairwise_net.{0,1}``, ``mlp.{0,1}``, ``fusion_net.{0,1}``, ``c``).

    s_ij = softmax over the edges of target i of  −‖e_i − e_j‖²        (e = pairwise_net(y); PyG softmax, :118-119)
    x^t  = ( z + (Σ_j s_ij x^{t−1}_j)·C )·(I + C)⁻¹ ,  z = unary_net(x) ,  C = cᵀc                       (:121-128)
    out  = fusion_net([ mlp(x^T) | y ])                                                                 (:130-131)

It runs on the same sm_100a mean-field kernels as the dense layer (csrc/crf.cu).  Those kernels take a [N, K] neighbour table:
a REGULAR edge list (every node the target of the same number of edges — what ``knn_graph(pos, k)`` produces, point_conv.py:267-280)
maps onto it directly; a ragged one (``radius_graph`` with ``max_num_neighbors``) is padded to its largest in-degree with an extra
node whose softmax weight is exactly zero (see ``dense_neighbours``).

``GuideGaussianCRFConv`` (:9-69) — the variant that builds its own ``radius_graph`` and runs the mean field at full width (out_channels
up to 256) — is provided on the ragged edge-list kernels of csrc/graph.cu (edge softmax + weighted aggregation, any channel count)
and the grid radius search of csrc/knn.cu (``graph_ops.radius_graph``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .common import _LinearBNAct


_FAR = 1.0e18       # embedding of the padding node: (1e18)² · F stays finite in fp32 and exp(−that) is exactly 0


def dense_neighbours(edge_index, num_nodes):
    """edge_index [2, E] (row 0 = target i, row 1 = source j, as the reference unpacks it at :114) → ([1, N', K+1] int64 table, padded).

    Column 0 of the table is the node itself (ignored by the kernels, like the self-match of the dense pipeline), columns 1..K its
    sources in the order given.  Regular graphs (every node the target of exactly K edges, e.g. ``knn_graph``) map one to one
    (N' = N, padded = False).  Ragged graphs (``radius_graph``) are padded to the largest in-degree with the index N of an extra
    PADDING NODE (N' = N + 1, padded = True) that the caller appends with a far-away embedding and a zero state: its softmax
    weight is exactly 0 in the forward and in the backward pass, so the result equals the ragged sum."""
    i, j = edge_index[0].to(torch.int64), edge_index[1].to(torch.int64)
    E = i.numel()
    if num_nodes <= 0:
        raise ValueError("crfconv_b200: edge-list CRF layer needs at least one node")
    dev = i.device
    K = E // num_nodes if E % num_nodes == 0 else -1
    if K > 0:
        expect = torch.arange(num_nodes, device=dev).repeat_interleave(K)
        regular = torch.equal(i, expect)
    else:
        regular = False
    if not regular:
        order = torch.sort(i, stable=True).indices            # group by target, keeping the given order inside a group
        i, j = i[order], j[order]
        regular = K > 0 and torch.equal(i, torch.arange(num_nodes, device=dev).repeat_interleave(K))
    if regular:
        nbr = torch.empty((1, num_nodes, K + 1), dtype=torch.int64, device=dev)
        nbr[0, :, 0] = torch.arange(num_nodes, device=dev)
        nbr[0, :, 1:] = j.view(num_nodes, K)
        return nbr, False
    counts = torch.bincount(i, minlength=num_nodes)
    kmax = max(int(counts.max()), 1) if E > 0 else 1
    start = torch.cumsum(counts, 0) - counts                   # first edge of every group (edges are grouped by target now)
    slot = torch.arange(E, device=dev) - start[i]
    nbr = torch.full((1, num_nodes + 1, kmax + 1), num_nodes, dtype=torch.int64, device=dev)
    nbr[0, :, 0] = torch.arange(num_nodes + 1, device=dev)
    nbr[0, i, 1 + slot] = j
    return nbr, True


class _MeanField(torch.autograd.Function):
    """x^T from (z, e, neighbour table, c): the loop of continuous_crf_conv.py:121-128 / continuous_crf_conv_big.py:68-72."""

    @staticmethod
    def forward(ctx, z, e, nbr, c, steps):
        if not z.is_cuda:
            raise RuntimeError("crfconv_b200 layers run on CUDA tensors only (no CPU fallback)")
        z2, e2 = ops.as2d(z), ops.as2d(e)
        _, N, K = nbr.shape
        F = z2.shape[1]
        cc = c.detach().contiguous().float()
        Cm, Minv = ops.crf_compat_fwd(cc)
        ones = torch.ones(F, dtype=torch.float32, device=z.device)
        xs = [z2]
        for _ in range(steps):
            xs.append(ops.crf_step_fwd(e2, ones, z2, xs[-1], nbr, Cm, Minv, 1, N, K))
        ctx.dims = (N, K, F, steps)
        ctx.save_for_backward(e2, nbr, cc, Cm, Minv, ones, *xs)
        return xs[-1].clone() if steps == 0 else xs[-1]

    @staticmethod
    def backward(ctx, g):
        e2, nbr, cc, Cm, Minv, ones, *xs = ctx.saved_tensors
        N, K, F, steps = ctx.dims
        dev = g.device
        z2 = xs[0]
        g = g.contiguous().float()
        if steps == 0:
            return g, torch.zeros_like(e2), None, torch.zeros_like(cc), None
        Gy = torch.zeros((N, F), dtype=torch.float32, device=dev)
        GC, GM = torch.zeros((F, F), dtype=torch.float32, device=dev), torch.zeros((F, F), dtype=torch.float32, device=dev)
        Gz, m_out, v_out, h_out = (torch.empty((N, F), dtype=torch.float32, device=dev) for _ in range(4))
        for t in range(steps, 0, -1):
            gprev = torch.zeros((N, F), dtype=torch.float32, device=dev)
            ops.crf_step_bwd(e2, ones, z2, xs[t - 1], nbr, Cm, Minv, g, Gz, gprev, Gy, m_out, v_out, h_out, t != steps, 1, N, K)
            ops.linear_bwd(m_out, None, None, 1.0, h_out, GC, dW=GC)          # GC += mᵀ·h
            ops.linear_bwd(v_out, None, None, 1.0, g, GM, dW=GM)              # GM += vᵀ·g
            g = gprev
        Gc = torch.zeros((F, F), dtype=torch.float32, device=dev)
        ops.crf_compat_bwd(cc, Minv, GC, GM, Gc)
        return Gz + g, Gy, None, Gc, None                                      # dL/dz = Σ_t h^t + g^0


def _lin_bn(x, lin: nn.Linear, bn: nn.BatchNorm1d, slope, training, x2=None):
    """nn.Linear(bias=False) + nn.BatchNorm1d (+ LeakyReLU(slope)) on [N, C] rows through the fused Linear+BN kernels."""
    y = _LinearBNAct.apply(x[None], x2[None] if x2 is not None else None, None, None, lin.weight, None, bn.weight, bn.bias, bn,
                           training or not bn.track_running_stats, slope)
    return y[0]


class ContinuousGaussianCRFConv(nn.Module):
    def __init__(self, unary_channels, pairwise_channels, hidden_channels=None, out_channels=None, steps=1):
        super(ContinuousGaussianCRFConv, self).__init__()
        self.unary_channels = unary_channels
        self.pairwise_channels = pairwise_channels
        self.out_channels = out_channels if out_channels is not None else pairwise_channels
        self.hidden_channels = hidden_channels if hidden_channels is not None else self.out_channels // 4
        self.steps = steps
        self.unary_net = nn.Sequential(nn.Linear(self.unary_channels, self.hidden_channels, bias=False), nn.BatchNorm1d(self.hidden_channels))
        self.pairwise_net = nn.Sequential(nn.Linear(self.pairwise_channels, self.hidden_channels, bias=False), nn.BatchNorm1d(self.hidden_channels))
        self.mlp = nn.Sequential(nn.Linear(self.hidden_channels, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels),
                                 nn.LeakyReLU(inplace=True))
        self.fusion_net = nn.Sequential(nn.Linear(self.out_channels * 2, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels),
                                        nn.LeakyReLU(inplace=True))
        self.c = nn.Parameter(torch.Tensor(self.hidden_channels, self.hidden_channels))
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.eye_(self.c)

    def forward(self, x, y, pos, edge_index):
        if self.pairwise_channels != self.out_channels:
            raise RuntimeError("ContinuousGaussianCRFConv: fusion_net needs pairwise_channels == out_channels (continuous_crf_conv.py:102,131)")
        if self.hidden_channels not in (4, 8, 16, 32, 64):
            raise RuntimeError("ContinuousGaussianCRFConv: hidden channels must be 4, 8, 16, 32 or 64")
        N = pos.shape[0]
        nbr, padded = dense_neighbours(edge_index, N)
        tr = self.training
        z = _lin_bn(x, self.unary_net[0], self.unary_net[1], 1.0, tr)
        e = _lin_bn(y, self.pairwise_net[0], self.pairwise_net[1], 1.0, tr)
        if padded:                                             # the padding node: far-away embedding, zero state (after the BatchNorms)
            zp = torch.cat([z, z.new_zeros(1, z.shape[1])], dim=0)
            ep = torch.cat([e, e.new_full((1, e.shape[1]), _FAR)], dim=0)
            xm = _MeanField.apply(zp, ep, nbr, self.c, self.steps)[:N]
        else:
            xm = _MeanField.apply(z, e, nbr, self.c, self.steps)
        o = _lin_bn(xm, self.mlp[0], self.mlp[1], self.mlp[2].negative_slope, tr)
        return _lin_bn(o, self.fusion_net[0], self.fusion_net[1], self.fusion_net[2].negative_slope, tr, x2=y)


class GuideGaussianCRFConv(nn.Module):
    """Drop-in for models/continuous_crf_conv.py:9-69: ``forward(x, y, pos, batch)`` builds the radius graph itself
    (``radius_graph(pos, r, batch, max_num_neighbors=kernel_size)``, :52), embeds x / y with one Linear + BatchNorm each (y also
    through LeakyReLU(0.01)), runs ``steps`` mean-field updates at ``out_channels`` width and returns ``leaky_relu(x)``.
    Sub-module names and ``state_dict`` keys are the reference's (``unary.{0,1}``, ``pairwise.{0,1}``, ``c``)."""

    def __init__(self, in_n_channels, in_e_channels, out_channels=None, radius=0.1, kernel_size=32, steps=1):
        super(GuideGaussianCRFConv, self).__init__()
        self.in_n_channels = in_n_channels
        self.in_e_channels = in_e_channels
        self.out_channels = out_channels if out_channels is not None else in_e_channels
        self.radius = radius
        self.kernel_size = kernel_size
        self.steps = steps
        self.unary = nn.Sequential(nn.Linear(self.in_n_channels, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels))
        self.pairwise = nn.Sequential(nn.Linear(self.in_e_channels, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels),
                                      nn.LeakyReLU(inplace=True))
        self.c = nn.Parameter(torch.Tensor(self.out_channels, self.out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.eye_(self.c)

    def _lin_bn(self, seq, t, slope):
        lin, bn = seq[0], seq[1]
        return _LinearBNAct.apply(t, None, None, None, lin.weight, None, bn.weight, bn.bias, bn, self.training or not bn.track_running_stats, slope)

    def forward(self, x, y, pos, batch=None, edge_index=None):
        """edge_index (optional, [source j, target i]) replaces the internally built radius graph — used by the parity tests."""
        from . import graph_ops
        if not x.is_cuda:
            raise RuntimeError("crfconv_b200 layers run on CUDA tensors only (no CPU fallback)")
        N = pos.shape[0]
        if edge_index is None:
            edge_index = graph_ops.radius_graph(pos, self.radius, batch, loop=False, max_num_neighbors=self.kernel_size)
        col, row = edge_index[0], edge_index[1]                                  # (:52) col = source j, row = target i
        eptr, colg, _ = graph_ops.csr_by_target(row.to(torch.int64), col.to(torch.int64), N)
        x = self._lin_bn(self.unary, x, 1.0)                                     # (:53)
        y = self._lin_bn(self.pairwise, y, float(self.pairwise[2].negative_slope))   # (:54)
        s = graph_ops.EdgeSoftmax.apply(y, eptr, colg)                           # (:55-56)
        z = x
        eye = torch.eye(self.out_channels, dtype=torch.float32, device=x.device)
        C = torch.mm(self.c.t(), self.c)                                         # (:60) F x F algebra stays in torch
        Minv = torch.linalg.inv(eye + C)
        for t in range(self.steps):                                              # (:62-66)
            m = graph_ops.SpMM.apply(s, x, eptr, colg)
            x = _LinearBNAct.apply(m, None, None, z, C.t(), None, None, None, None, False, 1.0)              # z + m·C
            last = t == self.steps - 1
            x = _LinearBNAct.apply(x, None, None, None, Minv.t(), None, None, None, None, False, 0.01 if last else 1.0)   # ·(I+C)^-1 (+ final leaky_relu, :69)
        if self.steps == 0:
            x = torch.nn.functional.leaky_relu(x)
        return x
