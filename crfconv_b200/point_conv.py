"""Edge-list form of the depthwise-separable point convolution — drop-in for ``DepthwiseSeparablePointConv`` of the reference's
PyG family (models/point_conv.py:12-66): same constructor, ``forward(x, pos | (pos_src, pos_dst), edge_index)``, sub-module names
and ``state_dict`` keys (``mlp1.{0,1,3,4}``, ``mlp2.{0,1}``, ``mlp3.{0,1}``, ``mlp4.{0,1}``).

    out_i = leaky_relu( mlp3( Σ_{j→i} mlp1(pos_i − pos_j) ⊙ mlp2(x)_j ) + residual_i )                      (:43-66)

``edge_index`` follows PyG's default flow: row 0 = source j, row 1 = target i (:50-52).  With a single ``pos`` tensor the reference
removes self loops and adds one per node (:46-48) and the residual is ``x`` itself; with a ``(pos_src, pos_dst)`` pair the graph is
bipartite, no self loops are touched, and the residual is the max over each target's sources (:50-53).  ``mlp4`` (:37-41) projects
the residual when the channel counts differ; its ``nn.Linear`` has a bias, which the BatchNorm that follows cancels in the output
(its gradient is identically zero and is returned as zeros) and which only shifts the running mean (``_lin_bias_bn``).

Runs on the dense PointConv kernels (csrc/pointconv.cu + the Linear/BN chains), which take a [N, K] neighbour table and compute
``mlp1``'s BatchNorm statistics over all N·K edges — so the graph must be REGULAR after the self-loop step (``knn_graph``);
ragged graphs are rejected loudly (padding edges would enter those statistics).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .common import _LinearBNAct
from .point_conv_big import _GatherMax, _PointConvFunction


def _regular_table(src, dst, num_dst):
    """[1, num_dst, K] table of sources per target (edges grouped stably by target); raises if in-degrees differ."""
    E = dst.numel()
    if num_dst <= 0 or E == 0 or E % num_dst != 0:
        raise NotImplementedError("crfconv_b200: the edge-list point convolution needs a regular graph (same in-degree for every target)")
    K = E // num_dst
    expect = torch.arange(num_dst, device=dst.device).repeat_interleave(K)
    if not torch.equal(dst, expect):
        order = torch.sort(dst, stable=True).indices
        src, dst = src[order], dst[order]
        if not torch.equal(dst, expect):
            raise NotImplementedError("crfconv_b200: the edge-list point convolution needs a regular graph (same in-degree for every target)")
    return src.view(1, num_dst, K).contiguous()


def _lin_bn(x, lin, bn, slope, training, residual=None):
    y = _LinearBNAct.apply(x[None], None, None, residual[None] if residual is not None else None, lin.weight, None, bn.weight, bn.bias, bn,
                           training or not bn.track_running_stats, slope)
    return y[0]


def _lin_bias_bn(x, lin, bn, training):
    """nn.Linear WITH bias followed by BatchNorm (mlp4, :37-41).  In training mode the bias cancels in the normalised output and only
    shifts the batch mean, so the fused bias-free kernel is used and the running mean is corrected by momentum·bias; in eval mode
    BN(Wx + b) with running mean rm equals the bias-free BN with rm − b."""
    use_batch = training or not bn.track_running_stats
    b = lin.bias.detach()
    if use_batch:
        y = _lin_bn(x, lin, bn, 1.0, training)
        if training and bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.add_(m * b)
        if torch.is_grad_enabled() and lin.bias.requires_grad:
            y = y + 0.0 * lin.bias.sum()      # the bias gradient is identically zero: hand autograd zeros (not None) like the reference, so
        return y                              # that an optimizer's weight decay treats the parameter the same way
    with torch.no_grad():
        bn.running_mean.sub_(b)
    try:
        return _lin_bn(x, lin, bn, 1.0, training)
    finally:
        with torch.no_grad():
            bn.running_mean.add_(b)


class DepthwiseSeparablePointConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super(DepthwiseSeparablePointConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.hidden_channels = out_channels // 4
        h = self.hidden_channels
        self.mlp1 = nn.Sequential(nn.Linear(3, h, bias=False), nn.BatchNorm1d(h), nn.LeakyReLU(inplace=True),
                                  nn.Linear(h, h, bias=False), nn.BatchNorm1d(h))
        self.mlp2 = nn.Sequential(nn.Linear(self.in_channels, h, bias=False), nn.BatchNorm1d(h), nn.LeakyReLU(inplace=True))
        self.mlp3 = nn.Sequential(nn.Linear(h, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels))
        if self.in_channels != self.out_channels:
            self.mlp4 = nn.Sequential(nn.Linear(self.in_channels, self.out_channels), nn.BatchNorm1d(self.out_channels))

    def forward(self, x, pos, edge_index):
        src, dst = edge_index[0].to(torch.int64), edge_index[1].to(torch.int64)
        tr = self.training
        if torch.is_tensor(pos):                       # symmetric graph: remove self loops, add one per node (:46-48)
            N = pos.size(0)
            keep = src != dst
            loops = torch.arange(N, device=src.device)
            src, dst = torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])
            support, centres = pos, pos
            table = _regular_table(src, dst, N)
            residual = x
        else:                                          # bipartite graph (:50-53)
            support, centres = pos
            table = _regular_table(src, dst, centres.size(0))
            residual = _GatherMax.apply(x[None], table)[0]
        if self.in_channels != self.out_channels:
            residual = _lin_bias_bn(residual, self.mlp4[0], self.mlp4[1], tr)
        h = _lin_bn(x, self.mlp2[0], self.mlp2[1], self.mlp2[2].negative_slope, tr)
        m = self.mlp1
        agg = _PointConvFunction.apply(h[None], support[None], centres[None], table, m[0].weight, m[1].weight, m[1].bias, m[3].weight,
                                       m[4].weight, m[4].bias, m[1], m[4], tr, m[2].negative_slope)[0]
        return _lin_bn(agg, self.mlp3[0], self.mlp3[1], 0.01, tr, residual=residual)       # F.leaky_relu(x + residual), default slope (:58)
