"""Drop-in for the reference's ``models/point_conv_big.py``: ``PointConv`` (:8-58), ``ResNetBBlock`` (:61-88), ``Upsampling``
(:91-107) and ``PointConvResNet`` (:110-167, exported as ``models.PointConvBig``), with identical constructor / forward
signatures and ``state_dict`` keys, running on the sm_100a kernels:

  * neighbour gathers never materialise [B,N',K,F] tensors or the repeated int64 index (csrc/pointconv.cu);
  * every Linear → BatchNorm → LeakyReLU goes through the tensor-core kernels with BN statistics in the epilogue and the
    previous BN + activation applied on the fly (csrc/linear*.cu); torch.cat inputs are two-segment GEMMs;
  * ``leaky_relu(lin_out(x) + residual)`` is one fused BN + add + activation pass.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (kept for API parity with the reference module)

from . import ops
from .common import MLP, Base, bn_forward_state
from .continuous_crf_conv_big import ContinuousGaussianCRFConv as CRFConv


class _PointConvFunction(torch.autograd.Function):
    """out[b,i,:] = Σ_k w[b,i,k,:] ⊙ x[b, idx[b,i,k], :],  w = BN(W2·lrelu(BN(W1·(centre_i − support_j)))).
    BatchNorm statistics run over all B·N'·K edges, like the reference's weight_nn on [B, N'·K, 3]."""

    @staticmethod
    def forward(ctx, x, support, centres, idx, W1, g1, b1, W2, g2, b2, bn1, bn2, training, slope1, gbuf=None):
        if not x.is_cuda:
            raise RuntimeError("crfconv_b200 layers run on CUDA tensors only (no CPU fallback)")
        B, Ns, d = x.shape
        Nq, K = idx.shape[1], idx.shape[2]
        E = B * Nq * K
        x2 = ops.as2d(x)
        sup = support.detach().float().contiguous()
        cen = centres.detach().float().contiguous()
        nbr = idx.detach().contiguous().to(torch.int64)
        W1c, W2c = W1.detach().contiguous().float(), W2.detach().contiguous().float()
        rel = ops.relpos(sup, cen, nbr)                                                   # [E, 3]
        st1, fin = bn_forward_state(d, x.device, E, bn1, training or not bn1.track_running_stats)
        H1 = ops.linear_fwd(rel, W1c, stats=st1.stats); fin()
        st2, fin = bn_forward_state(d, x.device, E, bn2, training or not bn2.track_running_stats)
        H2 = ops.linear_fwd(H1, W2c, scale1=st1.scale, shift1=st1.shift, slope1=slope1, stats=st2.stats); fin()
        out = ops.pointconv_aggregate_fwd(x2, H2, st2, nbr, B, Ns, Nq, K)
        ctx.dims, ctx.st, ctx.slope1 = (B, Ns, Nq, K, d), (st1, st2), slope1
        ctx.gbuf = gbuf if (gbuf is not None and all(b is not None for b in gbuf)) else None     # direct accumulation targets (common.direct_grad_buffers)
        ctx.save_for_backward(x2, nbr, rel, H1, H2, W1c, W2c)
        return out.view(B, Nq, d)

    @staticmethod
    def backward(ctx, g):
        x2, nbr, rel, H1, H2, W1c, W2c = ctx.saved_tensors
        B, Ns, Nq, K, d = ctx.dims
        st1, st2 = ctx.st
        dev = g.device
        g2 = ops.as2d(g)
        if ctx.gbuf is not None:                                                          # the kernels accumulate into the bound .grad views
            dW1, dg1, db1, dW2, dg2, db2 = ctx.gbuf
        else:
            small = ops.Flat(W1c.numel() + W2c.numel() + 4 * d, torch.float32, dev)
            dW1, dW2 = small.take(*W1c.shape), small.take(*W2c.shape)
            dg1, db1, dg2, db2 = (small.take(d) for _ in range(4))
        dx = torch.zeros_like(x2) if ctx.needs_input_grad[0] else None
        dWgt = ops.pointconv_aggregate_bwd(x2, H2, st2, nbr, g2, dx, B, Ns, Nq, K)        # [E, d]
        ops.bn_backward_prepare(dWgt, H2, st2, 1.0, dg2, db2)
        dA1 = torch.empty_like(H1)
        ops.linear_bwd(dWgt, H2, st2, 1.0, H1, W2c, scale1=st1.scale, shift1=st1.shift, slope1=ctx.slope1, dX1=dA1, dW=dW2)
        ops.bn_backward_prepare(dA1, H1, st1, ctx.slope1, dg1, db1)
        ops.linear_bwd(dA1, H1, st1, ctx.slope1, rel, W1c, dW=dW1)                        # positions carry no gradient
        if ctx.gbuf is not None:
            return (dx.view(B, Ns, d) if dx is not None else None,) + (None,) * 14
        return (dx.view(B, Ns, d) if dx is not None else None, None, None, None, dW1, dg1, db1, dW2, dg2, db2, None, None, None, None, None)


FUSED_EDGE_MLP = True      # module attribute (the tests compare both paths); no environment switch
_REL_CACHE = {"key": None, "val": None}


def _relpos_moments_cached(sup, cen, nbr):
    """Relative positions and their moments depend on the geometry only: the two ResNet blocks of a level (conv1_1 / conv1_2: same
    positions, same neighbour table) share one pass.  Single-entry cache keyed on the storage and version counters of the inputs
    (an in-place update, e.g. the input copy before a CUDA-graph replay, bumps the version; replays re-run the captured launch)."""
    key = (sup.data_ptr(), cen.data_ptr(), nbr.data_ptr(), tuple(sup.shape), tuple(cen.shape), tuple(nbr.shape), sup._version, cen._version,
           nbr._version, torch.cuda.is_current_stream_capturing())
    if _REL_CACHE["key"] == key:
        return _REL_CACHE["val"]
    mom = torch.zeros(ops.STAT_SLOTS * 9, dtype=torch.float32, device=sup.device)
    rel = ops.pcf_relpos_moments(sup, cen, nbr, mom)
    _REL_CACHE["key"], _REL_CACHE["val"] = key, (rel, mom, sup, cen, nbr)     # the inputs are kept alive so that the addresses stay theirs
    return _REL_CACHE["val"]


class _PointConvFusedFunction(torch.autograd.Function):
    """The same layer with the 3→d→d edge MLP recomputed from the relative positions in every pass (csrc/pointconv_fused.cu, d = 8 / 16):
    no [E, d] tensor is written — BN1 statistics come from the moments of r, BN2 statistics and out = sc2 ⊙ Σ h2 ⊙ x_j + sh2 ⊙ Σ x_j
    from one pass, the backward needs two passes and derives dW1 / dW2 from per-matrix sums."""

    @staticmethod
    def forward(ctx, x, support, centres, idx, W1, g1, b1, W2, g2, b2, bn1, bn2, slope1, gbuf=None):
        if not x.is_cuda:
            raise RuntimeError("crfconv_b200 layers run on CUDA tensors only (no CPU fallback)")
        B, Ns, d = x.shape
        Nq, K = idx.shape[1], idx.shape[2]
        E = B * Nq * K
        dev = x.device
        x2 = ops.as2d(x)
        sup = support.detach().float().contiguous()
        cen = centres.detach().float().contiguous()
        nbr = idx.detach().contiguous().to(torch.int64)
        W1c, W2c = W1.detach().contiguous().float(), W2.detach().contiguous().float()
        S = ops.STAT_SLOTS
        nf, _ = ops.pcf_scratch_floats(d)
        z = ops.Flat(nf, torch.float32, dev)                    # one zero fill: BN1 statistics | BN2 statistics | activation sums
        stats1, stats2, asum = z.take(S * 2 * d), z.take(S * 2 * d), z.take(S * (d + d * d))
        rel, mom = _relpos_moments_cached(sup, cen, nbr)[:2]                               # [E, 3], Σr, Σrrᵀ
        st1, fin1 = bn_forward_state(d, dev, E, bn1, True, stats1)
        ops.pcf_stats1(mom, W1c, stats1); fin1()                                          # h1 = W1·r is linear in r
        st2, fin2 = bn_forward_state(d, dev, E, bn2, True, stats2)
        P, Q = ops.pcf_fwd(x2, rel, nbr, W1c, W2c, st1, slope1, stats2, asum, B, Ns, Nq, K); fin2()
        out = ops.pcf_out(P, Q, st2)
        ctx.dims, ctx.st, ctx.slope1 = (B, Ns, Nq, K, d), (st1, st2), slope1
        ctx.gbuf = gbuf if (gbuf is not None and all(b is not None for b in gbuf)) else None
        ctx.save_for_backward(x2, nbr, rel, W1c, W2c, mom, asum)
        return out.view(B, Nq, d)

    @staticmethod
    def backward(ctx, g):
        x2, nbr, rel, W1c, W2c, mom, asum = ctx.saved_tensors
        B, Ns, Nq, K, d = ctx.dims
        st1, st2 = ctx.st
        dev = g.device
        g2 = ops.as2d(g)
        if ctx.gbuf is not None:
            dW1, dg1, db1, dW2, dg2, db2 = ctx.gbuf
        else:
            small = ops.Flat(W1c.numel() + W2c.numel() + 4 * d, torch.float32, dev)
            dW1, dW2 = small.take(*W1c.shape), small.take(*W2c.shape)
            dg1, db1, dg2, db2 = (small.take(d) for _ in range(4))
        S = ops.STAT_SLOTS
        _, nb = ops.pcf_scratch_floats(d)
        z = ops.Flat(nb, torch.float32, dev, scratch=True)
        sums2, mdw, sums1, s1 = z.take(S * 2 * d), z.take(S * d * d), z.take(S * 2 * d), z.take(S * 3 * d)
        dx = torch.zeros_like(x2) if ctx.needs_input_grad[0] else None
        ops.pcf_bwd1(x2, rel, nbr, g2, W1c, W2c, st1, ctx.slope1, st2, dx, sums2, mdw, B, Ns, Nq, K)
        ops.bn_finalize_bwd(sums2, st2, dg2, db2)
        ops.pcf_bwd2(x2, rel, nbr, g2, W1c, W2c, st1, ctx.slope1, st2, sums1, s1, B, Ns, Nq, K)
        ops.bn_finalize_bwd(sums1, st1, dg1, db1)
        ops.pcf_param_grads(mom, asum, mdw, s1, W1c, W2c, st1, st2, dW1, dW2)
        if ctx.gbuf is not None:
            return (dx.view(B, Ns, d) if dx is not None else None,) + (None,) * 13
        return (dx.view(B, Ns, d) if dx is not None else None, None, None, None, dW1, dg1, db1, dW2, dg2, db2, None, None, None, None)


class _GatherMax(torch.autograd.Function):
    """out[b,i,:] = max_k x[b, idx[b,i,k], :]   (ResNetBBlock.max_pooling, point_conv_big.py:74-77)."""

    @staticmethod
    def forward(ctx, x, idx):
        B, Ns, C = x.shape
        Nq, K = idx.shape[1], idx.shape[2]
        out, arg = ops.gather_max_fwd(ops.as2d(x), idx.detach().contiguous().to(torch.int64), B, Ns, Nq, K)
        ctx.dims = (B, Ns, C)
        ctx.save_for_backward(arg)
        return out.view(B, Nq, C)

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        B, Ns, C = ctx.dims
        dx = torch.zeros((B * Ns, C), dtype=torch.float32, device=g.device)
        ops.gather_max_bwd(ops.as2d(g), arg, dx)
        return dx.view(B, Ns, C), None


class PointConv(nn.Module):
    """
    Re-implementation of original used depth-wise separable point conv in paper,
    the new version will be tested on large scale dataset
    """

    def __init__(self, d_model):
        super(PointConv, self).__init__()
        self.weight_nn = nn.Sequential(
            MLP(3, d_model, activation=nn.LeakyReLU(negative_slope=0.1)),
            MLP(d_model, d_model, activation=None)
        )

    @staticmethod
    def gather_neighbors(x, idx):
        """
        :param x: [B, N, F]
        :param idx: [B, N', K]
        :return: [B, N', K, F]   (kept for API parity; the fused forward below never calls it)
        """
        B = x.shape[0]
        return x[torch.arange(B, device=x.device).view(B, 1, 1), idx]

    def forward(self, x, pos, neighbor_idx):
        if torch.is_tensor(pos):
            support, centres = pos, pos
        else:
            support, centres = pos
        m1, m2 = self.weight_nn[0], self.weight_nn[1]
        if m1.slope is None or m2.slope != 1.0:
            raise RuntimeError("PointConv: weight_nn must keep the reference's (LeakyReLU, None) activations to be fused")
        b1, b2 = m1.bn.batch_norm, m2.bn.batch_norm
        from .common import direct_grad_buffers
        if (FUSED_EDGE_MLP and x.is_cuda and ops.pcf_supported(x.shape[-1]) and x.dtype == torch.float32 and
                (self.training or not (b1.track_running_stats or b2.track_running_stats))):
            return _PointConvFusedFunction.apply(x, support, centres, neighbor_idx, m1.lin.weight, b1.weight, b1.bias, m2.lin.weight,
                                                 b2.weight, b2.bias, b1, b2, m1.slope,
                                                 direct_grad_buffers(m1.lin.weight, b1.weight, b1.bias, m2.lin.weight, b2.weight, b2.bias))
        return _PointConvFunction.apply(x, support, centres, neighbor_idx, m1.lin.weight, b1.weight, b1.bias, m2.lin.weight, b2.weight,
                                        b2.bias, b1, b2, self.training, m1.slope,
                                        direct_grad_buffers(m1.lin.weight, b1.weight, b1.bias, m2.lin.weight, b2.weight, b2.bias))


class ResNetBBlock(nn.Module):
    negative_slope = 0.01      # F.leaky_relu default used by the reference's final activation (point_conv_big.py:88)

    def __init__(self, in_channels, out_channels):
        super(ResNetBBlock, self).__init__()
        hidden_channels = out_channels // 4
        self.lin_in = MLP(in_channels, hidden_channels, activation=nn.LeakyReLU(negative_slope=0.1))
        self.lin_out = MLP(hidden_channels, out_channels, activation=None)
        if in_channels != out_channels:
            self.shortcut = MLP(in_channels, out_channels, activation=None)
        else:
            self.shortcut = nn.Identity()

        self.point_conv = PointConv(hidden_channels)

    @staticmethod
    def max_pooling(x, idx):
        return _GatherMax.apply(x, idx)

    def forward(self, x, pos, neighbor_idx):
        residual = self.shortcut(x)
        if not torch.is_tensor(pos):
            residual = self.max_pooling(residual, neighbor_idx)

        x = self.lin_in(x)
        x = self.point_conv(x, pos, neighbor_idx)
        return self.lin_out(x, residual=residual, slope=self.negative_slope)      # F.leaky_relu(x + residual) (:88)


class Upsampling(nn.Module):
    def __init__(self, down_channels, up_channels, out_channels):
        super(Upsampling, self).__init__()
        self.lin = MLP(down_channels, up_channels, activation=nn.LeakyReLU(negative_slope=0.1))
        self.fusion = MLP(up_channels * 2, out_channels, activation=nn.LeakyReLU(negative_slope=0.1))

    @staticmethod
    def upsampling(x, idx):
        B = x.shape[0]
        return x[torch.arange(B, device=x.device).view(B, 1), idx[:, :, 0]]

    def forward(self, x_down, x_up, up_idx, neighbor_idx=None):
        x_down = self.lin(x_down, gather_idx=up_idx)                # gather fused into the GEMM's A-operand load
        return self.fusion(x_up, x2=x_down)                         # torch.cat([x_up, x_down]) never materialised


class PointConvResNet(Base):
    def __init__(self, in_channels, n_classes, use_crf=True, steps=1):
        super(PointConvResNet, self).__init__()
        layers = [32, 64, 128, 256, 512]
        self.C = n_classes

        self.conv1_1 = ResNetBBlock(in_channels, layers[0])
        self.conv1_2 = ResNetBBlock(layers[0], layers[0])

        self.conv2_1 = ResNetBBlock(layers[0], layers[1])
        self.conv2_2 = ResNetBBlock(layers[1], layers[1])

        self.conv3_1 = ResNetBBlock(layers[1], layers[2])
        self.conv3_2 = ResNetBBlock(layers[2], layers[2])

        self.conv4_1 = ResNetBBlock(layers[2], layers[3])
        self.conv4_2 = ResNetBBlock(layers[3], layers[3])

        self.conv5_1 = ResNetBBlock(layers[3], layers[4])
        self.conv5_2 = ResNetBBlock(layers[4], layers[4])

        self.deconv4 = CRFConv(layers[4], layers[3], layers[3], steps=steps) if use_crf else Upsampling(layers[4], layers[3], layers[3])
        self.deconv3 = CRFConv(layers[3], layers[2], layers[2], steps=steps) if use_crf else Upsampling(layers[3], layers[2], layers[2])
        self.deconv2 = CRFConv(layers[2], layers[1], layers[1], steps=steps) if use_crf else Upsampling(layers[2], layers[1], layers[1])
        self.deconv1 = CRFConv(layers[1], layers[0], layers[0], steps=steps) if use_crf else Upsampling(layers[1], layers[0], layers[0])

        self.classifier = nn.Sequential(
            MLP(layers[0], layers[0] * 4, activation=nn.LeakyReLU(negative_slope=0.1)),
            nn.Dropout(p=0.5),
            nn.Linear(layers[0] * 4, n_classes)
        )

    def forward(self, data):
        from . import common as _common
        ops.scratch_begin_step(data.x.device)               # one memset instead of ≈130 zero-fill kernels (ops.zeros_scratch)
        owner = _common.deferred_counters_begin()           # one num_batches_tracked bump instead of 46
        try:
            return self._forward(data)
        finally:
            if owner:
                _common.deferred_counters_end()

    def _forward(self, data):
        x, multiscale = data.x, data.multiscale

        x1 = self.conv1_1(x, multiscale[0].pos, multiscale[0].neighbor_idx)
        x1 = self.conv1_2(x1, multiscale[0].pos, multiscale[0].neighbor_idx)

        x2 = self.conv2_1(x1, (multiscale[0].pos, multiscale[1].pos), multiscale[0].sub_idx)
        x2 = self.conv2_2(x2, multiscale[1].pos, multiscale[1].neighbor_idx)

        x3 = self.conv3_1(x2, (multiscale[1].pos, multiscale[2].pos), multiscale[1].sub_idx)
        x3 = self.conv3_2(x3, multiscale[2].pos, multiscale[2].neighbor_idx)

        x4 = self.conv4_1(x3, (multiscale[2].pos, multiscale[3].pos), multiscale[2].sub_idx)
        x4 = self.conv4_2(x4, multiscale[3].pos, multiscale[3].neighbor_idx)

        x = self.conv5_1(x4, (multiscale[3].pos, multiscale[4].pos), multiscale[3].sub_idx)
        x = self.conv5_2(x, multiscale[4].pos, multiscale[4].neighbor_idx)

        x = self.deconv4(x, x4, multiscale[3].up_idx, multiscale[3].neighbor_idx)
        x = self.deconv3(x, x3, multiscale[2].up_idx, multiscale[2].neighbor_idx)
        x = self.deconv2(x, x2, multiscale[1].up_idx, multiscale[1].neighbor_idx)
        x = self.deconv1(x, x1, multiscale[0].up_idx, multiscale[0].neighbor_idx)

        x = self.classifier[0](x)
        x = self.classifier[1](x)                                   # nn.Dropout: RNG stream, plain torch op
        x = _classifier_head(self.classifier[2], x)

        return x.reshape(-1, self.C)


def _classifier_head(lin: nn.Linear, x):
    """Final nn.Linear(128, n_classes) (point_conv_big.py:139) through the same kernels (plain Linear, bias, no BN)."""
    from .common import _LinearBNAct
    from .common import direct_grad_buffers
    return _LinearBNAct.apply(x, None, None, None, lin.weight, lin.bias, None, None, None, False, 1.0,
                              direct_grad_buffers(lin.weight, lin.bias, None, None))
