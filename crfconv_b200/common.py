"""Drop-in for the hot-path part of the reference's ``models/common.py``: ``MLP`` (:26-40) and ``Base`` (:89-97), plus
``FastBatchNorm1d`` (torch_points3d.core.common_modules, imported at common.py:6), with identical constructor
signatures and ``state_dict`` keys (``lin.weight``, ``lin.bias``, ``bn.batch_norm.{weight,bias,running_mean,running_var,
num_batches_tracked}``), so reference checkpoints load unchanged.

The arithmetic runs in the sm_100a kernels of csrc/linear.cu (tensor-core Linear with BatchNorm statistics in the
epilogue, BN affine + LeakyReLU applied on the fly, BN-backward folded into dgrad/wgrad).  CUDA tensors only — there is
no CPU / eager fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class FastBatchNorm1d(nn.Module):
    """Parameter container with torch_points3d's layout; the statistics/normalisation are computed by the fused kernels
    of the owning ``MLP`` (a standalone call normalises through the same kernels)."""

    def __init__(self, num_features, momentum=0.1, **kwargs):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(num_features, momentum=momentum, **kwargs)

    def forward(self, x):
        if x.dim() not in (2, 3):
            raise ValueError("Non supported number of dimensions {}".format(x.dim()))
        bn = self.batch_norm
        eye = torch.eye(bn.num_features, device=x.device, dtype=torch.float32)
        return _LinearBNAct.apply(x, eye, None, bn.weight, bn.bias, bn, self.training or not bn.track_running_stats, 1.0)


def _slope_of(activation):
    """LeakyReLU slope equivalent of an activation module, or None if it is not of that family."""
    if activation is None:
        return 1.0
    if isinstance(activation, nn.LeakyReLU):
        return float(activation.negative_slope)
    if isinstance(activation, nn.ReLU):
        return 0.0
    return None


def bn_forward_state(C, device, count, bn_module, training, stats=None):
    """Allocates the per-call BN scratch; returns (state, finalize) where finalize() must run after the producing GEMM."""
    st = ops.BN(C, device, stats)

    def finalize():
        ops.bn_finalize_fwd(st, count, bn_module.weight, bn_module.bias, bn_module.eps,
                            bn_module.momentum if bn_module.momentum is not None else 0.1, training,
                            bn_module.running_mean, bn_module.running_var)
        if training and bn_module.track_running_stats and bn_module.num_batches_tracked is not None:
            bn_module.num_batches_tracked.add_(1)
    return st, finalize


class _LinearBNAct(torch.autograd.Function):
    """y = lrelu( BN( x·Wᵀ ) , slope )  or, without BN,  y = lrelu( x·Wᵀ + b , slope )."""

    @staticmethod
    def forward(ctx, x, W, bias, gamma, beta, bn_module, training, slope):
        if not x.is_cuda:
            raise RuntimeError("crfconv_b200 layers run on CUDA tensors only (no CPU fallback)")
        x2 = ops.as2d(x)
        Wc = W.detach().contiguous().float()
        M, Cout = x2.shape[0], Wc.shape[0]
        if bn_module is not None:
            st, fin = bn_forward_state(Cout, x.device, M, bn_module, training)
            H = ops.linear_fwd(x2, Wc, stats=st.stats if training else None)
            fin()
            Y = ops.bn_act_fwd(H, st, slope)
        else:
            st = None
            H = ops.linear_fwd(x2, Wc, bias=bias.detach().contiguous().float() if bias is not None else None)
            if slope == 1.0:
                Y = H
            else:
                st = ops.BN(Cout, x.device)            # identity affine, used only for the activation kernel
                st.scale.fill_(1.0); st.shift.zero_(); st.mean.zero_(); st.invstd.fill_(1.0)
                Y = ops.bn_act_fwd(H, st, slope)
        ctx.has_bn, ctx.slope, ctx.st = bn_module is not None, slope, st
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, Wc, H)
        return Y.view(*x.shape[:-1], Cout)

    @staticmethod
    def backward(ctx, gy):
        x2, Wc, H = ctx.saved_tensors
        g2 = ops.as2d(gy)
        Cout, Cin = Wc.shape
        need_x = ctx.needs_input_grad[0]
        dx = torch.empty_like(x2) if need_x else None
        dW = torch.zeros_like(Wc)
        dgamma = dbeta = dbias = None
        if ctx.has_bn:
            dgamma = torch.zeros(Cout, device=g2.device)
            dbeta = torch.zeros(Cout, device=g2.device)
            ops.bn_backward_prepare(g2, H, ctx.st, ctx.slope, dgamma, dbeta)
            ops.linear_bwd(g2, H, ctx.st, ctx.slope, x2, Wc, dX1=dx, dW=dW)
        else:
            dbias = torch.zeros(Cout, device=g2.device) if ctx.has_bias else None
            if ctx.slope != 1.0:       # activation without BN: fixed affine (scale 1), k1 = k2 = 0
                ctx.st.k1.zero_(); ctx.st.k2.zero_()
                ops.linear_bwd(g2, H, ctx.st, ctx.slope, x2, Wc, dX1=dx, dW=dW, dbias=dbias)
            else:
                ops.linear_bwd(g2, H, None, 1.0, x2, Wc, dX1=dx, dW=dW, dbias=dbias)
        nig = ctx.needs_input_grad
        return (dx.view(ctx.x_shape) if need_x else None), (dW if nig[1] else None), (dbias if nig[2] else None), \
            (dgamma if nig[3] else None), (dbeta if nig[4] else None), None, None, None


class MLP(nn.Module):
    def __init__(self, in_channels, out_channels, bn=True, activation=None):
        super().__init__()
        bias = False if bn else True
        self.lin = nn.Linear(in_channels, out_channels, bias=bias)
        self.bn = FastBatchNorm1d(out_channels) if bn else None
        self.activation = activation

    @property
    def slope(self):
        return _slope_of(self.activation)

    def forward(self, x, *args, **kwargs):
        slope = self.slope
        fused = slope if slope is not None else 1.0
        if self.bn is not None:
            bnm = self.bn.batch_norm
            y = _LinearBNAct.apply(x, self.lin.weight, None, bnm.weight, bnm.bias, bnm, self.training or not bnm.track_running_stats, fused)
        else:
            y = _LinearBNAct.apply(x, self.lin.weight, self.lin.bias, None, None, None, False, fused)
        if slope is None:
            y = self.activation(y)
        return y


class Base(nn.Module):
    def __init__(self):
        super().__init__()

    def save(self, filename):
        torch.save(self.state_dict(), filename)

    def load(self, filename):
        self.load_state_dict(torch.load(filename))
