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
        return _LinearBNAct.apply(x, None, None, None, eye, None, bn.weight, bn.bias, bn, self.training or not bn.track_running_stats, 1.0)


def _slope_of(activation):
    """LeakyReLU slope equivalent of an activation module, or None if it is not of that family."""
    if activation is None:
        return 1.0
    if isinstance(activation, nn.LeakyReLU):
        return float(activation.negative_slope)
    if isinstance(activation, nn.ReLU):
        return 0.0
    return None


# When a list, BatchNorm layers append their num_batches_tracked buffers here instead of launching one `+= 1` kernel each; whoever set
# it (PointConvResNet.forward) bumps them all with one torch._foreach_add_ (46 launches → 1 per step).
_NBT_DEFER = None


def deferred_counters_begin():
    global _NBT_DEFER
    if _NBT_DEFER is not None:
        return False
    _NBT_DEFER = []
    return True


def deferred_counters_end():
    global _NBT_DEFER
    lst, _NBT_DEFER = _NBT_DEFER, None
    if lst:
        torch._foreach_add_(lst, 1)


# Weight gradients feed nothing downstream in the backward pass.  At the deep levels of the network (≤ 2,560 points per cloud) the input-
# gradient and weight-gradient kernels of a layer are latency-bound launches on 30-300 CTAs each: the weight gradient runs on its own
# stream (a parallel branch of the captured graph) and is joined before the layer's backward returns.
WGRAD_SIDE_MAX_ROWS = 16384
_WGRAD_STREAMS = {}


def _wgrad_side_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _WGRAD_STREAMS:
        _WGRAD_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _WGRAD_STREAMS[key]


def bn_forward_state(C, device, count, bn_module, training, stats=None, defer_counters=None):
    """Allocates the per-call BN scratch; returns (state, finalize) where finalize() must run after the producing GEMM.
    `defer_counters` (a list): collect the num_batches_tracked buffers instead of bumping each with its own kernel; the caller
    bumps them all with one torch._foreach_add_."""
    st = ops.BN(C, device, stats)
    if training and bn_module.track_running_stats and bn_module.momentum is None:
        # cumulative moving average: the factor 1 / num_batches_tracked lives in a device counter; reading it would synchronise every step
        raise RuntimeError("crfconv_b200: BatchNorm momentum=None (cumulative average) is not supported; use a float momentum")

    def finalize():
        ops.bn_finalize_fwd(st, count, bn_module.weight, bn_module.bias, bn_module.eps,
                            bn_module.momentum if bn_module.momentum is not None else 0.1, training,
                            bn_module.running_mean, bn_module.running_var)
        if training and bn_module.track_running_stats and bn_module.num_batches_tracked is not None:
            if _NBT_DEFER is not None:
                _NBT_DEFER.append(bn_module.num_batches_tracked)
            elif defer_counters is not None:
                defer_counters.append(bn_module.num_batches_tracked)
            else:
                bn_module.num_batches_tracked.add_(1)
    return st, finalize


class _LinearBNAct(torch.autograd.Function):
    """y = lrelu( BN( [x1[idx] | x2]·Wᵀ ) (+ R) , slope )   or, without BN,   y = lrelu( [x1 | x2]·Wᵀ + b , slope ).

    x1 [B,N1,C1] (rows gathered through idx [B,N,1] when given), x2 [B,N,C2] or None (the concat is never materialised),
    R [B,N,Cout] or None (residual added before the activation, point_conv_big.py:88)."""

    @staticmethod
    def forward(ctx, x1, x2, idx, R, W, bias, gamma, beta, bn_module, training, slope, gbuf=None):
        if not x1.is_cuda:
            raise RuntimeError("crfconv_b200 layers run on CUDA tensors only (no CPU fallback)")
        a1 = ops.as2d(x1)
        a2 = ops.as2d(x2) if x2 is not None else None
        r2 = ops.as2d(R) if R is not None else None
        Wc = W.detach().contiguous().float()
        Cout = Wc.shape[0]
        gidx, rows_dst, rows_src = None, 0, 0
        lead = x1.shape[:-1]
        if idx is not None:
            gidx = idx.detach().reshape(idx.shape[0], -1).contiguous().to(torch.int64)
            rows_dst, rows_src = gidx.shape[1], x1.shape[1]
            lead = (x1.shape[0], rows_dst)
        M = gidx.numel() if gidx is not None else a1.shape[0]
        if bn_module is not None:
            st, fin = bn_forward_state(Cout, x1.device, M, bn_module, training)
            H = ops.linear_fwd(a1, Wc, idx1=gidx, rows_dst=rows_dst, rows_src=rows_src, X2=a2, stats=st.stats if training else None, M=M)
            fin()
            Y = ops.bn_act_fwd(H, st, slope, R=r2)
        else:
            st = None
            H = ops.linear_fwd(a1, Wc, idx1=gidx, rows_dst=rows_dst, rows_src=rows_src, X2=a2,
                               bias=bias.detach().contiguous().float() if bias is not None else None, M=M)
            if slope == 1.0 and r2 is None:
                Y = H
            else:
                st = ops.BN(Cout, x1.device)           # identity affine, used only for the (residual +) activation kernel
                st.scale.fill_(1.0); st.shift.zero_(); st.mean.zero_(); st.invstd.fill_(1.0)
                st.count, st.training = M, False
                Y = ops.bn_act_fwd(H, st, slope, R=r2)
        ctx.has_bn, ctx.slope, ctx.st = bn_module is not None, slope, st
        ctx.gbuf = gbuf        # (dW, dbias, dgamma, dbeta) accumulation targets owned by the caller (see direct_grad_buffers), or None
        ctx.has_bias, ctx.has_R = bias is not None, R is not None
        ctx.shapes = (x1.shape, x2.shape if x2 is not None else None, R.shape if R is not None else None)
        ctx.gather = (rows_dst, rows_src)
        ctx.save_for_backward(a1, a2, gidx, Wc, H, Y if (R is not None or st is not None and not ctx.has_bn) else None)
        return Y.view(*lead, Cout)

    @staticmethod
    def backward(ctx, gy):
        a1, a2, gidx, Wc, H, Y = ctx.saved_tensors
        g2 = ops.as2d(gy)
        Cout = Wc.shape[0]
        dev = g2.device
        nig = ctx.needs_input_grad
        rows_dst, rows_src = ctx.gather
        M = g2.shape[0]
        dR = None
        slope = ctx.slope
        if ctx.has_R:                          # out = lrelu(V + R): dS feeds both the residual and the BN branch (slope 1)
            g2 = ops.lrelu_bwd(g2, Y, slope)
            dR, slope = g2, 1.0
        dX1 = torch.empty((M, a1.shape[1]), dtype=torch.float32, device=dev) if nig[0] else None
        dX2 = torch.empty_like(a2) if (a2 is not None and nig[1]) else None
        gb = ctx.gbuf or (None, None, None, None)
        direct = [b is not None for b in gb]                 # kernels accumulate (+=) straight into the caller's gradient buffers
        dW = gb[0] if direct[0] else torch.zeros_like(Wc)
        dgamma = dbeta = dbias = None
        st = ctx.st
        if ctx.has_bn:
            dgamma = gb[2] if direct[2] else torch.zeros(Cout, device=dev)
            dbeta = gb[3] if direct[3] else torch.zeros(Cout, device=dev)
            ops.bn_backward_prepare(g2, H, st, slope, dgamma, dbeta)
        else:
            dbias = (gb[1] if direct[1] else torch.zeros(Cout, device=dev)) if ctx.has_bias else None
            if st is not None and slope != 1.0:
                st.k1.zero_(); st.k2.zero_()   # activation without BN: fixed affine (scale 1)
            else:
                st = None
        join = None
        if (dX1 is not None or dX2 is not None) and M <= WGRAD_SIDE_MAX_ROWS:
            main, ws = torch.cuda.current_stream(dev), _wgrad_side_stream(dev)
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            ws.wait_event(fork)
            with torch.cuda.stream(ws):                      # dW (and dbias): a parallel branch
                ops.linear_bwd(g2, H, st, slope, a1, Wc, idx1=gidx, rows_dst=rows_dst, rows_src=rows_src, X2=a2, dW=dW, dbias=dbias)
                join.record(ws)
            ops.linear_bwd(g2, H, st, slope, a1, Wc, idx1=gidx, rows_dst=rows_dst, rows_src=rows_src, X2=a2, dX1=dX1, dX2=dX2)
        else:
            ops.linear_bwd(g2, H, st, slope, a1, Wc, idx1=gidx, rows_dst=rows_dst, rows_src=rows_src, X2=a2, dX1=dX1, dX2=dX2, dW=dW,
                           dbias=dbias)
        s1, s2, sR = ctx.shapes
        if dX1 is not None and gidx is not None:           # gradient wrt gathered rows → scatter onto the source rows
            full = torch.zeros((s1[0] * s1[1], s1[2]), dtype=torch.float32, device=dev)
            ops.scatter_add_rows(dX1, gidx, full, s1[0], rows_dst, rows_src)
            dX1 = full
        if join is not None:
            torch.cuda.current_stream(dev).wait_event(join)
        return (dX1.view(s1) if dX1 is not None else None, dX2.view(s2) if dX2 is not None else None, None,
                dR.view(sR) if (dR is not None and nig[3]) else None, dW if (nig[4] and not direct[0]) else None,
                dbias if (nig[5] and not direct[1]) else None, dgamma if (nig[6] and not direct[2]) else None,
                dbeta if (nig[7] and not direct[3]) else None, None, None, None, None)


def direct_grad_buffers(*params):
    """Gradient buffers that the backward kernels may accumulate into DIRECTLY, one per parameter (None where not applicable).

    A parameter opts in through ``p._crf_direct_grad = True`` — set by ``FlatGradients(module, direct=True)`` on parameters whose
    ``.grad`` is a pre-bound view of the flat gradient buffer.  The kernels already accumulate (dW += …, dγ += …), so handing them the
    bound views removes, per parameter and step, a zero-fill of a temporary and autograd's AccumulateGrad add kernel (≈450 tiny
    launches per PointConvResNet step).  The autograd Function then reports no gradient for that parameter; code that needs
    ``torch.autograd.grad`` or parameter hooks must not opt in."""
    out = []
    for p in params:
        g = getattr(p, "grad", None) if p is not None else None
        ok = (p is not None and getattr(p, "_crf_direct_grad", False) and g is not None and g.is_cuda and g.dtype == torch.float32
              and g.is_contiguous() and p.requires_grad and torch.is_grad_enabled())
        out.append(g if ok else None)
    return tuple(out) if any(b is not None for b in out) else None


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

    def forward(self, x, *args, x2=None, gather_idx=None, residual=None, slope=None, **kwargs):
        """x2: second input segment (the reference's torch.cat([x, x2], -1) without the copy); gather_idx: [B,N,1] row
        gather applied to x (Upsampling.upsampling); residual / slope: fused `leaky_relu(mlp(x) + residual, slope)`."""
        own = self.slope
        fused = own if own is not None else 1.0
        if residual is not None:
            assert own == 1.0 and slope is not None, "a fused residual needs an activation-free MLP and an explicit slope"
            fused = slope
        if self.bn is not None:
            bnm = self.bn.batch_norm
            y = _LinearBNAct.apply(x, x2, gather_idx, residual, self.lin.weight, None, bnm.weight, bnm.bias, bnm,
                                   self.training or not bnm.track_running_stats, fused,
                                   direct_grad_buffers(self.lin.weight, None, bnm.weight, bnm.bias))
        else:
            y = _LinearBNAct.apply(x, x2, gather_idx, residual, self.lin.weight, self.lin.bias, None, None, None, False, fused,
                                   direct_grad_buffers(self.lin.weight, self.lin.bias, None, None))
        if own is None:
            y = self.activation(y)
        return y


class Base(nn.Module):
    def __init__(self):
        super().__init__()

    def save(self, filename):
        torch.save(self.state_dict(), filename)

    def load(self, filename):
        self.load_state_dict(torch.load(filename))
