"""Training criterion of the reference (trainval.py:66-70,100-104: ``torch.nn.CrossEntropyLoss(weight=…, ignore_index=…)`` on the
``[B·N, n_classes]`` logits) as one forward and one backward kernel (csrc/loss.cu).  Same semantics as ``F.cross_entropy`` for
class-index targets: 'mean' divides by the summed weights of the non-ignored rows, ignored rows get zero gradient."""
from __future__ import annotations

import torch

from . import _lib


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, ignore_index, mean):
        if not logits.is_cuda:
            raise RuntimeError("crfconv_b200.losses.cross_entropy runs on CUDA tensors only (no CPU fallback)")
        x = logits.detach().contiguous().float()
        t = target.detach().contiguous().to(torch.int64)
        w = weight.detach().contiguous().float() if weight is not None else None
        M, C = x.shape
        if t.numel() != M:
            raise RuntimeError(f"cross_entropy: {t.numel()} targets for {M} rows")
        sums = torch.zeros(2, dtype=torch.float64, device=x.device)
        rc = _lib.lib().crfconv_cross_entropy_fwd(x.data_ptr(), t.data_ptr(), w.data_ptr() if w is not None else None, M, C, int(ignore_index),
                                                  sums.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "cross_entropy_fwd")
        ctx.save_for_backward(x, t, w, sums)
        ctx.cfg = (int(ignore_index), bool(mean))
        ctx.mark_non_differentiable(sums)
        loss = (sums[0] / sums[1]) if mean else sums[0]
        return loss.to(torch.float32), sums

    @staticmethod
    def backward(ctx, gout, _gsums):
        x, t, w, sums = ctx.saved_tensors
        ignore_index, mean = ctx.cfg
        M, C = x.shape
        dx = torch.empty_like(x)
        g = gout.detach().contiguous().float()
        rc = _lib.lib().crfconv_cross_entropy_bwd(x.data_ptr(), t.data_ptr(), w.data_ptr() if w is not None else None, M, C, ignore_index,
                                                  sums.data_ptr(), g.data_ptr(), int(mean), dx.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "cross_entropy_bwd")
        return dx, None, None, None, None


def cross_entropy(logits, target, weight=None, ignore_index=-100, reduction="mean", return_normaliser=False):
    """``F.cross_entropy(logits [M, C], target [M] int64, weight, ignore_index, reduction)`` for reduction in {'mean', 'sum'}.
    return_normaliser: also return Σ of the class weights of the non-ignored rows (a float64 CUDA scalar; what 'mean' divides by)."""
    if reduction not in ("mean", "sum"):
        raise ValueError("cross_entropy: reduction must be 'mean' or 'sum'")
    if logits.dim() != 2:
        raise ValueError("cross_entropy: logits must be [M, C]")
    loss, sums = _CrossEntropy.apply(logits, target, weight, ignore_index, reduction == "mean")
    return (loss, sums[1]) if return_normaliser else loss
