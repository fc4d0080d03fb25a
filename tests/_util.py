"""Shared comparison helpers for the parity tests."""
import numpy as np


def rel_err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor) — error relative to the tensor's scale.  `floor` guards quantities that are
    analytically zero (e.g. the gradient of a BatchNorm bias that feeds another BatchNorm), whose reference value is
    rounding noise."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-30))


def grad_floor(golden, tag, frac=1e-3):
    """frac × the largest parameter-gradient magnitude of the fixture `tag`."""
    return frac * max(float(np.abs(v).max()) for k, v in golden.items() if k.startswith(tag + ".gparam."))


def rel_l2(a, b, floor=0.0):
    """||a-b||_2 / max(||b||_2, floor·sqrt(n)) — robust to the handful of LeakyReLU kink flips that any two finite-precision
    implementations of a 10^7-activation layer disagree on (a flipped branch changes one gradient entry by O(1))."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor * np.sqrt(b.size), 1e-30))


def rel_err_trimmed(a, b, floor=0.0, outlier_frac=2e-5):
    """`rel_err` after setting aside the ceil(outlier_frac·n) largest deviations (at most 2 in 10^5 entries).
    Why: BatchNorm statistics are accumulated with floating-point atomics, so two runs of the SAME kernel differ in the last
    bit; an activation that sits within one ulp of a LeakyReLU kink then takes the other branch in one run out of a few dozen
    and changes ONE input-gradient entry by O(1).  Tests that use this also bound the L2 error over ALL entries."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.abs(a - b)
    k = int(np.ceil(outlier_frac * d.size)) if d.size >= 1000 else 0
    if k:
        d = np.partition(d, d.size - k - 1)[: d.size - k]
    return float(d.max() / max(np.abs(b).max(), floor, 1e-30))


def rel_err_rows_trimmed(a, b, floor=0.0, max_rows=None):
    """(max-norm error, L2 error, rows over 1e-3) of a [rows, C] gradient after setting aside the `max_rows` rows that deviate most
    (default: 0.4 % of the rows, at least 2).
    Why rows: a LeakyReLU unit whose pre-activation sits within rounding of 0 may take the other branch than the oracle (the batch
    statistics are summed in a different order); its backward factor changes by O(1) and with it the WHOLE row of the input
    gradient dX[i, :] = dH[i, :]·W of that point — C entries, more than the entry-wise trimming of `rel_err_trimmed` sets aside on
    wide layers with few rows.  Every other row must meet the tolerance."""
    a = np.asarray(a, dtype=np.float64)
    a = a.reshape(-1, a.shape[-1])
    b = np.asarray(b, dtype=np.float64).reshape(a.shape)
    d = np.abs(a - b).max(axis=1)
    if max_rows is None:
        max_rows = max(2, int(np.ceil(0.004 * len(d))))
    order = np.argsort(d)
    keep = order[: max(len(order) - max_rows, 1)]
    scale = max(np.abs(b).max(), floor, 1e-30)
    l2 = np.linalg.norm((a - b)[keep]) / max(np.linalg.norm(b[keep]), floor * np.sqrt(b[keep].size), 1e-30)
    return float(d[keep].max() / scale), float(l2), int((d > 1e-3 * scale).sum())
