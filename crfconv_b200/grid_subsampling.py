"""Drop-in for the reference's ``grid_subsampling`` CPython module (utils/cpp_wrappers/cpp_subsampling/wrapper.cpp:58-286),
re-exported there as ``utils.cpp_subsampling``; backed by the sm_100a radix-sort / segmented-sum kernels.

    compute(points, *, features=None, classes=None, sampleDl=0.1, method='barycenters', verbose=0)
        -> points | (points, features) | (points, classes) | (points, features, classes)       (wrapper.cpp:269-276)

Argument handling follows the wrapper: ``points`` is the only positional argument ("O|$OOfsi", wrapper.cpp:76);
inputs are converted to C-contiguous float32 / float32 / int32 (:104-106); shape errors raise ``RuntimeError`` with the
wrapper's messages (:109-190); ``method`` is validated and then ignored, exactly like the reference (:83-90 —
barycentres are always returned); 1-D classes come back as ``[M,1]`` (:170-172,241-243).

Extension (keyword-only, default reproduces the reference): ``order='reference'`` emits rows in the reference's
libstdc++ ``unordered_map`` iteration order; ``order='key'`` skips the host replay and emits ascending voxel-key order
(same set of rows, bit for bit).  CUDA tensors in ⇒ CUDA tensors out (no PCIe); numpy in ⇒ numpy out.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .nearest_neighbors import _is_cuda_tensor, _workspace


def _order_flag(order):
    if order not in ("reference", "key"):
        raise RuntimeError('Error parsing order. Valid orders are "reference" and "key"')
    return 1 if order == "reference" else 0


def compute(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0, order="reference"):
    if method not in ("barycenters", "voxelcenters"):
        raise RuntimeError('Error parsing method. Valid method names are "barycenters" and "voxelcenters" ')
    if _is_cuda_tensor(points):
        return _compute_cuda(points, features, classes, float(sampleDl), _order_flag(order))
    try:
        pts = np.ascontiguousarray(points, dtype=np.float32)
    except Exception:
        raise RuntimeError("Error converting input points to numpy arrays of type float32")
    feats = cls = None
    if features is not None:
        try:
            feats = np.ascontiguousarray(features, dtype=np.float32)
        except Exception:
            raise RuntimeError("Error converting input features to numpy arrays of type float32")
    if classes is not None:
        try:
            cls = np.ascontiguousarray(classes, dtype=np.int32)
        except Exception:
            raise RuntimeError("Error converting input classes to numpy arrays of type int32")
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    if feats is not None and feats.ndim != 2:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if cls is not None and cls.ndim > 2:
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    N = pts.shape[0]
    fdim = feats.shape[1] if feats is not None else 0
    ldim = (cls.shape[1] if cls.ndim == 2 else 1) if cls is not None else 0
    if feats is not None and feats.shape[0] != N:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if cls is not None and cls.shape[0] != N:
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    if N < 1:
        raise RuntimeError("Error")            # wrapper.cpp:225-229: empty result
    out_p = np.empty((N, 3), np.float32)
    out_f = np.empty((N, fdim), np.float32) if feats is not None else None
    out_c = np.empty((N, ldim), np.int32) if cls is not None else None
    M = C.c_int64(0)
    rc = _lib.lib().crfconv_grid_subsample_host(
        pts.ctypes.data, N, feats.ctypes.data if fdim else None, fdim, cls.ctypes.data if ldim else None, ldim,
        float(sampleDl), _order_flag(order), out_p.ctypes.data, out_f.ctypes.data if fdim else None,
        out_c.ctypes.data if ldim else None, C.byref(M))
    _lib.check(rc, "grid_subsampling.compute")
    m = M.value
    res = [out_p[:m].copy()]
    if feats is not None:
        res.append(out_f[:m].copy())
    if cls is not None:
        res.append(out_c[:m].copy())
    return res[0] if len(res) == 1 else tuple(res)


def _compute_cuda(points, features, classes, dl, order_flag, return_keys=False):
    import torch
    L = _lib.lib()
    pts = points.detach().to(torch.float32).contiguous()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    N = pts.shape[0]
    if N < 1:
        raise RuntimeError("Error")
    dev = pts.device
    feats = cls = None
    fdim = ldim = 0
    if features is not None:
        feats = features.detach().to(device=dev, dtype=torch.float32).contiguous()
        if feats.dim() != 2 or feats.shape[0] != N:
            raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
        fdim = feats.shape[1]
    if classes is not None:
        cls = classes.detach().to(device=dev, dtype=torch.int32).contiguous()
        if cls.dim() > 2 or cls.shape[0] != N:
            raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
        ldim = cls.shape[1] if cls.dim() == 2 else 1
    out_p = torch.empty((N, 3), dtype=torch.float32, device=dev)
    out_f = torch.empty((N, fdim), dtype=torch.float32, device=dev) if fdim else None
    out_c = torch.empty((N, ldim), dtype=torch.int32, device=dev) if ldim else None
    out_k = torch.empty((N,), dtype=torch.int64, device=dev) if return_keys else None
    M = C.c_int64(0)
    with torch.cuda.device(dev):
        nbytes = L.crfconv_grid_subsample_workspace_bytes(N, fdim, ldim)
        ws = _workspace(nbytes, dev)
        rc = L.crfconv_grid_subsample(pts.data_ptr(), N, feats.data_ptr() if fdim else None, fdim,
                                      cls.data_ptr() if ldim else None, ldim, dl, order_flag, out_p.data_ptr(),
                                      out_f.data_ptr() if fdim else None, out_c.data_ptr() if ldim else None,
                                      out_k.data_ptr() if return_keys else None, C.byref(M), ws.data_ptr(), ws.numel(),
                                      _lib.stream_ptr())
    _lib.check(rc, "grid_subsampling.compute")
    m = M.value
    res = [out_p[:m]]
    if fdim:
        res.append(out_f[:m])
    if ldim:
        res.append(out_c[:m])
    if return_keys:
        res.append(out_k[:m])
    return res[0] if len(res) == 1 else tuple(res)
