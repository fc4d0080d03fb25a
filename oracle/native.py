"""TEST INFRASTRUCTURE ONLY (oracle).  ctypes bindings for

* ``liboracle.so``  — our CPU restatement (oracle_native.cpp): brute-force exact kNN, voxel-grid subsampling;
* ``oracle/_ref/*`` — the unmodified reference C++ compiled by oracle/build.py (when it was built in the container).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

_oracle = None
_ref_knn = None
_ref_sub = None


def _lib():
    global _oracle
    if _oracle is None:
        _oracle = C.CDLL(_build.build_oracle())
        _oracle.oracle_knn_batch.restype = None
        _oracle.oracle_knn_batch.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, C.c_int64, C.c_int64, _i64p, C.c_void_p]
        _oracle.oracle_knn_distances.restype = None
        _oracle.oracle_knn_distances.argtypes = [_f32p, _f32p, C.c_int64, C.c_int64, _i64p, _f32p]
        _oracle.oracle_grid_subsample.restype = C.c_int64
        _oracle.oracle_grid_subsample.argtypes = [_f32p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                  C.c_float, C.c_int, _f32p, C.c_void_p, C.c_void_p, C.c_void_p]
        _oracle.oracle_knn_batch_distance_pick.restype = None
        _oracle.oracle_knn_batch_distance_pick.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, _i64p, _f32p]
        _oracle.oracle_fps.restype = None
        _oracle.oracle_fps.argtypes = [_f32p, _i64p, C.c_int64, _i64p, _i64p, _i64p, _i64p]
        _oracle.oracle_radius.restype = C.c_int64
        _oracle.oracle_radius.argtypes = [_f32p, _i64p, _f32p, _i64p, C.c_int64, C.c_float, C.c_int64, _i64p, _i64p]
    return _oracle


# --------------------------------------------------------------------------------------------------------- kNN
def knn_batch(pts, queries, K, return_dist=False):
    """Brute-force canonical kNN: ascending (squared f32 distance, index).  pts [B,N,3], queries [B,Q,3]."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    B, N, _ = pts.shape
    Q = queries.shape[1]
    idx = np.zeros((B, Q, K), dtype=np.int64)
    dist = np.zeros((B, Q, K), dtype=np.float32) if return_dist else None
    _lib().oracle_knn_batch(pts, B, N, queries, Q, K, idx, dist.ctypes.data if return_dist else None)
    return (idx, dist) if return_dist else idx


def knn_batch_distance_pick(pts, nqueries, K, seed):
    """knn.pyx:111-149 / knn_.cxx:138-203 with an explicit mt19937 seed.  Returns (indices [B,Q,K], queries [B,Q,3])."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    B, N, _ = pts.shape
    idx = np.zeros((B, nqueries, K), dtype=np.int64)
    q = np.zeros((B, nqueries, 3), dtype=np.float32)
    _lib().oracle_knn_batch_distance_pick(pts, B, N, nqueries, K, int(seed) & 0xFFFFFFFF, idx, q)
    return idx, q


def fps(pos, ptr, nsample, start=None):
    """Farthest point sampling per cloud (CSR offsets `ptr`); returns flat int64 indices into pos, cloud after cloud."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    B = len(ptr) - 1
    ns = np.ascontiguousarray(np.broadcast_to(np.asarray(nsample, dtype=np.int64), (B,)))
    st = np.zeros(B, dtype=np.int64) if start is None else np.ascontiguousarray(start, dtype=np.int64)
    optr = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    out = np.zeros(int(optr[-1]), dtype=np.int64)
    _lib().oracle_fps(pos, ptr, B, ns, st, out, optr)
    return out


def radius(x, ptr_x, y, ptr_y, r, max_num_neighbors):
    """(rows = query index into y, cols = support index into x): first max_num_neighbors supports within r, ascending index."""
    x, y = np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(y, dtype=np.float32)
    ptr_x, ptr_y = np.ascontiguousarray(ptr_x, dtype=np.int64), np.ascontiguousarray(ptr_y, dtype=np.int64)
    cap = y.shape[0] * int(max_num_neighbors)
    rows, cols = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
    e = _lib().oracle_radius(x, ptr_x, y, ptr_y, len(ptr_x) - 1, float(r), int(max_num_neighbors), rows, cols)
    return rows[:e].copy(), cols[:e].copy()


def knn(pts, queries, K, return_dist=False):
    r = knn_batch(np.asarray(pts)[None], np.asarray(queries)[None], K, return_dist)
    return (r[0][0], r[1][0]) if return_dist else r[0]


def knn_distances(pts, queries, idx):
    """Squared distances (reference arithmetic) from each query to the points named by idx [Q,K]."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.zeros(idx.shape, dtype=np.float32)
    _lib().oracle_knn_distances(pts, queries, idx.shape[0], idx.shape[1], idx, out)
    return out


# ------------------------------------------------------------------------------------------- grid subsampling
def grid_subsample(points, features=None, classes=None, sampleDl=0.1, order="key", return_keys=False):
    """order='key' → ascending voxel key (canonical set form); order='reference' → libstdc++ map iteration order."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    N = points.shape[0]
    fdim = ldim = 0
    f = c = None
    if features is not None:
        f = np.ascontiguousarray(features, dtype=np.float32)
        fdim = f.shape[1]
    if classes is not None:
        c = np.ascontiguousarray(classes, dtype=np.int32)
        ldim = 1 if c.ndim == 1 else c.shape[1]
    op = np.zeros((max(N, 1), 3), np.float32)
    of = np.zeros((max(N, 1), max(fdim, 1)), np.float32)
    oc = np.zeros((max(N, 1), max(ldim, 1)), np.int32)
    keys = np.zeros(max(N, 1), np.uint64)
    M = _lib().oracle_grid_subsample(points, N, f.ctypes.data if f is not None else None, fdim,
                                     c.ctypes.data if c is not None else None, ldim, float(sampleDl),
                                     1 if order == "reference" else 0, op, of.ctypes.data, oc.ctypes.data,
                                     keys.ctypes.data)
    res = [op[:M].copy()]
    if f is not None:
        res.append(of[:M, :fdim].copy())
    if c is not None:
        res.append(oc[:M, :ldim].copy())
    if return_keys:
        res.append(keys[:M].copy())
    return res[0] if len(res) == 1 else tuple(res)


# --------------------------------------------------------------------------- compiled reference (oracle/_ref)
def ref_paths():
    return _build.build_ref()


def have_ref_knn():
    return ref_paths()["knn"] is not None


def have_ref_subsample():
    return ref_paths()["subsample"] is not None


def _refknn():
    global _ref_knn
    if _ref_knn is None:
        p = ref_paths()["knn"]
        if p is None:
            raise RuntimeError("oracle/_ref/libref_knn.so missing (reference tree not mounted and no prebuilt copy)")
        _ref_knn = C.CDLL(p)
        _ref_knn.ref_knn.restype = None
        _ref_knn.ref_knn.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, _i64p, C.c_int]
        _ref_knn.ref_knn_batch.restype = None
        _ref_knn.ref_knn_batch.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, C.c_size_t,
                                           _i64p, C.c_int]
    return _ref_knn


def ref_knn(pts, queries, K, omp=False):
    """The reference's nearest_neighbors.knn (knn.pyx:33-69) minus Cython: same conversions, same C++ call."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    out = np.zeros((queries.shape[0], K), dtype=np.int64)
    _refknn().ref_knn(pts, pts.shape[0], pts.shape[1], queries, queries.shape[0], K, out, int(omp))
    return out


def ref_knn_batch_distance_pick(pts, nqueries, K, omp=False):
    """The reference's nearest_neighbors.knn_batch_distance_pick (knn.pyx:111-149) minus Cython (seeded with time(0) inside)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    lib = _refknn()
    lib.ref_knn_batch_distance_pick.restype = None
    lib.ref_knn_batch_distance_pick.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, _i64p, C.c_int]
    idx = np.zeros((pts.shape[0], nqueries, K), dtype=np.int64)
    q = np.zeros((pts.shape[0], nqueries, 3), dtype=np.float32)
    lib.ref_knn_batch_distance_pick(pts, pts.shape[0], pts.shape[1], pts.shape[2], q, nqueries, K, idx, int(omp))
    return idx, q


def ref_knn_batch(pts, queries, K, omp=False):
    """The reference's nearest_neighbors.knn_batch (knn.pyx:71-109) minus Cython."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    out = np.zeros((pts.shape[0], queries.shape[1], K), dtype=np.int64)
    _refknn().ref_knn_batch(pts, pts.shape[0], pts.shape[1], pts.shape[2], queries, queries.shape[1], K, out, int(omp))
    return out


def _refsub():
    global _ref_sub
    if _ref_sub is None:
        p = ref_paths()["subsample"]
        if p is None:
            raise RuntimeError("oracle/_ref/libref_subsample.so missing")
        _ref_sub = C.CDLL(p)
        _ref_sub.ref_grid_subsample_run.restype = C.c_long
        _ref_sub.ref_grid_subsample_run.argtypes = [_f32p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_float]
        _ref_sub.ref_grid_subsample_fetch.restype = None
        _ref_sub.ref_grid_subsample_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _ref_sub


def ref_grid_subsample(points, features=None, classes=None, sampleDl=0.1):
    """The reference grid_subsampling core with wrapper.cpp's conversions (f32 / f32 / int32) and return arity."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    N = points.shape[0]
    f = np.ascontiguousarray(features, dtype=np.float32) if features is not None else None
    c = np.ascontiguousarray(classes, dtype=np.int32) if classes is not None else None
    fdim = f.shape[1] if f is not None else 0
    ldim = (1 if c.ndim == 1 else c.shape[1]) if c is not None else 0
    lib = _refsub()
    M = lib.ref_grid_subsample_run(points, N, f.ctypes.data if f is not None else None, fdim,
                                   c.ctypes.data if c is not None else None, ldim, float(sampleDl))
    op = np.zeros((M, 3), np.float32)
    of = np.zeros((M, fdim), np.float32) if f is not None else None
    oc = np.zeros((M, ldim), np.int32) if c is not None else None
    lib.ref_grid_subsample_fetch(op.ctypes.data, of.ctypes.data if of is not None else None,
                                 oc.ctypes.data if oc is not None else None)
    res = [op]
    if of is not None:
        res.append(of)
    if oc is not None:
        res.append(oc)
    return res[0] if len(res) == 1 else tuple(res)
