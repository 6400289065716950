"""CPU oracle for the geodesic-guidance hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  It is never imported by geoformer_b200 (the product), which has no CPU fallback.

The arithmetic lives in oracle.c (C, explicit fmaf, -ffp-contract=off); this module is a thin numpy
front-end over it.  Function names follow the reference operators they restate.
Parity status: FPS / ball_query / gather / group / three_* pinned by the reference's own CUDA
kernels (oracle/_ref, GPU box); geodesic pinned by the reference's own cal_geodesic_vectorize on
CPU (tests/golden); kNN PARITY UNPINNED (faiss-gpu is an absent, un-pinned dependency).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile oracle.c with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "_build/liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP team size of the C routines (overrides an OMP_NUM_THREADS exported by a launcher)."""
    lib().orc_set_num_threads(ctypes.c_int(int(n)))
    return num_threads()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_ci = ctypes.c_int
_cf = ctypes.c_float


def furthest_point_sampling(xyz, npoint):
    """xyz (B,N,3) f32 -> idx (B,npoint) i32.  sampling.cpp:67-88 + sampling_gpu.cu:72-176."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), dtype=np.int32)
    lib().orc_fps(_p(xyz), _ci(B), _ci(N), _ci(npoint), _p(idx))
    return idx


def gather_points(points, idx):
    """points (B,C,N) f32, idx (B,m) i32 -> (B,C,m).  sampling_gpu.cu:11-23."""
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    m = idx.shape[1]
    out = np.zeros((B, C, m), dtype=np.float32)
    lib().orc_gather_points(_p(points), _p(idx), _ci(B), _ci(C), _ci(N), _ci(m), _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().orc_gather_points_grad(_p(grad_out), _p(idx), _ci(B), _ci(C), _ci(n), _ci(m), _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """centres first (bindings order, pointnet2_utils.py:260).  ball_query_gpu.cu:12-47."""
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, m, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = np.zeros((B, m, nsample), dtype=np.int32)
    lib().orc_ball_query(_p(new_xyz), _p(xyz), _ci(B), _ci(N), _ci(m), _cf(radius), _ci(nsample), _p(idx))
    return idx


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, np_, ns = idx.shape
    out = np.zeros((B, C, np_, ns), dtype=np.float32)
    lib().orc_group_points(_p(points), _p(idx), _ci(B), _ci(C), _ci(N), _ci(np_), _ci(ns), _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, np_, ns = grad_out.shape
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().orc_group_points_grad(_p(grad_out), _p(idx), _ci(B), _ci(C), _ci(n), _ci(np_), _ci(ns), _p(out))
    return out


def three_nn(unknown, known):
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    lib().orc_three_nn(_p(unknown), _p(known), _ci(B), _ci(n), _ci(m), _p(d2), _p(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().orc_three_interpolate(_p(points), _p(idx), _p(weight), _ci(B), _ci(C), _ci(m), _ci(n), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, m), dtype=np.float32)
    lib().orc_three_interpolate_grad(_p(grad_out), _p(idx), _p(weight), _ci(B), _ci(C), _ci(n), _ci(m), _p(out))
    return out


def knn_sq(x, k, queries=None):
    """Canonical exact kNN.  Returns (D2 (nq,k) f32 SQUARED, I (nq,k) i64)."""
    x = _f32(x)
    q = x if queries is None else _f32(queries)
    N, nq = x.shape[0], q.shape[0]
    D2 = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    lib().orc_knn(_p(x), _ci(N), _p(q), _ci(nq), _ci(k), _p(D2), _p(I))
    return D2, I


def find_knn(x, k):
    """geodesic_utils.py:11-24 contract: (sqrt distances (N,k) f32, indices (N,k) i64)."""
    D2, I = knn_sq(x, k)
    return np.sqrt(D2), I


def geodesic(D, I, seeds, radius, max_step, return_stats=False):
    """geodesic_utils.py:91-164 for one scene.  D (N,k) sqrt'ed distances, I (N,k) i64 (column 0
    dropped inside, :110-111), seeds (Q,) -> geo (Q,N) f32 with -1 = unreachable."""
    D, I, seeds = _f32(D), _i64(I), _i64(seeds)
    N, k = D.shape
    Q = seeds.shape[0]
    geo = np.empty((Q, N), dtype=np.float32)
    reached = ctypes.c_int64(0)
    levels = ctypes.c_int(0)
    r32 = float(np.float32(radius))
    lib().orc_geodesic(_p(D), _p(I), _ci(N), _ci(k), _p(seeds), _ci(Q), _cf(r32), _ci(max_step), _p(geo),
                       ctypes.byref(reached), ctypes.byref(levels))
    if return_stats:
        return geo, int(reached.value), int(levels.value)
    return geo


def cal_geodesic_vectorize(pre_enc_inds, locs_float, batch_offsets, max_step=128, neighbor=64, radius=0.05,
                           n_queries=128):
    """Whole reference entry point (geodesic_utils.py:91-164) on numpy inputs: list of (Q,N_b)."""
    out = []
    for b in range(pre_enc_inds.shape[0]):
        s, e = int(batch_offsets[b]), int(batch_offsets[b + 1])
        D, I = find_knn(locs_float[s:e], neighbor)
        out.append(geodesic(D, I, pre_enc_inds[b][:n_queries], radius, max_step))
    return out
