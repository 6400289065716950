"""Drop-in for model/geoformer/geodesic_utils.py (find_knn, cal_geodesic_vectorize,
unique_with_inds) plus the index object that stands where faiss.GpuIndexFlatL2 stood
(geoformer_fs.py:170-175).  All arithmetic runs in the sm_100a library; there is no torch
fallback path for the propagation.
"""
import ctypes

import torch

from . import _capi as C
from .guidance import BATCH_MAX_POINTS, geodesic_guidance_batch


class FlatL2Index:
    """Exact L2 index over 3-D points with the faiss protocol used by the reference
    (geodesic_utils.py:18-21): add(x) / search(x, k, D_out, I_out) / reset().
    search() writes SQUARED distances (float32) and int64 indices into the pre-allocated
    tensors, ordered by (distance, index); missing neighbours are (-1, +inf)."""

    def __init__(self, algo="grid"):
        self.algo = {"grid": 0, "brute": 1}[algo]
        self._x = None
        self._src = None

    @property
    def ntotal(self):
        return 0 if self._x is None else self._x.shape[0]

    def add(self, x):
        C.check_cuda_f32(x, "x")
        C.require(x.dim() == 2 and x.size(1) == 3, "x must be (N, 3)")
        # faiss copies what it is given: later changes of the caller's tensor must not change the index
        self._x = x.detach().clone() if self._x is None else torch.cat([self._x, x]).contiguous()
        self._src = (x.data_ptr(), tuple(x.shape)) if self.ntotal == x.size(0) else None

    def reset(self):
        self._x = None
        self._src = None

    def search(self, q, k, D_out, I_out):
        C.require(self._x is not None, "search on an empty index")
        C.check_cuda_f32(q, "q")
        C.check_cuda_f32(D_out, "D_out")
        C.require(I_out.is_cuda and I_out.is_contiguous() and I_out.dtype == torch.int64, "I_out must be int64 CUDA")
        nq = q.size(0)
        C.require(tuple(D_out.shape) == (nq, k) and tuple(I_out.shape) == (nq, k), "output shape must be (nq, k)")
        x = self._x
        # the reference searches a scene against itself right after adding it (geodesic_utils.py:18-19): then the
        # self-query path (no second copy of the points) applies
        same = self._src == (q.data_ptr(), tuple(q.shape))
        _knn_into(x, None if same else q, int(k), False, D_out, I_out, None, self.algo)
        return D_out, I_out


def _knn_into(x, queries, k, sqrt_out, dist, idx64, idx32, algo):
    L = C.lib()
    N = x.size(0)
    nq = N if queries is None else queries.size(0)
    with torch.cuda.device(x.device):
        nbytes = L.gf_knn_workspace_bytes(N, nq, k, algo)
        ws = C.workspace.get(x.device, "knn", nbytes) if nbytes else None
        C.check(L.gf_knn(C.ptr(x), N, C.ptr(queries), nq, k, 1 if sqrt_out else 0, C.ptr(dist), C.ptr(idx64),
                         C.ptr(idx32), algo, C.ptr(ws), nbytes, C.stream_of(x.device)), "knn")


def knn_graph(locs, neighbor, algo="grid", index_dtype=torch.int64):
    """kNN of a scene against itself: (sqrt distances (N,k) f32, indices (N,k) int64 or int32)."""
    C.check_cuda_f32(locs, "locs")
    N = locs.size(0)
    D = torch.empty((N, neighbor), dtype=torch.float32, device=locs.device)
    I = torch.empty((N, neighbor), dtype=index_dtype, device=locs.device)
    a = {"grid": 0, "brute": 1}[algo]
    if index_dtype == torch.int64:
        _knn_into(locs, None, int(neighbor), True, D, I, None, a)
    else:
        _knn_into(locs, None, int(neighbor), True, D, None, I, a)
    return D, I


@torch.no_grad()
def find_knn(gpu_index, locs, neighbor=32):
    """geodesic_utils.py:11-24: (sqrt distances (N,k) f32, indices (N,k) i64) through the index
    protocol.  With our own FlatL2Index the sqrt is fused into the search kernel."""
    if isinstance(gpu_index, FlatL2Index) or gpu_index is None:
        # the reference adds the scene, searches it against itself and resets the index (:18-21): points added
        # earlier would take part in the search, so an index that is not empty is an error here, not ignored
        C.require(gpu_index is None or gpu_index.ntotal == 0,
                  "find_knn: the index already holds %d points (the reference resets it after every call)"
                  % (0 if gpu_index is None else gpu_index.ntotal))
        algo = "grid" if gpu_index is None or gpu_index.algo == 0 else "brute"
        return knn_graph(locs.contiguous(), neighbor, algo=algo)
    n_points = locs.shape[0]
    D = torch.zeros(n_points, neighbor, device=locs.device, dtype=torch.float32)
    I = torch.zeros(n_points, neighbor, device=locs.device, dtype=torch.int64)
    gpu_index.add(locs)
    gpu_index.search(locs, neighbor, D, I)
    gpu_index.reset()
    return torch.sqrt(D), I


def unique_with_inds(x, dim=-1):
    """geodesic_utils.py:4-8: unique slices along `dim` and, for each, the index of its FIRST
    occurrence.  The reference relies on scatter_ with duplicate indices (last write wins), which is
    only deterministic on CPU; this version is deterministic on CUDA too (amin reduction).  It is
    not used by our propagation kernel -- kept because it is part of the module's surface."""
    unique, inverse = torch.unique(x, return_inverse=True, dim=dim)
    perm = torch.arange(inverse.size(0), dtype=inverse.dtype, device=inverse.device)
    first = torch.full((unique.size(dim),), inverse.size(0), dtype=inverse.dtype, device=inverse.device)
    first.scatter_reduce_(0, inverse, perm, reduce="amin", include_self=True)
    return unique, first


def geodesic_from_graph(D, I, seeds, radius, max_step, return_stats=False, row_max=None):
    """Propagation only (geodesic_utils.py:109-163) on a given kNN graph: D (N,k) sqrt'ed f32,
    I (N,k) int64/int32, seeds (Q,) int32/int64 -> (Q,N) f32, -1 = unreachable."""
    C.check_cuda_f32(D, "D")
    C.require(I.is_cuda and I.is_contiguous() and I.dtype in (torch.int64, torch.int32), "I must be int64/int32 CUDA")
    N, k = D.shape
    seeds = seeds.to(device=D.device, dtype=torch.int32).contiguous()
    Q = seeds.numel()
    geo = torch.empty((Q, N), dtype=torch.float32, device=D.device)
    stats = torch.zeros(2, dtype=torch.int64, device=D.device) if return_stats else None
    L = C.lib()
    with torch.cuda.device(D.device):
        nbytes = L.gf_geodesic_workspace_bytes(N, k, Q)
        ws = C.workspace.get(D.device, "geodesic", nbytes) if nbytes else None
        C.check(L.gf_geodesic(C.ptr(D), C.ptr(I), 1 if I.dtype == torch.int64 else 0, N, k, C.ptr(seeds), Q,
                              ctypes.c_float(float(radius)), int(max_step), C.ptr(geo), C.ptr(stats), C.ptr(row_max),
                              C.ptr(ws), nbytes, C.stream_of(D.device)), "geodesic")
    return (geo, stats) if return_stats else geo


def geodesic_from_points(locs, seeds, neighbor, radius, max_step, return_graph=False, return_stats=False,
                         row_max=None, ws_tag="guidance"):
    """kNN graph + propagation of one scene in one library call (the body of the reference loop,
    geodesic_utils.py:98-163)."""
    C.check_cuda_f32(locs, "locs")
    N = locs.size(0)
    seeds = seeds.to(device=locs.device, dtype=torch.int32).contiguous()
    Q = seeds.numel()
    geo = torch.empty((Q, N), dtype=torch.float32, device=locs.device)
    D = torch.empty((N, neighbor), dtype=torch.float32, device=locs.device) if return_graph else None
    I = torch.empty((N, neighbor), dtype=torch.int32, device=locs.device) if return_graph else None
    stats = torch.zeros(2, dtype=torch.int64, device=locs.device) if return_stats else None
    L = C.lib()
    with torch.cuda.device(locs.device):
        nbytes = L.gf_guidance_workspace_bytes(N, Q, int(neighbor))
        ws = C.workspace.get(locs.device, ws_tag, nbytes)
        C.check(L.gf_guidance_seeded(C.ptr(locs), N, C.ptr(seeds), Q, int(neighbor), ctypes.c_float(float(radius)),
                                     int(max_step), C.ptr(geo), C.ptr(D), C.ptr(I), C.ptr(stats), C.ptr(row_max),
                                     C.ptr(ws), nbytes, C.stream_of(locs.device)), "guidance_seeded")
    out = [geo]
    if return_graph:
        out += [D, I]
    if return_stats:
        out.append(stats)
    return out[0] if len(out) == 1 else tuple(out)


@torch.no_grad()
def cal_geodesic_vectorize(gpu_index, pre_enc_inds, locs_float_, batch_offset_, max_step=128, neighbor=64,
                           radius=0.05, n_queries=128):
    """geodesic_utils.py:91-164, same signature and return value: a list over the batch of
    (n_queries, N_b) float32 tensors on locs_float_.device, -1 = unreachable.

    gpu_index: a FlatL2Index (or None) -> kNN and propagation run fused in the library;
               any other object with add/search/reset (e.g. a real faiss GpuIndexFlatL2) -> its
               neighbours are used as they are and only the propagation runs here."""
    batch_size = pre_enc_inds.shape[0]
    offsets = batch_offset_.tolist() if isinstance(batch_offset_, torch.Tensor) else list(batch_offset_)
    own_index = gpu_index is None or isinstance(gpu_index, FlatL2Index)
    fused = own_index and (gpu_index is None or gpu_index.algo == 0)
    dev = locs_float_.device
    geo_dists = [None] * batch_size
    batch = []  # (b, points, seeds) of the scenes that go through the batched library call
    for b in range(batch_size):
        start, end = int(offsets[b]), int(offsets[b + 1])
        seeds = pre_enc_inds[b][:n_queries]
        if end - start == 0:
            geo_dists[b] = torch.empty((seeds.numel(), 0), dtype=torch.float32, device=dev)
            continue
        locs_b = locs_float_[start:end].contiguous()
        if fused and end - start <= BATCH_MAX_POINTS and seeds.numel() > 0:
            batch.append((b, locs_b, seeds))
        elif fused:
            geo_dists[b] = geodesic_from_points(locs_b, seeds, neighbor, radius, max_step)
        else:
            D, I = find_knn(gpu_index, locs_b, neighbor=neighbor)
            geo_dists[b] = geodesic_from_graph(D, I, seeds, radius, max_step)
    # The scenes of a batch are independent (the reference loops over them, :98): their graphs are built side by
    # side and ALL their (scene, seed) pairs are propagated by one launch (gf_guidance_batch) -- the per-seed runs
    # are chains of dependent latencies, so the GPU is only filled by many of them at once.
    # Scenes with the same number of seeds share a call (pre_enc_inds is one (B, >= n_queries) tensor: all of them).
    groups = {}
    for b, locs_b, seeds in batch:
        groups.setdefault(seeds.numel(), []).append((b, locs_b, seeds))
    for q_n, items in groups.items():
        _, geos = geodesic_guidance_batch([x for _, x, _ in items], q_n, neighbor, radius, max_step,
                                          seeds=[s for _, _, s in items])
        for (b, _, _), g in zip(items, geos):
            geo_dists[b] = g
    return geo_dists


_MAX_SCENES_IN_FLIGHT = 4
_lane_streams = {}


def _side_streams(device, n):
    """up to n cached side streams of `device` (created once per device and thread of first use)"""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    have = _lane_streams.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device=device))
    return have[:n]
