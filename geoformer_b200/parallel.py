"""Multi-GPU execution of the hot path: one process per GPU, torch.distributed for the plumbing.

The reference is single-GPU and has no collective anywhere on this path; the two
partitionings below are the ones the path itself offers:

  scene-parallel   scenes are independent (geodesic_utils.py:98 loops per scene): scene s runs on
                   rank s % world.  NO data-path collective.
  seed-sharded     one large scene: seeds are independent (rows of the output never interact), so
                   each rank propagates a contiguous block of seeds and the (Q/G, N) row blocks are
                   exchanged with ONE all-gather (NCCL over NVLink) -- or, fused (`SeedShardedRows`),
                   stored by the propagation kernel itself into every peer's matrix over NVLink as
                   each row finishes, so no collective follows.  FPS is inherently sequential
                   and the kNN graph is needed whole by every rank; both are cheap next to the
                   propagation of a million-point scene, so they are computed redundantly
                   (deterministic => identical on every rank, no broadcast needed).

The compute callables are injectable so that the partition / gather logic is exercised on CPU with
the gloo backend in tests; the defaults are the CUDA entry points (no CPU fallback).
"""
import torch
import torch.distributed as dist


def shard_scenes(n_scenes, rank, world):
    """indices of the scenes rank `rank` owns (round robin keeps ragged batches balanced)"""
    return list(range(rank, n_scenes, world))


def shard_seeds(n_seeds, rank, world):
    """contiguous block [q0, q1) of seeds of rank `rank`; blocks differ by at most one seed"""
    base, rem = divmod(n_seeds, world)
    q0 = rank * base + min(rank, rem)
    return q0, q0 + base + (1 if rank < rem else 0)


def _default_guidance(xyz, n_queries, neighbor, radius, max_step):
    from .guidance import geodesic_guidance

    return geodesic_guidance(xyz, n_queries, neighbor, radius, max_step)


def _default_fps(xyz, n_queries):
    from .pointnet2 import _ext

    return _ext.furthest_point_sampling(xyz[None].contiguous(), n_queries)[0]


def _default_geodesic(xyz, seeds, neighbor, radius, max_step):
    from .geodesic_utils import geodesic_from_points

    return geodesic_from_points(xyz, seeds, neighbor, radius, max_step)


def scene_parallel_guidance(scenes, n_queries, neighbor, radius, max_step, rank=None, world=None, guidance_fn=None):
    """scenes: list of (N_s, 3) tensors (every rank passes the same list, or at least its own
    entries).  Returns {scene index: (seeds, geo)} for the scenes this rank owns.  No collective."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    mine = shard_scenes(len(scenes), rank, world)
    if guidance_fn is not None or len(mine) < 2 or not scenes[mine[0]].is_cuda:
        fn = guidance_fn or _default_guidance
        return {s: fn(scenes[s], n_queries, neighbor, radius, max_step) for s in mine}
    # this rank's scenes are independent as well: one batched library call builds their graphs side by side and
    # propagates all their (scene, seed) pairs in one launch (scenes too large for it go one by one)
    from .guidance import BATCH_MAX_POINTS, geodesic_guidance, geodesic_guidance_batch

    small = [s for s in mine if scenes[s].size(0) <= BATCH_MAX_POINTS]
    out = {}
    if small:
        seeds, geos = geodesic_guidance_batch([scenes[s] for s in small], n_queries, neighbor, radius, max_step)
        out.update({s: (seeds[i], geos[i]) for i, s in enumerate(small)})
    for s in mine:
        if s not in out:
            out[s] = geodesic_guidance(scenes[s], n_queries, neighbor, radius, max_step)
    return out


class _DeviceView:
    """numba-style view of raw device memory, so torch can wrap it without a copy"""

    def __init__(self, address, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(address), False),
                                         "version": 2, "strides": None}


class SeedShardedRows:
    """The (Q, N) result matrix of a seed-sharded scene, allocated on every rank as peer-visible memory
    and mapped into every other rank (CUDA IPC over NVLink), so that the propagation kernel of rank r
    can store its finished rows straight into all of them (`gf_guidance_seeded_scatter`).
    Set up once per (Q, N) and reused; `close()` unmaps and frees."""

    def __init__(self, n_queries, n_points, group=None, device=None):
        import ctypes

        from . import _capi as C

        self.group, self.Q, self.N = group, int(n_queries), int(n_points)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        C.require(self.world - 1 <= 15, "seed-sharded scatter supports at most 16 ranks")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        L = C.lib()
        with torch.cuda.device(self.device):
            own = ctypes.c_void_p()
            C.check(L.gf_peer_alloc(ctypes.byref(own), self.Q * self.N * 4), "peer_alloc")
            self._own = own
            handle = (ctypes.c_ubyte * 64)()
            C.check(L.gf_peer_export(own, handle), "peer_export")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            every = torch.empty((self.world, 64), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            every = every.cpu()
            self._mapped = {}
            for r in range(self.world):
                if r == self.rank:
                    continue
                h = (ctypes.c_ubyte * 64)(*every[r].tolist())
                q = ctypes.c_void_p()
                C.check(L.gf_peer_open(h, ctypes.byref(q)), "peer_open(rank %d)" % r)
                self._mapped[r] = q
        self.geo = torch.as_tensor(_DeviceView(own.value, (self.Q, self.N)), device=self.device)
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)

    def row_pointers(self, row0):
        """(own pointer at row row0, ctypes array of the peers' pointers at the same row)"""
        import ctypes

        off = int(row0) * self.N * 4
        peers = [ctypes.c_void_p(p.value + off) for _, p in sorted(self._mapped.items())]
        return ctypes.c_void_p(self._own.value + off), (ctypes.c_void_p * max(len(peers), 1))(*peers), len(peers)

    def fence(self):
        """all ranks reach this point of their streams (a one-word all-reduce; orders peer stores and reads)"""
        dist.all_reduce(self._token, group=self.group)

    def close(self):
        from . import _capi as C

        L = C.lib()
        torch.cuda.synchronize(self.device)
        self.geo = None
        with torch.cuda.device(self.device):
            for q in self._mapped.values():
                L.gf_peer_close(q)
            self._mapped = {}
            dist.barrier(group=self.group)  # nobody frees memory a peer still has mapped
            if self._own is not None:
                L.gf_peer_free(self._own)
                self._own = None


def seed_sharded_guidance_fused(xyz, n_queries, neighbor, radius, max_step, rows, seeds=None):
    """One scene split by seed blocks, exchange fused into the propagation kernel.  `rows` is a
    `SeedShardedRows(n_queries, N)`; returns (seeds, rows.geo) with the full matrix valid on every
    rank once the stream reaches the end of this call."""
    import ctypes

    from . import _capi as C

    C.check_cuda_f32(xyz, "xyz")
    N = xyz.shape[0]
    C.require(xyz.dim() == 2 and xyz.shape[1] == 3 and N == rows.N and n_queries == rows.Q, "scene / rows mismatch")
    dev = xyz.device
    if seeds is None:
        seeds = _default_fps(xyz, n_queries)
    C.check_cuda_i32(seeds, "seeds")
    q0, q1 = shard_seeds(n_queries, rows.rank, rows.world)
    local = seeds[q0:q1].contiguous()
    L = C.lib()
    own, peers, n_peers = rows.row_pointers(q0)
    rows.fence()  # the peers have finished with the previous contents of their matrices
    if q1 > q0:
        with torch.cuda.device(dev):
            nbytes = L.gf_guidance_workspace_bytes(N, q1 - q0, neighbor)
            ws = C.workspace.get(dev, "guidance", nbytes)
            C.check(L.gf_guidance_seeded_scatter(C.ptr(xyz), N, C.ptr(local), q1 - q0, neighbor, float(radius),
                                                 int(max_step), own, ctypes.cast(peers, ctypes.c_void_p), n_peers,
                                                 None, C.ptr(ws), nbytes, C.stream_of(dev)), "guidance_seeded_scatter")
    rows.fence()  # every rank's kernel is complete: all rows have landed everywhere
    return seeds, rows.geo


def _guidance_shard(xyz, n_queries, q0, q1, neighbor, radius, max_step):
    """one library call: FPS of all seeds next to the kNN graph, then this rank's block of seeds (gf_guidance_shard)"""
    import ctypes

    from . import _capi as C

    C.check_cuda_f32(xyz, "xyz")
    N, dev = xyz.shape[0], xyz.device
    seeds = torch.empty((n_queries,), dtype=torch.int32, device=dev)
    block = torch.empty((q1 - q0, N), dtype=torch.float32, device=dev)
    L = C.lib()
    with torch.cuda.device(dev):
        nbytes = L.gf_guidance_workspace_bytes(N, n_queries, int(neighbor))
        ws = C.workspace.get(dev, "guidance", nbytes)
        C.check(L.gf_guidance_shard(C.ptr(xyz), N, n_queries, q0, q1, int(neighbor), ctypes.c_float(float(radius)),
                                    int(max_step), C.ptr(seeds), C.ptr(block), None, None, C.ptr(ws), nbytes,
                                    C.stream_of(dev)), "guidance_shard")
    return seeds, block


def seed_sharded_guidance(xyz, n_queries, neighbor, radius, max_step, group=None, fps_fn=None, geodesic_fn=None,
                          gather=True):
    """One scene split by seed blocks.  Returns (seeds (Q,), geo): geo is the full (Q, N) matrix on
    every rank when gather=True, else this rank's (Q_local, N) block.
    gather=False is the mode that SCALES: nothing is exchanged, so the consumer of the maps (the decoder's bias
    epilogue, the mask head) has to be sharded by query as well -- each of those is row-wise in the queries.
    Replicating the 2 GB result of a 1M-point scene on every rank costs more than the propagation it follows."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    q0, q1 = shard_seeds(n_queries, rank, world)
    if fps_fn is None and geodesic_fn is None and xyz.is_cuda:
        seeds, local = _guidance_shard(xyz, n_queries, q0, q1, neighbor, radius, max_step)
    else:
        seeds = (fps_fn or _default_fps)(xyz, n_queries)
        local = (geodesic_fn or _default_geodesic)(xyz, seeds[q0:q1].contiguous(), neighbor, radius, max_step)
    if not gather or world == 1:
        return seeds, local
    N = xyz.shape[0]
    base, rem = divmod(n_queries, world)
    if rem == 0:  # equal blocks: one all_gather straight into the (Q, N) output
        out = torch.empty((n_queries, N), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return seeds, out
    # ragged blocks: pad to the largest block, gather, drop the padding rows
    rows = base + 1
    padded = torch.empty((rows, N), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * rows, N), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_seeds(n_queries, r, world)
        parts.append(buf[r * rows: r * rows + (b - a)])
    return seeds, torch.cat(parts, dim=0)
