"""Multi-GPU execution of the hot path: one process per GPU, torch.distributed for the plumbing.

The reference is single-GPU and has no collective anywhere on this path; the two
partitionings below are the ones the path itself offers:

  scene-parallel   scenes are independent (geodesic_utils.py:98 loops per scene): scene s runs on
                   rank s % world.  NO data-path collective.
  seed-sharded     one large scene: seeds are independent (rows of the output never interact), so
                   each rank propagates a contiguous block of seeds and the (Q/G, N) row blocks are
                   exchanged with ONE all-gather (NCCL over NVLink).  FPS is inherently sequential
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
    fn = guidance_fn or _default_guidance
    return {s: fn(scenes[s], n_queries, neighbor, radius, max_step) for s in shard_scenes(len(scenes), rank, world)}


def seed_sharded_guidance(xyz, n_queries, neighbor, radius, max_step, group=None, fps_fn=None, geodesic_fn=None,
                          gather=True):
    """One scene split by seed blocks.  Returns (seeds (Q,), geo): geo is the full (Q, N) matrix on
    every rank when gather=True, else this rank's (Q_local, N) block."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    seeds = (fps_fn or _default_fps)(xyz, n_queries)
    q0, q1 = shard_seeds(n_queries, rank, world)
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
