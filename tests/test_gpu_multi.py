"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): NCCL process groups, one rank per GPU.
seed-sharded propagation + all-gather == the single-GPU result, scene-parallel ownership, bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from geoformer_b200.guidance import geodesic_guidance
        from geoformer_b200.parallel import scene_parallel_guidance, seed_sharded_guidance, shard_scenes
        from geoformer_b200.scenes import scene

        x = scene(60_000, 5).to(dev)
        for Q in (64, 37):  # even and ragged seed blocks
            seeds, geo = seed_sharded_guidance(x, Q, 16, 0.5, 24)
            ref_seeds, ref_geo = geodesic_guidance(x, Q, 16, 0.5, 24)
            assert torch.equal(seeds, ref_seeds)
            assert geo.shape == (Q, 60_000) and torch.equal(geo, ref_geo)
            # no exchange (gf_guidance_shard: FPS of all seeds, this rank's block of rows only)
            from geoformer_b200.parallel import shard_seeds

            seeds_b, block = seed_sharded_guidance(x, Q, 16, 0.5, 24, gather=False)
            q0, q1 = shard_seeds(Q, rank, world)
            assert torch.equal(seeds_b, ref_seeds) and block.shape == (q1 - q0, 60_000)
            assert torch.equal(block, ref_geo[q0:q1])
        # exchange fused into the propagation kernel (rows stored into the peers' matrices over NVLink)
        from geoformer_b200.parallel import SeedShardedRows, seed_sharded_guidance_fused

        for n_pts, Q in ((60_000, 64), (30_001, 37), (30_002, world)):  # aligned / unaligned rows, ragged blocks
            xs = scene(n_pts, 7).to(dev)
            rows = SeedShardedRows(Q, n_pts)
            for max_step in (24, 3, 24):  # reuse of the mapped matrices across calls
                seeds, geo = seed_sharded_guidance_fused(xs, Q, 16, 0.5, max_step, rows)
                ref_seeds, ref_geo = geodesic_guidance(xs, Q, 16, 0.5, max_step)
                assert torch.equal(seeds, ref_seeds)
                assert geo.shape == (Q, n_pts) and torch.equal(geo, ref_geo), (n_pts, Q, max_step)
            rows.close()
        scenes = [scene(20_000 + 1000 * s, 30 + s).to(dev) for s in range(5)]
        mine = scene_parallel_guidance(scenes, 32, 8, 0.5, 16)
        assert sorted(mine) == shard_scenes(5, rank, world)
        for s, (sd, g) in mine.items():
            rs, rg = geodesic_guidance(scenes[s], 32, 8, 0.5, 16)
            assert torch.equal(sd, rs) and torch.equal(g, rg)
        dist.barrier()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_seed_sharded_and_scene_parallel_nccl(tmp_path, cuda_lib):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))
