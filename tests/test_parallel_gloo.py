"""CPU tests of the multi-GPU host logic: world_size-2 process groups on the gloo backend, with the
oracle standing in for the CUDA compute callables (the partition / gather code is what is tested)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from geoformer_b200.parallel import shard_scenes, shard_seeds


def test_shard_helpers():
    assert shard_scenes(8, 1, 4) == [1, 5] and shard_scenes(3, 2, 4) == [2] and shard_scenes(2, 3, 4) == []
    for Q, world in ((256, 8), (10, 4), (3, 8), (512, 3)):
        blocks = [shard_seeds(Q, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == Q
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fps(xyz, q):
    import oracle

    return torch.from_numpy(oracle.furthest_point_sampling(xyz[None].numpy(), q)[0])


def _oracle_geodesic(xyz, seeds, k, radius, max_step):
    import oracle

    D, I = oracle.find_knn(xyz.numpy(), k)
    return torch.from_numpy(oracle.geodesic(D, I, seeds.numpy(), radius, max_step))


def _oracle_guidance(xyz, q, k, radius, max_step):
    seeds = _oracle_fps(xyz, q)
    return seeds, _oracle_geodesic(xyz, seeds, k, radius, max_step)


def _worker(rank, world, port, Q, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geoformer_b200.parallel import scene_parallel_guidance, seed_sharded_guidance
        from geoformer_b200.scenes import scene

        room = dict(L=(2.0, 1.5, 1.0), nbox=4)
        x = scene(3000, 11, **room)
        seeds, geo = seed_sharded_guidance(x, Q, 8, 0.2, 20, fps_fn=_oracle_fps, geodesic_fn=_oracle_geodesic)
        ref_seeds, ref_geo = _oracle_guidance(x, Q, 8, 0.2, 20)
        assert torch.equal(seeds, ref_seeds)
        assert geo.shape == (Q, 3000) and torch.equal(geo, ref_geo), "gathered rows differ from the single-process run"
        _, local = seed_sharded_guidance(x, Q, 8, 0.2, 20, fps_fn=_oracle_fps, geodesic_fn=_oracle_geodesic, gather=False)
        from geoformer_b200.parallel import shard_seeds

        q0, q1 = shard_seeds(Q, rank, world)
        assert torch.equal(local, ref_geo[q0:q1])
        scenes = [scene(1500 + 100 * s, 20 + s, **room) for s in range(3)]
        mine = scene_parallel_guidance(scenes, 8, 6, 0.2, 10, guidance_fn=_oracle_guidance)
        assert sorted(mine) == list(range(rank, 3, world))
        for s, (sd, g) in mine.items():
            rs, rg = _oracle_guidance(scenes[s], 8, 6, 0.2, 10)
            assert torch.equal(sd, rs) and torch.equal(g, rg)
        dist.barrier()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _run(Q, tmp_path):
    import oracle

    oracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, Q, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_seed_sharded_and_scene_parallel_world2_even(tmp_path):
    _run(16, tmp_path)


def test_seed_sharded_world2_ragged(tmp_path):
    _run(9, tmp_path)
