"""Synthetic ScanNet / S3DIS-shaped scenes (SURVEY.md App. C -- the data spec every number in
BASELINE.md / DESIGN.md is quoted on).  Pure data generation on the CPU generator, so the same
seed gives the same points on every platform.  The order of the random draws is part of the spec.
"""
import torch


def scene(n, seed, L=(8.0, 6.0, 3.0), nbox=20):
    """(n,3) contiguous float32: 40 % floor, 30 % walls, 30 % box faces, 5 mm jitter, mean-centred
    (ScanNet convention, data/scannetv2/prepare_data_inst.py:44), shuffled."""
    g = torch.Generator().manual_seed(seed)
    L = torch.tensor(L)
    nf = int(0.4 * n)
    nw = int(0.3 * n)
    nb = n - nf - nw
    fl = torch.rand(nf, 3, generator=g) * L
    fl[:, 2] = 0
    w = torch.rand(nw, 3, generator=g) * L
    side = torch.randint(0, 4, (nw,), generator=g)
    w[side == 0, 0] = 0
    w[side == 1, 0] = L[0]
    w[side == 2, 1] = 0
    w[side == 3, 1] = L[1]
    c = torch.rand(nbox, 3, generator=g) * L * torch.tensor([1, 1, 0.3])
    s = 0.3 + torch.rand(nbox, 3, generator=g) * 0.9
    bi = torch.randint(0, nbox, (nb,), generator=g)
    u = torch.rand(nb, 3, generator=g) - 0.5
    ax = torch.randint(0, 3, (nb,), generator=g)
    sg = (torch.randint(0, 2, (nb,), generator=g) * 2 - 1).float()
    u[torch.arange(nb), ax] = 0.5 * sg
    b = c[bi] + u * s[bi]
    b[:, 2] = b[:, 2].abs()
    x = torch.cat([fl, w, b]) + torch.randn(n, 3, generator=g) * 0.005
    x = x - x.mean(0)
    return x[torch.randperm(n, generator=g)].contiguous().float()


def room(n, seed):
    """S3DIS-room-shaped scene for config c4 (10 x 8 x 3.2 m, 60 boxes)."""
    return scene(n, seed, L=(10.0, 8.0, 3.2), nbox=60)


# BASELINE.json configs: name -> generator call and hot-path parameters.
CONFIGS = {
    "c1": dict(gen="scene", n=50_000, seed=1234, Q=128, k=8, radius=0.5, max_step=32),
    "c2": dict(gen="scene", n=100_000, seed=1234, Q=256, k=16, radius=0.5, max_step=32),
    "c4": dict(gen="room", n=1_000_000, seed=4321, Q=512, k=16, radius=0.5, max_step=32),
}


def make(name, scene_index=0):
    """Points of a named config; c3 = eight c2 scenes with seeds 1234+s."""
    cfg = CONFIGS["c2" if name == "c3" else name]
    gen = scene if cfg["gen"] == "scene" else room
    return gen(cfg["n"], cfg["seed"] + scene_index)
