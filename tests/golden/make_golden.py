"""Generates the committed golden fixtures in tests/golden/ (run in the BUILD container only).

    python tests/golden/make_golden.py

geodesic_*.npz  outputs of the reference's OWN cal_geodesic_vectorize
                (/root/reference/model/geoformer/geodesic_utils.py, loaded unmodified by
                oracle/ref_geodesic.py) on CPU, fed by the canonical kNN of oracle.c through a faiss-shaped
                stand-in index, with torch.sqrt replaced by a correctly rounded sqrt (what the reference
                gets on its real platform, CUDA; torch's CPU sqrt goes through MKL VML and is off by one ulp
                for ~0.7 % of inputs -- see oracle/ref_geodesic.py).  The unmodified-torch run is checked
                here too, to 1e-6 relative with identical unreachable sets.
The reference has no golden vectors or known-answer tests of its own for this path,
so these files ARE the pin of the geodesic oracle.  FPS / ball_query / gather / group / three_* fixtures
come from the reference's own CUDA kernels and are produced on the GPU box by
tests/golden/make_golden_ref_ext.py (the kernels are CUDA-only).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from geoformer_b200.scenes import scene  # noqa: E402
from oracle import ref_geodesic  # noqa: E402

# small rooms (2 x 1.5 x 1 m, 4 boxes) so that a few thousand points are as dense as a real scan
ROOM = dict(L=(2.0, 1.5, 1.0), nbox=4)
CASES = {
    # name: (N, scene seed, Q, k, radius, max_step, duplicates?)
    "geodesic_small": (2000, 4, 16, 8, 0.2, 64, False),     # runs to exhaustion
    "geodesic_dups": (2000, 5, 16, 8, 0.2, 64, True),       # exact duplicate points / seeds in own row
    "geodesic_dense": (4000, 6, 24, 16, 0.5, 6, False),     # level bound hit, many ties per level
    "geodesic_deep": (6000, 7, 8, 6, 0.05, 300, False),     # tight radius: long thin fronts
}


def make_case(name, N, seed, Q, k, radius, max_step, dups):
    x = scene(N, seed, **ROOM)
    if dups:
        x[10:20] = x[0:10]
        x[5] = x[N - 1]
    seeds = oracle.furthest_point_sampling(x[None].numpy(), Q)[0]
    if dups:
        seeds[:4] = [12, 5, N - 1, 0]
    D, I = oracle.find_knn(x.numpy(), k)
    args = (torch.from_numpy(seeds[None].copy()), x, torch.tensor([0, N], dtype=torch.int32), max_step, k, radius, Q)
    ref = ref_geodesic.reference_cal_geodesic(*args, ieee_sqrt=True)[0].numpy()
    raw = ref_geodesic.reference_cal_geodesic(*args, ieee_sqrt=False)[0].numpy()
    assert np.array_equal(raw < 0, ref < 0) and np.allclose(raw, ref, rtol=1e-6, atol=0), name
    mine = oracle.geodesic(D, I, seeds, radius, max_step)
    assert np.array_equal(mine, ref), "oracle.c disagrees with the reference on %s" % name
    np.savez_compressed(os.path.join(HERE, name + ".npz"), xyz=x.numpy(), seeds=seeds.astype(np.int32), knn_dist=D,
                        knn_idx=I, geo=ref, k=np.int32(k), radius=np.float32(radius), max_step=np.int32(max_step))
    lv = oracle.geodesic(D, I, seeds, radius, max_step, return_stats=True)[2]
    print("%-16s N=%d Q=%d k=%d r=%g max_step=%d reach=%.4f levels=%d (%d bytes)" % (
        name, N, Q, k, radius, max_step, float((ref >= 0).mean()), lv, os.path.getsize(os.path.join(HERE, name + ".npz"))))


if __name__ == "__main__":
    assert ref_geodesic.available(), "needs /root/reference (build container)"
    for name, cfg in CASES.items():
        make_case(name, *cfg)
