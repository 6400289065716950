"""The propagation kernel has three state layouts (both bitmaps on chip / visited bitmap only / no on-chip
state, chosen by scene size) and a global overflow area for frontiers beyond the on-chip queue.  The size
rule would exercise the last two only at ~1M points, so they are forced here through the kernel's test knob
(GF_GEO_NOBITMAP, read once per process => one subprocess per mode) on scenes the oracle finishes in
seconds, including frontiers of tens of thousands of points (k = 64, no radius filter)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
import numpy as np
import torch
sys.path.insert(0, %r)
import oracle
from geoformer_b200.geodesic_utils import geodesic_from_graph, knn_graph
from geoformer_b200.scenes import scene
dev = torch.device("cuda:0")
cases = [
    # N, room, Q, k, radius, max_step
    (30000, dict(), 8, 16, 0.5, 40),
    (50000, dict(L=(1.0, 1.0, 0.5), nbox=2), 6, 64, 10.0, 6),    # frontiers >> 4096: overflow area
    (20000, dict(L=(2.0, 1.5, 1.0), nbox=4), 300, 8, 0.2, 300),  # more seeds than CTAs: persistent loop, deep
    (4097, dict(L=(1.0, 1.0, 0.5), nbox=2), 5, 33, 10.0, 4),     # K = 32 exactly (KP = 32), N just over a queue
]
for N, room, Q, k, r, ms in cases:
    x = scene(N, 17, **room)
    seeds = oracle.furthest_point_sampling(x[None].numpy(), Q)[0]
    D, I = knn_graph(x.to(dev), k)
    want = oracle.geodesic(D.cpu().numpy(), I.cpu().numpy(), seeds, r, ms)
    rm = torch.zeros(Q, device=dev)
    got = geodesic_from_graph(D, I, torch.from_numpy(seeds).to(dev), r, ms, row_max=rm).cpu().numpy()
    assert np.array_equal(got, want), (N, Q, k, r, ms, int((got != want).sum()))
    assert np.array_equal(rm.cpu().numpy(), want.max(axis=1)), "row_max by-product"
    print("ok", N, Q, k, r, ms, "reached", int((want >= 0).sum()))
"""


# (GF_GEO_NOBITMAP, GF_GEO_ENC, GF_GEO_UNROLL, GF_GEO_THREADS): every kernel the size rule or a knob can select
VARIANTS = [("0", "2", "2", "0"),     # the default: batched two-bitmap kernel, thread count by the number of seeds
            ("0", "2", "1", "1024"), ("0", "2", "2", "512"), ("0", "2", "1", "256"), ("0", "2", "2", "256"),
            ("0", "0", "2", "1024"),  # the per-scene template kernel with plain edge targets (round 1)
            ("0", "0", "2", "512"),
            ("1", "2", "2", "1024"),  # no on-chip state (the layout of scenes beyond ~860k points)
            ("2", "2", "2", "1024")]  # visited bitmap only


@pytest.mark.parametrize("mode,enc,unroll,threads", VARIANTS)
def test_geodesic_state_layouts_and_overflow(cuda_lib, oracle_lib, mode, enc, unroll, threads):
    env = dict(os.environ)
    env["GF_GEO_NOBITMAP"] = mode  # 0 = size rule (both bitmaps here), 1 = no on-chip state, 2 = visited bitmap only
    env["GF_GEO_ENC"], env["GF_GEO_UNROLL"], env["GF_GEO_THREADS"] = enc, unroll, threads
    out = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, cwd=ROOT, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok ") == 4
