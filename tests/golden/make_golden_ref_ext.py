"""Runs the reference's OWN pointnet2 CUDA kernels (oracle/_ref, built unmodified from
/root/reference/lib/pointnet2/_ext_src by oracle/build_ref.py) on seeded inputs and stores inputs +
outputs as fixtures.  The kernels are CUDA-only, so this runs on the GPU box:

    gpurun -- python tests/golden/make_golden_ref_ext.py gpurun_out/ref_ext_golden.npz

The resulting file is committed as tests/golden/ref_ext_golden.npz; it pins oracle.c's FPS / ball_query /
gather / group / three_nn / three_interpolate restatements (tests/test_oracle_golden.py) and our kernels
(tests/test_gpu_vs_reference_ext.py) against the reference itself.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from geoformer_b200.scenes import scene  # noqa: E402
from oracle.build_ref import load_ref_ext  # noqa: E402


def lattice(n, seed):
    g = np.random.default_rng(seed)
    x = g.integers(-3, 4, size=(n, 3)).astype(np.float32) * 0.25
    x[g.integers(0, n, size=max(1, n // 50))] = 0.0
    x[g.integers(0, n, size=max(1, n // 50))] = np.float32(0.01)
    return x


def main(out_path):
    ext = load_ref_ext()
    assert ext is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    out = {}
    # FPS: scene-like clouds, tie-heavy lattices, tiny and non-power-of-two sizes
    fps_cases = {"scene20k": (scene(20000, 3).numpy()[None], 300), "scene3x511": (
        np.stack([scene(511, 10 + b).numpy() for b in range(3)]), 100),
        "lattice4096": (lattice(4096, 1)[None], 512), "lattice513": (lattice(513, 2)[None], 300),
        "lattice64": (lattice(64, 3)[None], 64), "tiny7": (np.random.default_rng(4).normal(size=(2, 7, 3)).astype(np.float32), 5),
        "room70k": (scene(70000, 9).numpy()[None], 128)}
    for name, (xyz, m) in fps_cases.items():
        idx = ext.furthest_point_sampling(torch.from_numpy(xyz).to(dev), m).cpu().numpy()
        out["fps_%s_xyz" % name], out["fps_%s_m" % name], out["fps_%s_idx" % name] = xyz, np.int32(m), idx
    # ball query + grouping + gather
    xyz = scene(20000, 7)[None]
    centres = xyz[:, torch.randperm(20000, generator=torch.Generator().manual_seed(2))[:512]].contiguous()
    centres[0, 0] += 100.0
    for tag, r, ns in (("a", 0.2, 64), ("b", 0.05, 16)):
        bq = ext.ball_query(centres.to(dev), xyz.to(dev), r, ns).cpu().numpy()
        out["bq_%s_idx" % tag], out["bq_%s_r" % tag], out["bq_%s_ns" % tag] = bq, np.float32(r), np.int32(ns)
    out["bq_xyz"], out["bq_centres"] = xyz.numpy(), centres.numpy()
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(1, 6, 20000, generator=g)
    bq_t = torch.from_numpy(out["bq_a_idx"]).to(dev)
    out["grp_feats"] = feats.numpy()
    out["grp_out"] = ext.group_points(feats.to(dev), bq_t).cpu().numpy()
    out["gat_out"] = ext.gather_points(feats.to(dev), bq_t[:, :, 0].contiguous()).cpu().numpy()
    # three_nn / three_interpolate
    unknown, known = scene(3000, 11)[None], scene(1500, 12)[None]
    d2, idx3 = ext.three_nn(unknown.to(dev), known.to(dev))
    w = torch.rand(1, 3000, 3, generator=g)
    f3 = torch.randn(1, 6, 1500, generator=g)
    out["tnn_unknown"], out["tnn_known"] = unknown.numpy(), known.numpy()
    out["tnn_d2"], out["tnn_idx"] = d2.cpu().numpy(), idx3.cpu().numpy()
    out["ti_feats"], out["ti_w"] = f3.numpy(), w.numpy()
    out["ti_out"] = ext.three_interpolate(f3.to(dev), idx3, w.to(dev)).cpu().numpy()
    d2s, idxs = ext.three_nn(unknown[:, :10].contiguous().to(dev), known[:, :2].contiguous().to(dev))
    out["tnn_small_d2"], out["tnn_small_idx"] = d2s.cpu().numpy(), idxs.cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_ext_golden.npz")
