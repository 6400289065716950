"""Generates tests/golden/bias_golden.npz (run in the BUILD container only; needs /root/reference).

    python tests/golden/make_golden_bias.py

Inputs and outputs of the reference's OWN two epilogues, executed from model/geoformer/geoformer_fs.py by
oracle/ref_bias.py (mask_heads_forward :263-300 and the statements :680-702 of forward_decoder).  The cases cover
unreachable entries, rows with no reachable entry at all (:276 / :693 take the global maximum), a matrix with no
reachable entry anywhere (the fill value is then -1: sqrt(-1) = NaN in the mask head), zero differences
(sign(0) = 0), a batch of two scenes."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_bias  # noqa: E402


def geo_like(gen, Q, N, frac_unreached, empty_rows=()):
    g = torch.rand(Q, N, generator=gen) * 5.0
    g[torch.rand(Q, N, generator=gen) < frac_unreached] = -1.0
    for r in empty_rows:
        g[r] = -1.0
    return g


def main():
    assert ref_bias.available(), "needs /root/reference"
    mask_fn, mask_lines = ref_bias.load_mask_heads_forward()
    dec_fn, dec_lines = ref_bias.load_decoder_relative_pos()
    gen = torch.Generator().manual_seed(23)
    out = {"mask_lines": np.array(mask_lines), "dec_lines": np.array(dec_lines)}
    # ---- mask head: (Q,N) -> (Q,3,N)
    cases = [(7, 301, 0.6, (2,)), (5, 64, 0.0, ()), (4, 50, 1.0, (0, 1, 2, 3)), (3, 1, 0.5, ())]
    for i, (Q, N, fr, empty) in enumerate(cases):
        geo = geo_like(gen, Q, N, fr, empty)
        coords = torch.randn(N, 3, generator=gen) * 2.0
        seeds_xyz = coords[torch.randint(0, N, (Q,), generator=gen)].clone()  # a seed is one of the points: exact zeros
        res = mask_fn(geo, coords, seeds_xyz)
        out.update({"m%d_geo" % i: geo.numpy(), "m%d_coords" % i: coords.numpy(), "m%d_seed_xyz" % i: seeds_xyz.numpy(),
                    "m%d_out" % i: res.numpy()})
    out["n_mask"] = np.array(len(cases))
    # ---- decoder: list of (Q,N_b), (B,C) indices -> (B,Q,C,3)
    dcases = [(2, 6, 40, (300, 450), 0.5, (1,)), (1, 9, 33, (200,), 0.9, ()), (1, 4, 16, (64,), 1.0, (0, 1, 2, 3))]
    for i, (B, Q, Cn, Ns, fr, empty) in enumerate(dcases):
        geos = [geo_like(gen, Q, n, fr, empty) for n in Ns]
        inds = torch.stack([torch.randperm(n, generator=gen)[:Cn] for n in Ns]).int()
        ctx = torch.randn(B, Cn, 3, generator=gen)
        qry = ctx[:, :Q].clone()
        res = dec_fn(geos, inds, qry, ctx)
        for b, g in enumerate(geos):
            out["d%d_geo%d" % (i, b)] = g.numpy()
        out.update({"d%d_inds" % i: inds.numpy(), "d%d_ctx" % i: ctx.numpy(), "d%d_qry" % i: qry.numpy(),
                    "d%d_out" % i: res.numpy(), "d%d_B" % i: np.array(B)})
    out["n_dec"] = np.array(len(dcases))
    path = os.path.join(HERE, "bias_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "mask_heads_forward lines", mask_lines, "forward_decoder lines", dec_lines)


if __name__ == "__main__":
    main()
