"""Generates tests/golden/fourier_golden.npz (run in the BUILD container only; needs /root/reference).

    python tests/golden/make_golden_fourier.py

The decoder's relative position embedding, geoformer_fs.py:680-712, produced by the reference's OWN
`PositionEmbeddingCoordsSine` (model/pos_embedding.py, imported unmodified; its `util.utils_pc` dependency
imports trimesh, which is not in this image and is not used on this path, so an empty stand-in module is
registered for the import) applied to the (B,Q,C,3) tensor that the reference's lines :680-702 produce.
Those lines live inside a method of the spconv-dependent model class and cannot be imported, so they are
executed here as the line-by-line restatement oracle/bias.py:decoder_relative_pos (bit-exact torch ops).
Small shapes; includes unreachable contexts, a query whose row has no reachable context at all (:693), a
batch of two scenes with different extents, and num_channels = d_pos."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("trimesh", types.ModuleType("trimesh"))

from model.pos_embedding import PositionEmbeddingCoordsSine  # noqa: E402  (the reference's class)

from oracle import bias as obias  # noqa: E402


def main():
    g = torch.Generator().manual_seed(11)
    B, Q, Cn, d_pos = 2, 12, 40, 64
    Ns = [500, 700]
    torch.manual_seed(5)
    pe = PositionEmbeddingCoordsSine(d_pos=d_pos, pos_type="fourier", normalize=True)  # geoformer_fs.py:116
    gauss_B = pe.gauss_B.clone()
    locs = [torch.rand(n, 3, generator=g) * torch.tensor([6.0, 4.0, 2.5]) - torch.tensor([3.0, 2.0, 0.2]) for n in Ns]
    pc_min = torch.stack([x.min(0)[0] for x in locs])
    pc_max = torch.stack([x.max(0)[0] for x in locs])
    inds = torch.stack([torch.randperm(n, generator=g)[:Cn] for n in Ns]).int()
    ctx = torch.stack([locs[b][inds[b].long()] for b in range(B)])
    qry = ctx[:, :Q].contiguous()
    geo = []
    for b in range(B):
        d = torch.rand(Q, Ns[b], generator=g) * 3.0
        d[torch.rand(Q, Ns[b], generator=g) < 0.4] = -1.0
        geo.append(d)
    geo[0][3, :] = -1.0  # a query that reaches nothing: row max < 0 -> global max (:693)
    rel3 = obias.decoder_relative_pos(geo, inds, qry, ctx)  # :680-702
    # :704-712 through the reference's module
    emb = pe(rel3.reshape(B, Q * Cn, -1), input_range=[pc_min, pc_max]).reshape(B, -1, Q, Cn).permute(2, 3, 0, 1)
    mine = obias.decoder_relative_embedding(geo, inds, qry, ctx, gauss_B, [pc_min, pc_max])
    assert torch.equal(mine, emb), "oracle/bias.py restatement differs from the reference module"
    out = os.path.join(HERE, "fourier_golden.npz")
    np.savez_compressed(out, gauss_B=gauss_B.numpy(), pc_min=pc_min.numpy(), pc_max=pc_max.numpy(),
                        inds=inds.numpy(), ctx=ctx.numpy(), qry=qry.numpy(), geo0=geo[0].numpy(), geo1=geo[1].numpy(),
                        emb=emb.contiguous().numpy())
    print("fourier_golden.npz  B=%d Q=%d C=%d d_pos=%d  (%d bytes)  oracle restatement == reference module: exact"
          % (B, Q, Cn, d_pos, os.path.getsize(out)))


if __name__ == "__main__":
    main()
