"""Generates tests/golden/aggregate_golden.npz (run in the BUILD container only; needs /root/reference).

    python tests/golden/make_golden_aggregate.py

The reference's OWN PointnetSAModuleVotesSeparate (lib/pointnet2/pointnet2_modules.py, imported unmodified on top of
this repo's pointnet2._ext stand-in, which is never called here) with the model's construction
(geoformer_fs.py:105-113: mlp [16,32,32,32], radius 0.2, nsample 64, normalize_xyz), random weights and running
statistics, eval mode.  Its `.mlp(grouped_features, grouped_xyz, pooling)` runs on CPU; the grouped tensors it is fed
are built with the statements of its own QueryAndGroup.forward (pointnet2_utils.py:330-341) on CPU gathers, from
ball-query indices of the C oracle."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import geoformer_b200.pointnet2 as p2  # noqa: E402

sys.modules["pointnet2"], sys.modules["pointnet2._ext"] = p2, p2._ext
sys.path.insert(0, "/root/reference")
from lib.pointnet2.pointnet2_modules import PointnetSAModuleVotesSeparate  # noqa: E402

import oracle  # noqa: E402
from geoformer_b200.scenes import scene  # noqa: E402


def main():
    out = {}
    cases = [(1, 6000, 96, 16, 0.2, 64, True), (2, 3000, 40, 5, 0.35, 16, False)]
    for i, (B, N, m, Cf, radius, ns, norm) in enumerate(cases):
        torch.manual_seed(300 + i)
        mod = PointnetSAModuleVotesSeparate(radius=radius, nsample=ns, npoint=m, mlp=[Cf, 32, 32, 32],
                                            normalize_xyz=norm).eval()
        for layer in mod.mlp_module:  # non-trivial running statistics and affine parameters
            bn = layer[1][0]
            bn.running_mean.normal_(0, 0.3)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.data.uniform_(0.5, 1.5)
            bn.bias.data.normal_(0, 0.2)
        xyz = torch.stack([scene(N, 400 + i * 10 + b, L=(3.0, 2.0, 1.5), nbox=4) for b in range(B)])
        feats = torch.randn(B, Cf, N)
        inds = oracle.furthest_point_sampling(xyz.numpy(), m)
        new_xyz = torch.stack([xyz[b][torch.from_numpy(inds[b]).long()] for b in range(B)])
        idx = torch.from_numpy(oracle.ball_query(new_xyz.numpy(), xyz.numpy(), radius, ns))
        ii = idx.long()
        # QueryAndGroup.forward after the ball query, pointnet2_utils.py:330-341, on CPU gathers
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = torch.stack([xyz_trans[b][:, ii[b]] for b in range(B)])
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if norm:
            grouped_xyz /= radius
        grouped_features = torch.stack([feats[b][:, ii[b]] for b in range(B)])
        new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
        with torch.no_grad():
            o_max = mod.mlp(new_features, grouped_xyz, pooling="max")
            o_avg = mod.mlp(new_features, grouped_xyz, pooling="avg")
        out.update({"c%d_xyz" % i: xyz.numpy(), "c%d_feats" % i: feats.numpy(), "c%d_inds" % i: inds,
                    "c%d_idx" % i: idx.numpy(), "c%d_max" % i: o_max.numpy(), "c%d_avg" % i: o_avg.numpy(),
                    "c%d_meta" % i: np.array([radius, ns, float(norm)], dtype=np.float64)})
        for l, layer in enumerate(mod.mlp_module):
            conv, bn = layer[0], layer[1][0]
            out["c%d_w%d" % (i, l)] = conv.weight.detach().reshape(conv.out_channels, -1).numpy()
            out["c%d_gamma%d" % (i, l)] = bn.weight.detach().numpy()
            out["c%d_beta%d" % (i, l)] = bn.bias.detach().numpy()
            out["c%d_mean%d" % (i, l)] = bn.running_mean.numpy()
            out["c%d_var%d" % (i, l)] = bn.running_var.numpy()
    out["n"] = np.array(len(cases))
    path = os.path.join(HERE, "aggregate_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
