"""GPU tests of the fused set_aggregator (group -> SharedMLP -> pool, gf_group_mlp_pool) against the fixture produced by
the reference's OWN PointnetSAModuleVotesSeparate.mlp on CPU, and end to end (FPS -> gather -> ball query -> fused
kernel) against the CPU restatement at the model's shapes.  fp32 FMA chains in a different order than cuDNN / MKL:
1e-4 relative + 1e-5 absolute."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = dict(rtol=1e-4, atol=1e-5)


@pytest.fixture(scope="module")
def dev(cuda_lib):
    return torch.device("cuda:0")


def _fold(layers, eps=1e-5):
    widths = [layers[0]["w"].shape[1]] + [ly["w"].shape[0] for ly in layers]
    Ws = [ly["w"] for ly in layers]
    scs = [ly["gamma"] / torch.sqrt(ly["var"] + eps) for ly in layers]
    shs = [ly["beta"] - ly["mean"] * s for ly, s in zip(layers, scs)]
    return widths, Ws, scs, shs


def test_fused_aggregator_matches_reference_module_fixture(dev):
    from test_oracle_golden import _aggregate_case

    from geoformer_b200.aggregate import group_mlp_pool

    g = np.load(os.path.join(HERE, "golden", "aggregate_golden.npz"))
    for i in range(int(g["n"])):
        xyz, new_xyz, feats, idx, radius, norm, layers = _aggregate_case(g, i)
        widths, Ws, scs, shs = _fold(layers)
        for pooling in ("max", "avg"):
            out = group_mlp_pool(xyz.to(dev), new_xyz.to(dev), feats.to(dev), idx.int().to(dev), radius, norm, True,
                                 widths, Ws, scs, shs, pooling=pooling)
            np.testing.assert_allclose(out.cpu().numpy(), g["c%d_%s" % (i, pooling)], **TOL)


def test_fused_aggregator_model_shapes_end_to_end(oracle_lib, dev):
    """the model's aggregator (geoformer_fs.py:105-113: 2048 centres, radius 0.2, 64 samples, 16 + 3 -> 32 -> 32 -> 32)
    on a 100k-point scene: our FPS / gather / ball query feed the fused kernel; the CPU restatement gets the same
    indices.  Also a torch.nn module built like the reference's goes through aggregate() (weight folding)."""
    from oracle import aggregate as oagg

    from geoformer_b200.aggregate import aggregate, fold_shared_mlp
    from geoformer_b200.scenes import scene

    torch.manual_seed(3)
    N, m, Cf, radius, ns = 100_000, 2048, 16, 0.2, 64
    xyz = scene(N, 1234)[None].contiguous()
    feats = torch.randn(1, Cf, N)

    class Grouper:  # the attributes aggregate() reads from QueryAndGroup
        pass

    class Module:
        pass

    def conv_bn(cin, cout):
        bn = torch.nn.BatchNorm2d(cout)
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
        return torch.nn.Sequential(torch.nn.Conv2d(cin, cout, 1, bias=False), torch.nn.Sequential(bn), torch.nn.ReLU())

    mod = Module()
    mod.npoint, mod.pooling = m, "max"
    mod.grouper = Grouper()
    mod.grouper.radius, mod.grouper.nsample, mod.grouper.normalize_xyz, mod.grouper.use_xyz = radius, ns, True, True
    mod.mlp_module = torch.nn.Sequential(conv_bn(Cf + 3, 32), conv_bn(32, 32), conv_bn(32, 32)).eval()
    new_xyz, out, inds = aggregate(mod, xyz.to(dev), feats.to(dev))
    ref_inds = oracle_lib.furthest_point_sampling(xyz.numpy(), m)
    assert np.array_equal(inds.cpu().numpy(), ref_inds)
    ref_new = xyz[0][torch.from_numpy(ref_inds[0]).long()][None]
    idx = torch.from_numpy(oracle_lib.ball_query(ref_new.numpy(), xyz.numpy(), radius, ns))
    layers = [{"w": l[0].weight.detach().reshape(l[0].out_channels, -1), "gamma": l[1][0].weight.detach(),
               "beta": l[1][0].bias.detach(), "mean": l[1][0].running_mean, "var": l[1][0].running_var}
              for l in mod.mlp_module]
    want = oagg.group_mlp_pool(xyz, ref_new, feats, idx, radius, True, True, layers, pooling="max")
    assert out.shape == (1, 32, m)
    torch.testing.assert_close(out.cpu(), want, **TOL)
    widths, _, _, _ = fold_shared_mlp(mod.mlp_module)
    assert widths == [19, 32, 32, 32]
    _, out_avg, _ = aggregate(mod, xyz.to(dev), feats.to(dev), inds=inds, pooling="avg")
    want_avg = oagg.group_mlp_pool(xyz, ref_new, feats, idx, radius, True, True, layers, pooling="avg")
    torch.testing.assert_close(out_avg.cpu(), want_avg, **TOL)
