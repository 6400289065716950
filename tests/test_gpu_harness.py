"""Config 5: the post-backbone few-shot forward (aggregator -> geodesic -> decoder -> mask head,
geoformer_fs.py:424-596) with a stub backbone and random weights, run twice: with the library's kernels ("b200") and
with the reference's torch formulation of the same pieces ("torch").  The geodesic maps and FPS indices are identical;
the decoder's cross-attention runs TF32 products in the library (2e-3 per layer), so mask logits are compared at
2 % of their scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_few_shot_forward_with_b200_path_matches_torch_formulation(cuda_lib):
    from geoformer_b200.harness import FewShotForward
    from geoformer_b200.scenes import scene

    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    B, N, m = 2, 20000, 16
    net = FewShotForward(m=m, nlayers=2, n_decode_point=512, n_query_points=64).to(dev)
    locs = torch.stack([scene(N, 90 + b, L=(4.0, 3.0, 2.0), nbox=6) for b in range(B)]).to(dev)
    feats = torch.randn(B, N, m, device=dev)
    support = torch.randn(B, 2 * m, device=dev)
    la, ga, ia = net(locs, feats, support, impl="b200", max_step=48, neighbor=16, geo_radius=0.2)
    lb, gb, ib = net(locs, feats, support, impl="torch", max_step=48, neighbor=16, geo_radius=0.2)
    assert torch.equal(ia, ib)
    for b in range(B):
        assert torch.equal(ga[b], gb[b]) and ga[b].shape == (64, N)
        assert (ga[b] >= 0).float().mean() > 0.01  # the propagation reached something
        assert la[b].shape == (64, N) and torch.isfinite(la[b]).all()
        scale = lb[b].abs().max().item()
        err = (la[b] - lb[b]).abs().max().item()
        assert err <= 0.02 * scale + 1e-6, (b, err, scale)
