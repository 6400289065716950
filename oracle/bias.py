"""CPU restatement of the two distance -> bias epilogues -- TEST INFRASTRUCTURE ONLY.

Plain torch ops on CPU tensors, following the reference line by line:
  decoder_relative_pos      : model/geoformer/geoformer_fs.py:680-702
  mask_head_relative_coords : model/geoformer/geoformer_fs.py:263-289
  decoder_relative_embedding: model/geoformer/geoformer_fs.py:680-712 with model/pos_embedding.py:88-114 and
                              util/utils_pc.py:35-61 (pinned by tests/golden/fourier_golden.npz, which was produced
                              by the reference's own PositionEmbeddingCoordsSine)
(only float32 add / sub / abs / max / sqrt / sign: every op is correctly rounded, so the CUDA
kernels are expected to match bit for bit; the tests allow 1e-6 relative as SURVEY 8(c) states).
PINNED: tests/golden/bias_golden.npz holds what the reference's own statements produce -- oracle/ref_bias.py
cuts geoformer_fs.py:263-300 and :680-702 out of the reference source with `ast` and executes them unmodified
(tests/golden/make_golden_bias.py); tests/test_oracle_golden.py asserts these functions == that fixture bit
for bit, and tests/test_gpu_parity.py asserts the CUDA kernels == the fixture.
"""
import numpy as np
import torch


def decoder_relative_pos(geo_dists, pre_enc_inds, query_locs, context_locs):
    B = context_locs.shape[0]
    rel = torch.abs(query_locs[:, :, None, :] - context_locs[:, None, :, :])  # :680-682
    Q, Cn = rel.shape[1], rel.shape[2]
    g = torch.stack([geo_dists[b][:, pre_enc_inds[b].long()] for b in range(B)], dim=0)  # :685-690
    m = torch.max(g, dim=2)[0]  # :691
    M = torch.max(m)  # :692
    m[m < 0] = M  # :693
    m = m[:, :, None, None].expand(B, Q, Cn, 3)
    g = g[:, :, :, None].repeat(1, 1, 1, 3)  # :699
    cond = g < 0
    g[cond] = m[cond] + rel[cond]  # :701-702
    return g


def fourier_embedding(xyz, gauss_B, pc_dims, num_channels=None):
    """PositionEmbeddingCoordsSine.get_fourier_embeddings with normalize=True (pos_embedding.py:88-114) and the
    shift_scale_points it calls (utils_pc.py:35-61, dst_range = [0, 1]).  xyz (B, n, 3) -> (B, d_pos, n)."""
    d_out = gauss_B.shape[1] if num_channels is None else num_channels // 2
    B, n = xyz.shape[0], xyz.shape[1]
    xyz = xyz.clone()
    src_min, src_max = pc_dims
    src_diff = src_max[:, None, :] - src_min[:, None, :]  # utils_pc.py:58
    dst_diff = torch.ones_like(src_diff) - torch.zeros_like(src_diff)  # :59
    xyz = (((xyz - src_min[:, None, :]) * dst_diff) / src_diff) + torch.zeros_like(src_diff)  # :60
    xyz *= 2 * np.pi  # pos_embedding.py:107
    xyz = xyz.float()
    proj = torch.mm(xyz.view(-1, 3), gauss_B[:, :d_out]).view(B, n, d_out)  # :109
    return torch.cat([proj.sin(), proj.cos()], dim=2).permute(0, 2, 1)  # :110-113


def decoder_relative_embedding(geo_dists, pre_enc_inds, query_locs, context_locs, gauss_B, pc_dims):
    """geoformer_fs.py:680-712: relative_embedding_pos, shape (Q, C, B, d_pos)."""
    g = decoder_relative_pos(geo_dists, pre_enc_inds, query_locs, context_locs)
    B, Q, Cn = g.shape[:3]
    e = fourier_embedding(g.reshape(B, Q * Cn, -1), gauss_B, pc_dims).reshape(B, -1, Q, Cn)  # :704-711
    return e.permute(2, 3, 0, 1)  # :712


def mask_head_relative_coords(geo_dist, coords, fps_sampling_coords):
    rel = fps_sampling_coords[:, None, :] - coords[None, :, :]  # :271
    Q, N = geo_dist.shape[:2]
    m = torch.max(geo_dist, dim=1)[0]  # :274-275
    M = torch.max(m)
    m = m.clone()
    m[m < 0] = M  # :276
    with np.errstate(invalid="ignore"):  # nothing reachable anywhere: the reference takes sqrt(-1) = NaN as well
        m = torch.from_numpy(np.sqrt(m.numpy()))  # :277 (numpy's sqrt is correctly rounded, like CUDA's)
    m = m[:, None, None].expand(Q, N, 3)
    cond = (geo_dist < 0).unsqueeze(-1).expand(Q, N, 3)
    rel = rel.clone()
    rel[cond] = rel[cond] + m[cond] * torch.sign(rel[cond])  # :284-286
    return rel.permute(0, 2, 1).contiguous()  # :288
