"""CPU restatement of the two distance -> bias epilogues -- TEST INFRASTRUCTURE ONLY.

Plain torch ops on CPU tensors, following the reference line by line:
  decoder_relative_pos      : model/geoformer/geoformer_fs.py:680-702
  mask_head_relative_coords : model/geoformer/geoformer_fs.py:263-289
(only float32 add / sub / abs / max / sqrt / sign: every op is correctly rounded, so the CUDA
kernels are expected to match bit for bit; the tests allow 1e-6 relative as SURVEY 8(c) states).
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


def mask_head_relative_coords(geo_dist, coords, fps_sampling_coords):
    rel = fps_sampling_coords[:, None, :] - coords[None, :, :]  # :271
    Q, N = geo_dist.shape[:2]
    m = torch.max(geo_dist, dim=1)[0]  # :274-275
    M = torch.max(m)
    m = m.clone()
    m[m < 0] = M  # :276
    m = torch.from_numpy(np.sqrt(m.numpy()))  # :277 (numpy's sqrt is correctly rounded, like CUDA's)
    m = m[:, None, None].expand(Q, N, 3)
    cond = (geo_dist < 0).unsqueeze(-1).expand(Q, N, 3)
    rel = rel.clone()
    rel[cond] = rel[cond] + m[cond] * torch.sign(rel[cond])  # :284-286
    return rel.permute(0, 2, 1).contiguous()  # :288
