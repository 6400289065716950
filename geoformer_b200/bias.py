"""The two distance -> bias epilogues that consume the geodesic maps, as fused kernels.

decoder_relative_pos   : geoformer_fs.py:680-702 (= geoformer.py:619-641)
decoder_relative_embedding : the same fused with the Fourier embedding that follows it, geoformer_fs.py:680-712
mask_head_relative_coords : geoformer_fs.py:263-292 (= geoformer.py:286-313)
"""
import ctypes

import torch

from . import _capi as C


def decoder_relative_pos(geo_dists, pre_enc_inds, query_locs, context_locs):
    """geo_dists: list of B tensors (Q, N_b) f32; pre_enc_inds (B,C) i32; query_locs (B,Q,3);
    context_locs (B,C,3)  ->  (B,Q,C,3) f32: the geodesic distance to each context (replicated on
    the 3 channels), unreachable contexts replaced by rowmax + |query - context|."""
    B, Q, _ = query_locs.shape
    Cn = context_locs.size(1)
    C.check_cuda_f32(query_locs, "query_locs")
    C.check_cuda_f32(context_locs, "context_locs")
    C.check_cuda_i32(pre_enc_inds, "pre_enc_inds")
    C.require(len(geo_dists) == B and tuple(pre_enc_inds.shape) == (B, Cn), "batch / context mismatch")
    for g in geo_dists:
        C.check_cuda_f32(g, "geo_dists[b]")
        C.require(g.size(0) == Q, "geo_dists[b] must have Q rows")
    dev = query_locs.device
    out = torch.empty((B, Q, Cn, 3), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * B)(*[g.data_ptr() for g in geo_dists])
    lds = (ctypes.c_int * B)(*[g.size(1) for g in geo_dists])
    L = C.lib()
    with torch.cuda.device(dev):
        nbytes = L.gf_bias_workspace_bytes(B, Q)
        ws = C.workspace.get(dev, "bias", nbytes)
        C.check(L.gf_bias_decoder(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(lds, ctypes.c_void_p),
                                  C.ptr(pre_enc_inds), C.ptr(query_locs), C.ptr(context_locs), B, Q, Cn, C.ptr(out),
                                  C.ptr(ws), nbytes, C.stream_of(dev)), "bias_decoder")
    return out


def decoder_relative_embedding(geo_dists, pre_enc_inds, query_locs, context_locs, gauss_B, pc_dims, num_channels=None):
    """geoformer_fs.py:680-712 in one pass: the decoder's `relative_embedding_pos`.

    gauss_B: the (3, d_pos/2) buffer of the reference's PositionEmbeddingCoordsSine(pos_type="fourier",
    normalize=True); pc_dims = [pc_mins (B,3), pc_maxs (B,3)].  Returns the (Q, C, B, d_pos) tensor the
    reference hands to the decoder (a permuted view of contiguous (B,Q,C,d_pos) memory, as there)."""
    B, Q, _ = query_locs.shape
    Cn = context_locs.size(1)
    C.check_cuda_f32(query_locs, "query_locs")
    C.check_cuda_f32(context_locs, "context_locs")
    C.check_cuda_i32(pre_enc_inds, "pre_enc_inds")
    C.check_cuda_f32(gauss_B, "gauss_B")
    pc_min, pc_max = pc_dims
    C.check_cuda_f32(pc_min, "pc_mins")
    C.check_cuda_f32(pc_max, "pc_maxs")
    C.require(len(geo_dists) == B and tuple(pre_enc_inds.shape) == (B, Cn), "batch / context mismatch")
    C.require(gauss_B.dim() == 2 and gauss_B.size(0) == 3, "gauss_B must be (3, d_pos/2)")
    C.require(tuple(pc_min.shape) == (B, 3) and tuple(pc_max.shape) == (B, 3), "pc_dims must be two (B,3) tensors")
    d_out = gauss_B.size(1) if num_channels is None else int(num_channels) // 2
    C.require(0 < d_out <= gauss_B.size(1), "num_channels out of range")  # pos_embedding.py:95-99
    for g in geo_dists:
        C.check_cuda_f32(g, "geo_dists[b]")
        C.require(g.size(0) == Q, "geo_dists[b] must have Q rows")
    dev = query_locs.device
    out = torch.empty((B, Q, Cn, 2 * d_out), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * B)(*[g.data_ptr() for g in geo_dists])
    lds = (ctypes.c_int * B)(*[g.size(1) for g in geo_dists])
    L = C.lib()
    with torch.cuda.device(dev):
        nbytes = L.gf_bias_workspace_bytes(B, Q)
        ws = C.workspace.get(dev, "bias", nbytes)
        C.check(L.gf_bias_decoder_fourier(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(lds, ctypes.c_void_p),
                                          C.ptr(pre_enc_inds), C.ptr(query_locs), C.ptr(context_locs), B, Q, Cn,
                                          C.ptr(gauss_B), d_out, gauss_B.size(1), C.ptr(pc_min), C.ptr(pc_max),
                                          C.ptr(out), C.ptr(ws), nbytes, C.stream_of(dev)), "bias_decoder_fourier")
    return out.permute(1, 2, 0, 3)


def mask_head_relative_coords(geo_dist, coords, fps_sampling_coords, row_max=None):
    """geo_dist (Q,N), coords (N,3), fps_sampling_coords (Q,3) -> (Q,3,N): seed - point, pushed
    outwards by sqrt(rowmax) along each axis where the point is unreachable from the seed.
    row_max: optional (Q,) row maxima as produced by the propagation (`geodesic_guidance(..., row_max=t)`);
    with it the maps are read once instead of twice."""
    C.check_cuda_f32(geo_dist, "geo_dist")
    C.check_cuda_f32(coords, "coords")
    C.check_cuda_f32(fps_sampling_coords, "fps_sampling_coords")
    Q, N = geo_dist.shape
    C.require(tuple(coords.shape) == (N, 3) and tuple(fps_sampling_coords.shape) == (Q, 3), "shape mismatch")
    if row_max is not None:
        C.check_cuda_f32(row_max, "row_max")
        C.require(tuple(row_max.shape) == (Q,), "row_max must be (Q,)")
    dev = geo_dist.device
    out = torch.empty((Q, 3, N), dtype=torch.float32, device=dev)
    L = C.lib()
    with torch.cuda.device(dev):
        nbytes = L.gf_bias_workspace_bytes(1, max(Q, 1))
        ws = C.workspace.get(dev, "bias", nbytes)
        C.check(L.gf_bias_mask_head(C.ptr(geo_dist), C.ptr(coords), C.ptr(fps_sampling_coords), Q, N, C.ptr(row_max),
                                    C.ptr(out),
                                    C.ptr(ws), nbytes, C.stream_of(dev)), "bias_mask_head")
    return out
