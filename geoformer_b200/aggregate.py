"""set_aggregator (PointnetSAModuleVotesSeparate, lib/pointnet2/pointnet2_modules.py:150-249) with its grouping and its
SharedMLP fused into one kernel: FPS -> gather -> ball_query on the library's operators, then gf_group_mlp_pool
instead of group_points x2 + three Conv2d/BatchNorm2d/ReLU layers + a pooling op.  Inference form only (batch norm
with running statistics); a module in training mode must use the reference's own `mlp`.
"""
import ctypes

import torch

from . import _capi as C
from .pointnet2 import _ext


def fold_shared_mlp(mlp_module):
    """(widths, weights, scales, shifts) of a reference SharedMLP (pytorch_utils.py:9-32) in eval mode: every layer is
    Conv2d 1x1 [+ BatchNorm2d] [+ ReLU]; batch norm folds into y = scale * (W x) + shift"""
    widths, Ws, scs, shs = [], [], [], []
    for layer in mlp_module:
        conv = bn = None
        for sub in layer:
            if isinstance(sub, torch.nn.Conv2d):
                conv = sub
            elif isinstance(sub, torch.nn.Sequential) or isinstance(sub, torch.nn.BatchNorm2d):
                bn = sub[0] if isinstance(sub, torch.nn.Sequential) else sub
        C.require(conv is not None and conv.kernel_size == (1, 1), "SharedMLP layers must be 1x1 convolutions")
        W = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        if bn is not None:
            C.require(not bn.training, "fused aggregator: batch norm must be in eval mode (running statistics)")
            scale = (bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)).float()
            shift = (bn.bias.detach() - bn.running_mean.detach() * scale).float()
        else:
            scale = torch.ones(conv.out_channels, device=W.device)
            shift = torch.zeros(conv.out_channels, device=W.device)
        if conv.bias is not None:
            shift = shift + conv.bias.detach().float() * scale
        if not widths:
            widths.append(conv.in_channels)
        widths.append(conv.out_channels)
        Ws.append(W.contiguous()), scs.append(scale.contiguous()), shs.append(shift.contiguous())
    return widths, Ws, scs, shs


def group_mlp_pool(xyz, new_xyz, features, idx, radius, normalize_xyz, use_xyz, widths, weights, scales, shifts,
                   pooling="max"):
    """xyz (B,N,3), new_xyz (B,m,3), features (B,C,N) or None, idx (B,m,nsample) i32 -> (B, widths[-1], m)"""
    C.check_cuda_f32(xyz, "xyz")
    C.check_cuda_f32(new_xyz, "new_xyz")
    C.check_cuda_i32(idx, "idx")
    C.require(pooling in ("max", "avg"), "pooling must be 'max' or 'avg' (rbf pooling is not fused)")
    B, N, _ = xyz.shape
    m, ns = idx.shape[1], idx.shape[2]
    Cn = 0
    if features is not None:
        C.check_cuda_f32(features, "features")
        C.require(features.shape[0] == B and features.shape[2] == N, "features must be (B, C, N)")
        Cn = features.shape[1]
    dev = xyz.device
    L = len(weights)
    ws = [w.to(device=dev, dtype=torch.float32).contiguous() for w in weights]
    sc = [s.to(device=dev, dtype=torch.float32).contiguous() for s in scales]
    sh = [s.to(device=dev, dtype=torch.float32).contiguous() for s in shifts]
    out = torch.empty((B, widths[-1], m), dtype=torch.float32, device=dev)
    arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])  # noqa: E731
    with torch.cuda.device(dev):
        C.check(C.lib().gf_group_mlp_pool(C.ptr(xyz), C.ptr(new_xyz), C.ptr(features), C.ptr(idx), B, N, m, ns, Cn,
                                          ctypes.c_float(float(radius)), 1 if normalize_xyz else 0, 1 if use_xyz else 0,
                                          L, (ctypes.c_int * (L + 1))(*widths), arr(ws), arr(sc), arr(sh),
                                          1 if pooling == "avg" else 0, C.ptr(out), C.stream_of(dev)), "group_mlp_pool")
    return out


def aggregate(module, xyz, features, inds=None, npoint_new=None, pooling=None):
    """module.group_points(...) followed by module.mlp(...) of a reference PointnetSAModuleVotesSeparate (the calls of
    geoformer_fs.py:643-657 / :413-416), fused: returns (new_xyz (B,m,3), new_features (B,C_out,m), inds (B,m) i32).
    xyz (B,N,3), features (B,C,N)."""
    npoint = npoint_new if npoint_new is not None else module.npoint
    C.require(npoint is not None, "GroupAll aggregators (npoint=None) are not fused")
    if inds is None:
        inds = _ext.furthest_point_sampling(xyz.contiguous(), int(npoint))
    else:
        C.require(inds.shape[1] == npoint, "inds must have npoint columns")
    new_xyz = _ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    g = module.grouper
    idx = _ext.ball_query(new_xyz, xyz.contiguous(), float(g.radius), int(g.nsample))
    widths, Ws, scs, shs = fold_shared_mlp(module.mlp_module)
    out = group_mlp_pool(xyz.contiguous(), new_xyz, features.contiguous() if features is not None else None, idx,
                         g.radius, g.normalize_xyz, g.use_xyz, widths, Ws, scs, shs,
                         pooling=pooling if pooling is not None else module.pooling)
    return new_xyz, out, inds
