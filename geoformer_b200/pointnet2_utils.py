"""Python surface of lib/pointnet2/pointnet2_utils.py on top of the sm_100a operators.

Same public names and call signatures as the reference (furthest_point_sample, gather_operation,
three_nn, three_interpolate, grouping_operation, ball_query, QueryAndGroup, GroupAll), so the
callers in lib/pointnet2/pointnet2_modules.py:200-226 and model/geoformer/geoformer_fs.py:619-660
work unchanged.  Only the operators differ: they come from geoformer_b200.pointnet2._ext.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from .pointnet2 import _ext


class FurthestPointSampling(Function):
    """pointnet2_utils.py:40-68: xyz (B,N,3) -> (B,npoint) int32, non-differentiable."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """pointnet2_utils.py:71-104: features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features, idx):
        _, C, N = features.size()
        ctx.for_backwards = (idx, C, N)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """pointnet2_utils.py:107-135: returns (sqrt distances (B,n,3), idx (B,n,3))."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """pointnet2_utils.py:138-187."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        m = features.size(2)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:190-236: features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features, idx):
        N = features.size(2)
        ctx.for_backwards = (idx, N)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """pointnet2_utils.py:239-269.  Note the argument order (radius, nsample, xyz, new_xyz)."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def _resample_rows_uniformly(idx, nsample):
    """In-place variant of pointnet2_utils.py:320-329 (not used by GeoFormer): every (batch, region)
    row keeps its distinct indices and tops itself up to `nsample` by drawing among them with
    replacement.  Returns the (B, npoint) count of distinct indices (float CPU tensor, as there)."""
    counts = torch.zeros(idx.shape[:2])
    for row, cnt in zip(idx.view(-1, idx.shape[-1]), counts.view(-1)):
        distinct = torch.unique(row)
        cnt.fill_(distinct.numel())
        extra = torch.randint(0, distinct.numel(), (nsample - distinct.numel(),), dtype=torch.long)
        row.copy_(torch.cat((distinct, distinct[extra])))
    return counts


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:272-356: ball query -> group xyz (centred, optionally /radius) -> group
    features -> concat (B, 3+C, npoint, nsample)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if self.sample_uniformly:
            unique_cnt = _resample_rows_uniformly(idx, self.nsample)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz /= self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """pointnet2_utils.py:359-401."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


def group_points(xyz, features, grouper, npoint, inds=None):
    """PointnetSAModuleVotesSeparate.group_points (pointnet2_modules.py:200-226) as a function:
    FPS (unless `inds` is given) -> gather centres -> grouper.  Returns
    (new_xyz (B,npoint,3), grouped_features, grouped_xyz, inds (B,npoint) i32)."""
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    if inds is None:
        inds = furthest_point_sample(xyz, npoint)
    else:
        assert inds.shape[1] == npoint
    new_xyz = gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous()
    grouped_features, grouped_xyz = grouper(xyz, new_xyz, features)
    return new_xyz, grouped_features, grouped_xyz, inds
