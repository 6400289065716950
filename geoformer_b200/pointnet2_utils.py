"""The Python surface GeoFormer expects from `pointnet2_utils` (lib/pointnet2/pointnet2_utils.py in the
reference), rebuilt on the sm_100a operators of `geoformer_b200.pointnet2._ext`.

Public names and call signatures are the reference's -- furthest_point_sample, gather_operation, three_nn,
three_interpolate, grouping_operation, ball_query, QueryAndGroup, GroupAll -- so that
lib/pointnet2/pointnet2_modules.py:200-226 and model/geoformer/geoformer_fs.py:619-660 run unchanged.
The implementation is not: the six autograd wrappers are stamped out by one small factory from a table
(forward operator, which outputs are index tensors, how the gradient is routed), and the two grouping
modules share their feature-assembly code.
"""
import torch
import torch.nn as nn

from .pointnet2 import _ext


def _autograd_op(name, doc, n_inputs, run, grad=None, index_outputs=()):
    """Builds a torch.autograd.Function subclass and returns its `.apply`.

    run(*inputs)            -> (outputs, saved) with outputs a tensor or tuple and `saved` anything the
                               gradient needs;
    grad(saved, *grad_outs) -> gradient of input 0 (only input 0 is ever differentiable here), or None;
    index_outputs           -> positions of integer outputs (marked non-differentiable)."""

    def forward(ctx, *inputs):
        outputs, ctx.saved_for_grad = run(*inputs)
        many = isinstance(outputs, tuple)
        for pos in index_outputs:
            ctx.mark_non_differentiable(outputs[pos] if many else outputs)
        return outputs

    def backward(ctx, *grad_outputs):
        first = None if grad is None else grad(ctx.saved_for_grad, *grad_outputs)
        return (first,) + (None,) * (n_inputs - 1)

    cls = type(name, (torch.autograd.Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward),
                                                  "__doc__": doc})
    return cls, cls.apply


# ---- sampling / gathering ----------------------------------------------------------------------------
FurthestPointSampling, furthest_point_sample = _autograd_op(
    "FurthestPointSampling", "(xyz (B,N,3), npoint) -> (B,npoint) int32 sample indices; reference :40-68", 2,
    run=lambda xyz, npoint: (_ext.furthest_point_sampling(xyz, npoint), None), index_outputs=(0,))

GatherOperation, gather_operation = _autograd_op(
    "GatherOperation", "(features (B,C,N), idx (B,npoint)) -> (B,C,npoint); reference :71-104", 2,
    run=lambda feats, idx: (_ext.gather_points(feats, idx), (idx, feats.size(2))),
    grad=lambda saved, g: _ext.gather_points_grad(g.contiguous(), saved[0], saved[1]))

# ---- three nearest neighbours and the interpolation built on them ---------------------------------------


def _three_nn_run(unknown, known):
    d2, idx = _ext.three_nn(unknown, known)
    return (torch.sqrt(d2), idx), None  # the reference returns distances, the kernel their squares (:126-128)


ThreeNN, three_nn = _autograd_op(
    "ThreeNN", "(unknown (B,n,3), known (B,m,3)) -> (dist (B,n,3), idx (B,n,3)); reference :107-135", 2,
    run=_three_nn_run, index_outputs=(1,))

ThreeInterpolate, three_interpolate = _autograd_op(
    "ThreeInterpolate", "(features (B,C,m), idx (B,n,3), weight (B,n,3)) -> (B,C,n); reference :138-187", 3,
    run=lambda feats, idx, w: (_ext.three_interpolate(feats, idx, w), (idx, w, feats.size(2))),
    grad=lambda saved, g: _ext.three_interpolate_grad(g.contiguous(), saved[0], saved[1], saved[2]))

# ---- neighbourhood queries ----------------------------------------------------------------------------------
GroupingOperation, grouping_operation = _autograd_op(
    "GroupingOperation", "(features (B,C,N), idx (B,npoint,nsample)) -> (B,C,npoint,nsample); reference :190-236", 2,
    run=lambda feats, idx: (_ext.group_points(feats, idx), (idx, feats.size(2))),
    grad=lambda saved, g: _ext.group_points_grad(g.contiguous(), saved[0], saved[1]))

# argument order of the reference: (radius, nsample, xyz, new_xyz) -- the operator takes the centres first
BallQuery, ball_query = _autograd_op(
    "BallQuery", "(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) int32; reference :239-269",
    4, run=lambda radius, nsample, xyz, centres: (_ext.ball_query(centres, xyz, radius, nsample), None),
    index_outputs=(0,))


def _resample_rows_uniformly(idx, nsample):
    """`sample_uniformly` of the reference's QueryAndGroup (:320-329; GeoFormer never turns it on): each
    (batch, region) row is rewritten in place as its distinct indices followed by draws, with replacement,
    among them.  Returns how many distinct indices every row had, as a (B, npoint) float CPU tensor."""
    flat = idx.view(-1, idx.shape[-1])
    distinct_counts = torch.zeros(idx.shape[:2])
    flat_counts = distinct_counts.view(-1)
    for r in range(flat.shape[0]):
        members = torch.unique(flat[r])
        flat_counts[r] = members.numel()
        refill = torch.randint(0, members.numel(), (nsample - members.numel(),), dtype=torch.long)
        flat[r] = torch.cat((members, members[refill]))
    return distinct_counts


def _assemble(local_xyz, grouped_feats, use_xyz):
    """what both groupers return as features: coordinates first (when asked for), then the gathered channels"""
    if grouped_feats is None:
        assert use_xyz, "Cannot have not features and not use xyz as a feature!"
        return local_xyz
    return torch.cat([local_xyz, grouped_feats], dim=1) if use_xyz else grouped_feats


class QueryAndGroup(nn.Module):
    """Ball query around every centre, neighbours' coordinates made relative to it (and divided by the radius
    when `normalize_xyz`), neighbours' features gathered: (B, 3+C, npoint, nsample).  Reference :272-356."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        assert sample_uniformly or not ret_unique_cnt  # counts only exist when rows are resampled
        self.radius = radius
        self.nsample = nsample
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt

    def forward(self, xyz, new_xyz, features=None):
        members = ball_query(self.radius, self.nsample, xyz, new_xyz)
        counts = _resample_rows_uniformly(members, self.nsample) if self.sample_uniformly else None
        local = grouping_operation(xyz.transpose(1, 2).contiguous(), members)  # (B,3,npoint,nsample)
        local -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            local /= self.radius
        gathered = None if features is None else grouping_operation(features, members)
        outputs = [_assemble(local, gathered, self.use_xyz)]
        if self.ret_grouped_xyz:
            outputs.append(local)
        if self.ret_unique_cnt:
            outputs.append(counts)
        return outputs[0] if len(outputs) == 1 else tuple(outputs)


class GroupAll(nn.Module):
    """One group holding every point: (B, 3+C, 1, N).  Reference :359-401."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        everything = xyz.transpose(1, 2).unsqueeze(2)
        merged = everything if features is None else _assemble(everything, features.unsqueeze(2), self.use_xyz)
        return (merged, everything) if self.ret_grouped_xyz else merged


def group_points(xyz, features, grouper, npoint, inds=None):
    """The set-aggregation front end of GeoFormer (PointnetSAModuleVotesSeparate.group_points,
    pointnet2_modules.py:200-226) as a function: centres by FPS unless their indices are handed in, then the
    grouper.  -> (centres (B,npoint,3), grouped features, grouped coordinates, centre indices (B,npoint) i32)."""
    if inds is None:
        inds = furthest_point_sample(xyz, npoint)
    assert inds.shape[1] == npoint
    centres = gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    grouped_feats, grouped_xyz = grouper(xyz, centres, features)
    return centres, grouped_feats, grouped_xyz, inds
