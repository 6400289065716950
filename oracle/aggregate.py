"""CPU restatement of set_aggregator's grouping + SharedMLP + pooling -- TEST INFRASTRUCTURE ONLY.
Plain fp32 torch ops following lib/pointnet2/pointnet2_utils.py:330-341 (QueryAndGroup after the ball query),
pytorch_utils.py:59-104 (1x1 Conv2d + BatchNorm2d in eval mode + ReLU) and pointnet2_modules.py:228-249 (.mlp).
PINNED by tests/golden/aggregate_golden.npz: outputs of the reference's OWN PointnetSAModuleVotesSeparate.mlp on CPU,
fed with grouped tensors built by its own QueryAndGroup arithmetic (tests/golden/make_golden_aggregate.py)."""
import torch
import torch.nn.functional as F


def group_mlp_pool(xyz, new_xyz, features, idx, radius, normalize_xyz, use_xyz, layers, pooling="max", eps=1e-5):
    """xyz (B,N,3), new_xyz (B,m,3), features (B,C,N) or None, idx (B,m,ns) int; layers: list of dicts with conv weight
    `w` (out,in), batch-norm `gamma`, `beta`, `mean`, `var` -> (B, C_out, m)"""
    B, m, ns = idx.shape
    ii = idx.long()
    xyz_t = xyz.transpose(1, 2).contiguous()  # (B,3,N)
    g_xyz = torch.stack([xyz_t[b][:, ii[b]] for b in range(B)])  # grouping_operation: (B,3,m,ns)
    g_xyz = g_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)  # :333
    if normalize_xyz:
        g_xyz = g_xyz / radius  # :335
    if features is not None:
        g_f = torch.stack([features[b][:, ii[b]] for b in range(B)])
        x = torch.cat([g_xyz, g_f], dim=1) if use_xyz else g_f  # :339-341
    else:
        x = g_xyz
    for ly in layers:
        x = F.conv2d(x, ly["w"][:, :, None, None])
        x = F.batch_norm(x, ly["mean"], ly["var"], ly["gamma"], ly["beta"], training=False, eps=eps)
        x = F.relu(x)
    if pooling == "max":
        x = F.max_pool2d(x, kernel_size=[1, x.size(3)])
    else:
        x = F.avg_pool2d(x, kernel_size=[1, x.size(3)])
    return x.squeeze(-1)
