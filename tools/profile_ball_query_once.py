import sys, torch
sys.path.insert(0, ".")
from geoformer_b200.scenes import scene
from geoformer_b200.pointnet2 import _ext as p2
dev = torch.device("cuda:0")
x = scene(100_000, 1234)[None].to(dev).contiguous()
inds = p2.furthest_point_sampling(x, 2048)
new_xyz = p2.gather_points(x.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
for _ in range(3): p2.ball_query(new_xyz, x, 0.2, 64)
torch.cuda.synchronize()
