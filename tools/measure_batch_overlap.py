import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from geoformer_b200.scenes import scene
from geoformer_b200.geodesic_utils import cal_geodesic_vectorize
import geoformer_b200.geodesic_utils as gu
from geoformer_b200.pointnet2 import _ext as p2
dev = torch.device("cuda:0")
B, N, Q = 8, 100_000, 256
pts = [scene(N, 1234 + i).to(dev) for i in range(B)]
locs = torch.cat(pts); offs = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=dev)
pre = torch.stack([p2.furthest_point_sampling(p[None].contiguous(), Q)[0] for p in pts])
for lanes in (1, 4):
    gu._MAX_SCENES_IN_FLIGHT = lanes
    for _ in range(3): out = cal_geodesic_vectorize(None, pre, locs, offs, max_step=32, neighbor=16, radius=0.5, n_queries=Q)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): out = cal_geodesic_vectorize(None, pre, locs, offs, max_step=32, neighbor=16, radius=0.5, n_queries=Q)
    b.record(); torch.cuda.synchronize()
    print("scenes in flight %d: %.4f ms per scene (batch of %d, seeds given)" % (lanes, a.elapsed_time(b) / 10 / B, B))
