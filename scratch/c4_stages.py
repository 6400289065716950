import torch, time
from geoformer_b200.scenes import room, scene
from geoformer_b200.pointnet2 import _ext as p2
from geoformer_b200.geodesic_utils import knn_graph
dev = torch.device("cuda", 0)
for name, x, Q in (("c4", room(1_000_000, 4321).to(dev), 512), ("c2", scene(100_000, 1234).to(dev), 256)):
    xb = x[None].contiguous()
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    print(name, "fps ms", t(lambda: p2.furthest_point_sampling(xb, Q)), "fps2048 ms", t(lambda: p2.furthest_point_sampling(xb, 2048)), "knn ms", t(lambda: knn_graph(x, 16)))
