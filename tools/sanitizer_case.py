"""A small pass over the hot path for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitizer_case.py [N of the single-scene call, default 20000]"""
import sys, torch
sys.path.insert(0, ".")
from geoformer_b200.guidance import geodesic_guidance_batch, geodesic_guidance
from geoformer_b200.geodesic_utils import knn_graph
from geoformer_b200.scenes import scene
dev = torch.device("cuda:0")
xs = [scene(n, 7 + i).to(dev) for i, n in enumerate((6000, 9000, 3000))]
out = geodesic_guidance_batch(xs, 40, 16, 0.5, 12)
torch.cuda.synchronize()
x = scene(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 3).to(dev)
s, g = geodesic_guidance(x, 300, 8, 0.5, 16)[:2]
D, I = knn_graph(x, 64, index_dtype=torch.int32)
torch.cuda.synchronize()
print("ok", float(g.sum()), int(I.sum()))
