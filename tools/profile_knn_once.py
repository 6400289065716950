import sys, torch
sys.path.insert(0, ".")
from geoformer_b200.scenes import scene
from geoformer_b200.geodesic_utils import knn_graph
x = scene(100_000, 1234).to("cuda:0")
for _ in range(3): knn_graph(x, 16, index_dtype=torch.int32)
torch.cuda.synchronize()
