"""Three batched hot-path calls of a workload, for `ncu -k regex:geo_bfs_batch --set full --import-source on`."""
import sys
import torch
sys.path.insert(0, ".")
from geoformer_b200.guidance import BatchGuidanceRunner
from geoformer_b200.scenes import CONFIGS, scene, room
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = CONFIGS[wl]
dev = torch.device("cuda:0")
gen = scene if cfg["gen"] == "scene" else room
r = BatchGuidanceRunner(cfg["n"], B, cfg["Q"], cfg["k"], cfg["radius"], cfg["max_step"], device=dev)
r.load([gen(cfg["n"], cfg["seed"] + s).to(dev) for s in range(B)])
for _ in range(3):
    r.replay()
    torch.cuda.synchronize()
