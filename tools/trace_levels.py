"""Development: per-seed and per-level phase trace of the propagation kernel (needs the -DGF_TRACE build:
python -m geoformer_b200.build --variant=trace -DGF_TRACE; GF_LIB=geoformer_b200/libgeoformer_b200_trace.so)."""
import sys
import torch
sys.path.insert(0, ".")
from geoformer_b200.guidance import GuidanceRunner
from geoformer_b200.scenes import CONFIGS, scene, room
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = CONFIGS[wl]
dev = torch.device("cuda:0")
gen = scene if cfg["gen"] == "scene" else room
x = gen(cfg["n"], cfg["seed"]).to(dev)
r = GuidanceRunner(cfg["n"], cfg["Q"], cfg["k"], cfg["radius"], cfg["max_step"], device=dev)
for _ in range(4):
    r.run(x)
    torch.cuda.synchronize()
