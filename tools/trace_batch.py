import sys, torch
sys.path.insert(0, ".")
from geoformer_b200.guidance import BatchGuidanceRunner
from geoformer_b200.scenes import CONFIGS, scene
cfg = CONFIGS["c2"]; dev = torch.device("cuda:0"); B = 10
xs = [scene(cfg["n"], cfg["seed"] + s).to(dev) for s in range(B)]
r = BatchGuidanceRunner(cfg["n"], B, cfg["Q"], cfg["k"], cfg["radius"], cfg["max_step"], device=dev, graph=False)
for _ in range(4):
    r.run(xs); torch.cuda.synchronize()
