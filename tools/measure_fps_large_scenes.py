import sys, torch
sys.path.insert(0, ".")
from geoformer_b200.scenes import room
from geoformer_b200.pointnet2 import _ext as p2
dev = torch.device("cuda:0")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for N, m in ((1_000_000, 512), (250_000, 2048), (400_000, 512), (1_300_000, 256)):
    x = room(N, 4321)[None].to(dev).contiguous()
    ms = t(lambda: p2.furthest_point_sampling(x, m))
    print("N=%d m=%d  %.4f ms  %.3f us/round" % (N, m, ms, 1e3 * ms / (m - 1)))
