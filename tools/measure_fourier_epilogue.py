import sys, time, torch
sys.path.insert(0, ".")
from geoformer_b200.scenes import scene
from geoformer_b200.bias import decoder_relative_embedding, decoder_relative_pos
from geoformer_b200.pointnet2 import _ext as p2
from geoformer_b200.guidance import geodesic_guidance
dev = torch.device("cuda:0")
x = scene(100_000, 1234).to(dev)
Q, Cn = 256, 2048
ctx = p2.furthest_point_sampling(x[None].contiguous(), Cn)
seeds, geo = geodesic_guidance(x, Q, 16, 0.5, 32)
ctx_xyz = x[ctx[0].long()][None].contiguous(); q_xyz = ctx_xyz[:, :Q].contiguous()
gb = torch.randn(3, 32, device=dev); pc = [x.min(0)[0][None].contiguous(), x.max(0)[0][None].contiguous()]
def torch_path():
    rel = decoder_relative_pos([geo], ctx, q_xyz, ctx_xyz).reshape(1, Q * Cn, 3)
    v = rel.clone()
    v = (((v - pc[0][:, None, :]) * 1.0) / (pc[1] - pc[0])[:, None, :]) + 0.0
    v *= 6.283185307179586
    pr = torch.mm(v.view(-1, 3), gb).view(1, Q * Cn, 32)
    return torch.cat([pr.sin(), pr.cos()], dim=2).permute(0, 2, 1).reshape(1, -1, Q, Cn).permute(2, 3, 0, 1)
for name, fn in (("fused", lambda: decoder_relative_embedding([geo], ctx, q_xyz, ctx_xyz, gb, pc)), ("torch ops after our (B,Q,C,3) kernel", torch_path)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(name, "gpu ms/call %.4f" % (a.elapsed_time(b) / 50), "host enqueue ms/call %.4f" % ((t1 - t0) * 1e3 / 50))
o1 = decoder_relative_embedding([geo], ctx, q_xyz, ctx_xyz, gb, pc); o2 = torch_path()
print("max abs diff vs torch CUDA ops", float((o1 - o2).abs().max()))
