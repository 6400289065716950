import sys, time, ctypes, torch
sys.path.insert(0, ".")
from geoformer_b200 import _capi as C
from geoformer_b200.guidance import GuidanceRunner
from geoformer_b200.scenes import scene, CONFIGS
dev = torch.device("cuda:0")
for name in ("c1", "c2"):
    cfg = CONFIGS[name]; N, Q, k = cfg["n"], cfg["Q"], cfg["k"]
    xs = [scene(N, 1234 + i).to(dev) for i in range(4)]
    runners = [GuidanceRunner(N, Q, k, cfg["radius"], cfg["max_step"], device=dev) for _ in range(4)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    ref = []
    for r, x, s in zip(runners, xs, streams):
        for _ in range(2): r.run(x, s)
        s.synchronize(); ref.append((r.seeds.clone(), r.geo.clone()))
    graphs = []
    for r, x, s in zip(runners, xs, streams):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            r.run(x, s)
        graphs.append(g)
    torch.cuda.synchronize()
    for r in runners: r.geo.zero_(); r.seeds.zero_()
    for g, s in zip(graphs, streams):
        with torch.cuda.stream(s): g.replay()
    torch.cuda.synchronize()
    ok = all(torch.equal(r.seeds, a) and torch.equal(r.geo, b) for r, (a, b) in zip(runners, ref))
    print(name, "graph replay identical:", ok)
    def bench(fn, K=256):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream(dev)
        e0.record(main)
        for s in streams: s.wait_event(e0)
        t0 = time.perf_counter()
        for i in range(K): fn(i)
        host = (time.perf_counter() - t0) * 1e3 / K
        for s in streams:
            d = torch.cuda.Event(); d.record(s); main.wait_event(d)
        e1.record(main); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K, host
    def plain(i): runners[i % 4].run(xs[i % 4], streams[i % 4])
    def graphed(i):
        with torch.cuda.stream(streams[i % 4]): graphs[i % 4].replay()
    for nm, fn in (("plain", plain), ("graph", graphed), ("plain", plain), ("graph", graphed)):
        ms, host = bench(fn)
        print(name, nm, "ms/step %.4f  host enqueue %.4f  maps/s %.0f" % (ms, host, Q / ms * 1e3))
