"""The fused hot path: FPS seeds -> kNN graph -> geodesic maps of one scene.

geodesic_guidance      : device tensors in, device tensors out (asynchronous on the current stream)
HostGuidance           : host buffers in, host buffers out through gf_guidance_host (the call the
                         benchmark's `e2e` number times: H2D of the points and D2H of the maps
                         are inside the call)
"""
import ctypes

import torch

from . import _capi as C


def geodesic_guidance(xyz, n_queries, neighbor, radius, max_step, return_graph=False, return_stats=False,
                      row_max=None, ws_tag="guidance"):
    """xyz (N,3) f32 CUDA -> (seeds (Q,) i32, geo (Q,N) f32 [, D (N,k) f32, I (N,k) i32][, stats (2,) i64]).
    row_max: optional (Q,) f32 CUDA tensor that receives the maximum of every row of geo (for the epilogues)."""
    C.check_cuda_f32(xyz, "xyz")
    C.require(xyz.dim() == 2 and xyz.size(1) == 3, "xyz must be (N, 3)")
    N, Q, k = xyz.size(0), int(n_queries), int(neighbor)
    dev = xyz.device
    seeds = torch.empty((Q,), dtype=torch.int32, device=dev)
    geo = torch.empty((Q, N), dtype=torch.float32, device=dev)
    D = torch.empty((N, k), dtype=torch.float32, device=dev) if return_graph else None
    I = torch.empty((N, k), dtype=torch.int32, device=dev) if return_graph else None
    stats = torch.zeros(2, dtype=torch.int64, device=dev) if return_stats else None
    L = C.lib()
    with torch.cuda.device(dev):
        nbytes = L.gf_guidance_workspace_bytes(N, Q, k)
        ws = C.workspace.get(dev, ws_tag, nbytes)
        C.check(L.gf_guidance(C.ptr(xyz), N, Q, k, ctypes.c_float(float(radius)), int(max_step), C.ptr(seeds),
                              C.ptr(geo), C.ptr(D), C.ptr(I), C.ptr(stats), C.ptr(row_max), C.ptr(ws), nbytes,
                              C.stream_of(dev)), "guidance")
    out = [seeds, geo]
    if return_graph:
        out += [D, I]
    if return_stats:
        out.append(stats)
    return tuple(out)


class GuidanceRunner:
    """Pre-allocated device-resident runner (no allocation per call): run(xyz) -> (seeds, geo) views
    of its own buffers, valid until the next run().  Used by bench.py for the kernel-only number.

    graph=True captures the 17 launches of one call (FPS on its forked stream included) into a CUDA graph at
    the first run() and replays it afterwards: one launch per scene on the host side (0.08 -> 0.013 ms of
    enqueue time at c1, where the 13 tiny kernels of the kNN grid build make the step launch bound: +6 %
    throughput; at c2 the plain path is ~4 % faster because graph kernel nodes lose the high stream priority
    that lets FPS start ahead of other scenes' work).  The points are copied into a fixed staging buffer when
    run() is given a different tensor than the one captured."""

    def __init__(self, N, n_queries, neighbor, radius, max_step, device="cuda", graph=False):
        self.N, self.Q, self.k = int(N), int(n_queries), int(neighbor)
        self.radius, self.max_step = float(radius), int(max_step)
        self.device = torch.device(device)
        self.seeds = torch.empty((self.Q,), dtype=torch.int32, device=self.device)
        self.geo = torch.empty((self.Q, self.N), dtype=torch.float32, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.row_max = torch.empty((self.Q,), dtype=torch.float32, device=self.device)
        self._L = C.lib()
        self._nbytes = self._L.gf_guidance_workspace_bytes(self.N, self.Q, self.k)
        self._ws = torch.empty(self._nbytes, dtype=torch.uint8, device=self.device)
        self._use_graph = bool(graph)
        self._graph = None
        self._xyz_captured = None
        self._capture_stream = None
        self.launches_per_run = None  # kernels inside one replay (the library's counter only sees the capture)

    def _capture(self, xyz):
        self._xyz_captured = xyz
        self._capture_stream = torch.cuda.Stream(device=self.device)
        self._capture_stream.wait_stream(torch.cuda.current_stream(self.device))
        for _ in range(2):  # warm-up on the capture stream: creates its auxiliary stream, sets kernel attributes
            self._launch(xyz, self._capture_stream)
        self._capture_stream.synchronize()
        g = torch.cuda.CUDAGraph()
        before = C.launch_count()
        with torch.cuda.graph(g, stream=self._capture_stream):
            self._launch(xyz, self._capture_stream)
        self.launches_per_run = C.launch_count() - before
        self._graph = g

    def run(self, xyz, stream=None):
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        if not self._use_graph:
            return self._launch(xyz, st)
        if self._graph is None:
            self._capture(xyz)
        with torch.cuda.stream(st):
            if xyz.data_ptr() != self._xyz_captured.data_ptr():
                self._xyz_captured.copy_(xyz, non_blocking=True)
            self._graph.replay()
        return self.seeds, self.geo

    def _launch(self, xyz, st):
        C.check(self._L.gf_guidance(C.ptr(xyz), self.N, self.Q, self.k, ctypes.c_float(self.radius), self.max_step,
                                    C.ptr(self.seeds), C.ptr(self.geo), None, None, C.ptr(self.stats),
                                    C.ptr(self.row_max), C.ptr(self._ws), self._nbytes,
                                    ctypes.c_void_p(st.cuda_stream)), "guidance")
        return self.seeds, self.geo


class HostGuidance:
    """Host-buffer entry point.  run(xyz_host) copies the points to the device, runs the hot path and
    copies seeds and maps back, all inside gf_guidance_host (blocking; releases the GIL, so several
    HostGuidance objects driven from different Python threads overlap copies with compute)."""

    def __init__(self, N, n_queries, neighbor, radius, max_step, device="cuda", pinned=True):
        self.N, self.Q, self.k = int(N), int(n_queries), int(neighbor)
        self.radius, self.max_step = float(radius), int(max_step)
        self.device = torch.device(device)
        self._L = C.lib()
        self._nbytes = self._L.gf_guidance_host_workspace_bytes(self.N, self.Q, self.k)
        self._ws = torch.empty(self._nbytes, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.seeds_host = torch.empty((self.Q,), dtype=torch.int32, pin_memory=pinned)
        self.geo_host = torch.empty((self.Q, self.N), dtype=torch.float32, pin_memory=pinned)
        self.h2d_bytes = self.N * 3 * 4
        self.d2h_bytes = self.Q * 4 + self.Q * self.N * 4

    def run(self, xyz_host):
        C.require(not xyz_host.is_cuda and xyz_host.dtype == torch.float32 and xyz_host.is_contiguous(),
                  "xyz_host must be a contiguous float32 CPU tensor")
        C.require(tuple(xyz_host.shape) == (self.N, 3), "xyz_host must be (N, 3)")
        with torch.cuda.device(self.device):
            C.check(self._L.gf_guidance_host(C.ptr(xyz_host), self.N, self.Q, self.k, ctypes.c_float(self.radius),
                                             self.max_step, C.ptr(self.seeds_host), C.ptr(self.geo_host),
                                             C.ptr(self._ws), self._nbytes,
                                             ctypes.c_void_p(self.stream.cuda_stream)), "guidance_host")
        return self.seeds_host, self.geo_host
