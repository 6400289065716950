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


BATCH_MAX = 32  # scenes per library call (GEO_BATCH_MAX)
BATCH_MAX_POINTS = 860_000  # beyond this a scene's bitmaps do not fit one CTA: per-scene calls


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])


def geodesic_guidance_batch(scenes, n_queries, neighbor, radius, max_step, seeds=None, return_stats=False,
                            row_max=False, ws_tag="guidance_batch"):
    """The hot path for a BATCH of scenes in one library call per <= 32 scenes: FPS and graph construction of the
    scenes side by side, ONE propagation launch over all (scene, seed) pairs (gf_guidance_batch).
    scenes: list of (N_b, 3) f32 CUDA tensors on one device.  seeds: optional list of (Q,) int tensors (then FPS is
    skipped -- the body of cal_geodesic_vectorize).  Returns (seeds list, geo list[, stats (B,2) i64][, row_max list])."""
    C.require(len(scenes) > 0, "empty batch")
    dev = scenes[0].device
    Q, k = int(n_queries), int(neighbor)
    for x in scenes:
        C.check_cuda_f32(x, "xyz")
        C.require(x.dim() == 2 and x.size(1) == 3 and x.size(0) >= 1 and x.device == dev, "scenes must be (N, 3) on one device")
    given = seeds is not None
    if given:
        seeds = [s.to(device=dev, dtype=torch.int32).contiguous() for s in seeds]
        C.require(len(seeds) == len(scenes) and all(s.numel() == Q for s in seeds), "one (Q,) seed tensor per scene")
    else:
        seeds = [torch.empty((Q,), dtype=torch.int32, device=dev) for _ in scenes]
    geo = [torch.empty((Q, x.size(0)), dtype=torch.float32, device=dev) for x in scenes]
    rmax = [torch.empty((Q,), dtype=torch.float32, device=dev) for _ in scenes] if row_max else None
    stats = torch.zeros((len(scenes), 2), dtype=torch.int64, device=dev) if return_stats else None
    L = C.lib()
    with torch.cuda.device(dev):
        st = C.stream_of(dev)
        for b0 in range(0, len(scenes), BATCH_MAX):
            sl = slice(b0, min(len(scenes), b0 + BATCH_MAX))
            B = sl.stop - sl.start
            Ns = (ctypes.c_int * B)(*[x.size(0) for x in scenes[sl]])
            nbytes = L.gf_guidance_batch_workspace_bytes(Ns, B, Q, k)
            C.require(nbytes > 0, "guidance_batch: bad sizes")
            ws = C.workspace.get(dev, ws_tag, nbytes)
            C.check(L.gf_guidance_batch(_ptr_array(scenes[sl]), Ns, B, Q, k, ctypes.c_float(float(radius)),
                                        int(max_step), _ptr_array(seeds[sl]), 1 if given else 0, _ptr_array(geo[sl]),
                                        _ptr_array(rmax[sl]) if rmax else None,
                                        stats[sl].data_ptr() if stats is not None else None, C.ptr(ws), nbytes, st),
                    "guidance_batch")
    out = [seeds, geo]
    if return_stats:
        out.append(stats)
    if row_max:
        out.append(rmax)
    return tuple(out)


class BatchGuidanceRunner:
    """Pre-allocated runner for batches of B equally sized scenes (no allocation per call), optionally replayed as
    ONE CUDA graph: run(list of B (N,3) tensors) -> (seeds (B,Q) i32, geo (B,Q,N) f32) views of its own buffers,
    valid until the next run().  The per-batch propagation launch is what bench.py's roofline line times: with
    stage_events=True two events bracket it in every call / replay (see propagation_ms())."""

    def __init__(self, N, B, n_queries, neighbor, radius, max_step, device="cuda", graph=False, stage_events=False):
        self.N, self.B, self.Q, self.k = int(N), int(B), int(n_queries), int(neighbor)
        C.require(1 <= self.B <= BATCH_MAX, "B must be in [1, %d]" % BATCH_MAX)
        self.radius, self.max_step = float(radius), int(max_step)
        self.device = torch.device(device)
        dev = self.device
        self.xyz = torch.empty((self.B, self.N, 3), dtype=torch.float32, device=dev)  # staging: the graph reads from here
        self.seeds = torch.empty((self.B, self.Q), dtype=torch.int32, device=dev)
        self.geo = torch.empty((self.B, self.Q, self.N), dtype=torch.float32, device=dev)
        self.row_max = torch.empty((self.B, self.Q), dtype=torch.float32, device=dev)
        self.stats = torch.zeros((self.B, 2), dtype=torch.int64, device=dev)
        self._L = C.lib()
        self._Ns = (ctypes.c_int * self.B)(*([self.N] * self.B))
        self._nbytes = self._L.gf_guidance_batch_workspace_bytes(self._Ns, self.B, self.Q, self.k)
        C.require(self._nbytes > 0, "scenes of %d points do not fit the batched call" % self.N)
        self._ws = torch.empty(self._nbytes, dtype=torch.uint8, device=dev)
        self._xyz_p = _ptr_array([self.xyz[b] for b in range(self.B)])
        self._seeds_p = _ptr_array([self.seeds[b] for b in range(self.B)])
        self._geo_p = _ptr_array([self.geo[b] for b in range(self.B)])
        self._rmax_p = _ptr_array([self.row_max[b] for b in range(self.B)])
        self._use_graph = bool(graph)
        self._graph = None
        self._capture_stream = None
        self.launches_per_run = None
        self._ev = None
        if stage_events:
            ev = [None, None, None, self._L.gf_event_create(), self._L.gf_event_create()]
            self._ev = (ctypes.c_void_p * 5)(*ev)

    def load(self, scenes, stream=None):
        """copy the batch's points into the staging buffer (asynchronous on `stream`)"""
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        C.require(len(scenes) == self.B, "need %d scenes" % self.B)
        with torch.cuda.stream(st):
            for b, x in enumerate(scenes):
                C.require(tuple(x.shape) == (self.N, 3), "scenes must be (N, 3) with the runner's N")
                self.xyz[b].copy_(x, non_blocking=True)

    def _launch(self, st):
        if self._ev is not None:
            self._L.gf_set_stage_events(self._ev, 5)
        C.check(self._L.gf_guidance_batch(self._xyz_p, self._Ns, self.B, self.Q, self.k, ctypes.c_float(self.radius),
                                          self.max_step, self._seeds_p, 0, self._geo_p, self._rmax_p,
                                          self.stats.data_ptr(), C.ptr(self._ws), self._nbytes,
                                          ctypes.c_void_p(st.cuda_stream)), "guidance_batch")

    def _capture(self, st):
        self._capture_stream = torch.cuda.Stream(device=self.device)
        self._capture_stream.wait_stream(st)
        for _ in range(2):  # warm-up on the capture stream: creates its auxiliary streams, sets kernel attributes
            self._launch(self._capture_stream)
        self._capture_stream.synchronize()
        g = torch.cuda.CUDAGraph()
        before = C.launch_count()
        with torch.cuda.graph(g, stream=self._capture_stream):
            self._launch(self._capture_stream)
        self.launches_per_run = C.launch_count() - before
        self._graph = g
        st.wait_stream(self._capture_stream)

    def replay(self, stream=None):
        """run the hot path on the staged points (asynchronous on `stream`)"""
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            if not self._use_graph:
                before = C.launch_count()
                self._launch(st)
                self.launches_per_run = C.launch_count() - before
            else:
                if self._graph is None:
                    self._capture(st)
                with torch.cuda.stream(st):
                    self._graph.replay()
        return self.seeds, self.geo

    def run(self, scenes, stream=None):
        self.load(scenes, stream)
        return self.replay(stream)

    def propagation_ms(self):
        """duration of the propagation launch of the LAST finished call / replay (needs stage_events=True)"""
        if self._ev is None:
            return None
        v = self._L.gf_event_elapsed_ms(self._ev[3], self._ev[4])
        return v if v >= 0 else None

    def close(self):
        if self._ev is not None:
            for h in (self._ev[3], self._ev[4]):
                self._L.gf_event_destroy(h)
            self._ev = None


class GuidanceRunner:
    """Pre-allocated device-resident runner (no allocation per call): run(xyz) -> (seeds, geo) views
    of its own buffers, valid until the next run().  Used by bench.py for the kernel-only number.

    graph=True captures the 17 launches of one call (FPS on its forked stream included) into a CUDA graph at
    the first run() and replays it afterwards: one launch per scene on the host side (0.08 -> 0.013 ms of
    enqueue time at c1, where the 13 tiny kernels of the kNN grid build make the step launch bound: +6 %
    throughput).  The graph reads the points from a private staging buffer of the runner: run() always copies its
    argument there on the run stream first, so the caller's tensors are never written and never aliased."""

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
        self._xyz_stage = None
        self._capture_stream = None
        self.launches_per_run = None  # kernels inside one replay (the library's counter only sees the capture)

    def _capture(self, xyz, st):
        self._xyz_stage = torch.empty((self.N, 3), dtype=torch.float32, device=self.device)
        self._capture_stream = torch.cuda.Stream(device=self.device)
        self._capture_stream.wait_stream(st)  # the caller's points may still be in flight on the run stream
        with torch.cuda.stream(self._capture_stream):
            self._xyz_stage.copy_(xyz, non_blocking=True)
        for _ in range(2):  # warm-up on the capture stream: creates its auxiliary stream, sets kernel attributes
            self._launch(self._xyz_stage, self._capture_stream)
        self._capture_stream.synchronize()
        g = torch.cuda.CUDAGraph()
        before = C.launch_count()
        with torch.cuda.graph(g, stream=self._capture_stream):
            self._launch(self._xyz_stage, self._capture_stream)
        self.launches_per_run = C.launch_count() - before
        self._graph = g
        st.wait_stream(self._capture_stream)

    def run(self, xyz, stream=None):
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        if not self._use_graph:
            return self._launch(xyz, st)
        C.check_cuda_f32(xyz, "xyz")
        C.require(tuple(xyz.shape) == (self.N, 3), "xyz must be (N, 3) with the runner's N")
        if self._graph is None:
            self._capture(xyz, st)
        with torch.cuda.stream(st):
            self._xyz_stage.copy_(xyz, non_blocking=True)
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


class HostBatchGuidance:
    """Host-buffer entry point for batches: run(list of B pinned (N,3) CPU tensors) copies the points to the device,
    runs the batched hot path and copies seeds and maps back, all inside gf_guidance_batch_host (blocking; releases
    the GIL).  Two objects driven from two Python threads keep the host link busy: the maps of one batch (4*Q*N bytes
    per scene) travel while the next batch is computed."""

    def __init__(self, N, B, n_queries, neighbor, radius, max_step, device="cuda", pinned=True):
        self.N, self.B, self.Q, self.k = int(N), int(B), int(n_queries), int(neighbor)
        self.radius, self.max_step = float(radius), int(max_step)
        self.device = torch.device(device)
        self._L = C.lib()
        self._Ns = (ctypes.c_int * self.B)(*([self.N] * self.B))
        self._nbytes = self._L.gf_guidance_batch_host_workspace_bytes(self._Ns, self.B, self.Q, self.k)
        C.require(self._nbytes > 0, "scenes of %d points do not fit the batched call" % self.N)
        self._ws = torch.empty(self._nbytes, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.seeds_host = torch.empty((self.B, self.Q), dtype=torch.int32, pin_memory=pinned)
        self.geo_host = torch.empty((self.B, self.Q, self.N), dtype=torch.float32, pin_memory=pinned)
        self._seeds_p = _ptr_array([self.seeds_host[b] for b in range(self.B)])
        self._geo_p = _ptr_array([self.geo_host[b] for b in range(self.B)])
        self.h2d_bytes = self.N * 3 * 4  # per scene
        self.d2h_bytes = self.Q * 4 + self.Q * self.N * 4

    def run(self, scenes_host):
        C.require(len(scenes_host) == self.B, "need %d scenes" % self.B)
        for x in scenes_host:
            C.require(not x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (self.N, 3),
                      "scenes must be contiguous float32 CPU tensors of shape (N, 3)")
        with torch.cuda.device(self.device):
            C.check(self._L.gf_guidance_batch_host(_ptr_array(scenes_host), self._Ns, self.B, self.Q, self.k,
                                                   ctypes.c_float(self.radius), self.max_step, self._seeds_p,
                                                   self._geo_p, C.ptr(self._ws), self._nbytes,
                                                   ctypes.c_void_p(self.stream.cuda_stream)), "guidance_batch_host")
        return self.seeds_host, self.geo_host
