// The fused hot path of BASELINE.json's north_star: FPS seeds -> kNN graph -> geodesic maps for one
// scene (what GeoFormerFS.forward does at geoformer_fs.py:630-645 + :497-506 through three separate
// libraries and a Python loop).  FPS (a <= 16-SM cluster kernel, latency bound) runs on a forked
// stream concurrently with the kNN graph construction (which fills the other SMs); the geodesic
// waits for both.
#include <stdlib.h>

#include "gf_geodesic.cuh"
#include "gf_knn.cuh"

namespace gf {

struct ForkJoin {
  cudaStream_t aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

// One auxiliary stream per (device, caller stream) of the calling thread, so that callers that keep several
// scenes in flight on several streams also get their FPS kernels (one 16-SM cluster each, latency bound)
// running side by side instead of queueing on a single hidden stream.  Highest priority: FPS is the longest
// dependency chain of a step, its cluster should be placed as soon as SMs free up.
static int get_fork_join(cudaStream_t user, ForkJoin **out) {
  constexpr int SLOTS = 16;
  struct Entry {
    int dev;
    cudaStream_t user;
    ForkJoin fj;
    bool used;
  };
  static thread_local Entry pool[SLOTS];
  static thread_local int next_victim = 0;
  int dev = 0;
  GF_CUDA(cudaGetDevice(&dev));
  Entry *e = nullptr;
  for (int i = 0; i < SLOTS; ++i)
    if (pool[i].used && pool[i].dev == dev && pool[i].user == user) e = &pool[i];
  if (!e) {
    for (int i = 0; i < SLOTS && !e; ++i)
      if (!pool[i].used) e = &pool[i];
    if (!e) {  // more caller streams than slots: share (correct, only less concurrent)
      e = &pool[next_victim];
      next_victim = (next_victim + 1) % SLOTS;
      if (e->dev != dev) e = nullptr;
    }
    if (e && !e->used) {
      int lo = 0, hi = 0;
      GF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = greatest priority
      GF_CUDA(cudaStreamCreateWithPriority(&e->fj.aux, cudaStreamNonBlocking, hi));
      GF_CUDA(cudaEventCreateWithFlags(&e->fj.fork, cudaEventDisableTiming));
      GF_CUDA(cudaEventCreateWithFlags(&e->fj.join, cudaEventDisableTiming));
      e->used = true;
      e->dev = dev;
    }
    if (e) e->user = user;
  }
  if (!e) {
    set_error("guidance: no auxiliary stream available on device %d", dev);
    return GF_ERR_CUDA;
  }
  *out = &e->fj;
  return GF_OK;
}

struct GuidancePlan {
  size_t fps, knn, geo, dist, idx, total;
};

static GuidancePlan plan_guidance(int N, int Q, int k) {
  GuidancePlan p;
  p.fps = align256(gf_fps_workspace_bytes(1, N, Q));
  p.knn = align256(knn_grid_workspace_bytes(N));
  p.geo = align256(geodesic_workspace_bytes(N, k, Q));
  p.dist = align256(sizeof(float) * (size_t)N * k);
  p.idx = align256(sizeof(int) * (size_t)N * k);
  p.total = p.fps + p.knn + p.geo + p.dist + p.idx + 1024;
  return p;
}

}  // namespace gf

using namespace gf;

extern "C" size_t gf_guidance_workspace_bytes(int N, int Q, int k) {
  if (N <= 0 || Q <= 0 || k <= 0) return 0;
  return plan_guidance(N, Q, k).total;
}

// seeds_given == 0: run FPS (forked stream) and write `seeds`; != 0: `seeds` is an input
static int guidance_impl(const float *xyz, int N, int Q, int k, float radius, int max_step, int *seeds,
                         int seeds_given, float *geo, float *knn_dist, int32_t *knn_idx32, int64_t *stats,
                         void *workspace, size_t workspace_bytes, void *stream, float *const *peer_geo = nullptr,
                         int n_peers = 0, float *row_max = nullptr) {
  GF_CHECK_ARG(N >= 1 && Q >= 1, "guidance: need N >= 1 and Q >= 1 (N=%d Q=%d)", N, Q);
  GF_CHECK_ARG(k >= 1 && k <= KNN_MAX_K, "guidance: k=%d outside [1,%d]", k, KNN_MAX_K);
  GF_CHECK_ARG(xyz && seeds && geo, "guidance: null pointer");
  GuidancePlan p = plan_guidance(N, Q, k);
  if (workspace == nullptr || workspace_bytes < p.total) {
    set_error("guidance: workspace too small (%zu bytes given, %zu needed)", workspace_bytes, p.total);
    return GF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char *w = (char *)workspace;
  void *ws_fps = w;
  w += p.fps;
  void *ws_knn = w;
  w += p.knn;
  void *ws_geo = w;
  w += p.geo;
  float *dist = knn_dist ? knn_dist : (float *)w;
  w += p.dist;
  int *idx = knn_idx32 ? knn_idx32 : (int *)w;

  ForkJoin *fj = nullptr;
  int rc = GF_OK;
  stage_mark(ST_BEGIN, st);
  if (!seeds_given) {
    rc = get_fork_join(st, &fj);
    if (rc) return rc;
    // fork: FPS on the auxiliary stream
    GF_CUDA(cudaEventRecord(fj->fork, st));
    GF_CUDA(cudaStreamWaitEvent(fj->aux, fj->fork, 0));
    rc = gf_furthest_point_sampling(xyz, 1, N, Q, seeds, ws_fps, p.fps, fj->aux);
    if (rc) return rc;
    GF_CUDA(cudaEventRecord(fj->join, fj->aux));
  }
  // main stream: kNN graph
  KnnGridBuffers gb;
  rc = knn_grid_build(xyz, N, k, ws_knn, p.knn, st, &gb);
  if (rc) return rc;
  stage_mark(ST_KNN_BUILT, st);
  rc = knn_grid_query(gb, nullptr, N, k, /*sqrt=*/1, dist, nullptr, idx, st);
  if (rc) return rc;
  // join, then propagate
  if (!seeds_given) GF_CUDA(cudaStreamWaitEvent(st, fj->join, 0));
  stage_mark(ST_KNN_DONE, st);
  return geodesic_run(dist, idx, /*is64=*/0, N, k, seeds, Q, radius, max_step, geo, stats, ws_geo, p.geo, st, peer_geo,
                      n_peers, row_max);
}

extern "C" int gf_guidance(const float *xyz, int N, int Q, int k, float radius, int max_step, int *seeds, float *geo,
                           float *knn_dist, int32_t *knn_idx32, int64_t *stats, float *row_max, void *workspace,
                           size_t workspace_bytes, void *stream) {
  return guidance_impl(xyz, N, Q, k, radius, max_step, seeds, 0, geo, knn_dist, knn_idx32, stats, workspace,
                       workspace_bytes, stream, nullptr, 0, row_max);
}

extern "C" int gf_guidance_seeded(const float *xyz, int N, const int *seeds, int Q, int k, float radius, int max_step,
                                  float *geo, float *knn_dist, int32_t *knn_idx32, int64_t *stats, float *row_max,
                                  void *workspace, size_t workspace_bytes, void *stream) {
  return guidance_impl(xyz, N, Q, k, radius, max_step, const_cast<int *>(seeds), 1, geo, knn_dist, knn_idx32, stats,
                       workspace, workspace_bytes, stream, nullptr, 0, row_max);
}

extern "C" int gf_guidance_seeded_scatter(const float *xyz, int N, const int *seeds, int Q, int k, float radius,
                                          int max_step, float *geo, float *const *peer_geo, int n_peers,
                                          int64_t *stats, void *workspace, size_t workspace_bytes, void *stream) {
  return guidance_impl(xyz, N, Q, k, radius, max_step, const_cast<int *>(seeds), 1, geo, nullptr, nullptr, stats,
                       workspace, workspace_bytes, stream, peer_geo, n_peers);
}

extern "C" size_t gf_guidance_host_workspace_bytes(int N, int Q, int k) {
  if (N <= 0 || Q <= 0 || k <= 0) return 0;
  return plan_guidance(N, Q, k).total + align256(sizeof(float) * (size_t)N * 3) + align256(sizeof(int) * (size_t)Q) +
         align256(sizeof(float) * (size_t)Q * N) + 1024;
}

extern "C" int gf_guidance_host(const float *xyz_host, int N, int Q, int k, float radius, int max_step,
                                int *seeds_host, float *geo_host, void *workspace, size_t workspace_bytes,
                                void *stream) {
  GF_CHECK_ARG(N >= 1 && Q >= 1, "guidance_host: need N >= 1 and Q >= 1");
  GF_CHECK_ARG(xyz_host && seeds_host && geo_host, "guidance_host: null host pointer");
  size_t need = gf_guidance_host_workspace_bytes(N, Q, k);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("guidance_host: workspace too small (%zu bytes given, %zu needed)", workspace_bytes, need);
    return GF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char *w = (char *)workspace;
  float *d_xyz = (float *)w;
  w += align256(sizeof(float) * (size_t)N * 3);
  int *d_seeds = (int *)w;
  w += align256(sizeof(int) * (size_t)Q);
  float *d_geo = (float *)w;
  w += align256(sizeof(float) * (size_t)Q * N);
  size_t inner = plan_guidance(N, Q, k).total;
  GF_CUDA(cudaMemcpyAsync(d_xyz, xyz_host, sizeof(float) * (size_t)N * 3, cudaMemcpyHostToDevice, st));
  int rc = gf_guidance(d_xyz, N, Q, k, radius, max_step, d_seeds, d_geo, nullptr, nullptr, nullptr, nullptr, w, inner, st);
  if (rc) return rc;
  GF_CUDA(cudaMemcpyAsync(seeds_host, d_seeds, sizeof(int) * (size_t)Q, cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaMemcpyAsync(geo_host, d_geo, sizeof(float) * (size_t)Q * N, cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  return GF_OK;
}
