// Internal interface of the geodesic propagation (shared with the fused hot path).
#pragma once
#include "gf_common.cuh"

namespace gf {

size_t geodesic_workspace_bytes(int N, int k, int Q);
int geodesic_run(const float *D, const void *I, int is64, int N, int k, const int *seeds, int Q, float radius,
                 int max_step, float *geo, int64_t *stats_out, void *workspace, size_t workspace_bytes,
                 cudaStream_t st, float *const *peer_rows = nullptr, int n_peers = 0, float *row_max = nullptr,
                 const int *rank = nullptr, const int *order = nullptr);

// ---- batched launch: the (scene, seed) pairs of several scenes in ONE kernel ----------------------------------
struct GeoSceneDesc {
  const int *tgt;    // packed edge table of the scene, in the format geodesic_edge_buffers reported (encoded)
  const float *len;
  const int *rank, *order;  // the numbering the table is written in (gf_knn.cuh: cell order); both null = original
  const int *seeds;  // (Q) device
  float *geo;        // (Q, N) device
  float *row_max;    // optional (Q)
  int64_t *stats;    // optional device [reached pairs, deepest level]; ACCUMULATED with atomics: clear before
  int N, Q;
};
size_t geodesic_batch_scratch_bytes(int maxN, long long items);
int geodesic_batch_launch(const GeoSceneDesc *scenes, int B, int k, int max_step, void *scratch, size_t scratch_bytes,
                          cudaStream_t st);
constexpr int GEO_BATCH_MAX = 32;

// where the packed edge table of a later geodesic_run(D = nullptr, ...) on the same workspace lives
int geodesic_edge_buffers(void *workspace, size_t workspace_bytes, int N, int k, int Q, int **tgt, float **len,
                          int *slot_bits, int *enc);

int geodesic_edge_format(int N);  // 1 = encoded targets (the batched kernel takes the scene), 0 = plain

}  // namespace gf
