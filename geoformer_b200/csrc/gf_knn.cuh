// Internal interface of the grid kNN (shared with the fused hot path in gf_guidance.cu).
#pragma once
#include "gf_common.cuh"

namespace gf {

constexpr int KNN_MAX_K = 64;  // register-resident top-k; GeoFormer uses neighbor <= 64 (geoformer_fs.py:503)

struct KnnGridBuffers {
  const void *grid;       // device KnnGrid
  const int *cell_start;  // (ncells+1) exclusive prefix of the cell populations
  const float4 *sorted;   // (N) points in cell order: x, y, z, bit-cast original index
  const int *order;       // (N) order[cell-order position] = original index
  const int *rank;        // (N) rank[original index] = cell-order position
};

// optional second output of the self query: the propagation's packed edge table (gf_geodesic.cuh)
struct KnnEdgeOut {
  int *tgt;    // (N + 1) << slot_bits
  float *len;  // (N + 1) << slot_bits
  float radius;
  int slot_bits;
  const int *rank;  // != nullptr: table in CELL ORDER (rows and targets are cell-order positions), targets stored
                    // encoded as (t >> 5) << 7 | (t & 31) (gf_geodesic.cu); nullptr: original indices, plain
};

size_t knn_grid_workspace_bytes(int N);
int knn_grid_build(const float *xyz, int N, int k, void *workspace, size_t workspace_bytes, cudaStream_t st,
                   KnnGridBuffers *out);
int knn_grid_query(const KnnGridBuffers &b, const float *queries, int nq, int k, int do_sqrt, float *dist,
                   long long *idx64, int *idx32, cudaStream_t st, const KnnEdgeOut *edges = nullptr);

}  // namespace gf
