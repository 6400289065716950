// Internal interface of the geodesic propagation (shared with the fused hot path).
#pragma once
#include "gf_common.cuh"

namespace gf {

size_t geodesic_workspace_bytes(int N, int k, int Q);
int geodesic_run(const float *D, const void *I, int is64, int N, int k, const int *seeds, int Q, float radius,
                 int max_step, float *geo, int64_t *stats_out, void *workspace, size_t workspace_bytes,
                 cudaStream_t st, float *const *peer_rows = nullptr, int n_peers = 0, float *row_max = nullptr);

}  // namespace gf
