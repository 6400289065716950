// Multi-source geodesic propagation over the kNN graph for sm_100a.
//
// Reference: model/geoformer/geodesic_utils.py:91-164 (cal_geodesic_vectorize).  NOT a shortest
// path: it is a level-synchronous, first-visit-wins BFS.  A point receives its distance at the
// first hop level at which it is reached and is never relaxed again; among several parents at
// that level the winner is the first candidate in the reference's candidate order, i.e. the
// parent with the smallest point index, then the smallest neighbour slot (geodesic_utils.py:
// 131-136 keep the first occurrence of each (point, query) pair; the candidate list is built from
// the lexicographically sorted winner list of the previous level, :131-135, :154).
//
// Formulation here ("pull, bit-parallel over seeds"):
//   * the kNN graph (radius-filtered, column 0 dropped) is transposed ONCE per scene into a
//     reverse CSR whose rows are sorted by (parent index, slot) -- the reference's tie order;
//   * visited / frontier sets are bit matrices [point][seed word]: one 32-bit word serves 32
//     seeds, a whole row of 256 seeds is one 32-byte sector;
//   * one persistent cooperative kernel runs all levels (grid barrier between levels, no host
//     synchronisation -- the reference syncs the host >= 3 times per level).  In a level every
//     (target, word) thread walks the target's in-edges in tie order and takes, for each seed bit
//     that is still unvisited, the FIRST parent whose frontier bit is set: no atomics, no
//     sort/unique, deterministic, and exactly the reference's winner.
//   * distances live directly in the (Q,N) output; one fp32 add per reached pair, as the reference.
// Points are optionally renumbered in the cell order of the kNN grid (order/rank) so that the
// frontier rows a warp touches are neighbours in memory; tie-breaking still uses ORIGINAL indices.
#include <cooperative_groups.h>

#include "gf_geodesic.cuh"

namespace cg = cooperative_groups;

namespace gf {

// ---- generic exclusive scan over n ints (n known on the host) ------------------------------------
constexpr int SCAN_BLOCKS = 512;

__global__ void __launch_bounds__(256) scan_a_kernel(const int *__restrict__ in, int n, int *__restrict__ bsum) {
  const int chunk = (n + SCAN_BLOCKS - 1) / SCAN_BLOCKS;
  const int b0 = blockIdx.x * chunk, b1 = min(n, b0 + chunk);
  int acc = 0;
  for (int i = b0 + threadIdx.x; i < b1; i += 256) acc += in[i];
  __shared__ int ws[8];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    bsum[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(SCAN_BLOCKS) scan_b_kernel(int *__restrict__ bsum) {
  __shared__ int s[SCAN_BLOCKS];
  int v = bsum[threadIdx.x];
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_BLOCKS; o <<= 1) {
    int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  bsum[threadIdx.x] = s[threadIdx.x] - v;
}

// out[i] = sum(in[0..i)), out[n] = total
__global__ void __launch_bounds__(256) scan_c_kernel(const int *__restrict__ in, int n, const int *__restrict__ bsum,
                                                     int *__restrict__ out) {
  const int chunk = (n + SCAN_BLOCKS - 1) / SCAN_BLOCKS;
  const int b0 = blockIdx.x * chunk, b1 = min(n, b0 + chunk);
  __shared__ int ws[8];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = bsum[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = b0; t0 < b1; t0 += 1024) {
    int i0 = t0 + threadIdx.x * 4;
    int c[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) c[e] = (i0 + e < b1) ? in[i0 + e] : 0;
    int tsum = c[0] + c[1] + c[2] + c[3];
    int inc = tsum;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += ws[w];
    int excl = carry + woff + inc - tsum;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (i0 + e < b1) out[i0 + e] = excl;
      excl += c[e];
    }
    __syncthreads();
    if (threadIdx.x == 255) carry = excl;
    __syncthreads();
  }
  if (b1 == n && b0 < n && threadIdx.x == 0) out[n] = carry;
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

static int exclusive_scan(const int *in, int n, int *out, int *bsum, cudaStream_t st) {
  scan_a_kernel<<<SCAN_BLOCKS, 256, 0, st>>>(in, n, bsum);
  GF_LAUNCHED();
  scan_b_kernel<<<1, SCAN_BLOCKS, 0, st>>>(bsum);
  GF_LAUNCHED();
  scan_c_kernel<<<SCAN_BLOCKS, 256, 0, st>>>(in, n, bsum, out);
  GF_LAUNCHED();
  return GF_OK;
}

// ---- reverse CSR construction ---------------------------------------------------------------------
__device__ __forceinline__ long long load_idx(const void *idx, int is64, size_t at) {
  return is64 ? ((const long long *)idx)[at] : (long long)((const int *)idx)[at];
}

// one thread per forward edge (p, slot j of the K = k-1 usable columns); valid iff D <= radius and
// 0 <= I < N (geodesic_utils.py:123,151)
__global__ void geo_count_edges_kernel(const float *__restrict__ D, const void *__restrict__ I, int is64, int N, int k,
                                       float radius, const int *__restrict__ rank, int *__restrict__ rev_count) {
  const int K = k - 1;
  const long long total = (long long)N * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int p = (int)(e / K), j = (int)(e - (long long)p * K);
    size_t at = (size_t)p * k + 1 + j;
    long long t = load_idx(I, is64, at);
    if (__ldg(D + at) <= radius && t >= 0 && t < N) atomicAdd(rev_count + (rank ? __ldg(rank + t) : (int)t), 1);
  }
}

__global__ void geo_fill_edges_kernel(const float *__restrict__ D, const void *__restrict__ I, int is64, int N, int k,
                                      float radius, const int *__restrict__ rank, const int *__restrict__ rev_start,
                                      int *__restrict__ cursor, unsigned long long *__restrict__ rev_key) {
  const int K = k - 1;
  const long long total = (long long)N * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int p = (int)(e / K), j = (int)(e - (long long)p * K);
    size_t at = (size_t)p * k + 1 + j;
    long long t = load_idx(I, is64, at);
    if (__ldg(D + at) <= radius && t >= 0 && t < N) {
      int ti = rank ? __ldg(rank + t) : (int)t;
      int s = atomicAdd(cursor + ti, 1);
      rev_key[(size_t)__ldg(rev_start + ti) + s] = ((unsigned long long)(unsigned)p << 8) | (unsigned)j;
    }
  }
}

// one thread per target row: order the in-edges by (original parent index, slot) -- the reference's
// tie order -- then rewrite each 64-bit key in place as {internal parent id, edge length bits}
__global__ void geo_sort_rows_kernel(const float *__restrict__ D, int N, int k, const int *__restrict__ rank,
                                     const int *__restrict__ rev_start, unsigned long long *__restrict__ rev) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
    const int e0 = rev_start[t], e1 = rev_start[t + 1];
    for (int a = e0 + 1; a < e1; ++a) {
      unsigned long long key = rev[a];
      int b = a - 1;
      while (b >= e0 && rev[b] > key) {
        rev[b + 1] = rev[b];
        --b;
      }
      rev[b + 1] = key;
    }
    for (int a = e0; a < e1; ++a) {
      unsigned long long key = rev[a];
      unsigned p = (unsigned)(key >> 8), j = (unsigned)(key & 255u);
      unsigned pi = rank ? (unsigned)__ldg(rank + p) : p;
      unsigned wb = __float_as_uint(__ldg(D + (size_t)p * k + 1 + j));
      rev[a] = ((unsigned long long)wb << 32) | pi;  // as uint2: .x = parent, .y = length bits
    }
  }
}

// ---- state initialisation ---------------------------------------------------------------------------
__global__ void geo_fill_kernel(float *__restrict__ geo, size_t n, float v) {
  // vectorised bulk + scalar head/tail (geo is only guaranteed 4-byte aligned)
  size_t head = ((16 - ((uintptr_t)geo & 15)) & 15) / 4;
  if (head > n) head = n;
  size_t nvec = (n - head) / 4;
  float4 *g4 = (float4 *)(geo + head);
  const float4 vv = make_float4(v, v, v, v);
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < nvec; i += stride) g4[i] = vv;
  if (tid < head) geo[tid] = v;
  size_t tail0 = head + nvec * 4;
  if (tid < n - tail0) geo[tail0 + tid] = v;
}

__global__ void geo_seed_kernel(const int *__restrict__ seeds, int Q, int N, int W, const int *__restrict__ rank,
                                float *__restrict__ geo, uint32_t *__restrict__ vis, uint32_t *__restrict__ fr) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int s = seeds[q];
  if (s < 0 || s >= N) return;  // the reference would raise an index error; the row stays -1
  geo[(size_t)q * N + s] = 0.f;  // geodesic_utils.py:118
  int si = rank ? rank[s] : s;
  atomicOr(vis + (size_t)si * W + (q >> 5), 1u << (q & 31));  // :119
  atomicOr(fr + (size_t)si * W + (q >> 5), 1u << (q & 31));
}

// ---- the level loop -----------------------------------------------------------------------------------
// flags[0..2]: rotating "this level reached something" flags; stats[0] reached pairs, stats[1] levels run
__global__ void __launch_bounds__(256)
    geo_levels_kernel(int N, int W, int Wshift, int Q, int max_step, const int *__restrict__ rev_start,
                      const uint2 *__restrict__ rev, const int *__restrict__ order, uint32_t *vis, uint32_t *fr0,
                      uint32_t *fr1, float *geo, int *flags, unsigned long long *stats) {
  cg::grid_group grid = cg::this_grid();
  const size_t total = (size_t)N << Wshift;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  uint32_t *cur = fr0, *nxt = fr1;
  __shared__ int s_any;
  __shared__ unsigned long long s_cnt;
  for (int level = 1; level <= max_step; ++level) {
    if (threadIdx.x == 0) {
      s_any = 0;
      s_cnt = 0ull;
    }
    __syncthreads();
    unsigned cnt = 0;
    for (size_t gid = tid0; gid < total; gid += stride) {
      const int t = (int)(gid >> Wshift), w = (int)(gid & (W - 1));
      const int qbase = w << 5;
      uint32_t qmask = qbase + 32 <= Q ? 0xffffffffu : (qbase < Q ? ((1u << (Q - qbase)) - 1u) : 0u);
      const uint32_t vold = vis[gid];
      // level 1 ignores the visited set (geodesic_utils.py:123 has no visited filter)
      uint32_t avail = (level == 1 ? 0xffffffffu : ~vold) & qmask;
      uint32_t newb = 0u;
      if (avail) {
        const int e0 = __ldg(rev_start + t), e1 = __ldg(rev_start + t + 1);
        const int to = order ? __ldg(order + t) : t;
        for (int e = e0; e < e1 && avail; ++e) {
          const uint2 en = __ldg(rev + e);
          uint32_t nb = cur[((size_t)en.x << Wshift) + w] & avail;
          if (nb) {
            avail &= ~nb;
            newb |= nb;
            const float wgt = __uint_as_float(en.y);
            const int po = order ? __ldg(order + en.x) : (int)en.x;
            do {
              int b = __ffs(nb) - 1;
              nb &= nb - 1;
              size_t rowq = (size_t)(qbase + b) * N;
              // level 1: the candidate distance is the edge itself (:127); later: edge + parent (:144)
              float d = level == 1 ? wgt : __fadd_rn(wgt, geo[rowq + po]);
              geo[rowq + to] = d;  // :139
            } while (nb);
          }
        }
      }
      nxt[gid] = newb;
      if (newb) {
        vis[gid] = vold | newb;  // :140
        cnt += __popc(newb);
      }
    }
    if (cnt) {
      s_any = 1;
      atomicAdd(&s_cnt, (unsigned long long)cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_any) {
        flags[level % 3] = 1;
        atomicAdd(stats, s_cnt);
      }
      if (blockIdx.x == 0) flags[(level + 1) % 3] = 0;
    }
    grid.sync();
    const int any = *(volatile int *)(flags + level % 3);
    if (!any) break;  // geodesic_utils.py:156-157
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[1] = (unsigned long long)level;
    uint32_t *tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
}

size_t geodesic_workspace_bytes(int N, int k, int Q) {
  int Qc = Q < GEO_MAX_Q_PER_PASS ? Q : GEO_MAX_Q_PER_PASS;
  int W = 1;
  while (W * 32 < Qc) W <<= 1;
  size_t K = k > 1 ? (size_t)(k - 1) : 0;
  size_t b = 0;
  b += align256(sizeof(int) * ((size_t)N + 1)) * 3;               // rev_count, rev_start, cursor
  b += align256(sizeof(unsigned long long) * ((size_t)N * K + 1));  // rev entries
  b += align256(sizeof(uint32_t) * (size_t)N * W) * 3;             // vis, fr0, fr1
  b += align256(sizeof(int) * SCAN_BLOCKS);
  b += align256(64);  // flags
  b += align256(64);  // internal stats
  return b + 1024;
}

int geodesic_run(const float *D, const void *I, int is64, int N, int k, const int *seeds, int Q, float radius,
                 int max_step, float *geo, const int *order, const int *rank, int64_t *stats_out, void *workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  const int K = k - 1;
  int Qc0 = Q < GEO_MAX_Q_PER_PASS ? Q : GEO_MAX_Q_PER_PASS;
  int W = 1, Wshift = 0;
  while (W * 32 < Qc0) W <<= 1, ++Wshift;
  Arena a(workspace, workspace_bytes);
  int *rev_count = a.take<int>((size_t)N + 1);
  int *rev_start = a.take<int>((size_t)N + 1);
  int *cursor = a.take<int>((size_t)N + 1);
  unsigned long long *rev = a.take<unsigned long long>((size_t)N * (K > 0 ? K : 0) + 1);
  uint32_t *vis = a.take<uint32_t>((size_t)N * W);
  uint32_t *fr0 = a.take<uint32_t>((size_t)N * W);
  uint32_t *fr1 = a.take<uint32_t>((size_t)N * W);
  int *bsum = a.take<int>(SCAN_BLOCKS);
  int *flags = a.take<int>(16);
  unsigned long long *stats = a.take<unsigned long long>(8);
  if (!a.ok) {
    set_error("geodesic: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              geodesic_workspace_bytes(N, k, Q));
    return GF_ERR_WORKSPACE;
  }
  const int nb = num_sms() * 8;
  GF_CUDA(cudaMemsetAsync(stats, 0, 64, st));
  // reverse CSR (once per scene)
  GF_CUDA(cudaMemsetAsync(rev_count, 0, sizeof(int) * ((size_t)N + 1), st));
  GF_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)N + 1), st));
  const long long nedge = (long long)N * K;
  if (nedge > 0) {
    int g = (int)((nedge + 255) / 256 < nb ? (nedge + 255) / 256 : nb);
    geo_count_edges_kernel<<<g, 256, 0, st>>>(D, I, is64, N, k, radius, rank, rev_count);
    GF_LAUNCHED();
  }
  int rc = exclusive_scan(rev_count, N, rev_start, bsum, st);
  if (rc) return rc;
  if (nedge > 0) {
    int g = (int)((nedge + 255) / 256 < nb ? (nedge + 255) / 256 : nb);
    geo_fill_edges_kernel<<<g, 256, 0, st>>>(D, I, is64, N, k, radius, rank, rev_start, cursor, rev);
    GF_LAUNCHED();
    geo_sort_rows_kernel<<<(N + 127) / 128 < nb ? (N + 127) / 128 : nb, 128, 0, st>>>(D, N, k, rank, rev_start, rev);
    GF_LAUNCHED();
  }
  // output fill (geodesic_utils.py:113)
  {
    size_t n = (size_t)Q * N;
    geo_fill_kernel<<<num_sms() * 16, 256, 0, st>>>(geo, n, -1.0f);
    GF_LAUNCHED();
  }
  // cooperative launch geometry
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, geo_levels_kernel, 256, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  for (int q0 = 0; q0 < Q; q0 += GEO_MAX_Q_PER_PASS) {
    const int Qc = Q - q0 < GEO_MAX_Q_PER_PASS ? Q - q0 : GEO_MAX_Q_PER_PASS;
    GF_CUDA(cudaMemsetAsync(vis, 0, sizeof(uint32_t) * (size_t)N * W, st));
    GF_CUDA(cudaMemsetAsync(fr0, 0, sizeof(uint32_t) * (size_t)N * W, st));
    GF_CUDA(cudaMemsetAsync(flags, 0, 64, st));
    float *geo_c = geo + (size_t)q0 * N;
    geo_seed_kernel<<<(Qc + 127) / 128, 128, 0, st>>>(seeds + q0, Qc, N, W, rank, geo_c, vis, fr0);
    GF_LAUNCHED();
    if (q0 == 0) stage_mark(ST_GEO_READY, st);
    if (max_step > 0 && K > 0) {
      size_t total = (size_t)N * W;
      int grid = num_sms() * blocks_per_sm;
      size_t want = (total + 255) / 256;
      if ((size_t)grid > want) grid = (int)(want > 0 ? want : 1);
      const uint2 *rev2 = (const uint2 *)rev;
      unsigned long long *st_ptr = stats;
      int Nn = N, Ww = W, Ws = Wshift, Qq = Qc, ms = max_step;
      void *args[] = {&Nn, &Ww, &Ws, &Qq, &ms, &rev_start, &rev2, &order, &vis, &fr0, &fr1, &geo_c, &flags, &st_ptr};
      GF_CUDA(cudaLaunchCooperativeKernel((const void *)geo_levels_kernel, dim3(grid), dim3(256), args, 0, st));
      count_launch();
    }
  }
  stage_mark(ST_GEO_DONE, st);
  if (stats_out) GF_CUDA(cudaMemcpyAsync(stats_out, stats, 16, cudaMemcpyDeviceToDevice, st));
  return GF_OK;
}

}  // namespace gf

using namespace gf;

extern "C" size_t gf_geodesic_workspace_bytes(int N, int k, int Q) {
  if (N <= 0 || Q <= 0) return 0;
  return geodesic_workspace_bytes(N, k, Q);
}

extern "C" int gf_geodesic(const float *knn_dist, const void *knn_idx, int idx_is_i64, int N, int k, const int *seeds,
                           int Q, float radius, int max_step, float *geo, const int *order, const int *rank,
                           int64_t *stats, void *workspace, size_t workspace_bytes, void *stream) {
  GF_CHECK_ARG(N >= 0 && Q >= 0, "geodesic: negative size");
  GF_CHECK_ARG(k >= 1 && k <= 256, "geodesic: k=%d outside [1,256]", k);
  GF_CHECK_ARG((order == nullptr) == (rank == nullptr), "geodesic: order and rank must be given together");
  if (N == 0 || Q == 0) return GF_OK;
  GF_CHECK_ARG(knn_dist && knn_idx && seeds && geo, "geodesic: null pointer");
  return geodesic_run(knn_dist, knn_idx, idx_is_i64, N, k, seeds, Q, radius, max_step, geo, order, rank, stats,
                      workspace, workspace_bytes, (cudaStream_t)stream);
}
