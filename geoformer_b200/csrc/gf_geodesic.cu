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
// Formulation: the Q seeds are independent BFS runs over the same graph, so ONE CTA owns ONE seed
// from its first level to its last (persistent CTAs pull seeds from a counter).  All level
// synchronisation is a __syncthreads() -- no grid barrier, no host round trip (the reference
// synchronises the host >= 3 times per level) -- and the state of a run lives on chip:
//   * visited set: a bitmap in shared memory (N bits);
//   * frontier: a compacted queue of (point, distance) pairs in shared memory (global overflow);
//   * the seed's own output row geo[q][:] doubles as the claim array.  An unvisited entry holds
//     -1.0f = 0xBF800000.  A candidate (parent p, slot j) -> t claims t with
//         atomicMin(bits(geo[q][t]), 0x80000000 | (p << SB | j))
//     which is smaller than "unvisited", larger than any finished distance (a non-negative float,
//     < 0x80000000), and ordered exactly like the reference's tie rule (parent index, then slot).
//     After a CTA barrier the candidate whose key is still in place is the winner: it overwrites
//     the key with the distance D[p][j] + dist(p) (one fp32 add, as the reference), sets the
//     visited bit and appends (t, distance) to the next frontier.
// Work is proportional to the edges actually expanded (R*K), the kNN rows are streamed straight
// from the forward graph (no transpose / sort / unique), and nothing is allocated per level.
#include <stdlib.h>

#include "gf_geodesic.cuh"

namespace gf {

constexpr int GEO_THREADS = 512;
constexpr int GEO_QCAP = 2048;                   // frontier entries kept in shared memory (per buffer)
constexpr int GEO_SCAP = 4096;                   // claimants of one level kept in shared memory
constexpr int GEO_UNROLL = 4;
constexpr uint32_t GEO_UNVISITED = 0xBF800000u;  // bits of -1.0f
constexpr uint32_t GEO_KEYBIT = 0x80000000u;
constexpr uint32_t GEO_KEYMAX = 0x3F800000u;  // keys must stay below "unvisited"

struct GeoArgs {
  const float *D;  // (N,k) sqrt'ed kNN distances
  const void *I;   // (N,k) int32 / int64 neighbour indices
  int N, k, Q, max_step;
  float radius;
  const int *seeds;
  float *geo;                 // (Q,N)
  int2 *overflow;             // per CTA: N+2 frontier entries beyond GEO_QCAP (two stacks, one per end)
  unsigned *seed_counter;     // work distribution
  unsigned long long *stats;  // [0] reached pairs, [1] deepest level (atomicMax)
  int bitmap_words;           // shared-memory visited bitmap size (0 = test the output row instead)
  int slot_bits;              // key layout / candidate indexing (slots padded to 2^slot_bits)
};

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Frontier entry i: the first GEO_QCAP live in shared memory, the rest in the CTA's global overflow
// area.  Consecutive levels hold disjoint point sets (F_L + F_{L+1} <= N + 1), so ONE buffer of N + 2
// entries serves both: odd levels grow up from index 0, even levels grow down from the top.
__device__ __forceinline__ size_t ovf_index(int i, int level_parity, int N) {
  const size_t j = (size_t)(i - GEO_QCAP);
  return level_parity ? j : (size_t)N + 1 - j;
}
__device__ __forceinline__ int2 frontier_get(const int2 *sq, const int2 *ovf, int i, int level_parity, int N) {
  return i < GEO_QCAP ? sq[i] : ovf[ovf_index(i, level_parity, N)];
}

// warp-aggregated append: one shared-memory atomic per warp instead of one per lane
__device__ __forceinline__ int warp_append_pos(bool want, int *counter) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  int base = 0;
  const unsigned lane = threadIdx.x & 31;
  if (m && lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, m ? __ffs(m) - 1 : 0);
  return base + __popc(m & ((1u << lane) - 1u));
}

template <bool IS64>
__global__ void __launch_bounds__(GEO_THREADS) geo_seed_bfs_kernel(const GeoArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int2 *q0 = reinterpret_cast<int2 *>(smem_raw);
  int2 *q1 = q0 + GEO_QCAP;
  int *sv_t = reinterpret_cast<int *>(q1 + GEO_QCAP);  // survivors of pass A: target, key, distance bits
  uint32_t *sv_key = reinterpret_cast<uint32_t *>(sv_t + GEO_SCAP);
  int *sv_d = reinterpret_cast<int *>(sv_key + GEO_SCAP);
  uint32_t *vis = reinterpret_cast<uint32_t *>(sv_d + GEO_SCAP);
  __shared__ int s_next_n, s_seed_q, s_surv_n;
  __shared__ unsigned long long s_reached;

  const int N = a.N, k = a.k, K = a.k - 1;
  const int KP = 1 << a.slot_bits;
  const float radius = a.radius;
  const int tid = threadIdx.x;
  int2 *ovf = a.overflow + (size_t)blockIdx.x * ((size_t)N + 2);
  unsigned long long reached_total = 0;  // thread 0 only
  int deepest = 0;

  for (;;) {
    if (tid == 0) s_seed_q = (int)atomicAdd(a.seed_counter, 1u);
    __syncthreads();
    const int q = s_seed_q;
    if (q >= a.Q) break;
    float *row = a.geo + (size_t)q * N;
    uint32_t *rowu = reinterpret_cast<uint32_t *>(row);
    // ---- init: row = -1 (geodesic_utils.py:113), visited = {} (:114) -------------------------------
    {
      const size_t head = ((16 - ((uintptr_t)row & 15)) & 15) / 4;
      const size_t h = head < (size_t)N ? head : (size_t)N;
      const size_t nvec = ((size_t)N - h) / 4;
      float4 *r4 = reinterpret_cast<float4 *>(row + h);
      const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
      for (size_t i = tid; i < nvec; i += GEO_THREADS) r4[i] = m1;
      if ((size_t)tid < h) row[tid] = -1.f;
      const size_t tail0 = h + nvec * 4;
      if ((size_t)tid < (size_t)N - tail0) row[tail0 + tid] = -1.f;
      for (int i = tid; i < a.bitmap_words; i += GEO_THREADS) vis[i] = 0u;
    }
    const int s = a.seeds[q];
    const bool seed_ok = s >= 0 && s < N;  // the reference would raise an index error; the row stays -1
    if (tid == 0) {
      s_next_n = 0;
      s_surv_n = 0;
      s_reached = 0ull;
      if (seed_ok) q0[0] = make_int2(s, __float_as_int(0.f));  // :118, distance of the seed
    }
    __syncthreads();
    int F = seed_ok ? 1 : 0;
    int2 *fq = q0, *nq = q1;
    // NOTE the seed is NOT marked visited before level 1: the reference's first expansion has no
    // visited filter (:123), so a seed that appears in its own neighbour row is re-won at level 1.
    int level = 0;
    while (F > 0 && level < a.max_step) {
      ++level;
      const int par = (level - 1) & 1;
      const long long ncand = (long long)F << a.slot_bits;
      // ---- pass A: every valid candidate claims its target; claimants are remembered -----------------
      // batches of GEO_UNROLL candidates per thread: all neighbour loads of a batch are in flight
      // together (the propagation is latency bound: short dependent chains, little work per level)
      // (loop bounds are warp-uniform: the appends below use full-warp ballots)
      for (long long wbase = tid & ~31; wbase < ncand; wbase += (long long)GEO_THREADS * GEO_UNROLL) {
        const long long base = wbase + (tid & 31);
        int2 pe[GEO_UNROLL];
        int slot[GEO_UNROLL];
        long long t[GEO_UNROLL];
        float w[GEO_UNROLL];
        bool ok[GEO_UNROLL];
#pragma unroll
        for (int u = 0; u < GEO_UNROLL; ++u) {
          const long long c = base + (long long)u * GEO_THREADS;
          slot[u] = (int)(c & (KP - 1));
          ok[u] = c < ncand && slot[u] < K;
          pe[u] = ok[u] ? frontier_get(fq, ovf, (int)(c >> a.slot_bits), par, N) : make_int2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < GEO_UNROLL; ++u) {
          const size_t at = (size_t)pe[u].x * k + 1 + slot[u];
          t[u] = ok[u] ? (IS64 ? ((const long long *)a.I)[at] : (long long)((const int *)a.I)[at]) : -1;
          w[u] = ok[u] ? __ldg(a.D + at) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < GEO_UNROLL; ++u) {
          bool go = ok[u] && (w[u] <= radius) && t[u] >= 0 && t[u] < N;  // :123 / :151
          if (go) {
            if (a.bitmap_words)
              go = !(vis[t[u] >> 5] & (1u << (t[u] & 31)));
            else
              go = ld_cg_u32(rowu + t[u]) >= GEO_KEYBIT;
          }
          const uint32_t key = GEO_KEYBIT | ((uint32_t)pe[u].x << a.slot_bits) | (uint32_t)slot[u];
          if (go) atomicMin(rowu + t[u], key);
          const int pos = warp_append_pos(go, &s_surv_n);
          if (go && pos < GEO_SCAP) {
            sv_t[pos] = (int)t[u];
            sv_key[pos] = key;
            // level 1: the distance is the edge itself (:127); later: edge + parent's distance (:144)
            sv_d[pos] = __float_as_int(level == 1 ? w[u] : __fadd_rn(w[u], __int_as_float(pe[u].y)));
          }
        }
      }
      __syncthreads();
      const int nsurv = s_surv_n;
      // ---- pass B: the claimant whose key survived is the reference's winner --------------------------
      if (nsurv <= GEO_SCAP) {
        for (int wi0 = tid & ~31; wi0 < nsurv; wi0 += GEO_THREADS * GEO_UNROLL) {
          const int i0 = wi0 + (tid & 31);
          uint32_t cur[GEO_UNROLL];
#pragma unroll
          for (int u = 0; u < GEO_UNROLL; ++u) {
            const int i = i0 + u * GEO_THREADS;
            cur[u] = i < nsurv ? ld_cg_u32(rowu + sv_t[i]) : 0u;
          }
#pragma unroll
          for (int u = 0; u < GEO_UNROLL; ++u) {
            const int i = i0 + u * GEO_THREADS;
            const bool win = i < nsurv && cur[u] == sv_key[i];
            int tt = 0, dd = 0;
            if (win) {
              tt = sv_t[i], dd = sv_d[i];
              rowu[tt] = (uint32_t)dd;  // :139
              if (a.bitmap_words) atomicOr(vis + (tt >> 5), 1u << (tt & 31));  // :140
            }
            const int pos = warp_append_pos(win, &s_next_n);
            if (win) {
              if (pos < GEO_QCAP)
                nq[pos] = make_int2(tt, dd);
              else
                ovf[ovf_index(pos, level & 1, N)] = make_int2(tt, dd);
            }
          }
        }
      } else {
        // more claimants than the on-chip list holds: walk the candidates again (same tests as pass A)
        for (long long c = tid; c < ncand; c += GEO_THREADS) {
          const int node = (int)(c >> a.slot_bits), sl = (int)(c & (KP - 1));
          if (sl >= K) continue;
          const int2 p1 = frontier_get(fq, ovf, node, par, N);
          const size_t at = (size_t)p1.x * k + 1 + sl;
          const long long t1 = IS64 ? ((const long long *)a.I)[at] : (long long)((const int *)a.I)[at];
          const float w1 = __ldg(a.D + at);
          if (!(w1 <= radius) || t1 < 0 || t1 >= N) continue;
          if (a.bitmap_words && (vis[t1 >> 5] & (1u << (t1 & 31)))) continue;
          const uint32_t key = GEO_KEYBIT | ((uint32_t)p1.x << a.slot_bits) | (uint32_t)sl;
          if (ld_cg_u32(rowu + t1) != key) continue;
          const float d = level == 1 ? w1 : __fadd_rn(w1, __int_as_float(p1.y));
          row[t1] = d;
          if (a.bitmap_words) atomicOr(vis + (t1 >> 5), 1u << (t1 & 31));
          const int pos = atomicAdd(&s_next_n, 1);
          const int2 ne = make_int2((int)t1, __float_as_int(d));
          if (pos < GEO_QCAP)
            nq[pos] = ne;
          else
            ovf[ovf_index(pos, level & 1, N)] = ne;
        }
      }
      __syncthreads();
      const int nextF = s_next_n;
      if (level == 1 && tid == 0) {
        // the seed joins the visited set now; if no self edge re-won it, its distance stays 0
        if (ld_cg_u32(rowu + s) == GEO_UNVISITED) row[s] = 0.f;
        if (a.bitmap_words) atomicOr(vis + (s >> 5), 1u << (s & 31));
      }
      if (nextF > 0 && level > deepest) deepest = level;
      __syncthreads();
      if (tid == 0) {
        s_reached += (unsigned long long)nextF;
        s_next_n = 0;
        s_surv_n = 0;
      }
      F = nextF;
      int2 *tq = fq;
      fq = nq;
      nq = tq;
    }
    if (level == 0 && seed_ok && tid == 0) row[s] = 0.f;  // max_step <= 0: only the seed entry (:118)
    __syncthreads();
    if (tid == 0) reached_total += s_reached;
  }
  if (tid == 0) {
    if (reached_total) atomicAdd(a.stats, reached_total);
    atomicMax(a.stats + 1, (unsigned long long)deepest);
  }
}

static int ceil_log2(int v) {
  int b = 0;
  while ((1 << b) < v) ++b;
  return b;
}

struct GeoPlan {
  int grid, bitmap_words;
  size_t smem;
};

// shared-memory plan, identical for sizing and launching: 227 KB usable per CTA and per SM on sm_100
static void geo_smem_plan(int N, int *bitmap_words, size_t *smem, int *ctas_per_sm) {
  const size_t queues = sizeof(int2) * 2 * GEO_QCAP + 12 * (size_t)GEO_SCAP;
  int words = (N + 31) / 32;
  size_t bytes = queues + sizeof(uint32_t) * (size_t)words;
  if (bytes > (size_t)226 * 1024) {  // scene too large for an on-chip bitmap: test the output row instead
    words = 0;
    bytes = queues;
  }
  int per_sm = (int)((size_t)(227 * 1024) / (bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;  // 4 x 512 threads = the SM's 2048
  *bitmap_words = words;
  *smem = bytes;
  *ctas_per_sm = per_sm;
}

static int plan_geo(int N, int Q, GeoPlan *p) {
  int per_sm = 1;
  geo_smem_plan(N, &p->bitmap_words, &p->smem, &per_sm);
  static int bps_cap = -1;
  if (bps_cap < 0) {
    const char *e = getenv("GF_GEO_BPS");  // experiment knob
    bps_cap = e ? atoi(e) : 0;
  }
  if (bps_cap > 0 && per_sm > bps_cap) per_sm = bps_cap;
  int grid = num_sms() * per_sm;
  if (grid > Q) grid = Q;
  p->grid = grid < 1 ? 1 : grid;
  return GF_OK;
}

size_t geodesic_workspace_bytes(int N, int k, int Q) {
  (void)k;
  int words = 0, per_sm = 1;
  size_t smem = 0;
  geo_smem_plan(N, &words, &smem, &per_sm);
  long long grid = (long long)num_sms() * per_sm;
  if (grid > Q) grid = Q;
  if (grid < 1) grid = 1;
  size_t b = 0;
  b += align256(sizeof(int2) * ((size_t)N + 2) * (size_t)grid);  // frontier overflow
  b += align256(64);                                             // seed counter
  b += align256(64);                                             // stats
  return b + 1024;
}

int geodesic_run(const float *D, const void *I, int is64, int N, int k, const int *seeds, int Q, float radius,
                 int max_step, float *geo, int64_t *stats_out, void *workspace, size_t workspace_bytes,
                 cudaStream_t st) {
  const int K = k - 1;
  GeoPlan p;
  int rc = plan_geo(N, Q, &p);
  if (rc) return rc;
  const int slot_bits = ceil_log2(K > 1 ? K : 1);
  if (((unsigned long long)N << slot_bits) >= GEO_KEYMAX) {
    set_error("geodesic: N=%d with k=%d does not fit the 30-bit claim key (N << %d must be < 2^30)", N, k, slot_bits);
    return GF_ERR_INVALID;
  }
  Arena a(workspace, workspace_bytes);
  int2 *overflow = a.take<int2>(((size_t)N + 2) * (size_t)p.grid);
  unsigned *counter = a.take<unsigned>(16);
  unsigned long long *stats = a.take<unsigned long long>(8);
  if (!a.ok) {
    set_error("geodesic: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              geodesic_workspace_bytes(N, k, Q));
    return GF_ERR_WORKSPACE;
  }
  GF_CUDA(cudaMemsetAsync(counter, 0, 64, st));
  GF_CUDA(cudaMemsetAsync(stats, 0, 64, st));
  stage_mark(ST_GEO_READY, st);
  GeoArgs ga;
  ga.D = D, ga.I = I, ga.N = N, ga.k = k, ga.Q = Q, ga.max_step = max_step, ga.radius = radius;
  ga.seeds = seeds, ga.geo = geo, ga.overflow = overflow, ga.seed_counter = counter, ga.stats = stats;
  ga.bitmap_words = p.bitmap_words, ga.slot_bits = slot_bits;
  if (is64) {
    GF_CUDA(cudaFuncSetAttribute(geo_seed_bfs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    geo_seed_bfs_kernel<true><<<p.grid, GEO_THREADS, p.smem, st>>>(ga);
  } else {
    GF_CUDA(cudaFuncSetAttribute(geo_seed_bfs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    geo_seed_bfs_kernel<false><<<p.grid, GEO_THREADS, p.smem, st>>>(ga);
  }
  GF_LAUNCHED();
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
                           int Q, float radius, int max_step, float *geo, int64_t *stats, void *workspace,
                           size_t workspace_bytes, void *stream) {
  GF_CHECK_ARG(N >= 0 && Q >= 0, "geodesic: negative size");
  GF_CHECK_ARG(k >= 1 && k <= 256, "geodesic: k=%d outside [1,256]", k);
  if (N == 0 || Q == 0) return GF_OK;
  GF_CHECK_ARG(knn_dist && knn_idx && seeds && geo, "geodesic: null pointer");
  return geodesic_run(knn_dist, knn_idx, idx_is_i64, N, k, seeds, Q, radius, max_step, geo, stats, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}
