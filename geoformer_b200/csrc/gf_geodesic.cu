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
#ifdef GF_TRACE
  long long *trace;  // development only: per-level timestamps of CTA 0
#endif
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

// One level of one seed.  SPILLED: part of the frontier lives in the CTA's global overflow area.
//   pass A  every valid candidate (parent p, slot j) -> t claims t with atomicMin on the output row;
//           the claimant that finds the row entry still "unvisited" is the first one for t at this
//           level and appends t to the next frontier (exactly once per new point).
//   pass B  (after a CTA barrier) for every new point the key left in its row entry IS the
//           reference's winner (smallest parent index, then slot): decode it, read the edge length
//           and the parent's finished distance, write the point's distance, mark it visited.
template <bool IS64, bool BITMAP, bool SPILLED>
__device__ __forceinline__ void geo_level(const GeoArgs &a, int level, int F, const int2 *fq, int2 *nq, int2 *ovf,
                                          uint32_t *vis, float *row, int *s_next_n) {
  uint32_t *rowu = reinterpret_cast<uint32_t *>(row);
  const int N = a.N;
  const unsigned k = (unsigned)a.k, K = k - 1, sb = (unsigned)a.slot_bits, KP = 1u << sb;
  const float radius = a.radius;
  const unsigned tid = threadIdx.x;
  const int par = (level - 1) & 1;
  // (F << slot_bits) < 2^30 and N * k < 2^31 are guaranteed by the host-side key check: 32-bit math
  const unsigned ncand = (unsigned)F << sb;
  // ---- pass A: batches of GEO_UNROLL candidates per thread, all neighbour loads of a batch in flight
  for (unsigned c0 = tid; c0 < ncand; c0 += GEO_THREADS * GEO_UNROLL) {
    unsigned key[GEO_UNROLL], t[GEO_UNROLL];
    float w[GEO_UNROLL];
#pragma unroll
    for (int u = 0; u < GEO_UNROLL; ++u) {
      const unsigned c = c0 + u * GEO_THREADS;
      const unsigned slot = c & (KP - 1), node = c >> sb;
      t[u] = 0xffffffffu;
      w[u] = 0.f;
      key[u] = 0u;
      if (c < ncand && slot < K) {
        const int p = SPILLED ? frontier_get(fq, ovf, (int)node, par, N).x : fq[node].x;
        const unsigned at = (unsigned)p * k + 1u + slot;
        if (IS64) {
          const long long tl = ((const long long *)a.I)[at];
          t[u] = (tl >= 0 && tl < N) ? (unsigned)tl : 0xffffffffu;
        } else {
          t[u] = (unsigned)((const int *)a.I)[at];
        }
        w[u] = __ldg(a.D + at);
        key[u] = GEO_KEYBIT | ((unsigned)p << sb) | slot;
      }
    }
#pragma unroll
    for (int u = 0; u < GEO_UNROLL; ++u) {
      if (key[u] && (w[u] <= radius) && t[u] < (unsigned)N) {  // :123 / :151
        const bool seen = BITMAP ? (vis[t[u] >> 5] >> (t[u] & 31)) & 1u : ld_cg_u32(rowu + t[u]) < GEO_KEYBIT;
        if (!seen) {
          const uint32_t old = atomicMin(rowu + t[u], key[u]);
          if (old == GEO_UNVISITED) {  // first claimant of t at this level
            const int pos = atomicAdd(s_next_n, 1);
            const int2 ne = make_int2((int)t[u], 0);
            if (pos < GEO_QCAP)
              nq[pos] = ne;
            else
              ovf[ovf_index(pos, level & 1, N)] = ne;
          }
        }
      }
    }
  }
  __syncthreads();
#ifdef GF_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0 && level < 300) {
    long long tnow;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
    a.trace[level * 4 + 1] = tnow;
    a.trace[level * 4 + 3] = ((long long)F << 32) | (unsigned)*s_next_n;
  }
#endif
  // ---- pass B: resolve the winners -----------------------------------------------------------------
  const int nextF = *s_next_n;
  for (int i = tid; i < nextF; i += GEO_THREADS) {
    int2 *slot_ptr = i < GEO_QCAP ? nq + i : ovf + ovf_index(i, level & 1, N);
    const int t = slot_ptr->x;
    const uint32_t key = ld_cg_u32(rowu + t);
    const unsigned p = (key & 0x7fffffffu) >> sb, slot = key & (KP - 1);
    const float w = __ldg(a.D + p * k + 1u + slot);
    // level 1: the distance is the edge itself (:127); later: edge + parent's distance (:144)
    const float d = level == 1 ? w : __fadd_rn(w, __uint_as_float(ld_cg_u32(rowu + p)));
    row[t] = d;                                                       // :139
    if (BITMAP) atomicOr(vis + ((unsigned)t >> 5), 1u << (t & 31));  // :140
    *slot_ptr = make_int2(t, __float_as_int(d));
  }
}

template <bool IS64, bool BITMAP>
__global__ void __launch_bounds__(GEO_THREADS) geo_seed_bfs_kernel(const GeoArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int2 *q0 = reinterpret_cast<int2 *>(smem_raw);
  int2 *q1 = q0 + GEO_QCAP;
  uint32_t *vis = reinterpret_cast<uint32_t *>(q1 + GEO_QCAP);
  __shared__ int s_next_n, s_seed_q;

  const int N = a.N;
  const int tid = threadIdx.x;
  int2 *ovf = a.overflow + (size_t)blockIdx.x * ((size_t)N + 2);
  unsigned long long reached_total = 0;
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
      if (BITMAP)
        for (int i = tid; i < a.bitmap_words; i += GEO_THREADS) vis[i] = 0u;
    }
    const int s = a.seeds[q];
    const bool seed_ok = s >= 0 && s < N;  // the reference would raise an index error; the row stays -1
    if (tid == 0) {
      s_next_n = 0;
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
#ifdef GF_TRACE
      if (blockIdx.x == 0 && tid == 0 && level < 300) {
        long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        a.trace[level * 4 + 0] = tnow;
      }
#endif
      if (F > GEO_QCAP)
        geo_level<IS64, BITMAP, true>(a, level, F, fq, nq, ovf, vis, row, &s_next_n);
      else
        geo_level<IS64, BITMAP, false>(a, level, F, fq, nq, ovf, vis, row, &s_next_n);
      const int nextF = s_next_n;
      __syncthreads();
#ifdef GF_TRACE
      if (blockIdx.x == 0 && tid == 0 && level < 300) {
        long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        a.trace[level * 4 + 2] = tnow;
      }
#endif
      if (tid == 0) {
        s_next_n = 0;
        if (level == 1) {
          // the seed joins the visited set now; if no self edge re-won it, its distance stays 0
          if (ld_cg_u32(rowu + s) == GEO_UNVISITED) row[s] = 0.f;
          if (BITMAP) atomicOr(vis + (s >> 5), 1u << (s & 31));
        }
      }
      if (nextF > 0) deepest = level > deepest ? level : deepest;
      reached_total += (unsigned long long)nextF;
      F = nextF;
      int2 *tq = fq;
      fq = nq;
      nq = tq;
      __syncthreads();
    }
    if (level == 0 && seed_ok && tid == 0) row[s] = 0.f;  // max_step <= 0: only the seed entry (:118)
  }
  if (tid == 0) {
    if (reached_total) atomicAdd(a.stats, reached_total);
    atomicMax(a.stats + 1, (unsigned long long)deepest);
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast variant (scenes whose visited AND claimed bitmaps fit in shared memory, N <~ 800k).
// The propagation is latency bound: a level is a short dependent chain of L2 accesses, so the kernel
// is organised to make that chain as short as possible and to run once per level:
//   * claims are fire-and-forget RED.MIN on the output row (no round trip);
//   * "first claimant of t at this level" is decided by a test-and-set on a CLAIMED bitmap in shared
//     memory (~100 cycles) instead of the atomic's return value (~1.5 us under load);
//   * the distance of a point is not needed to expand it, only to report it: the winner's key stays in
//     the row for one level and is resolved (edge + parent distance) by the slot-0 lane of the point
//     while the other lanes of the same warp already expand it -- one barrier pair per level.
// Frontier queues hold point ids only.
constexpr int GEO_FAST_UNROLL = 4;
constexpr int GEO_FAST_QCAP = 4096;

template <bool IS64>
__global__ void __launch_bounds__(GEO_THREADS, 2) geo_seed_bfs_fast_kernel(const GeoArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int *q0 = reinterpret_cast<int *>(smem_raw);
  int *q1 = q0 + GEO_FAST_QCAP;
  uint32_t *vis = reinterpret_cast<uint32_t *>(q1 + GEO_FAST_QCAP);
  uint32_t *clm = vis + a.bitmap_words;
  __shared__ int s_next_n, s_seed_q;

  const int N = a.N;
  const unsigned k = (unsigned)a.k, K = k - 1, sb = (unsigned)a.slot_bits, KP = 1u << sb;
  const float radius = a.radius;
  const unsigned tid = threadIdx.x;
  const unsigned slot = tid & (KP - 1);  // GEO_THREADS is a multiple of KP: a thread keeps its slot
  int *ovf = reinterpret_cast<int *>(a.overflow) + (size_t)blockIdx.x * 2 * ((size_t)N + 2);
  unsigned long long reached_total = 0;
  int deepest = 0;

  for (;;) {
    if (tid == 0) s_seed_q = (int)atomicAdd(a.seed_counter, 1u);
    __syncthreads();
    const int q = s_seed_q;
    if (q >= a.Q) break;
    float *row = a.geo + (size_t)q * N;
    uint32_t *rowu = reinterpret_cast<uint32_t *>(row);
    {  // init: row = -1 (geodesic_utils.py:113), visited = claimed = {} (:114)
      const size_t head = ((16 - ((uintptr_t)row & 15)) & 15) / 4;
      const size_t h = head < (size_t)N ? head : (size_t)N;
      const size_t nvec = ((size_t)N - h) / 4;
      float4 *r4 = reinterpret_cast<float4 *>(row + h);
      const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
      for (size_t i = tid; i < nvec; i += GEO_THREADS) r4[i] = m1;
      if ((size_t)tid < h) row[tid] = -1.f;
      const size_t tail0 = h + nvec * 4;
      if ((size_t)tid < (size_t)N - tail0) row[tail0 + tid] = -1.f;
      for (int i = tid; i < 2 * a.bitmap_words; i += GEO_THREADS) vis[i] = 0u;  // vis and clm are contiguous
    }
    const int s = a.seeds[q];
    const bool seed_ok = s >= 0 && s < N;
    if (tid == 0) {
      s_next_n = 0;
      if (seed_ok) q0[0] = s;
    }
    __syncthreads();
    int F = seed_ok ? 1 : 0;
    int *fq = q0, *nq = q1;
    int level = 0;
    // frontier(level) = points won at level-1 (their row entry still holds the winning key) ------------
    // A group of KP consecutive lanes expands one frontier point (lane = neighbour slot); a group keeps
    // GEO_FAST_UNROLL points in flight.  The lane whose slot is not a real neighbour column (slot == K
    // when K is not a power of two, else slot 0) resolves the point's own distance meanwhile.
    const unsigned group = tid >> sb, ngroups = GEO_THREADS >> sb;
    const bool cand_lane = slot < K;
    const unsigned rslot = K < KP ? K : 0u;
    while (F > 0 && level < a.max_step) {
      ++level;
      const int par = (level - 1) & 1;
      const bool resolve_lane = slot == rslot && level > 1;
#ifdef GF_TRACE
      if (blockIdx.x == 0 && tid == 0 && level < 300) {
        long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        a.trace[level * 4 + 0] = tnow;
      }
#endif
      const int Fs = F < GEO_FAST_QCAP ? F : GEO_FAST_QCAP;  // part of the frontier held in shared memory
      for (int n0 = (int)group; n0 < F; n0 += (int)ngroups * GEO_FAST_UNROLL) {
        int v[GEO_FAST_UNROLL];
        unsigned t[GEO_FAST_UNROLL];
        float w[GEO_FAST_UNROLL];
        uint32_t rkey[GEO_FAST_UNROLL];
        float rd[GEO_FAST_UNROLL];
#pragma unroll
        for (int u = 0; u < GEO_FAST_UNROLL; ++u) {
          const int node = n0 + u * (int)ngroups;
          v[u] = node < Fs ? fq[node] : (node < F ? ovf[(size_t)par * (N + 2) + (node - GEO_FAST_QCAP)] : -1);
        }
#pragma unroll
        for (int u = 0; u < GEO_FAST_UNROLL; ++u) {
          t[u] = 0xffffffffu;
          w[u] = 0.f;
          rkey[u] = 0u;
          if (v[u] >= 0) {
            if (cand_lane) {
              const unsigned at = (unsigned)v[u] * k + 1u + slot;
              if (IS64) {
                const long long tl = ((const long long *)a.I)[at];
                t[u] = (tl >= 0 && tl < N) ? (unsigned)tl : 0xffffffffu;
              } else {
                t[u] = (unsigned)((const int *)a.I)[at];
              }
              w[u] = __ldg(a.D + at);
            }
            if (resolve_lane) rkey[u] = ld_cg_u32(rowu + v[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < GEO_FAST_UNROLL; ++u) {
          if (rkey[u]) {  // resolve lanes only: edge length and parent distance of the point itself
            const unsigned pu = (rkey[u] & 0x7fffffffu) >> sb, sl = rkey[u] & (KP - 1);
            const float ww = __ldg(a.D + pu * k + 1u + sl);
            // points won at level 1 take the edge itself (:127); later ones edge + parent (:144)
            const float dp = level == 2 ? 0.f : __uint_as_float(ld_cg_u32(rowu + pu));
            rd[u] = level == 2 ? ww : __fadd_rn(ww, dp);
          }
        }
#pragma unroll
        for (int u = 0; u < GEO_FAST_UNROLL; ++u) {
          if (rkey[u]) row[v[u]] = rd[u];                    // :139
          if ((w[u] <= radius) && t[u] < (unsigned)N) {  // :123 / :151 (t = ~0 when inactive)
            const unsigned tw = t[u] >> 5, tb = 1u << (t[u] & 31);
            if (!(vis[tw] & tb)) {
              atomicMin(rowu + t[u], GEO_KEYBIT | ((unsigned)v[u] << sb) | slot);
              if (!(atomicOr(clm + tw, tb) & tb)) {  // first claimant of t at this level
                const int pos = atomicAdd(&s_next_n, 1);
                if (pos < GEO_FAST_QCAP)
                  nq[pos] = (int)t[u];
                else
                  ovf[(size_t)(level & 1) * (N + 2) + (pos - GEO_FAST_QCAP)] = (int)t[u];
              }
            }
          }
        }
      }
      __syncthreads();
#ifdef GF_TRACE
      if (blockIdx.x == 0 && tid == 0 && level < 300) {
        long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        a.trace[level * 4 + 1] = tnow;
        a.trace[level * 4 + 3] = ((long long)F << 32) | (unsigned)s_next_n;
      }
#endif
      // fold this level's claims into the visited set (:140)
      for (int i = tid; i < a.bitmap_words; i += GEO_THREADS) {
        const uint32_t cbits = clm[i];
        if (cbits) {
          vis[i] |= cbits;
          clm[i] = 0u;
        }
      }
      const int nextF = s_next_n;
      __syncthreads();
      if (tid == 0) {
        s_next_n = 0;
        if (level == 1) {
          // the seed joins the visited set now; if no self edge re-won it, its distance stays 0
          if (ld_cg_u32(rowu + s) == GEO_UNVISITED) row[s] = 0.f;
          vis[s >> 5] |= 1u << (s & 31);
        }
      }
      if (nextF > 0) deepest = level > deepest ? level : deepest;
      reached_total += (unsigned long long)nextF;
      F = nextF;
      int *tq = fq;
      fq = nq;
      nq = tq;
      __syncthreads();
#ifdef GF_TRACE
      if (blockIdx.x == 0 && tid == 0 && level < 300) {
        long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        a.trace[level * 4 + 2] = tnow;
      }
#endif
    }
    // the points won at the last executed level still hold their keys: resolve them
    if (level >= 1) {
      const int par = level & 1;
      for (int i = tid; i < F; i += GEO_THREADS) {
        const int vv = i < GEO_FAST_QCAP ? fq[i] : ovf[(size_t)par * (N + 2) + (i - GEO_FAST_QCAP)];
        const uint32_t key = ld_cg_u32(rowu + vv);
        const unsigned pu = (key & 0x7fffffffu) >> sb, sl = key & (KP - 1);
        const float ww = __ldg(a.D + pu * k + 1u + sl);
        row[vv] = level == 1 ? ww : __fadd_rn(ww, __uint_as_float(ld_cg_u32(rowu + pu)));
      }
    }
    if (level == 0 && seed_ok && tid == 0) row[s] = 0.f;  // max_step <= 0: only the seed entry (:118)
    __syncthreads();
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
  int grid, bitmap_words, fast;
  size_t smem;
};

// shared-memory plan, identical for sizing and launching: 227 KB usable per CTA and per SM on sm_100
static void geo_smem_plan(int N, int *bitmap_words, size_t *smem, int *ctas_per_sm, int *fast) {
  int words = (N + 31) / 32;
  const size_t fast_bytes = sizeof(int) * 2 * GEO_FAST_QCAP + sizeof(uint32_t) * 2 * (size_t)words;
  static int no_fast = -1;
  if (no_fast < 0) {
    const char *e = getenv("GF_GEO_NOFAST");  // experiment / test knob: force the general kernel
    no_fast = e ? atoi(e) : 0;
  }
  if (!no_fast && fast_bytes <= (size_t)226 * 1024) {  // visited + claimed bitmaps fit on chip
    int per = (int)((size_t)(227 * 1024) / (fast_bytes + 1024));
    *bitmap_words = words;
    *smem = fast_bytes;
    *ctas_per_sm = per < 1 ? 1 : (per > 2 ? 2 : per);  // __launch_bounds__(512, 2)
    *fast = 1;
    return;
  }
  *fast = 0;
  const size_t queues = sizeof(int2) * 2 * GEO_QCAP;
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
  geo_smem_plan(N, &p->bitmap_words, &p->smem, &per_sm, &p->fast);
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
  int words = 0, per_sm = 1, fast = 0;
  size_t smem = 0;
  geo_smem_plan(N, &words, &smem, &per_sm, &fast);
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
#ifdef GF_TRACE
  static long long *d_trace = nullptr;
  if (!d_trace) cudaMalloc(&d_trace, 8 * 4 * 300);
  cudaMemsetAsync(d_trace, 0, 8 * 4 * 300, st);
  ga.trace = d_trace;
#endif
#define GF_GEO_LAUNCH(I64, BM)                                                                                  \
  do {                                                                                                          \
    GF_CUDA(cudaFuncSetAttribute(geo_seed_bfs_kernel<I64, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                 (int)p.smem));                                                                 \
    geo_seed_bfs_kernel<I64, BM><<<p.grid, GEO_THREADS, p.smem, st>>>(ga);                                      \
  } while (0)
  if (p.fast) {
    if (is64) {
      GF_CUDA(cudaFuncSetAttribute(geo_seed_bfs_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)p.smem));
      geo_seed_bfs_fast_kernel<true><<<p.grid, GEO_THREADS, p.smem, st>>>(ga);
    } else {
      GF_CUDA(cudaFuncSetAttribute(geo_seed_bfs_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)p.smem));
      geo_seed_bfs_fast_kernel<false><<<p.grid, GEO_THREADS, p.smem, st>>>(ga);
    }
  } else if (is64 && p.bitmap_words)
    GF_GEO_LAUNCH(true, true);
  else if (is64)
    GF_GEO_LAUNCH(true, false);
  else if (p.bitmap_words)
    GF_GEO_LAUNCH(false, true);
  else
    GF_GEO_LAUNCH(false, false);
#undef GF_GEO_LAUNCH
  GF_LAUNCHED();
#ifdef GF_TRACE
  {
    static int calls = 0;
    if (++calls == 3) {
      long long h[4 * 300];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
      for (int l = 1; l < 300 && h[l * 4]; ++l)
        fprintf(stderr, "TRACE level %3d F=%6lld next=%6lld passA=%7.2fus passB=%7.2fus\n", l, h[l * 4 + 3] >> 32,
                h[l * 4 + 3] & 0xffffffffll, (h[l * 4 + 1] - h[l * 4]) * 1e-3, (h[l * 4 + 2] - h[l * 4 + 1]) * 1e-3);
    }
  }
#endif
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
