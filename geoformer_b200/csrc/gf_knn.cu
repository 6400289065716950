// Exact L2 k-nearest-neighbour graph for 3-D points on sm_100a.
//
// Replaces faiss.GpuIndexFlatL2.search as used by model/geoformer/geodesic_utils.py:11-24
// (call sites geoformer.py:172-177, geoformer_fs.py:170-175).  Result contract:
// neighbours ordered by (d2, index), d2 = fmaf(dz,dz, fmaf(dx,dx, dy*dy)) in fp32, k-th missing
// neighbour = (-1, +inf).
//
// Two algorithms, identical results:
//   algo 0  uniform-grid search.  Points are counting-sorted into cells of edge h (chosen on the
//           device from the measured cell occupancy, no host round trip); each query walks the
//           cell shells around its own cell and stops as soon as the k-th best distance is
//           provably smaller than the distance to every unvisited cell.  Work per query is
//           O(k) instead of O(N); threads are laid out in cell order so a warp shares its cells.
//   algo 1  brute-force tiled scan (shared-memory tiles, register-resident sorted top-k).  This
//           is the formulation named in BASELINE.json; kept as an independent cross-check.
// 3-D coordinates make this a compare/select problem, not a contraction: no tensor cores.
#include <stdlib.h>

#include "gf_knn.cuh"

namespace gf {

constexpr int KNN_MAX_CELLS = 1 << 22;  // dense cell table (2 x 16 MB of int32)
constexpr int KNN_PASS0_CELLS = 1 << 18;
constexpr int KNN_SCAN_BLOCKS = 512;
constexpr int KNN_MAX_DIM = 1024;

// ---- device-side grid description --------------------------------------------------------------
// The control block (KnnCtl: KnnGrid + the scan states) is cleared by ONE cudaMemsetAsync per build; every field is
// laid out so that zero is its initial value.
struct KnnGrid {
  float ox, oy, oz;  // origin = bbox min
  float h, inv_h;
  float slack;  // safety margin of the termination test (absolute length)
  int dx, dy, dz, ncells;
  uint32_t mm[6];             // max over the points of ~ord(v) for x,y,z (= inverted minimum), then of ord(v)
  unsigned long long sum_sq;  // sum over cells of count^2 (occupancy estimate of the coarse pass)
  unsigned done[2];           // last-block-done counters of the two kernels that end with a planning tail
  unsigned pad_[2];
};
struct KnnCtl {
  KnnGrid g;
  unsigned long long scan_state[KNN_SCAN_BLOCKS];  // bit 63 = published, low bits = the block's cell-count sum
};

// pass 0: coarse guess from the bounding-box volume; pass 1: rescale h so that the occupancy seen
// by a point (sum c^2 / N, measured with the pass-0 grid) becomes the target tau.
// Runs in ONE thread, as the tail of the last block of the kernel that produced its inputs.
__device__ void knn_plan(KnnGrid *g, int N, int pass, float tau, int cap) {
  volatile KnnGrid *vg = g;  // the inputs were written by other blocks of the same launch
  float lo[3], ext[3], maxext = 0.f, maxabs = 0.f;
  for (int a = 0; a < 3; ++a) {
    const uint32_t mn = ~vg->mm[a], mx = vg->mm[3 + a];
    bool empty = mn == 0xffffffffu && mx == 0u;
    float l = empty ? 0.f : ord2f(mn), hgh = empty ? 0.f : ord2f(mx);
    lo[a] = l;
    ext[a] = hgh - l;
    maxext = fmaxf(maxext, ext[a]);
    maxabs = fmaxf(maxabs, fmaxf(fabsf(l), fabsf(hgh)));
  }
  if (!(maxext > 0.f)) maxext = 1.f;
  float h;
  if (pass == 0) {
    float vol = 1.f;
    for (int a = 0; a < 3; ++a) vol *= fmaxf(ext[a], 1e-3f * maxext);
    h = cbrtf(vol * tau / (float)(N > 0 ? N : 1));
  } else {
    float occ = (float)((double)vg->sum_sq / (double)(N > 0 ? N : 1));
    float f = sqrtf(tau / fmaxf(occ, 1e-3f));  // surface-like scaling (occupancy ~ h^2)
    f = fminf(fmaxf(f, 0.125f), 4.f);
    h = vg->h * f;
  }
  h = fmaxf(h, maxext * 1e-6f);
  h = fmaxf(h, maxext / (float)KNN_MAX_DIM);
  int d[3];
  for (int it = 0; it < 64; ++it) {
    long long cells = 1;
    for (int a = 0; a < 3; ++a) {
      d[a] = (int)fminf(floorf(ext[a] / h) + 1.f, (float)KNN_MAX_DIM);
      if (d[a] < 1) d[a] = 1;
      cells *= d[a];
    }
    if (cells <= cap) break;
    h *= 1.2599210f;
  }
  g->ox = lo[0], g->oy = lo[1], g->oz = lo[2];
  g->h = h;
  g->inv_h = 1.f / h;
  g->dx = d[0], g->dy = d[1], g->dz = d[2];
  g->ncells = d[0] * d[1] * d[2];
  // cell membership is decided in fp32: a point may sit one rounding error on the wrong side of
  // a face.  The slack below is orders of magnitude above that error and costs < 1 % extra rings.
  g->slack = 1e-3f * h + 4e-6f * maxabs;
}

// true in every thread of the block that finished last (its view of the other blocks' results is ordered
// by the fence / counter pair, as in the CUDA threadFenceReduction sample)
__device__ __forceinline__ bool last_block_done(unsigned *counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = t == gridDim.x - 1;
    __threadfence();
  }
  __syncthreads();
  return s_last;
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int dim) {
  float f = floorf((v - o) * inv_h);
  int c = (f >= 0.f) ? (f < (float)dim ? (int)f : dim - 1) : 0;  // NaN -> 0
  return c;
}
__device__ __forceinline__ int cell_of(const KnnGrid &g, float x, float y, float z) {
  int cx = cell_coord(x, g.ox, g.inv_h, g.dx), cy = cell_coord(y, g.oy, g.inv_h, g.dy),
      cz = cell_coord(z, g.oz, g.inv_h, g.dz);
  return (cz * g.dy + cy) * g.dx + cx;
}

// build kernel 1 of 5: bounding box (+ tail: coarse grid plan) and the zero fill of both cell tables
__global__ void __launch_bounds__(256) knn_bbox_kernel(const float *__restrict__ xyz, int N, KnnGrid *g, float tau,
                                                       int cap_coarse, int4 *__restrict__ zero_a, int n4_a,
                                                       int4 *__restrict__ zero_b, int n4_b) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  const int4 z4 = make_int4(0, 0, 0, 0);
  for (int i = gtid; i < n4_a; i += gsz) zero_a[i] = z4;
  for (int i = gtid; i < n4_b; i += gsz) zero_b[i] = z4;
  uint32_t lo[3] = {0u, 0u, 0u}, hi[3] = {0u, 0u, 0u};  // lo holds the maximum of ~ord = the inverted minimum
  for (int i = gtid; i < N; i += gsz) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = __ldg(xyz + (size_t)i * 3 + a);
      if (v == v && fabsf(v) <= 3.0e38f) {  // ignore NaN / inf for the box
        uint32_t o = f2ord(v);
        lo[a] = max(lo[a], ~o);
        hi[a] = max(hi[a], o);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    uint32_t l = __reduce_max_sync(0xffffffffu, lo[a]);
    uint32_t h = __reduce_max_sync(0xffffffffu, hi[a]);
    if ((threadIdx.x & 31) == 0) {
      if (l) atomicMax(&g->mm[a], l);
      if (h) atomicMax(&g->mm[3 + a], h);
    }
  }
  if (last_block_done(&g->done[0]) && threadIdx.x == 0) knn_plan(g, N, 0, tau, cap_coarse);
}

// build kernels 2 and 3: population of every cell.  The coarse pass (cell_id == nullptr) also accumulates
// sum c^2 = sum over the points of (2 * slot + 1) from the slots its atomics return, and ends with the plan of
// the final grid; the final pass records every point's cell and slot for the scatter.
__global__ void __launch_bounds__(256) knn_count_kernel(const float *__restrict__ xyz, int N, KnnGrid *gp,
                                                        int *__restrict__ cell_count, int *__restrict__ cell_id,
                                                        int *__restrict__ slot_in_cell, float tau, int cap_fine) {
  const KnnGrid g = *gp;
  unsigned long long acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    float x = __ldg(xyz + (size_t)i * 3), y = __ldg(xyz + (size_t)i * 3 + 1), z = __ldg(xyz + (size_t)i * 3 + 2);
    int c = cell_of(g, x, y, z);
    int s = atomicAdd(cell_count + c, 1);
    if (cell_id) {
      cell_id[i] = c;
      slot_in_cell[i] = s;
    } else {
      acc += 2ull * (unsigned)s + 1ull;
    }
  }
  if (cell_id) return;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&gp->sum_sq, acc);
  if (last_block_done(&gp->done[1]) && threadIdx.x == 0) knn_plan(gp, N, 1, tau, cap_fine);
}

// build kernel 4: exclusive scan of the cell populations in one launch.  Block b sums its chunk and publishes
// the sum; its carry-in is the sum of the published sums of the blocks before it (they were scheduled earlier,
// so waiting for them cannot deadlock); then it rescans its chunk and writes the cell starts.
__global__ void __launch_bounds__(256) knn_scan_kernel(KnnCtl *ctl, const int *__restrict__ cnt,
                                                       int *__restrict__ start, int N) {
  const int n = ctl->g.ncells;
  const int chunk = (((n + KNN_SCAN_BLOCKS - 1) / KNN_SCAN_BLOCKS) + 3) & ~3;
  const int b0 = min(n, (int)blockIdx.x * chunk), b1 = min(n, b0 + chunk);
  __shared__ int ws[8];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int acc = 0;
  for (int i = b0 + threadIdx.x; i < b1; i += 256) acc += cnt[i];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) ws[warp] = acc;
  __syncthreads();
  volatile unsigned long long *state = ctl->scan_state;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    state[blockIdx.x] = (1ull << 63) | (unsigned long long)(unsigned)t;
  }
  int pre = 0;
  for (int j = threadIdx.x; j < (int)blockIdx.x; j += 256) {
    unsigned long long v;
    while (!((v = state[j]) >> 63)) __nanosleep(20);
    pre += (int)(unsigned)(v & 0xffffffffull);
  }
  for (int o = 16; o; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
  __syncthreads();  // ws is reused
  if (lane == 0) ws[warp] = pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    carry = t;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) start[n] = N;
  __syncthreads();
  for (int t0 = b0; t0 < b1; t0 += 1024) {
    int i0 = t0 + threadIdx.x * 4;
    int c[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) c[e] = (i0 + e < b1) ? cnt[i0 + e] : 0;
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
      if (i0 + e < b1) start[i0 + e] = excl;
      excl += c[e];
    }
    __syncthreads();
    if (threadIdx.x == 255) carry = excl;
    __syncthreads();
  }
}

// build kernel 5: the points in cell order, each with its original index
__global__ void knn_scatter_kernel(const float *__restrict__ xyz, int N, const int *__restrict__ cell_id,
                                   const int *__restrict__ slot_in_cell, const int *__restrict__ start,
                                   float4 *__restrict__ sorted, int *__restrict__ order, int *__restrict__ rank) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int pos = start[cell_id[i]] + slot_in_cell[i];
    float x = __ldg(xyz + (size_t)i * 3), y = __ldg(xyz + (size_t)i * 3 + 1), z = __ldg(xyz + (size_t)i * 3 + 2);
    sorted[pos] = make_float4(x, y, z, __int_as_float(i));
    order[pos] = i;  // order[cell-order position] = original index
    rank[i] = pos;   // rank[original index] = cell-order position
  }
}

// ---- register-resident sorted top-k -------------------------------------------------------------
// 64-bit keys (d2 bits << 32 | index): lexicographic (d2, index) order, all keys distinct.
// The list is kept ascending; slots [0, KT-k) are pinned to 0 so that the live part is always the
// last k entries and list[KT-1] is the current k-th best.
template <int KT>
struct TopK {
  unsigned long long key[KT];
  __device__ __forceinline__ void init(int k) {
#pragma unroll
    for (int i = 0; i < KT; ++i) key[i] = (i < KT - k) ? 0ull : ~0ull;
  }
  __device__ __forceinline__ void offer(float d2, int index) {
    unsigned long long kk = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)index;
    if (kk < key[KT - 1]) {
#ifdef KNN_SERIAL_INSERT
      key[KT - 1] = kk;
#pragma unroll
      for (int i = KT - 1; i > 0; --i) {
        unsigned long long a = key[i - 1], b = key[i];
        bool sw = b < a;
        key[i - 1] = sw ? b : a;
        key[i] = sw ? a : b;
      }
#else
      // every slot decides for itself (no compare-exchange chain through the moving key): slot i takes its lower
      // neighbour if the new key sorts below that one, the new key if it sorts below the slot's own key, else it
      // keeps what it has; top down, so a slot's lower neighbour is still the old one when it is read
      bool below = true;  // kk < key[KT - 1]
#pragma unroll
      for (int i = KT - 1; i > 0; --i) {
        const bool below_prev = kk < key[i - 1];
        key[i] = below_prev ? key[i - 1] : (below ? kk : key[i]);
        below = below_prev;
      }
      if (below) key[0] = kk;
#endif
    }
  }
  __device__ __forceinline__ void offer_key(unsigned long long kk) {
    if (kk < key[KT - 1]) {
      bool below = true;
#pragma unroll
      for (int i = KT - 1; i > 0; --i) {
        const bool below_prev = kk < key[i - 1];
        key[i] = below_prev ? key[i - 1] : (below ? kk : key[i]);
        below = below_prev;
      }
      if (below) key[0] = kk;
    }
  }
  __device__ __forceinline__ float kth_d2() const { return __uint_as_float((unsigned)(key[KT - 1] >> 32)); }
  // The row of the propagation's edge table (gf_geodesic.cu: geo_pack_edges_kernel, same rule): neighbour
  // 1 + e of the result as edge e if it may ever be used (sqrt(d2) <= radius, geodesic_utils.py:123,151),
  // else an edge to the sentinel point N; padded with such edges to KP = 1 << slot_bits entries.
  // rank != nullptr: the table lives in CELL ORDER (the caller passes the row of the query's cell-order position and
  // targets are translated through rank[]) with encoded targets; else original indices, plain.
  __device__ __forceinline__ void store_edges(int k, float radius, int N, int slot_bits, const int *__restrict__ rank,
                                              int *__restrict__ tgt, float *__restrict__ len) const {
    const int KP = 1 << slot_bits;
    auto code = [&](int t) {
      if (!rank) return t;
      const unsigned r = (unsigned)__ldg(rank + t);
      return (int)(((r >> 5) << 7) | (r & 31u));
    };
    const int none = rank ? (int)((((unsigned)N >> 5) << 7) | ((unsigned)N & 31u)) : N;
    if (k == KT && KP == KT) {  // the usual case (k a power of two): whole 16-byte stores
#pragma unroll
      for (int e0 = 0; e0 < KT; e0 += 4) {
        int t[4];
        float w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u;
          t[u] = none, w[u] = 0.f;
          if (e + 1 < KT) {
            const unsigned long long kk = key[e + 1];
            const float d = sqrtf(__uint_as_float((unsigned)(kk >> 32)));
            if (kk != ~0ull && d <= radius) t[u] = code((int)(unsigned)(kk & 0xffffffffu)), w[u] = d;
          }
        }
        reinterpret_cast<int4 *>(tgt)[e0 >> 2] = make_int4(t[0], t[1], t[2], t[3]);
        reinterpret_cast<float4 *>(len)[e0 >> 2] = make_float4(w[0], w[1], w[2], w[3]);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int e = i - (KT - k) - 1;
      if (e >= 0) {
        const unsigned long long kk = key[i];
        const float d = sqrtf(__uint_as_float((unsigned)(kk >> 32)));
        const bool ok = kk != ~0ull && d <= radius;
        tgt[e] = ok ? code((int)(unsigned)(kk & 0xffffffffu)) : none;
        len[e] = ok ? d : 0.f;
      }
    }
    for (int e = k - 1 > 0 ? k - 1 : 0; e < KP; ++e) tgt[e] = none, len[e] = 0.f;
  }
  __device__ __forceinline__ void store(int k, bool do_sqrt, float *__restrict__ dist, long long *__restrict__ i64,
                                        int *__restrict__ i32) const {
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      int j = i - (KT - k);
      if (j >= 0) {
        unsigned long long kk = key[i];
        bool missing = kk == ~0ull;
        float d2 = missing ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(kk >> 32));
        int id = missing ? -1 : (int)(unsigned)(kk & 0xffffffffu);
        if (dist) dist[j] = do_sqrt ? sqrtf(d2) : d2;
        if (i64) i64[j] = (long long)id;
        if (i32) i32[j] = id;
      }
    }
  }
};

// ---- grid query ---------------------------------------------------------------------------------
#ifndef KNN_MIN_BLOCKS
#define KNN_MIN_BLOCKS 7  // 72 registers for k <= 16 (a few spilled words): measured 1 % ahead of 80 registers / 6 blocks per SM
#endif
template <int KT>
__global__ void __launch_bounds__(128, KT <= 16 ? KNN_MIN_BLOCKS : 1)
    knn_grid_query_kernel(const KnnGrid *__restrict__ gp, const float4 *__restrict__ sorted,
                          const int *__restrict__ start, const float *__restrict__ queries, int nq, int k,
                          int do_sqrt, float *__restrict__ dist, long long *__restrict__ idx64,
                          int *__restrict__ idx32, const KnnEdgeOut eo) {
  const KnnGrid g = *gp;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nq) return;
  if (eo.tgt && s == 0) {  // the sentinel row N: only edges to N
    const int none = eo.rank ? (int)((((unsigned)nq >> 5) << 7) | ((unsigned)nq & 31u)) : nq;
    for (int e = 0; e < (1 << eo.slot_bits); ++e)
      eo.tgt[((size_t)nq << eo.slot_bits) + e] = none, eo.len[((size_t)nq << eo.slot_bits) + e] = 0.f;
  }
  float qx, qy, qz;
  int row;
  if (queries == nullptr) {
    float4 me = sorted[s];
    qx = me.x, qy = me.y, qz = me.z;
    row = __float_as_int(me.w);
  } else {
    qx = __ldg(queries + (size_t)s * 3), qy = __ldg(queries + (size_t)s * 3 + 1), qz = __ldg(queries + (size_t)s * 3 + 2);
    row = s;
  }
  const int cx = cell_coord(qx, g.ox, g.inv_h, g.dx), cy = cell_coord(qy, g.oy, g.inv_h, g.dy),
            cz = cell_coord(qz, g.oz, g.inv_h, g.dz);
  TopK<KT> top;
  top.init(k);
  const int rmax = max(g.dx, max(g.dy, g.dz));
#ifndef KNN_PLAIN_ROW_ORDER
  // distances from the query to the faces of its own cell: lower bounds for everything beyond them
  const float gzl = qz - (g.oz + (float)cz * g.h), gzh = (g.oz + (float)(cz + 1) * g.h) - qz;
  const float gyl = qy - (g.oy + (float)cy * g.h), gyh = (g.oy + (float)(cy + 1) * g.h) - qy;
#endif
  for (int R = 0; R <= rmax; ++R) {
    const int xl = cx - R, xr = cx + R;
    const int side = 2 * R + 1, nrows = side * side;
    for (int ri = 0; ri < nrows; ++ri) {
      // rows (dz, dy) of the shell.  R = 1 (where nearly all the work is): nearest rows first -- the row of the
      // query's own cells, the four rows sharing a face with it, the four diagonal ones -- so that the k-th key is
      // tight before the far cells are scanned, and a row whose slab lies beyond the current k-th distance is skipped.
      int dz, dy;
#ifndef KNN_PLAIN_ROW_ORDER
      if (R == 1) {
        // ri: 0 -> (0,0); 1..4 -> (0,-1) (0,1) (-1,0) (1,0); 5..8 -> (-1,-1) (-1,1) (1,-1) (1,1)
        dz = (int)((0x28215u >> (2 * ri)) & 3u) - 1;  // 2-bit codes of dz + 1, ri = 0 in the low bits
        dy = (int)((0x22161u >> (2 * ri)) & 3u) - 1;
      } else
#endif
      {
        dz = ri / side - R, dy = ri % side - R;
      }
      const int z = cz + dz, y = cy + dy;
      if (z < 0 || z >= g.dz || y < 0 || y >= g.dy) continue;
#ifndef KNN_PLAIN_ROW_ORDER
      if (R == 1) {  // every point of the row is at least (gap_z, gap_y) away from the query
        const float gz = dz < 0 ? gzl : (dz > 0 ? gzh : 0.f), gy = dy < 0 ? gyl : (dy > 0 ? gyh : 0.f);
        const float bz = fmaxf(gz - g.slack, 0.f), by = fmaxf(gy - g.slack, 0.f);
        if (bz * bz + by * by > top.kth_d2()) continue;  // (trimming the row's end cells as well: no further gain)
      }
#endif
      const bool face = dz == -R || dz == R || dy == -R || dy == R;
      const int rowbase = (z * g.dy + y) * g.dx;
      // on a face of the shell the whole x-run belongs to it; otherwise only its two end cells
      const int nseg = (face || R == 0) ? 1 : 2;
      for (int sgi = 0; sgi < nseg; ++sgi) {
        int xa, xb;
        if (nseg == 1) {
          xa = max(xl, 0), xb = min(xr, g.dx - 1);
        } else {
          xa = xb = sgi == 0 ? xl : xr;
          if (xa < 0 || xa >= g.dx) continue;
        }
        if (xa > xb) continue;
        const int p0 = __ldg(start + rowbase + xa), p1 = __ldg(start + rowbase + xb + 1);
        int p = p0;
        for (; p + 4 <= p1; p += 4) {  // four candidates in flight: the loads and distances overlap the offers
          const float4 c0 = __ldg(sorted + p), c1 = __ldg(sorted + p + 1), c2 = __ldg(sorted + p + 2),
                       c3 = __ldg(sorted + p + 3);
          const float e0 = sq3(c0.x - qx, c0.y - qy, c0.z - qz), e1 = sq3(c1.x - qx, c1.y - qy, c1.z - qz),
                      e2 = sq3(c2.x - qx, c2.y - qy, c2.z - qz), e3 = sq3(c3.x - qx, c3.y - qy, c3.z - qz);
          top.offer(e0, __float_as_int(c0.w));
          top.offer(e1, __float_as_int(c1.w));
          top.offer(e2, __float_as_int(c2.w));
          top.offer(e3, __float_as_int(c3.w));
        }
        for (; p < p1; ++p) {
          float4 c = __ldg(sorted + p);
          float d2 = sq3(c.x - qx, c.y - qy, c.z - qz);
          top.offer(d2, __float_as_int(c.w));
        }
      }
    }
    // distance from the query to the nearest face of the visited block that still has cells
    // behind it; every unvisited point is at least that far away (minus the fp32 slack)
    float mfd = __int_as_float(0x7f800000);
    if (cx - R > 0) mfd = fminf(mfd, qx - (g.ox + (float)(cx - R) * g.h));
    if (cx + R + 1 < g.dx) mfd = fminf(mfd, (g.ox + (float)(cx + R + 1) * g.h) - qx);
    if (cy - R > 0) mfd = fminf(mfd, qy - (g.oy + (float)(cy - R) * g.h));
    if (cy + R + 1 < g.dy) mfd = fminf(mfd, (g.oy + (float)(cy + R + 1) * g.h) - qy);
    if (cz - R > 0) mfd = fminf(mfd, qz - (g.oz + (float)(cz - R) * g.h));
    if (cz + R + 1 < g.dz) mfd = fminf(mfd, (g.oz + (float)(cz + R + 1) * g.h) - qz);
    if (mfd == __int_as_float(0x7f800000)) break;  // the whole grid has been visited
    float ms = mfd - g.slack;
    if (ms > 0.f && top.kth_d2() < ms * ms) break;
  }
  if (dist || idx64 || idx32)
    top.store(k, do_sqrt != 0, dist ? dist + (size_t)row * k : nullptr, idx64 ? idx64 + (size_t)row * k : nullptr,
              idx32 ? idx32 + (size_t)row * k : nullptr);
  if (eo.tgt) {  // self query only (nq == N): the propagation's edge rows, written straight from the registers
    const size_t er = (size_t)(eo.rank ? s : row) << eo.slot_bits;  // cell-order row (= this thread) or original row
    top.store_edges(k, eo.radius, nq, eo.slot_bits, eo.rank, eo.tgt + er, eo.len + er);
  }
}

// ---- grid query, warp-synchronous insertion --------------------------------------------------------
// The kernel above spends its instructions on DIVERGENT insertions: a lane inserts at ~56 of its ~195 candidates
// (k (1 + ln(n/k))), but the 32 lanes of a warp insert at different candidates, so the warp executes the
// insertion network at nearly every candidate (12 of 32 lanes active on average).  Here a candidate that beats the
// lane's current k-th key is only PUSHED onto a small per-lane stack in shared memory; the warp runs the
// insertion network when some lane's stack is nearly full, and then every lane with a pending key inserts one
// (LIFO: the order of insertions does not change the final set, keys are distinct; a key that no longer beats
// the k-th is dropped at the pop).  ~85 executions of the network per warp instead of ~200, each at ~25 lanes.
// Control flow is warp-uniform (votes), every lane walks its own sequence of candidate ranges and never idles
// before its shell is exhausted; the per-shell termination test is unchanged (after draining the stacks).
constexpr int KNN_PEND = 8;  // stack depth; a step pushes at most 4, the network runs while a stack holds > 4
#ifndef KNN_SYNC_MIN_BLOCKS
#define KNN_SYNC_MIN_BLOCKS 1
#endif
template <int KT>
__global__ void __launch_bounds__(128, KT <= 16 ? KNN_SYNC_MIN_BLOCKS : 1)
    knn_grid_query_sync_kernel(const KnnGrid *__restrict__ gp, const float4 *__restrict__ sorted,
                               const int *__restrict__ start, const float *__restrict__ queries, int nq, int k,
                               int do_sqrt, float *__restrict__ dist, long long *__restrict__ idx64,
                               int *__restrict__ idx32, const KnnEdgeOut eo) {
  __shared__ unsigned long long pend[KNN_PEND][128];
  constexpr unsigned FULL = 0xffffffffu;
  const KnnGrid g = *gp;
  const int tid = threadIdx.x;
  const int s = blockIdx.x * blockDim.x + tid;
  const bool live = s < nq;
  if (eo.tgt && s == 0) {  // the sentinel row N: only edges to N
    const int none = eo.rank ? (int)((((unsigned)nq >> 5) << 7) | ((unsigned)nq & 31u)) : nq;
    for (int e = 0; e < (1 << eo.slot_bits); ++e)
      eo.tgt[((size_t)nq << eo.slot_bits) + e] = none, eo.len[((size_t)nq << eo.slot_bits) + e] = 0.f;
  }
  float qx = 0.f, qy = 0.f, qz = 0.f;
  int row = 0;
  if (live) {
    if (queries == nullptr) {
      float4 me = sorted[s];
      qx = me.x, qy = me.y, qz = me.z;
      row = __float_as_int(me.w);
    } else {
      qx = __ldg(queries + (size_t)s * 3), qy = __ldg(queries + (size_t)s * 3 + 1), qz = __ldg(queries + (size_t)s * 3 + 2);
      row = s;
    }
  }
  const int cx = cell_coord(qx, g.ox, g.inv_h, g.dx), cy = cell_coord(qy, g.oy, g.inv_h, g.dy),
            cz = cell_coord(qz, g.oz, g.inv_h, g.dz);
  TopK<KT> top;
  top.init(k);
  int npend = 0;
  auto insert_round = [&]() {  // every lane with a pending key inserts its newest one
    if (npend > 0) top.offer_key(pend[--npend][tid]);
  };
  const int rmax = max(g.dx, max(g.dy, g.dz));
  bool active = live;
  for (int R = 0; R <= rmax; ++R) {
    if (!__any_sync(FULL, active)) break;
    const int z0 = max(cz - R, 0), z1 = min(cz + R, g.dz - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, g.dy - 1);
    const int xl = cx - R, xr = cx + R;
    // this lane's walk over the candidate ranges of shell R: rows (z, y) of the block; on a face of the shell the
    // whole x-run belongs to it (one range), otherwise only its two end cells (two ranges)
    int z = z0, y = y0, sg = 0, p = 0, p1 = 0;
    bool more = active;
    for (;;) {
      if (more && p >= p1) {
        for (;;) {
          if (z > z1) {
            more = false;
            break;
          }
          const bool face = R == 0 || z == cz - R || z == cz + R || y == cy - R || y == cy + R;
          const int rowbase = (z * g.dy + y) * g.dx;
          int xa, xb;
          bool last;
          if (face)
            xa = max(xl, 0), xb = min(xr, g.dx - 1), last = true;
          else
            xa = xb = sg == 0 ? xl : xr, last = sg == 1;
          if (last) {
            sg = 0;
            if (++y > y1) y = y0, ++z;
          } else {
            sg = 1;
          }
          if (xa < 0 || xb >= g.dx || xa > xb) continue;
          p = __ldg(start + rowbase + xa), p1 = __ldg(start + rowbase + xb + 1);
          if (p < p1) break;
        }
      }
      if (!__any_sync(FULL, more)) break;
      if (more) {  // up to four candidates of this lane's current range
        const int n = min(p1 - p, 4);
        float4 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u < n) c[u] = __ldg(sorted + p + u);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u < n) {
            const float d2 = sq3(c[u].x - qx, c[u].y - qy, c[u].z - qz);
            const unsigned long long kk = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(c[u].w);
            if (kk < top.key[KT - 1]) pend[npend++][tid] = kk;
          }
        p += 4;
      }
      while (__any_sync(FULL, npend > KNN_PEND - 4)) insert_round();
    }
    while (__any_sync(FULL, npend > 0)) insert_round();
    if (active) {
      // distance from the query to the nearest face of the visited block that still has cells
      // behind it; every unvisited point is at least that far away (minus the fp32 slack)
      float mfd = __int_as_float(0x7f800000);
      if (cx - R > 0) mfd = fminf(mfd, qx - (g.ox + (float)(cx - R) * g.h));
      if (cx + R + 1 < g.dx) mfd = fminf(mfd, (g.ox + (float)(cx + R + 1) * g.h) - qx);
      if (cy - R > 0) mfd = fminf(mfd, qy - (g.oy + (float)(cy - R) * g.h));
      if (cy + R + 1 < g.dy) mfd = fminf(mfd, (g.oy + (float)(cy + R + 1) * g.h) - qy);
      if (cz - R > 0) mfd = fminf(mfd, qz - (g.oz + (float)(cz - R) * g.h));
      if (cz + R + 1 < g.dz) mfd = fminf(mfd, (g.oz + (float)(cz + R + 1) * g.h) - qz);
      if (mfd == __int_as_float(0x7f800000)) active = false;  // the whole grid has been visited
      const float ms = mfd - g.slack;
      if (ms > 0.f && top.kth_d2() < ms * ms) active = false;
    }
  }
  if (!live) return;
  if (dist || idx64 || idx32)
    top.store(k, do_sqrt != 0, dist ? dist + (size_t)row * k : nullptr, idx64 ? idx64 + (size_t)row * k : nullptr,
              idx32 ? idx32 + (size_t)row * k : nullptr);
  if (eo.tgt) {  // self query only (nq == N): the propagation's edge rows, written straight from the registers
    const size_t er = (size_t)(eo.rank ? s : row) << eo.slot_bits;  // cell-order row (= this thread) or original row
    top.store_edges(k, eo.radius, nq, eo.slot_bits, eo.rank, eo.tgt + er, eo.len + er);
  }
}

// ---- brute force (algo 1) -----------------------------------------------------------------------
constexpr int BF_TILE = 1024;
template <int KT>
__global__ void __launch_bounds__(128)
    knn_brute_kernel(const float *__restrict__ xyz, int N, const float *__restrict__ queries, int nq, int k,
                     int do_sqrt, float *__restrict__ dist, long long *__restrict__ idx64, int *__restrict__ idx32) {
  __shared__ float sx[BF_TILE], sy[BF_TILE], sz[BF_TILE];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = q < nq;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) qx = __ldg(queries + (size_t)q * 3), qy = __ldg(queries + (size_t)q * 3 + 1), qz = __ldg(queries + (size_t)q * 3 + 2);
  TopK<KT> top;
  top.init(k);
  for (int base = 0; base < N; base += BF_TILE) {
    const int tn = min(BF_TILE, N - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tn * 3; t += blockDim.x) {
      float v = __ldg(xyz + (size_t)base * 3 + t);
      int pt = t / 3, ax = t - pt * 3;
      (ax == 0 ? sx : ax == 1 ? sy : sz)[pt] = v;
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int t = 0; t < tn; ++t) {
        float d2 = sq3(sx[t] - qx, sy[t] - qy, sz[t] - qz);
        top.offer(d2, base + t);
      }
    }
  }
  if (live)
    top.store(k, do_sqrt != 0, dist ? dist + (size_t)q * k : nullptr, idx64 ? idx64 + (size_t)q * k : nullptr,
              idx32 ? idx32 + (size_t)q * k : nullptr);
}

// ---- host orchestration -------------------------------------------------------------------------
// table sizes depend on N only (the workspace is sized before k is known)
static int knn_cap_fine(int N) {
  long long c = 8ll * N;
  if (c < (1 << 18)) c = 1 << 18;
  if (c > KNN_MAX_CELLS) c = KNN_MAX_CELLS;
  return (int)c;
}
static int knn_cap_coarse(int N) {
  int c = N < 4096 ? 4096 : N;
  return c > KNN_PASS0_CELLS ? KNN_PASS0_CELLS : c;
}

size_t knn_grid_workspace_bytes(int N) {
  size_t b = 0;
  b += align256(sizeof(KnnCtl));
  b += align256(sizeof(int) * (size_t)knn_cap_coarse(N));
  b += align256(sizeof(int) * (size_t)(knn_cap_fine(N) + 4)) * 2;  // cell_count, cell_start
  b += align256(sizeof(int) * (size_t)N) * 4;                        // cell_id, slot_in_cell, order, rank
  b += align256(sizeof(float4) * (size_t)N);
  return b + 1024;
}

template <int KT>
static void launch_grid_query(const KnnGrid *g, const float4 *sorted, const int *start, const float *queries, int nq,
                              int k, int do_sqrt, float *dist, long long *i64, int *i32, const KnnEdgeOut &eo,
                              cudaStream_t st) {
  // GF_KNN_SYNC=0|1 forces one kernel for every k.  Measured (c2, B200): the two execute the same number of warp
  // instructions at k = 16 (the saved insertions are spent on votes and the divergent range walk) and the
  // synchronous one is 10 % slower alone, 1 % slower in the batched pipeline; at k = 64, where an insertion is 64
  // compare-exchanges, it is 10-16 % faster.
  static const int force = [] {
    const char *e = getenv("GF_KNN_SYNC");
    return e ? (e[0] == '0' ? 0 : 1) : -1;
  }();
  const bool sync = force >= 0 ? force == 1 : KT > 32;
  if (sync)
    knn_grid_query_sync_kernel<KT><<<(nq + 127) / 128, 128, 0, st>>>(g, sorted, start, queries, nq, k, do_sqrt, dist, i64,
                                                                     i32, eo);
  else
    knn_grid_query_kernel<KT><<<(nq + 127) / 128, 128, 0, st>>>(g, sorted, start, queries, nq, k, do_sqrt, dist, i64, i32,
                                                                eo);
}
template <int KT>
static void launch_brute(const float *xyz, int N, const float *queries, int nq, int k, int do_sqrt, float *dist,
                         long long *i64, int *i32, cudaStream_t st) {
  knn_brute_kernel<KT><<<(nq + 127) / 128, 128, 0, st>>>(xyz, N, queries, nq, k, do_sqrt, dist, i64, i32);
}

// Five launches and one memset (round 1: thirteen launches): bbox + coarse plan | coarse count + final plan |
// final count | scan | scatter.  Nothing returns to the host; grid dimensions are read from the device.
int knn_grid_build(const float *xyz, int N, int k, void *workspace, size_t workspace_bytes, cudaStream_t st,
                   KnnGridBuffers *out) {
  const int cap_fine = knn_cap_fine(N), cap_coarse = knn_cap_coarse(N);
  Arena a(workspace, workspace_bytes);
  KnnCtl *ctl = a.take<KnnCtl>(1);
  KnnGrid *g = &ctl->g;
  int *coarse_count = a.take<int>(cap_coarse);
  int *cell_count = a.take<int>(cap_fine + 4);
  int *cell_start = a.take<int>(cap_fine + 4);
  int *cell_id = a.take<int>(N);
  int *slot = a.take<int>(N);
  int *order = a.take<int>(N);
  int *rank = a.take<int>(N);
  float4 *sorted = a.take<float4>(N);
  if (!a.ok) {
    set_error("knn: workspace too small (%zu bytes given, %zu needed)", workspace_bytes, knn_grid_workspace_bytes(N));
    return GF_ERR_WORKSPACE;
  }
  const int nb = num_sms() * 8;
  // target cell occupancy = tau_mul x k.  Measured (c2 / c1, B200): flat between 0.12 and 0.7 for k <= 32 (0.45 kept);
  // at k = 64 smaller cells win (0.2: 0.88 ms vs 1.03 ms per graph).  GF_KNN_TAU overrides (experiments).
  static const float tau_env = [] {
    const char *e = getenv("GF_KNN_TAU");
    const float v = e ? (float)atof(e) : 0.f;
    return v > 0.f ? v : 0.f;
  }();
  const float tau_mul = tau_env > 0.f ? tau_env : (k > 32 ? 0.2f : 0.45f);
  const float tau = fmaxf(2.f, tau_mul * (float)k);
  const int npt = min(nb, (N + 255) / 256);
  GF_CUDA(cudaMemsetAsync(ctl, 0, sizeof(KnnCtl), st));
  knn_bbox_kernel<<<nb, 256, 0, st>>>(xyz, N, g, tau, cap_coarse, (int4 *)coarse_count, (cap_coarse + 3) / 4,
                                      (int4 *)cell_count, (cap_fine + 4) / 4);
  GF_LAUNCHED();
  knn_count_kernel<<<npt, 256, 0, st>>>(xyz, N, g, coarse_count, nullptr, nullptr, tau, cap_fine);
  GF_LAUNCHED();
  knn_count_kernel<<<npt, 256, 0, st>>>(xyz, N, g, cell_count, cell_id, slot, tau, cap_fine);
  GF_LAUNCHED();
  knn_scan_kernel<<<KNN_SCAN_BLOCKS, 256, 0, st>>>(ctl, cell_count, cell_start, N);
  GF_LAUNCHED();
  knn_scatter_kernel<<<npt, 256, 0, st>>>(xyz, N, cell_id, slot, cell_start, sorted, order, rank);
  GF_LAUNCHED();
  out->grid = g;
  out->cell_start = cell_start;
  out->sorted = sorted;
  out->order = order;
  out->rank = rank;
  return GF_OK;
}

int knn_grid_query(const KnnGridBuffers &b, const float *queries, int nq, int k, int do_sqrt, float *dist,
                   long long *idx64, int *idx32, cudaStream_t st, const KnnEdgeOut *edges) {
  const KnnGrid *g = (const KnnGrid *)b.grid;
  KnnEdgeOut eo = {nullptr, nullptr, 0.f, 0, nullptr};
  if (edges && queries == nullptr) {
    eo = *edges;
    if (eo.rank) eo.rank = b.rank;  // any non-null value asks for the cell-order format
  }
  if (k <= 8)
    launch_grid_query<8>(g, b.sorted, b.cell_start, queries, nq, k, do_sqrt, dist, idx64, idx32, eo, st);
  else if (k <= 16)
    launch_grid_query<16>(g, b.sorted, b.cell_start, queries, nq, k, do_sqrt, dist, idx64, idx32, eo, st);
  else if (k <= 32)
    launch_grid_query<32>(g, b.sorted, b.cell_start, queries, nq, k, do_sqrt, dist, idx64, idx32, eo, st);
  else
    launch_grid_query<64>(g, b.sorted, b.cell_start, queries, nq, k, do_sqrt, dist, idx64, idx32, eo, st);
  GF_LAUNCHED();
  return GF_OK;
}

}  // namespace gf

using namespace gf;

extern "C" size_t gf_knn_workspace_bytes(int N, int nq, int k, int algo) {
  (void)nq;
  (void)k;
  if (algo == 1 || N <= 0) return 0;
  return knn_grid_workspace_bytes(N);
}

extern "C" int gf_knn(const float *xyz, int N, const float *queries, int nq, int k, int sqrt_out, float *dist,
                      int64_t *idx64, int32_t *idx32, int algo, void *workspace, size_t workspace_bytes,
                      void *stream) {
  GF_CHECK_ARG(N >= 0 && nq >= 0, "knn: negative size");
  GF_CHECK_ARG(k >= 1 && k <= KNN_MAX_K, "knn: k=%d outside [1,%d]", k, KNN_MAX_K);
  GF_CHECK_ARG(algo == 0 || algo == 1, "knn: unknown algo %d", algo);
  cudaStream_t st = (cudaStream_t)stream;
  if (queries == nullptr) nq = N;
  if (nq == 0) return GF_OK;
  GF_CHECK_ARG(xyz || N == 0, "knn: null xyz");
  if (algo == 1 || N == 0) {
    const float *q = queries ? queries : xyz;
    if (k <= 8)
      launch_brute<8>(xyz, N, q, nq, k, sqrt_out, dist, (long long *)idx64, idx32, st);
    else if (k <= 16)
      launch_brute<16>(xyz, N, q, nq, k, sqrt_out, dist, (long long *)idx64, idx32, st);
    else if (k <= 32)
      launch_brute<32>(xyz, N, q, nq, k, sqrt_out, dist, (long long *)idx64, idx32, st);
    else
      launch_brute<64>(xyz, N, q, nq, k, sqrt_out, dist, (long long *)idx64, idx32, st);
    GF_LAUNCHED();
    return GF_OK;
  }
  KnnGridBuffers b;
  int rc = knn_grid_build(xyz, N, k, workspace, workspace_bytes, st, &b);
  if (rc) return rc;
  return knn_grid_query(b, queries, nq, k, sqrt_out, dist, (long long *)idx64, idx32, st);
}
