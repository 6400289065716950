/*
 * oracle.c -- CPU restatement of the GeoFormer geodesic-guidance hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library.  Nothing under geoformer_b200/
 * links, imports or calls it; the product path fails loudly without its CUDA library.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Arithmetic notes that matter for index-exactness:
 *   - the reference CUDA kernels are compiled by nvcc with its default -fmad=true, so
 *     a*a + b*b + c*c becomes FMUL, FFMA, FFMA (checked in the SASS of the sm_100 build,
 *     see DESIGN.md "FMA contraction").  This file is compiled with -ffp-contract=off and
 *     spells the contraction out with fmaf().
 *   - PARITY PIN: FPS / ball_query / gather / group / three_* are pinned against the
 *     reference's own kernels (oracle/_ref, run on the GPU box, tests/test_gpu_vs_reference_ext.py
 *     and the committed fixtures tests/golden/ref_ext_*.npz).  The geodesic is pinned against
 *     the reference's own cal_geodesic_vectorize run on CPU (tests/golden/geodesic_*.npz,
 *     made by tests/golden/make_golden.py).  kNN: PARITY UNPINNED -- the reference uses
 *     faiss-gpu (un-vendored, un-pinned, docs/INSTALL.md:72); the canonical exact fp32 kNN
 *     of SURVEY.md App. A.4 is restated here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* nvcc's contraction of  ax*ax + ay*ay + az*az  as seen in the SASS of the sm_100 build of
 * sampling_gpu.cu / ball_query_gpu.cu / interpolate_gpu.cu:
 *     FMUL t, y, y ;  FFMA t, x, x, t ;  FFMA t, z, z, t
 * i.e. the MIDDLE product is the rounded one:  fma(az,az, fma(ax,ax, ay*ay)). */
static inline float sq3(float ax, float ay, float az) {
  return fmaf(az, az, fmaf(ax, ax, ay * ay));
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The timing legs of bench.py set the team size themselves: launchers such as torch.distributed.run
 * export OMP_NUM_THREADS=1 to every rank, which would silently turn the host baseline into one core. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* lib/pointnet2/_ext_src/include/cuda_utils.h:15-21  opt_n_threads */
static int opt_n_threads(int work_size) {
  int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}

/* ------------------------------------------------------------------------------------------
 * FPS.  sampling_gpu.cu:72-176 (kernel), sampling.cpp:67-88 (host: idx zero-init, temp=1e10).
 * The block of `bs` threads is simulated literally: thread tid scans k = tid, tid+bs, ... with a
 * strict '>' (:111-112); the shared-memory tree (:118-171) keeps the lower slot on ties (:62-68).
 * ---------------------------------------------------------------------------------------- */
void orc_fps(const float *xyz, int B, int N, int m, int *idx) {
  if (m <= 0) return; /* :76 */
  int bs = opt_n_threads(N);
  float *temp = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
  float dists[512];
  int dists_i[512];
  for (int b = 0; b < B; ++b) {
    const float *ds = xyz + (size_t)b * N * 3;
    int *out = idx + (size_t)b * m;
    for (int k = 0; k < N; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
      for (int t = 0; t < bs; ++t) {
        dists[t] = -1.0f;
        dists_i[t] = 0;
      }
      for (int k = 0; k < N; ++k) {
        int t = k % bs;
        float x2 = ds[k * 3 + 0], y2 = ds[k * 3 + 1], z2 = ds[k * 3 + 2];
        float mag = sq3(x2, y2, z2);
        if ((double)mag <= 1e-3) continue; /* :103-104, compare done in double */
        float d = sq3(x2 - x1, y2 - y1, z2 - z1);
        float d2 = fminf(d, temp[k]);
        temp[k] = d2;
        if (d2 > dists[t]) { /* :111-112 */
          dists_i[t] = k;
          dists[t] = d2;
        }
      }
      for (int s = bs / 2; s >= 1; s >>= 1) { /* :118-171 */
        for (int t = 0; t < s; ++t) {
          float v1 = dists[t], v2 = dists[t + s];
          int i1 = dists_i[t], i2 = dists_i[t + s];
          dists[t] = v1 > v2 ? v1 : (v2 > v1 ? v2 : v1); /* max(v1,v2) on non-NaN data */
          dists_i[t] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(temp);
}

/* sampling_gpu.cu:11-23   out[b,c,j] = points[b,c,idx[b,j]] */
void orc_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        out[((size_t)b * C + c) * m + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * m + j]];
}

/* sampling_gpu.cu:37-50   atomicAdd scatter; serial j order here (fp32 add order is not
 * reproducible in the reference either) */
void orc_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        grad_points[((size_t)b * C + c) * N + idx[(size_t)b * m + j]] += grad_out[((size_t)b * C + c) * m + j];
}

/* ball_query_gpu.cu:12-47; host ball_query.cpp:11-35 (idx zero-init).  Centres first. */
void orc_ball_query(const float *new_xyz, const float *xyz, int B, int N, int m, float radius, int nsample, int *idx) {
  float r2 = radius * radius;
  memset(idx, 0, sizeof(int) * (size_t)B * m * nsample);
  for (int b = 0; b < B; ++b) {
    const float *P = xyz + (size_t)b * N * 3;
    const float *Cn = new_xyz + (size_t)b * m * 3;
    int *o = idx + (size_t)b * m * nsample;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < m; ++j) {
      float nx = Cn[j * 3 + 0], ny = Cn[j * 3 + 1], nz = Cn[j * 3 + 2];
      int cnt = 0;
      for (int k = 0; k < N && cnt < nsample; ++k) {
        float d2 = sq3(nx - P[k * 3 + 0], ny - P[k * 3 + 1], nz - P[k * 3 + 2]);
        if (d2 < r2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[(size_t)j * nsample + l] = k;
          o[(size_t)j * nsample + cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* group_points_gpu.cu:11-31   out[b,c,j,s] = points[b,c,idx[b,j,s]] */
void orc_group_points(const float *points, const int *idx, int B, int C, int N, int np, int ns, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < np; ++j)
        for (int s = 0; s < ns; ++s)
          out[(((size_t)b * C + c) * np + j) * ns + s] =
              points[((size_t)b * C + c) * N + idx[((size_t)b * np + j) * ns + s]];
}

/* group_points_gpu.cu:46-67 */
void orc_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int np, int ns, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < np; ++j)
        for (int s = 0; s < ns; ++s)
          grad_points[((size_t)b * C + c) * N + idx[((size_t)b * np + j) * ns + s]] +=
              grad_out[(((size_t)b * C + c) * np + j) * ns + s];
}

/* interpolate_gpu.cu:12-62.  best1..3 are doubles initialised to 1e40 and compared against a
 * float d (:38-54); written back as float (:56-58) -> +inf when fewer than 3 known points. */
void orc_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx) {
  for (int b = 0; b < B; ++b) {
    const float *U = unknown + (size_t)b * n * 3;
    const float *K = known + (size_t)b * m * 3;
    for (int j = 0; j < n; ++j) {
      float ux = U[j * 3 + 0], uy = U[j * 3 + 1], uz = U[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int b1 = 0, b2 = 0, b3 = 0;
      for (int k = 0; k < m; ++k) {
        float d = sq3(ux - K[k * 3 + 0], uy - K[k * 3 + 1], uz - K[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
        } else if (d < best2) {
          best3 = best2; b3 = b2; best2 = d; b2 = k;
        } else if (d < best3) {
          best3 = d; b3 = k;
        }
      }
      size_t o = ((size_t)b * n + j) * 3;
      dist2[o + 0] = (float)best1; dist2[o + 1] = (float)best2; dist2[o + 2] = (float)best3;
      idx[o + 0] = b1; idx[o + 1] = b2; idx[o + 2] = b3;
    }
  }
}

/* interpolate_gpu.cu:75-104.  nvcc contracts p1*w1 + p2*w2 + p3*w3 to
 * fma(p3,w3, fma(p1,w1, p2*w2))  (SASS: FMUL on the +4 operands, then FFMA +0, FFMA +8). */
void orc_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m, int n, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < n; ++j) {
        size_t o = ((size_t)b * n + j) * 3;
        const float *P = points + ((size_t)b * C + c) * m;
        out[((size_t)b * C + c) * n + j] =
            fmaf(P[idx[o + 2]], weight[o + 2], fmaf(P[idx[o + 0]], weight[o + 0], P[idx[o + 1]] * weight[o + 1]));
      }
}

/* interpolate_gpu.cu:119-146 */
void orc_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C, int n, int m, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * m);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < n; ++j) {
        size_t o = ((size_t)b * n + j) * 3;
        float g = grad_out[((size_t)b * C + c) * n + j];
        float *G = grad_points + ((size_t)b * C + c) * m;
        G[idx[o + 0]] += g * weight[o + 0];
        G[idx[o + 1]] += g * weight[o + 1];
        G[idx[o + 2]] += g * weight[o + 2];
      }
}

/* ------------------------------------------------------------------------------------------
 * Canonical exact kNN (replaces faiss.GpuIndexFlatL2, call sites
 * model/geoformer/geodesic_utils.py:18-21, geoformer_fs.py:170-175).  PARITY UNPINNED (see header).
 * For each query row i of `q` (nq,3) against the database `x` (N,3): order all j by (d2, j),
 * d2 = fmaf(dz,dz, fmaf(dx,dx, dy*dy)) with dx = x_j - q_i (same contraction as the pointnet2
 * kernels, see sq3); keep the first k.
 * D2 receives SQUARED distances (faiss contract; find_knn applies sqrt at geodesic_utils.py:22).
 * Missing neighbours (N < k): I = -1, D2 = +inf.
 * row0/nrows: optional row window so that a bounded sample can be timed.
 * ---------------------------------------------------------------------------------------- */
void orc_knn(const float *x, int N, const float *q, int nq, int k, float *D2, int64_t *I) {
#pragma omp parallel
  {
    float *bd = (float *)malloc(sizeof(float) * (size_t)k);
    int *bi = (int *)malloc(sizeof(int) * (size_t)k);
    float *d2row = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float qx = q[i * 3 + 0], qy = q[i * 3 + 1], qz = q[i * 3 + 2];
      for (int j = 0; j < N; ++j) d2row[j] = sq3(x[j * 3 + 0] - qx, x[j * 3 + 1] - qy, x[j * 3 + 2] - qz);
      int cnt = 0;
      for (int j = 0; j < N; ++j) {
        float d = d2row[j];
        /* ascending j: an equal d2 never displaces an earlier (smaller) index */
        if (cnt == k && !(d < bd[k - 1])) continue;
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && d < bd[pos - 1]) {
          bd[pos] = bd[pos - 1];
          bi[pos] = bi[pos - 1];
          --pos;
        }
        bd[pos] = d;
        bi[pos] = j;
        if (cnt < k) ++cnt;
      }
      for (int s = 0; s < k; ++s) {
        D2[(size_t)i * k + s] = s < cnt ? bd[s] : INFINITY;
        I[(size_t)i * k + s] = s < cnt ? (int64_t)bi[s] : (int64_t)-1;
      }
    }
    free(bd);
    free(bi);
    free(d2row);
  }
}

/* ------------------------------------------------------------------------------------------
 * Multi-source geodesic = level-synchronous, first-visit-wins BFS.
 * model/geoformer/geodesic_utils.py:91-164 (cal_geodesic_vectorize), restated per SURVEY App. A.5.
 *   D (N,k) f32: kNN distances (already sqrt'ed, geodesic_utils.py:22), I (N,k) i64; column 0 is
 *   dropped (:110-111).  geo (Q,N) f32, -1 = unreachable (:113).
 * Per level the reference dedupes candidates with unique_with_inds (:4-8,:131-136) which keeps the
 * FIRST occurrence in candidate order; candidate order for a fixed query is (parent index
 * ascending, slot ascending) because the winner list is the lexicographically sorted unique
 * (point, query) list (:131-135) and torch.nonzero is row-major (:154).  Seeds are independent,
 * so the restatement runs one query at a time.
 * Returns the number of reached (q,p) pairs excluding the seeds' own zero entries that were
 * never re-written (R of SURVEY 8(d)) through *reached.
 * ---------------------------------------------------------------------------------------- */
static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

void orc_geodesic(const float *D, const int64_t *I, int N, int k, const int64_t *seeds, int Q, float radius,
                  int max_step, float *geo, int64_t *reached, int *levels_run) {
  int K = k - 1;
  int64_t total = 0;
  int maxlev = 0;
#pragma omp parallel reduction(+ : total) reduction(max : maxlev)
  {
    size_t n1 = (size_t)(N > 0 ? N : 1);
    unsigned char *vis = (unsigned char *)malloc(n1);
    int *claimed = (int *)malloc(sizeof(int) * n1);     /* level stamp of the last claim */
    float *cand_d = (float *)malloc(sizeof(float) * n1); /* distance carried by the winning candidate */
    int *front = (int *)malloc(sizeof(int) * n1);
    int *cand = (int *)malloc(sizeof(int) * n1);
#pragma omp for schedule(dynamic, 1)
    for (int q = 0; q < Q; ++q) {
      float *g = geo + (size_t)q * N;
      for (int t = 0; t < N; ++t) {
        g[t] = -1.0f; /* :113 */
        vis[t] = 0;   /* :114 */
        claimed[t] = 0;
      }
      int s = (int)seeds[q];
      g[s] = 0.0f; /* :118 */
      vis[s] = 1;  /* :119 */
      /* level-1 candidates from the seed row, NO visited filter (:121-127) */
      int nc = 0;
      int level = 1;
      for (int j = 0; j < K; ++j) {
        float d = D[(size_t)s * k + 1 + j];
        int64_t t = I[(size_t)s * k + 1 + j];
        if (d <= radius && t >= 0 && claimed[t] != level) { /* first occurrence wins (:131-136) */
          claimed[t] = level;
          cand_d[t] = d;
          cand[nc++] = (int)t;
        }
      }
      for (int step = 0; step < max_step; ++step) { /* :129 */
        /* :131-140  commit the (already deduped) winners of this level */
        qsort(cand, (size_t)nc, sizeof(int), cmp_int); /* unique() output is sorted by point */
        int nf = nc;
        memcpy(front, cand, sizeof(int) * (size_t)nf);
        for (int a = 0; a < nf; ++a) {
          g[front[a]] = cand_d[front[a]]; /* :139 */
          vis[front[a]] = 1;              /* :140 */
        }
        total += nf;
        if (nf > 0 && level > maxlev) maxlev = level;
        /* :143-154  new candidates: parent ascending, slot ascending, visited filter */
        nc = 0;
        ++level;
        for (int a = 0; a < nf; ++a) {
          int p = front[a];
          float dp = g[p];
          for (int j = 0; j < K; ++j) {
            float d = D[(size_t)p * k + 1 + j];
            int64_t t = I[(size_t)p * k + 1 + j];
            if (d <= radius && t >= 0 && !vis[t] && claimed[t] != level) {
              claimed[t] = level;
              cand_d[t] = d + dp; /* :144, one fp32 add */
              cand[nc++] = (int)t;
            }
          }
        }
        if (nc == 0) break; /* :156-157 */
      }
    }
    free(vis);
    free(claimed);
    free(cand_d);
    free(front);
    free(cand);
  }
  if (reached) *reached = total;
  if (levels_run) *levels_run = maxlev;
}
