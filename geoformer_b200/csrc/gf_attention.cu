// Decoder vector cross-attention over the geodesic relative-position embedding, for sm_100a (tcgen05 / TMEM).
//
// Reference: model/transformer_detr.py:443-454 (TransformerDecoderLayer.forward_pre_rel, MLPs defined at :384-396):
//     x    = tgt2[q] - memory[c] + relative_pos[q,c]                  (64)      for every (query q, context c)
//     sim  = W2 relu(W1 x + b1) + b2                                   attn_mlp
//     attn = softmax(sim / sqrt(64), over the CONTEXTS, per channel)   :449 (dim=1)
//     v2   = Wv (memory[c] + relative_pos[q,c]) + bv                   v_mlp
//     out  = relu(Wo (sum_c attn * v2) + bo)                           :452-453 (einsum, out_mlp)
// The reference materialises three (Q,C,B,64) tensors besides the embedding (134 MB each at Q=256, C=2048).
// Here one CTA owns one (query, batch element) and walks the contexts in tiles of 128:
//   * The (q,c)-dependent part of both first layers is ONE tensor-core product per tile: [Wv; W1] (128 x 64) times
//     the tile of the embedding (128 contexts x 64), M = 128, N = 128, K = 64 -> TMEM columns 0..127.  The query
//     part W1 tgt2[q] and the context parts W1 memory[c], Wv memory[c] are small matrices computed once per call
//     (aq, P1, Pv) and added in the epilogue: W1 x = W1 tgt2[q] - W1 memory[c] + W1 rel.
//   * Channel-major accumulators: TMEM lane = output channel, column = context.  Lanes 0..63 hold v2, lanes 64..127
//     the hidden layer; the second product [W2; 0] times the hidden tile lands in lanes 0..63 of columns 128..255,
//     i.e. in the SAME threads that hold v2 -- so the softmax over the contexts and the weighted sum are plain
//     per-thread loops over TMEM columns (online softmax, no shuffles, nothing written back).
//   * Operands are fp32 in shared memory, multiplied as TF32 (kind::tf32), accumulated in fp32; K-major tiles in
//     the 128-byte swizzle the UMMA descriptors describe, written by the threads themselves (the embedding is
//     either loaded from the (Q,C,B,64) tensor or -- fused with the decoder epilogue a10 -- computed on the fly
//     from the geodesic maps, so that it never exists in memory).
// tcgen05.mma is issued by one thread; completion is signalled through an mbarrier by tcgen05.commit.
#include "gf_common.cuh"

namespace gf {

constexpr int ATT_D = 64;        // channels (dec_dim of the model, geoformer_fs.py:116)
constexpr int ATT_TILE = 128;    // contexts per tile = MMA N
constexpr int ATT_THREADS = 128;
constexpr int ATT_MAX_B = 8;     // batch elements of the fused variant (map pointers travel in the kernel parameters)
constexpr int ATT_HALF_BYTES = ATT_TILE * 128;  // one K-half (32 floats = 128 bytes per row) of a 128-row tile
constexpr int ATT_TILE_BYTES = 2 * ATT_HALF_BYTES;
constexpr int ATT_SMEM = 4 * ATT_TILE_BYTES + 1024;  // A1, A2, embedding tile, hidden tile (+ alignment slack)

struct AttArgs {
  const float *aq;   // (B, Q, 64)  W1 tgt2[q] + b1
  const float *p1;   // (B, C, 64)  W1 memory[c]
  const float *pv;   // (B, C, 64)  Wv memory[c] + bv
  const float *w1, *w2, *wv, *wo;  // (64, 64) row-major (out, in): nn.Linear.weight
  const float *b2, *bo;            // (64)
  int Q, C, B;
  // the embedding: either the tensor (Q, C, B, 64) ...
  const float *rel;
  // ... or its ingredients (decoder epilogue a10 + Fourier features, gf_bias.cu: bias_ctx_fourier_kernel)
  const float *geo[ATT_MAX_B];  // (Q, N_b) maps, one per batch element
  int geo_ld[ATT_MAX_B];        // their row strides
  const int *ctx_idx;            // (B, C)
  const float *query_xyz, *ctx_xyz;  // (B, Q, 3), (B, C, 3)
  const uint32_t *rowmax, *gmax;     // ordered-uint row maxima (B, Q) and the global maximum (bias_ctx_rowmax_kernel)
  const float *gauss_B;              // (3, ldb), 32 frequencies used
  int ldb;
  const float *pc_min, *pc_max;      // (B, 3)
  float *out;                        // (Q, B, 64)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major tile with 128-byte swizzle: row r, fp32 column k (0..63) -> byte offset inside the tile
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  const int half = k >> 5, kk = k & 31;
  return (uint32_t)(half * ATT_HALF_BYTES + r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2)));
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t ATT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((ATT_TILE >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(ATT_IDESC), "r"(accumulate)
      : "memory");
}

// D (128 x 128, TMEM) = A (128 x 64) * B^T (128 x 64): eight K = 8 steps, 32 bytes apart inside a swizzled row,
// the second half of K in the second half-tile
__device__ __forceinline__ void umma_tile(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint32_t koff = (uint32_t)((kk >> 2) * ATT_HALF_BYTES + (kk & 3) * 32);
    umma_tf32(d_tmem, umma_desc(a_saddr + koff), umma_desc(b_saddr + koff), kk > 0);
  }
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W;\n\t}"
      :
      : "r"(mbar), "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool FUSED>
__global__ void __launch_bounds__(ATT_THREADS, 1) rel_cross_attention_kernel(const AttArgs a) {
  extern __shared__ unsigned char att_smem_raw[];
  __shared__ __align__(8) unsigned long long s_mbar[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_out[ATT_D];
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  const int q = blockIdx.x, b = blockIdx.y;
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA1 = smem, *sA2 = smem + ATT_TILE_BYTES, *sR = smem + 2 * ATT_TILE_BYTES, *sH = smem + 3 * ATT_TILE_BYTES;
  const uint32_t mbar0 = smem_u32(&s_mbar[0]), mbar1 = smem_u32(&s_mbar[1]);

  // ---- one-time set-up: weights into swizzled K-major tiles, TMEM, barriers -----------------------------------
  // A1 rows 0..63 = Wv, rows 64..127 = W1;  A2 rows 0..63 = W2, rows 64..127 = 0
  for (int e = tid; e < 128 * ATT_D; e += ATT_THREADS) {
    const int r = e >> 6, k = e & 63;
    *reinterpret_cast<float *>(sA1 + sw128_off(r, k)) = r < 64 ? __ldg(a.wv + r * 64 + k) : __ldg(a.w1 + (r - 64) * 64 + k);
    *reinterpret_cast<float *>(sA2 + sw128_off(r, k)) = r < 64 ? __ldg(a.w2 + r * 64 + k) : 0.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the weight tiles were written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t t_lane = tmem + ((warp * 32u) << 16);  // this warp's 32 TMEM lanes

  // per-thread constants: threads 0..63 own output channel f = tid (v2, sim); threads 64..127 hidden channel i
  const int ch = tid & 63;
  const float aq = tid >= 64 ? __ldg(a.aq + ((size_t)b * a.Q + q) * 64 + ch) : 0.f;
  const float b2 = tid < 64 ? __ldg(a.b2 + ch) : 0.f;
  const float *p1 = a.p1 + (size_t)b * a.C * 64, *pv = a.pv + (size_t)b * a.C * 64;
  float m_run = -INFINITY, s_run = 0.f, acc = 0.f;  // online softmax over the contexts, per channel

  // ingredients of the fused embedding (gf_bias.cu: bias_ctx_fourier_kernel)
  float fm = 0.f, qx = 0.f, qy = 0.f, qz = 0.f, mn0 = 0.f, mn1 = 0.f, mn2 = 0.f, df0 = 1.f, df1 = 1.f, df2 = 1.f;
  const float *geo_row = nullptr;
  if (FUSED) {
    fm = ord2f(a.rowmax[b * a.Q + q]);
    if (fm < 0.f) fm = ord2f(*a.gmax);  // geoformer_fs.py:693
    qx = a.query_xyz[((size_t)b * a.Q + q) * 3 + 0], qy = a.query_xyz[((size_t)b * a.Q + q) * 3 + 1],
    qz = a.query_xyz[((size_t)b * a.Q + q) * 3 + 2];
    mn0 = a.pc_min[b * 3 + 0], mn1 = a.pc_min[b * 3 + 1], mn2 = a.pc_min[b * 3 + 2];
    df0 = __fsub_rn(a.pc_max[b * 3 + 0], mn0), df1 = __fsub_rn(a.pc_max[b * 3 + 1], mn1),
    df2 = __fsub_rn(a.pc_max[b * 3 + 2], mn2);
    geo_row = a.geo[b] + (size_t)q * a.geo_ld[b];
  }

  const int ntiles = (a.C + ATT_TILE - 1) / ATT_TILE;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int c0 = tile * ATT_TILE;
    const uint32_t parity = (uint32_t)(tile & 1);
    // ---- the embedding tile: thread t writes row t (context c0 + t), 64 fp32, swizzled ------------------------
    {
      const int c = c0 + (int)tid;
      if (!FUSED) {
        const float4 *src = reinterpret_cast<const float4 *>(a.rel + (((size_t)q * a.C + (c < a.C ? c : 0)) * a.B + b) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 v = c < a.C ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4 *>(sR + sw128_off((int)tid, 4 * j)) = v;
        }
      } else {
        float n0 = 0.f, n1 = 0.f, n2 = 0.f;
        if (c < a.C) {
          const float g = __ldg(geo_row + __ldg(a.ctx_idx + (size_t)b * a.C + c));
          float v0 = g, v1 = g, v2 = g;
          if (g < 0.f) {  // :699-702
            const float *cx = a.ctx_xyz + ((size_t)b * a.C + c) * 3;
            v0 = __fadd_rn(fm, fabsf(__fsub_rn(qx, cx[0])));
            v1 = __fadd_rn(fm, fabsf(__fsub_rn(qy, cx[1])));
            v2 = __fadd_rn(fm, fabsf(__fsub_rn(qz, cx[2])));
          }
          const float two_pi = 6.2831855f;
          n0 = __fmul_rn(__fdiv_rn(__fsub_rn(v0, mn0), df0), two_pi);
          n1 = __fmul_rn(__fdiv_rn(__fsub_rn(v1, mn1), df1), two_pi);
          n2 = __fmul_rn(__fdiv_rn(__fsub_rn(v2, mn2), df2), two_pi);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {  // four frequencies at a time: sin -> columns j, cos -> columns 32 + j
          float sn[4], cs[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            const float p = fmaf(n2, __ldg(a.gauss_B + 2 * a.ldb + j), fmaf(n1, __ldg(a.gauss_B + a.ldb + j), __fmul_rn(n0, __ldg(a.gauss_B + j))));
            sincosf(p, &sn[e], &cs[e]);
            if (c >= a.C) sn[e] = cs[e] = 0.f;
          }
          *reinterpret_cast<float4 *>(sR + sw128_off((int)tid, 4 * j4)) = make_float4(sn[0], sn[1], sn[2], sn[3]);
          *reinterpret_cast<float4 *>(sR + sw128_off((int)tid, 32 + 4 * j4)) = make_float4(cs[0], cs[1], cs[2], cs[3]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // tile complete; every thread has finished reading the previous tile's accumulators
    // ---- product 1: lanes 0..63 <- Wv rel, lanes 64..127 <- W1 rel (TMEM columns 0..127) ----------------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      umma_tile(tmem, smem_u32(sA1), smem_u32(sR));
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar0) : "memory");
    }
    mbar_wait(mbar0, parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- hidden layer (threads 64..127): h = relu(W1 rel + W1 tgt2[q] + b1 - W1 memory[c]) -> hidden tile --------
    if (tid >= 64) {
#pragma unroll 1
      for (int cc = 0; cc < ATT_TILE; cc += 32) {
        float d[32];
        tmem_ld32(t_lane + (uint32_t)cc, d);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int c = c0 + cc + e;
          const float h = c < a.C ? fmaxf(d[e] + aq - __ldg(p1 + (size_t)c * 64 + ch), 0.f) : 0.f;
          *reinterpret_cast<float *>(sH + sw128_off(cc + e, ch)) = h;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- product 2: lanes 0..63 <- W2 h (TMEM columns 128..255) ------------------------------------------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      umma_tile(tmem + 128u, smem_u32(sA2), smem_u32(sH));
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar1) : "memory");
    }
    mbar_wait(mbar1, parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- softmax over the contexts and weighted sum (threads 0..63: channel f = tid) ---------------------------
    if (tid < 64) {
#pragma unroll 1
      for (int cc = 0; cc < ATT_TILE; cc += 32) {
        float sv[32], vv[32];
        tmem_ld32(t_lane + 128u + (uint32_t)cc, sv);
        tmem_ld32(t_lane + (uint32_t)cc, vv);
        float mx = m_run;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          sv[e] = c0 + cc + e < a.C ? (sv[e] + b2) * 0.125f : -INFINITY;  // / sqrt(64), :449
          mx = fmaxf(mx, sv[e]);
        }
        if (mx > -INFINITY) {
          const float scale = __expf(m_run - mx);  // exp(-inf) = 0 on the first tile
          s_run *= scale, acc *= scale;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int c = c0 + cc + e;
            if (c < a.C) {
              const float w = __expf(sv[e] - mx);
              s_run += w;
              acc = fmaf(w, vv[e] + __ldg(pv + (size_t)c * 64 + ch), acc);
            }
          }
          m_run = mx;
        }
      }
    }
  }
  // ---- out_mlp: relu(Wo (acc / s) + bo) ---------------------------------------------------------------------------
  if (tid < 64) s_out[tid] = s_run > 0.f ? acc / s_run : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 64) {
    float o = __ldg(a.bo + tid);
#pragma unroll 8
    for (int f = 0; f < 64; ++f) o = fmaf(__ldg(a.wo + tid * 64 + f), s_out[f], o);
    a.out[((size_t)q * a.B + b) * 64 + tid] = fmaxf(o, 0.f);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace gf

using namespace gf;

// tables of the query / context parts (three small matrix products, fp32 on the CUDA cores):
//   aq[b,q,:] = W1 tgt2[q,b,:] + b1,   p1[b,c,:] = W1 memory[c,b,:],   pv[b,c,:] = Wv memory[c,b,:] + bv
__global__ void att_tables_kernel(const float *__restrict__ tgt2, const float *__restrict__ memory, int Q, int C, int B,
                                  const float *__restrict__ w1, const float *__restrict__ b1,
                                  const float *__restrict__ wv, const float *__restrict__ bv, float *__restrict__ aq,
                                  float *__restrict__ p1, float *__restrict__ pv) {
  __shared__ float x[64];
  const int row = blockIdx.x, b = blockIdx.y, o = threadIdx.x;  // rows 0..Q-1: queries, Q..Q+C-1: contexts
  const bool is_q = row < Q;
  const float *src = is_q ? tgt2 + ((size_t)row * B + b) * 64 : memory + ((size_t)(row - Q) * B + b) * 64;
  x[o] = src[o];
  __syncthreads();
  float s1 = 0.f, sv = 0.f;
#pragma unroll 8
  for (int i = 0; i < 64; ++i) {
    s1 = fmaf(__ldg(w1 + o * 64 + i), x[i], s1);
    if (!is_q) sv = fmaf(__ldg(wv + o * 64 + i), x[i], sv);
  }
  if (is_q) {
    aq[((size_t)b * Q + row) * 64 + o] = s1 + b1[o];
  } else {
    p1[((size_t)b * C + (row - Q)) * 64 + o] = s1;
    pv[((size_t)b * C + (row - Q)) * 64 + o] = sv + bv[o];
  }
}

extern "C" size_t gf_rel_cross_attention_workspace_bytes(int Q, int C, int B) {
  if (Q <= 0 || C <= 0 || B <= 0) return 0;
  return align256(sizeof(float) * 64 * (size_t)B * Q) + 2 * align256(sizeof(float) * 64 * (size_t)B * C) +
         align256(sizeof(uint32_t) * ((size_t)B * Q + 1)) + 1024;
}

namespace gf {
// row maxima of the gathered maps, gf_bias.cu
int bias_ctx_rowmax(const float *const *geo_ptrs, const int *geo_ld, const int *ctx_idx, int B, int Q, int C,
                    uint32_t *rowmax, uint32_t *gmax, cudaStream_t st);
}  // namespace gf

static int att_launch(AttArgs &a, const float *tgt2, const float *memory, const float *b1, const float *bv,
                      const float *const *geo_ptrs, const int *geo_ld, void *workspace, size_t workspace_bytes,
                      cudaStream_t st) {
  const bool fused = geo_ptrs != nullptr;
  Arena ar(workspace, workspace_bytes);
  float *aq = ar.take<float>(64 * (size_t)a.B * a.Q);
  float *p1 = ar.take<float>(64 * (size_t)a.B * a.C);
  float *pv = ar.take<float>(64 * (size_t)a.B * a.C);
  uint32_t *rm = ar.take<uint32_t>((size_t)a.B * a.Q + 1);
  if (!ar.ok) {
    set_error("rel_cross_attention: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              gf_rel_cross_attention_workspace_bytes(a.Q, a.C, a.B));
    return GF_ERR_WORKSPACE;
  }
  att_tables_kernel<<<dim3(a.Q + a.C, a.B), 64, 0, st>>>(tgt2, memory, a.Q, a.C, a.B, a.w1, b1, a.wv, bv, aq, p1, pv);
  GF_LAUNCHED();
  a.aq = aq, a.p1 = p1, a.pv = pv;
  if (fused) {
    for (int b = 0; b < a.B; ++b) a.geo[b] = geo_ptrs[b], a.geo_ld[b] = geo_ld[b];
    int rc = bias_ctx_rowmax(geo_ptrs, geo_ld, a.ctx_idx, a.B, a.Q, a.C, rm, rm + (size_t)a.B * a.Q, st);
    if (rc) return rc;
    a.rowmax = rm, a.gmax = rm + (size_t)a.B * a.Q;
  }
  static int done[64] = {0};
  int dev = 0;
  GF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!__atomic_load_n(&done[dev], __ATOMIC_ACQUIRE)) {
    GF_CUDA(cudaFuncSetAttribute(rel_cross_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    GF_CUDA(cudaFuncSetAttribute(rel_cross_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    __atomic_store_n(&done[dev], 1, __ATOMIC_RELEASE);
  }
  if (fused)
    rel_cross_attention_kernel<true><<<dim3(a.Q, a.B), ATT_THREADS, ATT_SMEM, st>>>(a);
  else
    rel_cross_attention_kernel<false><<<dim3(a.Q, a.B), ATT_THREADS, ATT_SMEM, st>>>(a);
  GF_LAUNCHED();
  return GF_OK;
}

#define ATT_COMMON_CHECKS(name)                                                                            \
  GF_CHECK_ARG(Q >= 1 && C >= 1 && B >= 1, name ": need Q, C, B >= 1");                                      \
  GF_CHECK_ARG(tgt2 && memory && w1 && b1 && w2 && b2 && wv && bv && wo && bo && out, name ": null pointer")

extern "C" int gf_rel_cross_attention(const float *tgt2, const float *memory, const float *relative_pos, int Q, int C,
                                      int B, const float *w1, const float *b1, const float *w2, const float *b2,
                                      const float *wv, const float *bv, const float *wo, const float *bo, float *out,
                                      void *workspace, size_t workspace_bytes, void *stream) {
  ATT_COMMON_CHECKS("rel_cross_attention");
  GF_CHECK_ARG(relative_pos, "rel_cross_attention: null relative_pos");
  AttArgs a = {};
  a.w1 = w1, a.w2 = w2, a.wv = wv, a.wo = wo, a.b2 = b2, a.bo = bo, a.Q = Q, a.C = C, a.B = B, a.rel = relative_pos, a.out = out;
  return att_launch(a, tgt2, memory, b1, bv, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gf_rel_cross_attention_fused(const float *tgt2, const float *memory, const float *const *geo_ptrs,
                                            const int *geo_ld, const int *ctx_idx, const float *query_xyz,
                                            const float *ctx_xyz, const float *gauss_B, int gauss_ld, const float *pc_min,
                                            const float *pc_max, int Q, int C, int B, const float *w1, const float *b1,
                                            const float *w2, const float *b2, const float *wv, const float *bv,
                                            const float *wo, const float *bo, float *out, void *workspace,
                                            size_t workspace_bytes, void *stream) {
  ATT_COMMON_CHECKS("rel_cross_attention_fused");
  GF_CHECK_ARG(geo_ptrs && geo_ld && ctx_idx && query_xyz && ctx_xyz && gauss_B && pc_min && pc_max,
               "rel_cross_attention_fused: null pointer");
  GF_CHECK_ARG(gauss_ld >= 32, "rel_cross_attention_fused: the embedding has 32 frequencies (gauss_B must be (3, >= 32))");
  GF_CHECK_ARG(B <= ATT_MAX_B, "rel_cross_attention_fused: B=%d, at most %d batch elements per call", B, ATT_MAX_B);
  AttArgs a = {};
  a.w1 = w1, a.w2 = w2, a.wv = wv, a.wo = wo, a.b2 = b2, a.bo = bo, a.Q = Q, a.C = C, a.B = B, a.out = out;
  a.ctx_idx = ctx_idx, a.query_xyz = query_xyz, a.ctx_xyz = ctx_xyz;
  a.gauss_B = gauss_B, a.ldb = gauss_ld, a.pc_min = pc_min, a.pc_max = pc_max;
  return att_launch(a, tgt2, memory, b1, bv, geo_ptrs, geo_ld, workspace, workspace_bytes, (cudaStream_t)stream);
}
