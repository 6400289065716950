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
constexpr int ATT_TILE = 64;     // contexts per tile = MMA N
constexpr int ATT_THREADS = 256;
constexpr int ATT_MAX_B = 8;     // batch elements of the fused variant (map pointers travel in the kernel parameters)
constexpr int ATT_W_HALF = 128 * 128;           // one K-half (32 floats = 128 bytes per row) of a 128-row weight tile
constexpr int ATT_W_BYTES = 2 * ATT_W_HALF;     // [Wv; W1] and [W2; 0]: 128 x 64 fp32
constexpr int ATT_T_HALF = ATT_TILE * 128;      // one K-half of a 64-row operand tile (embedding / hidden)
constexpr int ATT_T_BYTES = 2 * ATT_T_HALF;
constexpr int ATT_SMEM = 2 * ATT_W_BYTES + 4 * ATT_T_BYTES + 1024;  // A1, A2, 2 embedding tiles, 2 hidden tiles
constexpr uint32_t ATT_D1_COL = 0, ATT_D2_COL = 3 * ATT_TILE;       // TMEM: three D1 buffers, two D2 buffers

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

// K-major tile with 128-byte swizzle, `rows` rows: row r, fp32 column k (0..63) -> byte offset inside the tile
template <int ROWS>
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  const int half = k >> 5, kk = k & 31;
  return (uint32_t)(half * (ROWS * 128) + r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2)));
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = 64
constexpr uint32_t ATT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((ATT_TILE >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(ATT_IDESC), "r"(accumulate)
      : "memory");
}

// D (128 x 64, TMEM) = A (128 x 64) * B^T (64 x 64): eight K = 8 steps, 32 bytes apart inside a swizzled row, the
// second half of K in the second half-tile; then the completion of everything issued so far arrives on `mbar`
__device__ __forceinline__ void umma_tile(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, uint32_t mbar) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint32_t ka = (uint32_t)((kk >> 2) * ATT_W_HALF + (kk & 3) * 32), kb = (uint32_t)((kk >> 2) * ATT_T_HALF + (kk & 3) * 32);
    umma_tf32(d_tmem, umma_desc(a_saddr + ka), umma_desc(b_saddr + kb), kk > 0);
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
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

// Software pipeline over the tiles of one (query, batch element), one __syncthreads per tile.  In iteration i
//   every thread    loads tile i+2 of the embedding into the operand buffer tile i has just released
//   warps 2,3,6,7   (TMEM lanes 64..127) turn product 1 of tile i into the hidden tile  (two warps per lane
//                   quadrant, 32 contexts each)
//   warps 0,1,4,5   (TMEM lanes 0..63) run softmax + weighted sum of tile i-1 from product 2 and the v2 rows of product 1
//   thread 0        then issues product 2 of tile i and product 1 of tile i+2; both complete under the next iteration.
// Product 1 of a tile is issued two iterations before it is consumed, product 2 one iteration: three D1 and two D2
// accumulator buffers in TMEM, one mbarrier per buffer.
template <bool FUSED>
__global__ void __launch_bounds__(ATT_THREADS, 1) rel_cross_attention_kernel(const AttArgs a) {
  extern __shared__ unsigned char att_smem_raw[];
  __shared__ __align__(8) unsigned long long s_mbar[5];  // [0..2] product 1 into D1[b], [3..4] product 2 into D2[b]
  __shared__ uint32_t s_tmem;
  __shared__ float s_part[3][2][ATT_D];  // partial softmax states (max, sum, weighted sum) of the two column halves
  __shared__ float s_out[ATT_D];
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const int q = blockIdx.x, b = blockIdx.y;
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA1 = smem, *sA2 = smem + ATT_W_BYTES, *sR = smem + 2 * ATT_W_BYTES, *sH = sR + 2 * ATT_T_BYTES;
  uint32_t mbar[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) mbar[i] = smem_u32(&s_mbar[i]);

  // ---- one-time set-up: weights into swizzled K-major tiles, TMEM, barriers -----------------------------------
  // A1 rows 0..63 = Wv, rows 64..127 = W1;  A2 rows 0..63 = W2, rows 64..127 = 0
  for (int e = tid; e < 128 * ATT_D; e += ATT_THREADS) {
    const int r = e >> 6, k = e & 63;
    *reinterpret_cast<float *>(sA1 + sw128_off<128>(r, k)) = r < 64 ? __ldg(a.wv + r * 64 + k) : __ldg(a.w1 + (r - 64) * 64 + k);
    *reinterpret_cast<float *>(sA2 + sw128_off<128>(r, k)) = r < 64 ? __ldg(a.w2 + r * 64 + k) : 0.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar[i]) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // roles and per-thread constants
  const bool hidden_role = (warp & 2u) != 0;     // warps 2,3,6,7: TMEM lanes 64..127
  const int ch = (int)(((warp & 1u) << 5) | lane);  // channel inside the role's 64 lanes
  const int colhalf = (int)(warp >> 2) * 32;      // which 32 of the tile's 64 contexts this warp handles
  const float aq = hidden_role ? __ldg(a.aq + ((size_t)b * a.Q + q) * 64 + ch) : 0.f;
  const float b2 = hidden_role ? 0.f : __ldg(a.b2 + ch);
  const float *p1 = a.p1 + (size_t)b * a.C * 64, *pv = a.pv + (size_t)b * a.C * 64;
  // hidden-tile store addresses: half-tile and word of column ch, plus the eight swizzled chunk offsets
  const uint32_t h_base = (uint32_t)((ch >> 5) * ATT_T_HALF + colhalf * 128 + ((ch & 3) << 2));
  uint32_t h_sw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h_sw[j] = (uint32_t)((((ch & 31) >> 2) ^ j) << 4);
  float m_run = -INFINITY, s_run = 0.f, acc = 0.f;  // online softmax over this warp's contexts, per channel

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

  // The embedding tile in two steps, so that its global loads are in flight while the thread does its role's work:
  // fetch(tile) issues them (a quarter of row t / 4: 16 floats, or for the fused variant the gathered map entry and the
  // context's coordinates), stash(buf) writes the operand buffer (fused: after the sin / cos of its eight frequencies).
  float4 pre[4];
  float n0 = 0.f, n1 = 0.f, n2 = 0.f;
  bool pre_live = false;
  const int lr = (int)(tid >> 2), qd = (int)(tid & 3u);
  // shared-space addresses with the swizzle folded into per-thread constants: no address arithmetic per store
  const uint32_t sR_s = smem_u32(sR), sH_s = smem_u32(sH);
  uint32_t st_off[4];  // this thread's four 16-byte chunks of row lr (unfused: columns 16 qd + 4 j; fused: see stash)
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_off[j] = FUSED ? sw128_off<ATT_TILE>(lr, (j & 1) * 32 + 8 * qd + 4 * (j >> 1)) : sw128_off<ATT_TILE>(lr, 16 * qd + 4 * j);
  auto sts128 = [](uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  };
  auto fetch = [&](int tile) {
    const int c = tile * ATT_TILE + lr;
    pre_live = c < a.C;
    if (!FUSED) {
      const float4 *src = reinterpret_cast<const float4 *>(a.rel + (((size_t)q * a.C + (pre_live ? c : 0)) * a.B + b) * 64) + 4 * qd;
#pragma unroll
      for (int j = 0; j < 4; ++j) pre[j] = pre_live ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      n0 = n1 = n2 = 0.f;
      if (pre_live) {
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
    }
  };
  auto stash = [&](int buf) {
    const uint32_t dst = sR_s + (uint32_t)buf * ATT_T_BYTES;
    if (!FUSED) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sts128(dst + st_off[j], pre[j]);
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 2; ++j4) {  // this thread's eight frequencies: sin -> columns j, cos -> columns 32 + j
        float sn[4], cs[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 8 * qd + 4 * j4 + e;
          const float p = fmaf(n2, __ldg(a.gauss_B + 2 * a.ldb + j), fmaf(n1, __ldg(a.gauss_B + a.ldb + j), __fmul_rn(n0, __ldg(a.gauss_B + j))));
          sincosf(p, &sn[e], &cs[e]);
          if (!pre_live) sn[e] = cs[e] = 0.f;
        }
        sts128(dst + st_off[2 * j4], make_float4(sn[0], sn[1], sn[2], sn[3]));
        sts128(dst + st_off[2 * j4 + 1], make_float4(cs[0], cs[1], cs[2], cs[3]));
      }
    }
  };
  auto load_tile = [&](int tile, int buf) {
    fetch(tile);
    stash(buf);
  };

  const int T = (a.C + ATT_TILE - 1) / ATT_TILE;
  load_tile(0, 0);
  if (T > 1) load_tile(1, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // tiles were written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t t_lane = tmem + (((warp & 3u) * 32u) << 16);  // this warp's 32 TMEM lanes
  if (tid == 0) {
    umma_tile(tmem + ATT_D1_COL, smem_u32(sA1), smem_u32(sR), mbar[0]);
    if (T > 1) umma_tile(tmem + ATT_D1_COL + ATT_TILE, smem_u32(sA1), smem_u32(sR + ATT_T_BYTES), mbar[1]);
  }

  for (int i = 0; i <= T; ++i) {
    // global loads first: tile i + 2 of the embedding and this thread's 32 table entries of the tile it works on
    const bool more = i + 2 < T;
    if (more) fetch(i + 2);
    float tab[32];
    {
      const int t = hidden_role ? i : i - 1;
      const float *src = (hidden_role ? p1 : pv) + ch;
      const int c0 = t * ATT_TILE + colhalf;
      if (t >= 0 && t < T) {
#pragma unroll
        for (int e = 0; e < 32; ++e) tab[e] = c0 + e < a.C ? __ldg(src + (size_t)(c0 + e) * 64) : 0.f;
      }
    }
    if (i < T) {  // product 1 of tile i has landed in D1[i % 3]; the operand buffer i % 2 is free again
      mbar_wait(mbar[i % 3], (uint32_t)((i / 3) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (hidden_role) {
      if (i < T) {  // h = relu(W1 rel + (W1 tgt2[q] + b1) - W1 memory[c]) -> hidden tile i % 2, row = context, column = ch
        float d[32];
        tmem_ld32(t_lane + ATT_D1_COL + (uint32_t)((i % 3) * ATT_TILE + colhalf), d);
        // row colhalf + e, column ch: (colhalf + e) & 7 == e & 7, so the swizzled chunk is one of eight per-thread
        // constants and the row is an immediate offset
        const uint32_t dst = sH_s + (uint32_t)(i & 1) * ATT_T_BYTES + h_base;
        const int c0 = i * ATT_TILE + colhalf;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float h = c0 + e < a.C ? fmaxf(d[e] + aq - tab[e], 0.f) : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + h_sw[e & 7] + (uint32_t)e * 128u), "f"(h) : "memory");
        }
      }
    } else if (i >= 1) {  // softmax over the contexts and weighted sum of tile i - 1
      const int t = i - 1;
      mbar_wait(mbar[3 + (t & 1)], (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float sv[32], vv[32];
      tmem_ld32(t_lane + ATT_D2_COL + (uint32_t)((t & 1) * ATT_TILE + colhalf), sv);
      tmem_ld32(t_lane + ATT_D1_COL + (uint32_t)((t % 3) * ATT_TILE + colhalf), vv);
      const int c0 = t * ATT_TILE + colhalf;
      float mx = m_run;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        sv[e] = c0 + e < a.C ? (sv[e] + b2) * 0.125f : -INFINITY;  // / sqrt(64), :449
        mx = fmaxf(mx, sv[e]);
      }
      if (mx > -INFINITY) {
        const float scale = __expf(m_run - mx);  // exp(-inf) = 0 for the first contexts
        s_run *= scale, acc *= scale;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (c0 + e < a.C) {
            const float w = __expf(sv[e] - mx);
            s_run += w;
            acc = fmaf(w, vv[e] + tab[e], acc);
          }
        }
        m_run = mx;
      }
    }
    if (more) stash(i & 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (i < T)  // product 2: lanes 0..63 <- W2 h
        umma_tile(tmem + ATT_D2_COL + (uint32_t)((i & 1) * ATT_TILE), smem_u32(sA2), smem_u32(sH + (i & 1) * ATT_T_BYTES), mbar[3 + (i & 1)]);
      if (i + 2 < T)  // product 1 of tile i + 2: lanes 0..63 <- Wv rel, lanes 64..127 <- W1 rel
        umma_tile(tmem + ATT_D1_COL + (uint32_t)(((i + 2) % 3) * ATT_TILE), smem_u32(sA1), smem_u32(sR + (i & 1) * ATT_T_BYTES), mbar[(i + 2) % 3]);
    }
  }
  // ---- merge the two column halves, then out_mlp: relu(Wo (acc / s) + bo) ----------------------------------------
  if (!hidden_role) {
    const int hh = (int)(warp >> 2);
    s_part[0][hh][ch] = m_run, s_part[1][hh][ch] = s_run, s_part[2][hh][ch] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 64) {
    const float m0 = s_part[0][0][tid], m1 = s_part[0][1][tid];
    const float mx = fmaxf(m0, m1);
    float s = 0.f, av = 0.f;
    if (mx > -INFINITY) {
      const float e0 = __expf(m0 - mx), e1 = __expf(m1 - mx);
      s = s_part[1][0][tid] * e0 + s_part[1][1][tid] * e1;
      av = s_part[2][0][tid] * e0 + s_part[2][1][tid] * e1;
    }
    s_out[tid] = s > 0.f ? av / s : 0.f;
  }
  __syncthreads();
  if (tid < 64) {
    float o = __ldg(a.bo + tid);
#pragma unroll 8
    for (int f = 0; f < 64; ++f) o = fmaf(__ldg(a.wo + tid * 64 + f), s_out[f], o);
    a.out[((size_t)q * a.B + b) * 64 + tid] = fmaxf(o, 0.f);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
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
