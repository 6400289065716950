// set_aggregator.mlp fused with the grouping that feeds it, for sm_100a.
//
// Reference: lib/pointnet2/pointnet2_modules.py:200-249 (PointnetSAModuleVotesSeparate.group_points + .mlp) with
// QueryAndGroup.forward (pointnet2_utils.py:303-356) and SharedMLP (pytorch_utils.py:9-32): for every centre j and
// every one of its nsample ball-query neighbours s
//     x   = [ (xyz[idx[j,s]] - new_xyz[j]) / radius  |  features[:, idx[j,s]] ]          (3 + C values, :333-341)
//     h_l = relu(bn_l(W_l h_{l-1}))     l = 1..L   (1x1 Conv2d without bias + BatchNorm2d + ReLU, pytorch_utils.py:59-104)
//     out[:, j] = max_s / mean_s h_L                                                       (pointnet2_modules.py:233-236)
// The reference materialises (B, 3+C, npoint, nsample) grouped tensors and an activation of that size per layer
// (19 -> 32 -> 32 -> 32 channels x 2048 x 64 in the model).  Here one warp owns one centre: lane = output channel,
// the layer inputs are broadcast lane to lane with shuffles, every lane keeps its rows of all weight matrices in
// registers, and the pooled value is the only thing written.  BatchNorm is applied in its inference form
// (running statistics folded into a per-channel scale and shift by the host); widths up to 32 per layer.
#include "gf_common.cuh"

namespace gf {

constexpr int AGG_MAX_LAYERS = 4;
constexpr int AGG_W = 32;  // widest layer (one lane per channel)

struct AggArgs {
  const float *xyz;       // (B, N, 3)
  const float *new_xyz;   // (B, m, 3)
  const float *features;  // (B, C, N) or null
  const int *idx;         // (B, m, ns)
  int B, N, m, ns, C;
  float radius;           // divides the relative coordinates when `normalize` (pointnet2_utils.py:335)
  int normalize;
  int use_xyz, n_layers, pool_avg;
  int width[AGG_MAX_LAYERS + 1];        // channels of the input and of every layer
  const float *w[AGG_MAX_LAYERS];       // (width[l+1], width[l]) row-major
  const float *scale[AGG_MAX_LAYERS];   // (width[l+1]) folded batch-norm scale
  const float *shift[AGG_MAX_LAYERS];   // (width[l+1]) folded batch-norm shift (or the conv bias when there is no bn)
  float *out;                           // (B, width[L], m)
};

template <int L>
__global__ void __launch_bounds__(256) group_mlp_pool_kernel(const AggArgs a) {
  const unsigned lane = threadIdx.x & 31u;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  // this lane's row of every weight matrix, its scale and shift
  float w[L][AGG_W], sc[L], sh[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const bool live = (int)lane < a.width[l + 1];
#pragma unroll
    for (int i = 0; i < AGG_W; ++i) w[l][i] = live && i < a.width[l] ? __ldg(a.w[l] + lane * a.width[l] + i) : 0.f;
    sc[l] = live ? __ldg(a.scale[l] + lane) : 0.f;
    sh[l] = live ? __ldg(a.shift[l] + lane) : 0.f;
  }
  const int c_in = a.width[0], c_out = a.width[L];
  const int xoff = a.use_xyz ? 3 : 0;
  for (int cj = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cj < a.B * a.m; cj += warps_per_grid) {
    const int b = cj / a.m, j = cj - b * a.m;
    const float *xyz = a.xyz + (size_t)b * a.N * 3;
    const int *idx = a.idx + ((size_t)b * a.m + j) * a.ns;
    const float centre = lane < 3u ? __ldg(a.new_xyz + ((size_t)b * a.m + j) * 3 + lane) : 0.f;
    float pooled = a.pool_avg ? 0.f : -INFINITY;
    for (int s = 0; s < a.ns; ++s) {
      const int p = __ldg(idx + s);
      // lane i holds input channel i: relative, optionally normalised coordinates, then the point's features
      float x = 0.f;
      if (a.use_xyz && lane < 3u) {
        x = __fsub_rn(__ldg(xyz + (size_t)p * 3 + lane), centre);  // :333
        if (a.normalize) x = __fdiv_rn(x, a.radius);  // :335
      } else if ((int)lane >= xoff && (int)lane < c_in) {
        x = __ldg(a.features + ((size_t)b * a.C + (lane - xoff)) * a.N + p);
      }
#pragma unroll
      for (int l = 0; l < L; ++l) {
        float accv = 0.f;
#pragma unroll
        for (int i = 0; i < AGG_W; ++i) accv = fmaf(w[l][i], __shfl_sync(0xffffffffu, x, i), accv);
        x = fmaxf(fmaf(accv, sc[l], sh[l]), 0.f);  // batch norm (inference form) + ReLU
      }
      pooled = a.pool_avg ? pooled + x : fmaxf(pooled, x);
    }
    if ((int)lane < c_out) a.out[((size_t)b * c_out + lane) * a.m + j] = a.pool_avg ? pooled / (float)a.ns : pooled;
  }
}

}  // namespace gf

using namespace gf;

extern "C" int gf_group_mlp_pool(const float *xyz, const float *new_xyz, const float *features, const int *idx, int B,
                                 int N, int m, int nsample, int C, float radius, int normalize_xyz, int use_xyz,
                                 int n_layers, const int *widths, const float *const *weights,
                                 const float *const *scales, const float *const *shifts, int pool_avg, float *out,
                                 void *stream) {
  GF_CHECK_ARG(B >= 0 && N >= 1 && m >= 0 && nsample >= 1, "group_mlp_pool: bad sizes");
  GF_CHECK_ARG(n_layers >= 1 && n_layers <= AGG_MAX_LAYERS, "group_mlp_pool: %d layers, 1..%d supported", n_layers, AGG_MAX_LAYERS);
  GF_CHECK_ARG(xyz && new_xyz && idx && widths && weights && scales && shifts && out, "group_mlp_pool: null pointer");
  GF_CHECK_ARG(features || C == 0, "group_mlp_pool: null features");
  GF_CHECK_ARG(use_xyz || C > 0, "group_mlp_pool: neither coordinates nor features");
  GF_CHECK_ARG(widths[0] == (use_xyz ? 3 : 0) + C, "group_mlp_pool: first width %d != %d input channels", widths[0],
               (use_xyz ? 3 : 0) + C);
  for (int l = 0; l <= n_layers; ++l)
    GF_CHECK_ARG(widths[l] >= 1 && widths[l] <= AGG_W, "group_mlp_pool: width %d of layer %d outside [1,%d]", widths[l], l, AGG_W);
  if (B == 0 || m == 0) return GF_OK;
  AggArgs a = {};
  a.xyz = xyz, a.new_xyz = new_xyz, a.features = features, a.idx = idx, a.B = B, a.N = N, a.m = m, a.ns = nsample, a.C = C;
  a.radius = radius, a.normalize = normalize_xyz ? 1 : 0;
  a.use_xyz = use_xyz, a.n_layers = n_layers, a.pool_avg = pool_avg;
  for (int l = 0; l <= n_layers; ++l) a.width[l] = widths[l];
  for (int l = 0; l < n_layers; ++l) {
    GF_CHECK_ARG(weights[l] && scales[l] && shifts[l], "group_mlp_pool: null weights of layer %d", l);
    a.w[l] = weights[l], a.scale[l] = scales[l], a.shift[l] = shifts[l];
  }
  a.out = out;
  const long long warps = (long long)B * m;
  long long blocks = (warps + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_layers) {
    case 1: group_mlp_pool_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    case 2: group_mlp_pool_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    case 3: group_mlp_pool_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    default: group_mlp_pool_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(a); break;
  }
  GF_LAUNCHED();
  return GF_OK;
}
