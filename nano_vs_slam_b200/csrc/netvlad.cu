// Fused NetVLAD (modules/aggregators/netvlad.py:79-106; the loop form :158-193 is the same math):
//   xh = x / max(||x||_C, 1e-12);  a = softmax_K(W xh);  V[k,c] = sum_s a[k,s] xh[c,s] - cent[k,c] sum_s a[k,s]
//   V <- V / max(||V[k,:]||, 1e-12)  (intra)  ->  flatten k-major  ->  / max(||.||, 1e-12)
// The reference materialises the (N,K,C,S) residual tensor (78.6 MB per 240x320 frame); here one CTA
// streams a slice of S through shared memory in 64-pixel chunks and keeps its share of the K x C
// accumulator in registers.  Kernel 1 writes per-slice partial sums, kernel 2 reduces and normalises.
#include "common.cuh"

namespace nvs {

constexpr int VP = 64;  // pixels per chunk

template <int C, int K>
struct VladCfg {
  static constexpr int NCT = C / 4;           // channel tiles (4 channels each)
  static constexpr int KG = 16 * NCT <= 256 ? 16 : 256 / NCT;  // cluster groups in the aggregation phase
  static constexpr int TK = K / KG;           // clusters per thread in the aggregation phase
  static constexpr int NT = 256;
  static constexpr int AGG_THREADS = KG * NCT;  // <= 256
  static_assert(K % KG == 0 && AGG_THREADS <= 256 && C % 4 == 0 && K % 4 == 0, "NetVLAD tiling");
  static constexpr int KPT = K / 4;           // clusters per thread in the assignment phase
  static constexpr int XP = VP + 1;           // pitch of xs[c][p]
  static constexpr int AP = K + 1;            // pitch of as[p][k]
  static constexpr size_t SMEM = sizeof(float) * (C * XP + K * C + VP * AP + VP + 4 * VP + 4 * VP);
};

template <int C, int K>
__global__ void __launch_bounds__(256) netvlad_partial_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ w_assign,
                                                              float* __restrict__ partial, int S,
                                                              int px_per_split) {
  using Cfg = VladCfg<C, K>;
  constexpr int XP = Cfg::XP, AP = Cfg::AP, KPT = Cfg::KPT, TK = Cfg::TK, NCT = Cfg::NCT;
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                  // [C][XP]   normalised descriptors of the chunk
  float* ws = xs + C * XP;         // [K][C]    soft-assignment weights
  float* as = ws + K * C;          // [VP][AP]  soft assignments
  float* sinv = as + VP * AP;      // [VP]
  float* smax = sinv + VP;         // [4][VP]
  float* ssum = smax + 4 * VP;     // [4][VP]

  const int tid = threadIdx.x;
  const int split = blockIdx.x, b = blockIdx.y, nsplit = gridDim.x;
  const int s_begin = split * px_per_split;
  const int s_end = min(S, s_begin + px_per_split);
  const float* xb = x + (size_t)b * C * S;

  for (int i = tid; i < K * C; i += 256) ws[i] = w_assign[i];

  // aggregation-phase ownership
  const int kt = tid / NCT, ct = tid % NCT;
  const bool agg = tid < Cfg::AGG_THREADS;
  float acc[TK][4];
  float asum[TK];
#pragma unroll
  for (int i = 0; i < TK; ++i) {
    asum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  // assignment-phase ownership
  const int p = tid & (VP - 1), kg = tid / VP;

  for (int s0 = s_begin; s0 < s_end; s0 += VP) {
    const int np = min(VP, s_end - s0);
    __syncthreads();  // previous chunk fully consumed (also covers the ws fill)
    for (int i = tid; i < C * VP; i += 256) {
      const int c = i / VP, pp = i - c * VP;
      xs[c * XP + pp] = pp < np ? xb[(size_t)c * S + s0 + pp] : 0.f;
    }
    __syncthreads();
    if (tid < VP) {
      float ss = 0.f;
      for (int c = 0; c < C; ++c) ss = fmaf(xs[c * XP + tid], xs[c * XP + tid], ss);
      sinv[tid] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
    __syncthreads();
    for (int i = tid; i < C * VP; i += 256) {
      const int c = i / VP, pp = i - c * VP;
      xs[c * XP + pp] *= sinv[pp];
    }
    __syncthreads();
    // ---- soft assignment: thread (p, kg) computes KPT logits ----
    float lg[KPT];
#pragma unroll
    for (int i = 0; i < KPT; ++i) lg[i] = 0.f;
    for (int c = 0; c < C; c += 4) {
      const float x0 = xs[(c + 0) * XP + p], x1 = xs[(c + 1) * XP + p];
      const float x2 = xs[(c + 2) * XP + p], x3 = xs[(c + 3) * XP + p];
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + (kg * KPT + i) * C + c);
        lg[i] = fmaf(x0, wv.x, lg[i]);
        lg[i] = fmaf(x1, wv.y, lg[i]);
        lg[i] = fmaf(x2, wv.z, lg[i]);
        lg[i] = fmaf(x3, wv.w, lg[i]);
      }
    }
    float m = lg[0];
#pragma unroll
    for (int i = 1; i < KPT; ++i) m = fmaxf(m, lg[i]);
    smax[kg * VP + p] = m;
    __syncthreads();
    m = fmaxf(fmaxf(smax[p], smax[VP + p]), fmaxf(smax[2 * VP + p], smax[3 * VP + p]));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      lg[i] = expf(lg[i] - m);
      sum += lg[i];
    }
    ssum[kg * VP + p] = sum;
    __syncthreads();
    sum = (ssum[p] + ssum[VP + p]) + (ssum[2 * VP + p] + ssum[3 * VP + p]);
    const float inv = (p < np) ? 1.f / sum : 0.f;  // padded pixels contribute nothing
#pragma unroll
    for (int i = 0; i < KPT; ++i) as[p * AP + kg * KPT + i] = lg[i] * inv;
    __syncthreads();
    // ---- aggregation: V[k][c] += a[k][p] * xh[c][p] ----
    if (agg) {
      for (int pp = 0; pp < VP; ++pp) {
        float av[TK], xv[4];
#pragma unroll
        for (int i = 0; i < TK; ++i) av[i] = as[pp * AP + kt * TK + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xs[(ct * 4 + j) * XP + pp];
#pragma unroll
        for (int i = 0; i < TK; ++i) {
          asum[i] += av[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
      }
    }
  }
  if (agg) {
    float* out = partial + ((size_t)b * nsplit + split) * (K * C + K);
#pragma unroll
    for (int i = 0; i < TK; ++i) {
      const int k = kt * TK + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) out[k * C + ct * 4 + j] = acc[i][j];
      if (ct == 0) out[K * C + k] = asum[i];
    }
  }
}

// one CTA per frame: reduce the slices, subtract centroid term, intra-normalise, global normalise
__global__ void __launch_bounds__(256) netvlad_finish_kernel(const float* __restrict__ partial,
                                                             const float* __restrict__ cent,
                                                             float* __restrict__ vlad, int C, int K,
                                                             int nsplit) {
  extern __shared__ float sv[];  // [K*C]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* pb = partial + (size_t)b * nsplit * (K * C + K);
  const int stride = K * C + K;
  float gsum = 0.f;
  for (int k = warp; k < K; k += 8) {
    float asum = 0.f;
    for (int s = 0; s < nsplit; ++s) asum += pb[(size_t)s * stride + K * C + k];
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      float v = 0.f;
      for (int s = 0; s < nsplit; ++s) v += pb[(size_t)s * stride + k * C + c];
      v -= cent[k * C + c] * asum;
      sv[k * C + c] = v;
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    float s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = sv[k * C + c] * inv;
      sv[k * C + c] = v;
      s2 = fmaf(v, v, s2);
    }
    gsum += s2;
  }
  gsum = warp_sum(gsum);
  if (lane == 0) red[warp] = gsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float ginv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
  for (int i = tid; i < K * C; i += 256) vlad[(size_t)b * K * C + i] = sv[i] * ginv;
}

static int pick_splits(int B, int S) {
  int want = (296 + B - 1) / B;
  int maxs = (S + VP - 1) / VP;
  if (want < 1) want = 1;
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  return want;
}

template <int C, int K>
static int run_netvlad(const float* x, const float* w, const float* cent, float* vlad, float* ws, int B,
                       int S, cudaStream_t st) {
  using Cfg = VladCfg<C, K>;
  const int splits = pick_splits(B, S);
  int per = (S + splits - 1) / splits;
  per = (per + VP - 1) / VP * VP;
  auto kern = netvlad_partial_kernel<C, K>;
  NVS_OPT_IN_SMEM(kern, Cfg::SMEM);
  kern<<<dim3(splits, B), 256, Cfg::SMEM, st>>>(x, w, ws, S, per);
  NVS_CHECK_LAUNCH();
  netvlad_finish_kernel<<<B, 256, sizeof(float) * K * C, st>>>(ws, cent, vlad, C, K, splits);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

}  // namespace nvs

extern "C" size_t nvs_netvlad_workspace_bytes(int32_t B, int32_t C, int32_t K, int32_t S) {
  if (B <= 0 || C <= 0 || K <= 0 || S <= 0) return 0;
  return sizeof(float) * (size_t)B * nvs::pick_splits(B, S) * ((size_t)K * C + K);
}

extern "C" int nvs_netvlad(const float* x, const float* w_assign, const float* centroids, float* vlad,
                           void* workspace, size_t workspace_bytes, int32_t B, int32_t C, int32_t K,
                           int32_t S, void* stream) {
  if (!x || !w_assign || !centroids || !vlad || !workspace) return NVS_ERR_ARG;
  if (B <= 0 || S <= 0 || B > 65535) return NVS_ERR_ARG;
  if (workspace_bytes < nvs_netvlad_workspace_bytes(B, C, K, S)) return NVS_ERR_ARG;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C == 64 && K == 64) return nvs::run_netvlad<64, 64>(x, w_assign, centroids, vlad, ws, B, S, st);
  if (C == 48 && K == 32) return nvs::run_netvlad<48, 32>(x, w_assign, centroids, vlad, ws, B, S, st);
  if (C == 48 && K == 64) return nvs::run_netvlad<48, 64>(x, w_assign, centroids, vlad, ws, B, S, st);
  if (C == 64 && K == 32) return nvs::run_netvlad<64, 32>(x, w_assign, centroids, vlad, ws, B, S, st);
  if (C == 128 && K == 64) return nvs::run_netvlad<128, 64>(x, w_assign, centroids, vlad, ws, B, S, st);  // letter F
  return NVS_ERR_UNSUPPORTED;
}
