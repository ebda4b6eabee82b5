// Fused NetVLAD (modules/aggregators/netvlad.py:79-106; the loop form :158-193 is the same math):
//   xh = x / max(||x||_C, 1e-12);  a = softmax_K(W xh);  V[k,c] = sum_s a[k,s] xh[c,s] - cent[k,c] sum_s a[k,s]
//   V <- V / max(||V[k,:]||, 1e-12)  (intra)  ->  flatten k-major  ->  / max(||.||, 1e-12)
// The reference materialises the (N,K,C,S) residual tensor (78.6 MB per 240x320 frame); here one CTA
// streams a slice of S through shared memory in 64-pixel chunks and keeps its share of the K x C
// accumulator in registers.  Kernel 1 writes per-slice partial sums, kernel 2 reduces and normalises.
#include "common.cuh"

namespace nvs {

constexpr int VP = 64;  // pixels per chunk

// Kernel 1, per (frame, slice of S): 64-pixel chunks stream through a double-buffered, pixel-major shared-memory tile
// (cp.async, the next chunk lands while this one is computed).  Two phases per chunk, both on packed FFMA2:
//   assignment   thread = (pixel pair p / p + 32, group of K/8 clusters): raw logits W x and |x|^2 in one pass over the
//                channels (the normalisation is folded in afterwards: W xh = (W x) / |x|), softmax over the 8 groups
//                through two small exchanges, a' = a / |x| stored duplicated (a', a') = the packed operand as loaded
//   aggregation  thread = 4 clusters x 4 channels: V[k][c] += a'[k][p] x[c][p] (= a xh), 8 FFMA2 per pixel from three
//                128-bit loads
// sum_p a[k][p] (the centroid term) is accumulated per thread in the assignment phase and reduced over the lanes once,
// at the end.  The previous version normalised the tile in place (two extra passes), computed the norms with 64 of the
// 256 threads and synchronised seven times per chunk: 1.04 ms per 256 frames, 27 % of the FMA rate.
template <int C, int K>
struct VladCfg {
  static_assert(C % 4 == 0 && K % 8 == 0 && (K / 8) % 2 == 0, "NetVLAD tiling");
  static constexpr int XP = C + 4;            // pitch of xs[p][c]: 128-bit reads of 4 channels are conflict-free
  static constexpr int AP = 2 * K + 4;        // pitch of as2[p][k] (float2 per entry)
  static constexpr int KPT = K / 8;           // clusters per thread in the assignment phase
  static constexpr int TILES = (K / 4) * (C / 4);             // 4 x 4 output tiles of the aggregation phase
  static constexpr int TPT = (TILES + 255) / 256;             // tiles per thread
  static constexpr size_t SMEM = sizeof(float) * (2 * VP * XP + C * K + VP * AP + 2 * 8 * VP);
};

template <int C, int K>
__global__ void __launch_bounds__(256, 2) netvlad_partial_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ w_assign,
                                                              float* __restrict__ partial, int S,
                                                              int px_per_split) {
  using Cfg = VladCfg<C, K>;
  constexpr int XP = Cfg::XP, AP = Cfg::AP, KPT = Cfg::KPT, TPT = Cfg::TPT;
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                      // [2][VP][XP]  raw descriptors of the chunk, pixel-major
  float* wt = xs + 2 * VP * XP;        // [C][K]       soft-assignment weights, transposed (clusters contiguous)
  float* as2 = wt + C * K;             // [VP][AP]     (a', a') pairs
  float* smax = as2 + VP * AP;         // [8][VP]
  float* ssum = smax + 8 * VP;         // [8][VP]

  const int tid = threadIdx.x, lane = tid & 31;
  const int split = blockIdx.x, b = blockIdx.y, nsplit = gridDim.x;
  const int s_begin = split * px_per_split;
  const int s_end = min(S, s_begin + px_per_split);
  const float* xb = x + (size_t)b * C * S;

  auto load_chunk = [&](int buf, int s0) {  // 4-byte async copies: coalesced rows of x in, transposed tile out
    float* dst = xs + buf * VP * XP;
    for (int i = tid; i < C * VP; i += 256) {
      const int c = i / VP, pp = i - c * VP;
      const bool ok = s0 + pp < s_end;
      cp_async4(dst + pp * XP + c, ok ? xb + (size_t)c * S + s0 + pp : xb, ok);
    }
    cp_async_commit();
  };
  if (s_begin < s_end) load_chunk(0, s_begin);
  for (int i = tid; i < K * C; i += 256) {
    const int k = i / C, c = i - k * C;
    wt[c * K + k] = w_assign[i];
  }

  // aggregation-phase ownership: tile t -> clusters [4 kt, 4 kt + 4), channels [4 ct, 4 ct + 4)
  float2 acc[TPT][4][2];
#pragma unroll
  for (int t = 0; t < TPT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[t][i][0] = acc[t][i][1] = make_float2(0.f, 0.f);
  // assignment-phase ownership: pixels pa, pa + 32; clusters [kg KPT, kg KPT + KPT); kg = warp: weights are broadcasts
  const int pa = lane, kg = tid >> 5;
  float asum[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) asum[i] = 0.f;

  int buf = 0;
  for (int s0 = s_begin; s0 < s_end; s0 += VP, buf ^= 1) {
    const int np = min(VP, s_end - s0);
    cp_async_wait<0>();
    __syncthreads();  // this chunk has landed; the previous chunk (other buffer, as2) is fully consumed
    if (s0 + VP < s_end) load_chunk(buf ^ 1, s0 + VP);
    const float* xc = xs + buf * VP * XP;
    // ---- soft assignment ----
    float2 lg[2][KPT / 2];
#pragma unroll
    for (int i = 0; i < KPT / 2; ++i) lg[0][i] = lg[1][i] = make_float2(0.f, 0.f);
    float ss0 = 0.f, ss1 = 0.f;
#pragma unroll 2
    for (int c = 0; c < C; c += 4) {
      const float4 x0 = *reinterpret_cast<const float4*>(xc + pa * XP + c);
      const float4 x1 = *reinterpret_cast<const float4*>(xc + (pa + 32) * XP + c);
      const float v0[4] = {x0.x, x0.y, x0.z, x0.w}, v1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        ss0 = fmaf(v0[cc], v0[cc], ss0);
        ss1 = fmaf(v1[cc], v1[cc], ss1);
        const float2 d0 = make_float2(v0[cc], v0[cc]), d1 = make_float2(v1[cc], v1[cc]);
        const float* wr = wt + (c + cc) * K + kg * KPT;
#pragma unroll
        for (int i = 0; i < KPT / 2; ++i) {
          const float2 w2 = *reinterpret_cast<const float2*>(wr + 2 * i);
          lg[0][i] = __ffma2_rn(d0, w2, lg[0][i]);
          lg[1][i] = __ffma2_rn(d1, w2, lg[1][i]);
        }
      }
    }
    const float sinv0 = 1.f / fmaxf(sqrtf(ss0), 1e-12f), sinv1 = 1.f / fmaxf(sqrtf(ss1), 1e-12f);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < KPT / 2; ++i) {
      lg[0][i].x *= sinv0; lg[0][i].y *= sinv0;
      lg[1][i].x *= sinv1; lg[1][i].y *= sinv1;
      m0 = fmaxf(m0, fmaxf(lg[0][i].x, lg[0][i].y));
      m1 = fmaxf(m1, fmaxf(lg[1][i].x, lg[1][i].y));
    }
    smax[kg * VP + pa] = m0;
    smax[kg * VP + pa + 32] = m1;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      m0 = fmaxf(m0, smax[g * VP + pa]);
      m1 = fmaxf(m1, smax[g * VP + pa + 32]);
    }
    float e0 = 0.f, e1 = 0.f;
#pragma unroll
    for (int i = 0; i < KPT / 2; ++i) {
      lg[0][i].x = expf(lg[0][i].x - m0); lg[0][i].y = expf(lg[0][i].y - m0);
      lg[1][i].x = expf(lg[1][i].x - m1); lg[1][i].y = expf(lg[1][i].y - m1);
      e0 += lg[0][i].x + lg[0][i].y;
      e1 += lg[1][i].x + lg[1][i].y;
    }
    ssum[kg * VP + pa] = e0;
    ssum[kg * VP + pa + 32] = e1;
    __syncthreads();
    e0 = e1 = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      e0 += ssum[g * VP + pa];
      e1 += ssum[g * VP + pa + 32];
    }
    const float inv0 = pa < np ? 1.f / e0 : 0.f, inv1 = pa + 32 < np ? 1.f / e1 : 0.f;  // padded pixels contribute nothing
    {
      float* a0 = as2 + pa * AP + 2 * kg * KPT;
      float* a1 = as2 + (pa + 32) * AP + 2 * kg * KPT;
#pragma unroll
      for (int i = 0; i < KPT / 2; ++i) {
        const float ax = lg[0][i].x * inv0, ay = lg[0][i].y * inv0, bx = lg[1][i].x * inv1, by = lg[1][i].y * inv1;
        asum[2 * i] += ax + bx;
        asum[2 * i + 1] += ay + by;
        *reinterpret_cast<float4*>(a0 + 4 * i) = make_float4(ax * sinv0, ax * sinv0, ay * sinv0, ay * sinv0);
        *reinterpret_cast<float4*>(a1 + 4 * i) = make_float4(bx * sinv1, bx * sinv1, by * sinv1, by * sinv1);
      }
    }
    __syncthreads();
    // ---- aggregation: V[k][c] += a'[k][p] * x[c][p] ----
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int tile = tid + 256 * t;
      if (tile < Cfg::TILES) {
        const int kt = tile / (C / 4), ct = tile - kt * (C / 4);
        const float* ap = as2 + 8 * kt;
        const float* xp = xc + 4 * ct;
#pragma unroll 4
        for (int pp = 0; pp < VP; ++pp) {
          const float4 a01 = *reinterpret_cast<const float4*>(ap + pp * AP);
          const float4 a23 = *reinterpret_cast<const float4*>(ap + pp * AP + 4);
          const float4 xv = *reinterpret_cast<const float4*>(xp + pp * XP);
          const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
          const float2 a0 = make_float2(a01.x, a01.y), a1 = make_float2(a01.z, a01.w);
          const float2 a2 = make_float2(a23.x, a23.y), a3 = make_float2(a23.z, a23.w);
          acc[t][0][0] = __ffma2_rn(a0, x01, acc[t][0][0]); acc[t][0][1] = __ffma2_rn(a0, x23, acc[t][0][1]);
          acc[t][1][0] = __ffma2_rn(a1, x01, acc[t][1][0]); acc[t][1][1] = __ffma2_rn(a1, x23, acc[t][1][1]);
          acc[t][2][0] = __ffma2_rn(a2, x01, acc[t][2][0]); acc[t][2][1] = __ffma2_rn(a2, x23, acc[t][2][1]);
          acc[t][3][0] = __ffma2_rn(a3, x01, acc[t][3][0]); acc[t][3][1] = __ffma2_rn(a3, x23, acc[t][3][1]);
        }
      }
    }
  }
  float* out = partial + ((size_t)b * nsplit + split) * (K * C + K);
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    const int tile = tid + 256 * t;
    if (tile < Cfg::TILES) {
      const int kt = tile / (C / 4), ct = tile - kt * (C / 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(out + (4 * kt + i) * C + 4 * ct) =
            make_float4(acc[t][i][0].x, acc[t][i][0].y, acc[t][i][1].x, acc[t][i][1].y);
    }
  }
#pragma unroll
  for (int i = 0; i < KPT; ++i) {
    const float v = warp_sum(asum[i]);
    if (lane == 0) out[K * C + kg * KPT + i] = v;
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
  int want = (4 * 296 + B - 1) / B;  // ~4 waves of the 2 x 148 resident CTAs: the last wave's idle tail stays small
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
