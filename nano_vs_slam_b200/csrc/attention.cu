// Streaming-softmax attention core of EfficientSelfAttention (modules/segformer.py:113-133).
// The reference materialises sim = q k^T ((B*heads) x Nq x Nk fp32: 92 MB per 240x320 frame, 4.3 GB per
// 512x1024 frame); here one thread owns one query, K/V of the (frame, head) stream through shared
// memory in blocks of KB keys and the softmax is kept online (running max / sum), fp32 throughout.
// head_dim is 16 (S) or 12 (N): far too small for tensor-core tiles to pay, the kernel is FFMA/MUFU bound.
#include "common.cuh"

namespace nvs {

constexpr int ATT_THREADS = 128;  // queries per CTA
constexpr int ATT_KB = 256;       // keys staged per block
constexpr int ATT_G = 8;          // keys per online-softmax group

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const float* __restrict__ q,
                                                                const float* __restrict__ kv,
                                                                float* __restrict__ out, int C, int Nq,
                                                                int Nk, float scale_log2e) {
  constexpr int DP = 20;  // row pitch: 16-byte aligned rows, 4-way (not 16-way) conflicts on the fill
  static_assert(D <= 16 && D % 4 == 0, "head_dim 12 or 16");
  __shared__ __align__(16) float ks[ATT_KB][DP];
  __shared__ __align__(16) float vs[ATT_KB][DP];
  const int head = blockIdx.y, b = blockIdx.z;
  const int n = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool active = n < Nq;
  const float* qb = q + ((size_t)b * C + head * D) * Nq;
  const float* kb = kv + ((size_t)b * 2 * C + head * D) * Nk;
  const float* vb = kv + ((size_t)b * 2 * C + C + head * D) * Nk;

  float qr[D], o[D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    qr[c] = active ? qb[(size_t)c * Nq + n] * scale_log2e : 0.f;
    o[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int j0 = 0; j0 < Nk; j0 += ATT_KB) {
    const int nk = min(ATT_KB, Nk - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < D * ATT_KB; i += ATT_THREADS) {
      const int c = i / ATT_KB, j = i - c * ATT_KB;  // coalesced along keys
      const bool ok = j < nk;
      ks[j][c] = ok ? kb[(size_t)c * Nk + j0 + j] : 0.f;
      vs[j][c] = ok ? vb[(size_t)c * Nk + j0 + j] : 0.f;
    }
    __syncthreads();
    for (int g0 = 0; g0 < nk; g0 += ATT_G) {
      float s[ATT_G];
      float gm = -INFINITY;
#pragma unroll
      for (int g = 0; g < ATT_G; ++g) {
        float a = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 kk = *reinterpret_cast<const float4*>(&ks[g0 + g][c4]);
          a = fmaf(qr[c4], kk.x, a);
          a = fmaf(qr[c4 + 1], kk.y, a);
          a = fmaf(qr[c4 + 2], kk.z, a);
          a = fmaf(qr[c4 + 3], kk.w, a);
        }
        s[g] = (g0 + g < nk) ? a : -INFINITY;
        gm = fmaxf(gm, s[g]);
      }
      const float mn = fmaxf(m, gm);
      const float corr = exp2f(m - mn);  // m = -inf on the first group -> 0
      l *= corr;
#pragma unroll
      for (int c = 0; c < D; ++c) o[c] *= corr;
      m = mn;
#pragma unroll
      for (int g = 0; g < ATT_G; ++g) {
        const float pw = exp2f(s[g] - mn);
        l += pw;
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(&vs[g0 + g][c4]);
          o[c4] = fmaf(pw, vv.x, o[c4]);
          o[c4 + 1] = fmaf(pw, vv.y, o[c4 + 1]);
          o[c4 + 2] = fmaf(pw, vv.z, o[c4 + 2]);
          o[c4 + 3] = fmaf(pw, vv.w, o[c4 + 3]);
        }
      }
    }
  }
  if (active) {
    const float inv = 1.f / l;
    float* ob = out + ((size_t)b * C + head * D) * Nq;
#pragma unroll
    for (int c = 0; c < D; ++c) ob[(size_t)c * Nq + n] = o[c] * inv;
  }
}

}  // namespace nvs

extern "C" int nvs_attention(const float* q, const float* kv, float* out, int32_t B, int32_t C, int32_t heads,
                             int32_t Nq, int32_t Nk, void* stream) {
  using namespace nvs;
  if (!q || !kv || !out || B <= 0 || C <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return NVS_ERR_ARG;
  if (C % heads != 0 || B > 65535) return NVS_ERR_ARG;
  const int d = C / heads;
  const float scale_log2e = (float)(1.0 / sqrt((double)d) * 1.4426950408889634);
  dim3 grid((Nq + ATT_THREADS - 1) / ATT_THREADS, heads, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d) {
    case 16: attention_kernel<16><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e); break;
    case 12: attention_kernel<12><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e); break;
    default: return NVS_ERR_UNSUPPORTED;
  }
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
