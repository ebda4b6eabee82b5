// Streaming-softmax attention core of EfficientSelfAttention (modules/segformer.py:113-133).
// The reference materialises sim = q k^T ((B*heads) x Nq x Nk fp32: 92 MB per 240x320 frame, 4.3 GB per
// 512x1024 frame); here K/V of a (frame, head) stream through shared memory in blocks of KB keys and the
// softmax is kept online (running max / sum), fp32 throughout.
//
// head_dim is 16 (S) or 12 (N): the P.V product would be an N = 16 GEMM, far too narrow for tcgen05 tiles
// (an MMA costs >= 50 cycles whatever N is), so this stays on the CUDA cores and is organised around their
// limits instead: every thread owns TWO queries (K/V rows are warp-broadcast LDS.128, so two queries halve the
// shared-memory instructions per FMA) and all multiply-adds are packed FFMA2 (fma.rn.f32x2, two fp32 FMAs per
// issue slot on sm_100) so the FMA pipe, not the issue port, is the bound; exp2 with pre-scaled logits.
// The large letters (D, D_A: C = 256, head_dim 64) use the same kernel with one query per thread and 64-key
// blocks (the query and output rows alone are 128 registers).
#include <stdlib.h>

#include "common.cuh"

namespace nvs {

constexpr int ATT_THREADS = 128;              // threads per CTA
constexpr int ATT_G = 8;                      // keys per online-softmax group

// ATT_QPT queries per thread, ATT_KB keys staged per block
template <int D, int ATT_QPT, int ATT_KB>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const float* __restrict__ q,
                                                                const float* __restrict__ kv,
                                                                float* __restrict__ out, int C, int Nq,
                                                                int Nk, float scale_log2e) {
  constexpr int ATT_QPB = ATT_THREADS * ATT_QPT;  // queries per CTA
  constexpr int DP = D + 4;  // row pitch: 16-byte aligned rows, 4-way (not 16-way) conflicts on the fill
  constexpr int H = D / 2;
  static_assert(D % 4 == 0 && 2 * ATT_KB * DP * 4 <= 48 * 1024, "head_dim multiple of 4; K/V block fits static smem");
  __shared__ __align__(16) float ks[ATT_KB][DP];
  __shared__ __align__(16) float vs[ATT_KB][DP];
  const int head = blockIdx.y, b = blockIdx.z;
  const float* qb = q + ((size_t)b * C + head * D) * Nq;
  const float* kb = kv + ((size_t)b * 2 * C + head * D) * Nk;
  const float* vb = kv + ((size_t)b * 2 * C + C + head * D) * Nk;

  int n[ATT_QPT];
  bool active[ATT_QPT];
  float2 qr[ATT_QPT][H], o[ATT_QPT][H];
  float m[ATT_QPT], l[ATT_QPT];
#pragma unroll
  for (int t = 0; t < ATT_QPT; ++t) {
    n[t] = blockIdx.x * ATT_QPB + t * ATT_THREADS + threadIdx.x;  // coalesced along queries
    active[t] = n[t] < Nq;
#pragma unroll
    for (int c = 0; c < H; ++c) {
      qr[t][c].x = active[t] ? qb[(size_t)(2 * c) * Nq + n[t]] * scale_log2e : 0.f;
      qr[t][c].y = active[t] ? qb[(size_t)(2 * c + 1) * Nq + n[t]] * scale_log2e : 0.f;
      o[t][c] = make_float2(0.f, 0.f);
    }
    m[t] = -INFINITY;
    l[t] = 0.f;
  }

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
      float s[ATT_QPT][ATT_G];
      float gm[ATT_QPT];
#pragma unroll
      for (int t = 0; t < ATT_QPT; ++t) gm[t] = -INFINITY;
#pragma unroll
      for (int g = 0; g < ATT_G; ++g) {
        float2 kk[H];
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 k4 = *reinterpret_cast<const float4*>(&ks[g0 + g][c4]);
          kk[c4 / 2] = make_float2(k4.x, k4.y);
          kk[c4 / 2 + 1] = make_float2(k4.z, k4.w);
        }
        const bool kvalid = g0 + g < nk;
#pragma unroll
        for (int t = 0; t < ATT_QPT; ++t) {
          float2 a = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < H; ++c) a = __ffma2_rn(qr[t][c], kk[c], a);
          s[t][g] = kvalid ? a.x + a.y : -INFINITY;
          gm[t] = fmaxf(gm[t], s[t][g]);
        }
      }
#pragma unroll
      for (int t = 0; t < ATT_QPT; ++t) {
        const float mn = fmaxf(m[t], gm[t]);
        const float corr = exp2f(m[t] - mn);  // m = -inf on the first group -> 0
        l[t] *= corr;
        const float2 c2 = make_float2(corr, corr);
#pragma unroll
        for (int c = 0; c < H; ++c) o[t][c] = __fmul2_rn(o[t][c], c2);
        m[t] = mn;
      }
#pragma unroll
      for (int g = 0; g < ATT_G; ++g) {
        float2 vv[H];
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(&vs[g0 + g][c4]);
          vv[c4 / 2] = make_float2(v4.x, v4.y);
          vv[c4 / 2 + 1] = make_float2(v4.z, v4.w);
        }
#pragma unroll
        for (int t = 0; t < ATT_QPT; ++t) {
          const float pw = exp2f(s[t][g] - m[t]);
          l[t] += pw;
          const float2 p2 = make_float2(pw, pw);
#pragma unroll
          for (int c = 0; c < H; ++c) o[t][c] = __ffma2_rn(p2, vv[c], o[t][c]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < ATT_QPT; ++t) {
    if (active[t]) {
      const float inv = 1.f / l[t];
      float* ob = out + ((size_t)b * C + head * D) * Nq;
#pragma unroll
      for (int c = 0; c < H; ++c) {
        ob[(size_t)(2 * c) * Nq + n[t]] = o[t][c].x * inv;
        ob[(size_t)(2 * c + 1) * Nq + n[t]] = o[t][c].y * inv;
      }
    }
  }
}

// attention_tc.cu: Q.K^T on the tensor cores (head_dim 12 / 16)
int attention_tc_launch(const float* q, const float* kv, float* out, int B, int C, int heads, int Nq, int Nk,
                        float scale_log2e, cudaStream_t st);

}  // namespace nvs

extern "C" int nvs_attention(const float* q, const float* kv, float* out, int32_t B, int32_t C, int32_t heads,
                             int32_t Nq, int32_t Nk, void* stream) {
  using namespace nvs;
  if (!q || !kv || !out || B <= 0 || C <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return NVS_ERR_ARG;
  if (C % heads != 0 || B > 65535) return NVS_ERR_ARG;
  const int d = C / heads;
  const float scale_log2e = (float)(1.0 / sqrt((double)d) * 1.4426950408889634);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // head_dim 12 / 16: logits on the tensor cores (attention_tc.cu); NVS_ATT_BACKEND=ffma keeps the all-FFMA kernels
  // below for A/B measurements (they remain the path of head_dim 64)
  static int tc_env = -1;
  if (tc_env < 0) {
    const char* e = getenv("NVS_ATT_BACKEND");
    tc_env = (e && e[0] == 'f') ? 0 : 1;
  }
  if (tc_env && (d == 12 || d == 16)) return attention_tc_launch(q, kv, out, B, C, heads, Nq, Nk, scale_log2e, st);
  // queries per thread for head_dim 12 / 16: 4 halves the K/V shared-memory loads per FMA again (+7 % at 32 k
  // tokens, config 4) but leaves too few CTAs for the small maps (4800 tokens: -2 %)
  static int qpt_env = -1;
  if (qpt_env < 0) {
    const char* e = getenv("NVS_ATT_QPT");
    qpt_env = e ? atoi(e) : 0;
    if (qpt_env != 2 && qpt_env != 4) qpt_env = 0;
  }
  const int qpt_small = qpt_env ? qpt_env : (Nq >= 16384 ? 4 : 2);
  const int qpb = ATT_THREADS * (d <= 16 ? qpt_small : 1);
  dim3 grid((Nq + qpb - 1) / qpb, heads, B);
  if (d <= 16 && qpt_small == 4) {
    if (d == 16) attention_kernel<16, 4, 256><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e);
    else if (d == 12) attention_kernel<12, 4, 256><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e);
    else return NVS_ERR_UNSUPPORTED;
    NVS_CHECK_LAUNCH();
    return NVS_OK;
  }
  switch (d) {
    case 16: attention_kernel<16, 2, 256><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e); break;
    case 12: attention_kernel<12, 2, 256><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e); break;
    case 64: attention_kernel<64, 1, 64><<<grid, ATT_THREADS, 0, st>>>(q, kv, out, C, Nq, Nk, scale_log2e); break;
    default: return NVS_ERR_UNSUPPORTED;
  }
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
