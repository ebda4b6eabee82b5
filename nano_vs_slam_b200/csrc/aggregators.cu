// The two alternative global-descriptor aggregators of VPRHead (modules/decoders/vpr.py:70-76):
//   GeM over PixelUnshuffle(4)   modules/aggregators/gem.py:21-33   (letters GEM_N, GEM_S_A)
//   ConvAP                       modules/aggregators/convap.py:29-37 (letters CONVAP_S_A, V3 CONVAP_S_A)
// Both read the (B,C,H,W) encoder map once; outputs are B x 16C floats, so these are latency-sized kernels.
#include "common.cuh"

namespace nvs {

// out[b, c*16 + (y%4)*4 + x%4] = ( mean_{cells} max(x, eps)^p )^(1/p): PixelUnshuffle(4) moves the position inside
// every 4x4 cell into the channel index (gem.py:23-24), the pooling window is the whole unshuffled map (:30).
__global__ void __launch_bounds__(256) gem_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H,
                                                  int W, float p, float inv_p, float eps) {
  __shared__ float red[16][17];
  const int c = blockIdx.x, b = blockIdx.y;
  const int cls = threadIdx.x & 15, grp = threadIdx.x >> 4;  // 16 threads per class stride over its cells
  const int i = cls >> 2, j = cls & 3;
  const int ch = H >> 2, cw = W >> 2, cells = ch * cw;
  const float* xp = x + ((size_t)b * C + c) * H * W;
  float s = 0.f;
  for (int q = grp; q < cells; q += 16) {
    const int cy = q / cw, cx = q - cy * cw;
    s += powf(fmaxf(xp[(size_t)(4 * cy + i) * W + 4 * cx + j], eps), p);
  }
  red[cls][grp] = s;
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) t += red[threadIdx.x][g];
    out[((size_t)b * C + c) * 16 + threadIdx.x] = powf(t / (float)cells, inv_p);
  }
}

// AdaptiveAvgPool2d((S1,S2)) of the INPUT map: bin i covers rows [floor(i*H/S1), ceil((i+1)*H/S1)).  The 1x1
// channel_pool conv is linear, so conv -> pool (convap.py:30-33) equals pool -> conv; pooling first reads x once
// and leaves a C x (S1*S2) matrix for the channel mix.
__global__ void __launch_bounds__(128) aap_kernel(const float* __restrict__ x, float* __restrict__ pooled, int C, int H,
                                                  int W, int S1, int S2) {
  const int c = blockIdx.x, b = blockIdx.y;
  const float* xp = x + ((size_t)b * C + c) * H * W;
  const int nb = S1 * S2;
  for (int bin = threadIdx.x >> 3; bin < nb; bin += 16) {  // 8 threads per bin
    const int bi = bin / S2, bj = bin - bi * S2;
    const int y0 = (bi * H) / S1, y1 = ((bi + 1) * H + S1 - 1) / S1;
    const int x0 = (bj * W) / S2, x1 = ((bj + 1) * W + S2 - 1) / S2;
    const int rw = x1 - x0, n = (y1 - y0) * rw;
    float s = 0.f;
    for (int q = threadIdx.x & 7; q < n; q += 8) {
      const int yy = q / rw, xx = q - yy * rw;
      s += xp[(size_t)(y0 + yy) * W + x0 + xx];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((threadIdx.x & 7) == 0) pooled[((size_t)b * C + c) * nb + bin] = s / (float)n;
  }
}

// out[b, co*nb + bin] = (sum_ci w[co,ci] * pooled[b,ci,bin] + bias[co]) / max(||.||_2, 1e-12)   (convap.py:30-36)
__global__ void __launch_bounds__(256) convap_finish_kernel(const float* __restrict__ pooled, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            int Cin, int Cout, int nb) {
  extern __shared__ float sm[];
  float* ps = sm;             // [Cin][nb]
  float* os = sm + Cin * nb;  // [Cout][nb]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < Cin * nb; i += 256) ps[i] = pooled[(size_t)b * Cin * nb + i];
  __syncthreads();
  float ss = 0.f;
  for (int o = tid; o < Cout * nb; o += 256) {
    const int co = o / nb, bin = o - co * nb;
    float a = bias[co];
    for (int ci = 0; ci < Cin; ++ci) a = fmaf(w[(size_t)co * Cin + ci], ps[ci * nb + bin], a);
    os[o] = a;
    ss += a * a;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
  for (int o = tid; o < Cout * nb; o += 256) out[(size_t)b * Cout * nb + o] = os[o] * inv;
}

}  // namespace nvs

extern "C" int nvs_gem(const float* x, float* out, int32_t B, int32_t C, int32_t H, int32_t W, float p, float eps,
                       void* stream) {
  if (!x || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535) return NVS_ERR_ARG;
  if ((H % 4) != 0 || (W % 4) != 0) return NVS_ERR_ARG;  // PixelUnshuffle(4) raises in the reference as well
  if (!(p > 0.f)) return NVS_ERR_ARG;
  nvs::gem_kernel<<<dim3(C, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, C, H, W, p, 1.0f / p, eps);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" size_t nvs_convap_workspace_bytes(int32_t B, int32_t Cin, int32_t s1, int32_t s2) {
  if (B <= 0 || Cin <= 0 || s1 <= 0 || s2 <= 0) return 0;
  return sizeof(float) * (size_t)B * Cin * s1 * s2;
}

extern "C" int nvs_convap(const float* x, const float* weight, const float* bias, float* out, void* workspace,
                          size_t workspace_bytes, int32_t B, int32_t Cin, int32_t Cout, int32_t H, int32_t W,
                          int32_t s1, int32_t s2, void* stream) {
  if (!x || !weight || !bias || !out || !workspace) return NVS_ERR_ARG;
  if (B <= 0 || Cin <= 0 || Cout <= 0 || H < s1 || W < s2 || s1 <= 0 || s2 <= 0 || B > 65535) return NVS_ERR_ARG;
  if (workspace_bytes < nvs_convap_workspace_bytes(B, Cin, s1, s2)) return NVS_ERR_ARG;
  const size_t smem = sizeof(float) * (size_t)(Cin + Cout) * s1 * s2;
  if (smem > 48 * 1024) return NVS_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* pooled = static_cast<float*>(workspace);
  nvs::aap_kernel<<<dim3(Cin, B), 128, 0, st>>>(x, pooled, Cin, H, W, s1, s2);
  NVS_CHECK_LAUNCH();
  nvs::convap_finish_kernel<<<B, 256, smem, st>>>(pooled, weight, bias, out, Cin, Cout, s1 * s2);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

namespace nvs {
// torch's upsample_bilinear2d (align_corners=False): src = scale * (dst + 0.5) - 0.5 clamped at 0, i1 = i0 + 1 if
// inside, weights (1 - l, l), rows first then columns:  h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                                            int Hin, int Win, int Hout, int Wout, float sh, float sw) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= Hout * Wout) return;
  const int oy = idx / Wout, ox = idx - oy * Wout;
  const unsigned char* ib = img + (size_t)b * Hin * Win * 3 + c;
  float v;
  if (Hin == Hout && Win == Wout) {
    v = __fdiv_rn((float)ib[((size_t)oy * Win + ox) * 3], 255.f);
  } else {
    const float fy = fmaxf(__fsub_rn(__fmul_rn(sh, (float)oy + 0.5f), 0.5f), 0.f);
    const float fx = fmaxf(__fsub_rn(__fmul_rn(sw, (float)ox + 0.5f), 0.5f), 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float v00 = __fdiv_rn((float)ib[((size_t)y0 * Win + x0) * 3], 255.f);
    const float v01 = __fdiv_rn((float)ib[((size_t)y0 * Win + x1) * 3], 255.f);
    const float v10 = __fdiv_rn((float)ib[((size_t)y1 * Win + x0) * 3], 255.f);
    const float v11 = __fdiv_rn((float)ib[((size_t)y1 * Win + x1) * 3], 255.f);
    const float top = __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01));
    const float bot = __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11));
    v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
  }
  out[(((size_t)b * 3 + c) * Hout + oy) * Wout + ox] = __fmul_rn(__fsub_rn(v, 0.5f), 2.f);
}
}  // namespace nvs

extern "C" int nvs_preprocess_u8(const uint8_t* img, float* out, int32_t B, int32_t Hin, int32_t Win, int32_t Hout,
                                 int32_t Wout, void* stream) {
  if (!img || !out || B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || B > 65535) return NVS_ERR_ARG;
  const dim3 grid((Hout * Wout + 255) / 256, 3, B);
  nvs::preprocess_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, out, Hin, Win, Hout, Wout, (float)Hin / (float)Hout, (float)Win / (float)Wout);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
