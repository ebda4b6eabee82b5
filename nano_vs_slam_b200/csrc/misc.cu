// Small HBM/latency-bound kernels around the conv chain: depthwise 3x3, channel LayerNorm,
// channel softmax, channel L2 norm, seg argmax and the keypoint decode (post_processing).
#include "common.cuh"

namespace nvs {

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 + bias (modules/segformer.py:46-54).  One thread per output pixel, coalesced in x.
// ---------------------------------------------------------------------------------------------
__global__ void dwconv3x3_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                 const float* __restrict__ bias, float* __restrict__ dst, int C, int H,
                                 int W, size_t total) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {  // 32-bit index math: 64-bit div/mod is emulated
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const unsigned bc = i / (unsigned)(W * H);
    const int c = (int)(bc % (unsigned)C);
    const float* s = src + bc * (size_t)H * W;
    const float* k = w + c * 9;
    float acc = bias ? bias[c] : 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(__ldg(s + (size_t)yy * W + xx), k[(dy + 1) * 3 + dx + 1], acc);
      }
    }
    dst[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// per-pixel statistics over channels.  One thread per pixel; every channel read is coalesced across
// the warp (NCHW planes).  MODE 0: LayerNorm (x-mean)/(std+eps)*g+b  (segformer.py:70-73)
//                          MODE 1: softmax over channels (Softmax2d)
//                          MODE 2: x / max(||x||_2, 1e-12)  (F.normalize, vpr.py:85-86)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void channel_stat_kernel(const float* __restrict__ src, const float* __restrict__ g,
                                    const float* __restrict__ bb, float* __restrict__ dst, int C, int HW,
                                    size_t npix, float eps) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)npix; i += gridDim.x * blockDim.x) {
    const size_t b = i / HW;
    const int s = (int)(i - b * HW);
    const float* p = src + b * (size_t)C * HW + s;
    float* d = dst + b * (size_t)C * HW + s;
    if (MODE == 0) {
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += p[(size_t)c * HW];
      const float mean = sum / C;
      float var = 0.f;
      for (int c = 0; c < C; ++c) {
        const float t = p[(size_t)c * HW] - mean;
        var = fmaf(t, t, var);
      }
      const float inv = 1.f / (sqrtf(var / C) + eps);
      for (int c = 0; c < C; ++c) d[(size_t)c * HW] = (p[(size_t)c * HW] - mean) * inv * g[c] + bb[c];
    } else if (MODE == 1) {
      float m = -INFINITY;
      for (int c = 0; c < C; ++c) m = fmaxf(m, p[(size_t)c * HW]);
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(p[(size_t)c * HW] - m);
      const float inv = 1.f / sum;
      for (int c = 0; c < C; ++c) d[(size_t)c * HW] = expf(p[(size_t)c * HW] - m) * inv;
    } else {
      float ss = 0.f;
      for (int c = 0; c < C; ++c) {
        const float t = p[(size_t)c * HW];
        ss = fmaf(t, t, ss);
      }
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      for (int c = 0; c < C; ++c) d[(size_t)c * HW] = p[(size_t)c * HW] * inv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// argmax over classes -> int64 (first maximal index, like torch.argmax on distinct values).
// Optional nearest sampling at keypoint coordinates (grid_sample mode='nearest', align_corners=True).
// ---------------------------------------------------------------------------------------------
__global__ void seg_argmax_kernel(const float* __restrict__ seg, const float* __restrict__ coord,
                                  int64_t* __restrict__ out, int C, int Hs, int Ws, int Hc, int Wc, int H,
                                  int W, size_t total) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {  // 32-bit index math: 64-bit div/mod is emulated
    size_t b;
    int sy, sx;
    bool inb = true;
    if (coord) {
      const int ncell = Hc * Wc;
      b = i / ncell;
      const int cell = (int)(i - b * ncell);
      const float cx = coord[(b * 2 + 0) * ncell + cell];
      const float cy = coord[(b * 2 + 1) * ncell + cell];
      const float gx = cx / ((float)(W - 1) / 2.f) - 1.f;
      const float gy = cy / ((float)(H - 1) / 2.f) - 1.f;
      const float fx = (gx + 1.f) * ((float)(Ws - 1) / 2.f);
      const float fy = (gy + 1.f) * ((float)(Hs - 1) / 2.f);
      sx = (int)nearbyintf(fx);
      sy = (int)nearbyintf(fy);
      inb = sx >= 0 && sx < Ws && sy >= 0 && sy < Hs;
    } else {
      const int hw = Hs * Ws;
      b = i / hw;
      const int s = (int)(i - b * hw);
      sy = s / Ws;
      sx = s - sy * Ws;
    }
    int best = 0;
    if (inb) {
      const float* p = seg + b * (size_t)C * Hs * Ws + (size_t)sy * Ws + sx;
      float bv = p[0];
      for (int c = 1; c < C; ++c) {
        const float v = p[(size_t)c * Hs * Ws];
        if (v > bv) {
          bv = v;
          best = c;
        }
      }
    }
    out[i] = best;
  }
}

// ---------------------------------------------------------------------------------------------
// keypoint decode (kp2dtiny.py:593-647): one thread per cell.
//   score  <- score * border_mask                                    (:520-528)
//   coord  <- clamp(cell_xy*cell + step + shift*cross_ratio*step)    (:597-614)
//   feat   <- grid_sample(feat, coord_norm, bilinear, align_corners=True, zeros) / ||.||_2   (:627-631)
// Arithmetic follows ATen's grid_sampler (unnormalise ((g+1)/2)*(size-1); weights from floor corners)
// ---------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void decode_kernel(const float* __restrict__ score, const float* __restrict__ shift,
                              const float* __restrict__ feat, float* __restrict__ out_score,
                              float* __restrict__ out_coord, float* __restrict__ out_feat, int Hc, int Wc,
                              int D, int Hf, int Wf, int H, int W, float cell, float step, float cross,
                              size_t total) {
  const int ncell = Hc * Wc;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {  // 32-bit index math: 64-bit div/mod is emulated
    const size_t b = i / ncell;
    const int cidx = (int)(i - b * ncell);
    const int cy = cidx / Wc, cx = cidx - cy * Wc;
    const bool border = (cy == 0) | (cy == Hc - 1) | (cx == 0) | (cx == Wc - 1);
    out_score[i] = border ? 0.f : score[i];
    const float sx = shift[(b * 2 + 0) * ncell + cidx];
    const float sy = shift[(b * 2 + 1) * ncell + cidx];
    // base + shift * (cross_ratio * step): separate multiply and add, as torch evaluates it
    float px = __fadd_rn(__fadd_rn(__fmul_rn((float)cx, cell), step), __fmul_rn(sx, cross * step));
    float py = __fadd_rn(__fadd_rn(__fmul_rn((float)cy, cell), step), __fmul_rn(sy, cross * step));
    px = fminf(fmaxf(px, 0.f), (float)(W - 1));
    py = fminf(fmaxf(py, 0.f), (float)(H - 1));
    out_coord[(b * 2 + 0) * ncell + cidx] = px;
    out_coord[(b * 2 + 1) * ncell + cidx] = py;
    if (!feat) continue;
    // normalize_coord (:642-647) then ATen grid_sampler (align_corners): (g + 1) * ((size - 1) / 2)
    const float gx = __fsub_rn(__fdiv_rn(px, (float)(W - 1) / 2.f), 1.f);
    const float gy = __fsub_rn(__fdiv_rn(py, (float)(H - 1) / 2.f), 1.f);
    const float ix = __fmul_rn(__fadd_rn(gx, 1.f), (float)(Wf - 1) / 2.f);
    const float iy = __fmul_rn(__fadd_rn(gy, 1.f), (float)(Hf - 1) / 2.f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float ww = ix - fx0, we = 1.f - ww, wn_ = iy - fy0, ws = 1.f - wn_;
    const float wnw = ws * we, wne = ws * ww, wsw = wn_ * we, wse = wn_ * ww;
    const bool vx0 = x0 >= 0 && x0 < Wf, vx1 = x1 >= 0 && x1 < Wf;
    const bool vy0 = y0 >= 0 && y0 < Hf, vy1 = y1 >= 0 && y1 < Hf;
    const float* f = feat + b * (size_t)D * Hf * Wf;
    float v[DMAX];
    float ss = 0.f;
#pragma unroll 4
    for (int c = 0; c < DMAX; ++c) {
      if (c < D) {
        const float* fc = f + (size_t)c * Hf * Wf;
        const float tnw = (vy0 && vx0) ? __ldg(fc + (size_t)y0 * Wf + x0) : 0.f;
        const float tne = (vy0 && vx1) ? __ldg(fc + (size_t)y0 * Wf + x1) : 0.f;
        const float tsw = (vy1 && vx0) ? __ldg(fc + (size_t)y1 * Wf + x0) : 0.f;
        const float tse = (vy1 && vx1) ? __ldg(fc + (size_t)y1 * Wf + x1) : 0.f;
        const float a = tnw * wnw + tne * wne + tsw * wsw + tse * wse;
        v[c] = a;
        ss = fmaf(a, a, ss);
      }
    }
    const float nrm = sqrtf(ss);  // no epsilon in the reference (:629-630)
#pragma unroll 4
    for (int c = 0; c < DMAX; ++c)
      if (c < D) out_feat[(b * D + c) * (size_t)ncell + cidx] = v[c] / nrm;
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 conv with 1..4 output channels from a channels-last map (score / location heads, heads.py:33).
// 0.03 % of the model's FLOPs: one thread per output pixel, 16-byte loads of the pixel's channel vector
// for each tap (neighbouring threads re-hit L1), weights broadcast from shared memory.  Output NCHW.
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128) conv_small_kernel(const float* __restrict__ src,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ dst,
                                                         int H, int W, int cin, int act, size_t total) {
  // A warp handles 8 consecutive pixels (linear index over b,y,x).  16 lanes share one pixel: lane q = l%16
  // owns the 16-byte channel chunks q, q+16, ... of that pixel, so every load instruction of the warp reads
  // whole 128-byte lines (the first version gave each lane its own pixel = 32 lines per load and ran 10x
  // slower than its byte count).  Partial dot products are reduced over the 16 lanes by shuffles.
  extern __shared__ __align__(16) float ws[];  // [9][COUT][cin]
  for (int i = threadIdx.x; i < 9 * COUT * cin; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, q = lane & 15, half = lane >> 4;
  const int chunks = cin >> 2;  // float4 chunks per pixel
  const unsigned warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned utotal = (unsigned)total, uW = (unsigned)W, uH = (unsigned)H;
  for (unsigned p0 = warp_id * 8; p0 < utotal; p0 += n_warps * 8) {
    float acc[4][COUT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[i][o] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned pix = p0 + 2 * i + half;
      if (pix >= utotal) continue;
      const int x = (int)(pix % uW);
      const int y = (int)((pix / uW) % uH);
      const size_t b = pix / (uW * uH);
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const float4* s = reinterpret_cast<const float4*>(src + ((b * H + yy) * (size_t)W + xx) * cin);
        const float4* wt = reinterpret_cast<const float4*>(ws + tap * COUT * cin);
        for (int c4 = q; c4 < chunks; c4 += 16) {
          const float4 v = __ldg(s + c4);
#pragma unroll
          for (int o = 0; o < COUT; ++o) {
            const float4 k = wt[o * chunks + c4];
            acc[i][o] = fmaf(v.x, k.x, acc[i][o]);
            acc[i][o] = fmaf(v.y, k.y, acc[i][o]);
            acc[i][o] = fmaf(v.z, k.z, acc[i][o]);
            acc[i][o] = fmaf(v.w, k.w, acc[i][o]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        float v = acc[i][o];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        acc[i][o] = v;
      }
    if (q == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned pix = p0 + 2 * i + half;
        if (pix >= utotal) continue;
        const int x = (int)(pix % uW);
        const int y = (int)((pix / uW) % uH);
        const size_t b = pix / (uW * uH);
#pragma unroll
        for (int o = 0; o < COUT; ++o)
          dst[((b * COUT + o) * H + y) * (size_t)W + x] = apply_act(acc[i][o] + bias[o], act, o);
      }
    }
  }
}

#define NVS_REQUIRE_32BIT(total) do { if ((total) >= 0x7fffffffULL) return NVS_ERR_UNSUPPORTED; } while (0)

static inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 32;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace nvs

using namespace nvs;

extern "C" int nvs_dwconv3x3(const float* src, const float* w, const float* bias, float* dst, int32_t B,
                             int32_t C, int32_t H, int32_t W, void* stream) {
  if (!src || !w || !dst || B <= 0 || C <= 0 || H <= 0 || W <= 0) return NVS_ERR_ARG;
  const size_t total = (size_t)B * C * H * W;
  NVS_REQUIRE_32BIT(total);
  dwconv3x3_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, w, bias, dst, C, H, W, total);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_channel_layernorm(const float* src, const float* g, const float* b, float* dst, int32_t B,
                                     int32_t C, int32_t HW, float eps, void* stream) {
  if (!src || !g || !b || !dst || B <= 0 || C <= 0 || HW <= 0) return NVS_ERR_ARG;
  const size_t npix = (size_t)B * HW;
  channel_stat_kernel<0><<<grid_for(npix, 128), 128, 0, (cudaStream_t)stream>>>(src, g, b, dst, C, HW, npix, eps);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_softmax_channels(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, void* stream) {
  if (!src || !dst || B <= 0 || C <= 0 || HW <= 0) return NVS_ERR_ARG;
  const size_t npix = (size_t)B * HW;
  channel_stat_kernel<1><<<grid_for(npix, 128), 128, 0, (cudaStream_t)stream>>>(src, nullptr, nullptr, dst, C, HW, npix, 0.f);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_l2norm_channels(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, void* stream) {
  if (!src || !dst || B <= 0 || C <= 0 || HW <= 0) return NVS_ERR_ARG;
  const size_t npix = (size_t)B * HW;
  channel_stat_kernel<2><<<grid_for(npix, 128), 128, 0, (cudaStream_t)stream>>>(src, nullptr, nullptr, dst, C, HW, npix, 0.f);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_seg_argmax(const float* seg, const float* coord, int64_t* out, int32_t B, int32_t C,
                              int32_t Hs, int32_t Ws, int32_t Hc, int32_t Wc, int32_t H, int32_t W,
                              void* stream) {
  if (!seg || !out || B <= 0 || C <= 0 || Hs <= 0 || Ws <= 0) return NVS_ERR_ARG;
  if (coord && (Hc <= 0 || Wc <= 0 || H <= 1 || W <= 1)) return NVS_ERR_ARG;
  const size_t total = coord ? (size_t)B * Hc * Wc : (size_t)B * Hs * Ws;
  NVS_REQUIRE_32BIT(total);
  seg_argmax_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(seg, coord, out, C, Hs, Ws, Hc, Wc, H, W, total);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_decode(const float* score, const float* shift, const float* feat, float* out_score,
                          float* out_coord, float* out_feat, int32_t B, int32_t Hc, int32_t Wc, int32_t D,
                          int32_t Hf, int32_t Wf, int32_t H, int32_t W, int32_t cell, float cross_ratio,
                          void* stream) {
  if (!score || !shift || !out_score || !out_coord || B <= 0 || Hc <= 0 || Wc <= 0) return NVS_ERR_ARG;
  if (feat && (!out_feat || D <= 0 || Hf <= 0 || Wf <= 0 || H <= 1 || W <= 1)) return NVS_ERR_ARG;
  if (feat && D > 128) return NVS_ERR_UNSUPPORTED;
  const size_t total = (size_t)B * Hc * Wc;
  const float step = (cell - 1) / 2.0f;
  cudaStream_t st = (cudaStream_t)stream;
  if (!feat || D <= 32)
    decode_kernel<32><<<grid_for(total, 128), 128, 0, st>>>(score, shift, feat, out_score, out_coord, out_feat, Hc, Wc, D, Hf, Wf, H, W, (float)cell, step, cross_ratio, total);
  else
    decode_kernel<128><<<grid_for(total, 128), 128, 0, st>>>(score, shift, feat, out_score, out_coord, out_feat, Hc, Wc, D, Hf, Wf, H, W, (float)cell, step, cross_ratio, total);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_conv_small(const float* src, const float* weight, const float* bias, float* dst, int32_t B,
                              int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t act, void* stream) {
  if (!src || !weight || !bias || !dst || B <= 0 || H <= 0 || W <= 0) return NVS_ERR_ARG;
  if (cin <= 0 || (cin % 4) != 0 || cout < 1 || cout > 4) return NVS_ERR_UNSUPPORTED;
  const size_t total = (size_t)B * H * W;
  const size_t smem = sizeof(float) * 9 * cout * cin;
  if (smem > 48 * 1024) return NVS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((total + 7) / 8 * 32, 128);  // one warp per 8 pixels
  switch (cout) {
    case 1: conv_small_kernel<1><<<grid, 128, smem, st>>>(src, weight, bias, dst, H, W, cin, act, total); break;
    case 2: conv_small_kernel<2><<<grid, 128, smem, st>>>(src, weight, bias, dst, H, W, cin, act, total); break;
    case 3: conv_small_kernel<3><<<grid, 128, smem, st>>>(src, weight, bias, dst, H, W, cin, act, total); break;
    case 4: conv_small_kernel<4><<<grid, 128, smem, st>>>(src, weight, bias, dst, H, W, cin, act, total); break;
  }
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
