// Direct fp32 convolution (3x3 pad 1, or 1x1) with folded BatchNorm, activation and fused
// pool / pixel-shuffle / concat-read epilogues.  sm_100a, FFMA pipe.
//
// Why FFMA and not tcgen05: parity is 1e-4 relative on fp32 features and >= 99.9 % segmentation argmax
// agreement through ~20 chained layers; single-pass TF32/BF16 tensor-core products miss that
// (SURVEY.md §7 "Precision vs tensor cores"), so the layers run exact fp32 multiply-adds.  The kernel is
// therefore math-pipe bound (K = 9*Cin is contracted per output), and everything below is organised to
// keep the FFMA pipe issuing: 64 accumulators per thread, 30 shared-memory loads per 576 FFMAs,
// conflict-free LDS.64 activation reads, warp-broadcast LDS.128 weight reads, cp.async double buffering.
//
// Tiling
//   thread : 4 rows x 2 cols of output pixels x 8 output channels (64 accumulators)
//   warp   : lanes 2x16 -> 8 rows x 32 cols   (WIDE)   or lanes 4x8 -> 16 rows x 16 cols (NARROW)
//   CTA    : WM pixel-warps stacked in y  x  WN channel-warps (8 channels each): CT = 8*WN channels
//   K loop : input channels in chunks of CK (4 or 8), double buffered in shared memory:
//            s_in[2][CK][TH+2][PITCH] (halo tile, zero filled = conv padding) and s_w[2][CK][taps][CT]
#include <cuda_fp16.h>

#include "common.cuh"

namespace nvs {

// split channels-last format of the 3xFP16 tensor-core convs (conv_rs.cu): (x, y) -> packed fp16 pair of the values
// rounded to fp16 and packed fp16 pair of the exact remainders
__device__ __forceinline__ void split_pair16(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct ConvP {
  const float* src0;
  const float* src1;
  const float* w;
  const float* bias;
  float* dst;
  float* dst2;
  int c0_total, c0_off, c0;
  int c1_total, c1_off, c1;
  int dst_c_total, dst_c_off;
  int dst2_c_total, dst2_c_off;
  int H, W, inH, inW;
  int cin_pad, cout, cout_pad;
  int act, out_mode, in_mode;
  int dst_nhwc, dst2_nhwc;  // store layout of dst / dst2: 0 = NCHW planes, 1 = NHWC fp32 (feeds the 3xTF32 convs),
                            // 2 (dst only) = split fp16 hi / lo channels-last (feeds the 3xFP16 convs, conv_rs.cu)
  int tiles_x;
};

__device__ __noinline__ float slow_act(float v, int act, int channel) { return apply_act(v, act, channel); }

template <int KS, int CK, int WN, int WM, bool NARROW>
struct ConvCfg {
  static constexpr int TAPS = KS * KS;
  static constexpr int HALO = (KS == 3) ? 1 : 0;
  static constexpr int LX = NARROW ? 8 : 16;
  static constexpr int LY = 32 / LX;
  static constexpr int TW = LX * 2;
  static constexpr int TH = LY * 4 * WM;
  static constexpr int IW = TW + 2 * HALO;
  static constexpr int IH = TH + 2 * HALO;
  // NARROW: a half-warp spans two lane-rows (4 tile rows apart); PITCH % 8 == 4 puts them on
  // disjoint bank pairs.  WIDE: a half-warp is one lane-row, any even pitch is conflict free.
  static constexpr int PITCH = NARROW ? 20 : (KS == 3 ? 36 : 32);
  static constexpr int CT = WN * 8;
  static constexpr int NT = 32 * WN * WM;
  static constexpr int IN_ELEMS = CK * IH * PITCH;
  static constexpr int W_ELEMS = CK * TAPS * CT;
  static constexpr size_t SMEM = sizeof(float) * 2 * (IN_ELEMS + W_ELEMS);
};

template <int KS, int CK, int WN, int WM, bool NARROW>
__global__ void __launch_bounds__(32 * WN * WM, 2) conv_kernel(const ConvP p) {
  using C = ConvCfg<KS, CK, WN, WM, NARROW>;
  constexpr int TAPS = C::TAPS, HALO = C::HALO, PITCH = C::PITCH, CT = C::CT, NT = C::NT;
  constexpr int WR = 4 + 2 * HALO;  // window rows per thread
  constexpr int WC = 2 + 2 * HALO;  // window cols per thread

  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                    // [2][CK][IH][PITCH]
  float* s_w = smem + 2 * C::IN_ELEMS;   // [2][CK][TAPS][CT]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wn = warp % WN, wm = warp / WN;
  const int lx = lane % C::LX, ly = lane / C::LX;
  const int py = wm * (C::LY * 4) + ly * 4;  // thread's first output row inside the tile
  const int px = lx * 2;

  const int tile = blockIdx.x;
  const int x0 = (tile % p.tiles_x) * C::TW;
  const int y0 = (tile / p.tiles_x) * C::TH;
  const int ct0 = blockIdx.y * CT;  // first output channel of this CTA
  const int b = blockIdx.z;

  const int cin = p.c0 + p.c1;
  const int nchunks = p.cin_pad / CK;
  const size_t in_plane = (size_t)p.inH * p.inW;

  auto load_chunk = [&](int chunk, int buf) {
    float* di = s_in + buf * C::IN_ELEMS;
    // ---- activations: CK x IH x IW elements, 4-byte cp.async with zero fill outside the image ----
    for (int idx = tid; idx < CK * C::IH * C::IW; idx += NT) {
      const int c = idx / (C::IH * C::IW);
      const int rem = idx - c * (C::IH * C::IW);
      const int yy = rem / C::IW, xx = rem - yy * C::IW;
      const int cg = chunk * CK + c;
      const float* src = p.src0;
      bool ok = cg < cin;
      size_t off = 0;
      if (p.in_mode == NVS_IN_S2D) {
        // virtual channel cg = ci*4 + i*2 + j reads in[ci][2y+i][2x+j]
        const int ci = cg >> 2, i = (cg >> 1) & 1, j = cg & 1;
        const int gy = 2 * (y0 + yy) + i, gx = 2 * (x0 + xx) + j;
        ok = ok && (y0 + yy) < p.H && (x0 + xx) < p.W;
        off = ((size_t)b * p.c0_total + p.c0_off + ci) * in_plane + (size_t)gy * p.inW + gx;
      } else {
        const int gy = y0 - HALO + yy, gx = x0 - HALO + xx;
        ok = ok && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        if (cg < p.c0) {
          off = ((size_t)b * p.c0_total + p.c0_off + cg) * in_plane + (size_t)gy * p.W + gx;
        } else {
          src = p.src1;
          off = ((size_t)b * p.c1_total + p.c1_off + (cg - p.c0)) * in_plane + (size_t)gy * p.W + gx;
        }
      }
      cp_async4(di + (c * C::IH + yy) * PITCH + xx, ok ? (src + off) : p.src0, ok);
    }
    // ---- weights: CK x TAPS rows of CT floats, 16-byte cp.async ----
    float* dw = s_w + buf * C::W_ELEMS;
    const float* gw = p.w + ((size_t)chunk * CK * TAPS) * p.cout_pad + ct0;
    for (int idx = tid; idx < CK * TAPS * (CT / 4); idx += NT) {
      const int row = idx / (CT / 4), q = idx - row * (CT / 4);
      cp_async16(dw + row * CT + q * 4, gw + (size_t)row * p.cout_pad + q * 4);
    }
  };

  float acc[4][2][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[r][c][o] = 0.f;

  load_chunk(0, 0);
  cp_async_commit();

  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) {
      load_chunk(ch + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const float* si = s_in + buf * C::IN_ELEMS + py * PITCH + px;
    const float* sw = s_w + buf * C::W_ELEMS + wn * 8;
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
      float a[WR][WC];
#pragma unroll
      for (int r = 0; r < WR; ++r) {
#pragma unroll
        for (int q = 0; q < WC / 2; ++q) {
          const float2 v = *reinterpret_cast<const float2*>(si + (c * C::IH + r) * PITCH + 2 * q);
          a[r][2 * q] = v.x;
          a[r][2 * q + 1] = v.y;
        }
      }
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
        const int ky = t / KS, kx = t % KS;
        const float4 w0 = *reinterpret_cast<const float4*>(sw + (c * TAPS + t) * CT);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + (c * TAPS + t) * CT + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const float v = a[r + ky][cc + kx];
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[r][cc][o] = fmaf(v, wv[o], acc[r][cc][o]);
          }
      }
    }
    __syncthreads();
  }

  // ------------------------------------ epilogue ------------------------------------
  const int co0 = ct0 + wn * 8;  // this thread's first output channel
  if (co0 >= p.cout) return;
  const int gy0 = y0 + py, gx0 = x0 + px;
  if (gy0 >= p.H || gx0 >= p.W) return;

  float bias[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) bias[o] = p.bias[co0 + o];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float v = acc[r][c][o] + bias[o];
        // LeakyReLU / ReLU / identity inline; transcendental activations (tiny head layers) out of line
        acc[r][c][o] = p.act == NVS_ACT_LRELU  ? (v > 0.f ? v : 0.01f * v)
                       : p.act == NVS_ACT_NONE ? v
                       : p.act == NVS_ACT_RELU ? fmaxf(v, 0.f)
                                               : slow_act(v, p.act, co0 + o);
      }

  const bool x1ok = gx0 + 1 < p.W;

  if ((p.out_mode == NVS_OUT_PLAIN || p.out_mode == NVS_OUT_BOTH) && p.dst_nhwc) {
    // NHWC: the thread's 8 channels of one pixel are contiguous (2 x 16-byte stores)
    const bool full8 = co0 + 8 <= p.cout && ((p.dst_c_total | p.dst_c_off) & 3) == 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (gy0 + r >= p.H) break;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c == 1 && !x1ok) break;
        float* d = p.dst + (((size_t)b * p.H + gy0 + r) * p.W + gx0 + c) * p.dst_c_total + p.dst_c_off + co0;
        if (p.dst_nhwc == 2) {  // split format: 8 fp16 a_hi at channel offset c, 8 fp16 a_lo at c_total + c (host checks % 8)
          uint8_t* px = reinterpret_cast<uint8_t*>(p.dst) +
                        (((size_t)b * p.H + gy0 + r) * p.W + gx0 + c) * (size_t)p.dst_c_total * 4;
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const float v0 = co0 + 2 * o < p.cout ? acc[r][c][2 * o] : 0.f, v1 = co0 + 2 * o + 1 < p.cout ? acc[r][c][2 * o + 1] : 0.f;
            split_pair16(v0, v1, hi[o], lo[o]);
          }
          *reinterpret_cast<uint4*>(px + (p.dst_c_off + co0) * 2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(px + (p.dst_c_total + p.dst_c_off + co0) * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        } else if (full8) {
          reinterpret_cast<float4*>(d)[0] = make_float4(acc[r][c][0], acc[r][c][1], acc[r][c][2], acc[r][c][3]);
          reinterpret_cast<float4*>(d)[1] = make_float4(acc[r][c][4], acc[r][c][5], acc[r][c][6], acc[r][c][7]);
        } else {
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (co0 + o < p.cout) d[o] = acc[r][c][o];
        }
      }
    }
  } else if (p.out_mode == NVS_OUT_PLAIN || p.out_mode == NVS_OUT_BOTH) {
    const bool vec = x1ok && ((p.W & 1) == 0);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (co0 + o >= p.cout) break;
      float* d = p.dst + (((size_t)b * p.dst_c_total + p.dst_c_off + co0 + o) * p.H + gy0) * p.W + gx0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (gy0 + r >= p.H) break;
        if (vec) {
          *reinterpret_cast<float2*>(d + (size_t)r * p.W) = make_float2(acc[r][0][o], acc[r][1][o]);
        } else {
          d[(size_t)r * p.W] = acc[r][0][o];
          if (x1ok) d[(size_t)r * p.W + 1] = acc[r][1][o];
        }
      }
    }
  }
  if (p.out_mode == NVS_OUT_POOL || p.out_mode == NVS_OUT_BOTH) {
    // MaxPool2d(2,2), floor: thread-local because the thread owns aligned 2x2 blocks.
    const int Hp = p.H >> 1, Wp = p.W >> 1;
    const int qx = gx0 >> 1;
    if (qx < Wp && p.dst2_nhwc) {
      const bool full8 = co0 + 8 <= p.cout && ((p.dst2_c_total | p.dst2_c_off) & 3) == 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int qy = (gy0 >> 1) + h;
        if (qy >= Hp) break;
        float m[8];
#pragma unroll
        for (int o = 0; o < 8; ++o)
          m[o] = fmaxf(fmaxf(acc[2 * h][0][o], acc[2 * h][1][o]), fmaxf(acc[2 * h + 1][0][o], acc[2 * h + 1][1][o]));
        float* d = p.dst2 + (((size_t)b * Hp + qy) * Wp + qx) * p.dst2_c_total + p.dst2_c_off + co0;
        if (full8) {
          reinterpret_cast<float4*>(d)[0] = make_float4(m[0], m[1], m[2], m[3]);
          reinterpret_cast<float4*>(d)[1] = make_float4(m[4], m[5], m[6], m[7]);
        } else {
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (co0 + o < p.cout) d[o] = m[o];
        }
      }
    } else if (qx < Wp) {
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        if (co0 + o >= p.cout) break;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int qy = (gy0 >> 1) + h;
          if (qy >= Hp) break;
          const float m = fmaxf(fmaxf(acc[2 * h][0][o], acc[2 * h][1][o]),
                                fmaxf(acc[2 * h + 1][0][o], acc[2 * h + 1][1][o]));
          p.dst2[(((size_t)b * p.dst2_c_total + p.dst2_c_off + co0 + o) * Hp + qy) * Wp + qx] = m;
        }
      }
    }
  }
  if (p.out_mode == NVS_OUT_SHUFFLE) {
    // PixelShuffle(2): channel co -> (co/4, i=(co%4)/2, j=co%2); thread's 8 channels = 2 shuffled channels.
    const int H2 = p.H * 2, W2 = p.W * 2;
    const bool vec = x1ok && ((p.W & 1) == 0);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (co0 + 4 * s >= p.cout) break;
      const int cs = (co0 >> 2) + s;
      float* d = p.dst + (((size_t)b * p.dst_c_total + p.dst_c_off + cs) * H2) * W2;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (gy0 + r >= p.H) break;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float* row = d + (size_t)(2 * (gy0 + r) + i) * W2 + 2 * gx0;
          const float v0 = acc[r][0][4 * s + 2 * i], v1 = acc[r][0][4 * s + 2 * i + 1];
          const float v2 = acc[r][1][4 * s + 2 * i], v3 = acc[r][1][4 * s + 2 * i + 1];
          if (vec) {
            *reinterpret_cast<float4*>(row) = make_float4(v0, v1, v2, v3);
          } else {
            row[0] = v0;
            row[1] = v1;
            if (x1ok) {
              row[2] = v2;
              row[3] = v3;
            }
          }
        }
      }
    }
  }
}

template <int KS, int CK, int WN, int WM, bool NARROW>
static int launch_conv(const ConvP& p0, int B, cudaStream_t st) {
  using C = ConvCfg<KS, CK, WN, WM, NARROW>;
  ConvP p = p0;
  p.tiles_x = (p.W + C::TW - 1) / C::TW;
  const int tiles_y = (p.H + C::TH - 1) / C::TH;
  dim3 grid(p.tiles_x * tiles_y, p.cout_pad / C::CT, B);
  auto kern = conv_kernel<KS, CK, WN, WM, NARROW>;
  NVS_OPT_IN_SMEM(kern, C::SMEM);
  kern<<<grid, C::NT, C::SMEM, st>>>(p);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

// padding waste of a tile width against the plane width
static inline bool prefer_narrow(int H, int W) {
  auto waste = [](int n, int t) { return (double)(((n + t - 1) / t) * t) / n; };
  const double wide = waste(W, 32) * waste(H, 8);
  const double narrow = waste(W, 16) * waste(H, 16);
  return narrow < wide - 1e-9;
}

template <int KS, int CK>
static int dispatch_ct(const ConvP& p, int B, int ct, cudaStream_t st) {
  const bool narrow = prefer_narrow(p.H, p.W);
#define NVS_CASE(WN_, WM_)                                                   \
  return narrow ? launch_conv<KS, CK, WN_, WM_, true>(p, B, st)              \
                : launch_conv<KS, CK, WN_, WM_, false>(p, B, st)
  switch (ct) {
    case 8: NVS_CASE(1, 8);
    case 16: NVS_CASE(2, 4);
    case 24: NVS_CASE(3, 2);
    case 32: NVS_CASE(4, 2);
    case 48: NVS_CASE(6, 1);
    case 64: NVS_CASE(8, 1);
  }
#undef NVS_CASE
  return NVS_ERR_UNSUPPORTED;
}


// ------------------------------------------------------------------ stem layer (3 -> 16, NCHW image -> NHWC)
// backbone.conv1a (encoders.py:105-107): K = 27 is too thin for the tiled kernel above (one 4-channel chunk,
// all overhead) and for a TMA row of the tensor-core kernel.  Dedicated direct kernel: a 64x8 pixel tile per
// 128-thread CTA, the 3x10x66 input window and the 27x16 weights in shared memory; every thread computes all 16
// output channels of FOUR consecutive pixels of a row: the 6 inputs of a (channel, ky) row segment are loaded
// once and reused by the 3 kx taps, and each broadcast weight LDS.128 feeds 8 packed FFMA2.
constexpr int STEM_TX = 64, STEM_TY = 8, STEM_CO = 16, STEM_PX = 4;
__global__ void __launch_bounds__(128) stem_conv_kernel(ConvP p) {
  // every input value is stored twice, (v, v): it is the packed FFMA2 operand as loaded, no register MOVs
  // (the buffer is reused afterwards to stage the results)
  __shared__ __align__(16) unsigned char smem_raw[STEM_TY * STEM_TX * STEM_CO * 4];
  static_assert(sizeof(smem_raw) >= 3 * (STEM_TY + 2) * (STEM_TX + 4) * sizeof(float2), "input window fits");
  float2 (*tile)[STEM_TY + 2][STEM_TX + 4] = reinterpret_cast<float2 (*)[STEM_TY + 2][STEM_TX + 4]>(smem_raw);  // pitch 68
  __shared__ __align__(16) float ws[27][STEM_CO];
  const int b = blockIdx.z, x0 = blockIdx.x * STEM_TX, y0 = blockIdx.y * STEM_TY;
  const float* src = p.src0 + ((size_t)b * p.c0_total + p.c0_off) * p.H * p.W;
  // NVS_IN_U8_HWC: the frame is the camera's uint8 HWC image; /255 (visual_odometry.py:283) and (x - 0.5) * 2
  // (frontend.py:79) happen here, in the same fp32 operations, so the fp32 NCHW copy of the input never exists
  const unsigned char* src8 = reinterpret_cast<const unsigned char*>(p.src0) + (size_t)b * p.H * p.W * 3;
  const bool u8 = p.in_mode == NVS_IN_U8_HWC;
  const bool unit = p.in_mode == NVS_IN_UNIT;  // fp32 frames in [0,1]: x.sub(0.5).mul(2.0) of frontend.py:79 on load
  // Input window: ALL of a thread's global loads are issued before the first one is consumed (two fully unrolled
  // loops).  As one load-convert-store loop every iteration exposed a full global-load latency: ncu's source counters put
  // 66 % of the kernel's stall samples in this prologue and 16 % in the FFMA2 loop.
  {
    constexpr int WIN = (STEM_TY + 2) * (STEM_TX + 2), TOT = 3 * WIN, NLD = (TOT + 127) / 128;
    float v[NLD];
    int so[NLD];  // float2 offset inside `tile`, -1 = nothing to store
    // branch-free: clamped (always valid) addresses, the padding decision is kept in so[]; the uniform input-mode branch
    // sits outside the loops, so the sixteen loads of a thread issue back to back
    size_t ga[NLD];
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
      const int i = threadIdx.x + j * 128;
      const int c0 = i / WIN, r = i - c0 * WIN;
      const int c = c0 < 3 ? c0 : 2;
      const int yy = r / (STEM_TX + 2), xx = r - yy * (STEM_TX + 2);
      const int gy = y0 + yy - 1, gx = x0 + xx - 1;
      const bool live = i < TOT;
      const bool inside = live && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      const int o = (c * (STEM_TY + 2) + yy) * (STEM_TX + 4) + xx;
      so[j] = !live ? -1 : (inside ? o : -2 - o);  // -2 - o: zero padding, stored as 0 without the input conversion
      const int cy = min(max(gy, 0), p.H - 1), cx = min(max(gx, 0), p.W - 1);
      ga[j] = u8 ? ((size_t)cy * p.W + cx) * 3 + c : ((size_t)c * p.H + cy) * p.W + cx;
    }
    if (u8) {
#pragma unroll
      for (int j = 0; j < NLD; ++j) v[j] = (float)__ldg(src8 + ga[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NLD; ++j) v[j] = __ldg(src + ga[j]);
    }
    float2* tl = &tile[0][0][0];
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
      if (so[j] == -1) continue;
      float w = v[j];
      int o = so[j];
      if (o >= 0) {
        if (u8) w = __fmul_rn(__fsub_rn(__fdiv_rn(w, 255.f), 0.5f), 2.f);
        if (unit) w = __fmul_rn(__fsub_rn(w, 0.5f), 2.f);
      } else {
        o = -2 - o;
        w = 0.f;
      }
      tl[o] = make_float2(w, w);
    }
  }
  for (int i = threadIdx.x; i < 27 * STEM_CO; i += 128) {
    const int k = i / STEM_CO, co = i - k * STEM_CO;  // k = ci * 9 + tap; packed weights are [cin_pad][9][cout_pad]
    ws[k][co] = p.w[(size_t)k * p.cout_pad + co];
  }
  __syncthreads();
  const int lx = (threadIdx.x & 15) * STEM_PX, ly = threadIdx.x >> 4;  // pixels (lx .. lx+3, ly)
  float2 acc[STEM_PX][STEM_CO / 2];
#pragma unroll
  for (int j = 0; j < STEM_CO / 2; ++j) {
    const float2 bj = make_float2(p.bias[2 * j], p.bias[2 * j + 1]);
#pragma unroll
    for (int i = 0; i < STEM_PX; ++i) acc[i][j] = bj;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      float2 in[STEM_PX + 2];
#pragma unroll
      for (int i = 0; i < (STEM_PX + 2) / 2; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(&tile[c][ly + ky][lx + 2 * i]);
        in[2 * i] = make_float2(t.x, t.y);
        in[2 * i + 1] = make_float2(t.z, t.w);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int k = c * 9 + ky * 3 + kx;
#pragma unroll
        for (int q = 0; q < STEM_CO / 4; ++q) {
          const float4 w4 = *reinterpret_cast<const float4*>(&ws[k][4 * q]);
          const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
#pragma unroll
          for (int i = 0; i < STEM_PX; ++i) {
            acc[i][2 * q] = __ffma2_rn(in[i + kx], wa, acc[i][2 * q]);
            acc[i][2 * q + 1] = __ffma2_rn(in[i + kx], wb, acc[i][2 * q + 1]);
          }
        }
      }
    }
  // stage the tile's results (XOR-swizzled 16-byte chunks) and write them out as fully coalesced 16-byte
  // chunks: a thread owns 256 contiguous bytes of NHWC output, which as direct stores is 32 partial sectors
  // per instruction
  float4 (*stage)[STEM_TX][STEM_CO / 4] = reinterpret_cast<float4 (*)[STEM_TX][STEM_CO / 4]>(smem_raw);
  __syncthreads();  // every thread is done reading the input window
  const float neg_slope = p.act == NVS_ACT_NONE ? 1.f : (p.act == NVS_ACT_LRELU ? 0.01f : 0.f);
#pragma unroll
  for (int i = 0; i < STEM_PX; ++i) {
    const int x = lx + i;
#pragma unroll
    for (int q = 0; q < STEM_CO / 4; ++q) {
      float o[4] = {acc[i][2 * q].x, acc[i][2 * q].y, acc[i][2 * q + 1].x, acc[i][2 * q + 1].y};
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f) + neg_slope * fminf(o[e], 0.f);  // none / LeakyReLU / ReLU
      if (p.dst_nhwc == 2) {
        // split format, 16 channels: bytes [0, 32) of the pixel = 16 fp16 a_hi, bytes [32, 64) = 16 fp16 a_lo
        uint32_t h0, l0, h1, l1;
        split_pair16(o[0], o[1], h0, l0);
        split_pair16(o[2], o[3], h1, l1);
        uint2* ch_hi = reinterpret_cast<uint2*>(&stage[ly][x][(q >> 1) ^ ((x >> 2) & 3)]);
        uint2* ch_lo = reinterpret_cast<uint2*>(&stage[ly][x][(2 + (q >> 1)) ^ ((x >> 2) & 3)]);
        ch_hi[q & 1] = make_uint2(h0, h1);
        ch_lo[q & 1] = make_uint2(l0, l1);
      } else {
        stage[ly][x][q ^ ((x >> 2) & 3)] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  __syncthreads();
  constexpr int CH = STEM_CO / 4;
  for (int idx = threadIdx.x; idx < STEM_TY * STEM_TX * CH; idx += 128) {
    const int row = idx / (STEM_TX * CH), rem = idx - row * (STEM_TX * CH);
    const int x = rem / CH, q = rem - x * CH;
    const int gy = y0 + row, gx = x0 + x;
    if (gy < p.H && gx < p.W) {
      float4* d = reinterpret_cast<float4*>(p.dst + (((size_t)b * p.H + gy) * p.W + gx) * p.dst_c_total + p.dst_c_off);
      d[q] = stage[row][x][q ^ ((x >> 2) & 3)];
    }
  }
}

}  // namespace nvs

extern "C" int nvs_conv_cout_tile(int32_t cout) {
  if (cout <= 0) return NVS_ERR_ARG;
  if (cout <= 8) return 8;
  if (cout <= 16) return 16;
  if (cout <= 24) return 24;
  if (cout <= 32) return 32;
  if (cout <= 48) return 48;
  if (cout <= 64) return 64;
  if (cout % 64 == 0) return 64;
  if (cout % 48 == 0) return 48;
  return 64;
}

extern "C" int nvs_conv_cin_chunk(int32_t cin) { return cin <= 4 ? 4 : 8; }

extern "C" int nvs_conv(const NvsConvArgs* a, void* stream) {
  using namespace nvs;
  if (!a || !a->src0 || !a->weight || !a->bias) return NVS_ERR_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0 || a->cout <= 0 || a->c0 <= 0) return NVS_ERR_ARG;
  if (a->B > 65535) return NVS_ERR_ARG;
  if (a->ksize != 1 && a->ksize != 3) return NVS_ERR_UNSUPPORTED;
  if (a->c1 > 0 && !a->src1) return NVS_ERR_ARG;
  if (a->out_mode != NVS_OUT_POOL && !a->dst) return NVS_ERR_ARG;
  if ((a->out_mode == NVS_OUT_POOL || a->out_mode == NVS_OUT_BOTH) && !a->dst2) return NVS_ERR_ARG;
  if (a->in_mode == NVS_IN_S2D && (a->ksize != 1 || a->c1 != 0)) return NVS_ERR_UNSUPPORTED;
  const int cin = (a->in_mode == NVS_IN_S2D) ? 4 * a->c0 : a->c0 + a->c1;
  const int ck = nvs_conv_cin_chunk(cin);
  if (a->c1 > 0 && (a->c0 % ck) != 0) return NVS_ERR_UNSUPPORTED;  // chunk must not straddle sources
  const int ct = nvs_conv_cout_tile(a->cout);
  if (a->out_mode == NVS_OUT_SHUFFLE && (a->cout % 4) != 0) return NVS_ERR_ARG;

  ConvP p;
  p.src0 = a->src0; p.src1 = a->src1; p.w = a->weight; p.bias = a->bias;
  p.dst = a->dst; p.dst2 = a->dst2;
  p.c0_total = a->c0_total; p.c0_off = a->c0_off;
  p.c0 = (a->in_mode == NVS_IN_S2D) ? 4 * a->c0 : a->c0;  // virtual channel count under S2D
  p.c1_total = a->c1_total; p.c1_off = a->c1_off; p.c1 = a->c1;
  p.dst_c_total = a->dst_c_total; p.dst_c_off = a->dst_c_off;
  p.dst2_c_total = a->dst2_c_total; p.dst2_c_off = a->dst2_c_off;
  p.H = a->H; p.W = a->W;
  p.inH = (a->in_mode == NVS_IN_S2D) ? a->in_H : a->H;
  p.inW = (a->in_mode == NVS_IN_S2D) ? a->in_W : a->W;
  if (a->in_mode == NVS_IN_S2D && (a->in_H < 2 * a->H || a->in_W < 2 * a->W)) return NVS_ERR_ARG;
  p.cin_pad = (cin + ck - 1) / ck * ck;
  p.cout = a->cout;
  p.cout_pad = (a->cout + ct - 1) / ct * ct;
  p.act = a->act; p.out_mode = a->out_mode; p.in_mode = a->in_mode;
  p.dst_nhwc = a->dst_nhwc; p.dst2_nhwc = a->dst2_nhwc;
  if (a->out_mode == NVS_OUT_SHUFFLE && a->dst_nhwc) return NVS_ERR_UNSUPPORTED;
  // dst_nhwc == 2: split channels-last format (16-byte stores of 8 fp16 channels); plain outputs only
  if (a->dst_nhwc == 2 && ((a->dst_c_total % 8) || (a->dst_c_off % 8) || a->out_mode != NVS_OUT_PLAIN)) return NVS_ERR_ARG;
  if (a->dst2_nhwc == 2) return NVS_ERR_UNSUPPORTED;
  p.tiles_x = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((a->in_mode == NVS_IN_U8_HWC || a->in_mode == NVS_IN_UNIT) &&
      !(a->ksize == 3 && cin == 3 && a->c0_total == 3 && a->c0_off == 0 && a->c1 == 0 && a->out_mode == NVS_OUT_PLAIN &&
        a->dst_nhwc && a->cout == STEM_CO))
    return NVS_ERR_UNSUPPORTED;  // uint8 / [0,1] frames are read by the stem kernel only
  if (a->ksize == 3 && cin == 3 && a->c1 == 0 &&
      (a->in_mode == NVS_IN_PLAIN || a->in_mode == NVS_IN_U8_HWC || a->in_mode == NVS_IN_UNIT) &&
      a->out_mode == NVS_OUT_PLAIN &&
      a->dst_nhwc && a->cout == STEM_CO && (a->dst_c_total % 4) == 0 && (a->dst_c_off % 4) == 0 &&
      (a->act == NVS_ACT_NONE || a->act == NVS_ACT_LRELU || a->act == NVS_ACT_RELU)) {
    dim3 grid((a->W + STEM_TX - 1) / STEM_TX, (a->H + STEM_TY - 1) / STEM_TY, a->B);
    stem_conv_kernel<<<grid, 128, 0, st>>>(p);
    NVS_CHECK_LAUNCH();
    return NVS_OK;
  }
  if (a->ksize == 3) {
    return ck == 4 ? dispatch_ct<3, 4>(p, a->B, ct, st) : dispatch_ct<3, 8>(p, a->B, ct, st);
  }
  return ck == 4 ? dispatch_ct<1, 4>(p, a->B, ct, st) : dispatch_ct<1, 8>(p, a->B, ct, st);
}
