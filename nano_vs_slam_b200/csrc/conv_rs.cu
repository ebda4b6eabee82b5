// 3x3 convolution, row-stationary implicit GEMM on tcgen05 with "3xFP16" operands (fp32-grade accuracy):
//
//     a = a_hi + a_lo,  w * 2^t = w_hi + w_lo   (hi = value rounded to fp16, lo = fp16 of the exact remainder: 22
//     significant bits, what 3xTF32 keeps),  D += a_hi w_hi + a_hi w_lo + a_lo w_hi,  fp32 accumulation in TMEM.
//
// Compared with the 3xTF32 kernels of conv_tc.cu a kind::f16 MMA contracts K = 16 per instruction instead of 8 (half
// the instructions for the same work -- the MMA count is what bounds those kernels) and the operands are half as wide.
// fp16 has a 5-bit exponent: weights are scaled per layer by a power of two 2^t (undone in the epilogue, exact) so that
// neither part is subnormal; activations are used as they are -- |a| must stay below 65504 (the epilogue raises a
// sticky device flag when an output it writes does not), and the remainder of a small activation is quantised at
// 2^-25 ABSOLUTE, far below the 1e-4-of-the-map's-maximum the outputs are held to.
//
// Row-stationary formulation (the ROW3 variant of conv_tc.cu, here for every layer): the GEMM row is an input position,
// the three taps of a kernel row are one N = 3 x Cout operand ([kx = 0 | 1 | 2] x Cout) and
// out[i] = D_0[i-1] + D_1[i] + D_2[i+1] is formed by two warp shuffles per channel in the epilogue.  Tile = 4 image
// rows x 30 columns (32 positions per row with the halo columns, one row per TMEM lane quadrant).  What is new:
//
//   * the A operand is read from SHARED memory (SS-mode MMA).  The 6 x 32 halo box of a 32-channel chunk is one
//     contiguous K-major matrix of 192 rows, and the operand of kernel row ky is simply its rows [32 ky, 32 ky + 128):
//     every input value is converted ONCE per tile and chunk (the TMEM-fed kernels convert it once per tap, 9x / 3x),
//     by three converter warps that read the fp32 box rows TMA delivered and write the fp16 hi / lo tiles in the
//     canonical SWIZZLE_64B layout;
//   * all three products go to ONE accumulator (no scaled correction half): 3 x Cout columns per stage, two stages;
//     Cout = 32 layers issue a_hi x [W_hi ; W_lo] as one N = 192 instruction (the epilogue adds the halves);
//   * one MMA-issuing thread (N = 192 MMAs cost their nominal 96 cycles, more than a thread needs to queue one):
//     the accumulation order is fixed, results are bit-reproducible run to run;
//   * weights stay resident in shared memory for the whole kernel when the layer's 3 x chunks tiles fit (up to
//     64 input channels at Cout = 64), else they stream through a ring.
//
// Warps: 0-7 epilogue (quadrant = warp % 4, channel half = warp / 4), 8-10 converters (box rows round-robin; each also
// issues the TMA loads of its rows), 11 MMA issuer (allocates TMEM), 12 weight TMA.  Persistent CTAs, one per SM.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace nvs {
namespace rs {

constexpr int TX = 30, TY = 4;           // output pixels per tile
constexpr int HX = 32, HY = 6;           // halo box: positions per row (= GEMM rows per quadrant), rows
constexpr int KC = 32;                   // input channels per chunk
constexpr int ROW_BYTES = HX * KC * 4;   // one fp32 box row: 32 positions x 128 B
#ifndef NVS_RS_NB
#define NVS_RS_NB 6
#endif
#ifndef NVS_RS_CONV_WARPS
#define NVS_RS_CONV_WARPS 3
#endif
constexpr int NB = NVS_RS_NB;            // box rows in flight
constexpr int A_HALF = HX * HY * KC * 2; // one of the hi / lo operand matrices: 192 rows x 64 B
constexpr int A_BYTES = 2 * A_HALF;
#ifndef NVS_RS_NA
#define NVS_RS_NA 2
#endif
constexpr int NA = NVS_RS_NA;            // converted chunks in flight
constexpr int ACC_STAGES = 2;
constexpr int EPI_WARPS = 8;
// NB is a multiple of CONV_WARPS: a ring slot is then always read by the same converter warp, so a warp can never wait
// for a phase of a row barrier that is two ahead of the barrier's current one (a parity wait cannot tell that from
// "already complete": with 4 warps on 6 slots a fast warp read a row that had not landed yet)
constexpr int WARP_CONV = EPI_WARPS, CONV_WARPS = NVS_RS_CONV_WARPS;
static_assert(NB % CONV_WARPS == 0 && HY >= CONV_WARPS, "row ring / converter warps");
constexpr int WARP_MMA = WARP_CONV + CONV_WARPS, WARP_TMA_W = WARP_MMA + 1;
constexpr int THREADS = 32 * (WARP_TMA_W + 1);
constexpr int ROWS_AHEAD = NB / CONV_WARPS;  // row loads a converter warp keeps in flight (its slots of the ring)
constexpr unsigned long long PLAN_MAGIC = 0x7273506C616E0001ull;  // first word of an rs::Plan

template <int CO>
struct Cfg {
  static_assert(CO == 32 || CO == 64, "output channels per launch");
  static constexpr int NW = 3 * CO;                 // GEMM N of one of W_hi / W_lo: [kx][cout]
  static constexpr int W_HALF = NW * KC * 2;        // one of W_hi / W_lo of one (chunk, ky): NW rows x 64 B
  static constexpr int W_STAGE = 2 * W_HALF;
  static constexpr bool CONCAT = CO == 32;          // a_hi x [W_hi ; W_lo] as one N = 2 NW instruction
  static constexpr int ACC_COLS = 192;              // per stage: CONCAT 2 x 96, else 192
  static constexpr int CW = CO / 2;                 // output channels per epilogue warp
  static constexpr int SM_ROWS = 0;
  static constexpr int SM_A = SM_ROWS + NB * ROW_BYTES;
  static constexpr int SM_BIAS = SM_A + NA * A_BYTES;
  static constexpr int SM_POOL = SM_BIAS + CO * 4;
  static constexpr int POOL_BYTES = 2 * 2 * 15 * 16 * 4;   // max-pool exchange: (channel half, quadrant pair) x 15 x 16
  static constexpr int SM_BAR = SM_POOL + POOL_BYTES;
  static constexpr int MAX_WS = 16;                  // barrier slots reserved for the weight ring
  static constexpr int N_BARS = 2 * NB + 2 * NA + 2 * MAX_WS + 4 + 1;
  static constexpr int SM_W = ((SM_BAR + 8 * N_BARS + 16 + 1023) / 1024) * 1024;
  static constexpr int MAX_W_STAGES = (227 * 1024 - 1024 - SM_W) / W_STAGE;
  static constexpr int smem_bytes(int w_stages) { return SM_W + w_stages * W_STAGE + 1024; }
  static_assert(MAX_W_STAGES >= 3 && MAX_W_STAGES <= MAX_WS, "weight ring");
  static constexpr uint32_t idesc(int n) {  // kind::f16: D fp32, A / B fp16, both K-major
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  }
};

struct Params {
  const float* bias;
  float* dst;
  float* dst_pool;
  int* range_flag;   // sticky: set when an output written in an activation layout exceeds the fp16 range
  int c0_off, c0_chunks, c1_off, c1_chunks;
  int dst_c_total, dst_c_off, dst_layout, dst_mode;
  int pool_c_total, pool_c_off;
  int B, H, W, cout, act;
  int tiles_x, tiles_y, n_tiles;
  int nk_last0, nk_last1;  // k-steps (16 channels) of the last chunk of source 0 / 1 that can be non-zero
  int w_stages;            // weight ring depth; >= 3 * chunks means resident (every tile loaded once)
  float w_scale;           // 2^-t: undoes the weights' power-of-two scale
  long long* dbg;          // env NVS_RS_DBG=1: per-role clock64 stamps of CTA 0 (nvs_conv_rs_debug_buffer), 256 per role
  int knock;               // bottleneck experiments (env NVS_RS_KNOCK, results are then garbage): 1 converters only
                           // pass barriers, 2 epilogue only passes barriers, 4 no MMAs, 8 no activation TMA
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();  // never hang the box on a protocol bug
  }
}
// single-thread roles with slack (TMA producers): back off so the poll does not take issue slots from working warps
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifndef NVS_RS_NO_SLEEP
    __nanosleep(64);
#endif
    if (++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand with 64-byte rows (32 fp16), SWIZZLE_64B (layout 4), 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* a) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = __uint_as_float(r[i]);
}
// two 16-column loads (the two accumulator halves of the same channels), one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr_a, uint32_t taddr_b, float* a, float* b) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr_a)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
        "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr_b)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(q[i]);
  }
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

template <int CO>
__global__ void __launch_bounds__(THREADS, 1)
conv_rs_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
               const Params p) {
  using C = Cfg<CO>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(sm + C::SM_BIAS);
  const uint32_t bar0 = base + C::SM_BAR;
  auto rfull = [&](int i) { return bar0 + 8u * i; };
  auto afull = [&](int i) { return bar0 + 8u * (2 * NB + i); };
  auto aempty = [&](int i) { return bar0 + 8u * (2 * NB + NA + i); };
  auto wfull = [&](int i) { return bar0 + 8u * (2 * NB + 2 * NA + i); };
  auto wempty = [&](int i) { return bar0 + 8u * (2 * NB + 2 * NA + C::MAX_WS + i); };
  auto accfull = [&](int i) { return bar0 + 8u * (2 * NB + 2 * NA + 2 * C::MAX_WS + i); };
  auto accempty = [&](int i) { return bar0 + 8u * (2 * NB + 2 * NA + 2 * C::MAX_WS + 2 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::SM_BAR + 8 * C::N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = p.c0_chunks + p.c1_chunks;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int w_tiles = 3 * chunks;                 // (chunk, ky) weight tiles of the layer
  const bool w_resident = p.w_stages >= w_tiles;  // every tile has its own stage: loaded once

  for (int i = threadIdx.x; i < CO; i += THREADS) bias_s[i] = p.bias[i];
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
    for (int i = 0; i < NB; ++i) {
      mbar_init(rfull(i), 1);
    }
    for (int i = 0; i < NA; ++i) {
#ifdef NVS_RS_WARP_ARRIVE
      mbar_init(afull(i), HY);          // lane 0 of the converter warp, after every lane's proxy fence and a warp barrier
#else
      mbar_init(afull(i), HY * 32);     // every converter lane, after its own proxy fence
#endif
      mbar_init(aempty(i), 1);          // tcgen05.commit of the chunk's MMAs
    }
    for (int i = 0; i < C::MAX_WS; ++i) {
      mbar_init(wfull(i), 1);
      mbar_init(wempty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(accfull(i), 1);
      mbar_init(accempty(i), EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARPS) {
    // =========================== epilogue ===========================
    // quadrant (= image row of the tile) warp % 4, channel half h = warp / 4; lane = position of the 32-wide strip
    // (lane 0 / 31 are halo positions).  Accumulator columns of a stage: [kx 0 | kx 1 | kx 2] x CO channels (CONCAT:
    // then the same three blocks of a_hi w_lo + a_lo w_hi).  out[i] = D_0[i-1] + D_1[i] + D_2[i+1].
    int acc = 0;
    uint32_t aph = 0;
    const int quad = warp & 3, h = warp >> 2;
    const float neg_slope = p.act == NVS_ACT_LRELU ? 0.01f : (p.act == NVS_ACT_RELU ? 0.f : 1.f);
    float seen_max = 0.f;
    int dbg_i = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && dbg_i < 256) p.dbg[dbg_i++] = clock64();
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
      const int gx = tx * TX - 1 + lane, gy = ty * TY + quad;
      const bool valid = lane >= 1 && lane <= TX && gx < p.W && gy < p.H;
      if (!(p.knock & 32)) mbar_wait(accfull(acc), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * C::ACC_COLS) + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int part = 0; part < ((p.knock & 2) ? 0 : C::CW / 16); ++part) {
        const int cbase = C::CW * h + 16 * part;  // first of this pass's 16 output channels
        if (p.dst_mode == 3 && cbase != 0) continue;  // keypoint heads: 3 channels in all
        if (cbase >= p.cout) continue;
        float o[16], u[16];
        if (C::CONCAT) {
          float w[16];
          tmem_ld16x2(taddr + (uint32_t)cbase, taddr + (uint32_t)(C::NW + cbase), u, w);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __shfl_up_sync(0xffffffffu, u[j] + w[j], 1);
          tmem_ld16x2(taddr + (uint32_t)(CO + cbase), taddr + (uint32_t)(C::NW + CO + cbase), u, w);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] += u[j] + w[j];
          tmem_ld16x2(taddr + (uint32_t)(2 * CO + cbase), taddr + (uint32_t)(C::NW + 2 * CO + cbase), u, w);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] += __shfl_down_sync(0xffffffffu, u[j] + w[j], 1);
        } else {
          tmem_ld16(taddr + (uint32_t)cbase, u);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __shfl_up_sync(0xffffffffu, u[j], 1);
          tmem_ld16(taddr + (uint32_t)(CO + cbase), u);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] += u[j];
          tmem_ld16(taddr + (uint32_t)(2 * CO + cbase), u);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] += __shfl_down_sync(0xffffffffu, u[j], 1);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = fmaf(o[j], p.w_scale, bias_s[cbase + j]);
          o[j] = fmaxf(a, 0.f) + neg_slope * fminf(a, 0.f);
        }
        if (p.act == NVS_ACT_SIGMOID && cbase == 0) {  // depth heads: cout <= 4
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = 1.f / (1.f + expf(-o[j]));
        }
        if (p.dst_mode == 3) {  // keypoint heads: sigmoid -> score, tanh -> centre shift (kp2dtiny.py:574-575, 927-935)
          if (valid) {
            const size_t plane = (size_t)p.H * p.W, pix = (size_t)gy * p.W + gx;
            p.dst[(size_t)b * plane + pix] = 1.f / (1.f + expf(-o[0]));
            p.dst_pool[((size_t)b * 2 + 0) * plane + pix] = tanhf(o[1]);
            p.dst_pool[((size_t)b * 2 + 1) * plane + pix] = tanhf(o[2]);
          }
          continue;
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) seen_max = fmaxf(seen_max, fabsf(o[j]));
        }
        if (p.dst_pool != nullptr) {
          // MaxPool2d(2,2): x partner = next lane (strips start at even x, so pairs are lanes (1,2), (3,4), ...),
          // y partner = the row of the next quadrant's warp: odd quadrants hand their x-pooled values to the even
          // quadrant below them through shared memory (one buffer and one named barrier per (h, quadrant pair);
          // the pair's second barrier keeps the next write behind this read)
          float* ps = reinterpret_cast<float*>(sm + C::SM_POOL) + (h * 2 + (quad >> 1)) * (15 * 16);
          const int bar_id = 2 + h * 2 + (quad >> 1);
          float m[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = fmaxf(o[j], __shfl_down_sync(0xffffffffu, o[j], 1));
          const int pc = (lane - 1) >> 1;  // pooled column inside the strip, odd lanes 1..29 -> 0..14
          const bool owner = (lane & 1) && lane <= 29;
          if ((quad & 1) && owner) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<float4*>(ps + pc * 16)[q] = make_float4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          const int qx = gx >> 1, qy = gy >> 1;
          if (!(quad & 1) && owner && qx < (p.W >> 1) && qy < (p.H >> 1)) {
            float4* d = reinterpret_cast<float4*>(
                p.dst_pool + (((size_t)b * (p.H >> 1) + qy) * (p.W >> 1) + qx) * p.pool_c_total + p.pool_c_off + cbase);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 r = reinterpret_cast<const float4*>(ps + pc * 16)[q];
              if (cbase + 4 * q < p.cout)
                d[q] = make_float4(fmaxf(m[4 * q], r.x), fmaxf(m[4 * q + 1], r.y), fmaxf(m[4 * q + 2], r.z),
                                   fmaxf(m[4 * q + 3], r.w));
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        }
        if (valid && p.dst_mode == 1) {
          if (p.dst_layout == 0) {  // NHWC
            float4* d = reinterpret_cast<float4*>(
                p.dst + (((size_t)b * p.H + gy) * p.W + gx) * p.dst_c_total + p.dst_c_off + cbase);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (cbase + 4 * q < p.cout) d[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          } else {  // NCHW: a warp writes 30 consecutive x of one channel row
            float* d = p.dst + (((size_t)b * p.dst_c_total + p.dst_c_off + cbase) * p.H + gy) * p.W + gx;
            const size_t plane = (size_t)p.H * p.W;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cbase + j < p.cout) d[j * plane] = o[j];
          }
        } else if (valid && p.dst_mode == 2) {  // PixelShuffle(2) -> NHWC (B, 2H, 2W, cout/4): channel c -> (c%4/2, c%2, c/4)
          const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float4* d = reinterpret_cast<float4*>(
                  p.dst + (((size_t)b * H2 + 2 * gy + i) * W2 + 2 * gx + j) * p.dst_c_total + p.dst_c_off + cbase / 4);
              const int s = 2 * i + j;
              d[0] = make_float4(o[s], o[4 + s], o[8 + s], o[12 + s]);
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && !(p.knock & 32)) mbar_arrive(accempty(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        aph ^= 1;
      }
    }
    // an activation beyond the fp16 range would make the next layer's a_hi infinite: report it (sticky flag)
    if (p.range_flag != nullptr && p.dst_layout == 0 && p.dst_mode != 3) {
      seen_max = warp_max(seen_max);
      if (lane == 0 && !(seen_max < 60000.f)) atomicOr(p.range_flag, 1);
    }
  } else if (warp < WARP_MMA) {
    // =========================== converters: fp32 box row -> fp16 hi / lo rows of the A tile ===========================
    const int cw = warp - WARP_CONV;
    // ring positions are kept incrementally (no divisions: the single-thread roles are latency critical, and a 64-bit
    // division by a run-time value costs hundreds of cycles).  This warp takes rows cw, cw + CONV_WARPS, ... of the
    // CTA's row sequence; a chunk has HY rows, a row slot ring NB entries, the A ring NA entries.
    int rs = cw % NB, row = cw % HY, as = 0;
    uint32_t rph = 0, aph = 0;
    const int my_chunks = my_tiles * chunks;
    int cdone = 0;
    auto advance = [&]() {  // to this warp's next row
      rs += CONV_WARPS;
      if (rs >= NB) {
        rs -= NB;
        rph ^= 1;
      }
      row += CONV_WARPS;
      if (row >= HY) {       // next chunk (CONV_WARPS <= HY: at most one chunk boundary per step)
        row -= HY;
        ++cdone;
        if (++as == NA) {
          as = 0;
          aph ^= 1;
        }
      }
    };
    // The warp also LOADS its rows: lane 0 issues the TMA of the row that will use a slot next as soon as the warp has
    // read the slot (a separate producer thread needs a round trip through an "empty" barrier per row, and one thread
    // issuing six loads per chunk plus their address arithmetic was the slowest stage of the kernel).
    int p_row = cw % HY, p_ch = 0, p_t = blockIdx.x, p_cdone = 0;
    int p_tx = p_t % p.tiles_x, p_ty = (p_t / p.tiles_x) % p.tiles_y, p_b = p_t / (p.tiles_x * p.tiles_y);
    auto issue_next = [&](int slot) {  // lane 0: load this warp's next not yet requested row into `slot`
      if (p_cdone >= my_chunks) return;
      if (p.knock & 8) {
        mbar_arrive(rfull(slot));
      } else {
        mbar_expect_tx(rfull(slot), ROW_BYTES);
        const uint32_t dst = base + C::SM_ROWS + slot * ROW_BYTES;
        if (p_ch < p.c0_chunks)
          tma_load_4d(dst, &map_a0, rfull(slot), p.c0_off + p_ch * KC, p_tx * TX - 1, p_ty * TY - 1 + p_row, p_b);
        else
          tma_load_4d(dst, &map_a1, rfull(slot), p.c1_off + (p_ch - p.c0_chunks) * KC, p_tx * TX - 1,
                      p_ty * TY - 1 + p_row, p_b);
      }
      p_row += CONV_WARPS;
      if (p_row >= HY) {
        p_row -= HY;
        ++p_cdone;
        if (++p_ch == chunks) {
          p_ch = 0;
          p_t += gridDim.x;
          p_tx = p_t % p.tiles_x;
          p_ty = (p_t / p.tiles_x) % p.tiles_y;
          p_b = p_t / (p.tiles_x * p.tiles_y);
        }
      }
    };
    if (lane == 0) {
#pragma unroll 1
      for (int d = 0; d < ROWS_AHEAD; ++d) issue_next((cw + d * CONV_WARPS) % NB);
    }
    int dbg_i = 0;
    while (cdone < my_chunks) {
      if (p.dbg && blockIdx.x == 0 && cw == 0 && lane == 0 && dbg_i < 256) p.dbg[256 + dbg_i++] = clock64();
      mbar_wait(rfull(rs), rph);
      if (p.knock & 1) {
        if (!(p.knock & 16)) {
          if (row < CONV_WARPS) mbar_wait(aempty(as), aph ^ 1);
          mbar_arrive(afull(as));
        }
        __syncwarp();
        if (lane == 0) issue_next(rs);
        advance();
        continue;
      }
      // position `lane` of the row: 128 bytes, 16-byte chunk c at (c ^ (lane & 7)) (SWIZZLE_128B)
      const uint8_t* src = sm + C::SM_ROWS + rs * ROW_BYTES + lane * 128;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(src + ((c ^ (lane & 7)) << 4));
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        hi[2 * c] = h2_bits(h0);
        hi[2 * c + 1] = h2_bits(h1);
        lo[2 * c] = h2_bits(__floats2half2_rn(v.x - f0.x, v.y - f0.y));
        lo[2 * c + 1] = h2_bits(__floats2half2_rn(v.z - f1.x, v.w - f1.y));
      }
      // a warp's first row of a chunk is one of the chunk's first CONV_WARPS rows: the MMAs that last read this A
      // slot (two chunks ago) must have retired before it is overwritten
      if (row < CONV_WARPS) mbar_wait(aempty(as), aph ^ 1);
      // A tile row r = 32 * row + lane: 64 bytes, 16-byte chunk j (8 channels) at (j ^ ((r >> 1) & 3)) (SWIZZLE_64B)
      const int r = 32 * row + lane;
      uint8_t* dh = sm + C::SM_A + as * A_BYTES + r * 64;
      const int key = (r >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(dh + ((j ^ key) << 4)) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        *reinterpret_cast<uint4*>(dh + A_HALF + ((j ^ key) << 4)) =
            make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
      }
      proxy_fence();            // generic-proxy stores -> visible to the tensor core's operand reads
#ifdef NVS_RS_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) mbar_arrive(afull(as));
#else
      mbar_arrive(afull(as));
      __syncwarp();
#endif
      if (lane == 0) issue_next(rs);  // every lane's reads of the slot precede its proxy fence and the warp barrier above
      advance();
    }
  } else if (warp == WARP_TMA_W) {
    // =========================== weight producer: [W_hi ; W_lo] of one (chunk, ky) per stage ===========================
    if (lane == 0 && my_tiles > 0 && !(p.knock & 64)) {
      const int rounds = w_resident ? 1 : my_tiles;
      int ws = 0;
      uint32_t wph = 0;
      for (int r = 0; r < rounds; ++r)
        for (int ch = 0; ch < chunks; ++ch)
          for (int ky = 0; ky < 3; ++ky) {
            if (!w_resident) mbar_wait_relaxed(wempty(ws), wph ^ 1u);
            const uint32_t dst = base + C::SM_W + ws * C::W_STAGE;
            mbar_expect_tx(wfull(ws), C::W_STAGE);
            tma_load_3d(dst, &map_whi, wfull(ws), ch * KC, 0, ky);
            tma_load_3d(dst + C::W_HALF, &map_wlo, wfull(ws), ch * KC, 0, ky);
            if (++ws == p.w_stages) {
              ws = 0;
              wph ^= 1;
            }
          }
    }
  } else if (lane == 0 && my_tiles > 0) {
    // =========================== MMA issuer ===========================
    int dbg_i = 0;
    int ws = 0, as = 0, acc = 0;        // weight stage, A slot, accumulator stage: kept incrementally
    uint32_t wph = 0, aph = 0, cph = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      if (!(p.knock & 32)) mbar_wait(accempty(acc), cph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C::ACC_COLS);
      for (int ch = 0; ch < chunks; ++ch) {
        if (p.dbg && blockIdx.x == 0 && dbg_i < 256) p.dbg[512 + dbg_i++] = clock64();
        if (!(p.knock & 16)) mbar_wait(afull(as), aph);
        tc_fence_after();
        const uint32_t a_base = base + C::SM_A + as * A_BYTES;
        const int nk = ch == p.c0_chunks - 1 ? p.nk_last0 : (ch == chunks - 1 ? p.nk_last1 : 2);
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
          if ((!w_resident || tl == 0) && !(p.knock & 64)) mbar_wait(wfull(ws), wph);
          const uint64_t a_hi = make_desc64(a_base + ky * (HX * 64)), a_lo = a_hi + (uint64_t)(A_HALF >> 4);
          const uint64_t w_hi = make_desc64(base + C::SM_W + ws * C::W_STAGE), w_lo = w_hi + (uint64_t)(C::W_HALF >> 4);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < nk && !(p.knock & 4)) {
              const uint64_t o = (uint64_t)(2 * k);  // 16 fp16 = 32 bytes along K = +2 in the (address >> 4) field
              const uint32_t first = (ch | ky | k) != 0 ? 1u : 0u;
              if (C::CONCAT) {
                tc_mma_f16(d_tmem, a_hi + o, w_hi + o, C::idesc(2 * C::NW), first);            // x [W_hi ; W_lo]
                tc_mma_f16(d_tmem + (uint32_t)C::NW, a_lo + o, w_hi + o, C::idesc(C::NW), 1u);
              } else {
                tc_mma_f16(d_tmem, a_hi + o, w_hi + o, C::idesc(C::NW), first);
                tc_mma_f16(d_tmem, a_hi + o, w_lo + o, C::idesc(C::NW), 1u);
                tc_mma_f16(d_tmem, a_lo + o, w_hi + o, C::idesc(C::NW), 1u);
              }
            }
          }
          if (!w_resident && !(p.knock & 64)) tc_commit(wempty(ws));
          if (++ws == p.w_stages) {
            ws = 0;
            wph ^= 1;
          }
        }
        if (!(p.knock & 16)) tc_commit(aempty(as));
        if (++as == NA) {
          as = 0;
          aph ^= 1;
        }
      }
      if (!(p.knock & 32)) tc_commit(accfull(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        cph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

struct alignas(64) Plan {
  unsigned long long magic;   // tells an rs::Plan from a tc::Plan (whose first bytes are a tensor map) in the same memory
  int cout_tpl;
  alignas(64) CUtensorMap a0, a1, whi, wlo;
  Params p;
};

// NHWC fp32 activations (B,H,W,Ct): box = (32 channels, 32 positions, 1 row, 1 frame); channels beyond Ct (the
// 16-channel stem output) and positions outside the image are zero filled by TMA
static int encode_act(CUtensorMap* m, const float* ptr, int B, int H, int W, int Ct) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Ct * 4, (cuuint64_t)W * Ct * 4, (cuuint64_t)H * W * Ct * 4};
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)HX, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}
// packed fp16 weights [3 ky][3 x cout_pad rows][cin]: box = (32 channels, 3 x cout_pad, 1)
static int encode_w(CUtensorMap* m, const void* ptr, int cin, int rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)rows, 3};
  cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)rows * cin * 2};
  cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}

static long long* g_dbg = nullptr;  // nvs_conv_rs_debug_buffer

template <int CO>
static int launch(const Plan& pl, const Params& p, cudaStream_t st) {
  using C = Cfg<CO>;
  auto kern = conv_rs_kernel<CO>;
  NVS_OPT_IN_SMEM(kern, 227 * 1024);
  const int sms = nvs_sm_count();
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  Params q = p;
  const int w_tiles = 3 * (p.c0_chunks + p.c1_chunks);
  q.w_stages = w_tiles <= C::MAX_W_STAGES ? w_tiles : C::MAX_W_STAGES;
  {
    const char* e = getenv("NVS_RS_KNOCK");
    q.knock = e ? atoi(e) : 0;
    q.dbg = g_dbg;
  }
  kern<<<grid, THREADS, C::smem_bytes(q.w_stages), st>>>(pl.a0, pl.a1, pl.whi, pl.wlo, q);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

// sticky per-device flag the epilogues raise when an activation leaves the fp16 range (nvs_conv_rs_range_flag)
static int* range_flag_ptr() {
  static int* flags[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!flags[dev]) {
    int* f = nullptr;
    if (cudaMalloc(&f, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(f, 0, sizeof(int));
    flags[dev] = f;
  }
  return flags[dev];
}

size_t plan_bytes() { return sizeof(Plan) + 64; }

bool is_plan(const void* plan_mem) {
  const Plan* pl = reinterpret_cast<const Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  return pl->magic == PLAN_MAGIC;
}

int plan_init(void* plan_mem, const NvsConvTcArgs* a) {
  if (a->cout <= 0 || a->cout > 64) return NVS_ERR_UNSUPPORTED;
  const int c0 = a->c0 == 16 ? 32 : a->c0;  // a 16-channel source is read as one chunk whose upper half TMA zero-fills
  if (a->c0 == 16 && (a->c1 != 0 || a->c0_off != 0 || a->c0_total != 16)) return NVS_ERR_ARG;
  if ((c0 % KC) || (a->c1 % KC)) return NVS_ERR_UNSUPPORTED;
  if (!(a->w_scale > 0.f)) return NVS_ERR_ARG;
  Plan* pl = reinterpret_cast<Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  const int cpad = a->cout <= 32 ? 32 : 64;
  if (a->dst_mode == 2 && (a->cout % 16) != 0) return NVS_ERR_UNSUPPORTED;
  int rc = encode_act(&pl->a0, a->src0, a->B, a->H, a->W, a->c0_total);
  if (rc != NVS_OK) return rc;
  rc = encode_act(&pl->a1, a->c1 > 0 ? a->src1 : a->src0, a->B, a->H, a->W, a->c1 > 0 ? a->c1_total : a->c0_total);
  if (rc != NVS_OK) return rc;
  const int cin = c0 + a->c1;
  rc = encode_w(&pl->whi, a->w_hi, cin, 3 * cpad);
  if (rc != NVS_OK) return rc;
  rc = encode_w(&pl->wlo, a->w_lo, cin, 3 * cpad);
  if (rc != NVS_OK) return rc;
  Params& p = pl->p;
  p.bias = a->bias; p.dst = a->dst; p.dst_pool = a->dst_pool;
  p.range_flag = range_flag_ptr();
  p.c0_off = a->c0_off; p.c0_chunks = c0 / KC; p.c1_off = a->c1_off; p.c1_chunks = a->c1 / KC;
  p.dst_c_total = a->dst_c_total; p.dst_c_off = a->dst_c_off; p.dst_layout = a->dst_layout; p.dst_mode = a->dst_mode;
  p.pool_c_total = a->pool_c_total; p.pool_c_off = a->pool_c_off;
  p.B = a->B; p.H = a->H; p.W = a->W; p.cout = a->cout; p.act = a->act;
  p.tiles_x = (a->W + TX - 1) / TX; p.tiles_y = (a->H + TY - 1) / TY;
  p.n_tiles = p.tiles_x * p.tiles_y * a->B;
  auto last_nk = [&](int c, int real) {  // k-steps (16 channels) of the last chunk that can be non-zero
    if (real <= 0 || real >= c) return 2;
    const int in_last = real - (c / KC - 1) * KC;
    if (in_last <= 0) return 2;
    return (in_last + 15) / 16;
  };
  p.nk_last0 = a->c0 == 16 ? 1 : last_nk(c0, a->c0_real);
  p.nk_last1 = a->c1 > 0 ? last_nk(a->c1, a->c1_real) : p.nk_last0;
  p.w_stages = 0;
  p.w_scale = a->w_scale;
  p.knock = 0;
  p.dbg = nullptr;
  pl->cout_tpl = cpad;
  pl->magic = PLAN_MAGIC;
  return NVS_OK;
}

int run(const void* plan_mem, float* dst_override, float* dst2_override, cudaStream_t st) {
  const Plan* pl = reinterpret_cast<const Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  Params p = pl->p;
  if (dst_override) p.dst = dst_override;
  if (dst2_override) p.dst_pool = dst2_override;
  if (p.dst_mode == 3 && (!p.dst || !p.dst_pool)) return NVS_ERR_ARG;
  if (p.dst_mode != 0 && !p.dst) return NVS_ERR_ARG;
  return pl->cout_tpl == 32 ? launch<32>(*pl, p, st) : launch<64>(*pl, p, st);
}

}  // namespace rs
}  // namespace nvs

// 1 if any row-stationary conv on the current device has written an activation beyond the fp16 range since the last
// reset (the following layer's operands were then infinite: rerun with NVS_CONV_MATH=tf32); *synchronises the device*.
// debugging aid (tools/rs_timeline.py): device buffer of 768 clock64 stamps written by CTA 0 of every following launch
extern "C" void nvs_conv_rs_debug_buffer(long long* dev_buf) { nvs::rs::g_dbg = dev_buf; }

extern "C" int nvs_conv_rs_range_flag(int32_t reset) {
  int* f = nvs::rs::range_flag_ptr();
  if (!f) return -1;
  int v = 0;
  if (cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (reset && v) cudaMemset(f, 0, sizeof(int));
  return v;
}
