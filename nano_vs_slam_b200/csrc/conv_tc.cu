// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores with fp32-grade accuracy
// ("3xTF32": a = a_hi + a_lo, w = w_hi + w_lo, D += a_hi w_hi + a_lo w_hi + a_hi w_lo, fp32 accumulation
// in TMEM; the dropped a_lo w_lo term and the tf32 rounding of the lo parts are ~2^-20 relative, the same
// order as fp32 FFMA accumulation error over K = 9*Cin terms).
//
//   GEMM view:  M = 128 output pixels (8 rows x 16 cols of one frame),  N = Cout,  K = 9 taps x Cin
//   A operand:  NHWC activations.  For tap (ky,kx) and a 32-channel chunk, ONE 4-D TMA box
//               (32 ch, 16 x, 8 y, 1 frame) at (x0+kx-1, y0+ky-1) lands as 128 rows x 128 bytes in the
//               canonical K-major SWIZZLE_128B layout; out-of-image coordinates are zero filled by TMA,
//               which *is* the convolution's zero padding.  No im2col buffer, no halo bookkeeping.
//   B operand:  packed weights [tap][Cout][Cin] (hi and lo parts), 3-D TMA box (32, Cout, 1).
//   split:      4 converter warps read the landed fp32 tile (one 128-byte swizzled row per thread), form
//               hi (low 13 mantissa bits cleared) and lo = a - hi and write both straight into TENSOR MEMORY
//               with tcgen05.st (lane = pixel, column = channel): only ONE copy of the activations crosses
//               L2/HBM, and the A operand never goes back through shared memory (the first version of this
//               kernel was shared-memory-bandwidth bound: TMA writes + split round trip + 3 UMMA operand reads).
//   MMA:        one thread issues 12 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=Cout, K=8) per stage with
//               A from TMEM and B (weights) from shared memory
//   epilogue:   4 warps read TMEM (thread = pixel, columns = channels), add bias, activation, then store
//               NHWC / NCHW, 2x2 max-pool via warp shuffles (a warp owns 2 image rows x 16 cols) or
//               pixel-shuffled NHWC.
//
// Warp roles: 0-3 epilogue (TMEM lane quadrants), then CONV_GROUPS x 4 converter warps (groups take pipeline
// stages round-robin), one TMA producer warp, one TMEM-allocator + MMA-issuer warp.  Persistent CTAs, static round-robin over (frame, tile_y, tile_x).
#include <cuda.h>

#include "common.cuh"

namespace nvs {
namespace tc {

constexpr int TM = 128;            // pixels per tile
constexpr int TX = 16, TY = 8;     // tile shape
constexpr int KC = 32;             // channels per stage (128 bytes of fp32)
constexpr int A_BYTES = TM * KC * 4;   // 16 KiB
constexpr int CONV_GROUPS = 2;     // converter warp groups (4 warps each) taking pipeline stages round-robin
constexpr int WARP_TMA = 4 + 4 * CONV_GROUPS;
constexpr int WARP_MMA = WARP_TMA + 1;
constexpr int THREADS = 32 * (WARP_MMA + 1);

template <int COUT>
struct Cfg {
  static constexpr int W_BYTES = COUT * KC * 4;
  static constexpr int STAGE_BYTES = A_BYTES + 2 * W_BYTES;  // fp32 activation tile + W_hi + W_lo
  // tcgen05.mma costs max(N/2, ~50) cycles per instruction (profiles/microbench/mma_rate.cu), so for
  // Cout <= 64 the two products that share A_hi are issued as ONE instruction with N = 2*Cout against the
  // concatenated [W_hi ; W_lo] tile (they are contiguous in shared memory): columns [0,Cout) collect
  // a_hi w_hi + a_lo w_hi, columns [Cout,2Cout) collect a_hi w_lo, and the epilogue adds the two halves.
  static constexpr bool CONCAT = COUT <= 64;
  static constexpr int ACC_STAGE_COLS = CONCAT ? 2 * COUT : COUT;
  static constexpr int ACC_COLS = 2 * ACC_STAGE_COLS;   // double-buffered accumulator
  static constexpr int A_COLS = 2 * KC;                 // per stage: 32 columns hi + 32 columns lo
  // pipeline depth: bounded by TMEM (ACC_COLS + STAGES*A_COLS <= 512) and by ~200 KB of shared memory
  static constexpr int STAGES = (COUT <= 32) ? 6 : 4;
  static constexpr int TMEM_COLS = 512;
  static_assert(ACC_COLS + STAGES * A_COLS <= 512, "TMEM budget");
  static_assert(STAGES * STAGE_BYTES <= 200 * 1024, "smem budget");
  static constexpr int SM_BIAS = STAGES * STAGE_BYTES;
  static constexpr int SM_BAR = SM_BIAS + COUT * 4;
  static constexpr int SMEM_BYTES = SM_BAR + 256 + 1024;
  static constexpr uint32_t idesc_n(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  }
  static constexpr uint32_t IDESC = idesc_n(COUT);
  static constexpr uint32_t IDESC2 = idesc_n(2 * COUT);
};

struct Params {
  const float* bias;
  float* dst;
  float* dst_pool;
  int c0_off, c0_chunks, c1_off, c1_chunks;
  int dst_c_total, dst_c_off, dst_layout, dst_mode;
  int pool_c_total, pool_c_off;
  int B, H, W, cout, act;
  int tiles_x, tiles_y, n_tiles;
};

// ------------------------------------------------------------------ PTX wrappers (see retrieval.cu)
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
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();  // never hang the box on a protocol bug
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
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row m, one tf32 element per 32-bit column), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major SWIZZLE_128B descriptor: 128-byte rows, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int COUT>
__global__ void __launch_bounds__(THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
               const Params p) {
  using C = Cfg<COUT>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(sm + C::SM_BIAS);
  const uint32_t bar0 = base + C::SM_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (3 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (3 * STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::SM_BAR + 8 * (3 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < COUT; i += THREADS) bias_s[i] = p.bias[i];
  if (warp == WARP_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunks = p.c0_chunks + p.c1_chunks;
  const int ksteps = 9 * chunks;  // pipeline stages per tile

  if (warp == WARP_TMA) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        const int x0 = tx * TX, y0 = ty * TY;
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t s_base = base + stage * C::STAGE_BYTES;
            mbar_expect_tx(full_bar(stage), A_BYTES + 2 * C::W_BYTES);
            if (ch < p.c0_chunks)
              tma_load_4d(s_base, &map_a0, full_bar(stage), p.c0_off + ch * KC, x0 + kx - 1, y0 + ky - 1, b);
            else
              tma_load_4d(s_base, &map_a1, full_bar(stage), p.c1_off + (ch - p.c0_chunks) * KC, x0 + kx - 1,
                          y0 + ky - 1, b);
            tma_load_3d(s_base + A_BYTES, &map_whi, full_bar(stage), ch * KC, 0, tap);
            tma_load_3d(s_base + A_BYTES + C::W_BYTES, &map_wlo, full_bar(stage), ch * KC, 0, tap);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < WARP_TMA) {
    // =========================== hi / lo split -> TMEM ===========================
    const int cw = warp & 3;            // TMEM lane quadrant this warp may write (warp % 4)
    const int grp = (warp - 4) >> 2;    // converter group: handles k-steps with (global index % CONV_GROUPS == grp)
    const int r = cw * 32 + lane;       // tile row (pixel) == TMEM lane
    int stage = 0, turn = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      for (int ks = 0; ks < ksteps; ++ks) {
        if (turn != grp) {
          if (++turn == CONV_GROUPS) turn = 0;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        if (++turn == CONV_GROUPS) turn = 0;
        mbar_wait(full_bar(stage), phase);
        // row r of the TMA box: 128 bytes, 16-byte chunk c stored at chunk (c ^ (r & 7)) (SWIZZLE_128B); the
        // XOR makes the 8 lanes of a quarter-warp hit 8 different bank groups -> conflict-free LDS.128
        const uint8_t* rowp = sm + stage * C::STAGE_BYTES + r * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
          const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t h = __float_as_uint(f[e]) & 0xFFFFE000u;
            hi[4 * c + e] = h;
            lo[4 * c + e] = __float_as_uint(f[e] - __uint_as_float(h));
          }
        }
        const uint32_t ta = tmem_base + (uint32_t)(C::ACC_COLS + stage * C::A_COLS) + ((uint32_t)(cw * 32) << 16);
        tmem_st32(ta, hi);
        tmem_st32(ta + KC, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(conv_bar(stage));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C::ACC_STAGE_COLS);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(conv_bar(stage), phase);
          tc_fence_after();
          const uint32_t s_base = base + stage * C::STAGE_BYTES;
          const uint32_t a_hi = tmem_base + (uint32_t)(C::ACC_COLS + stage * C::A_COLS), a_lo = a_hi + KC;
          const uint64_t w_hi = make_sdesc(s_base + A_BYTES), w_lo = make_sdesc(s_base + A_BYTES + C::W_BYTES);
#pragma unroll
          for (int k = 0; k < KC / 8; ++k) {
            // A: 8 tf32 = 8 TMEM columns; B: 8 tf32 = 32 bytes along K = +2 in the descriptor's (addr >> 4)
            const uint64_t o = (uint64_t)(2 * k);
            if (C::CONCAT) {
              tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, w_hi + o, C::IDESC2, (ks | k) != 0 ? 1u : 0u);  // x [W_hi;W_lo]
              tc_mma_tf32_ts(d_tmem, a_lo + 8 * k, w_hi + o, C::IDESC, 1u);
            } else {
              tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, w_hi + o, C::IDESC, (ks | k) != 0 ? 1u : 0u);
              tc_mma_tf32_ts(d_tmem, a_lo + 8 * k, w_hi + o, C::IDESC, 1u);
              tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, w_lo + o, C::IDESC, 1u);
            }
          }
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < 4) {
    // =========================== epilogue ===========================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int ly = 2 * warp + (lane >> 4), lx = lane & 15;  // pixel inside the tile (TMEM lane = ly*16 + lx)
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
      const int gx = tx * TX + lx, gy = ty * TY + ly;
      const bool valid = gx < p.W && gy < p.H;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * C::ACC_STAGE_COLS) + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < COUT / 32; ++cc) {
        float v[32];
        __syncwarp();
        tmem_ld32(taddr + (uint32_t)(cc * 32), v);
        if (C::CONCAT) {
          float v2[32];
          tmem_ld32(taddr + (uint32_t)(COUT + cc * 32), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = v[j] + bias_s[cc * 32 + j];
          if (p.act == NVS_ACT_LRELU) a = a > 0.f ? a : 0.01f * a;
          else if (p.act == NVS_ACT_RELU) a = fmaxf(a, 0.f);
          v[j] = a;
        }
        const int cbase = cc * 32;
        if (p.dst_mode == 1 && valid) {
          if (p.dst_layout == 0) {  // NHWC
            float4* d = reinterpret_cast<float4*>(
                p.dst + (((size_t)b * p.H + gy) * p.W + gx) * p.dst_c_total + p.dst_c_off + cbase);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (cbase + 4 * q < p.cout) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {  // NCHW
            float* d = p.dst + (((size_t)b * p.dst_c_total + p.dst_c_off + cbase) * p.H + gy) * p.W + gx;
            const size_t plane = (size_t)p.H * p.W;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cbase + j < p.cout) d[j * plane] = v[j];
          }
        } else if (p.dst_mode == 2 && valid) {  // PixelShuffle(2) -> NHWC (B, 2H, 2W, cout/4)
          const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float4* d = reinterpret_cast<float4*>(
                  p.dst + (((size_t)b * H2 + 2 * gy + i) * W2 + 2 * gx + j) * p.dst_c_total + p.dst_c_off + cbase / 4);
              const int o = 2 * i + j;
              d[0] = make_float4(v[o], v[4 + o], v[8 + o], v[12 + o]);
              d[1] = make_float4(v[16 + o], v[20 + o], v[24 + o], v[28 + o]);
            }
        }
        if (p.dst_pool != nullptr) {  // MaxPool2d(2,2): partners are lanes ^1 (x) and ^16 (y) of this warp
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
          }
          const int qx = gx >> 1, qy = gy >> 1;
          if ((lane & 17) == 0 && qx < (p.W >> 1) && qy < (p.H >> 1)) {
            float4* d = reinterpret_cast<float4*>(
                p.dst_pool + (((size_t)b * (p.H >> 1) + qy) * (p.W >> 1) + qx) * p.pool_c_total + p.pool_c_off + cbase);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (cbase + 4 * q < p.cout) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
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
  CUtensorMap a0, a1, whi, wlo;
  Params p;
  int cout_tpl;
  int magic;
};
constexpr int PLAN_MAGIC = 0x7C0DE5;

static int encode_act(CUtensorMap* m, const float* ptr, int B, int H, int W, int Ct) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Ct * 4, (cuuint64_t)W * Ct * 4, (cuuint64_t)H * W * Ct * 4};
  cuuint32_t box[4] = {KC, TX, TY, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}
static int encode_w(CUtensorMap* m, const float* ptr, int cin, int cout_pad) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout_pad, 9};
  cuuint64_t strides[2] = {(cuuint64_t)cin * 4, (cuuint64_t)cout_pad * cin * 4};
  cuuint32_t box[3] = {KC, (cuuint32_t)cout_pad, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}

template <int COUT>
static int launch(const Plan& pl, const Params& p, cudaStream_t st) {
  using C = Cfg<COUT>;
  static bool done = false;
  static int sms = 0;
  if (!done) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    done = true;
  }
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  conv_tc_kernel<COUT><<<grid, THREADS, C::SMEM_BYTES, st>>>(pl.a0, pl.a1, pl.whi, pl.wlo, p);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

}  // namespace tc
}  // namespace nvs

using namespace nvs;

extern "C" int32_t nvs_conv_tc_cout_pad(int32_t cout) {
  if (cout <= 0 || cout > 128) return 0;
  return cout <= 32 ? 32 : (cout <= 64 ? 64 : 128);
}

extern "C" int32_t nvs_conv_tc_supported(int32_t c0, int32_t c1, int32_t cout) {
  if (c0 <= 0 || c0 % tc::KC != 0 || c1 < 0 || c1 % tc::KC != 0) return 0;
  return nvs_conv_tc_cout_pad(cout) != 0 ? 1 : 0;
}

extern "C" size_t nvs_conv_tc_plan_bytes(void) { return sizeof(tc::Plan) + 64; }

extern "C" int nvs_conv_tc_plan_init(void* plan_mem, const NvsConvTcArgs* a) {
  if (!plan_mem || !a || !a->src0 || !a->w_hi || !a->w_lo || !a->bias) return NVS_ERR_ARG;
  if (!nvs_conv_tc_supported(a->c0, a->c1, a->cout)) return NVS_ERR_UNSUPPORTED;
  if (a->c1 > 0 && !a->src1) return NVS_ERR_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0) return NVS_ERR_ARG;
  if ((a->c0_total % 4) || (a->c0_off % 4) || (a->c1 > 0 && ((a->c1_total % 4) || (a->c1_off % 4)))) return NVS_ERR_ARG;
  if (a->dst_mode != 0 && ((a->dst_c_total % 4) || (a->dst_c_off % 4)) && a->dst_layout == 0) return NVS_ERR_ARG;
  if (a->dst_mode == 2 && (a->cout % 32) != 0) return NVS_ERR_UNSUPPORTED;
  if (a->act != NVS_ACT_NONE && a->act != NVS_ACT_LRELU && a->act != NVS_ACT_RELU) return NVS_ERR_UNSUPPORTED;
  tc::Plan* pl = reinterpret_cast<tc::Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  const int cpad = nvs_conv_tc_cout_pad(a->cout);
  const int cin = a->c0 + a->c1;
  int rc = tc::encode_act(&pl->a0, a->src0, a->B, a->H, a->W, a->c0_total);
  if (rc != NVS_OK) return rc;
  rc = tc::encode_act(&pl->a1, a->c1 > 0 ? a->src1 : a->src0, a->B, a->H, a->W, a->c1 > 0 ? a->c1_total : a->c0_total);
  if (rc != NVS_OK) return rc;
  rc = tc::encode_w(&pl->whi, a->w_hi, cin, cpad);
  if (rc != NVS_OK) return rc;
  rc = tc::encode_w(&pl->wlo, a->w_lo, cin, cpad);
  if (rc != NVS_OK) return rc;
  tc::Params& p = pl->p;
  p.bias = a->bias; p.dst = a->dst; p.dst_pool = a->dst_pool;
  p.c0_off = a->c0_off; p.c0_chunks = a->c0 / tc::KC; p.c1_off = a->c1_off; p.c1_chunks = a->c1 / tc::KC;
  p.dst_c_total = a->dst_c_total; p.dst_c_off = a->dst_c_off; p.dst_layout = a->dst_layout; p.dst_mode = a->dst_mode;
  p.pool_c_total = a->pool_c_total; p.pool_c_off = a->pool_c_off;
  p.B = a->B; p.H = a->H; p.W = a->W; p.cout = a->cout; p.act = a->act;
  p.tiles_x = (a->W + tc::TX - 1) / tc::TX; p.tiles_y = (a->H + tc::TY - 1) / tc::TY;
  p.n_tiles = p.tiles_x * p.tiles_y * a->B;
  pl->cout_tpl = cpad;
  pl->magic = tc::PLAN_MAGIC;
  return NVS_OK;
}

extern "C" int nvs_conv_tc_run(const void* plan_mem, float* dst_override, void* stream) {
  if (!plan_mem) return NVS_ERR_ARG;
  const tc::Plan* pl = reinterpret_cast<const tc::Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  if (pl->magic != tc::PLAN_MAGIC) return NVS_ERR_ARG;
  tc::Params p = pl->p;
  if (dst_override) p.dst = dst_override;
  if (p.dst_mode != 0 && !p.dst) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pl->cout_tpl) {
    case 32: return tc::launch<32>(*pl, p, st);
    case 64: return tc::launch<64>(*pl, p, st);
    case 128: return tc::launch<128>(*pl, p, st);
  }
  return NVS_ERR_UNSUPPORTED;
}
