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
// ACTIVATION FORMAT ("split" channels-last, NVS_NHWC_SPLIT16): between these layers an activation tensor (B,H,W,C) keeps
// its fp32 footprint (4 C bytes per pixel) but holds, per pixel, C fp16 values a_hi followed by C fp16 values a_lo.
// The producing layer's epilogue (or the stem / 1x1 FFMA kernels of conv.cu) performs the split once per value; the
// consuming layer's TMA loads then deliver MMA-ready operand tiles: no conversion pass, no staging buffer.
//
// Row-stationary formulation: the GEMM row is an input position, the three taps of a kernel row are one N = 3 x Cout
// operand ([kx = 0 | 1 | 2] x Cout) and out[i] = D_0[i-1] + D_1[i] + D_2[i+1] is formed by two warp shuffles per
// channel in the epilogue.  Tile = 4 image rows x 30 columns (32 positions per row with the halo columns, one row per
// TMEM lane quadrant).  The 6 x 32 halo box of a 32-channel chunk is ONE TMA box per operand part (fp16, 64-byte rows,
// SWIZZLE_64B): a contiguous K-major matrix of 192 rows, and the A operand of kernel row ky is its rows
// [32 ky, 32 ky + 128) -- a descriptor offset.  All three products go to one accumulator (3 x Cout columns per stage,
// two stages); Cout = 32 layers issue a_hi x [W_hi ; W_lo] as one N = 192 instruction (the epilogue adds the halves).
// A 16-channel tensor (the stem's output) stores [hi | lo] in one 64-byte row: one box, a_lo is the second k-step.
//
// A thread issues dependent instructions at ~8 cycles apiece, so the single-thread roles are instruction-latency bound
// (measured with the clock64 timeline, tools/rs_timeline.py: one MMA thread needed ~3100 cycles per chunk for 1700
// cycles of MMA time).  Hence: three MMA-issuing threads (one per kernel row; issuer 0 queues the tile's first,
// overwriting MMA and releases the others through the astart barrier), descriptors advanced by 32-bit immediates, ring
// positions kept incrementally (no divisions).  flags bit 0 selects one issuer: fixed accumulation order,
// bit-reproducible results.  Weights stay resident in shared memory for the whole kernel when the layer's 3 x chunks
// tiles fit (up to 64 input channels at Cout = 64), else they stream through a ring.
//
// Warps: Cout / 4 epilogue warps (quadrant = warp % 4, group of 16 channels = warp / 4), then the activation TMA warp,
// the weight TMA warp and up to three MMA-issuer warps (the first allocates TMEM).  Persistent CTAs, one per SM.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace nvs {
namespace rs {

constexpr int TX = 30, TY = 4;           // output pixels per tile
constexpr int HX = 32, HY = 6;           // halo box: positions per row (= GEMM rows per quadrant), rows
constexpr int KC = 32;                   // input channels per chunk
constexpr int A_HALF = HX * HY * KC * 2; // one of the hi / lo operand matrices: 192 rows x 64 B
constexpr int A_BYTES = 2 * A_HALF;
constexpr int NA_MAX = 6;                // activation chunks in flight: as many as fit next to the weights (2 .. 6): a chunk
                                         // load takes ~2 us to land, and a one-chunk tile has only ~0.5 us of MMA work
constexpr int ACC_STAGES = 2;
constexpr int MAX_ISSUERS = 3;
constexpr unsigned long long PLAN_MAGIC = 0x7273506C616E0002ull;  // first word of an rs::Plan

template <int CO>
struct Cfg {
  static_assert(CO == 32 || CO == 64, "output channels per launch");
  static constexpr int NW = 3 * CO;                 // GEMM N of one of W_hi / W_lo: [kx][cout]
  static constexpr int W_HALF = NW * KC * 2;        // one of W_hi / W_lo of one (chunk, ky): NW rows x 64 B
  static constexpr int W_STAGE = 2 * W_HALF;
  static constexpr bool CONCAT = CO == 32;          // a_hi x [W_hi ; W_lo] as one N = 2 NW instruction
  static constexpr int ACC_COLS = 192;              // per stage: CONCAT 2 x 96, else 192
  // epilogue warps: TMEM lane quadrant (image row of the tile) x one of four channel groups.  The epilogue is a chain
  // of long-latency steps (TMEM loads, shuffles, scattered stores): it needs warps, not instructions per warp
  static constexpr int EPI_WARPS = 16;
  // Cout = 32: TWO epilogue warp sets of 8 warps (quadrant x half of the channels), set s owns accumulator stage s =
  // every other tile.  Such tiles have too little MMA work (<= 2 chunks, ~1500 cycles) to hide one epilogue chain of
  // ~2300 cycles (TMEM loads -> shuffles -> stores; clock64 timelines in profiles/): with two tiles in flight the chains
  // overlap, and a thread owns 16 channels = one full 32-byte sector per store instead of half of one.
  static constexpr bool SETS = CO == 32;
  static constexpr int SET_WARPS = SETS ? EPI_WARPS / 2 : EPI_WARPS;  // warps that drain one accumulator stage
  static constexpr int CPW = 16;                    // output channels per epilogue warp
  static constexpr int LPW = SETS ? 8 : 16;         // channels per batch of TMEM loads (registers: 3 or 6 x LPW)
  static constexpr int WARP_TMA_A = EPI_WARPS, WARP_TMA_W = WARP_TMA_A + 1, WARP_MMA = WARP_TMA_W + 1;
  static constexpr int THREADS = 32 * (WARP_MMA + MAX_ISSUERS);
  // output staging: per TMEM lane quadrant one operand part (a_hi or a_lo, CO fp16 per position) of its 30 outputs
  static constexpr int ROW_OUT = CO * 2;
  static constexpr int QSTG_BYTES = ((TX * ROW_OUT + 127) / 128) * 128;
  static constexpr int STG_BYTES = ((4 * QSTG_BYTES + 1023) / 1024) * 1024;
  static constexpr int SM_STG = 0;
  static constexpr int SM_BIAS = SM_STG + STG_BYTES;
  static constexpr int SM_POOL = SM_BIAS + CO * 4;
  // max-pool exchange: (set, channel group, quadrant pair) x 15 x CPW; SETS: double buffered (one barrier per tile)
  static constexpr int POOL_BYTES = (SETS ? 2 : 1) * (EPI_WARPS / 4) * 2 * 15 * CPW * 4;
  static constexpr int SM_BAR = SM_POOL + POOL_BYTES;
  static constexpr int MAX_WS = 16;                  // barrier slots reserved for the weight ring
  static constexpr int N_BARS = 2 * NA_MAX + 2 * MAX_WS + 6 + 1;
  static constexpr int SM_A = ((SM_BAR + 8 * N_BARS + 16 + 1023) / 1024) * 1024;   // then `na` A slots, then the weight stages
  static constexpr int SMEM_MAX = 227 * 1024 - 1024;
  static constexpr int max_w_stages(int na) {
    return (SMEM_MAX - SM_A - na * A_BYTES) / W_STAGE < MAX_WS ? (SMEM_MAX - SM_A - na * A_BYTES) / W_STAGE : MAX_WS;
  }
  static constexpr int smem_bytes(int na, int w_stages) { return SM_A + na * A_BYTES + w_stages * W_STAGE + 1024; }
  static_assert(max_w_stages(3) >= 3, "weight ring");
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
  int c0_lo, c1_lo;  // channel coordinate of a source's a_lo block (= its total channel count); unused when pair16
  int pair16;        // source 0 is a 16-channel tensor: [hi | lo] in one 64-byte row, one box, a_lo = second k-step
  int dst_c_total, dst_c_off, dst_layout, dst_mode;
  int pool_c_total, pool_c_off;
  int B, H, W, cout, act;
  int tiles_x, tiles_y, n_tiles;
  int nk_last0, nk_last1;  // k-steps (16 channels) of the last chunk of source 0 / 1 that can be non-zero
  int w_stages;            // weight ring depth; >= 3 * chunks means resident (every tile loaded once)
  int na;                  // A slots (2 or 3)
  int store_nhwc, store_pool;  // channels-last output / pooled output (split format)
  int ps_staged;               // PixelShuffle outputs through the quadrant staging cells (Cout = 64 launches, full width)
  int staged;                  // 1: through the quadrant staging buffers (coalesced lines, two barriers per part); 0:
                               // one 32-byte store per thread and part (no barriers: better for one-chunk tiles)
  int issuers;             // MMA-issuing threads: 3, or 1 (fixed accumulation order)
  float w_scale;           // 2^-t: undoes the weights' power-of-two scale
  long long* dbg;          // nvs_conv_rs_debug_buffer: per-role clock64 stamps of CTA 0, 256 per role
  int knock;               // bottleneck experiments (env NVS_RS_KNOCK, results are then garbage): 2 epilogue only
                           // passes barriers, 4 no MMAs, 8 no activation TMA, 64 no weight TMA
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
// tcgen05.ld is asynchronous: tmem_ld16_issue starts a 16-column load, tmem_ld_wait waits for every load of the thread,
// and tmem_pin (an empty volatile asm that takes the registers as read-write operands, ordered after the wait like every
// volatile asm) keeps either compiler from moving a use of the loaded registers above the wait.  Several loads are
// issued back to back and waited for once: a load-wait pair costs a few hundred cycles of latency.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t* r) {
  static_assert(N == 8 || N == 16, "columns per load");
  if (N == 16) tmem_ld16_issue(taddr, r);
  else tmem_ld8_issue(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_pin8(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) : : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_pin(uint32_t* r) {
  tmem_pin8(r);
  if (N == 16) tmem_pin8(r + 8);
}
// one lane of the (converged) warp; the compiler then knows the guarded code is single-threaded AND that warp-uniform
// values stay uniform, so tcgen05 operands move to uniform registers without a per-lane election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (x, y) -> packed fp16 pair of the values rounded to fp16 and packed fp16 pair of the exact remainders
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  hi = h2_bits(h);
  lo = h2_bits(__floats2half2_rn(x - f.x, y - f.y));
}
// 16 consecutive channels [c, c + 16) of one pixel of a split-format tensor with c_total channels: a_hi as 32 bytes at
// (pixel * 4 c_total + 2 c), a_lo 2 c_total bytes further; only the leading n_valid channels (in steps of 8) are written
__device__ __forceinline__ void store_split16(float* basep, size_t pixel, int c_total, int c, const float* v, int n_valid) {
  uint8_t* px = reinterpret_cast<uint8_t*>(basep) + pixel * (size_t)c_total * 4;
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_pair(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
#pragma unroll
  for (int q = 0; q < 2; ++q)
    if (8 * q < n_valid) {
      *reinterpret_cast<uint4*>(px + (c + 8 * q) * 2) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
      *reinterpret_cast<uint4*>(px + (c_total + c + 8 * q) * 2) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
}

// N (16 or 8) consecutive channels [c, c + N) of one pixel of a split-format tensor, straight from registers: a_hi as ONE
// store of 2 N bytes (N = 16: a full 32-byte sector with st.global.v8.b32, new with sm_100), a_lo as another
template <int N>
__device__ __forceinline__ void store_split_direct(float* basep, size_t pixel, int c_total, int c, const float* v) {
  uint8_t* px = reinterpret_cast<uint8_t*>(basep) + pixel * (size_t)c_total * 4;
  uint32_t hi[N / 2], lo[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) split_pair(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  if (N == 16) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(px + c * 2), "r"(hi[0]), "r"(hi[1]),
                 "r"(hi[2]), "r"(hi[3]), "r"(hi[4 % (N / 2)]), "r"(hi[5 % (N / 2)]), "r"(hi[6 % (N / 2)]), "r"(hi[7 % (N / 2)])
                 : "memory");
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(px + (c_total + c) * 2), "r"(lo[0]),
                 "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4 % (N / 2)]), "r"(lo[5 % (N / 2)]), "r"(lo[6 % (N / 2)]),
                 "r"(lo[7 % (N / 2)])
                 : "memory");
  } else {
    *reinterpret_cast<uint4*>(px + c * 2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(px + (c_total + c) * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// (tx, ty, frame) of the tiles blockIdx.x, blockIdx.x + gridDim.x, ... without a division per tile: three 32-bit divisions
// by run-time values cost a thread ~300 cycles, a large part of a one-chunk tile's budget in every role
struct TileIter {
  int tx, ty, b, sx, sy, sb, nx, ny;
  __device__ __forceinline__ TileIter(int t0, int step, int tiles_x, int tiles_y) : nx(tiles_x), ny(tiles_y) {
    tx = t0 % tiles_x; ty = (t0 / tiles_x) % tiles_y; b = t0 / (tiles_x * tiles_y);
    sx = step % tiles_x; sy = (step / tiles_x) % tiles_y; sb = step / (tiles_x * tiles_y);
  }
  __device__ __forceinline__ void next() {
    tx += sx;
    int carry = 0;
    if (tx >= nx) { tx -= nx; carry = 1; }
    ty += sy + carry;
    carry = 0;
    if (ty >= ny) { ty -= ny; carry = 1; }
    b += sb + carry;
  }
};

template <int CO>
__global__ void __launch_bounds__((Cfg<CO>::THREADS), 1)
conv_rs_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
               const Params p) {
  using C = Cfg<CO>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(sm + C::SM_BIAS);
  const uint32_t bar0 = base + C::SM_BAR;
  auto afull = [&](int i) { return bar0 + 8u * i; };
  auto aempty = [&](int i) { return bar0 + 8u * (NA_MAX + i); };
  auto wfull = [&](int i) { return bar0 + 8u * (2 * NA_MAX + i); };
  auto wempty = [&](int i) { return bar0 + 8u * (2 * NA_MAX + C::MAX_WS + i); };
  auto accfull = [&](int i) { return bar0 + 8u * (2 * NA_MAX + 2 * C::MAX_WS + i); };
  auto accempty = [&](int i) { return bar0 + 8u * (2 * NA_MAX + 2 * C::MAX_WS + 2 + i); };
  auto astart = [&](int i) { return bar0 + 8u * (2 * NA_MAX + 2 * C::MAX_WS + 4 + i); };
  const int NA = p.na;
  const uint32_t sm_w = (uint32_t)(C::SM_A + NA * A_BYTES);  // first weight stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::SM_BAR + 8 * C::N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = p.c0_chunks + p.c1_chunks;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int w_tiles = 3 * chunks;                 // (chunk, ky) weight tiles of the layer
  const bool w_resident = p.w_stages >= w_tiles;  // every tile has its own stage: loaded once

  for (int i = threadIdx.x; i < CO; i += C::THREADS) bias_s[i] = p.bias[i];
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
    for (int i = 0; i < NA_MAX; ++i) {
      mbar_init(afull(i), 1);           // the activation TMA's expect_tx arrive
      mbar_init(aempty(i), p.issuers);  // tcgen05.commit of every issuer's MMAs of the chunk
    }
    for (int i = 0; i < C::MAX_WS; ++i) {
      mbar_init(wfull(i), 1);
      mbar_init(wempty(i), 1);          // the issuer that used the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(accfull(i), p.issuers);
      mbar_init(accempty(i), C::SET_WARPS);
      mbar_init(astart(i), 1);          // issuer 0 has queued the tile's first (overwriting) MMA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C::WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < C::EPI_WARPS) {
    // =========================== epilogue ===========================
    // quadrant (= image row of the tile) warp % 4, channel group h = warp / 4 (16 channels); lane = position of the 32-wide strip
    // (lane 0 / 31 are halo positions).  Accumulator columns of a stage: [kx 0 | kx 1 | kx 2] x CO channels (CONCAT:
    // then the same three blocks of a_hi w_lo + a_lo w_hi).  out[i] = D_0[i-1] + D_1[i] + D_2[i+1].
    const int quad = warp & 3;
    const int set = C::SETS ? (warp >> 2) & 1 : 0;  // accumulator stage (= tile parity) this warp drains
    const int h = C::SETS ? warp >> 3 : warp >> 2;  // group of CPW channels
    int acc = set;
    uint32_t aph = 0;
    const float neg_slope = p.act == NVS_ACT_LRELU ? 0.01f : (p.act == NVS_ACT_RELU ? 0.f : 1.f);
    int bad = 0, out_of_range = 0;
    // channels-last outputs feed the next layer's fp16 operands: flag |x| >= 60000; other outputs: flag non-finite values
    const float range_limit = (p.dst_layout == 0 && p.dst_mode != 3) ? 60000.f : 3.0e38f;
    int dbg_i = 0;
    // Channels-last outputs go through a small staging buffer per TMEM lane quadrant (= image row of the tile), one
    // operand part at a time (a_hi, then a_lo).  Written straight from the registers a warp store instruction would hit
    // 32 different 128-byte lines with 16 bytes each (thread = pixel, pixels 4 C bytes apart): the stores, not the math,
    // bounded the epilogue and with it most layers.  Instead the warps of a quadrant (one per 16 channels) write their
    // 32 bytes per position into the swizzled staging rows, meet at the quadrant's named barrier, and copy the rows out
    // with consecutive threads on consecutive 16-byte chunks: full lines.  `v`: this thread's 16 channels, `row`: its
    // position in the staged row set (n_rows positions), `px0`: pixel index of staged row 0, `x_lim`: rows >= x_lim lie
    // outside the image.
    constexpr int CPW = C::CPW;
    constexpr int QWARPS = C::EPI_WARPS / 4, QTHREADS = QWARPS * 32, CH = C::ROW_OUT / 16;
    uint8_t* stg_q = sm + C::SM_STG + quad * C::QSTG_BYTES;
    const int qtid = h * 32 + lane;
    auto staged_store = [&](float* dstp, const float* v, bool writes, int row, int n_rows, int c_off, int c_total,
                            size_t px0, int x_lim, bool row_ok) {
      uint32_t hi[CPW / 2], lo[CPW / 2];
#pragma unroll
      for (int j = 0; j < CPW / 2; ++j) split_pair(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
      auto key = [](int r) { return C::ROW_OUT == 128 ? (r & 7) : ((r >> 1) & 3); };
      uint8_t* r0 = stg_q + row * C::ROW_OUT;
      const int k0 = key(row);
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const uint32_t* w = part ? lo : hi;
        if (writes) {  // this warp's channel group = 16-byte chunk(s) CPW / 8 * h ... of the row
          if (CPW == 16) {
            *reinterpret_cast<uint4*>(r0 + (((2 * h) ^ k0) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(r0 + (((2 * h + 1) ^ k0) << 4)) =
                make_uint4(w[4 % (CPW / 2)], w[5 % (CPW / 2)], w[6 % (CPW / 2)], w[7 % (CPW / 2)]);
          } else {
            *reinterpret_cast<uint4*>(r0 + ((h ^ k0) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(10 + quad), "n"(QTHREADS) : "memory");
        if (row_ok) {
          uint8_t* gbase = reinterpret_cast<uint8_t*>(dstp) + px0 * (size_t)c_total * 4 + (size_t)(c_off + part * c_total) * 2;
          for (int idx = qtid; idx < n_rows * CH; idx += QTHREADS) {
            const int r = idx / CH, j = idx - r * CH;
            if (r < x_lim)
              *reinterpret_cast<uint4*>(gbase + (size_t)r * c_total * 4 + j * 16) =
                  *reinterpret_cast<const uint4*>(stg_q + r * C::ROW_OUT + ((j ^ key(r)) << 4));
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(10 + quad), "n"(QTHREADS) : "memory");
      }
    };
    const int t_step = (C::SETS ? 2 : 1) * (int)gridDim.x;
    TileIter ti((int)blockIdx.x + set * (int)gridDim.x, t_step, p.tiles_x, p.tiles_y);
    for (int t = (int)blockIdx.x + set * (int)gridDim.x; t < p.n_tiles; t += t_step, ti.next()) {
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && dbg_i < 256) p.dbg[dbg_i++] = clock64();
      const int tx = ti.tx, ty = ti.ty, b = ti.b;
      const int gx = tx * TX - 1 + lane, gy = ty * TY + quad;
      const bool valid = lane >= 1 && lane <= TX && gx < p.W && gy < p.H;
      mbar_wait(accfull(acc), aph);
      tc_fence_after();
      bool released = false;
      const uint32_t taddr = tmem_base + (uint32_t)(acc * C::ACC_COLS) + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int part = 0; part < ((p.knock & 2) ? 0 : 1); ++part) {
        const int cbase = CPW * h;  // first of this warp's CPW output channels
        if (p.dst_mode == 3 && cbase != 0) continue;  // keypoint heads: 3 channels in all
        // (channels beyond cout are zero weights + zero bias: a staged store writes them, as exact zeros, with the rest)
        if (cbase >= p.cout && !(p.store_nhwc | p.store_pool)) continue;
        float o[CPW];
        constexpr int LPW = C::LPW;
#pragma unroll
        for (int sb = 0; sb < CPW / LPW; ++sb) {
          const int cb = cbase + sb * LPW;
          uint32_t u0[LPW], u1[LPW], u2[LPW];
          tmem_ld_issue<LPW>(taddr + (uint32_t)cb, u0);
          tmem_ld_issue<LPW>(taddr + (uint32_t)(CO + cb), u1);
          tmem_ld_issue<LPW>(taddr + (uint32_t)(2 * CO + cb), u2);
          if (C::CONCAT) {  // ... and the same three blocks of a_hi w_lo + a_lo w_hi
            uint32_t w0[LPW], w1[LPW], w2[LPW];
            tmem_ld_issue<LPW>(taddr + (uint32_t)(C::NW + cb), w0);
            tmem_ld_issue<LPW>(taddr + (uint32_t)(C::NW + CO + cb), w1);
            tmem_ld_issue<LPW>(taddr + (uint32_t)(C::NW + 2 * CO + cb), w2);
            tmem_ld_wait();
            tmem_pin<LPW>(u0); tmem_pin<LPW>(u1); tmem_pin<LPW>(u2); tmem_pin<LPW>(w0); tmem_pin<LPW>(w1); tmem_pin<LPW>(w2);
#pragma unroll
            for (int j = 0; j < LPW; ++j) {
              const float d0 = __uint_as_float(u0[j]) + __uint_as_float(w0[j]);
              const float d1 = __uint_as_float(u1[j]) + __uint_as_float(w1[j]);
              const float d2 = __uint_as_float(u2[j]) + __uint_as_float(w2[j]);
              o[sb * LPW + j] = __shfl_up_sync(0xffffffffu, d0, 1) + d1 + __shfl_down_sync(0xffffffffu, d2, 1);
            }
          } else {
            tmem_ld_wait();
            tmem_pin<LPW>(u0); tmem_pin<LPW>(u1); tmem_pin<LPW>(u2);
#pragma unroll
            for (int j = 0; j < LPW; ++j)
              o[sb * LPW + j] = __shfl_up_sync(0xffffffffu, __uint_as_float(u0[j]), 1) + __uint_as_float(u1[j]) +
                                __shfl_down_sync(0xffffffffu, __uint_as_float(u2[j]), 1);
          }
        }
        // the accumulator is in registers: hand the stage back to the MMA role before the activation / pooling / stores.
        // Only for tiles with little MMA work (same-box A/B, call 64: conv1b -7 %, 64 -> 64 -4 %, but the three-chunk
        // 96 -> 64 layers and the staged PixelShuffle layers +3..5 % -- there the epilogue runs better undisturbed)
        if (chunks <= 2 && !p.ps_staged) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accempty(acc));
          released = true;
        }
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const float a = fmaf(o[j], p.w_scale, bias_s[cbase + j]);
          // range check BEFORE the activation (max / min would turn a NaN into 0): an activation beyond fp16's range --
          // or, in any output mode, a non-finite value (an operand of this layer was already out of range)
          out_of_range |= !(fabsf(a) < range_limit);
          o[j] = fmaxf(a, 0.f) + neg_slope * fminf(a, 0.f);
        }
        if (!valid) out_of_range = 0;  // halo positions / positions outside the image are never stored
        bad |= out_of_range;
        out_of_range = 0;
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
        if (p.knock & 128) {  // experiment: everything but the stores (the dependency keeps the math alive)
#pragma unroll
          for (int j = 0; j < CPW; ++j) bad |= o[j] == 12345.678f;
          continue;
        }
        if (p.dst_pool != nullptr) {
          // MaxPool2d(2,2): x partner = next lane (strips start at even x, so pairs are lanes (1,2), (3,4), ...),
          // y partner = the row of the next quadrant's warp: odd quadrants hand their x-pooled values to the even
          // quadrant below them through shared memory (one buffer and one named barrier per (h, quadrant pair);
          // the pair's second barrier keeps the next write behind this read)
          const int pool_grp = (C::SETS ? set * 2 + h : h) * 2 + (quad >> 1);  // 0 .. 7
          float* ps = reinterpret_cast<float*>(sm + C::SM_POOL) + (pool_grp + (C::SETS ? 8 * (int)(aph & 1u) : 0)) * (15 * CPW);
          const int bar_id = 2 + pool_grp;
          float m[CPW];
#pragma unroll
          for (int j = 0; j < CPW; ++j) m[j] = fmaxf(o[j], __shfl_down_sync(0xffffffffu, o[j], 1));
          const int pc = (lane - 1) >> 1;  // pooled column inside the strip, odd lanes 1..29 -> 0..14
          const bool owner = (lane & 1) && lane <= 29;
          if ((quad & 1) && owner) {
#pragma unroll
            for (int q = 0; q < CPW / 4; ++q)
              reinterpret_cast<float4*>(ps + pc * CPW)[q] = make_float4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          const bool pool_owner = !(quad & 1) && owner;
          if (pool_owner) {
#pragma unroll
            for (int q = 0; q < CPW / 4; ++q) {
              const float4 r = reinterpret_cast<const float4*>(ps + pc * CPW)[q];
              m[4 * q] = fmaxf(m[4 * q], r.x);
              m[4 * q + 1] = fmaxf(m[4 * q + 1], r.y);
              m[4 * q + 2] = fmaxf(m[4 * q + 2], r.z);
              m[4 * q + 3] = fmaxf(m[4 * q + 3], r.w);
            }
          }
          // pooled row (15 positions) of the even quadrants; the odd ones have nothing to store
          if (p.staged == 0) {
            const int qx = gx >> 1, qy = gy >> 1;
            if (pool_owner && qx < (p.W >> 1) && qy < (p.H >> 1)) {
              const size_t ppx = ((size_t)b * (p.H >> 1) + qy) * (p.W >> 1) + qx;
              if (p.pool_c_off & 15) store_split16(p.dst_pool, ppx, p.pool_c_total, p.pool_c_off + cbase, m, CPW);  // 16-byte stores
              else store_split_direct<CPW>(p.dst_pool, ppx, p.pool_c_total, p.pool_c_off + cbase, m);
            }
          } else if (!(quad & 1)) {
            const int Hp = p.H >> 1, Wp = p.W >> 1, qy = gy >> 1, qx0 = tx * (TX / 2);
            staged_store(p.dst_pool, m, owner, pc, TX / 2, p.pool_c_off, p.pool_c_total,
                         ((size_t)b * Hp + qy) * Wp + qx0, Wp - qx0, qy < Hp);
          }
          // (SETS: the exchange buffer alternates per tile; the next tile's first barrier orders its reuse)
          if (!C::SETS) asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        }
        if (valid && p.dst_mode == 1) {
          if (p.dst_layout == 0) {  // channels-last, split format: through the staging buffer (below)
          } else {  // NCHW: a warp writes 30 consecutive x of one channel row
            float* d = p.dst + (((size_t)b * p.dst_c_total + p.dst_c_off + cbase) * p.H + gy) * p.W + gx;
            const size_t plane = (size_t)p.H * p.W;
#pragma unroll
            for (int j = 0; j < CPW; ++j)
              if (cbase + j < p.cout) d[j * plane] = o[j];
          }
        }
        if (p.store_nhwc && p.staged == 0) {
          if (valid) {
            const size_t opx = ((size_t)b * p.H + gy) * p.W + gx;
            if (p.dst_c_off & 15) store_split16(p.dst, opx, p.dst_c_total, p.dst_c_off + cbase, o, CPW);  // 16-byte stores
            else store_split_direct<CPW>(p.dst, opx, p.dst_c_total, p.dst_c_off + cbase, o);
          }
        } else if (p.store_nhwc)  // this quadrant's image row: 30 positions
          staged_store(p.dst, o, lane >= 1 && lane <= TX, lane - 1, TX, p.dst_c_off, p.dst_c_total,
                       ((size_t)b * p.H + gy) * p.W + tx * TX, p.W - tx * TX, gy < p.H);
        if (p.dst_mode == 2 && p.ps_staged && CPW == 16) {
          // PixelShuffle(2), staged: thread = input pixel holds 4 sub-pixels x 4 shuffled channels of its warp's 16
          // conv channels -- written directly that is 8 bytes to each of 4 output pixels 256 bytes apart (four warps
          // per 32-byte sector).  Instead the four warps of the quadrant fill 8-byte cells [h][i * 60 + xo] (xo = output
          // x inside the tile's 60-pixel output row i; conflict-free 16-byte stores of the (j = 0, j = 1) cell pair) and,
          // after the quadrant's barrier, lane pairs write each output pixel's 32 bytes (one full sector) per part.
          const int H2 = 2 * p.H, W2 = 2 * p.W;
          const bool wr = lane >= 1 && lane <= TX;
          uint32_t hi4[4][2], lo4[4][2];
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            split_pair(o[s4], o[4 + s4], hi4[s4][0], lo4[s4][0]);
            split_pair(o[(8 + s4) % CPW], o[(12 + s4) % CPW], hi4[s4][1], lo4[s4][1]);
          }
          const size_t opx0 = ((size_t)b * H2 + 2 * gy) * W2 + 2 * tx * TX;  // output pixel of (i = 0, xo = 0)
          const int xo_lim = W2 - 2 * tx * TX;
          const bool row_ok = gy < p.H;
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            if (wr) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint32_t(*w4)[2] = part ? lo4 : hi4;
                *reinterpret_cast<uint4*>(stg_q + (size_t)(h * 120 + i * 60 + 2 * (lane - 1)) * 8) =
                    make_uint4(w4[2 * i][0], w4[2 * i][1], w4[2 * i + 1][0], w4[2 * i + 1][1]);
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(10 + quad), "n"(QTHREADS) : "memory");
            if (row_ok) {
              uint8_t* gb = reinterpret_cast<uint8_t*>(p.dst) + (size_t)(p.dst_c_off + part * p.dst_c_total) * 2;
              for (int idx = qtid; idx < 240; idx += QTHREADS) {
                const int r = idx >> 1, jc = idx & 1, i = r >= 60 ? 1 : 0, xo = r - 60 * i;
                if (xo < xo_lim) {
                  const uint2 c0 = *reinterpret_cast<const uint2*>(stg_q + (size_t)((2 * jc) * 120 + r) * 8);
                  const uint2 c1 = *reinterpret_cast<const uint2*>(stg_q + (size_t)((2 * jc + 1) * 120 + r) * 8);
                  *reinterpret_cast<uint4*>(gb + (opx0 + (size_t)i * W2 + xo) * (size_t)p.dst_c_total * 4 + jc * 16) =
                      make_uint4(c0.x, c0.y, c1.x, c1.y);
                }
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(10 + quad), "n"(QTHREADS) : "memory");
          }
        } else if (valid && p.dst_mode == 2) {
          // PixelShuffle(2) -> channels-last (B, 2H, 2W, cout/4), split format: channel c -> (c%4/2, c%2, c/4)
          const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int s = 2 * i + j;
              uint8_t* px = reinterpret_cast<uint8_t*>(p.dst) +
                            (((size_t)b * H2 + 2 * gy + i) * W2 + 2 * gx + j) * (size_t)p.dst_c_total * 4;
              const int c = p.dst_c_off + cbase / 4;
              if (CPW == 16) {
                uint32_t hi2[2], lo2[2];
                split_pair(o[s], o[4 + s], hi2[0], lo2[0]);
                split_pair(o[(8 + s) % CPW], o[(12 + s) % CPW], hi2[1], lo2[1]);
                *reinterpret_cast<uint2*>(px + c * 2) = make_uint2(hi2[0], hi2[1]);
                *reinterpret_cast<uint2*>(px + (p.dst_c_total + c) * 2) = make_uint2(lo2[0], lo2[1]);
              } else {
                uint32_t hi1, lo1;
                split_pair(o[s], o[4 + s], hi1, lo1);
                *reinterpret_cast<uint32_t*>(px + c * 2) = hi1;
                *reinterpret_cast<uint32_t*>(px + (p.dst_c_total + c) * 2) = lo1;
              }
            }
        }
      }
      if (!released) {  // (warps with no channels to load in this launch)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty(acc));
      }
      if (C::SETS) {
        aph ^= 1;  // my stage again two tiles later
      } else if (++acc == ACC_STAGES) {
        acc = 0;
        aph ^= 1;
      }
    }
    // an activation beyond the fp16 range makes the next layer's a_hi infinite: report it (sticky flag)
    if (p.range_flag != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.range_flag, 1);
  } else if (warp == C::WARP_TMA_A) {
    // =========================== activation producer: the hi and the lo box of one chunk per A slot ===============
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      const uint32_t bytes = p.pair16 ? A_HALF : A_BYTES;
      TileIter ti((int)blockIdx.x, (int)gridDim.x, p.tiles_x, p.tiles_y);
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ti.next()) {
        const int b = ti.b;
        const int x0 = ti.tx * TX - 1, y0 = ti.ty * TY - 1;
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(aempty(as), aph ^ 1u);  // (spinning: this thread's latency bounds one-chunk tiles)
          const uint32_t dst = base + C::SM_A + as * A_BYTES;
          if (p.knock & 8) {
            mbar_arrive(afull(as));
          } else {
            mbar_expect_tx(afull(as), bytes);
            if (ch < p.c0_chunks) {
              const int c = p.c0_off + ch * KC;
              tma_load_4d(dst, &map_a0, afull(as), c, x0, y0, b);
              if (!p.pair16) tma_load_4d(dst + A_HALF, &map_a0, afull(as), p.c0_lo + c, x0, y0, b);
            } else {
              const int c = p.c1_off + (ch - p.c0_chunks) * KC;
              tma_load_4d(dst, &map_a1, afull(as), c, x0, y0, b);
              tma_load_4d(dst + A_HALF, &map_a1, afull(as), p.c1_lo + c, x0, y0, b);
            }
          }
          if (++as == NA) {
            as = 0;
            aph ^= 1;
          }
        }
      }
    }
  } else if (warp == C::WARP_TMA_W) {
    // =========================== weight producer: [W_hi ; W_lo] of one (chunk, ky) per stage ===========================
    if (lane == 0 && my_tiles > 0 && !(p.knock & 64)) {
      const int rounds = w_resident ? 1 : my_tiles;
      int ws = 0;
      uint32_t wph = 0;
      for (int r = 0; r < rounds; ++r)
        for (int ch = 0; ch < chunks; ++ch)
          for (int ky = 0; ky < 3; ++ky) {
            if (!w_resident) mbar_wait_relaxed(wempty(ws), wph ^ 1u);
            const uint32_t dst = base + sm_w + ws * C::W_STAGE;
            mbar_expect_tx(wfull(ws), C::W_STAGE);
            tma_load_3d(dst, &map_whi, wfull(ws), ch * KC, 0, ky);
            tma_load_3d(dst + C::W_HALF, &map_wlo, wfull(ws), ch * KC, 0, ky);
            if (++ws == p.w_stages) {
              ws = 0;
              wph ^= 1;
            }
          }
    }
  } else if (my_tiles > 0 && warp - C::WARP_MMA < p.issuers) {
    // =========================== MMA issuers ===========================
    // Issuer `me` of `nis` takes the kernel rows ky = me, me + nis, ... of every chunk.  Descriptors: the upper word is
    // the same constant for every operand (64-byte rows, SWIZZLE_64B, 8-row groups 512 B apart); the lower word is
    // (address >> 4) | LBO and is advanced by 32-bit adds of compile-time constants.  The whole warp runs the loop
    // (converged); one elected lane executes the tcgen05 / mbarrier-arrive instructions.
    const int me = __shfl_sync(0xffffffffu, warp, 0) - C::WARP_MMA, nis = p.issuers;
    constexpr uint32_t DESC_HI = (uint32_t)(512u >> 4) | (1u << 14) | (4u << 29);
    auto desc = [](uint32_t lo32) { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo32; };
    const uint32_t a_lo0 = (((base + C::SM_A) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_lo0 = (((base + sm_w) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_part = p.pair16 ? 2u : (uint32_t)(A_HALF >> 4);  // a_lo relative to a_hi
    int dbg_i = 0;
    int ws = me, as = 0, acc = 0;  // weight stage of my next kernel row, A slot, accumulator stage
    uint32_t wph = 0, aph = 0, cph = 0;
    while (ws >= p.w_stages) {  // (w_stages >= 3 > me: never taken; keeps the invariant explicit)
      ws -= p.w_stages;
      wph ^= 1;
    }
    // Fast path (the default configuration: one issuer, weights resident): the barrier / election skeleton of the general
    // loop below costs ~920 cycles per one-chunk tile with NO work in it (knock-outs 2 | 4 | 8, clock64 timeline), more
    // than the tile's MMAs (~400).  Here a chunk is ONE elected region: all three kernel rows' MMAs and the commits
    // back to back, no per-row election / __syncwarp, no astart hand-off, the weight barriers waited for once.
    const bool fast = nis == 1 && w_resident && !(p.knock & 4) && !(p.knock & 256);
    if (fast) {
      for (int i = 0; i < w_tiles; ++i) mbar_wait(wfull(i), 0u);
      for (int tl = 0; tl < my_tiles; ++tl) {
        mbar_wait(accempty(acc), cph ^ 1u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C::ACC_COLS);
        for (int ch = 0; ch < chunks; ++ch) {
          if (p.dbg && blockIdx.x == 0 && lane == 0 && dbg_i < 256) p.dbg[512 + dbg_i] = clock64();
          ++dbg_i;
          mbar_wait(afull(as), aph);
          tc_fence_after();
          const int nk = p.pair16 ? 1 : (ch == p.c0_chunks - 1 ? p.nk_last0 : (ch == chunks - 1 ? p.nk_last1 : 2));
          const uint32_t a_slot = a_lo0 + (uint32_t)as * (uint32_t)(A_BYTES >> 4);
          const uint32_t w_ch = w_lo0 + (uint32_t)(3 * ch) * (uint32_t)(C::W_STAGE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t a_hi = a_slot + (uint32_t)ky * (uint32_t)((HX * 64) >> 4), a_lo = a_hi + lo_part;
              const uint32_t w_hi = w_ch + (uint32_t)ky * (uint32_t)(C::W_STAGE >> 4), w_lo = w_hi + (uint32_t)(C::W_HALF >> 4);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                if (k < nk) {
                  const uint32_t o = 2u * k;  // 16 fp16 = 32 bytes along K = +2 in the (address >> 4) field
                  const uint32_t accum = (ch | ky | k) != 0 ? 1u : 0u;
                  if (C::CONCAT) {
                    tc_mma_f16(d_tmem, desc(a_hi + o), desc(w_hi + o), C::idesc(2 * C::NW), accum);  // x [W_hi ; W_lo]
                    tc_mma_f16(d_tmem + (uint32_t)C::NW, desc(a_lo + o), desc(w_hi + o), C::idesc(C::NW), 1u);
                  } else {
                    tc_mma_f16(d_tmem, desc(a_hi + o), desc(w_hi + o), C::idesc(C::NW), accum);
                    tc_mma_f16(d_tmem, desc(a_hi + o), desc(w_lo + o), C::idesc(C::NW), 1u);
                    tc_mma_f16(d_tmem, desc(a_lo + o), desc(w_hi + o), C::idesc(C::NW), 1u);
                  }
                }
              }
            }
            tc_commit(aempty(as));
            if (ch == chunks - 1) tc_commit(accfull(acc));
          }
          __syncwarp();
          if (++as == NA) {
            as = 0;
            aph ^= 1;
          }
        }
        if (++acc == ACC_STAGES) {
          acc = 0;
          cph ^= 1;
        }
      }
    }
    for (int tl = fast ? my_tiles : 0; tl < my_tiles; ++tl) {
      mbar_wait(accempty(acc), cph ^ 1u);
      if (me != 0) mbar_wait(astart(acc), cph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C::ACC_COLS);
      for (int ch = 0; ch < chunks; ++ch) {
        if (p.dbg && blockIdx.x == 0 && me == 0 && lane == 0 && dbg_i < 256) p.dbg[512 + dbg_i] = clock64();
        ++dbg_i;
        mbar_wait(afull(as), aph);
        tc_fence_after();
        const int nk = p.pair16 ? 1 : (ch == p.c0_chunks - 1 ? p.nk_last0 : (ch == chunks - 1 ? p.nk_last1 : 2));
        const uint32_t a_slot = a_lo0 + (uint32_t)as * (uint32_t)(A_BYTES >> 4);
#pragma unroll 1
        for (int ky = me; ky < 3; ky += nis) {
          if (!w_resident || tl == 0) mbar_wait(wfull(ws), wph);
          const uint32_t a_hi = a_slot + (uint32_t)ky * (uint32_t)((HX * 64) >> 4), a_lo = a_hi + lo_part;
          const uint32_t w_hi = w_lo0 + (uint32_t)ws * (uint32_t)(C::W_STAGE >> 4), w_lo = w_hi + (uint32_t)(C::W_HALF >> 4);
          if (!(p.knock & 4) && elect_one()) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (k < nk) {
                const uint32_t o = 2u * k;  // 16 fp16 = 32 bytes along K = +2 in the (address >> 4) field
                const uint32_t accum = (ch | ky | k) != 0 ? 1u : 0u;
                if (C::CONCAT) {
                  tc_mma_f16(d_tmem, desc(a_hi + o), desc(w_hi + o), C::idesc(2 * C::NW), accum);  // x [W_hi ; W_lo]
                  if (accum == 0u) mbar_arrive(astart(acc));
                  tc_mma_f16(d_tmem + (uint32_t)C::NW, desc(a_lo + o), desc(w_hi + o), C::idesc(C::NW), 1u);
                } else {
                  tc_mma_f16(d_tmem, desc(a_hi + o), desc(w_hi + o), C::idesc(C::NW), accum);
                  if (accum == 0u) mbar_arrive(astart(acc));
                  tc_mma_f16(d_tmem, desc(a_hi + o), desc(w_lo + o), C::idesc(C::NW), 1u);
                  tc_mma_f16(d_tmem, desc(a_lo + o), desc(w_hi + o), C::idesc(C::NW), 1u);
                }
              }
            }
          } else if ((p.knock & 4) && ch == 0 && ky == 0) {
            if (elect_one()) mbar_arrive(astart(acc));
          }
          __syncwarp();
          if (!w_resident && elect_one()) tc_commit(wempty(ws));
          ws += nis;
          if (ws >= p.w_stages) {
            ws -= p.w_stages;
            wph ^= 1;
          }
        }
        if (elect_one()) tc_commit(aempty(as));
        if (++as == NA) {
          as = 0;
          aph ^= 1;
        }
      }
      if (elect_one()) tc_commit(accfull(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        cph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::WARP_MMA) {
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

// split-format activations (B,H,W,Ct): per pixel Ct fp16 a_hi then Ct fp16 a_lo (4 Ct bytes).  As an fp16 tensor with
// 2 Ct "channels": box = (32 channels, 32 positions, 6 rows, 1 frame) = one operand part of a chunk's halo box;
// positions outside the image are zero filled by TMA = the convolution's zero padding
static int encode_act(CUtensorMap* m, const float* ptr, int B, int H, int W, int Ct) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)2 * Ct, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Ct * 4, (cuuint64_t)W * Ct * 4, (cuuint64_t)H * W * Ct * 4};
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)HX, (cuuint32_t)HY, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<float*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
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
  // weights resident when all 3 x chunks tiles fit next to >= 2 A slots (up to 64 -> 64 channels), else a 4-stage ring;
  // then as many A slots as fit (<= NA_MAX): the activation loads are what one- and two-chunk tiles wait for
  const int w_tiles = 3 * (p.c0_chunks + p.c1_chunks);
  q.w_stages = w_tiles <= C::max_w_stages(2) ? w_tiles : (C::max_w_stages(3) < 4 ? C::max_w_stages(3) : 4);
  q.na = (C::SMEM_MAX - C::SM_A - q.w_stages * C::W_STAGE) / A_BYTES;
  q.na = q.na > NA_MAX ? NA_MAX : q.na;
  {
    // staged stores pay off when a tile has enough MMA work to hide the quadrant barriers (measured: 96 -> 64 channels
    // 1.62 -> 1.46 ms, but 32 -> 32 0.50 -> 0.60 ms); NVS_RS_STORE = 1 / 2 forces staged / direct
    static int mode = -1;
    if (mode < 0) {
      const char* e = getenv("NVS_RS_STORE");
      mode = e ? atoi(e) : 0;
    }
    q.staged = mode == 1 ? 1 : (mode == 2 ? 0 : (p.c0_chunks + p.c1_chunks >= 3 ? 1 : 0));
    if (C::SETS) q.staged = 0;  // two epilogue sets: full-sector stores straight from the registers, no quadrant barriers
  }
  {
    // PixelShuffle outputs: staged full-sector stores (NVS_RS_PS=0: the direct 8-byte stores) when the launch writes
    // all of its 64 channels and the destination channel offsets keep 16-byte alignment
    static int ps = -1;
    if (ps < 0) {
      const char* e = getenv("NVS_RS_PS");
      ps = e ? atoi(e) : 1;
    }
    q.ps_staged = (ps && CO == 64 && p.dst_mode == 2 && p.cout == CO && p.dst_c_off % 8 == 0 && p.dst_c_total % 4 == 0 &&
                   C::QSTG_BYTES >= 4 * 120 * 8) ? 1 : 0;
  }
  {
    const char* e = getenv("NVS_RS_KNOCK");
    q.knock = e ? atoi(e) : 0;
    q.dbg = g_dbg;
  }
  kern<<<grid, C::THREADS, C::smem_bytes(q.na, q.w_stages), st>>>(pl.a0, pl.a1, pl.whi, pl.wlo, q);
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
  const int c0 = a->c0 == 16 ? 32 : a->c0;  // a 16-channel source ([hi | lo] in one 64-byte row) is one chunk
  if (a->c0 == 16 && (a->c1 != 0 || a->c0_off != 0 || a->c0_total != 16)) return NVS_ERR_ARG;
  // split-format outputs: 16-byte stores of 8 fp16 channels (pixel shuffle: 8-byte stores of 4)
  if (a->dst_mode == 1 && a->dst_layout == 0 && ((a->dst_c_total % 8) || (a->dst_c_off % 8))) return NVS_ERR_ARG;
  if (a->dst_pool && a->dst_mode != 3 && ((a->pool_c_total % 8) || (a->pool_c_off % 8))) return NVS_ERR_ARG;
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
  p.store_nhwc = (a->dst_mode == 1 && a->dst_layout == 0) ? 1 : 0;
  p.store_pool = (a->dst_pool != nullptr && a->dst_mode != 3) ? 1 : 0;
  // a staged store writes all cpad channels of the launch (padding channels as zeros)
  if (p.store_nhwc && a->dst_c_off + cpad > a->dst_c_total) return NVS_ERR_ARG;
  if (p.store_pool && a->pool_c_off + cpad > a->pool_c_total) return NVS_ERR_ARG;
  p.range_flag = range_flag_ptr();
  p.c0_off = a->c0_off; p.c0_chunks = c0 / KC; p.c1_off = a->c1_off; p.c1_chunks = a->c1 / KC;
  p.c0_lo = a->c0_total; p.c1_lo = a->c1_total; p.pair16 = a->c0 == 16 ? 1 : 0;
  {
    static int default_issuers = 0;
    if (!default_issuers) {
      const char* e = getenv("NVS_RS_ISSUERS");
      default_issuers = (e && atoi(e) == 3) ? MAX_ISSUERS : 1;
    }
    p.issuers = (a->flags & 1) ? 1 : default_issuers;
  }
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
  p.na = NA_MAX;
  p.staged = 0;
  p.ps_staged = 0;
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
namespace nvs {
namespace rs {
// thread = 8 channels of one pixel
__global__ void split16_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_pixels, int C, int fwd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = C / 8;
  if (i >= n_pixels * groups) return;
  const long long px = i / groups;
  const int g = (int)(i - px * groups);
  if (fwd) {
    const float4 a = reinterpret_cast<const float4*>(in + px * C + 8 * g)[0], b = reinterpret_cast<const float4*>(in + px * C + 8 * g)[1];
    uint32_t hi[4], lo[4];
    split_pair(a.x, a.y, hi[0], lo[0]);
    split_pair(a.z, a.w, hi[1], lo[1]);
    split_pair(b.x, b.y, hi[2], lo[2]);
    split_pair(b.z, b.w, hi[3], lo[3]);
    uint8_t* o = reinterpret_cast<uint8_t*>(out) + px * (long long)C * 4;
    *reinterpret_cast<uint4*>(o + 16 * g) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + 2 * C + 16 * g) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  } else {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(in) + px * (long long)C * 4;
    const uint4 h = *reinterpret_cast<const uint4*>(s + 16 * g), l = *reinterpret_cast<const uint4*>(s + 2 * C + 16 * g);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
      const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
      v[2 * j] = fh.x + fl.x;
      v[2 * j + 1] = fh.y + fl.y;
    }
    reinterpret_cast<float4*>(out + px * C + 8 * g)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(out + px * C + 8 * g)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}
static int split_launch(const float* in, float* out, long long n_pixels, int C, int fwd, cudaStream_t st) {
  if (!in || !out || n_pixels < 0 || C <= 0 || (C % 8)) return NVS_ERR_ARG;
  if (n_pixels == 0) return NVS_OK;
  if (in == out) return NVS_ERR_ARG;  // a pixel's groups are read and written by different threads
  const long long n = n_pixels * (C / 8);
  split16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n_pixels, C, fwd);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
}  // namespace rs
}  // namespace nvs

extern "C" int nvs_split16(const float* in, float* out, int64_t n_pixels, int32_t C, void* stream) {
  return nvs::rs::split_launch(in, out, n_pixels, C, 1, static_cast<cudaStream_t>(stream));
}
extern "C" int nvs_unsplit16(const float* in, float* out, int64_t n_pixels, int32_t C, void* stream) {
  return nvs::rs::split_launch(in, out, n_pixels, C, 0, static_cast<cudaStream_t>(stream));
}

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
