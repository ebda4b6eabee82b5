// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores with fp32-grade accuracy
// ("3xTF32": a = a_hi + a_lo, w = w_hi + w_lo, D += a_hi w_hi + a_lo w_hi + a_hi w_lo, fp32 accumulation
// in TMEM; the dropped a_lo w_lo term and the tf32 rounding of the lo parts are ~2^-20 relative, the same
// order as fp32 FFMA accumulation error over K = 9*Cin terms).
//
//   GEMM view:  M = 128 output pixels (8 rows x 16 cols of one frame),  N = Cout,  K = 9 taps x Cin
//   activations: NHWC.  Per tile and KC-channel chunk ONE 4-D TMA box (KC ch, 18 x, 10 y, 1 frame) -- the tile
//               plus its 1-pixel halo -- lands in shared memory; out-of-image coordinates are zero filled by
//               TMA, which *is* the convolution's zero padding.  The nine taps are nine shifted windows of that
//               one box: every activation byte crosses L2 once per layer, not nine times.
//   split:      converter warps (thread = output pixel) read the window row of the current tap from the halo
//               box (un-swizzling the 16-byte chunks), form hi = rn_tf32(a) and lo = rn_tf32(a - hi)
//               and write both straight into TENSOR MEMORY with tcgen05.st (lane = pixel, column = channel):
//               the A operand never returns to shared memory (an earlier version that staged hi/lo in smem
//               was shared-memory-bandwidth bound).
//   weights:    packed [tap][Cout][Cin] (hi and lo parts), 3-D TMA box (KC, Cout, 1) per tap into the slot ring
//               (4 or 8 slots = weight stage + TMEM A slot), canonical K-major swizzled layout for the UMMA
//               descriptor.
//   MMA:        up to four threads (one per issuer warp, steps round-robin) issue tcgen05.mma.cta_group::1.kind::tf32
//               (M=128, K=8) with A from TMEM and B from shared memory.  From one thread an instruction costs
//               max(N/2, ~50) cycles (profiles/microbench/mma_rate.cu); in the running kernel ~86 cycles each
//               whatever N <= 128 is (the measured sustained tensor rate of this pool), so
//               for Cout <= 64 the two products sharing a_hi are ONE instruction with N = 2*Cout against the
//               contiguous [W_hi ; W_lo] tile; columns [0,Cout) collect a_hi w_hi + a_lo w_hi, columns
//               [Cout,2Cout) collect a_hi w_lo and the epilogue adds the halves.
//   epilogue:   4 warps read TMEM (thread = pixel, columns = channels), add bias, activation, then store
//               NHWC / NCHW, 2x2 max-pool via warp shuffles (a warp owns 2 image rows x 16 cols) or
//               pixel-shuffled NHWC.
//
// Warp roles: 0-3 epilogue (TMEM lane quadrants), CONV_GROUPS x 4 converter warps (groups take pipeline steps
// round-robin), halo producer warp, weight producer warp, MAX_ISSUERS MMA-issuer warps (the first allocates TMEM).
// Persistent CTAs, static round-robin over (frame, tile_y, tile_x).  Four mbarrier rings: halo boxes
// (producer <-> converters), weight tiles (producer <-> MMA), TMEM A slots (converters <-> MMA), accumulators
// (MMA <-> epilogue).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace nvs {
namespace tc {

constexpr int TM = 128;            // GEMM rows (TMEM lanes) per tile
constexpr int CONV_GROUPS = 2;     // converter warp groups (4 warps each); group g takes the steps with index % CONV_GROUPS == g
constexpr int WARP_HALO = 4 + 4 * CONV_GROUPS;
constexpr int WARP_W = WARP_HALO + 1;
constexpr int MAX_ISSUERS = 4;
constexpr int ACC_STAGES = 2;      // accumulator stages in TMEM: the epilogue of a tile overlaps the next tile's MMAs
                                   // (1 stage frees columns for 6 slots instead of 4: measured no faster)
constexpr int WARP_MMA = WARP_W + 1;      // first of MAX_ISSUERS MMA-issuing warps (they take pipeline steps round-robin)
constexpr int THREADS = 32 * (WARP_MMA + MAX_ISSUERS);
constexpr int WARP_EPI2 = WARP_MMA + MAX_ISSUERS;  // ROW3 only: four more epilogue warps (the second 16 channels)
constexpr int THREADS_ROW3 = THREADS + 128;

// PAIR: 16-channel inputs (the stem's second conv).  One pipeline step then carries TWO taps: the K = 32 operand
// row is [tap 2t channels 0-15 | tap 2t+1 channels 0-15] (weights packed that way by the host, zero for the
// missing tenth tap), so a tile takes 5 steps instead of 9 -- the per-step cost is fixed overhead, not MMA time.
// ROW3 (output channels <= 32): "row-stationary" formulation.  The three taps of one kernel ROW share their A operand:
// GEMM row = input position, N = [kx = 0 | kx = 1 | kx = 2] x Cout, so a pipeline step is one kernel row (3 steps per
// tile and chunk instead of 9) and its MMAs are N = 192 ([W_hi ; W_lo] of three taps) and N = 96 wide -- wide enough
// to cost their nominal time instead of the ~86-cycle floor a narrow MMA pays.  The tile is 4 image rows x 32
// consecutive positions (one row per TMEM lane quadrant = per warp); D[position j][kx block] is w[ky,kx] . in[j] summed
// over ky and channels, and the output at position i is D_0[i-1] + D_1[i] + D_2[i+1]: two warp shuffles per channel
// in the epilogue, lanes 0 and 31 are halo positions (30 valid outputs per strip of 32).
template <int COUT, int KC, bool PAIR = false, bool ROW3 = false>
struct Cfg {
  static constexpr int TX = ROW3 ? 30 : 16, TY = ROW3 ? 4 : 8;             // output pixels per tile
  static constexpr int HX = ROW3 ? 32 : TX + 2, HY = TY + 2;               // halo box (ROW3: its 32 columns ARE the GEMM rows)
  static constexpr int HC = PAIR ? 16 : KC;                                // channels per halo-box row
  static constexpr int TAPS = ROW3 ? 3 : (PAIR ? 5 : 9);                   // pipeline steps per (tile, chunk)
  static constexpr int ROW_BYTES = KC * 4;                                 // weight / A row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  static constexpr int HALO_ROW_BYTES = HC * 4;
  static constexpr int HALO_BYTES = ((HX * HY * HALO_ROW_BYTES + 1023) / 1024) * 1024;
  static_assert(!PAIR || KC == 32, "paired taps fill a 32-wide K row");
  static_assert(!ROW3 || (COUT == 32 && !PAIR), "row-stationary variant: 32 output channels, 32- or 16-channel chunks");
  static constexpr int NW = ROW3 ? 3 * COUT : COUT;                        // GEMM N of one of W_hi / W_lo
  static constexpr int W_BYTES = NW * ROW_BYTES;                           // one of W_hi / W_lo
  static constexpr int W_STAGE = 2 * W_BYTES;
  static constexpr bool CONCAT = COUT <= 64;
  static constexpr int ACC_STAGE_COLS = CONCAT ? 2 * NW : NW;
  static constexpr int ACC_COLS = ACC_STAGES * ACC_STAGE_COLS;
  static constexpr int A_COLS = 2 * KC;                                    // TMEM slot: KC columns hi + KC columns lo
  // One ring of NS "slots": slot s = weight stage s in shared memory + A slot s in tensor memory, guarded by ONE
  // full barrier (4 converter warps + the weight TMA) and ONE empty barrier (tcgen05.commit).  The MMA-issuing
  // thread is the critical resource (it stalls while the tensor pipe is busy and the pipe drains while it polls
  // barriers), so it must do exactly one wait and one commit per pipeline step.
  // NS is a multiple of MAX_ISSUERS: slot s is then always consumed by issuer s % issuers, which keeps every
  // parity wait at most one phase behind (with 3 issuers on 4 slots an issuer lapped a slow converter warp).
  static constexpr int NS_TMEM = (512 - ACC_COLS) / A_COLS;
  static constexpr int NS_SMEM = (227 * 1024 - 8192 - 3 * HALO_BYTES) / W_STAGE;
  static constexpr int NS_FIT = NS_TMEM < NS_SMEM ? NS_TMEM : NS_SMEM;
  static constexpr int NS = NS_FIT >= 8 ? 8 : (NS_FIT >= 6 ? 6 : (NS_FIT >= 4 ? 4 : 2));
  static constexpr int NH = 3;                                             // halo boxes in flight
  static constexpr int KSTEPS = KC / 8;                                    // MMAs (K = 8 tf32) per operand pair
  // shared memory: halo boxes | bias, pool exchange, mbarriers, TMEM slot | weight stages (last: their number is a
  // launch parameter -- the NS-deep ring, or EVERY (chunk, tap) tile of the layer when the weights stay resident)
  static constexpr int SM_BIAS = NH * HALO_BYTES;
  static constexpr int POOL_BYTES = ROW3 ? 2 * 2 * 15 * 16 * 4 : 0;        // ROW3 max-pool: rows of a pair live in two warps
  static constexpr int SM_POOL = SM_BIAS + COUT * 4;
  static constexpr int SM_BAR = SM_POOL + ((POOL_BYTES + 7) / 8) * 8;
  static constexpr int N_BARS = 2 * NH + 2 * NS + 7;
  static constexpr int SM_W = ((SM_BAR + 8 * N_BARS + 16 + 1023) / 1024) * 1024;
  static constexpr int smem_bytes(int w_stages) { return SM_W + w_stages * W_STAGE + 1024; }
  static constexpr int SMEM_BYTES = smem_bytes(NS);
  // weight tiles that fit next to the halo boxes when they stay resident for the whole kernel
  static constexpr int MAX_W_RESIDENT = (227 * 1024 - 1024 - SM_W) / W_STAGE;
  static_assert(NS <= NS_FIT && NS % CONV_GROUPS == 0, "TMEM budget / converter ring");
  static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");
  static_assert(W_BYTES % 1024 == 0 || (KC == 16 && W_BYTES % 512 == 0), "[W_hi;W_lo] must keep the swizzle phase");
  static constexpr uint32_t idesc_n(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  }
  static constexpr int N_THREADS = ROW3 ? THREADS_ROW3 : THREADS;
  static constexpr int EPI_WARPS = ROW3 ? 8 : 4;                           // arrivals that free an accumulator stage
  static constexpr uint32_t IDESC = idesc_n(NW);
  static constexpr uint32_t IDESC2 = idesc_n(2 * NW);
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
  int nk_last0, nk_last1;  // MMA k-steps (8 channels each) in the LAST chunk of source 0 / 1: skips all-padding k-steps
  int issuers;     // MMA-issuing threads; 1 gives a fixed fp32 accumulation order (bit-reproducible)
  int w_res;       // 1: every (chunk, tap) weight tile of the layer is loaded once and stays in shared memory (the
                   // slot ring then only carries the TMEM A slots); 0: weight tiles stream through the slot ring
  long long* dbg;  // optional timeline dump of CTA 0 (NVS_TC_DEBUG builds only)
  int knock;       // NVS_TC_DEBUG builds: stage knock-out bits for bottleneck experiments (results are then garbage)
};
#ifdef NVS_TC_DEBUG
#define NVS_KNOCK(bit) ((p.knock & (bit)) != 0)
#else
#define NVS_KNOCK(bit) false
#endif

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
#ifdef NVS_TC_DEBUG
__device__ long long* g_timeout_log = nullptr;  // [0] = count, then 4 words per record
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity, int line) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 40000000LL) {  // ~20 ms: record who is stuck where and carry on (results are garbage)
      if (g_timeout_log) {
        const unsigned long long i = atomicAdd(reinterpret_cast<unsigned long long*>(g_timeout_log), 1ull);
        if (i < 200) {
          long long* r = g_timeout_log + 1 + 4 * i;
          r[0] = t0; r[1] = ((long long)blockIdx.x << 32) | threadIdx.x; r[2] = ((long long)bar << 8) | parity; r[3] = line;
        }
      }
      return;
    }
  }
}
#define mbar_wait(b, ph) mbar_wait_((b), (ph), __LINE__)
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();  // never hang the box on a protocol bug
  }
}
#endif
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t* r) {
  static_assert(N == 8 || N == 16 || N == 32, "columns per store");
  if (N == 32) tmem_st32(taddr, r);
  else if (N == 16) tmem_st16(taddr, r);
  else tmem_st8(taddr, r);
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
// two 16-column loads (e.g. the two accumulator halves of the same channels), one wait
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

// K-major swizzled shared-memory matrix descriptor.  KC = 32: 128-byte rows, SWIZZLE_128B (layout 2), 8-row
// groups 1024 B apart.  KC = 16: 64-byte rows, SWIZZLE_64B (layout 4), 8-row groups 512 B apart.
template <int KC>
__device__ __forceinline__ uint64_t make_wdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((KC == 32 ? 1024u : 512u) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(KC == 32 ? 2 : 4) << 61;
  return d;
}

template <int COUT, int KC, bool PAIR, bool ROW3>
__global__ void __launch_bounds__((Cfg<COUT, KC, PAIR, ROW3>::N_THREADS), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
               const Params p) {
  using C = Cfg<COUT, KC, PAIR, ROW3>;
  constexpr int NH = C::NH, NS = C::NS;
  constexpr int TX = C::TX, TY = C::TY, HX = C::HX, HY = C::HY;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(sm + C::SM_BIAS);
  const uint32_t bar0 = base + C::SM_BAR;
  auto hfull = [&](int i) { return bar0 + 8u * i; };
  auto hempty = [&](int i) { return bar0 + 8u * (NH + i); };
  auto sfull = [&](int i) { return bar0 + 8u * (2 * NH + i); };
  auto sempty = [&](int i) { return bar0 + 8u * (2 * NH + NS + i); };
  auto afull = [&](int i) { return bar0 + 8u * (2 * NH + 2 * NS + i); };
  auto aempty = [&](int i) { return bar0 + 8u * (2 * NH + 2 * NS + 2 + i); };
  auto astart = [&](int i) { return bar0 + 8u * (2 * NH + 2 * NS + 4 + i); };
  const uint32_t wres = bar0 + 8u * (2 * NH + 2 * NS + 6);  // resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::SM_BAR + 8 * C::N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < COUT; i += C::N_THREADS) bias_s[i] = p.bias[i];
  if (warp == WARP_HALO && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
    for (int i = 0; i < NH; ++i) {
      mbar_init(hfull(i), 1);
      mbar_init(hempty(i), C::TAPS * 4);  // one elected lane per converter warp per tap releases the box
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(sfull(i), p.w_res ? 4 : 4 + 1);   // 4 converter warps (A slot written) [+ the weight TMA's expect_tx arrive]
      mbar_init(sempty(i), 1);      // tcgen05.commit of the step that used the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(afull(i), p.issuers);   // every MMA issuer commits its share of the tile
      mbar_init(aempty(i), C::EPI_WARPS);
      mbar_init(astart(i), 1);  // the issuer of a tile's first step has queued the overwriting (accumulate=0) MMA
    }
    mbar_init(wres, 1);
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

  const int chunks = p.c0_chunks + p.c1_chunks;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (ROW3 && (warp < 4 || warp >= WARP_EPI2)) {
    // =========================== epilogue, row-stationary variant ===========================
    // Eight warps: TMEM lane quadrant (= image row of the tile) warp % 4, channel half h = 0 (warps 0-3) / 1 (warps
    // 18-21): a tile's epilogue -- six TMEM loads, two shuffles per channel -- is what bounds this kernel, so it is
    // split over twice the warps.  lane = position of the 32-wide strip (lane 0 / 31: halo positions).  Accumulator
    // columns of a stage: [kx0 | kx1 | kx2] x 32 channels of a_hi w_hi, then the same three blocks of the correction
    // products.  out[i] = D_0[i-1] + D_1[i] + D_2[i+1].
    int acc = 0;
    uint32_t aph = 0;
    const int quad = warp & 3, h = warp >= WARP_EPI2 ? 1 : 0;
    const float neg_slope = p.act == NVS_ACT_LRELU ? 0.01f : (p.act == NVS_ACT_RELU ? 0.f : 1.f);
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
      const int gx = tx * TX - 1 + lane, gy = ty * TY + quad;
      const bool valid = lane >= 1 && lane <= TX && gx < p.W && gy < p.H;
      mbar_wait(afull(acc), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * C::ACC_STAGE_COLS) + ((uint32_t)(quad * 32) << 16);
      if (!NVS_KNOCK(4) && !(p.dst_mode == 3 && h == 1)) {  // 16 output channels per warp (keypoint heads: 3 in all)
        float o[16], u[16], w[16];
        tmem_ld16x2(taddr + (uint32_t)(16 * h), taddr + (uint32_t)(C::NW + 16 * h), u, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __shfl_up_sync(0xffffffffu, u[j] + w[j], 1);
        tmem_ld16x2(taddr + (uint32_t)(COUT + 16 * h), taddr + (uint32_t)(C::NW + COUT + 16 * h), u, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] += u[j] + w[j];
        tmem_ld16x2(taddr + (uint32_t)(2 * COUT + 16 * h), taddr + (uint32_t)(C::NW + 2 * COUT + 16 * h), u, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] += __shfl_down_sync(0xffffffffu, u[j] + w[j], 1);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = o[j] + bias_s[16 * h + j];
          o[j] = fmaxf(a, 0.f) + neg_slope * fminf(a, 0.f);
        }
        if (p.act == NVS_ACT_SIGMOID && h == 0) {  // depth heads: cout <= 4
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = 1.f / (1.f + expf(-o[j]));
        }
        const int cbase = 16 * h;
        if (p.dst_mode == 3) {  // keypoint heads: sigmoid -> score, tanh -> centre shift (see the other epilogue)
          if (valid) {
            const size_t plane = (size_t)p.H * p.W, pix = (size_t)gy * p.W + gx;
            p.dst[(size_t)b * plane + pix] = 1.f / (1.f + expf(-o[0]));
            p.dst_pool[((size_t)b * 2 + 0) * plane + pix] = tanhf(o[1]);
            p.dst_pool[((size_t)b * 2 + 1) * plane + pix] = tanhf(o[2]);
          }
        } else {
          if (p.dst_pool != nullptr) {
            // MaxPool2d(2,2): x partner = next lane (strips start at even x, so pairs are lanes (1,2), (3,4), ...),
            // y partner = the row of the next quadrant's warp: odd quadrants hand their x-pooled values to the even
            // quadrant below them through shared memory (one buffer and one named barrier per (h, quadrant pair);
            // the pair's second barrier keeps the next tile's write behind this tile's read)
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
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(aempty(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        aph ^= 1;
      }
    }
  } else if (warp == WARP_HALO) {
    // =========================== halo producer: one box per (tile, chunk) ===========================
    if (lane == 0) {
      int hb = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(hempty(hb), ph ^ 1);
          if (NVS_KNOCK(64)) {
            mbar_arrive(hfull(hb));
          } else {
          mbar_expect_tx(hfull(hb), HX * HY * C::HALO_ROW_BYTES);
          if (ch < p.c0_chunks)
            tma_load_4d(base + hb * C::HALO_BYTES, &map_a0, hfull(hb), p.c0_off + ch * C::HC, tx * TX - 1, ty * TY - 1, b);
          else
            tma_load_4d(base + hb * C::HALO_BYTES, &map_a1, hfull(hb), p.c1_off + (ch - p.c0_chunks) * C::HC,
                        tx * TX - 1, ty * TY - 1, b);
          }
          if (++hb == NH) {
            hb = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == WARP_W) {
    // =========================== weight producer: one [W_hi;W_lo] tile per (chunk, tap) ===========================
    if (lane == 0 && p.w_res) {
      // resident weights: every (chunk, tap) tile once, one barrier for all of them
      if (my_tiles > 0) {
        mbar_expect_tx(wres, (uint32_t)(chunks * C::TAPS * C::W_STAGE));
        for (int ch = 0; ch < chunks; ++ch)
          for (int tap = 0; tap < C::TAPS; ++tap) {
            const uint32_t dst = base + C::SM_W + (ch * C::TAPS + tap) * C::W_STAGE;
            tma_load_3d(dst, &map_whi, wres, ch * KC, 0, tap);
            tma_load_3d(dst + C::W_BYTES, &map_wlo, wres, ch * KC, 0, tap);
          }
      }
    } else if (lane == 0) {
      int sl = 0;
      uint32_t ph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int ch = 0; ch < chunks; ++ch) {
          for (int tap = 0; tap < C::TAPS; ++tap) {
            mbar_wait(sempty(sl), ph ^ 1);
            const uint32_t dst = base + C::SM_W + sl * C::W_STAGE;
            if (NVS_KNOCK(32)) {
              mbar_arrive(sfull(sl));
            } else {
            mbar_expect_tx(sfull(sl), C::W_STAGE);
            tma_load_3d(dst, &map_whi, sfull(sl), ch * KC, 0, tap);
            tma_load_3d(dst + C::W_BYTES, &map_wlo, sfull(sl), ch * KC, 0, tap);
            }
            if (++sl == NS) {
              sl = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < WARP_HALO) {
    // =========================== converters: halo window -> hi / lo -> TMEM ===========================
    const int cw = warp & 3;            // TMEM lane quadrant this warp may write (warp % 4)
    const int grp = (warp - 4) >> 2;    // converter group: handles pipeline steps with (index % CONV_GROUPS == grp)
    const int r = cw * 32 + lane;       // GEMM row inside the tile == TMEM lane
    // 8 x 16 output pixels, or (ROW3) 4 image rows x 32 positions of the halo box (position 0 / 31 = halo columns)
    const int ly = ROW3 ? cw : (r >> 4), lx = ROW3 ? lane : (r & 15);
    int hb = 0, sl = 0, turn = 0;
    uint32_t hph = 0, sph = 0;
    // Splitting the channels of every step over both groups instead (half the latency per step, same work) was
    // measured slower: the two barrier waits, the tcgen05.wait::st and the fences are per warp and step.
    for (int t = 0; t < my_tiles; ++t) {
      for (int ch = 0; ch < chunks; ++ch) {
        for (int tap = 0; tap < C::TAPS; ++tap) {
          const bool mine = turn == grp;
          if (++turn == CONV_GROUPS) turn = 0;
          if (mine) {
#ifdef NVS_TC_DEBUG
            const long long k0 = clock64();
#endif
            if (tap < CONV_GROUPS) mbar_wait(hfull(hb), hph);  // first tap of this chunk for my group
#ifdef NVS_TC_DEBUG
            const long long k1 = clock64();
#endif
            // 16-byte chunk c of halo row rh sits at chunk (c ^ key): SWIZZLE_128B key = rh & 7, SWIZZLE_64B key =
            // (rh >> 1) & 3.  Consecutive pixels of a quarter-warp read consecutive rows -> distinct keys ->
            // conflict-free LDS.128.
            const uint8_t* hbase = sm + hb * C::HALO_BYTES;
            const int ky = PAIR ? 0 : (ROW3 ? tap : tap / 3), kx = (PAIR || ROW3) ? 0 : tap - ky * 3;
            const int rh = (ly + ky) * HX + lx + kx;  // row of the halo box (single-tap steps)
            const uint8_t* rowp = hbase + rh * C::HALO_ROW_BYTES;
            const int key = C::HC == 32 ? (rh & 7) : ((rh >> 1) & 3);
            uint32_t hi[KC], lo[KC];
            if (!NVS_KNOCK(2)) {
#pragma unroll
              for (int c = 0; c < KC / 4; ++c) {
                float4 v;
                if (PAIR) {  // chunks 0-3: tap 2*tap, chunks 4-7: tap 2*tap + 1 (the tenth tap is all zero)
                  const int t2 = 2 * tap + (c >> 2);
                  const int ky2 = t2 / 3, kx2 = t2 - ky2 * 3;
                  const int rh2 = (ly + ky2) * HX + lx + kx2;
                  v = t2 < 9 ? *reinterpret_cast<const float4*>(hbase + rh2 * C::HALO_ROW_BYTES +
                                                                (((c & 3) ^ ((rh2 >> 1) & 3)) << 4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                  v = NVS_KNOCK(16) ? make_float4(1.f * tap, 2.f, 3.f * lane, 4.f)
                                    : *reinterpret_cast<const float4*>(rowp + ((c ^ key) << 4));
                }
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // hi = a rounded to nearest tf32 (13 low mantissa bits), lo = a - hi exactly, then lo rounded
                  // to nearest tf32 too: the tensor core would otherwise TRUNCATE lo, a one-sided (biased) error
                  const uint32_t h = (__float_as_uint(f[e]) + 0x1000u) & 0xFFFFE000u;
                  hi[4 * c + e] = h;
                  lo[4 * c + e] = __float_as_uint(f[e] - __uint_as_float(h)) + 0x1000u;  // low 13 bits: dropped by the MMA
                }
              }
            }
            // the loads and the split do not need the slot: only now wait for the MMAs that last read it
            mbar_wait(sempty(sl), sph ^ 1);
            tc_fence_after();
#ifdef NVS_TC_DEBUG
            const long long k2 = clock64();
#endif
            if (!NVS_KNOCK(2)) {
              const uint32_t ta = tmem_base + (uint32_t)(C::ACC_COLS + sl * C::A_COLS) + ((uint32_t)(cw * 32) << 16);
              if (!NVS_KNOCK(8)) {
                tmem_st<KC>(ta, hi);
                tmem_st<KC>(ta + KC, lo);
              } else {
                uint32_t acc_x = 0;
#pragma unroll
                for (int c = 0; c < KC; ++c) acc_x ^= hi[c] + lo[c];
                if (acc_x == 0x12345678u) p.dbg[4095] = 1;
              }
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(sfull(sl));   // this warp's 32 TMEM lanes of the slot are written
              mbar_arrive(hempty(hb));  // ... and its reads of the halo box for this tap are done
            }
#ifdef NVS_TC_DEBUG
            {
              const int gs = (t * chunks + ch) * C::TAPS + tap;
              if (p.dbg && blockIdx.x == 0 && warp == 4 + 4 * grp && lane == 0 && gs < 480) {
                long long* d = p.dbg + 2048 + gs * 4;
                d[0] = k0; d[1] = k1; d[2] = k2; d[3] = clock64();
              }
            }
#endif
          }
          if (++sl == NS) {
            sl = 0;
            sph ^= 1;
          }
        }
        if (++hb == NH) {
          hb = 0;
          hph ^= 1;
        }
      }
    }
  } else if (warp >= WARP_MMA) {
    // =========================== MMA issuers (p.issuers threads) ===========================
    // A tcgen05.mma instruction blocks its issuing thread while the tensor pipe's short queue (2-3 MMAs) is
    // full, one thread cannot queue MMAs faster than ~80 cycles apiece (the pipe retires an N=64 MMA in 32), and
    // every pipeline step costs its issuer ~400 cycles of mbarrier polling, tcgen05.commit and loop overhead
    // during which the pipe drains (NVS_TC_DEBUG timeline, tools/tc_timeline.py).  Several issuers therefore
    // take the pipeline steps round-robin (global step index mod p.issuers): their MMAs interleave in the
    // pipe at its full rate and their overheads overlap.  Accumulation order inside a tile is irrelevant
    // except for the very first (overwriting) MMA, which is ordered by the astart barrier; each issuer commits
    // the slots it consumed, and all of them commit the tile's accumulator (afull count p.issuers).
    const int me = warp - WARP_MMA;
    if (lane == 0 && my_tiles > 0 && me < p.issuers) {
      const int nis = p.issuers;
      const int steps = C::TAPS * chunks;
      const int total = my_tiles * steps;
      const uint64_t wdesc0 = make_wdesc<KC>(base + C::SM_W);
      const uint32_t a0 = tmem_base + (uint32_t)C::ACC_COLS;
      int tile = 0, ks = me;  // my current step: tile index (CTA-local) and step inside the tile
      while (ks >= steps) {
        ks -= steps;
        ++tile;
      }
      int last_tile_synced = -1;
      if (p.w_res) mbar_wait(wres, 0);
      for (int g = me; g < total; g += nis) {
        const int acc = tile % ACC_STAGES;
        const uint32_t aph = (uint32_t)(tile / ACC_STAGES) & 1u;
        const int sl = g % NS;
        const uint32_t sph = (uint32_t)(g / NS) & 1u;
        if (last_tile_synced != tile) {  // my first step in this tile
          mbar_wait(aempty(acc), aph ^ 1);
          if (ks != 0) mbar_wait(astart(acc), aph);  // another issuer queued the tile's first MMA
          last_tile_synced = tile;
        }
#ifdef NVS_TC_DEBUG
        const long long c0 = clock64();
#endif
        mbar_wait(sfull(sl), sph);
        tc_fence_after();
#ifdef NVS_TC_DEBUG
        const long long c1 = clock64();
#endif
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C::ACC_STAGE_COLS);
        const uint32_t a_hi = a0 + (uint32_t)(sl * C::A_COLS), a_lo = a_hi + KC;
        // weight tile: stage sl of the ring, or (resident) the tile of this step = (chunk, tap) = ks
        const uint64_t w_hi = wdesc0 + (uint64_t)((((p.w_res ? ks : sl)) * C::W_STAGE) >> 4),
                       w_lo = w_hi + (uint64_t)(C::W_BYTES >> 4);
        auto issue_kstep = [&](int k) {
          // A: 8 tf32 = 8 TMEM columns; B: 8 tf32 = 32 bytes along K = +2 in the descriptor's (addr >> 4)
          const uint64_t o = (uint64_t)(2 * k);
          if (C::CONCAT) {
            tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, w_hi + o, C::IDESC2, (ks | k) != 0 ? 1u : 0u);  // x [W_hi;W_lo]
            if (k == 0 && ks == 0) mbar_arrive(astart(acc));
            // a_lo w_hi goes to the SECOND half as well: the tensor core's fp32 adder truncates, one (biased)
            // rounding per accumulation, so the big a_hi w_hi sums take half as many of them and the small
            // correction terms are rounded at their own, 2^-11 times smaller, magnitude
            tc_mma_tf32_ts(d_tmem + (uint32_t)C::NW, a_lo + 8 * k, w_hi + o, C::IDESC, 1u);
          } else {
            tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, w_hi + o, C::IDESC, (ks | k) != 0 ? 1u : 0u);
            if (k == 0 && ks == 0) mbar_arrive(astart(acc));
            tc_mma_tf32_ts(d_tmem, a_lo + 8 * k, w_hi + o, C::IDESC, 1u);
            tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, w_lo + o, C::IDESC, 1u);
          }
        };
        // k-steps of this chunk that can be non-zero: the last chunk of a zero-padded source (N letters: 24 real
        // channels in a 32-channel row ...) skips its all-padding k-steps.  The full chunk is the straight-line
        // common path: the issuing thread's instruction count per MMA is what this kernel is most sensitive to.
        int nk = C::KSTEPS;
        if (p.nk_last0 != C::KSTEPS || p.nk_last1 != C::KSTEPS) {
          const int chk = ks / C::TAPS;
          nk = chk == p.c0_chunks - 1 ? p.nk_last0 : (chk == chunks - 1 ? p.nk_last1 : C::KSTEPS);
        }
        if (NVS_KNOCK(1)) {
          if (ks == 0) mbar_arrive(astart(acc));
        } else if (nk == C::KSTEPS) {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k) issue_kstep(k);
        } else {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k)
            if (k < nk) issue_kstep(k);
        }
#ifdef NVS_TC_DEBUG
        const long long c2 = clock64();
#endif
        tc_commit(sempty(sl));  // weight stage + TMEM A slot are free once these MMAs retire
        if (ks >= steps - nis) tc_commit(afull(acc));  // my last step of this tile
#ifdef NVS_TC_DEBUG
        if (p.dbg && blockIdx.x == 0 && g < 512) {
          long long* d = p.dbg + g * 4;
          d[0] = c0; d[1] = c1; d[2] = c2; d[3] = clock64();
        }
#endif
        ks += nis;
        while (ks >= steps) {
          ks -= steps;
          ++tile;
        }
      }
    }
  } else if (warp < 4) {
    // =========================== epilogue ===========================
    int acc = 0;
    uint32_t aph = 0;
    const int ly = 2 * warp + (lane >> 4), lx = lane & 15;  // pixel inside the tile (TMEM lane = ly*16 + lx)
    const float neg_slope = p.act == NVS_ACT_LRELU ? 0.01f : (p.act == NVS_ACT_RELU ? 0.f : 1.f);
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
      const int gx = tx * TX + lx, gy = ty * TY + ly;
      const bool valid = gx < p.W && gy < p.H;
      mbar_wait(afull(acc), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * C::ACC_STAGE_COLS) + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < (NVS_KNOCK(4) ? 0 : COUT / 32); ++cc) {
        float v[32];
        __syncwarp();
        tmem_ld32(taddr + (uint32_t)(cc * 32), v);
        if (C::CONCAT) {
          float v2[32];
          tmem_ld32(taddr + (uint32_t)(COUT + cc * 32), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
        // none / LeakyReLU(0.01) / ReLU as max(a,0) + slope * min(a,0): branch free
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float a = v[j] + bias_s[cc * 32 + j];
          v[j] = fmaxf(a, 0.f) + neg_slope * fminf(a, 0.f);
        }
        if (p.act == NVS_ACT_SIGMOID) {  // depth heads (kp2dtiny.py:589, :956): cout <= 4, kept off the common path
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
        }
        const int cbase = cc * 32;
        if (p.dst_mode == 3) {
          // keypoint heads (kp2dtiny.py:574-575, 927-935): channel 0 -> sigmoid -> score (B,1,H,W);
          // channels 1,2 -> tanh -> centre shift (B,2,H,W), two NCHW outputs (dst / dst_pool pointer)
          if (valid && cc == 0) {
            const size_t plane = (size_t)p.H * p.W, pix = (size_t)gy * p.W + gx;
            p.dst[(size_t)b * plane + pix] = 1.f / (1.f + expf(-v[0]));
            p.dst_pool[((size_t)b * 2 + 0) * plane + pix] = tanhf(v[1]);
            p.dst_pool[((size_t)b * 2 + 1) * plane + pix] = tanhf(v[2]);
          }
          continue;
        }
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
      if (lane == 0) mbar_arrive(aempty(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        aph ^= 1;
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
  CUtensorMap a0, a1, whi, wlo;
  Params p;
  int cout_tpl, kc, pair, row3;
  int magic;
};
constexpr int PLAN_MAGIC = 0x7C0DE6;

// NHWC activations (B,H,W,Ct): box = (kc channels, hx, hy, 1 frame) = output tile + 1-pixel halo (18 x 10, or 32 x 6
// for the row-stationary variant)
static int encode_act(CUtensorMap* m, const float* ptr, int B, int H, int W, int Ct, int kc, int hx, int hy) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Ct * 4, (cuuint64_t)W * Ct * 4, (cuuint64_t)H * W * Ct * 4};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)hx, (cuuint32_t)hy, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}
// packed weights [taps][rows][cin]: box = (kc, rows, 1); rows = cout_pad, or 3 * cout_pad ([kx][cout]) with taps = 3
// kernel rows for the row-stationary variant
static int encode_w(CUtensorMap* m, const float* ptr, int cin, int cout_pad, int kc, int taps = 9) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)cin * 4, (cuuint64_t)cout_pad * cin * 4};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)cout_pad, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}

template <int COUT, int KC, bool PAIR = false, bool ROW3 = false>
static int launch(const Plan& pl, const Params& p, cudaStream_t st) {
  using C = Cfg<COUT, KC, PAIR, ROW3>;
  auto kern = conv_tc_kernel<COUT, KC, PAIR, ROW3>;
  NVS_OPT_IN_SMEM(kern, 227 * 1024);
  const int sms = nvs_sm_count();
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  Params q = p;
  // ROW3: a tile lasts ~1 us, far too short to hide a 24 KB weight TMA behind two slots, and the whole layer's weights
  // (3 tiles per 32-channel chunk) usually fit next to the halo boxes: load them once per CTA
  const int w_tiles = C::TAPS * (p.c0_chunks + p.c1_chunks);
  static const bool allow_res = getenv("NVS_TC_WRES") == nullptr || atoi(getenv("NVS_TC_WRES")) != 0;
  q.w_res = (ROW3 && allow_res && w_tiles <= C::MAX_W_RESIDENT) ? 1 : 0;
  const int smem = C::smem_bytes(q.w_res ? w_tiles : C::NS);
  // a slot must always be consumed by the same issuer, and every issuer needs a step in every tile (all of them commit
  // the tile's accumulator): issuers | ring size, issuers <= steps per tile
  const int steps = C::TAPS * (p.c0_chunks + p.c1_chunks);
  while (q.issuers > 1 && (C::NS % q.issuers != 0 || q.issuers > steps)) --q.issuers;
  kern<<<grid, C::N_THREADS, smem, st>>>(pl.a0, pl.a1, pl.whi, pl.wlo, q);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

#ifdef NVS_TC_DEBUG
static long long* g_dbg = nullptr;
static int g_knock = 0;
#endif

static inline int pick_kc(int c0, int c1) {
  if (c0 % 32 == 0 && c1 % 32 == 0) return 32;
  if (c0 % 16 == 0 && c1 % 16 == 0) return 16;
  return 0;
}

}  // namespace tc
}  // namespace nvs

namespace nvs {
namespace rs {  // conv_rs.cu: the "3xFP16" row-stationary kernel (NvsConvTcArgs.flags bit 4)
size_t plan_bytes();
bool is_plan(const void* plan_mem);
int plan_init(void* plan_mem, const NvsConvTcArgs* a);
int run(const void* plan_mem, float* dst_override, float* dst2_override, cudaStream_t st);
}  // namespace rs
}  // namespace nvs

using namespace nvs;

extern "C" int32_t nvs_conv_tc_cout_pad(int32_t cout) {
  if (cout <= 0 || cout > 128) return 0;
  return cout <= 32 ? 32 : (cout <= 64 ? 64 : 128);
}

extern "C" int32_t nvs_conv_tc_supported(int32_t c0, int32_t c1, int32_t cout) {
  if (c0 <= 0 || c1 < 0) return 0;
  const int kc = tc::pick_kc(c0, c1);
  const int cpad = nvs_conv_tc_cout_pad(cout);
  if (kc == 0 || cpad == 0) return 0;
  if (kc == 16 && cpad != 32) return 0;  // the 16-channel variant is instantiated for the stem layer only
  return 1;
}

extern "C" size_t nvs_conv_tc_plan_bytes(void) {
  const size_t a = sizeof(tc::Plan) + 64, b = rs::plan_bytes();
  return a > b ? a : b;
}

extern "C" int nvs_conv_tc_plan_init(void* plan_mem, const NvsConvTcArgs* a) {
  if (!plan_mem || !a || !a->src0 || !a->w_hi || !a->w_lo || !a->bias) return NVS_ERR_ARG;
  if (a->flags & 16) {  // "3xFP16" row-stationary kernel
    if (a->c1 > 0 && !a->src1) return NVS_ERR_ARG;
    if (a->B <= 0 || a->H <= 0 || a->W <= 0) return NVS_ERR_ARG;
    if ((a->c0_total % 4) || (a->c0_off % 4) || (a->c1 > 0 && ((a->c1_total % 4) || (a->c1_off % 4)))) return NVS_ERR_ARG;
    if (a->dst_mode != 0 && ((a->dst_c_total % 4) || (a->dst_c_off % 4)) && a->dst_layout == 0) return NVS_ERR_ARG;
    if (a->act != NVS_ACT_NONE && a->act != NVS_ACT_LRELU && a->act != NVS_ACT_RELU && a->act != NVS_ACT_SIGMOID)
      return NVS_ERR_UNSUPPORTED;
    if (a->dst_mode == 3 && (a->cout != 3 || a->act != NVS_ACT_NONE)) return NVS_ERR_ARG;
    if (a->act == NVS_ACT_SIGMOID && a->cout > 4) return NVS_ERR_UNSUPPORTED;
    if (a->dst_mode == 0 && a->dst_pool == nullptr) return NVS_ERR_ARG;
    return rs::plan_init(plan_mem, a);
  }
  if (!nvs_conv_tc_supported(a->c0, a->c1, a->cout)) return NVS_ERR_UNSUPPORTED;
  if (a->c1 > 0 && !a->src1) return NVS_ERR_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0) return NVS_ERR_ARG;
  if ((a->c0_total % 4) || (a->c0_off % 4) || (a->c1 > 0 && ((a->c1_total % 4) || (a->c1_off % 4)))) return NVS_ERR_ARG;
  if (a->dst_mode != 0 && ((a->dst_c_total % 4) || (a->dst_c_off % 4)) && a->dst_layout == 0) return NVS_ERR_ARG;
  if (a->dst_mode == 2 && (a->cout % 32) != 0) return NVS_ERR_UNSUPPORTED;
  if (a->act != NVS_ACT_NONE && a->act != NVS_ACT_LRELU && a->act != NVS_ACT_RELU && a->act != NVS_ACT_SIGMOID)
    return NVS_ERR_UNSUPPORTED;
  if (a->dst_mode == 3 && (a->cout != 3 || a->act != NVS_ACT_NONE)) return NVS_ERR_ARG;
  if (a->act == NVS_ACT_SIGMOID && a->cout > 4) return NVS_ERR_UNSUPPORTED;  // single-channel depth outputs only
  tc::Plan* pl = reinterpret_cast<tc::Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  const int cpad = nvs_conv_tc_cout_pad(a->cout);
  const int cin = a->c0 + a->c1;
  const int pair = (a->flags & 2) ? 1 : 0;  // weights given in the paired-tap layout [5][cout_pad][32] (c0 = 16, no src1)
  if (pair && (a->c0 != 16 || a->c1 != 0 || cpad != 32)) return NVS_ERR_ARG;
  int kc = tc::pick_kc(a->c0, a->c1);
  // bit 2: the row-stationary kernel (see Cfg: ROW3); bit 3: ... with 16-channel chunks (four A slots instead of two)
  const int row3 = (a->flags & 4) ? 1 : 0;
  if (row3 && (a->flags & 8) && kc != 0) kc = 16;
  if (row3 && (pair || cpad != 32 || kc == 0 || a->dst_mode == 2 || (a->dst_mode == 0 && a->dst_pool == nullptr)))
    return NVS_ERR_ARG;
  using R3 = tc::Cfg<32, 32, false, true>;
  using R9 = tc::Cfg<32, 32, false, false>;
  const int hx = row3 ? R3::HX : R9::HX, hy = row3 ? R3::HY : R9::HY;
  int rc = tc::encode_act(&pl->a0, a->src0, a->B, a->H, a->W, a->c0_total, kc, hx, hy);
  if (rc != NVS_OK) return rc;
  rc = tc::encode_act(&pl->a1, a->c1 > 0 ? a->src1 : a->src0, a->B, a->H, a->W, a->c1 > 0 ? a->c1_total : a->c0_total, kc,
                      hx, hy);
  if (rc != NVS_OK) return rc;
  const int wrows = row3 ? 3 * cpad : cpad, wtaps = row3 ? 3 : 9;
  rc = pair ? tc::encode_w(&pl->whi, a->w_hi, 32, cpad, 32, 5) : tc::encode_w(&pl->whi, a->w_hi, cin, wrows, kc, wtaps);
  if (rc != NVS_OK) return rc;
  rc = pair ? tc::encode_w(&pl->wlo, a->w_lo, 32, cpad, 32, 5) : tc::encode_w(&pl->wlo, a->w_lo, cin, wrows, kc, wtaps);
  if (rc != NVS_OK) return rc;
  tc::Params& p = pl->p;
  p.bias = a->bias; p.dst = a->dst; p.dst_pool = a->dst_pool;
  p.c0_off = a->c0_off; p.c0_chunks = a->c0 / kc; p.c1_off = a->c1_off; p.c1_chunks = a->c1 / kc;
  p.dst_c_total = a->dst_c_total; p.dst_c_off = a->dst_c_off; p.dst_layout = a->dst_layout; p.dst_mode = a->dst_mode;
  p.pool_c_total = a->pool_c_total; p.pool_c_off = a->pool_c_off;
  p.B = a->B; p.H = a->H; p.W = a->W; p.cout = a->cout; p.act = a->act;
  const int tile_x = row3 ? R3::TX : R9::TX, tile_y = row3 ? R3::TY : R9::TY;
  p.tiles_x = (a->W + tile_x - 1) / tile_x; p.tiles_y = (a->H + tile_y - 1) / tile_y;
  p.n_tiles = p.tiles_x * p.tiles_y * a->B;
  {
    static int default_issuers = 0;
    if (!default_issuers) {
      const char* e = getenv("NVS_TC_ISSUERS");
      default_issuers = e ? atoi(e) : 4;
      // a slot must always be consumed by the same issuer (parity waits tolerate no lapping): issuers | ring size
      if (default_issuers < 1 || default_issuers > tc::MAX_ISSUERS) default_issuers = tc::MAX_ISSUERS;
    }
    p.issuers = (a->flags & 1) ? 1 : default_issuers;
  }
  {
    // real channels of the last chunk of each source -> k-steps worth issuing (paired-tap rows interleave two taps:
    // all four k-steps carry data)
    const int ksteps = (pair ? 32 : kc) / 8;
    auto last_nk = [&](int c, int real) {
      if (pair || real <= 0 || real >= c) return ksteps;
      const int in_last = real - (c / kc - 1) * kc;  // real channels that fall into the last chunk
      if (in_last <= 0) return ksteps;               // (a whole chunk of padding is not expected; keep it exact)
      return (in_last + 7) / 8;
    };
    p.nk_last0 = last_nk(a->c0, a->c0_real);
    p.nk_last1 = a->c1 > 0 ? last_nk(a->c1, a->c1_real) : ksteps;
    if (a->c1 == 0) p.nk_last1 = p.nk_last0;  // single source: its last chunk is the last chunk
  }
  p.dbg = nullptr;
  p.knock = 0;
  p.w_res = 0;  // decided per launch (tc::launch)
  pl->cout_tpl = cpad;
  pl->kc = kc;
  pl->pair = pair;
  pl->row3 = row3;
  pl->magic = tc::PLAN_MAGIC;
  return NVS_OK;
}

extern "C" int nvs_conv_tc_run(const void* plan_mem, float* dst_override, float* dst2_override, void* stream) {
  if (!plan_mem) return NVS_ERR_ARG;
  if (rs::is_plan(plan_mem)) return rs::run(plan_mem, dst_override, dst2_override, static_cast<cudaStream_t>(stream));
  const tc::Plan* pl = reinterpret_cast<const tc::Plan*>(((uintptr_t)plan_mem + 63) & ~(uintptr_t)63);
  if (pl->magic != tc::PLAN_MAGIC) return NVS_ERR_ARG;
  tc::Params p = pl->p;
  if (dst_override) p.dst = dst_override;
  if (dst2_override) p.dst_pool = dst2_override;
  if (p.dst_mode == 3 && (!p.dst || !p.dst_pool)) return NVS_ERR_ARG;
#ifdef NVS_TC_DEBUG
  p.dbg = tc::g_dbg;
  p.knock = tc::g_knock;
#endif
  if (p.dst_mode != 0 && !p.dst) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl->row3) return pl->kc == 16 ? tc::launch<32, 16, false, true>(*pl, p, st) : tc::launch<32, 32, false, true>(*pl, p, st);
  if (pl->pair) return tc::launch<32, 32, true>(*pl, p, st);
  if (pl->kc == 16) return pl->cout_tpl == 32 ? tc::launch<32, 16>(*pl, p, st) : NVS_ERR_UNSUPPORTED;
  switch (pl->cout_tpl) {
    case 32: return tc::launch<32, 32>(*pl, p, st);
    case 64: return tc::launch<64, 32>(*pl, p, st);
    case 128: return tc::launch<128, 32>(*pl, p, st);
  }
  return NVS_ERR_UNSUPPORTED;
}

#ifdef NVS_TC_DEBUG
extern "C" void nvs_conv_tc_set_debug(long long* dev_buf) { nvs::tc::g_dbg = dev_buf; }
extern "C" void nvs_conv_tc_set_knock(int bits) { nvs::tc::g_knock = bits; }
extern "C" void nvs_conv_tc_set_timeout_log(long long* dev_buf) {
  cudaMemcpyToSymbol(nvs::tc::g_timeout_log, &dev_buf, sizeof(dev_buf));
}
#endif
