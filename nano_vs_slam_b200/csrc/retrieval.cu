// Exact L2 top-k retrieval = faiss.IndexFlatL2.add / search (src/evaluation/global_descriptor.py:55-60).
//
//   d2(q, x) = |q|^2 + |x|^2 - 2 q.x  in fp32,  k smallest, ascending, ties -> lower id, int64 labels.
//
// The Q x N x D contraction is the one genuinely tensor-core shaped piece of the hot path
// (10k x 1M x 4096 = 8.2e13 FLOP).  It runs in half precision as a SCREEN with a proven error bound, and everything
// the screen cannot decide is re-ranked with the literal fp32 formula, so the result is the fp32 result:
//
//   add     rows -> fp16 (scaled by one power of two so the largest component sits in [2^14, 2^15); fp16 subnormals
//           are flushed here, not by the hardware) + |x|^2 + per-shard max |x| and max |x - fp16(x)| (the residual).
//   search  queries -> fp16 (per-row power-of-two scale) with |q|, |q - fp16(q)|.  By Cauchy-Schwarz
//             |q.x - q~.x~| <= |q - q~| |x| + |q~| |x - x~|
//           which, with the fp32 accumulation of the tensor core and of the re-rank bounded generously, gives a
//           per-query eps with |s~ - s| <= eps for every row, s = |x|^2 - 2 q.x.  If tau~ is ANY upper bound of the
//           k-th smallest s~, every row of the true top-k has s~ <= tau~ + 2 eps.  So:
//   GEMM    warp-specialised tcgen05 kernel (TMA SWIZZLE_128B -> smem ring -> tcgen05.mma kind::f16 -> 2 x 256-column
//           TMEM accumulators).  Clusters of 8 / 4 / 2 CTAs work on the same 256-row database tile, every CTA on its own
//           128-query block; each loads a slice of the tile for all (TMA multicast); CTA pairs issue ONE M = 256
//           cta_group::2 MMA over both query blocks.  A cluster owns one contiguous range of (query group, tile) steps;
//           the part of it inside one query group is a SEGMENT.  The epilogue keeps, per query row and segment, the
//           L >= k smallest s~ that are <= (published bound + 2 eps) in shared memory (append + warp-cooperative
//           compaction); every compaction and every finished segment publish the k-th smallest so far (an upper
//           bound of tau~).
//   select  per query: exact k-th smallest s~ over all segment lists by radix select = tau~; candidates = everything with
//           s~ <= tau~ + 2 eps (as many as the data puts there, not a constant).  A query is FLAGGED when the proof
//           has a hole: a segment list that is full with all entries inside the slack (it may have dropped a row), or
//           more candidates than the re-rank buffer holds.
//   rerank  exact fp32 distances of the candidates, k smallest by (distance, id).
//   scan    flagged queries only (normally none): exact fp32 distances against EVERY row, merged into the result
//           under a per-query lock.  Slow, never wrong: dense near-ties or duplicate rows cost time, not exactness.
//
// The answer therefore does not depend on how the work was scheduled, and equals an fp32 IndexFlatL2 wherever the
// fp32 distances themselves are distinct.
#include <cuda.h>
#include <cuda_fp16.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace nvs {
namespace rt {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;       // 16 KiB
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int THREADS = 256;                // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 epilogue
constexpr int KMAX = 64;                    // largest k served
constexpr int CAP = 256;                    // candidates re-ranked exactly per query (more -> exact scan)
constexpr int MAX_CAND = 12288;             // list slots * L the selection kernel holds in shared memory (96 KiB)

// shared memory map (dynamic, 1024-byte aligned base): STAGES operand stages, then the per-row lists.
// PAIR = false: a CTA holds the whole 256-row database tile (48 KiB stages): 4 stages leave room for lists of 31 entries,
// 3 stages for 79, 2 for 127 (odd pitch: conflict-free row-per-thread access).  PAIR = true (tcgen05 cta_group::2): a CTA
// holds its half of the tile (32 KiB stages): 6 stages -> 31, 5 -> 63, 4 -> 95, 3 -> 127.
template <int STAGES, bool PAIR>
struct Smem {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;                 // database rows of a tile in this CTA's stages
  static constexpr int STAGE_BYTES = A_BYTES + B_ROWS * BK * 2;
  static constexpr int LROOM = (227 * 1024 - 1024 - 256 - ACC_STAGES * BN * 4 - STAGES * STAGE_BYTES) / (BM * 8);
  static constexpr int LMAX = LROOM >= 127 ? 127 : ((LROOM - 1) | 1);
  static constexpr int LIST_D = STAGES * STAGE_BYTES;               // float [128][LMAX]
  static constexpr int LIST_I = LIST_D + BM * LMAX * 4;             // int   [128][LMAX]
  static constexpr int XN = LIST_I + BM * LMAX * 4;                 // float [2][256]
  static constexpr int BAR = XN + ACC_STAGES * BN * 4;              // mbarriers
  static constexpr int TOTAL = BAR + 256;
  static constexpr int BYTES = TOTAL + 1024;                        // slack for manual 1024 B alignment
  static_assert(BYTES <= 227 * 1024, "shared memory budget");
};

// ---------------------------------------------------------------- PTX wrappers
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
// Bounded wait: a protocol bug must not hang the GPU box -- after ~4 s the kernel traps instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// the same, adding the cycles spent waiting to `acc` when the wait profile is on (nvs_flat_debug_buffer)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool on, long long& acc) {
  if (!on) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0,
                                               int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// ---- cta_group::2 (CTA pair = the two SMs of a TPC work on one 256-row MMA; operands come from both shared memories)
// TMA load whose completion is counted on the barrier `bar_cluster` (a shared::cluster address: the pair LEADER's barrier)
// L2 eviction policies of the TMA loads (the encodings createpolicy.fractional produces for fraction 1.0)
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull,
                   L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// the same, delivered to every CTA of cta_mask; each destination counts it on the barrier of ITS pair leader
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster,
                                                   int c0, int c1, uint16_t cta_mask, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "h"(cta_mask), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc2(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta_rank) {  // -> shared::cluster address
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// wait on a barrier that CTAs of the cluster arrive on with ordinary (generic-proxy) arrives
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster_t(uint32_t bar, uint32_t parity, bool on, long long& acc) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
  if (on) acc += clock64() - t0;
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of the (converged) warp.  Code guarded by it is single-threaded AND the compiler knows that warp-uniform
// values stay uniform: tcgen05 operands go to uniform registers directly.  Issued from a plain `if (lane == 0)` region
// every tcgen05.mma was wrapped in a per-lane ELECT / R2UR.BROADCAST loop (~15 SASS instructions, ~100 cycles of a
// lone thread's issue latency per MMA -- as long as the 128-cycle MMA itself).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_mma_f16_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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
// 32 lanes x 32 columns of 32-bit: thread t of the warp receives columns [c, c+32) of lane (quadrant*32 + t)
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 UMMA): rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO); LBO unused for swizzled K-major; version 1; layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = f32 (bit 4), A = B = fp16 (format 0), both K-major, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// cta_group::2: M = 256 (128 rows in each CTA of the pair), N = 256 (128 database rows from each CTA's shared memory)
constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

struct Params {
  const float* xnorm;   // [N]    |x|^2 (fp32, from the fp32 rows)
  const float* qcoef;   // [Qpad] -2 / (query scale * database scale): accumulator -> -2 q~.x~
  const float* qslack;  // [Qpad] 2 eps of the query
  float* gbound;        // [Qpad] running upper bound of each query's k-th smallest s~, init +inf
  float* cand_d;        // [Qpad][n_slots][L]  s~ = |x|^2 - 2 q~.x~
  int32_t* cand_i;      // [Qpad][n_slots][L]  row index inside this shard (-1 = empty; preset by the host)
  int Q, N, kblocks;    // kblocks = Dpad / 64
  int k, L;             // neighbours wanted, list length (>= k; the buffers hold L + 32 or more)
  int n_mblk, n_slots, n_tiles;  // n_slots: lists per query = most clusters that can share one query-block group
  int n_mgrp;           // ceil(n_mblk / cs): a cluster of cs CTAs owns query blocks cs*grp .. cs*grp + cs - 1
  int cs;               // cluster size (2, 4 or 8): CTAs sharing every database tile by multicast
  int hint;             // L2 eviction hints of the operand loads (NVS_RETR_HINT)
  int knock;            // NVS_RETR_KNOCK (timing experiments, results are WRONG): 1 no list work, 2 no TMEM loads, 4 no TMA
  long long* dbg;       // nvs_flat_debug_buffer: 8 counters per CTA (wait profile), or NULL
};

template <int STAGES, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
flat_l2_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                    const Params p) {
  using S = Smem<STAGES, PAIR>;
  constexpr int STAGE_BYTES = S::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* list_d = reinterpret_cast<float*>(sm + S::LIST_D);
  int32_t* list_i = reinterpret_cast<int32_t*>(sm + S::LIST_I);
  float* xn_s = reinterpret_cast<float*>(sm + S::XN);
  const uint32_t bar0 = base + S::BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + ACC_STAGES + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + S::BAR + 8 * (2 * STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cluster of cs CTAs (launch attribute): each owns one query block and TMA-loads 1/cs of every database tile for all
  const uint32_t crank = cluster_ctarank();
  // PAIR: CTAs (2i, 2i+1) of the cluster form a cta_group::2 pair -- one M = 256 MMA over both query blocks, issued by the
  // even CTA (the leader); a CTA holds HALF of the database tile (rows [128 h, 128 h + 128), h = rank & 1), loaded in
  // cs / 2 slices by the CTAs of its parity.  Every operand byte is written to and read from shared memory once per
  // pair instead of once per CTA (64 instead of 96 KB of shared-memory traffic per SM and k-block).
  const int cs = p.cs, b_rows = BN / cs;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  const uint32_t leader_rank = PAIR ? (crank & ~1u) : crank;
  const bool leader = crank == leader_rank;
  const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      // the peers multicast into this stage too: every CTA's (PAIR: every pair's) MMAs must have retired
      mbar_init(empty_bar(s), PAIR ? cs / 2 : cs);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), PAIR ? 8 : 4);  // one arrive per epilogue warp (PAIR: of both CTAs, on the leader's)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {  // both CTAs of the pair issue it, same warp, same slot address
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work = n_mgrp x n_tiles tile steps (one database tile against the cs query blocks of a group), group-major.  Cluster
  // c owns the contiguous range [W c / nc, W (c + 1) / nc): balanced to one tile step, and a cluster stays with one
  // query group for as long as possible -- its per-row lists live on across all those tiles (few restarts, tight
  // thresholds, few lists per query to select from) and its query blocks stay hot in L2.  A SEGMENT = the tiles
  // [t_begin, t_end) of one group inside the range; its lists go to slot (cluster - first cluster of the group).
  const long long W = (long long)p.n_mgrp * p.n_tiles;
  const long long w_begin = W * cluster_id / n_clusters, w_end = W * (cluster_id + 1) / n_clusters;
  const bool prof = p.dbg != nullptr;
  long long w0 = 0, w1 = 0, nkb = 0;
  const long long t_start = prof ? clock64() : 0;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // NVS_RETR_HINT (PAIR): 1 = queries evict-last (they are re-read for every strip), 2 = database rows evict-first
      const uint64_t hint_q = (p.hint & 1) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
      const uint64_t hint_x = (p.hint & 2) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
      for (long long w = w_begin; w < w_end;) {
        const int grp = (int)(w / p.n_tiles), mblk = cs * grp + (int)crank;
        const int t_begin = (int)(w - (long long)grp * p.n_tiles);
        const int t_end = (int)min((long long)p.n_tiles, t_begin + (w_end - w));
        w += t_end - t_begin;
        for (int t = t_begin; t < t_end; ++t) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait_t(empty_bar(stage), phase ^ 1, prof, w0);
            const uint32_t a_dst = base + stage * STAGE_BYTES;
            if (p.knock & 4) {
              if (leader) mbar_arrive(full_bar(stage));
            } else if constexpr (PAIR) {
              // all bytes of the pair (2 x (A + half of B)) are counted on the LEADER's barrier
              const uint32_t fb = map_to_cta(full_bar(stage), leader_rank);
              if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
              tma_load_2d_2sm(a_dst, &tmap_q, fb, kb * BK, mblk * BM, hint_q);  // (rows past Q: zero filled)
              const uint32_t h = crank & 1u, pi = crank >> 1;  // my half of the tile; my slice (256 / cs rows) of it
              const uint32_t b_dst = a_dst + A_BYTES + pi * (uint32_t)(b_rows * BK * 2);
              const int row0 = t * BN + (int)h * (BN / 2) + (int)pi * b_rows;
              if (cs == 2) tma_load_2d_2sm(b_dst, &tmap_x, fb, kb * BK, row0, hint_x);
              else tma_load_2d_2sm_mc(b_dst, &tmap_x, fb, kb * BK, row0, (uint16_t)((0x5555u << h) & cmask), hint_x);
            } else {
              mbar_expect_tx(full_bar(stage), STAGE_BYTES);  // own A + own slice of B + the peers' slices of B
              tma_load_2d(a_dst, &tmap_q, full_bar(stage), kb * BK, mblk * BM);  // (rows past Q: zero filled)
              // my slice (256 / cs rows) of the 256-row database tile, delivered to EVERY CTA of the cluster
              tma_load_2d_mc(a_dst + A_BYTES + crank * (uint32_t)(b_rows * BK * 2), &tmap_x, full_bar(stage), kb * BK,
                             t * BN + (int)crank * b_rows, cmask);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      if (prof) p.dbg[blockIdx.x * 16 + 1] = w0;  // producer: waiting for a free stage
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp runs the loop, one elected lane issues =================
    if (leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      {
        for (long long t = w_begin; t < w_end; ++t) {  // tile steps: the MMA role does not care about segments
          // epilogue (PAIR: of both CTAs) has drained this accumulator
          if constexpr (PAIR) mbar_wait_cluster_t(tempty_bar(acc), acc_phase ^ 1, prof, w1);
          else mbar_wait_t(tempty_bar(acc), acc_phase ^ 1, prof, w1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait_t(full_bar(stage), phase, prof, w0);
            ++nkb;
            tc_fence_after();
            const uint32_t a_addr = base + stage * STAGE_BYTES;
            const uint64_t adesc = make_sdesc(a_addr);
            const uint64_t bdesc = make_sdesc(a_addr + A_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                // advance 16 halves = 32 bytes along K inside the 128-byte swizzle span: +2 in (addr >> 4)
                if constexpr (PAIR)
                  tc_mma_f16_2(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC2,
                               (kb | k) != 0 ? 1u : 0u);
                else
                  tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC,
                             (kb | k) != 0 ? 1u : 0u);
              }
              // frees this stage in EVERY CTA when these MMAs retire
              if constexpr (PAIR) tc_commit_mc2(empty_bar(stage), cmask);
              else tc_commit_mc(empty_bar(stage), cmask);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (elect_one()) {  // accumulator complete -> epilogue (PAIR: of both CTAs)
            if constexpr (PAIR) tc_commit_mc2(tfull_bar(acc), (uint16_t)(0x3u << leader_rank));
            else tc_commit(tfull_bar(acc));
          }
          if (++acc == ACC_STAGES) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
      if (prof && lane == 0) {
        p.dbg[blockIdx.x * 16 + 0] = clock64() - t_start;  // MMA role: total cycles,
        p.dbg[blockIdx.x * 16 + 2] = w0;                   // waiting for operand stages,
        p.dbg[blockIdx.x * 16 + 3] = w1;                   // waiting for a drained accumulator,
        p.dbg[blockIdx.x * 16 + 6] = nkb;                  // k-blocks issued
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: per-row list of the L smallest s~ (append + lazy warp-cooperative compaction) =========
    // A thread owns one query row.  Rows that pass  d < thr && d <= bound  are APPENDED to the row's buffer (two stores);
    // when a buffer could overflow in the next 32 columns the whole warp compacts it to its L smallest entries (sorted)
    // and thr becomes the L-th smallest.  (Keeping a running maximum instead cost a serial O(L) scan per insertion in
    // ONE lane with 31 lanes idle: 30 % of the kernel's cycles, knock-out measurement in profiles/.)
    constexpr int CAPL = S::LMAX;       // buffer capacity = pitch (odd: conflict-free row-per-thread access)
    constexpr int EPL = (CAPL + 31) / 32;
    const int ew = warp - 4;            // == warp % 4: TMEM lane quadrant this warp may read
    const int row = ew * 32 + lane;     // accumulator row (query inside the block)
    const int et = threadIdx.x - 128;   // 0..127
    float* my_d = list_d + row * CAPL;
    int32_t* my_i = list_i + row * CAPL;
    float* warp_d = list_d + ew * 32 * CAPL;
    int32_t* warp_i = list_i + ew * 32 * CAPL;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_bar = 0, w_ld = 0, n_ins = 0, n_cmp = 0;
    // keep the L smallest of row r's n entries, ascending, the earlier entry first among equals (all lanes take part)
    auto compact_row = [&](int r, int n) {
      float* rd = warp_d + r * CAPL;
      int32_t* ri = warp_i + r * CAPL;
      float d[EPL];
      int32_t id[EPL];
      int rank[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int i = lane + 32 * e;
        d[e] = i < n ? rd[i] : INFINITY;
        id[e] = i < n ? ri[i] : -1;
        rank[e] = 0;
      }
      for (int j = 0; j < n; ++j) {
        const float w = rd[j];
#pragma unroll
        for (int e = 0; e < EPL; ++e) rank[e] += (w < d[e] || (w == d[e] && j < lane + 32 * e)) ? 1 : 0;
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        if (lane + 32 * e < n && rank[e] < p.L) {
          rd[rank[e]] = d[e];
          ri[rank[e]] = id[e];
        }
      }
      __syncwarp();
    };
    auto publish = [&](int qrow, float kth) {  // tighten the query's published bound of its k-th smallest s~
      if (kth >= 0.f) atomicMin(reinterpret_cast<int*>(p.gbound + qrow), __float_as_int(kth));
      else atomicMax(reinterpret_cast<unsigned int*>(p.gbound + qrow), __float_as_uint(kth));
    };
    for (long long w = w_begin; w < w_end;) {
      const int grp = (int)(w / p.n_tiles), mblk = cs * grp + (int)crank;
      const int t_begin = (int)(w - (long long)grp * p.n_tiles);
      const int t_end = (int)min((long long)p.n_tiles, t_begin + (w_end - w));
      w += t_end - t_begin;
      // list slot of this segment: clusters sharing the group are consecutive; c0 = the one owning the group's first tile
      const long long g0 = (long long)grp * p.n_tiles;
      int c0 = (int)(g0 * n_clusters / W);
      while (W * (c0 + 1) / n_clusters <= g0) ++c0;
      while (W * c0 / n_clusters > g0) --c0;
      const int slot = cluster_id - c0;
      if (slot < 0 || slot >= p.n_slots) __trap();  // (the host sized the slots for the shortest possible range)
      int cnt = 0;           // entries in my buffer
      float thr = INFINITY;  // L-th smallest at the last compaction: +inf before the first
      const int qrow = mblk * BM + row;
      const bool live = qrow < p.Q;
      const float coef = live ? __ldg(p.qcoef + qrow) : 0.f;
      const float slack = live ? __ldg(p.qslack + qrow) : 0.f;
      for (int t = t_begin; t < t_end; ++t) {
        const int n0 = t * BN;
        // upper bound of the query's k-th smallest s~ published so far (by any strip, any CTA), plus the slack: a row
        // above it cannot be one of the k nearest (see the header); +inf while nothing is published
        float gb2 = live ? __ldcg(p.gbound + qrow) + slack : -INFINITY;
        // stage |x|^2 of this tile (+inf past the end of the shard so padded columns never enter a list)
        float* xs = xn_s + acc * BN;
#pragma unroll
        for (int i = 0; i < BN / 128; ++i) {
          const int n = n0 + et + i * 128;
          xs[et + i * 128] = n < p.N ? __ldg(p.xnorm + n) : INFINITY;
        }
        const long long tb0 = prof ? clock64() : 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (prof) w_bar += clock64() - tb0;
        mbar_wait_t(tfull_bar(acc), acc_phase, prof, w0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < ((p.knock & 2) ? 0 : BN / 32); ++c) {
          // room for 32 more entries in every buffer of the warp
          unsigned need = __ballot_sync(0xffffffffu, cnt > CAPL - 32);
          while (need) {
            const int r = __ffs(need) - 1;
            need &= need - 1;
            compact_row(r, __shfl_sync(0xffffffffu, cnt, r));
            if (lane == r) {
              cnt = p.L;
              thr = my_d[p.L - 1];
              const float kth = my_d[p.k - 1];  // k-th smallest of the strip so far: a valid bound already
              if (live) publish(qrow, kth);
              gb2 = fminf(gb2, kth + slack);
              ++n_cmp;
            }
          }
          float v[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after divergent appends
          const long long tl0 = prof ? clock64() : 0;
          tmem_ld32(taddr + (uint32_t)(c * 32), v);
          if (prof) w_ld += clock64() - tl0;
          const float4* xs4 = reinterpret_cast<const float4*>(xs + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x4 = xs4[j];
            v[4 * j + 0] = fmaf(coef, v[4 * j + 0], x4.x);
            v[4 * j + 1] = fmaf(coef, v[4 * j + 1], x4.y);
            v[4 * j + 2] = fmaf(coef, v[4 * j + 2], x4.z);
            v[4 * j + 3] = fmaf(coef, v[4 * j + 3], x4.w);
          }
          float m[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fminf(fminf(v[4 * j], v[4 * j + 1]), fminf(v[4 * j + 2], v[4 * j + 3]));
          const float mn = fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])), fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
          if (mn < thr && mn <= gb2 && !(p.knock & 1)) {  // rare per lane: some column of this chunk belongs in the list
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (v[j] < thr && v[j] <= gb2) {
                my_d[cnt] = v[j];
                my_i[cnt] = n0 + c * 32 + j;
                ++cnt;
              }
            }
            ++n_ins;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(map_to_cta(tempty_bar(acc), leader_rank));
          else mbar_arrive(tempty_bar(acc));
        }
        if (++acc == ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      // flush this segment's lists: sort every row's buffer (its L smallest, ascending); the k-th smallest
      // entry tightens the published bound
      __syncwarp();
      for (int r = 0; r < 32; ++r) {
        const int n = __shfl_sync(0xffffffffu, cnt, r);
        if (n > 1) compact_row(r, n);
      }
      cnt = cnt < p.L ? cnt : p.L;
      if (live) {
        if (cnt >= p.k) publish(qrow, my_d[p.k - 1]);
        const size_t o = ((size_t)qrow * p.n_slots + slot) * p.L;
        for (int j = 0; j < p.L; ++j) {
          p.cand_d[o + j] = j < cnt ? my_d[j] : INFINITY;
          p.cand_i[o + j] = j < cnt ? my_i[j] : -1;
        }
      }
      __syncwarp();
    }
    if (prof && ew == 0) {
      const long long ins_sum = warp_sum_ll(n_ins);
      if (lane == 0) {
        p.dbg[blockIdx.x * 16 + 4] = w0;                   // epilogue warp 0: waiting for a finished accumulator,
        p.dbg[blockIdx.x * 16 + 5] = clock64() - t_start;  // its total,
        p.dbg[blockIdx.x * 16 + 7] = w_ld;                 // inside tcgen05.ld + wait::ld,
        p.dbg[blockIdx.x * 16 + 8] = w_bar;                // at the named barrier of the four epilogue warps,
        p.dbg[blockIdx.x * 16 + 9] = n_ins;                // 32-column chunks with an append, one row
        p.dbg[blockIdx.x * 16 + 10] = ins_sum;             // and the warp's 32 rows
        p.dbg[blockIdx.x * 16 + 11] = n_cmp;               // compactions of one row
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer may still multicast into / arrive on this CTA's shared memory until it is done
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
  }
}

// ---------------------------------------------------------------- helpers around the GEMM
// fp16 conversion with one power-of-two scale per database / per query row: v -> fp16(v * 2^e).  Values whose scaled
// magnitude falls below the fp16 normal range are flushed to zero HERE (their whole value lands in the residual), so
// the error bound does not depend on how the tensor core treats fp16 subnormals.
__device__ __forceinline__ int scale_exp(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 0;
  int e = 14 - ilogbf(amax);  // amax * 2^e in [2^14, 2^15)
  return e < -60 ? -60 : (e > 60 ? 60 : e);
}
__device__ __forceinline__ __half to_half_scaled(float v, float sc, float inv, float* resid) {
  const float s = v * sc;
  __half h = fabsf(s) < 6.103515625e-05f ? __float2half_rn(0.f) : __float2half_rn(s);  // 2^-14: smallest normal
  *resid = v - __half2float(h) * inv;
  return h;
}
__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {  // v >= 0
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void absmax_kernel(const float* __restrict__ x, long long total, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float a = fabsf(x[i]);
    m = a > m ? a : m;  // NaN never wins
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_pos(out, m);
}

// database rows -> fp16 rows padded to dpad (multiple of 64) + |x|^2; stats[0] = max |x|^2, stats[1] = max |x - x~|^2,
// stats[2] = scale exponent, stats[3] = max |component| (filled by absmax_kernel before).  One warp per row.
__global__ void db_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ xh, float* __restrict__ norms,
                                 float* __restrict__ stats, long long n, int d, int dpad) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int e = scale_exp(stats[3]);
  const float sc = ldexpf(1.f, e), inv = ldexpf(1.f, -e);
  const float* xr = x + row * d;
  __half* br = xh + row * dpad;
  float ss = 0.f, rr = 0.f;
  for (int c = lane; c < dpad; c += 32) {
    const float v = c < d ? xr[c] : 0.f;
    float r;
    br[c] = to_half_scaled(v, sc, inv, &r);
    ss = fmaf(v, v, ss);
    rr = fmaf(r, r, rr);
  }
  ss = warp_sum(ss);
  rr = warp_sum(rr);
  if (lane == 0) {
    norms[row] = ss;
    atomic_max_pos(stats + 0, ss);
    atomic_max_pos(stats + 1, rr);
    if (row == 0) stats[2] = (float)e;
  }
}

// queries -> fp16 rows (own scale per row), |q|^2, the accumulator coefficient and the slack 2 eps of the header.
__global__ void q_to_f16_kernel(const float* __restrict__ q, __half* __restrict__ qh, float* __restrict__ qn,
                                float* __restrict__ qcoef, float* __restrict__ qslack,
                                const float* __restrict__ db_stats, int nq, int d, int dpad) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  const float* qr = q + (size_t)row * d;
  __half* br = qh + (size_t)row * dpad;
  float amax = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float a = fabsf(qr[c]);
    amax = a > amax ? a : amax;
  }
  amax = warp_max(amax);
  const int e = scale_exp(amax);
  const float sc = ldexpf(1.f, e), inv = ldexpf(1.f, -e);
  float ss = 0.f, rr = 0.f;
  for (int c = lane; c < dpad; c += 32) {
    const float v = c < d ? qr[c] : 0.f;
    float r;
    br[c] = to_half_scaled(v, sc, inv, &r);
    ss = fmaf(v, v, ss);
    rr = fmaf(r, r, rr);
  }
  ss = warp_sum(ss);
  rr = warp_sum(rr);
  if (lane == 0) {
    qn[row] = ss;
    const float up = 1.f + 1e-5f;  // the sums above carry ~d * 2^-24 relative rounding
    const float xmax = sqrtf(db_stats[0]) * up, rxmax = sqrtf(db_stats[1]) * up;
    const int ex = (int)db_stats[2];
    const float nq_ = sqrtf(ss) * up, rq = sqrtf(rr) * up;
    qcoef[row] = ldexpf(-2.f, -(e + ex));
    // |q.x - q~.x~| <= |q - q~| |x| + |q~| |x - x~|,  |q~| <= |q| + |q - q~|
    const float e_round = rq * xmax + (nq_ + rq) * rxmax;
    // fp32 accumulation inside the tensor core: products of two fp16 are exact in fp32; every one of the dpad/16
    // accumulations (and the 16-term sum in front of it) is allowed two units of 2^-23 of the magnitude sum
    const float gamma = (float)(dpad / 16 + 32) * 2.384185791015625e-07f;
    const float e_acc = gamma * (nq_ + rq) * (xmax + rxmax);
    // fp32 evaluation of s~ here and of the exact distance in the re-rank (64 ulp of the largest term: generous)
    const float e_f32 = 64.f * 5.9604644775390625e-08f * (ss + xmax * xmax + 2.f * nq_ * xmax);
    const float eps = 2.f * (e_round + e_acc) * 1.001f + e_f32;
    qslack[row] = 2.f * eps;
  }
}

__global__ void fill_inf_kernel(float* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = INFINITY;
}

__device__ __forceinline__ uint32_t ord_key(float f) {  // unsigned order == float order (no NaNs in the lists)
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// One digit of the radix selects below, 256 threads, hist[tid] = occupancy of bin tid: the bin in which the cumulative
// count reaches *s_remaining (parallel prefix sum; a lone thread scanning the 256 bins cost ~7 000 cycles per digit) goes
// into *s_prefix at `shift`, the rank inside that bin into *s_remaining.  Ends with a __syncthreads.
__device__ __forceinline__ void radix_pick_digit(const int* hist, int* s_wsum, uint32_t* s_prefix, int* s_remaining, int shift) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = hist[tid], rem = *s_remaining;
  const uint32_t prefix = *s_prefix;
  int inc = h;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_wsum[w];
  const int cum_incl = base + inc, cum_excl = cum_incl - h;
  // (tid 255 also takes a rank beyond the total: the serial form stopped at the last bin)
  if ((cum_excl < rem && rem <= cum_incl) || (tid == 255 && rem > cum_incl)) {
    *s_remaining = rem - cum_excl;
    *s_prefix = prefix | ((uint32_t)tid << shift);
  }
  __syncthreads();
}

// Per query: tau~ = exact k-th smallest s~ over the n_strips * L listed rows (radix select on the ordered bit
// pattern), candidates = every listed row with s~ <= tau~ + 2 eps, flag = the lists may be missing a row that also
// satisfies that (a strip list full of such rows) or there are more candidates than CAP.
// Sharded search: out_u != nullptr -> only publish, per query, u_j = s~_j + eps for the shard's k smallest listed s~
// (unordered; +inf padding) and return: every u_j is an upper bound of the EXACT distance of a distinct row, so the
// k-th smallest B of the union of all shards' u (kth_union_kernel, after one all_gather of (nq,k) floats) is >= the
// global k-th exact distance.  ext_bound = B: a row of the global top-k has exact distance <= B, hence s~ <= B + eps:
// the cut becomes min(tau~ + 2 eps, B + eps) and a shard re-ranks only what can still matter globally (about k + a few
// rows over ALL shards instead of per shard: the fp32 re-rank, a gather of 16 KB rows, is the part of a sharded
// search that does not shrink otherwise).
__global__ void __launch_bounds__(256) select_candidates_kernel(const float* __restrict__ cand_d,
                                                                const int32_t* __restrict__ cand_i, int n_strips,
                                                                int L, int k, const float* __restrict__ qslack,
                                                                int32_t* __restrict__ sel_i, int32_t* __restrict__ sel_n,
                                                                int32_t* __restrict__ flag_count,
                                                                int32_t* __restrict__ flag_list,
                                                                const float* __restrict__ ext_bound,
                                                                float* __restrict__ out_u) {
  extern __shared__ uint8_t sraw[];
  const int n_cand = n_strips * L;
  float* sd = reinterpret_cast<float*>(sraw);
  int32_t* si = reinterpret_cast<int32_t*>(sraw) + n_cand;
  __shared__ int hist[256];
  __shared__ int s_wsum[8];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_valid, s_cnt, s_flag;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { s_valid = 0; s_cnt = 0; s_flag = 0; s_prefix = 0; }
  __syncthreads();
  int nv = 0;
  for (int i = tid; i < n_cand; i += 256) {
    const float d = cand_d[(size_t)q * n_cand + i];
    const int id = cand_i[(size_t)q * n_cand + i];
    sd[i] = d;
    si[i] = id;
    nv += id >= 0 ? 1 : 0;
  }
  nv = (int)warp_sum((float)nv);  // exact: counts are far below 2^24
  if ((tid & 31) == 0) atomicAdd(&s_valid, nv);
  __syncthreads();
  float T = INFINITY;
  if (s_valid > k) {
    if (tid == 0) s_remaining = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int i = tid; i < n_cand; i += 256) {
        if (si[i] < 0) continue;
        const uint32_t key = ord_key(sd[i]);
        if ((key & mask) == (prefix & mask)) atomicAdd(&hist[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      radix_pick_digit(hist, s_wsum, &s_prefix, &s_remaining, shift);
    }
    T = ord_float(s_prefix) + qslack[q];
  }
  if (out_u != nullptr) {  // first phase of a sharded search: the k smallest s~ (+ eps), unordered
    const float eps = 0.5f * qslack[q];
    const float tau = T - qslack[q];  // +inf: at most k rows listed, publish them all
    float* o = out_u + (size_t)q * k;
    for (int i = tid; i < n_cand; i += 256) {
      if (si[i] >= 0 && sd[i] < tau) {
        const int pos = atomicAdd(&s_cnt, 1);
        if (pos < k) o[pos] = sd[i] + eps;
      }
    }
    __syncthreads();
    const int c = s_cnt < k ? s_cnt : k;  // strictly below tau~: at most k - 1; the rest of the k smallest equal tau~
    for (int j = c + tid; j < k; j += 256) o[j] = tau + eps;
    return;
  }
  if (ext_bound != nullptr) T = fminf(T, ext_bound[q] + 0.5f * qslack[q]);
  // candidates inside the slack (order irrelevant: the re-rank orders by exact distance, then id)
  for (int i = tid; i < n_cand; i += 256) {
    if (si[i] >= 0 && sd[i] <= T) {
      const int pos = atomicAdd(&s_cnt, 1);
      if (pos < CAP) sel_i[(size_t)q * CAP + pos] = si[i];
    }
  }
  // a strip list that is full and lies entirely inside the slack may have dropped a row that belongs there too
  for (int s = tid; s < n_strips; s += 256) {
    bool full_inside = true;
    for (int j = 0; j < L; ++j) {
      const int i = s * L + j;
      if (si[i] < 0 || !(sd[i] <= T)) { full_inside = false; break; }
    }
    if (full_inside) s_flag = 1;
  }
  __syncthreads();
  if (tid == 0) {
    const int cnt = s_cnt;
    sel_n[q] = cnt < CAP ? cnt : CAP;
    if (s_flag || cnt > CAP) flag_list[atomicAdd(flag_count, 1)] = q;
  }
}

// Per query: B = k-th smallest of the parts * k published bounds (radix select, as above; +inf entries sort last).
__global__ void __launch_bounds__(256) kth_union_kernel(const float* __restrict__ u, int parts, int nq, int k,
                                                        float* __restrict__ out_b) {
  extern __shared__ float su[];
  __shared__ int hist[256];
  __shared__ int s_wsum[8];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  const int q = blockIdx.x, tid = threadIdx.x, n = parts * k;
  for (int i = tid; i < n; i += 256) su[i] = u[((size_t)(i / k) * nq + q) * k + (i % k)];
  if (tid == 0) { s_prefix = 0; s_remaining = k; }
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < n; i += 256) {
      const uint32_t key = ord_key(su[i]);
      if ((key & mask) == (prefix & mask)) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    radix_pick_digit(hist, s_wsum, &s_prefix, &s_remaining, shift);
  }
  if (tid == 0) out_b[q] = ord_float(s_prefix);
}

// Exact fp32 q.x of one row by one warp (all lanes return the sum).  The summation order is part of the result: the
// re-rank and the exact scan both use this function, so a row gets the same distance on either path.
__device__ __forceinline__ float exact_dot(const float* __restrict__ sq, const float* __restrict__ xr, int d, int lane) {
  float dot = 0.f;
  if ((d & 3) == 0) {
    // 16 KB row at d = 4096: 128-bit loads, four rows of the loop in flight per lane (a gather: latency, not math)
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    const int n4 = d >> 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int e = lane;
    for (; e + 96 < n4; e += 128) {
      const float4 v0 = __ldg(x4 + e), v1 = __ldg(x4 + e + 32), v2 = __ldg(x4 + e + 64), v3 = __ldg(x4 + e + 96);
      const float4 w0 = q4[e], w1 = q4[e + 32], w2 = q4[e + 64], w3 = q4[e + 96];
      a0 = fmaf(w0.x, v0.x, fmaf(w0.y, v0.y, fmaf(w0.z, v0.z, fmaf(w0.w, v0.w, a0))));
      a1 = fmaf(w1.x, v1.x, fmaf(w1.y, v1.y, fmaf(w1.z, v1.z, fmaf(w1.w, v1.w, a1))));
      a2 = fmaf(w2.x, v2.x, fmaf(w2.y, v2.y, fmaf(w2.z, v2.z, fmaf(w2.w, v2.w, a2))));
      a3 = fmaf(w3.x, v3.x, fmaf(w3.y, v3.y, fmaf(w3.z, v3.z, fmaf(w3.w, v3.w, a3))));
    }
    for (; e < n4; e += 32) {
      const float4 v = __ldg(x4 + e), w = q4[e];
      a0 = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, a0))));
    }
    dot = (a0 + a1) + (a2 + a3);
  } else {
    for (int e = lane; e < d; e += 32) dot = fmaf(sq[e], __ldg(xr + e), dot);
  }
  return warp_sum(dot);
}
__device__ __forceinline__ float exact_d2(float qn, float xn, float dot) {
  return __fmaf_rn(-2.f, dot, __fadd_rn(qn, xn));  // |q|^2 + |x|^2 - 2 q.x, one fixed evaluation order
}
// |q|^2 of a query staged in shared memory by a 256-thread block (same order wherever it is used)
__device__ __forceinline__ float block_query_norm(const float* __restrict__ qr, float* __restrict__ sq, int d,
                                                  float* wsum /* [8] shared */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float ss = 0.f;
  for (int c = tid; c < d; c += 256) {
    const float v = qr[c];
    sq[c] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) wsum[warp] = ss;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += wsum[w];
  __syncthreads();
  return t;
}

// Per query: exact fp32 distances of its candidates (one warp per candidate), then the k smallest ascending
// (ties -> lower id) with global int64 labels; ranks past the candidate count are filled with (+inf, -1).
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ q, const float* __restrict__ x,
                                                     const float* __restrict__ xnorm, const int32_t* __restrict__ sel_i,
                                                     const int32_t* __restrict__ sel_n, int k, int d,
                                                     long long id_offset, float* __restrict__ out_d,
                                                     int64_t* __restrict__ out_i) {
  extern __shared__ __align__(16) float sq[];  // [d] query, then [CAP] distances
  float* dist = sq + d;
  __shared__ float wsum[8];
  const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = sel_n[qi];
  const int32_t* ci = sel_i + (size_t)qi * CAP;
  const float qn = block_query_norm(q + (size_t)qi * d, sq, d, wsum);
  for (int c = warp; c < n; c += 8) {
    const int id = ci[c];
    const float dot = exact_dot(sq, x + (size_t)id * d, d, lane);
    if (lane == 0) dist[c] = exact_d2(qn, xnorm[id], dot);
  }
  __syncthreads();
  if (tid < n) {
    // rank of candidate tid among the n (by distance, then id); ranks < k are written
    const float dm = dist[tid];
    const int im = ci[tid];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float dj = dist[j];
      const int ij = ci[j];
      if (dj < dm || (dj == dm && ij < im)) ++rank;
    }
    if (rank < k) {
      out_d[(size_t)qi * k + rank] = dm;
      out_i[(size_t)qi * k + rank] = (int64_t)im + id_offset;
    }
  }
  for (int r = n + tid; r < k; r += 256) {
    out_d[(size_t)qi * k + r] = INFINITY;
    out_i[(size_t)qi * k + r] = -1;
  }
}

// Flagged queries only: exact distances against every row of the shard.  A row that beats the query's current k-th
// result (by distance, then id) is merged into the sorted result under a per-query lock; the re-ranked candidates
// are the starting point, so in the normal flagged case only the few rows the lists dropped are ever inserted.
constexpr int SCAN_FQ = 4;  // flagged queries staged per pass (shared memory: FQ * d floats)
__device__ __forceinline__ unsigned long long pack_bound(float d, int row) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)row;
}
__global__ void __launch_bounds__(256) exact_scan_kernel(const float* __restrict__ q, const float* __restrict__ x,
                                                         const float* __restrict__ xnorm, long long n_db, int d,
                                                         int k, long long id_offset, int fq,
                                                         const int32_t* __restrict__ flag_count,
                                                         const int32_t* __restrict__ flag_list, int* __restrict__ locks,
                                                         float* out_d, int64_t* out_i) {
  extern __shared__ __align__(16) float sq[];  // [fq][d]
  __shared__ float wsum[8];
  __shared__ float s_qn[SCAN_FQ];
  // the query's current k-th result as ONE 64-bit word (distance bits << 32 | shard-local row): warps update it
  // concurrently, and a reader must never pair one update's distance with another's row
  __shared__ unsigned long long s_bound[SCAN_FQ];
  __shared__ int s_q[SCAN_FQ];
  const int nf = *flag_count;
  if (nf == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long chunk = (n_db + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * chunk, r1 = r0 + chunk < n_db ? r0 + chunk : n_db;
  for (int g0 = 0; g0 < nf; g0 += fq) {
    const int ng = nf - g0 < fq ? nf - g0 : fq;
    __syncthreads();
    for (int j = 0; j < ng; ++j) {
      const int qi = flag_list[g0 + j];
      const float qn = block_query_norm(q + (size_t)qi * d, sq + (size_t)j * d, d, wsum);
      if (tid == 0) {
        s_q[j] = qi;
        s_qn[j] = qn;
        // other CTAs may already be merging rows into this query's result: read the (distance, label) pair under the lock
        while (atomicCAS(&locks[qi], 0, 1) != 0) {}
        __threadfence();
        const float bd0 = *(volatile float*)(out_d + (size_t)qi * k + k - 1);
        const long long bi0 = *(volatile long long*)(out_i + (size_t)qi * k + k - 1);
        __threadfence();
        atomicExch(&locks[qi], 0);
        s_bound[j] = pack_bound(bd0, bi0 < 0 ? -1 : (int)(bi0 - id_offset));
      }
    }
    __syncthreads();
    for (long long row = r0 + warp; row < r1; row += 8) {
      const float xn = xnorm[row];
      const long long gid = row + id_offset;
      for (int j = 0; j < ng; ++j) {
        const float dv = exact_d2(s_qn[j], xn, exact_dot(sq + (size_t)j * d, x + (size_t)row * d, d, lane));
        const unsigned long long bw = *(volatile unsigned long long*)&s_bound[j];
        const float bd = __uint_as_float((unsigned)(bw >> 32));
        const int bi = (int)(unsigned)(bw & 0xffffffffull);
        // beats the current k-th result?  (an empty slot is (+inf, -1): anything finite beats it)
        if (!(dv < bd || (dv == bd && (bi < 0 || row <= (long long)bi)))) continue;
        if (lane == 0) {
          const int qi = s_q[j];
          while (atomicCAS(&locks[qi], 0, 1) != 0) {}
          __threadfence();
          volatile float* D = out_d + (size_t)qi * k;
          volatile long long* I = reinterpret_cast<volatile long long*>(out_i) + (size_t)qi * k;
          bool dup = false;
          for (int t = 0; t < k; ++t) dup |= I[t] == gid;
          if (!dup) {
            int pos = k;
            for (int t = 0; t < k; ++t)
              if (I[t] < 0 || dv < D[t] || (dv == D[t] && gid < I[t])) { pos = t; break; }
            if (pos < k) {
              for (int t = k - 1; t > pos; --t) { D[t] = D[t - 1]; I[t] = I[t - 1]; }
              D[pos] = dv;
              I[pos] = gid;
            }
          }
          // any word ever stored here is some past k-th result of the query: a valid (possibly stale) bound
          const long long ik = I[k - 1];
          *(volatile unsigned long long*)&s_bound[j] = pack_bound(D[k - 1], ik < 0 ? -1 : (int)(ik - id_offset));
          __threadfence();
          atomicExch(&locks[qi], 0);
        }
        __syncwarp();
      }
    }
  }
}

// Merge `parts` sorted (Q,k) lists (e.g. one per GPU shard after the NCCL allgather) into the global top-k.
// Part s starts stride_d floats / stride_i int64s after part s-1 (so D and I may share one gathered buffer).
__global__ void merge_parts_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int parts, int Q,
                                   int k, long long stride_d, long long stride_i, float* __restrict__ out_d,
                                   int64_t* __restrict__ out_i) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int head[16];
  for (int s = 0; s < parts; ++s) head[s] = 0;
  for (int r = 0; r < k; ++r) {
    float bd = INFINITY;
    int64_t bi = -1;
    int bs = -1;
    for (int s = 0; s < parts; ++s) {
      if (head[s] >= k) continue;
      const size_t o = (size_t)q * k + head[s];
      const float d = D[(size_t)s * stride_d + o];
      const int64_t id = I[(size_t)s * stride_i + o];
      if (id < 0) continue;
      if (bs < 0 || d < bd || (d == bd && id < bi)) {
        bd = d; bi = id; bs = s;
      }
    }
    out_d[(size_t)q * k + r] = bs >= 0 ? bd : INFINITY;
    out_i[(size_t)q * k + r] = bs >= 0 ? bi : -1;
    if (bs >= 0) ++head[bs];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 [rows][dpad] row-major, box = 64 columns (128 bytes) x box_rows, SWIZZLE_128B, OOB rows -> zeros
static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t dpad, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[2] = {dpad, rows};
  cuuint64_t strides[1] = {dpad * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }
static inline int dpad_of(int d) { return (d + BK - 1) / BK * BK; }

struct Layout {
  int dpad, n_mblk, n_tiles, n_strips /* list slots per query */, qpad, L, stages, cs, pair;
  size_t off_qb, off_qn, off_qc, off_qs, off_gb, off_cd, off_ci, off_si, off_sn, off_fl, off_lk, total;
};

static int forced_stages() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NVS_RETR_STAGES");
    v = e ? atoi(e) : 0;
    if (v < 2 || v > 6) v = 0;
  }
  return v;
}

// cta_group::2 pairs (default) or one MMA per CTA (NVS_RETR_PAIR=0, the kernel of the earlier rounds, kept for A/B runs)
static int pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NVS_RETR_PAIR");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v;
}

// CTAs per cluster.  Every CTA of a cluster works on its own 128-query block against the SAME database tile and loads
// 1/cs of it for all (TMA multicast), so a k-block costs 16 KB (A) + 32/cs KB (B slice) of L2 -> SM traffic per CTA:
// 32 KB at cs = 2, 24 KB at 4, 20 KB at 8 against 512 cycles of MMA time.  The kernel runs against the board's power
// limit (SM clock 1.0-1.5 GHz under load), so bytes moved per FLOP decide the speed: measured at 10k x 1M x 4096 on
// cta_group::2 pairs 977 / 1037 / 1070 TFLOP/s with cs = 2 / 4 / 8 (SM clock 1.03 / 1.40 / 1.43 GHz) although only
// 132 / 120 of the 148 SMs can hold clusters of 4 / 8.  NVS_RETR_CLUSTER = 2 / 4 / 8 for A/B runs.
static int cluster_size_for(int n_mblk) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("NVS_RETR_CLUSTER");
    const int v = e ? atoi(e) : 0;
    forced = (v == 2 || v == 4 || v == 8) ? v : 0;
  }
  if (forced) return forced;
  return n_mblk >= 24 ? 8 : (n_mblk >= 8 ? 4 : 2);  // few query blocks: small clusters waste fewer CTAs on padding
}

static Layout make_layout(long long n_db, int nq, int d, int k) {
  Layout L;
  L.dpad = dpad_of(d);
  L.n_mblk = (nq + BM - 1) / BM;
  L.qpad = L.n_mblk * BM;
  L.n_tiles = (int)((n_db + BN - 1) / BN);
  // per-row list length: k plus room for rows inside the slack (a list that fills up with them flags the query for
  // the exact scan).  NVS_RETR_STAGES forces the ring depth for A/B runs.
  int want_l = k + (k / 4 > 6 ? k / 4 : 6);
  L.pair = pair_mode();
  const int fs = forced_stages();
  // The epilogue appends to per-row buffers of LMAX entries and compacts them to L when fewer than 32 slots are left:
  // LMAX >= L + 32 is required, L + 64 keeps compactions rare.  PAIR (32 KiB stages): 5 stages -> 63, 4 -> 95, 3 -> 127;
  // one MMA per CTA (48 KiB stages): 3 -> 79, 2 -> 127.
  int lmax;
  if (L.pair) {
    auto lmax_of = [](int st) { return st == 5 ? Smem<5, true>::LMAX : (st == 4 ? Smem<4, true>::LMAX : Smem<3, true>::LMAX); };
    L.stages = want_l + 64 <= lmax_of(4) ? 4 : 3;
    if (fs >= 3 && fs <= 5 && lmax_of(fs) >= k + 32) L.stages = fs;
    lmax = lmax_of(L.stages);
  } else {
    auto lmax_of = [](int st) { return st == 3 ? Smem<3, false>::LMAX : Smem<2, false>::LMAX; };
    L.stages = want_l + 32 <= lmax_of(3) ? 3 : 2;
    if (fs >= 2 && fs <= 3 && lmax_of(fs) >= k + 32) L.stages = fs;
    lmax = lmax_of(L.stages);
  }
  lmax -= 32;
  L.L = want_l < lmax ? want_l : lmax;
  // List slots per query: the clusters that can share one query-block group.  The n_mgrp x n_tiles tile steps are cut
  // into one contiguous range per cluster (group-major); a group of n_tiles steps meets at most
  // ceil(n_tiles / shortest range) + 1 ranges.  (10k queries, clusters of 8 on 148 SMs: 18 clusters, 10 groups -> 3.)
  L.cs = cluster_size_for(L.n_mblk);
  const int n_mgrp = (L.n_mblk + L.cs - 1) / L.cs;
  {
    const long long W = (long long)n_mgrp * L.n_tiles;
    const int sms = nvs_sm_count();
    long long nc = (sms > 0 ? sms : 148) / L.cs;  // (no device: size the workspace for a B200)
    if (nc < 1) nc = 1;
    if (nc > W) nc = W;
    const long long per = W / nc;  // >= 1
    L.n_strips = (int)((L.n_tiles + per - 1) / per) + 1;
  }
  size_t o = 0;
  L.off_qb = o; o += al256((size_t)L.qpad * L.dpad * 2);
  L.off_qn = o; o += al256((size_t)L.qpad * 4);
  L.off_qc = o; o += al256((size_t)L.qpad * 4);
  L.off_qs = o; o += al256((size_t)L.qpad * 4);
  L.off_gb = o; o += al256((size_t)L.qpad * 4);
  L.off_cd = o; o += al256((size_t)L.qpad * L.n_strips * L.L * 4);
  L.off_ci = o; o += al256((size_t)L.qpad * L.n_strips * L.L * 4);
  L.off_si = o; o += al256((size_t)nq * CAP * 4);
  L.off_sn = o; o += al256((size_t)nq * 4);
  L.off_fl = o; o += al256((size_t)(nq + 1) * 4);   // flag_count, then flag_list
  L.off_lk = o; o += al256((size_t)nq * 4);
  L.total = o;
  return L;
}

static long long* g_dbg = nullptr;

template <int STAGES, bool PAIR>
static int launch_gemm(const CUtensorMap& mq, const CUtensorMap& mx, const Params& p, long long n_steps, cudaStream_t st) {
  auto kern = flat_l2_topk_kernel<STAGES, PAIR>;
  NVS_OPT_IN_SMEM(kern, (Smem<STAGES, PAIR>::BYTES));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = Smem<STAGES, PAIR>::BYTES;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // clusters that can be resident at once (one CTA per SM; a cluster lives inside one GPC): per device and size
  static int resident[64][9] = {{0}};
  int dev = 0;
  cudaGetDevice(&dev);
  int n_res = __atomic_load_n(&resident[dev & 63][p.cs], __ATOMIC_ACQUIRE);
  if (n_res == 0) {
    cfg.gridDim = dim3((unsigned)(nvs_sm_count() / p.cs * p.cs));
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n_res, kern, &cfg);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    if (n_res <= 0) return NVS_ERR_CUDA;
    __atomic_store_n(&resident[dev & 63][p.cs], n_res, __ATOMIC_RELEASE);
  }
  const int sms = nvs_sm_count();
  if (n_res > sms / p.cs) n_res = sms / p.cs;
  cfg.gridDim = dim3((unsigned)(p.cs * (n_steps < n_res ? (int)n_steps : n_res)));  // never more clusters than tile steps
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mq, mx, p);
  if (e != cudaSuccess) return nvs_set_cuda_error(e);
  return NVS_OK;
}

}  // namespace rt
}  // namespace nvs

using namespace nvs;
using namespace nvs::rt;

extern "C" int32_t nvs_flat_padded_dim(int32_t d) { return d > 0 ? dpad_of(d) : 0; }

extern "C" int nvs_flat_prepare(const float* x, int64_t n, int32_t d, void* x_f16, float* norms, float* stats,
                                void* stream) {
  if (!x || !x_f16 || !norms || !stats || n <= 0 || d <= 0) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dp = dpad_of(d);
  cudaError_t e = cudaMemsetAsync(stats, 0, 4 * sizeof(float), st);
  if (e != cudaSuccess) return nvs_set_cuda_error(e);
  const long long total = (long long)n * d;
  const long long want = (total + 255) / 256;
  absmax_kernel<<<(unsigned)(want < 4096 ? want : 4096), 256, 0, st>>>(x, total, stats + 3);
  NVS_CHECK_LAUNCH();
  const long long blocks = (n + 7) / 8;
  db_to_f16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, (__half*)x_f16, norms, stats, n, d, dp);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int32_t nvs_flat_max_k(void) { return KMAX; }

// host-only: list slots per query and cluster size the search of this shape will use (tests check the work split
// against them; without a device the layout is sized for the 148 SMs of a B200)
extern "C" int32_t nvs_flat_list_slots(int64_t n_db, int32_t nq, int32_t d, int32_t k, int32_t* cluster_size) {
  if (n_db <= 0 || nq <= 0 || d <= 0 || k <= 0 || k > KMAX) return 0;
  const Layout L = make_layout(n_db, nq, d, k);
  if (cluster_size) *cluster_size = L.cs;
  return L.n_strips;
}

// debugging aid (tools/retr_waits.py): 16 int64 per CTA written by every following GEMM launch -- MMA-role cycles, producer
// wait, MMA wait for operands, MMA wait for an accumulator, epilogue wait, epilogue cycles, k-blocks, unused
extern "C" void nvs_flat_debug_buffer(long long* dev_buf) { nvs::rt::g_dbg = dev_buf; }

extern "C" size_t nvs_flat_search_workspace_bytes(int64_t n_db, int32_t nq, int32_t d, int32_t k) {
  if (n_db <= 0 || nq <= 0 || d <= 0 || k <= 0 || k > KMAX) return 0;
  return make_layout(n_db, nq, d, k).total;
}

// phase 1: query conversion, GEMM + per-row lists [, u = tau~ + eps per query]; phase 2: selection [inside the global
// bound], fp32 re-rank, exact scan of flagged queries.  The workspace carries the lists between the phases.
static int flat_search_impl(const float* db, const void* db_f16, const float* db_norms, const float* db_stats,
                            int64_t n_db, const float* q, int32_t nq, int32_t d, int32_t k, int64_t id_offset,
                            float* out_D, int64_t* out_I, void* workspace, size_t workspace_bytes, void* ev_gemm_start,
                            void* ev_gemm_stop, void* stream, int phase /* 0 both, 1 begin, 2 end */, float* out_u,
                            const float* ext_bound) {
  if (!db || !db_f16 || !db_norms || !db_stats || !q || !workspace) return NVS_ERR_ARG;
  if (phase != 1 && (!out_D || !out_I)) return NVS_ERR_ARG;
  if (phase == 1 && !out_u) return NVS_ERR_ARG;
  if (n_db <= 0 || nq <= 0 || d <= 0 || k <= 0) return NVS_ERR_ARG;
  if (k > KMAX) return NVS_ERR_UNSUPPORTED;   // per-row lists live in shared memory
  if (k > n_db) return NVS_ERR_ARG;
  if (n_db > 0x7fffffffLL - BN) return NVS_ERR_UNSUPPORTED;
  if ((size_t)(d + CAP) * 4 > 96 * 1024) return NVS_ERR_UNSUPPORTED;  // query row staged in shared memory
  const Layout L = make_layout(n_db, nq, d, k);
  if (workspace_bytes < L.total) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* qb = reinterpret_cast<__half*>(ws + L.off_qb);
  float* qn = reinterpret_cast<float*>(ws + L.off_qn);
  float* qc = reinterpret_cast<float*>(ws + L.off_qc);
  float* qs = reinterpret_cast<float*>(ws + L.off_qs);
  float* gb = reinterpret_cast<float*>(ws + L.off_gb);
  float* cd = reinterpret_cast<float*>(ws + L.off_cd);
  int32_t* ci = reinterpret_cast<int32_t*>(ws + L.off_ci);
  int32_t* si = reinterpret_cast<int32_t*>(ws + L.off_si);
  int32_t* sn = reinterpret_cast<int32_t*>(ws + L.off_sn);
  int32_t* fl = reinterpret_cast<int32_t*>(ws + L.off_fl);
  int* lk = reinterpret_cast<int*>(ws + L.off_lk);
  const int n_cand = L.n_strips * L.L;
  if (n_cand > MAX_CAND) return NVS_ERR_UNSUPPORTED;  // the selection kernel holds a query's lists in shared memory
  NVS_OPT_IN_SMEM(select_candidates_kernel, 96 * 1024);

  if (phase != 2) {
    // queries -> fp16 (rows past nq stay zero: they only feed padded accumulator rows that are never flushed);
    // flag counter + list and the scan locks (adjacent in the workspace) start at zero
    cudaError_t e = cudaMemsetAsync(qb, 0, (size_t)L.qpad * L.dpad * 2, st);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    e = cudaMemsetAsync(fl, 0, (L.off_lk - L.off_fl) + al256((size_t)nq * 4), st);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    q_to_f16_kernel<<<(nq + 7) / 8, 256, 0, st>>>(q, qb, qn, qc, qs, db_stats, nq, d, L.dpad);
    NVS_CHECK_LAUNCH();
    fill_inf_kernel<<<(L.qpad + 255) / 256, 256, 0, st>>>(gb, L.qpad);
    NVS_CHECK_LAUNCH();

    CUtensorMap mq, mx;
    if (make_map(&mq, qb, (uint64_t)L.qpad, (uint64_t)L.dpad, BM) != NVS_OK) return NVS_ERR_CUDA;
    if (make_map(&mx, db_f16, (uint64_t)n_db, (uint64_t)L.dpad, BN / L.cs) != NVS_OK) return NVS_ERR_CUDA;  // tile slices

    Params p;
    p.xnorm = db_norms; p.qcoef = qc; p.qslack = qs; p.gbound = gb; p.cand_d = cd; p.cand_i = ci;
    p.Q = nq; p.N = (int)n_db; p.kblocks = L.dpad / BK;
    p.k = k; p.L = L.L;
    p.n_mblk = L.n_mblk; p.n_slots = L.n_strips; p.n_tiles = L.n_tiles;
    p.cs = L.cs;
    p.dbg = g_dbg;
    {
      const char* e = getenv("NVS_RETR_KNOCK");
      p.knock = e ? atoi(e) : 0;
      const char* h = getenv("NVS_RETR_HINT");
      p.hint = h ? atoi(h) : 2;  // measured at 10k x 1M x 4096: 1061 / 1102 / 1111 / 1103 TFLOP/s with hints 0 / 1 / 2 / 3
    }
    p.n_mgrp = (L.n_mblk + L.cs - 1) / L.cs;
    const long long n_units = (long long)p.n_mgrp * L.n_tiles;  // tile steps, cut into one range per cluster
    // empty list slots (a group met by fewer clusters than slots) must read as empty: labels -1
    if (cudaMemsetAsync(ci, 0xFF, (size_t)L.qpad * L.n_strips * L.L * 4, st) != cudaSuccess) return NVS_ERR_CUDA;

    if (ev_gemm_start) cudaEventRecord(static_cast<cudaEvent_t>(ev_gemm_start), st);
    int rc;
    if (L.pair)
      rc = L.stages == 5 ? launch_gemm<5, true>(mq, mx, p, n_units, st)
                         : (L.stages == 4 ? launch_gemm<4, true>(mq, mx, p, n_units, st)
                                          : launch_gemm<3, true>(mq, mx, p, n_units, st));
    else
      rc = L.stages == 3 ? launch_gemm<3, false>(mq, mx, p, n_units, st) : launch_gemm<2, false>(mq, mx, p, n_units, st);
    if (rc != NVS_OK) return rc;
    if (ev_gemm_stop) cudaEventRecord(static_cast<cudaEvent_t>(ev_gemm_stop), st);
    if (phase == 1) {
      select_candidates_kernel<<<nq, 256, (size_t)n_cand * 8, st>>>(cd, ci, L.n_strips, L.L, k, qs, si, sn, fl, fl + 1,
                                                                    nullptr, out_u);
      NVS_CHECK_LAUNCH();
      return NVS_OK;
    }
  }

  NVS_OPT_IN_SMEM(rerank_kernel, 96 * 1024);
  NVS_OPT_IN_SMEM(exact_scan_kernel, 96 * 1024);
  select_candidates_kernel<<<nq, 256, (size_t)n_cand * 8, st>>>(cd, ci, L.n_strips, L.L, k, qs, si, sn, fl, fl + 1,
                                                                ext_bound, nullptr);
  NVS_CHECK_LAUNCH();
  rerank_kernel<<<nq, 256, (size_t)(d + CAP) * 4, st>>>(q, db, db_norms, si, sn, k, d, (long long)id_offset, out_D, out_I);
  NVS_CHECK_LAUNCH();
  int fq = 16384 / d;
  fq = fq < 1 ? 1 : (fq > SCAN_FQ ? SCAN_FQ : fq);
  exact_scan_kernel<<<2 * nvs_sm_count(), 256, (size_t)fq * d * 4, st>>>(q, db, db_norms, (long long)n_db, d, k,
                                                                        (long long)id_offset, fq, fl, fl + 1, lk,
                                                                        out_D, out_I);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_flat_search(const float* db, const void* db_f16, const float* db_norms, const float* db_stats,
                               int64_t n_db, const float* q, int32_t nq, int32_t d, int32_t k, int64_t id_offset,
                               float* out_D, int64_t* out_I, void* workspace, size_t workspace_bytes,
                               void* ev_gemm_start, void* ev_gemm_stop, void* stream) {
  return flat_search_impl(db, db_f16, db_norms, db_stats, n_db, q, nq, d, k, id_offset, out_D, out_I, workspace,
                          workspace_bytes, ev_gemm_start, ev_gemm_stop, stream, 0, nullptr, nullptr);
}

extern "C" int nvs_flat_search_begin(const float* db, const void* db_f16, const float* db_norms, const float* db_stats,
                                     int64_t n_db, const float* q, int32_t nq, int32_t d, int32_t k, float* out_bound,
                                     void* workspace, size_t workspace_bytes, void* ev_gemm_start, void* ev_gemm_stop,
                                     void* stream) {
  return flat_search_impl(db, db_f16, db_norms, db_stats, n_db, q, nq, d, k, 0, nullptr, nullptr, workspace,
                          workspace_bytes, ev_gemm_start, ev_gemm_stop, stream, 1, out_bound, nullptr);
}

extern "C" int nvs_flat_bound_merge(const float* bounds, int32_t parts, int32_t nq, int32_t k, float* out_bound,
                                    void* stream) {
  if (!bounds || !out_bound || parts <= 0 || parts > 64 || nq <= 0 || k <= 0 || k > KMAX) return NVS_ERR_ARG;
  kth_union_kernel<<<nq, 256, (size_t)parts * k * 4, static_cast<cudaStream_t>(stream)>>>(bounds, parts, nq, k, out_bound);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_flat_search_end(const float* db, const void* db_f16, const float* db_norms, const float* db_stats,
                                   int64_t n_db, const float* q, int32_t nq, int32_t d, int32_t k, int64_t id_offset,
                                   const float* global_bound, float* out_D, int64_t* out_I, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!global_bound) return NVS_ERR_ARG;
  return flat_search_impl(db, db_f16, db_norms, db_stats, n_db, q, nq, d, k, id_offset, out_D, out_I, workspace,
                          workspace_bytes, nullptr, nullptr, stream, 2, nullptr, global_bound);
}

extern "C" int nvs_topk_merge(const float* D_parts, const int64_t* I_parts, int32_t parts, int32_t nq, int32_t k,
                              int64_t part_stride_d, int64_t part_stride_i, float* out_D, int64_t* out_I,
                              void* stream) {
  if (!D_parts || !I_parts || !out_D || !out_I || parts <= 0 || parts > 16 || nq <= 0 || k <= 0) return NVS_ERR_ARG;
  const long long sd = part_stride_d > 0 ? part_stride_d : (long long)nq * k;
  const long long si = part_stride_i > 0 ? part_stride_i : (long long)nq * k;
  merge_parts_kernel<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream>>>(D_parts, I_parts, parts, nq, k, sd, si, out_D,
                                                                         out_I);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
