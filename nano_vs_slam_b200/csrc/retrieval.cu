// Exact L2 top-k retrieval = faiss.IndexFlatL2.add / search (src/evaluation/global_descriptor.py:55-60).
//
//   d2(q, x) = |q|^2 + |x|^2 - 2 q.x ,  k smallest, ascending, int64 labels.
//
// The Q x N x D contraction is the one genuinely tensor-core shaped piece of the hot path
// (10k x 1M x 4096 = 8.2e13 FLOP): it runs as a warp-specialised tcgen05 GEMM
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages bf16 tiles of queries (A, 128 x 64) and
//     database rows (B, 256 x 64) into a 4-deep shared-memory ring.  The kernel is bound by the L2->SM
//     fabric (~42 B/cycle/SM chip-wide), so CTAs run as clusters of two that work on the same database
//     tiles for two different query blocks: each CTA fetches HALF of every B tile and TMA-multicasts it
//     into both CTAs' shared memory (48 -> 32 KB of L2 reads per CTA per stage),
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) into a
//     double-buffered 128 x 256 fp32 accumulator in TMEM (2 x 256 columns),
//   * four epilogue warps read the accumulator with tcgen05.ld, form |x|^2 - 2 q.x and keep a running
//     per-row top-k in shared memory -- the score matrix never leaves the SM,
// followed by a small candidate merge and an fp32 re-rank with the literal formula above, so the
// reported distances and the order are those of an fp32 implementation (bf16 is only a filter).
#include <cuda.h>
#include <cuda_bf16.h>
#include <limits.h>

#include "common.cuh"

namespace nvs {
namespace rt {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;       // 16 KiB
constexpr int B_BYTES = BN * BK * 2;       // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int THREADS = 256;                // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 epilogue
constexpr int KMAX = 31;                    // per-row list length limit (odd pitch, conflict free)

// shared memory map (dynamic, 1024-byte aligned base)
constexpr int SM_TILES = 0;
constexpr int SM_LIST_D = STAGES * STAGE_BYTES;                // float [128][KP]
constexpr int SM_LIST_I = SM_LIST_D + BM * KMAX * 4;           // int   [128][KP]
constexpr int SM_XN = SM_LIST_I + BM * KMAX * 4;               // float [2][256]
constexpr int SM_BAR = SM_XN + ACC_STAGES * BN * 4;            // mbarriers
constexpr int SM_TOTAL = SM_BAR + 256;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;                    // slack for manual 1024 B alignment

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
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
// instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Params {
  const float* xnorm;   // [N]   |x|^2 (fp32, from the fp32 rows)
  float* gbound;        // [Qpad] running upper bound of each query's k-th smallest (|x|^2 - 2 q.x), init +inf
  float* cand_d;        // [Qpad][n_strips][k]  approximate |x|^2 - 2 q.x
  int32_t* cand_i;      // [Qpad][n_strips][k]  row index inside this shard (-1 = empty)
  int Q, N, kblocks;    // kblocks = Dpad / 64
  int k, kp;            // list length and its (odd) pitch
  int n_mblk, n_strips, tiles_per_strip, n_tiles;
  int n_mpair;          // ceil(n_mblk / 2): a cluster of two CTAs owns query blocks (2*pair, 2*pair + 1)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
flat_l2_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                    const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* list_d = reinterpret_cast<float*>(sm + SM_LIST_D);
  int32_t* list_i = reinterpret_cast<int32_t*>(sm + SM_LIST_I);
  float* xn_s = reinterpret_cast<float*>(sm + SM_XN);
  const uint32_t bar0 = base + SM_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + ACC_STAGES + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_BAR + 8 * (2 * STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0 / 1 inside the CTA pair
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);  // the peer multicasts into this stage too: both CTAs' MMAs must have retired
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_units = p.n_mpair * p.n_strips;  // (strip, query-block pair) units, round-robin over clusters

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters) {
        const int strip = u / p.n_mpair, mblk = 2 * (u - strip * p.n_mpair) + (int)crank;
        const int t_begin = strip * p.tiles_per_strip;
        const int t_end = min(p.n_tiles, t_begin + p.tiles_per_strip);
        for (int t = t_begin; t < t_end; ++t) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t a_dst = base + SM_TILES + stage * STAGE_BYTES;
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);  // own A + own half of B + the peer's half of B
            tma_load_2d(a_dst, &tmap_q, full_bar(stage), kb * BK, mblk * BM);  // (rows past Q: zero filled)
            // my half (128 rows) of the 256-row database tile, delivered to BOTH CTAs of the pair
            tma_load_2d_mc(a_dst + A_BYTES + crank * (B_BYTES / 2), &tmap_x, full_bar(stage), kb * BK,
                           t * BN + (int)crank * (BN / 2), (uint16_t)0x3);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters) {
        const int strip = u / p.n_mpair;
        const int t_begin = strip * p.tiles_per_strip;
        const int t_end = min(p.n_tiles, t_begin + p.tiles_per_strip);
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = base + SM_TILES + stage * STAGE_BYTES;
            const uint64_t adesc = make_sdesc(a_addr);
            const uint64_t bdesc = make_sdesc(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle span: +2 in (addr >> 4)
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC,
                          (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit_mc(empty_bar(stage), (uint16_t)0x3);  // frees this stage in BOTH CTAs when these MMAs retire
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
          if (++acc == ACC_STAGES) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: running per-row top-k =================
    const int ew = warp - 4;            // == warp % 4: TMEM lane quadrant this warp may read
    const int row = ew * 32 + lane;     // accumulator row (query inside the block)
    const int et = threadIdx.x - 128;   // 0..127
    float* my_d = list_d + row * p.kp;
    int32_t* my_i = list_i + row * p.kp;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = cluster_id; u < n_units; u += n_clusters) {
      const int strip = u / p.n_mpair, mblk = 2 * (u - strip * p.n_mpair) + (int)crank;
      const int t_begin = strip * p.tiles_per_strip;
      const int t_end = min(p.n_tiles, t_begin + p.tiles_per_strip);
      for (int j = 0; j < p.k; ++j) {
        my_d[j] = INFINITY;
        my_i[j] = -1;
      }
      float thr = INFINITY;
      int pmax = 0;
      const int qrow = mblk * BM + row;
      // upper bound on this query's final k-th distance, published by units that finished earlier
      const float gbound = qrow < p.Q ? __ldcg(p.gbound + qrow) : INFINITY;
      for (int t = t_begin; t < t_end; ++t) {
        const int n0 = t * BN;
        // stage |x|^2 of this tile (+inf past the end of the shard so padded columns never enter a list)
        float* xs = xn_s + acc * BN;
#pragma unroll
        for (int i = 0; i < BN / 128; ++i) {
          const int n = n0 + et + i * 128;
          xs[et + i * 128] = n < p.N ? __ldg(p.xnorm + n) : INFINITY;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after divergent list insertions
          tmem_ld32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = fmaf(-2.f, v[j], xs[c * 32 + j]);
            // two filters: the row's own k-th best so far (strict) and the bound published by strips that
            // already finished for this query (<=, ties must survive); insertions are rare after warm-up
            if (d < thr && d <= gbound) {
              my_d[pmax] = d;
              my_i[pmax] = n0 + c * 32 + j;
              thr = my_d[0];
              pmax = 0;
              for (int q = 1; q < p.k; ++q) {
                const float w = my_d[q];
                if (w > thr) {
                  thr = w;
                  pmax = q;
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      // flush this (query block, strip) candidate list; a full list tightens the published bound
      if (qrow < p.Q) {
        if (thr < INFINITY) {
          if (thr >= 0.f) atomicMin(reinterpret_cast<int*>(p.gbound + qrow), __float_as_int(thr));
          else atomicMax(reinterpret_cast<unsigned int*>(p.gbound + qrow), __float_as_uint(thr));
        }
        const size_t o = ((size_t)qrow * p.n_strips + strip) * p.k;
        for (int j = 0; j < p.k; ++j) {
          p.cand_d[o + j] = my_d[j];
          p.cand_i[o + j] = my_i[j];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer may still multicast into / arrive on this CTA's shared memory until it is done
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- helpers around the GEMM
// fp32 rows -> bf16 rows padded to dpad (multiple of 64) + fp32 squared norms.  One warp per row.
__global__ void to_bf16_norm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                    float* __restrict__ norms, long long n, int d, int dpad) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + row * d;
  __nv_bfloat16* br = xb + row * dpad;
  float ss = 0.f;
  for (int c = lane; c < dpad; c += 32) {
    const float v = c < d ? xr[c] : 0.f;
    ss = fmaf(v, v, ss);
    br[c] = __float2bfloat16_rn(v);
  }
  ss = warp_sum(ss);
  if (lane == 0) norms[row] = ss;
}

__global__ void fill_inf_kernel(float* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = INFINITY;
}

// Per query: keep the kk smallest of n_cand (approximate distance, index) pairs, ascending (d, i).
__global__ void __launch_bounds__(256) select_candidates_kernel(const float* __restrict__ cand_d,
                                                                const int32_t* __restrict__ cand_i, int n_cand,
                                                                int kk, float* __restrict__ out_d,
                                                                int32_t* __restrict__ out_i) {
  extern __shared__ uint8_t sraw[];
  float* sd = reinterpret_cast<float*>(sraw);
  int32_t* si = reinterpret_cast<int32_t*>(sraw) + n_cand;
  __shared__ float rd[8];
  __shared__ int ri[8], rp[8];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_cand; i += 256) {
    sd[i] = cand_d[(size_t)q * n_cand + i];
    si[i] = cand_i[(size_t)q * n_cand + i];
  }
  __syncthreads();
  for (int r = 0; r < kk; ++r) {
    float bd = INFINITY;
    int bi = INT_MAX, bp = -1;
    for (int i = tid; i < n_cand; i += 256) {
      const float d = sd[i];
      const int id = si[i];
      if (id >= 0 && (d < bd || (d == bd && id < bi))) {
        bd = d; bi = id; bp = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (op >= 0 && (bp < 0 || od < bd || (od == bd && oi < bi))) {
        bd = od; bi = oi; bp = op;
      }
    }
    if (lane == 0) { rd[warp] = bd; ri[warp] = bi; rp[warp] = bp; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (rp[w] >= 0 && (rp[0] < 0 || rd[w] < rd[0] || (rd[w] == rd[0] && ri[w] < ri[0]))) {
          rd[0] = rd[w]; ri[0] = ri[w]; rp[0] = rp[w];
        }
      out_d[(size_t)q * kk + r] = rp[0] >= 0 ? rd[0] : INFINITY;
      out_i[(size_t)q * kk + r] = rp[0] >= 0 ? ri[0] : -1;
      if (rp[0] >= 0) si[rp[0]] = -1;  // consume
    }
    __syncthreads();
  }
}

// Per query: exact fp32 d2 = |q|^2 + |x|^2 - 2 q.x for kk candidates (one warp per candidate), then the k
// smallest ascending (ties -> lower id) with global int64 labels.
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ q, const float* __restrict__ x,
                                                     const float* __restrict__ xnorm, const int32_t* __restrict__ ci,
                                                     int kk, int k, int d, long long id_offset,
                                                     float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  extern __shared__ __align__(16) float sq[];  // [d] query, then [kk] distances
  float* dist = sq + d;
  __shared__ float qn_s;
  const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qr = q + (size_t)qi * d;
  float ss = 0.f;
  for (int c = tid; c < d; c += 256) {
    const float v = qr[c];
    sq[c] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  __shared__ float wsum[8];
  if (lane == 0) wsum[warp] = ss;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    qn_s = t;
  }
  __syncthreads();
  for (int c = warp; c < kk; c += 8) {
    const int id = ci[(size_t)qi * kk + c];
    float dv = INFINITY;
    if (id >= 0) {
      const float* xr = x + (size_t)id * d;
      float dot = 0.f;
      if ((d & 3) == 0) {
        // 16 KB row per candidate at d = 4096: 128-bit loads, four rows of the loop in flight per lane (the
        // kernel is a gather of kk random rows per query: latency, not arithmetic)
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
      dot = warp_sum(dot);
      dv = qn_s + xnorm[id] - 2.f * dot;
    }
    if (lane == 0) dist[c] = dv;
  }
  __syncthreads();
  if (tid < kk) {
    // rank of candidate tid among the kk (stable by (distance, id)); ranks < k are written
    const float dm = dist[tid];
    const int im = ci[(size_t)qi * kk + tid];
    int rank = 0;
    for (int j = 0; j < kk; ++j) {
      const float dj = dist[j];
      const int ij = ci[(size_t)qi * kk + j];
      if (ij >= 0 && (dj < dm || (dj == dm && (ij < im || (ij == im && j < tid))))) ++rank;
    }
    if (im >= 0 && rank < k) {
      out_d[(size_t)qi * k + rank] = dm;
      out_i[(size_t)qi * k + rank] = (int64_t)im + id_offset;
    }
  }
}

// Merge `parts` sorted (Q,k) lists (e.g. one per GPU shard after the NCCL allgather) into the global top-k.
__global__ void merge_parts_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int parts, int Q,
                                   int k, float* __restrict__ out_d, int64_t* __restrict__ out_i) {
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
      const size_t o = ((size_t)s * Q + q) * k + head[s];
      const float d = D[o];
      const int64_t id = I[o];
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

// bf16 [rows][dpad] row-major, box = 64 columns (128 bytes) x box_rows, SWIZZLE_128B, OOB rows -> zeros
static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t dpad, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return NVS_ERR_CUDA;
  cuuint64_t dims[2] = {dpad, rows};
  cuuint64_t strides[1] = {dpad * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVS_OK : NVS_ERR_CUDA;
}

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }
static inline int dpad_of(int d) { return (d + BK - 1) / BK * BK; }

struct Layout {
  int dpad, n_mblk, n_tiles, tiles_per_strip, n_strips, kk, qpad;
  size_t off_qb, off_qn, off_gb, off_cd, off_ci, off_sd, off_si, total;
};

static Layout make_layout(long long n_db, int nq, int d, int k) {
  Layout L;
  L.dpad = dpad_of(d);
  L.n_mblk = (nq + BM - 1) / BM;
  L.qpad = L.n_mblk * BM;
  L.n_tiles = (int)((n_db + BN - 1) / BN);
  // Strips (each (strip, query-block pair) is one work unit of a CTA pair and yields k candidates per query):
  // enough units for ~32 rounds over the 74 clusters of a B200 (load balance to ~2 %), no more -- every strip
  // restarts its per-row lists and adds k candidates per query to the selection pass -- at most 256, at least
  // 16 tiles (4096 rows) each.
  const int n_mpair = (L.n_mblk + 1) / 2;
  int want = (32 * 74 + n_mpair - 1) / n_mpair;
  if (want > 256) want = 256;
  int tps = (L.n_tiles + want - 1) / want;
  if (tps < 16) tps = 16;
  L.tiles_per_strip = tps;
  L.n_strips = (L.n_tiles + tps - 1) / tps;
  L.kk = 2 * k > k + 32 ? 2 * k : k + 32;  // candidates re-ranked exactly
  if (L.kk > L.n_strips * k) L.kk = L.n_strips * k;
  if (L.kk > 256) L.kk = 256;
  size_t o = 0;
  L.off_qb = o; o += al256((size_t)L.qpad * L.dpad * 2);
  L.off_qn = o; o += al256((size_t)L.qpad * 4);
  L.off_gb = o; o += al256((size_t)L.qpad * 4);
  L.off_cd = o; o += al256((size_t)L.qpad * L.n_strips * k * 4);
  L.off_ci = o; o += al256((size_t)L.qpad * L.n_strips * k * 4);
  L.off_sd = o; o += al256((size_t)nq * L.kk * 4);
  L.off_si = o; o += al256((size_t)nq * L.kk * 4);
  L.total = o;
  return L;
}

}  // namespace rt
}  // namespace nvs

using namespace nvs;
using namespace nvs::rt;

extern "C" int32_t nvs_flat_padded_dim(int32_t d) { return d > 0 ? dpad_of(d) : 0; }

extern "C" int nvs_flat_prepare(const float* x, int64_t n, int32_t d, void* x_bf16, float* norms, void* stream) {
  if (!x || !x_bf16 || !norms || n <= 0 || d <= 0) return NVS_ERR_ARG;
  const int dp = dpad_of(d);
  const long long blocks = (n + 7) / 8;
  to_bf16_norm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)x_bf16, norms, n, d, dp);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" size_t nvs_flat_search_workspace_bytes(int64_t n_db, int32_t nq, int32_t d, int32_t k) {
  if (n_db <= 0 || nq <= 0 || d <= 0 || k <= 0) return 0;
  return make_layout(n_db, nq, d, k).total;
}

extern "C" int nvs_flat_search(const float* db, const void* db_bf16, const float* db_norms, int64_t n_db,
                               const float* q, int32_t nq, int32_t d, int32_t k, int64_t id_offset, float* out_D,
                               int64_t* out_I, void* workspace, size_t workspace_bytes, void* ev_gemm_start,
                               void* ev_gemm_stop, void* stream) {
  if (!db || !db_bf16 || !db_norms || !q || !out_D || !out_I || !workspace) return NVS_ERR_ARG;
  if (n_db <= 0 || nq <= 0 || d <= 0 || k <= 0) return NVS_ERR_ARG;
  if (k > KMAX) return NVS_ERR_UNSUPPORTED;   // per-row list lives in shared memory
  if (k > n_db) return NVS_ERR_ARG;
  if (n_db > 0x7fffffffLL - BN) return NVS_ERR_UNSUPPORTED;
  const Layout L = make_layout(n_db, nq, d, k);
  if (workspace_bytes < L.total) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>(ws + L.off_qb);
  float* qn = reinterpret_cast<float*>(ws + L.off_qn);
  float* gb = reinterpret_cast<float*>(ws + L.off_gb);
  float* cd = reinterpret_cast<float*>(ws + L.off_cd);
  int32_t* ci = reinterpret_cast<int32_t*>(ws + L.off_ci);
  float* sd = reinterpret_cast<float*>(ws + L.off_sd);
  int32_t* si = reinterpret_cast<int32_t*>(ws + L.off_si);

  // queries -> bf16 (rows past nq stay zero: they only feed padded accumulator rows that are never flushed)
  cudaError_t e = cudaMemsetAsync(qb, 0, (size_t)L.qpad * L.dpad * 2, st);
  if (e != cudaSuccess) return nvs_set_cuda_error(e);
  to_bf16_norm_kernel<<<(nq + 7) / 8, 256, 0, st>>>(q, qb, qn, nq, d, L.dpad);
  NVS_CHECK_LAUNCH();
  fill_inf_kernel<<<(L.qpad + 255) / 256, 256, 0, st>>>(gb, L.qpad);
  NVS_CHECK_LAUNCH();

  CUtensorMap mq, mx;
  if (make_map(&mq, qb, (uint64_t)L.qpad, (uint64_t)L.dpad, BM) != NVS_OK) return NVS_ERR_CUDA;
  if (make_map(&mx, db_bf16, (uint64_t)n_db, (uint64_t)L.dpad, BN / 2) != NVS_OK) return NVS_ERR_CUDA;  // half tiles

  Params p;
  p.xnorm = db_norms; p.gbound = gb; p.cand_d = cd; p.cand_i = ci;
  p.Q = nq; p.N = (int)n_db; p.kblocks = L.dpad / BK;
  p.k = k; p.kp = k | 1;
  p.n_mblk = L.n_mblk; p.n_strips = L.n_strips; p.tiles_per_strip = L.tiles_per_strip; p.n_tiles = L.n_tiles;
  p.n_mpair = (L.n_mblk + 1) / 2;

  static int sm_count = 0;
  static bool attr_done = false;
  if (!attr_done) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    e = cudaFuncSetAttribute(flat_l2_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    attr_done = true;
  }
  const int n_units = ((L.n_mblk + 1) / 2) * L.n_strips;           // work units of CTA pairs
  int grid = 2 * (n_units < sm_count / 2 ? n_units : sm_count / 2);  // clusters of two CTAs
  if (ev_gemm_start) cudaEventRecord(static_cast<cudaEvent_t>(ev_gemm_start), st);
  flat_l2_topk_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(mq, mx, p);
  NVS_CHECK_LAUNCH();
  if (ev_gemm_stop) cudaEventRecord(static_cast<cudaEvent_t>(ev_gemm_stop), st);

  const int n_cand = L.n_strips * k;
  static bool sel_attr = false;
  if (!sel_attr) {
    e = cudaFuncSetAttribute(select_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    e = cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    sel_attr = true;
  }
  if ((size_t)n_cand * 8 > 96 * 1024 || (size_t)(d + L.kk) * 4 > 96 * 1024) return NVS_ERR_UNSUPPORTED;
  select_candidates_kernel<<<nq, 256, (size_t)n_cand * 8, st>>>(cd, ci, n_cand, L.kk, sd, si);
  NVS_CHECK_LAUNCH();
  rerank_kernel<<<nq, 256, (size_t)(d + L.kk) * 4, st>>>(q, db, db_norms, si, L.kk, k, d, (long long)id_offset,
                                                         out_D, out_I);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

extern "C" int nvs_topk_merge(const float* D_parts, const int64_t* I_parts, int32_t parts, int32_t nq, int32_t k,
                              float* out_D, int64_t* out_I, void* stream) {
  if (!D_parts || !I_parts || !out_D || !out_I || parts <= 0 || parts > 16 || nq <= 0 || k <= 0) return NVS_ERR_ARG;
  merge_parts_kernel<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream>>>(D_parts, I_parts, parts, nq, k, out_D, out_I);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
