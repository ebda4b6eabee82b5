// Efficient-self-attention core (modules/segformer.py:113-133) for head_dim 12 / 16 with the Q.K^T contraction on the
// 5th-generation tensor cores.
//
// Work per (query, key) pair is 2*d FMAs for the logit, one exp and 2*d FMAs for P.V.  P.V is an N = head_dim GEMM
// (16 columns: a tcgen05.mma costs >= 50 cycles whatever N is, so it would be slower than the FMA pipe) but Q.K^T is a
// full-width one: M = 128 queries, N = 64 keys per instruction, K = head_dim (two tf32 k-steps).  So the logits are
// computed as a 3xTF32 tcgen05 product (q = q_hi + q_lo, k = k_hi + k_lo, S = q_hi k_hi + q_lo k_hi + q_hi k_lo, fp32
// accumulation in TMEM: fp32-grade, like conv_tc.cu) and only the streaming softmax and P.V stay on the CUDA cores,
// which halves their FMA work -- the bound of the all-FFMA kernel in attention.cu.
//
//   CTA          256 queries (two M = 128 tiles: a thread owns TMEM lane r of both, i.e. two queries, so every V row it
//                loads from shared memory feeds two queries) of one (frame, head); 2 CTAs per SM (256 TMEM columns each)
//   Q            loaded once, pre-scaled by d^-1/2 log2(e), split hi / lo, stored in shared memory in the canonical
//                K-major SWIZZLE_64B layout (64-byte rows): A operand through a shared-memory descriptor
//   K / V        stream in blocks of 64 keys through a 4-stage ring filled by a producer warp: K is read from the NCHW
//                planes (coalesced along keys), split hi / lo and stored as the K-major B operand; V is stored as
//                plain [key][16] rows for the softmax warps' broadcast LDS.128
//   S            2 stages x 2 tiles x 64 fp32 columns in TMEM: the MMAs of block i+1 run while block i is consumed
//   softmax      4 warps (warp w = TMEM lane quadrant w), tcgen05.ld 16 logits of each of the two queries, online
//                max / sum (rescale only when the maximum moves), ex2.approx, packed FFMA2 for P.V
//
// Warps: 0-3 softmax, 4 K/V producer, 5 MMA issuer (allocates TMEM).
#include <stdlib.h>

#include "common.cuh"

namespace nvs {
namespace att {

constexpr int KB = 64;                 // keys per block = N of one MMA
constexpr int NST_MAX = 4;             // K/V stages (4 with two CTAs per SM, 3 with three)
constexpr int DK = 16;                 // contraction length (head_dim 12 is zero padded)
constexpr int QT = 128;                // queries per tile (TMEM lanes)
constexpr int QPB = 2 * QT;            // queries per CTA
constexpr int THREADS = 6 * 32;
constexpr int Q_BYTES = QT * DK * 4;   // one of q_hi / q_lo of one tile
constexpr int K_BYTES = KB * DK * 4;   // one of k_hi / k_lo / v of one block
constexpr int SM_Q = 0;                              // [tile][hi, lo]
constexpr int SM_KV = SM_Q + 4 * Q_BYTES;            // [stage][k_hi, k_lo, v]
constexpr int N_BARS = 2 * NST_MAX + 4 + 1;
constexpr int sm_bar(int nst) { return SM_KV + nst * 3 * K_BYTES; }
constexpr int smem_bytes(int nst) { return sm_bar(nst) + 8 * N_BARS + 16 + 1024; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
// the same for the single-thread roles (K/V producer, MMA issuer), whose waits are long and have slack: back off
// between polls so that the spin does not take issue slots from the softmax warp on the same scheduler (ncu: the
// issuer's try_wait loop alone executed a quarter as many instructions as all FFMA2 of the kernel)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(100);
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// K-major operand with 64-byte rows (16 tf32), SWIZZLE_64B (layout 4), 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KB >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);

// 16 logits of one query tile: tcgen05.ld is asynchronous -- the registers may only be read after tmem_ld_wait, which
// takes them as read-write operands so that neither compiler moves a use above it
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t* a, uint32_t* b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                 "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// hi = x rounded to nearest tf32, lo = x - hi (exact in fp32) rounded to nearest tf32 (the MMA drops the low 13 bits)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  hi = __uint_as_float(h);
  lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}
// byte offset of 16-byte chunk c of row r inside a K-major SWIZZLE_64B tile (64-byte rows)
__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// CTAS = CTAs per SM the variant is sized for: 2 -> two S stages (256 TMEM columns), four K/V stages; 3 -> one S stage
// (128 columns), three K/V stages, <= 113 registers: the MMAs of a block then wait for the softmax warps to drain the
// previous one, which the two other CTAs cover (measured slower: 57.7 vs 67.3 TFLOP/s at 32k x 8k tokens).
// PREFETCH: the TMEM loads of the next 16-key chunk are in flight while a chunk is processed (168 registers).
template <int D, int CTAS, bool PREFETCH>
__global__ void __launch_bounds__(THREADS, CTAS) attention_tc_kernel(const float* __restrict__ q,
                                                                  const float* __restrict__ kv,
                                                                  float* __restrict__ out, int C, int Nq, int Nk,
                                                                  float scale_log2e) {
  static_assert(D == 12 || D == 16, "head_dim 12 / 16");
  static_assert(CTAS == 2 || CTAS == 3, "CTAs per SM");
  constexpr int H = D / 2;
  constexpr int NST = CTAS == 2 ? 4 : 3;
  constexpr int SS = CTAS == 2 ? 2 : 1;             // S stages
  constexpr uint32_t TMEM_COLS = SS * 2 * KB;       // stages x 2 tiles x 64 keys
  constexpr int SM_BAR = sm_bar(NST);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + SM_BAR;
  auto kv_full = [&](int i) { return bar0 + 8u * i; };
  auto kv_empty = [&](int i) { return bar0 + 8u * (NST + i); };
  auto s_full = [&](int i) { return bar0 + 8u * (2 * NST + i); };
  auto s_empty = [&](int i) { return bar0 + 8u * (2 * NST + 2 + i); };
  const uint32_t q_full = bar0 + 8u * (2 * NST + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_BAR + 8 * N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * QPB;
  const float* qb = q + ((size_t)b * C + head * D) * Nq;
  const float* kb = kv + ((size_t)b * 2 * C + head * D) * Nk;
  const float* vb = kv + ((size_t)b * 2 * C + C + head * D) * Nk;
  const int n_blocks = (Nk + KB - 1) / KB;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(kv_full(i), 32);   // the producer warp's lanes
      mbar_init(kv_empty(i), 4);   // the softmax warps (their arrival also implies the block's MMAs have retired)
    }
    for (int i = 0; i < SS; ++i) {
      mbar_init(s_full(i), 1);     // tcgen05.commit
      mbar_init(s_empty(i), 4);
    }
    mbar_init(q_full, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== softmax + P.V: thread = query r of tile 0 and of tile 1 ===========================
    const int r = threadIdx.x;  // TMEM lane
    int n[2];
    bool active[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      n[t] = q0 + t * QT + r;
      active[t] = n[t] < Nq;
      float hi[DK], lo[DK];
#pragma unroll
      for (int c = 0; c < DK; ++c) {
        const float x = (c < D && active[t]) ? qb[(size_t)c * Nq + n[t]] * scale_log2e : 0.f;  // coalesced along queries
        split_tf32(x, hi[c], lo[c]);
      }
      uint8_t* th = sm + SM_Q + (2 * t) * Q_BYTES;
#pragma unroll
      for (int c4 = 0; c4 < DK / 4; ++c4) {
        const uint32_t o = sw64(r, c4);
        *reinterpret_cast<float4*>(th + o) = make_float4(hi[4 * c4], hi[4 * c4 + 1], hi[4 * c4 + 2], hi[4 * c4 + 3]);
        *reinterpret_cast<float4*>(th + Q_BYTES + o) = make_float4(lo[4 * c4], lo[4 * c4 + 1], lo[4 * c4 + 2], lo[4 * c4 + 3]);
      }
    }
    proxy_fence();  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
    mbar_arrive(q_full);

    float2 o0[H], o1[H];
#pragma unroll
    for (int c = 0; c < H; ++c) o0[c] = o1[c] = make_float2(0.f, 0.f);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

    for (int i = 0; i < n_blocks; ++i) {
      const int st = i % NST, ss = i % SS;
      const int nvalid = min(KB, Nk - i * KB);
      const int n_chunks = (nvalid + 15) >> 4;
      mbar_wait(kv_full(st), (uint32_t)(i / NST) & 1u);  // V rows of this block (acquire of the producer's stores)
      mbar_wait(s_full(ss), (uint32_t)(i / SS) & 1u);    // logits of this block
      tc_fence_after();
      const float4* vrow = reinterpret_cast<const float4*>(sm + SM_KV + (st * 3 + 2) * K_BYTES);
      const uint32_t s_addr = lane_addr + (uint32_t)(ss * 2 * KB);
      // one chunk = 16 keys of both queries; the TMEM load of chunk ch + 1 is in flight while chunk ch is processed
      auto issue = [&](uint32_t* ra, uint32_t* rb, int ch) {
        tmem_ld16_issue(s_addr + (uint32_t)(ch * 16), ra);
        tmem_ld16_issue(s_addr + (uint32_t)(KB + ch * 16), rb);
      };
      auto release_s = [&]() {  // every logit of this stage is in registers: the MMAs of block i + 2 may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(ss));
      };
      auto process = [&](const uint32_t* ra, const uint32_t* rb, int ch) {
        float s0[16], s1[16];
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          s0[g] = __uint_as_float(ra[g]);
          s1[g] = __uint_as_float(rb[g]);
        }
        const int left = nvalid - ch * 16;
        if (left < 16) {
#pragma unroll
          for (int g = 0; g < 16; ++g)
            if (g >= left) s0[g] = s1[g] = -INFINITY;
        }
        float g0 = s0[0], g1 = s1[0];
#pragma unroll
        for (int g = 1; g < 16; ++g) {
          g0 = fmaxf(g0, s0[g]);
          g1 = fmaxf(g1, s1[g]);
        }
        if (g0 > m0) {  // the running maximum moves: rescale (rare after the first blocks)
          const float corr = ex2(m0 - g0);  // m = -inf at the start -> 0
          l0 *= corr;
          const float2 c2 = make_float2(corr, corr);
#pragma unroll
          for (int c = 0; c < H; ++c) o0[c] = __fmul2_rn(o0[c], c2);
          m0 = g0;
        }
        if (g1 > m1) {
          const float corr = ex2(m1 - g1);
          l1 *= corr;
          const float2 c2 = make_float2(corr, corr);
#pragma unroll
          for (int c = 0; c < H; ++c) o1[c] = __fmul2_rn(o1[c], c2);
          m1 = g1;
        }
        const float4* vr = vrow + (ch * 16) * (DK / 4);
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          float2 vv[H];
#pragma unroll
          for (int c4 = 0; c4 < D / 4; ++c4) {
            const float4 v4 = vr[g * (DK / 4) + c4];  // warp-broadcast LDS.128
            vv[2 * c4] = make_float2(v4.x, v4.y);
            vv[2 * c4 + 1] = make_float2(v4.z, v4.w);
          }
          const float p0 = ex2(s0[g] - m0), p1 = ex2(s1[g] - m1);
          l0 += p0;
          l1 += p1;
          const float2 a = make_float2(p0, p0), bq = make_float2(p1, p1);
#pragma unroll
          for (int c = 0; c < H; ++c) {
            o0[c] = __ffma2_rn(a, vv[c], o0[c]);
            o1[c] = __ffma2_rn(bq, vv[c], o1[c]);
          }
        }
      };
      if (PREFETCH) {
        uint32_t a0[16], a1[16], b0[16], b1[16];
        issue(a0, a1, 0);
#pragma unroll 1
        for (int ch = 0; ch < n_chunks; ch += 2) {
          tmem_ld_wait(a0, a1);
          const bool more1 = ch + 1 < n_chunks;
          if (more1) issue(b0, b1, ch + 1);
          else release_s();
          process(a0, a1, ch);
          if (more1) {
            tmem_ld_wait(b0, b1);
            if (ch + 2 < n_chunks) issue(a0, a1, ch + 2);
            else release_s();
            process(b0, b1, ch + 1);
          }
        }
      } else {
#pragma unroll 1
        for (int ch = 0; ch < n_chunks; ++ch) {
          uint32_t a0[16], a1[16];
          issue(a0, a1, ch);
          tmem_ld_wait(a0, a1);
          if (ch == n_chunks - 1) release_s();
          process(a0, a1, ch);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(kv_empty(st));
    }
    float* ob = out + ((size_t)b * C + head * D) * Nq;
    if (active[0]) {
      const float inv = 1.f / l0;
#pragma unroll
      for (int c = 0; c < H; ++c) {
        ob[(size_t)(2 * c) * Nq + n[0]] = o0[c].x * inv;
        ob[(size_t)(2 * c + 1) * Nq + n[0]] = o0[c].y * inv;
      }
    }
    if (active[1]) {
      const float inv = 1.f / l1;
#pragma unroll
      for (int c = 0; c < H; ++c) {
        ob[(size_t)(2 * c) * Nq + n[1]] = o1[c].x * inv;
        ob[(size_t)(2 * c + 1) * Nq + n[1]] = o1[c].y * inv;
      }
    }
  } else if (warp == 4) {
    // =========================== K/V producer: lane = key l and key l + 32 of the block ===========================
    for (int i = 0; i < n_blocks; ++i) {
      const int st = i % NST;
      const int j0 = i * KB + lane, j1 = j0 + 32;
      const bool ok0 = j0 < Nk, ok1 = j1 < Nk;
      float k0[DK], k1[DK], v0[DK], v1[DK];
#pragma unroll
      for (int c = 0; c < DK; ++c) {  // global loads first (coalesced along keys), independent of the ring
        k0[c] = (c < D && ok0) ? kb[(size_t)c * Nk + j0] : 0.f;
        k1[c] = (c < D && ok1) ? kb[(size_t)c * Nk + j1] : 0.f;
        v0[c] = (c < D && ok0) ? vb[(size_t)c * Nk + j0] : 0.f;
        v1[c] = (c < D && ok1) ? vb[(size_t)c * Nk + j1] : 0.f;
      }
      mbar_wait_relaxed(kv_empty(st), ((uint32_t)(i / NST) & 1u) ^ 1u);
      uint8_t* kh = sm + SM_KV + (st * 3) * K_BYTES;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float* kk = half ? k1 : k0;
        const float* vv = half ? v1 : v0;
        const int row = lane + 32 * half;
#pragma unroll
        for (int c4 = 0; c4 < DK / 4; ++c4) {
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32(kk[4 * c4 + e], hi[e], lo[e]);
          const uint32_t o = sw64(row, c4);
          *reinterpret_cast<float4*>(kh + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(kh + K_BYTES + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<float4*>(kh + 2 * K_BYTES + row * 64 + c4 * 16) =
              make_float4(vv[4 * c4], vv[4 * c4 + 1], vv[4 * c4 + 2], vv[4 * c4 + 3]);
        }
      }
      proxy_fence();
      mbar_arrive(kv_full(st));
    }
  } else if (lane == 0) {
    // =========================== MMA issuer ===========================
    mbar_wait_relaxed(q_full, 0);
    tc_fence_after();
    for (int i = 0; i < n_blocks; ++i) {
      const int st = i % NST, ss = i % SS;
      mbar_wait_relaxed(kv_full(st), (uint32_t)(i / NST) & 1u);
      mbar_wait_relaxed(s_empty(ss), ((uint32_t)(i / SS) & 1u) ^ 1u);
      tc_fence_after();
      const uint64_t k_hi = make_desc64(base + SM_KV + (st * 3) * K_BYTES), k_lo = k_hi + (uint64_t)(K_BYTES >> 4);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t d = tmem_base + (uint32_t)(ss * 2 * KB + t * KB);
        const uint64_t q_hi = make_desc64(base + SM_Q + (2 * t) * Q_BYTES), q_lo = q_hi + (uint64_t)(Q_BYTES >> 4);
        // 8 tf32 = 32 bytes along K = +2 in the descriptor's (address >> 4) field
        tc_mma_tf32(d, q_hi, k_hi, IDESC, 0u);
        tc_mma_tf32(d, q_hi + 2, k_hi + 2, IDESC, 1u);
        tc_mma_tf32(d, q_lo, k_hi, IDESC, 1u);
        tc_mma_tf32(d, q_lo + 2, k_hi + 2, IDESC, 1u);
        tc_mma_tf32(d, q_hi, k_lo, IDESC, 1u);
        tc_mma_tf32(d, q_hi + 2, k_lo + 2, IDESC, 1u);
      }
      tc_commit(s_full(ss));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace att

template <int D, int CTAS, bool PF>
static int launch_tc(const float* q, const float* kv, float* out, int B, int C, int heads, int Nq, int Nk,
                     float scale_log2e, cudaStream_t st) {
  constexpr int smem = att::smem_bytes(CTAS == 2 ? 4 : 3);
  auto kern = att::attention_tc_kernel<D, CTAS, PF>;
  NVS_OPT_IN_SMEM(kern, smem);
  dim3 grid((Nq + att::QPB - 1) / att::QPB, heads, B);
  kern<<<grid, att::THREADS, smem, st>>>(q, kv, out, C, Nq, Nk, scale_log2e);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

// called by nvs_attention (attention.cu) for head_dim 12 / 16
int attention_tc_launch(const float* q, const float* kv, float* out, int B, int C, int heads, int Nq, int Nk,
                        float scale_log2e, cudaStream_t st) {
  const int d = C / heads;
  // variant: NVS_ATT_CTAS = 2 (default) | 3 CTAs per SM, NVS_ATT_PREFETCH = 0 (default) | 1 (two CTAs only)
  static int ctas = 0, pf = 0;
  if (!ctas) {
    const char* e = getenv("NVS_ATT_CTAS");
    const char* f = getenv("NVS_ATT_PREFETCH");
    pf = (f && atoi(f) == 1) ? 1 : 0;
    ctas = (e && atoi(e) == 3) ? 3 : 2;
  }
#define NVS_ATT_GO(D_) \
  return ctas == 3 ? launch_tc<D_, 3, false>(q, kv, out, B, C, heads, Nq, Nk, scale_log2e, st) \
       : (pf ? launch_tc<D_, 2, true>(q, kv, out, B, C, heads, Nq, Nk, scale_log2e, st) \
             : launch_tc<D_, 2, false>(q, kv, out, B, C, heads, Nq, Nk, scale_log2e, st))
  if (d == 16) { NVS_ATT_GO(16); }
  if (d == 12) { NVS_ATT_GO(12); }
#undef NVS_ATT_GO
  return NVS_ERR_UNSUPPORTED;
}

}  // namespace nvs
