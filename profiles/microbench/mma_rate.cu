// Micro-benchmark: issue rate / latency of tcgen05.mma for the shapes used by retrieval.cu and conv_tc.cu.
// One CTA per SM; one thread issues N back-to-back MMAs into one TMEM accumulator, commits, waits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate.bin mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sdesc(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND>  // 0: f16(bf16) SS, 1: tf32 SS, 2: tf32 TS
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint32_t ta, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int N, int NMMA, int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) ((uint32_t*)sm)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t fmt = KIND == 0 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t a = make_sdesc(base), b = make_sdesc(base + 32768);
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < NMMA; ++i) {
      mma<KIND>(tm, a + 2 * (i & 3), tm + 256 + 8 * (i & 3), b + 2 * (i & 3), idesc, i ? 1u : 0u);
      if (COMMIT_EVERY > 0 && (i % COMMIT_EVERY) == COMMIT_EVERY - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int KIND, int N, int CE = 0>
void run(const char* name, long long* d) {
  constexpr int NMMA = 2048;
  auto kern = k<KIND, N, NMMA, CE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int grid : {1, 148}) {
    kern<<<grid, 128, 100 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (CE) printf("[commit every %d] ", CE);
    printf("%-28s N=%3d grid=%3d  issue %.1f cyc/mma   complete %.1f cyc/mma  (%s)\n", name, N, grid, h[0] / (double)NMMA,
           h[1] / (double)NMMA, cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<0, 256>("bf16 SS M128 K16", d);
  run<0, 128>("bf16 SS M128 K16", d);
  run<0, 64>("bf16 SS M128 K16", d);
  run<1, 256>("tf32 SS M128 K8", d);
  run<1, 128>("tf32 SS M128 K8", d);
  run<1, 64>("tf32 SS M128 K8", d);
  run<1, 32>("tf32 SS M128 K8", d);
  run<2, 256>("tf32 TS(A in TMEM) M128 K8", d);
  run<2, 128>("tf32 TS(A in TMEM) M128 K8", d);
  run<2, 64>("tf32 TS(A in TMEM) M128 K8", d);
  run<2, 32>("tf32 TS(A in TMEM) M128 K8", d);
  run<2, 128, 8>("tf32 TS", d);
  run<2, 128, 4>("tf32 TS", d);
  run<2, 128, 2>("tf32 TS", d);
  run<2, 64, 8>("tf32 TS", d);
  run<2, 64, 2>("tf32 TS", d);
  run<0, 256, 4>("bf16 SS", d);
  return 0;
}
