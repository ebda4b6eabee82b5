// Shared device helpers for the nanovs sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nanovs.h"

#define NVS_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return nvs_set_cuda_error(e__);  \
  } while (0)

int nvs_set_cuda_error(cudaError_t e);  // cabi.cu: records the message, returns NVS_ERR_CUDA
int nvs_sm_count(void);                 // cabi.cu: SM count of the CURRENT device (cached per device)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: opt in once per (kernel, device).
// One 64-bit mask per call site (bit = device ordinal); the benign race just repeats an idempotent call.
#define NVS_OPT_IN_SMEM(kernel, bytes)                                                              \
  do {                                                                                              \
    static unsigned long long done_mask__ = 0ull;                                                   \
    int dev__ = 0;                                                                                  \
    cudaGetDevice(&dev__);                                                                          \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                                          \
    if (!(__atomic_load_n(&done_mask__, __ATOMIC_ACQUIRE) & bit__)) {                               \
      cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
      if (e__ != cudaSuccess) return nvs_set_cuda_error(e__);                                       \
      __atomic_fetch_or(&done_mask__, bit__, __ATOMIC_RELEASE);                                     \
    }                                                                                               \
  } while (0)

namespace nvs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 4-byte async copy global->shared; src_bytes = 0 zero-fills (used for conv zero padding).
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(valid ? 4 : 0));
}
// 16-byte async copy (weights; L2-only caching: every CTA re-reads the same rows).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act, int channel) {
  switch (act) {
    case NVS_ACT_LRELU: return v > 0.f ? v : 0.01f * v;
    case NVS_ACT_RELU: return fmaxf(v, 0.f);
    case NVS_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case NVS_ACT_TANH: return tanhf(v);
    case NVS_ACT_SIGMOID_TANH: return channel == 0 ? 1.f / (1.f + expf(-v)) : tanhf(v);
    case NVS_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    default: return v;
  }
}

}  // namespace nvs
