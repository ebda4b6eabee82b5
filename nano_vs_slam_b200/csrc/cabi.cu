// Error plumbing + device probe for the C ABI.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

int nvs_set_cuda_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
  return NVS_ERR_CUDA;
}

extern "C" const char* nvs_last_error(void) { return g_err; }
extern "C" int nvs_abi_version(void) { return NVS_ABI_VERSION; }

extern "C" int nvs_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    nvs_set_cuda_error(e);
    return NVS_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    nvs_set_cuda_error(e);
    return NVS_ERR_NO_DEVICE;
  }
  if (prop.major != 10) {
    snprintf(g_err, sizeof(g_err), "device %d is sm_%d%d; libnanovs is built for sm_100a only", dev,
             prop.major, prop.minor);
    return NVS_ERR_NO_DEVICE;
  }
  return NVS_OK;
}

// SM count of the current device, cached per device ordinal (a process may drive several GPUs).
int nvs_sm_count(void) {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int v = __atomic_load_n(&cache[dev & 63], __ATOMIC_RELAXED);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    __atomic_store_n(&cache[dev & 63], v, __ATOMIC_RELAXED);
  }
  return v;
}
