// Library-wide plumbing of libfmm_b200: last-error text, device query, version.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace fmm {

__device__ int g_wait_prof_enable = 0;
__device__ unsigned long long g_wait_prof[32];

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace fmm

extern "C" {

const char* fmm_last_error(void) { return fmm::g_last_error; }

int fmm_version(void) { return 100; }

// dev aid: enable/reset (enable >= 0) and read back the per-wait-site blocked-cycle counters
int fmm_debug_wait_profile(int enable, unsigned long long* out32) {
  if (out32) {
    if (cudaMemcpyFromSymbol(out32, fmm::g_wait_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return FMM_ERR_CUDA;
  }
  if (enable >= 0) {
    unsigned long long zero[32] = {0};
    if (cudaMemcpyToSymbol(fmm::g_wait_prof, zero, sizeof(zero)) != cudaSuccess) return FMM_ERR_CUDA;
    if (cudaMemcpyToSymbol(fmm::g_wait_prof_enable, &enable, sizeof(int)) != cudaSuccess) return FMM_ERR_CUDA;
  }
  return FMM_OK;
}

// 1 if the current device can run the sm_100a kernels of this library.
int fmm_device_supported(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

}  // extern "C"
