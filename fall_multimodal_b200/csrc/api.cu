// Library-wide plumbing of libfmm_b200: last-error text, device query, version.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

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

static int g_sm_limit = 0;   // fmm_set_sm_limit: 0 = all SMs

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    // dev knob: persistent kernels size their grids by this count; FMM_SM_LIMIT=74 lets the kernels of two concurrent trunks
    // (fusion models, concurrent_streams) sit side by side on half of the SMs each instead of taking turns
    if (const char* e = getenv("FMM_SM_LIMIT")) {
      const int v = atoi(e);
      if (v >= 2 && v < n) n = v & ~1;
    }
    cached[dev] = n;
  }
  const int lim = g_sm_limit;
  return (lim >= 2 && lim < cached[dev]) ? lim : cached[dev];
}

}  // namespace fmm

extern "C" {

const char* fmm_last_error(void) { return fmm::g_last_error; }

int fmm_version(void) { return 100; }

// Grid budget of the persistent kernels (tap-conv / graph-conv / weight-gradient engines, TMA-staged streaming kernels), which
// size their grids by the SM count: n = 0 (default) all SMs; n = SMs / 2 lets the kernels of two concurrently running trunks of
// a fusion model sit side by side instead of taking turns on the whole chip (measured: 13.25 -> 13.07 ms per step at two trunks
// x 256 clips). Rounded down to an even count (CTA pairs). Process-wide, read at launch time (so it is baked into a captured
// graph); returns the previous value.
int fmm_set_sm_limit(int n) {
  const int prev = fmm::g_sm_limit;
  fmm::g_sm_limit = n >= 2 ? (n & ~1) : 0;
  return prev;
}

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
