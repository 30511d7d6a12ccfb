// Shared helpers for all kernels of libfmm_b200: status codes, dtype tags, small device utils.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define FMM_OK 0
#define FMM_ERR_ARG (-1)
#define FMM_ERR_CUDA (-2)
#define FMM_ERR_SMEM (-3)

#define FMM_DT_BF16 0
#define FMM_DT_F32 1

namespace fmm {

void set_last_error(const char* fmt, ...);

#define FMM_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      fmm::set_last_error(__VA_ARGS__);   \
      return FMM_ERR_ARG;                 \
    }                                     \
  } while (0)

#define FMM_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      fmm::set_last_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return FMM_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

int num_sms();

template <typename T>
struct ActTraits;
template <>
struct ActTraits<__nv_bfloat16> {
  static constexpr int kParts = 1;  // operand parts fed to the bf16 tensor cores
  static constexpr int kDt = FMM_DT_BF16;
};
template <>
struct ActTraits<float> {
  static constexpr int kParts = 3;  // x = p0 + p1 + p2 (bf16 each): fp32-grade products
  static constexpr int kDt = FMM_DT_F32;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// load 8 consecutive activations (16B for bf16, 32B for fp32) as fp32; pointer 16B aligned
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// The same 8 activations as RAW bits: streaming kernels issue the loads of several rows back to back into Raw8 registers
// and convert afterwards. With load8 in a rolled-up loop the compiler keeps ONE 16-byte load in flight per thread (load,
// wait, convert, accumulate, next load: ncu r02, every row-walk kernel sat at 2.2-3.9 TB/s on exactly that).
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
};
template <>
struct Raw8<float> {
  float4 a, b;
};
__device__ __forceinline__ void ldraw8(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) {
  r.u = __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void ldraw8(const float* p, Raw8<float>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p + 4));
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float (&f)[8]) {
  const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&f)[8]) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
  f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}

// pack 8 fp32 -> 8 bf16 (one 16-byte chunk)
__device__ __forceinline__ uint4 pack8_bf16(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
// split 8 fp32 into the next bf16 part: out = bf16(f), f -= float(out)
__device__ __forceinline__ uint4 split8_bf16(float (&f)[8]) {
  uint4 u;
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2bfloat16_rn(f[i]);
    f[i] -= __bfloat162float(h[i]);
  }
  return u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace fmm
