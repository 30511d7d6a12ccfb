// Window producer shared by the forward (tapconv) and weight-gradient (wgrad) engines.
//
// A "window chunk" is [atoms = time steps][8 columns][64 channels] bf16 in the SWIZZLE_128B image
// (1024 bytes per time step). Each producer thread owns one 16-byte piece (8 channels) of one
// column and walks over the time steps with a fixed stride; global loads are issued in batches
// (all loads of a batch first, then transform + store) so that every thread keeps several
// 16/32-byte requests in flight: the loop is latency-bound otherwise.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace fmm {

template <typename T>
struct RawPiece;
template <>
struct RawPiece<__nv_bfloat16> {
  uint4 u;
  static constexpr int kBatch = 8;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void to_f32(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct RawPiece<float> {
  float4 a, b;
  static constexpr int kBatch = 4;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0, 0, 0, 0); }
  __device__ __forceinline__ void to_f32(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// Produce atoms a0, a0+astep, ... < natoms of one chunk.
//   col_base : pointer to element [n][0][v][cb] of the source (row pitch `pitch_t` elements per time step)
//   t_lo     : time index of atom 0; times outside [0, Tn) give zeros (zero padding AFTER the transform)
//   nvalid   : number of valid channels at cb (<= 0: all-zero piece; < 8: ragged tail, scalar loads)
//   sdst     : shared address of this thread's piece in atom 0 (part 0); parts are `part_bytes` apart
template <typename T, int kParts, bool kAffine>
__device__ __forceinline__ void produce_chunk(const T* __restrict__ col_base, size_t pitch_t, bool col_ok, int Tn,
                                              int t_lo, int natoms, int a0, int astep, int nvalid, bool vec_ok,
                                              const float (&sc)[8], const float (&sh)[8], bool relu, uint32_t sdst,
                                              uint32_t part_bytes) {
  constexpr int kBatch = RawPiece<T>::kBatch;
  for (int ab = a0; ab < natoms; ab += astep * kBatch) {
    RawPiece<T> raw[kBatch];
    bool ok[kBatch];
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int a = ab + b * astep;
      const int ti = t_lo + a;
      ok[b] = col_ok && a < natoms && ti >= 0 && ti < Tn && nvalid > 0;
      if (ok[b] && vec_ok && nvalid >= 8) {
        raw[b].load(col_base + static_cast<size_t>(ti) * pitch_t);
      } else {
        raw[b].zero();
      }
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int a = ab + b * astep;
      if (a >= natoms) break;
      float f[8];
      raw[b].to_f32(f);
      if (ok[b] && !(vec_ok && nvalid >= 8)) {  // ragged / unaligned rows: scalar loads
        const T* src = col_base + static_cast<size_t>(t_lo + a) * pitch_t;
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = i < nvalid ? to_f32(src[i]) : 0.f;
      }
      if (kAffine) {
        if (ok[b]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float y = fmaf(f[i], sc[i], sh[i]);
            if (relu) y = fmaxf(y, 0.f);
            f[i] = i < nvalid ? y : 0.f;
          }
        }
      }
      const uint32_t dst = sdst + static_cast<uint32_t>(a) * 1024u;
      if (kParts == 1) {
        sts128(dst, pack8_bf16(f));
      } else {
#pragma unroll
        for (int part = 0; part < kParts; ++part) sts128(dst + part * part_bytes, split8_bf16(f));
      }
    }
  }
}

}  // namespace fmm

namespace fmm {

// ------------------------------------------------------------------------------------------
// cp.async (LDGSTS) path for bf16 windows with 16-byte aligned rows: the copies of several
// window slots are in flight at once (no registers held), which is what hides the global-memory
// latency; an optional per-channel affine(+ReLU) is then applied IN PLACE by the thread that
// issued the copy (its own completed cp.async data is visible to it after wait_group).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

// issue the copies of atoms a0, a0+astep, ... of one chunk (zero fill outside [0,Tn) / invalid column).
// The producer warps are instruction bound (one LDGSTS moves only 16 bytes per lane), so the common
// case - the whole window inside the clip - runs without per-copy predicates or address multiplies.
__device__ __forceinline__ void cpasync_issue_chunk(const __nv_bfloat16* __restrict__ col_base, size_t pitch_t,
                                                    bool col_ok, int Tn, int t_lo, int natoms, int a0, int astep,
                                                    uint32_t sdst) {
  const int n_it = (natoms - a0 + astep - 1) / astep;
  uint32_t dst = sdst + static_cast<uint32_t>(a0) * 1024u;
  const uint32_t dstep = static_cast<uint32_t>(astep) * 1024u;
  if (col_ok && t_lo >= 0 && t_lo + natoms <= Tn) {
    const __nv_bfloat16* src = col_base + static_cast<size_t>(t_lo + a0) * pitch_t;
    const size_t sstep = static_cast<size_t>(astep) * pitch_t;
#pragma unroll 4
    for (int i = 0; i < n_it; ++i) {
      cp_async16(dst, src, 16u);
      src += sstep;
      dst += dstep;
    }
  } else {
    int ti = t_lo + a0;
    for (int i = 0; i < n_it; ++i) {
      const bool ok = col_ok && ti >= 0 && ti < Tn;
      const __nv_bfloat16* src = ok ? col_base + static_cast<size_t>(ti) * pitch_t : col_base;
      cp_async16(dst, src, ok ? 16u : 0u);
      ti += astep;
      dst += dstep;
    }
  }
}

// y = relu?(x*sc + sh) in place on the pieces this thread copied (padding pieces stay zero)
__device__ __forceinline__ void inplace_affine_chunk(bool col_ok, int Tn, int t_lo, int natoms, int a0, int astep,
                                                     const float (&sc)[8], const float (&sh)[8], bool relu,
                                                     uint32_t sdst) {
  for (int a = a0; a < natoms; a += astep) {
    const int ti = t_lo + a;
    if (!(col_ok && ti >= 0 && ti < Tn)) continue;
    const uint32_t addr = sdst + static_cast<uint32_t>(a) * 1024u;
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
    float f[8];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(f[i], sc[i], sh[i]);
      f[i] = relu ? fmaxf(y, 0.f) : y;
    }
    sts128(addr, pack8_bf16(f));
  }
}

}  // namespace fmm
