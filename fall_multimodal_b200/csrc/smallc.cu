// 1x1 channel mixes whose narrow side has only a handful of channels: the first GSTCAN block mixes
// K*Cin = 9 (joints stream) or 6 (motion stream) aggregated input channels into 64 (stgcan.py:50-56 with
// in_channels = 3 / 2). A 64-channel tensor-core k-chunk would be 86 % padding and the kernels end up
// bound by streaming the 64-channel side, so these three run on CUDA cores, one pass over that tensor:
//   fwd   : out[r][co] = bias[v][co] + sum_k x[r][k] W[co][k]
//   dgrad : p[r][k]    = sum_co dy[r][co] W[co][k]
//   wgrad : dW[co][k] += sum_r dy[r][co] x[r][k]
// r = (n,t,v) rows of channels-last activations; W is addressed like the tap-conv weights
// (co*s_co + (k/K2)*s_k1 + (k%K2)*s_k2) so the k-major graph-conv weight needs no repacking.
#include "common.cuh"

namespace fmm {

constexpr int kMaxSmallC = 16;

struct SmallCParams {
  const void* x;     // [rows][Ck]  (fwd, wgrad)   / out of dgrad
  const void* y;     // [rows][Cw]  dy (dgrad, wgrad) / out of fwd
  const float* w;    // fp32 weights
  float* dw;         // wgrad output (atomics)
  const float* bias; // fwd: [V][Cw] or [Cw] (bias_vstride = 0) or null
  long long rows;
  int V, Ck, Cw, K2, bias_vstride;
  long long s_co, s_k1, s_k2;
};

__device__ __forceinline__ long long woff(const SmallCParams& p, int co, int k) {
  return co * p.s_co + (k / p.K2) * p.s_k1 + (k % p.K2) * p.s_k2;
}

// ---- forward: 8 output channels per thread, Cw/8 threads per row ----
template <typename T>
__global__ void __launch_bounds__(256) smallc_fwd_kernel(const SmallCParams p) {
  extern __shared__ float ws[];  // [Ck][Cw]
  for (int i = threadIdx.x; i < p.Ck * p.Cw; i += blockDim.x) ws[i] = p.w[woff(p, i % p.Cw, i / p.Cw)];
  __syncthreads();
  const T* __restrict__ X = reinterpret_cast<const T*>(p.x);
  T* __restrict__ O = reinterpret_cast<T*>(const_cast<void*>(p.y));
  const int c8n = p.Cw >> 3;
  const long long items = p.rows * c8n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const long long r = it / c8n;
    const int c0 = (int)(it % c8n) * 8;
    float acc[8];
    if (p.bias) {
      const int v = (int)(r % p.V);
      load8(p.bias + (long long)v * p.bias_vstride + c0, acc);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    }
    const T* xr = X + r * p.Ck;
    for (int k = 0; k < p.Ck; ++k) {
      const float xv = to_f32(xr[k]);
      const float* wr = ws + k * p.Cw + c0;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += xv * wr[e];
    }
    store8(O + r * p.Cw + c0, acc);
  }
}

// ---- dgrad: 8 threads per row, each 8 channels of dy, shuffle-reduced ----
template <typename T>
__global__ void __launch_bounds__(256) smallc_dgrad_kernel(const SmallCParams p) {
  extern __shared__ float ws[];  // [Ck][Cw]
  for (int i = threadIdx.x; i < p.Ck * p.Cw; i += blockDim.x) ws[i] = p.w[woff(p, i % p.Cw, i / p.Cw)];
  __syncthreads();
  const T* __restrict__ DY = reinterpret_cast<const T*>(p.y);
  T* __restrict__ P = reinterpret_cast<T*>(const_cast<void*>(p.x));
  const int c8n = p.Cw >> 3;  // threads per row (power of two, <= 32)
  const int sub = threadIdx.x % c8n;
  const int rows_per_block = blockDim.x / c8n;
  for (long long r0 = (long long)blockIdx.x * rows_per_block; r0 < p.rows; r0 += (long long)gridDim.x * rows_per_block) {
    const long long r = r0 + threadIdx.x / c8n;
    float acc[kMaxSmallC];
#pragma unroll
    for (int k = 0; k < kMaxSmallC; ++k) acc[k] = 0.f;
    if (r < p.rows) {
      float d[8];
      load8(DY + r * p.Cw + sub * 8, d);
#pragma unroll
      for (int k = 0; k < kMaxSmallC; ++k) {
        if (k < p.Ck) {
          const float* wr = ws + k * p.Cw + sub * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[k] += d[e] * wr[e];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxSmallC; ++k) {
      if (k < p.Ck) {
        for (int o = c8n >> 1; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      }
    }
    if (r < p.rows && sub == 0) {
#pragma unroll
      for (int k = 0; k < kMaxSmallC; ++k)
        if (k < p.Ck) P[r * p.Ck + k] = from_f32<T>(acc[k]);
    }
  }
}

// ---- wgrad: thread = (8 channels of dy, one of 32 row lanes); block partials in shared memory, then atomics ----
template <typename T>
__global__ void __launch_bounds__(256) smallc_wgrad_kernel(const SmallCParams p, long long rows_per_block) {
  extern __shared__ float red[];  // [Ck][Cw]
  for (int i = threadIdx.x; i < p.Ck * p.Cw; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const T* __restrict__ X = reinterpret_cast<const T*>(p.x);
  const T* __restrict__ DY = reinterpret_cast<const T*>(p.y);
  const int c8n = p.Cw >> 3;
  const int sub = threadIdx.x % c8n, rl = threadIdx.x / c8n, nrl = blockDim.x / c8n;
  float acc[kMaxSmallC][8];
#pragma unroll
  for (int k = 0; k < kMaxSmallC; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  const long long r_end = min(p.rows, r_begin + rows_per_block);
  for (long long r = r_begin + rl; r < r_end; r += nrl) {
    float d[8];
    load8(DY + r * p.Cw + sub * 8, d);
    const T* xr = X + r * p.Ck;
#pragma unroll
    for (int k = 0; k < kMaxSmallC; ++k) {
      if (k < p.Ck) {
        const float xv = to_f32(xr[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[k][e] += xv * d[e];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxSmallC; ++k) {
    if (k < p.Ck) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&red[k * p.Cw + sub * 8 + e], acc[k][e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.Ck * p.Cw; i += blockDim.x) atomicAdd(p.dw + woff(p, i % p.Cw, i / p.Cw), red[i]);
}

static int check_smallc(const SmallCParams& p, int dtype, const char* who) {
  FMM_CHECK_ARG(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "%s: bad dtype %d", who, dtype);
  FMM_CHECK_ARG(p.rows > 0 && p.V > 0 && p.Ck > 0 && p.Ck <= kMaxSmallC && p.K2 > 0, "%s: narrow side must have 1..%d channels, got %d",
                who, kMaxSmallC, p.Ck);
  const int c8n = p.Cw >> 3;
  FMM_CHECK_ARG(p.Cw >= 8 && (p.Cw % 8) == 0 && c8n <= 32 && (c8n & (c8n - 1)) == 0,
                "%s: wide side must be 8, 16, 32, 64, 128 or 256 channels, got %d", who, p.Cw);
  return FMM_OK;
}

}  // namespace fmm

using namespace fmm;

extern "C" {

// out[r][co] = bias[v][co] + sum_k x[r][k] W[co][k];  x [rows][Ck], out [rows][Cw], rows = N*T*V
int fmm_smallc_fwd(const void* x, void* out, const float* w, const float* bias, int bias_per_joint, long long rows, int V,
                   int Ck, int Cw, int K2, long long s_co, long long s_k1, long long s_k2, int dtype, cudaStream_t stream) {
  SmallCParams p{x, out, w, nullptr, bias, rows, V, Ck, Cw, K2, bias_per_joint ? Cw : 0, s_co, s_k1, s_k2};
  FMM_CHECK_ARG(x && out && w, "smallc_fwd: null pointer");
  int rc = check_smallc(p, dtype, "smallc_fwd");
  if (rc != FMM_OK) return rc;
  const size_t smem = sizeof(float) * Ck * Cw;
  long long blocks = (rows * (Cw >> 3) + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == FMM_DT_BF16) smallc_fwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, stream>>>(p);
  else smallc_fwd_kernel<float><<<(unsigned)blocks, 256, smem, stream>>>(p);
  FMM_CHECK_LAUNCH("smallc_fwd");
  return FMM_OK;
}

// p[r][k] = sum_co dy[r][co] W[co][k];  dy [rows][Cw], p [rows][Ck]
int fmm_smallc_dgrad(const void* dy, void* pout, const float* w, long long rows, int Ck, int Cw, int K2, long long s_co,
                     long long s_k1, long long s_k2, int dtype, cudaStream_t stream) {
  SmallCParams p{pout, dy, w, nullptr, nullptr, rows, 1, Ck, Cw, K2, 0, s_co, s_k1, s_k2};
  FMM_CHECK_ARG(dy && pout && w, "smallc_dgrad: null pointer");
  int rc = check_smallc(p, dtype, "smallc_dgrad");
  if (rc != FMM_OK) return rc;
  const size_t smem = sizeof(float) * Ck * Cw;
  const int rpb = 256 / (Cw >> 3);
  long long blocks = (rows + rpb - 1) / rpb;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == FMM_DT_BF16) smallc_dgrad_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, stream>>>(p);
  else smallc_dgrad_kernel<float><<<(unsigned)blocks, 256, smem, stream>>>(p);
  FMM_CHECK_LAUNCH("smallc_dgrad");
  return FMM_OK;
}

// dW[co][k] += sum_r dy[r][co] x[r][k]   (fp32 atomics into dw addressed like w: zero it first)
int fmm_smallc_wgrad(const void* x, const void* dy, float* dw, long long rows, int Ck, int Cw, int K2, long long s_co,
                     long long s_k1, long long s_k2, int dtype, cudaStream_t stream) {
  SmallCParams p{x, dy, nullptr, dw, nullptr, rows, 1, Ck, Cw, K2, 0, s_co, s_k1, s_k2};
  FMM_CHECK_ARG(x && dy && dw, "smallc_wgrad: null pointer");
  int rc = check_smallc(p, dtype, "smallc_wgrad");
  if (rc != FMM_OK) return rc;
  const size_t smem = sizeof(float) * Ck * Cw;
  long long blocks = (long long)num_sms() * 4;
  long long rpb = (rows + blocks - 1) / blocks;
  if (rpb < 32) rpb = 32;
  blocks = (rows + rpb - 1) / rpb;
  if (dtype == FMM_DT_BF16) smallc_wgrad_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, stream>>>(p, rpb);
  else smallc_wgrad_kernel<float><<<(unsigned)blocks, 256, smem, stream>>>(p, rpb);
  FMM_CHECK_LAUNCH("smallc_wgrad");
  return FMM_OK;
}

}  // extern "C"
