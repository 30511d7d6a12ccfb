// Memory-bound kernels of the musa `Model` (Multimodal_Fall3/model/musa_model.py), channels-last [N][T][V][C], C % 8 == 0:
//   dwconv_*      depthwise (k x 1) temporal convolution, groups = C (SepTemporal_Block.depth_conv :165-168,
//                 DepthWiseSeparableConv_{3x1,1x1}_1x1 :426, :446): forward, data gradient, weight + bias gradient
//   affine_act    y = act(a[c]*x + b[c] (+ res)) — a folded BatchNorm2d followed by the residual add and the activation
//                 (tanh / relu / leaky-relu 0.01 / identity), e.g. act(bn(x) + res) of SpatialGraphConv.forward :144-146
//   bn_act_bwd_*  its backward: dz = dy*act'(y); per-channel sum(dz), sum(dz*xhat); dx = a*(dz - mean(dz) - xhat*mean(dz*xhat))
//                 (train) or a*dz (eval); dres = dz
// The 1x1 convolutions in between are strided batched GEMMs (bgemm.cu).
#include "common.cuh"

namespace fmm {

constexpr int kMaxTaps = 9;

__device__ __forceinline__ float act_fwd(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return tanhf(v);
    case 3: return v > 0.f ? v : 0.01f * v;
    default: return v;
  }
}
// derivative expressed through the OUTPUT y
__device__ __forceinline__ float act_bwd(float y, int act) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : 0.f;
    case 2: return 1.f - y * y;
    case 3: return y > 0.f ? 1.f : 0.01f;
    default: return 1.f;
  }
}

// ---- depthwise temporal conv: out[n,to,v,c] = b[c] + sum_j w[c][j] * x[n, to*s + j - pad, v, c] ----
template <typename T>
__global__ void dwconv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                  T* __restrict__ out, int N, int Tin, int Tout, int V, int C, int k, int stride, int pad) {
  const int c8n = C >> 3;
  const long long items = (long long)N * Tout * V * c8n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(it % c8n) * 8;
    long long r = it / c8n;
    const int v = (int)(r % V);
    r /= V;
    const int to = (int)(r % Tout);
    const int n = (int)(r / Tout);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = b ? b[c0 + e] : 0.f;
    for (int j = 0; j < k; ++j) {
      const int t = to * stride + j - pad;
      if (t < 0 || t >= Tin) continue;
      float f[8];
      load8(x + (((long long)n * Tin + t) * V + v) * C + c0, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(w[(c0 + e) * k + j], f[e], acc[e]);
    }
    store8(out + (((long long)n * Tout + to) * V + v) * C + c0, acc);
  }
}

// dx[n,t,v,c] = sum_j w[c][j] * dy[n,(t + pad - j)/s, v, c]  (when divisible and in range)
template <typename T>
__global__ void dwconv_bwd_data_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, int N,
                                       int Tin, int Tout, int V, int C, int k, int stride, int pad) {
  const int c8n = C >> 3;
  const long long items = (long long)N * Tin * V * c8n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(it % c8n) * 8;
    long long r = it / c8n;
    const int v = (int)(r % V);
    r /= V;
    const int t = (int)(r % Tin);
    const int n = (int)(r / Tin);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int j = 0; j < k; ++j) {
      const int num = t + pad - j;
      if (num < 0 || (num % stride) != 0) continue;
      const int to = num / stride;
      if (to >= Tout) continue;
      float f[8];
      load8(dy + (((long long)n * Tout + to) * V + v) * C + c0, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(w[(c0 + e) * k + j], f[e], acc[e]);
    }
    store8(dx + (((long long)n * Tin + t) * V + v) * C + c0, acc);
  }
}

// dw[c][j] += sum dy[n,to,v,c]*x[n,to*s+j-pad,v,c] ; db[c] += sum dy.  Block = (n, a chunk of output frames); thread =
// (8 channels, row lane); partial sums combined in shared memory, one atomic per (c, j) per block.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_bwd_weight_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                float* __restrict__ dw, float* __restrict__ db, int Tin,
                                                                int Tout, int V, int C, int k, int stride, int pad, int tchunk) {
  extern __shared__ float red[];  // [C][k+1]
  const int n = blockIdx.y;
  const int c8n = C >> 3;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  for (int i = threadIdx.x; i < C * (k + 1); i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[kMaxTaps + 1][8];
#pragma unroll
  for (int j = 0; j <= kMaxTaps; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  const int r0 = blockIdx.x * tchunk * V, r1 = min((blockIdx.x + 1) * tchunk, Tout) * V;
  for (int r = r0 + rl; r < r1; r += RL) {
    const int to = r / V, v = r % V;
    float d[8];
    load8(dy + (((long long)n * Tout + to) * V + v) * C + c8 * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[kMaxTaps][e] += d[e];
#pragma unroll
    for (int j = 0; j < kMaxTaps; ++j) {
      if (j < k) {
        const int t = to * stride + j - pad;
        if (t >= 0 && t < Tin) {
          float f[8];
          load8(x + (((long long)n * Tin + t) * V + v) * C + c8 * 8, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(d[e], f[e], acc[j][e]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxTaps; ++j)
    if (j < k)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&red[(c8 * 8 + e) * (k + 1) + j], acc[j][e]);
#pragma unroll
  for (int e = 0; e < 8; ++e) atomicAdd(&red[(c8 * 8 + e) * (k + 1) + k], acc[kMaxTaps][e]);
  __syncthreads();
  for (int i = threadIdx.x; i < C * (k + 1); i += blockDim.x) {
    const int c = i / (k + 1), j = i % (k + 1);
    if (j < k) atomicAdd(dw + c * k + j, red[i]);
    else if (db) atomicAdd(db + c, red[i]);
  }
}

// ---- y = act(a*x + b (+ res)) ----
template <typename T>
__global__ void affine_act_kernel(const T* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                  const T* __restrict__ res, T* __restrict__ y, long long rows, int C, int act) {
  const int c8n = C >> 3;
  const long long items = rows * c8n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(it % c8n) * 8;
    float f[8], r[8];
    load8(x + it * 8, f);
    if (res) load8(res + it * 8, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(a[c0 + e], f[e], b[c0 + e]);
      if (res) v += r[e];
      f[e] = act_fwd(v, act);
    }
    store8(y + it * 8, f);
  }
}

// ---- backward reduce: S1[c] += sum dz, S2[c] += sum dz*xhat, dz = dy*act'(y), xhat = (x - mean)*rstd ----
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                                const T* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, double* __restrict__ S1,
                                                                double* __restrict__ S2, long long rows, int C, int act,
                                                                long long rows_per_block) {
  extern __shared__ float red[];  // [C][2]
  const int c8n = C >> 3;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float mu[8], rs[8], a1[8], a2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mu[e] = mean[c8 * 8 + e];
    rs[e] = rstd[c8 * 8 + e];
    a1[e] = a2[e] = 0.f;
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (long long r = r0 + rl; r < r1; r += RL) {
    float d[8], yy[8], xx[8];
    load8(dy + r * C + c8 * 8, d);
    load8(x + r * C + c8 * 8, xx);
    if (act) load8(y + r * C + c8 * 8, yy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = act ? d[e] * act_bwd(yy[e], act) : d[e];
      a1[e] += dz;
      a2[e] = fmaf(dz, (xx[e] - mu[e]) * rs[e], a2[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&red[(c8 * 8 + e) * 2], a1[e]);
    atomicAdd(&red[(c8 * 8 + e) * 2 + 1], a2[e]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(S1 + c, static_cast<double>(red[2 * c]));
    atomicAdd(S2 + c, static_cast<double>(red[2 * c + 1]));
  }
}

// ---- backward apply: dx = a*(dz - m1 - xhat*m2) (training) or a*dz ; dres = dz ----
template <typename T>
__global__ void bn_act_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                                        const float* __restrict__ a, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const double* __restrict__ S1,
                                        const double* __restrict__ S2, double inv_count, int training, T* __restrict__ dx,
                                        T* __restrict__ dres, long long rows, int C, int act) {
  const int c8n = C >> 3;
  const long long items = rows * c8n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(it % c8n) * 8;
    float d[8], yy[8], xx[8], o[8];
    load8(dy + it * 8, d);
    load8(x + it * 8, xx);
    if (act) load8(y + it * 8, yy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = act ? d[e] * act_bwd(yy[e], act) : d[e];
      d[e] = dz;
      if (training) {
        const float m1 = static_cast<float>(S1[c0 + e] * inv_count), m2 = static_cast<float>(S2[c0 + e] * inv_count);
        const float xh = (xx[e] - mean[c0 + e]) * rstd[c0 + e];
        o[e] = a[c0 + e] * (dz - m1 - xh * m2);
      } else {
        o[e] = a[c0 + e] * dz;
      }
    }
    store8(dx + it * 8, o);
    if (dres) store8(dres + it * 8, d);
  }
}

static int ew_grid(long long items) {
  long long b = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace fmm

using namespace fmm;
#define MU_DISPATCH(dtype, ...)   \
  if ((dtype) == FMM_DT_BF16) {   \
    using T = __nv_bfloat16;      \
    __VA_ARGS__                   \
  } else {                        \
    using T = float;              \
    __VA_ARGS__                   \
  }
#define MU_CHECK(cond, ...) FMM_CHECK_ARG(cond, __VA_ARGS__)

extern "C" {

int fmm_dwconv_fwd(const void* x, const float* w, const float* b, void* out, int N, int Tin, int Tout, int V, int C, int k,
                   int stride, int pad, int dtype, cudaStream_t stream) {
  MU_CHECK(x && w && out && N > 0 && Tin > 0 && Tout > 0 && V > 0 && C % 8 == 0 && k >= 1 && k <= kMaxTaps && stride >= 1,
           "dwconv_fwd: bad arguments (C must be a multiple of 8, 1 <= k <= %d)", kMaxTaps);
  MU_CHECK((Tout - 1) * stride - pad + (k - 1) < Tin + pad && (dtype == FMM_DT_BF16 || dtype == FMM_DT_F32), "dwconv_fwd: bad geometry / dtype");
  const long long items = (long long)N * Tout * V * (C / 8);
  MU_DISPATCH(dtype, dwconv_fwd_kernel<T><<<ew_grid(items), 256, 0, stream>>>((const T*)x, w, b, (T*)out, N, Tin, Tout, V, C, k, stride, pad);)
  FMM_CHECK_LAUNCH("dwconv_fwd");
  return FMM_OK;
}

int fmm_dwconv_bwd_data(const void* dy, const float* w, void* dx, int N, int Tin, int Tout, int V, int C, int k, int stride,
                        int pad, int dtype, cudaStream_t stream) {
  MU_CHECK(dy && w && dx && N > 0 && Tin > 0 && Tout > 0 && V > 0 && C % 8 == 0 && k >= 1 && k <= kMaxTaps && stride >= 1,
           "dwconv_bwd_data: bad arguments");
  MU_CHECK(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "dwconv_bwd_data: bad dtype");
  const long long items = (long long)N * Tin * V * (C / 8);
  MU_DISPATCH(dtype, dwconv_bwd_data_kernel<T><<<ew_grid(items), 256, 0, stream>>>((const T*)dy, w, (T*)dx, N, Tin, Tout, V, C, k, stride, pad);)
  FMM_CHECK_LAUNCH("dwconv_bwd_data");
  return FMM_OK;
}

int fmm_dwconv_bwd_weight(const void* x, const void* dy, float* dw, float* db, int N, int Tin, int Tout, int V, int C, int k,
                          int stride, int pad, int dtype, cudaStream_t stream) {
  MU_CHECK(x && dy && dw && N > 0 && N <= 65535 && C % 8 == 0 && C <= 2048 && k >= 1 && k <= kMaxTaps, "dwconv_bwd_weight: bad arguments");
  MU_CHECK(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "dwconv_bwd_weight: bad dtype");
  int tchunk = Tout;
  while (tchunk > 2 && (long long)N * ((Tout + tchunk - 1) / tchunk) < 2LL * num_sms()) tchunk = (tchunk + 1) / 2;
  const int c8n = C / 8;
  const int threads = c8n >= 256 ? c8n : (256 / c8n) * c8n;
  dim3 grid((Tout + tchunk - 1) / tchunk, N);
  const size_t smem = sizeof(float) * C * (k + 1);
  MU_DISPATCH(dtype, dwconv_bwd_weight_kernel<T><<<grid, threads, smem, stream>>>((const T*)x, (const T*)dy, dw, db, Tin, Tout, V, C, k, stride, pad, tchunk);)
  FMM_CHECK_LAUNCH("dwconv_bwd_weight");
  return FMM_OK;
}

int fmm_affine_act(const void* x, const float* a, const float* b, const void* res, void* y, long long rows, int C, int act,
                   int dtype, cudaStream_t stream) {
  MU_CHECK(x && a && b && y && rows > 0 && C % 8 == 0 && act >= 0 && act <= 3, "affine_act: bad arguments");
  MU_CHECK(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "affine_act: bad dtype");
  MU_DISPATCH(dtype, affine_act_kernel<T><<<ew_grid(rows * (C / 8)), 256, 0, stream>>>((const T*)x, a, b, (const T*)res, (T*)y, rows, C, act);)
  FMM_CHECK_LAUNCH("affine_act");
  return FMM_OK;
}

int fmm_bn_act_bwd_reduce(const void* dy, const void* y, const void* x, const float* mean, const float* rstd, double* S1,
                          double* S2, long long rows, int C, int act, int dtype, cudaStream_t stream) {
  MU_CHECK(dy && x && mean && rstd && S1 && S2 && rows > 0 && C % 8 == 0 && C <= 2048 && (act == 0 || y), "bn_act_bwd_reduce: bad arguments");
  MU_CHECK(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "bn_act_bwd_reduce: bad dtype");
  const int c8n = C / 8;
  const int threads = c8n >= 256 ? c8n : (256 / c8n) * c8n;
  long long blocks = (long long)num_sms() * 4;
  long long rpb = (rows + blocks - 1) / blocks;
  if (rpb < 64) rpb = 64;
  blocks = (rows + rpb - 1) / rpb;
  MU_DISPATCH(dtype, bn_act_bwd_reduce_kernel<T><<<(unsigned)blocks, threads, sizeof(float) * 2 * C, stream>>>(
      (const T*)dy, (const T*)y, (const T*)x, mean, rstd, S1, S2, rows, C, act, rpb);)
  FMM_CHECK_LAUNCH("bn_act_bwd_reduce");
  return FMM_OK;
}

int fmm_bn_act_bwd_apply(const void* dy, const void* y, const void* x, const float* a, const float* mean, const float* rstd,
                         const double* S1, const double* S2, double inv_count, int training, void* dx, void* dres,
                         long long rows, int C, int act, int dtype, cudaStream_t stream) {
  MU_CHECK(dy && x && a && mean && rstd && dx && rows > 0 && C % 8 == 0 && (act == 0 || y) && (!training || (S1 && S2)),
           "bn_act_bwd_apply: bad arguments");
  MU_CHECK(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "bn_act_bwd_apply: bad dtype");
  MU_DISPATCH(dtype, bn_act_bwd_apply_kernel<T><<<ew_grid(rows * (C / 8)), 256, 0, stream>>>(
      (const T*)dy, (const T*)y, (const T*)x, a, mean, rstd, S1, S2, inv_count, training, (T*)dx, (T*)dres, rows, C, act);)
  FMM_CHECK_LAUNCH("bn_act_bwd_apply");
  return FMM_OK;
}

}  // extern "C"
