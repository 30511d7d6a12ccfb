// Memory-bound kernels of the time-axis transformer of the TRAGCN family (reference: TA.py:40-108):
// row softmax, LayerNorm(a+b), positional encoding, ReLU masking, the time<->feature transpose.
// The matrix products in between run in bgemm.cu / tapconv.cu / wgrad.cu; the graph-GRU glue is gru_cell.cu.
#include "common.cuh"

namespace fmm {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dsilu_(float x) {
  float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}

// ---------------------------------------------------------------------------------------------
// row softmax over the first L entries of rows with pitch Lp (the pad is written as 0); one warp per row
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void softmax_fwd_kernel(T* __restrict__ x, long long rows, int L, int Lp) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  T* p = x + row * Lp;
  float mx = -INFINITY;
  for (int i = lane; i < L; i += 32) mx = fmaxf(mx, to_f32(p[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int i = lane; i < L; i += 32) s += __expf(to_f32(p[i]) - mx);
  s = warp_sum(s);
  float inv = 1.f / s;
  for (int i = lane; i < Lp; i += 32) p[i] = from_f32<T>(i < L ? __expf(to_f32(p[i]) - mx) * inv : 0.f);
}

// dx = p * (dp - sum(p*dp)), written over dp
template <typename T>
__global__ void softmax_bwd_kernel(const T* __restrict__ pr, T* __restrict__ dp, long long rows, int L, int Lp) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  const T* p = pr + row * Lp;
  T* d = dp + row * Lp;
  float s = 0.f;
  for (int i = lane; i < L; i += 32) s += to_f32(p[i]) * to_f32(d[i]);
  s = warp_sum(s);
  for (int i = lane; i < Lp; i += 32) d[i] = from_f32<T>(i < L ? to_f32(p[i]) * (to_f32(d[i]) - s) : 0.f);
}

// ---- vectorised fast paths: every global access is a 16-byte (bf16) / 32-byte (fp32) vector, each row is read once ----
// softmax: one warp per row, up to 2 vectors (16 values) per lane held in registers (Lp <= 512, Lp % 8 == 0)
template <typename T>
__global__ void softmax_fwd_vec_kernel(T* __restrict__ x, long long rows, int L, int Lp) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nv = Lp >> 3;
  T* p = x + row * Lp;
  float v[2][8];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nv) load8(p + vi * 8, v[k]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (vi >= nv || vi * 8 + e >= L) v[k][e] = -INFINITY;
      mx = fmaxf(mx, v[k][e]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[k][e] = __expf(v[k][e] - mx);   // exp(-inf) = 0 for the masked tail
      s += v[k][e];
    }
  const float inv = 1.f / warp_sum(s);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nv) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[k][e] *= inv;
      store8(p + vi * 8, v[k]);
    }
  }
}

template <typename T>
__global__ void softmax_bwd_vec_kernel(const T* __restrict__ pr, T* __restrict__ dp, long long rows, int L, int Lp) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nv = Lp >> 3;
  const T* p = pr + row * Lp;
  T* d = dp + row * Lp;
  float pv[2][8], dv[2][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nv) {
      load8(p + vi * 8, pv[k]);
      load8(d + vi * 8, dv[k]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (vi >= nv || vi * 8 + e >= L) pv[k][e] = dv[k][e] = 0.f;
      s = fmaf(pv[k][e], dv[k][e], s);
    }
  }
  s = warp_sum(s);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nv) {
#pragma unroll
      for (int e = 0; e < 8; ++e) dv[k][e] = pv[k][e] * (dv[k][e] - s);
      store8(d + vi * 8, dv[k]);
    }
  }
}

// LayerNorm, C = 8 * LPR with LPR (lanes per row) a power of two <= 32: 32/LPR rows per warp, reductions over LPR lanes
template <typename T>
__global__ void ln_fwd_vec_kernel(const T* __restrict__ a, const T* __restrict__ b2, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean,
                                  float* __restrict__ rstd, long long rows, int C, float eps) {
  const int lpr = C >> 3, rpw = 32 / lpr;
  const int lane = threadIdx.x & 31, sub = lane % lpr;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw + lane / lpr;
  const bool ok = row < rows;
  float v[8], g[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (ok) {
    load8(a + row * C + sub * 8, v);
    if (b2) {
      float w[8];
      load8(b2 + row * C + sub * 8, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = to_f32(from_f32<T>(v[e] + w[e]));  // the sum is a rounded tensor in the reference
    }
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += v[e];
  for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / C;
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) q += (v[e] - mu) * (v[e] - mu);
  for (int o = lpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = rsqrtf(q / C + eps);
  if (!ok) return;
  if (sub == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  load8(gamma + sub * 8, g);
  load8(beta + sub * 8, be);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (v[e] - mu) * rs * g[e] + be[e];
  store8(y + row * C + sub * 8, v);
}

template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ a,
                                                         const T* __restrict__ b2, const float* __restrict__ gamma,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         T* __restrict__ dx, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, long long rows, int C, int iters) {
  __shared__ float sg[256][9], sb[256][9];
  const int lpr = C >> 3, rpw = 32 / lpr;
  const int lane = threadIdx.x & 31, sub = lane % lpr;
  float g[8], ag[8], ab[8];
  load8(gamma + sub * 8, g);
#pragma unroll
  for (int e = 0; e < 8; ++e) ag[e] = ab[e] = 0.f;
  const long long rows_per_block = (long long)(blockDim.x >> 5) * rpw * iters;
  for (int it = 0; it < iters; ++it) {
    const long long row = (long long)blockIdx.x * rows_per_block + ((long long)it * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw + lane / lpr;
    const bool ok = row < rows;
    float v[8], d[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = d[e] = 0.f;
    float mu = 0.f, rs = 0.f;
    if (ok) {
      load8(a + row * C + sub * 8, v);
      if (b2) {
        float w[8];
        load8(b2 + row * C + sub * 8, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = to_f32(from_f32<T>(v[e] + w[e]));
      }
      load8(dy + row * C + sub * 8, d);
      mu = mean[row];
      rs = rstd[row];
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = (v[e] - mu) * rs;            // xhat
      ag[e] = fmaf(d[e], v[e], ag[e]);
      ab[e] += d[e];
      d[e] *= g[e];
      s1 += d[e];
      s2 = fmaf(d[e], v[e], s2);
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= C;
    s2 /= C;
    if (ok) {
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = rs * (d[e] - s1 - v[e] * s2);
      store8(dx + row * C + sub * 8, d);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sg[threadIdx.x][e] = ag[e];
    sb[threadIdx.x][e] = ab[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int s8 = c >> 3, e = c & 7;
    float gs = 0.f, bs = 0.f;
    for (int t = s8; t < (int)blockDim.x; t += lpr) {   // all threads whose channel group is s8
      gs += sg[t][e];
      bs += sb[t][e];
    }
    atomicAdd(&dgamma[c], gs);
    atomicAdd(&dbeta[c], bs);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over C (<= 256, multiple of 32) of (a + b); one warp per row
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void ln_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b2, const float* __restrict__ gamma,
                              const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean,
                              float* __restrict__ rstd, long long rows, int C, float eps) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31, per = C >> 5;
  float v[8];
  float s = 0.f;
  for (int i = 0; i < per; ++i) {
    long long o = row * C + lane + 32 * i;
    v[i] = to_f32(a[o]) + (b2 ? to_f32(b2[o]) : 0.f);
    if (b2) v[i] = to_f32(from_f32<T>(v[i]));  // the sum is a rounded tensor in the reference
    s += v[i];
  }
  float mu = warp_sum(s) / C;
  float q = 0.f;
  for (int i = 0; i < per; ++i) q += (v[i] - mu) * (v[i] - mu);
  float rs = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  for (int i = 0; i < per; ++i) {
    int c = lane + 32 * i;
    y[row * C + c] = from_f32<T>((v[i] - mu) * rs * gamma[c] + beta[c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ a,
                                                     const T* __restrict__ b2, const float* __restrict__ gamma,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     T* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, long long rows, int C, int rows_per_warp) {
  __shared__ float sg[8][256], sb[8][256];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, per = C >> 5;
  float ag[8], ab[8];
  for (int i = 0; i < 8; ++i) ag[i] = ab[i] = 0.f;
  long long r0 = ((long long)blockIdx.x * 8 + warp) * rows_per_warp;
  for (long long row = r0; row < r0 + rows_per_warp && row < rows; ++row) {
    float xh[8], g[8];
    float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < per; ++i) {
      int c = lane + 32 * i;
      long long o = row * C + c;
      float v = to_f32(a[o]) + (b2 ? to_f32(b2[o]) : 0.f);
      if (b2) v = to_f32(from_f32<T>(v));
      xh[i] = (v - mu) * rs;
      float d = to_f32(dy[o]);
      g[i] = d * gamma[c];
      ag[i] += d * xh[i];
      ab[i] += d;
      s1 += g[i];
      s2 += g[i] * xh[i];
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int i = 0; i < per; ++i) dx[row * C + lane + 32 * i] = from_f32<T>(rs * (g[i] - s1 - xh[i] * s2));
  }
  for (int i = 0; i < per; ++i) {
    sg[warp][lane + 32 * i] = ag[i];
    sb[warp][lane + 32 * i] = ab[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g = 0.f, bb = 0.f;
    for (int w = 0; w < 8; ++w) {
      g += sg[w][c];
      bb += sb[w][c];
    }
    atomicAdd(&dgamma[c], g);
    atomicAdd(&dbeta[c], bb);
  }
}

// y[b,t,v,c] = x[b,t,v,c] + pe[t,c]
template <typename T>
__global__ void add_pe_kernel(const T* __restrict__ x, const float* __restrict__ pe, T* __restrict__ y, long long total,
                              int Tn, int V, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int t = (int)((i / ((long long)C * V)) % Tn);
    y[i] = from_f32<T>(to_f32(x[i]) + pe[t * C + c]);
  }
}

template <typename T>
__global__ void relu_mask_kernel(T* __restrict__ dx, const T* __restrict__ y, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    if (!(to_f32(y[i]) > 0.f)) dx[i] = from_f32<T>(0.f);
}

// batched 2-D transpose: in[g][r][c] (row stride in_rs, c contiguous) -> out[g][c][r] (row stride out_rs),
// r in [0, Rp) with zeros for r >= R.  g = (g1, g2) with separate strides on both sides.
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int Cc, int Rp,
                                                        long long in_g1, long long in_g2, long long in_rs,
                                                        long long out_g1, long long out_g2, long long out_rs, int G2,
                                                        int tiles_r, int tiles_c) {
  __shared__ float tile[32][33];
  int bid = blockIdx.x;
  const int tr = bid % tiles_r;
  bid /= tiles_r;
  const int tc = bid % tiles_c;
  const int g = bid / tiles_c;
  const int g1 = g / G2, g2 = g % G2;
  const T* src = in + g1 * in_g1 + g2 * in_g2;
  T* dst = out + g1 * out_g1 + g2 * out_g2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    int r = tr * 32 + i, c = tc * 32 + tx;
    tile[i][tx] = (r < R && c < Cc) ? to_f32(src[r * in_rs + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int c = tc * 32 + i, r = tr * 32 + tx;
    if (c < Cc && r < Rp) dst[c * out_rs + r] = from_f32<T>(tile[tx][i]);
  }
}

static int ew_blocks(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)num_sms() * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace fmm

using namespace fmm;
#define TG_DISPATCH(dtype, ...)                                   \
  if ((dtype) == FMM_DT_BF16) {                                   \
    using T = __nv_bfloat16;                                      \
    __VA_ARGS__                                                   \
  } else {                                                        \
    using T = float;                                              \
    __VA_ARGS__                                                   \
  }
#define TG_CHECK_DT(dtype, name) FMM_CHECK_ARG((dtype) == FMM_DT_BF16 || (dtype) == FMM_DT_F32, name ": bad dtype %d", dtype)

extern "C" {

int fmm_tg_softmax_fwd(void* x, long long rows, int L, int Lp, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_softmax_fwd");
  FMM_CHECK_ARG(rows > 0 && L > 0 && Lp >= L, "tg_softmax_fwd: bad shape");
  const bool vec = (Lp % 8) == 0 && Lp <= 512 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec) {
    TG_DISPATCH(dtype, softmax_fwd_vec_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((T*)x, rows, L, Lp);)
  } else {
    TG_DISPATCH(dtype, softmax_fwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((T*)x, rows, L, Lp);)
  }
  FMM_CHECK_LAUNCH("tg_softmax_fwd");
  return FMM_OK;
}

int fmm_tg_softmax_bwd(const void* p, void* dp, long long rows, int L, int Lp, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_softmax_bwd");
  FMM_CHECK_ARG(rows > 0 && L > 0 && Lp >= L, "tg_softmax_bwd: bad shape");
  const bool vec = (Lp % 8) == 0 && Lp <= 512 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(dp)) & 15) == 0;
  if (vec) {
    TG_DISPATCH(dtype, softmax_bwd_vec_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const T*)p, (T*)dp, rows, L, Lp);)
  } else {
    TG_DISPATCH(dtype, softmax_bwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const T*)p, (T*)dp, rows, L, Lp);)
  }
  FMM_CHECK_LAUNCH("tg_softmax_bwd");
  return FMM_OK;
}

int fmm_tg_ln_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                  long long rows, int C, float eps, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_ln_fwd");
  FMM_CHECK_ARG(rows > 0 && C >= 32 && C <= 256 && (C % 32) == 0, "tg_ln_fwd: C must be a multiple of 32 up to 256, got %d", C);
  const int lpr = C / 8;
  const bool vec = (C % 8) == 0 && lpr <= 32 && (lpr & (lpr - 1)) == 0 &&
                   ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (vec) {
    const long long rpb = 8ll * (32 / lpr);
    TG_DISPATCH(dtype, ln_fwd_vec_kernel<T><<<(unsigned)((rows + rpb - 1) / rpb), 256, 0, (cudaStream_t)stream>>>(
        (const T*)a, (const T*)b, gamma, beta, (T*)y, mean, rstd, rows, C, eps);)
  } else {
    TG_DISPATCH(dtype, ln_fwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        (const T*)a, (const T*)b, gamma, beta, (T*)y, mean, rstd, rows, C, eps);)
  }
  FMM_CHECK_LAUNCH("tg_ln_fwd");
  return FMM_OK;
}

int fmm_tg_ln_bwd(const void* dy, const void* a, const void* b, const float* gamma, const float* mean, const float* rstd,
                  void* dx, float* dgamma, float* dbeta, long long rows, int C, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_ln_bwd");
  FMM_CHECK_ARG(rows > 0 && C >= 32 && C <= 256 && (C % 32) == 0, "tg_ln_bwd: C must be a multiple of 32 up to 256, got %d", C);
  const int lpr = C / 8;
  const bool vec = (C % 8) == 0 && lpr <= 32 && (lpr & (lpr - 1)) == 0 &&
                   ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(dy) |
                     reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (vec) {
    const long long rows_per_pass = 8ll * (32 / lpr);                       // rows a 256-thread block covers per iteration
    long long blocks = (long long)num_sms() * 16;
    int iters = (int)((rows + blocks * rows_per_pass - 1) / (blocks * rows_per_pass));
    if (iters < 1) iters = 1;
    blocks = (rows + rows_per_pass * iters - 1) / (rows_per_pass * iters);
    TG_DISPATCH(dtype, ln_bwd_vec_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const T*)dy, (const T*)a, (const T*)b, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows, C, iters);)
  } else {
    long long warps = (long long)num_sms() * 8 * 4;
    int rpw = (int)((rows + warps - 1) / warps);
    if (rpw < 1) rpw = 1;
    long long blocks = (rows + 8ll * rpw - 1) / (8ll * rpw);
    TG_DISPATCH(dtype, ln_bwd_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const T*)dy, (const T*)a, (const T*)b, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows, C, rpw);)
  }
  FMM_CHECK_LAUNCH("tg_ln_bwd");
  return FMM_OK;
}

int fmm_tg_add_pe(const void* x, const float* pe, void* y, int B, int Tn, int V, int C, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_add_pe");
  long long total = (long long)B * Tn * V * C;
  FMM_CHECK_ARG(total > 0, "tg_add_pe: empty");
  TG_DISPATCH(dtype, add_pe_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, pe, (T*)y, total, Tn, V, C);)
  FMM_CHECK_LAUNCH("tg_add_pe");
  return FMM_OK;
}

int fmm_tg_relu_mask(void* dx, const void* y, long long total, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_relu_mask");
  FMM_CHECK_ARG(total > 0, "tg_relu_mask: empty");
  TG_DISPATCH(dtype, relu_mask_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>((T*)dx, (const T*)y, total);)
  FMM_CHECK_LAUNCH("tg_relu_mask");
  return FMM_OK;
}

int fmm_tg_transpose(const void* in, void* out, int R, int Cc, int Rp, long long in_g1, long long in_g2, long long in_rs,
                     long long out_g1, long long out_g2, long long out_rs, int G1, int G2, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_transpose");
  FMM_CHECK_ARG(R > 0 && Cc > 0 && Rp >= R && G1 > 0 && G2 > 0, "tg_transpose: bad shape");
  const int tiles_r = (Rp + 31) / 32, tiles_c = (Cc + 31) / 32;
  long long blocks = (long long)tiles_r * tiles_c * G1 * G2;
  FMM_CHECK_ARG(blocks < (1ll << 31), "tg_transpose: too many tiles");
  TG_DISPATCH(dtype, transpose_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const T*)in, (T*)out, R, Cc, Rp, in_g1, in_g2, in_rs, out_g1, out_g2, out_rs, G2, tiles_r, tiles_c);)
  FMM_CHECK_LAUNCH("tg_transpose");
  return FMM_OK;
}

}  // extern "C"
