// Accelerometer 1-D CNN branch (notebook CNN1D, /root/reference/GSTCAN_HAR_conv_10kfold.ipynb#cell2:L6-27):
//   [Conv1d(k=5, pad=2) -> BatchNorm1d -> ReLU -> MaxPool1d(2)] x 2
// on channels-last windows x[N][L][C] (the layout the sensor stream already has on the host), fp32.
// The branch is ~0.01% of the step's FLOPs; it is written as direct kernels (one thread per output
// element) so the whole train step stays on the device with no library calls. BatchNorm statistics
// and the BatchNorm/ReLU backward reuse the generic kernels (colstats / bn_finalize / bn1_bwd_*)
// with V = 1.
#include "common.cuh"

namespace fmm {

// y[n][l][co] = b[co] + sum_{dk<5, ci} w[co][ci][dk] * x[n][l+dk-2][ci]
__global__ void conv1d_k5_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ b, float* __restrict__ y, int N, int L, int Ci,
                                     int Co) {
  const long long total = static_cast<long long>(N) * L * Co;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Co);
    const int l = static_cast<int>((i / Co) % L);
    const int n = static_cast<int>(i / (static_cast<long long>(Co) * L));
    float acc = b[co];
    for (int dk = 0; dk < 5; ++dk) {
      const int ll = l + dk - 2;
      if (ll < 0 || ll >= L) continue;
      const float* xr = x + (static_cast<size_t>(n) * L + ll) * Ci;
      const float* wr = w + static_cast<size_t>(co) * Ci * 5 + dk;
      for (int ci = 0; ci < Ci; ++ci) acc = fmaf(wr[ci * 5], xr[ci], acc);
    }
    y[i] = acc;
  }
}

// out[n][j][c] = max(relu(a*y[n][2j][c]+b), relu(a*y[n][2j+1][c]+b)),  j < L/2
__global__ void bn_relu_pool2_fwd_kernel(const float* __restrict__ y, const float* __restrict__ a,
                                         const float* __restrict__ b, float* __restrict__ out, int N, int L, int C) {
  const int Lo = L / 2;
  const long long total = static_cast<long long>(N) * Lo * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int j = static_cast<int>((i / C) % Lo);
    const int n = static_cast<int>(i / (static_cast<long long>(C) * Lo));
    const float* yr = y + (static_cast<size_t>(n) * L + 2 * j) * C + c;
    const float v0 = fmaxf(fmaf(a[c], yr[0], b[c]), 0.f);
    const float v1 = fmaxf(fmaf(a[c], yr[C], b[c]), 0.f);
    out[i] = fmaxf(v0, v1);
  }
}

// Max-pool backward onto the pre-pool grid: dh[n][l][c] = dout[n][l/2][c] if l is the arg-max of its
// pair (first index on ties, as torch), else 0; positions beyond 2*(L/2) get 0.
__global__ void pool2_bwd_kernel(const float* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b,
                                 const float* __restrict__ dout, float* __restrict__ dh, int N, int L, int C) {
  const int Lo = L / 2;
  const long long total = static_cast<long long>(N) * L * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int l = static_cast<int>((i / C) % L);
    const int n = static_cast<int>(i / (static_cast<long long>(C) * L));
    const int j = l >> 1;
    float g = 0.f;
    if (j < Lo) {
      const float* yr = y + (static_cast<size_t>(n) * L + 2 * j) * C + c;
      const float v0 = fmaxf(fmaf(a[c], yr[0], b[c]), 0.f);
      const float v1 = fmaxf(fmaf(a[c], yr[C], b[c]), 0.f);
      const int arg = v1 > v0 ? 1 : 0;
      if ((l & 1) == arg) g = dout[(static_cast<size_t>(n) * Lo + j) * C + c];
    }
    dh[i] = g;
  }
}

// dx[n][l][ci] = sum_{dk, co} w[co][ci][dk] * dy[n][l-dk+2][co]
__global__ void conv1d_k5_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                       float* __restrict__ dx, int N, int L, int Ci, int Co) {
  const long long total = static_cast<long long>(N) * L * Ci;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Ci);
    const int l = static_cast<int>((i / Ci) % L);
    const int n = static_cast<int>(i / (static_cast<long long>(Ci) * L));
    float acc = 0.f;
    for (int dk = 0; dk < 5; ++dk) {
      const int ll = l - dk + 2;
      if (ll < 0 || ll >= L) continue;
      const float* dr = dy + (static_cast<size_t>(n) * L + ll) * Co;
      for (int co = 0; co < Co; ++co) acc = fmaf(w[(static_cast<size_t>(co) * Ci + ci) * 5 + dk], dr[co], acc);
    }
    dx[i] = acc;
  }
}

// dw[co][ci][dk] = sum_{n,l} x[n][l+dk-2][ci] * dy[n][l][co]; db[co] = sum dy   (grid = Co*Ci*5 (+Co))
__global__ void conv1d_k5_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                       float* __restrict__ dw, float* __restrict__ db, int N, int L, int Ci, int Co) {
  __shared__ float red[32];
  const int id = blockIdx.x;
  const int nw = Co * Ci * 5;
  float acc = 0.f;
  const long long rows = static_cast<long long>(N) * L;
  if (id < nw) {
    const int dk = id % 5;
    const int ci = (id / 5) % Ci;
    const int co = id / (5 * Ci);
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
      const int l = static_cast<int>(r % L);
      const int ll = l + dk - 2;
      if (ll < 0 || ll >= L) continue;
      acc = fmaf(x[(r - l + ll) * Ci + ci], dy[r * Co + co], acc);
    }
  } else {
    const int co = id - nw;
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) acc += dy[r * Co + co];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int wv = 0; wv < (blockDim.x >> 5); ++wv) t += red[wv];
    if (id < nw)
      dw[id] = t;
    else
      db[id - nw] = t;
  }
}

static inline int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = 8LL * num_sms();
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fmm

using namespace fmm;

extern "C" {

int fmm_conv1d_k5_fwd(const float* x, const float* w, const float* b, float* y, int N, int L, int Ci, int Co,
                      cudaStream_t stream) {
  FMM_CHECK_ARG(x && w && b && y && N > 0 && L > 0 && Ci > 0 && Co > 0, "conv1d_k5_fwd: bad args");
  conv1d_k5_fwd_kernel<<<grid_for(static_cast<long long>(N) * L * Co), 256, 0, stream>>>(x, w, b, y, N, L, Ci, Co);
  FMM_CHECK_LAUNCH("conv1d_k5_fwd");
  return FMM_OK;
}

int fmm_bn_relu_pool2_fwd(const float* y, const float* a, const float* b, float* out, int N, int L, int C,
                          cudaStream_t stream) {
  FMM_CHECK_ARG(y && a && b && out && N > 0 && L > 1 && C > 0, "bn_relu_pool2_fwd: bad args");
  bn_relu_pool2_fwd_kernel<<<grid_for(static_cast<long long>(N) * (L / 2) * C), 256, 0, stream>>>(y, a, b, out, N, L, C);
  FMM_CHECK_LAUNCH("bn_relu_pool2_fwd");
  return FMM_OK;
}

int fmm_pool2_bwd(const float* y, const float* a, const float* b, const float* dout, float* dh, int N, int L, int C,
                  cudaStream_t stream) {
  FMM_CHECK_ARG(y && a && b && dout && dh && N > 0 && L > 1 && C > 0, "pool2_bwd: bad args");
  pool2_bwd_kernel<<<grid_for(static_cast<long long>(N) * L * C), 256, 0, stream>>>(y, a, b, dout, dh, N, L, C);
  FMM_CHECK_LAUNCH("pool2_bwd");
  return FMM_OK;
}

int fmm_conv1d_k5_bwd(const float* x, const float* dy, const float* w, float* dx, float* dw, float* db, int N, int L,
                      int Ci, int Co, cudaStream_t stream) {
  FMM_CHECK_ARG(x && dy && w && dw && db && N > 0 && L > 0 && Ci > 0 && Co > 0, "conv1d_k5_bwd: bad args");
  if (dx) conv1d_k5_dgrad_kernel<<<grid_for(static_cast<long long>(N) * L * Ci), 256, 0, stream>>>(dy, w, dx, N, L, Ci, Co);
  conv1d_k5_wgrad_kernel<<<Co * Ci * 5 + Co, 256, 0, stream>>>(x, dy, dw, db, N, L, Ci, Co);
  FMM_CHECK_LAUNCH("conv1d_k5_bwd");
  return FMM_OK;
}

}  // extern "C"
