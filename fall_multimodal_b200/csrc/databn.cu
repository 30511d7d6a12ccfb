// data_bn: the input pre-normalisation of STGCAN.forward (reference stgcan.py:212-218).
//
// The reference permutes the clip (N,C,T,V) -> (N,V,C,T), views it as (N, V*C, T), applies BatchNorm1d(V*C) over (N,T)
// (channel index v*C + c) and permutes back: two .contiguous() copies around cuDNN batch norm. Here the (N,C,T,V) fp32 input
// is read in place and the normalised clip is written ONCE, already in the channels-last (N,T,V,C) activation layout and
// dtype the graph-conv kernels consume:
//
//   databn_stats   per-(v,c) sum / sum of squares over (n,t)            -> fp64 [V*C] x 2     (then fmm_bn_finalize)
//   databn_apply   y[n][t][v][c] = x[n][c][t][v] * a[v*C+c] + b[v*C+c]                         (layout change fused)
//   databn_bwd     dgamma[v*C+c] = sum_{n,t} dy * xhat,  dbeta[v*C+c] = sum_{n,t} dy           (the clip needs no gradient)
//
// Memory bound and tiny (a clip is 6 336 elements); the point is one pass and no torch glue in the captured step.
#include "common.cuh"

namespace fmm {

// grid = (N*C), block = 128: thread v-lane walks the (T,V) plane of one (n,c) with stride blockDim (coalesced over v)
__global__ void databn_stats_kernel(const float* __restrict__ x, double* __restrict__ sum, double* __restrict__ sq, int C, int Tn, int V) {
  extern __shared__ float red[];   // [2][V]
  const int n = blockIdx.x / C, c = blockIdx.x % C;
  const float* plane = x + (static_cast<size_t>(n) * C + c) * Tn * V;
  for (int i = threadIdx.x; i < 2 * V; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  // thread owns plane elements i = tid, tid + B, ...; with B a multiple of... V is arbitrary, so accumulate per element
  // into the (v) slot through shared-memory atomics after a register pre-reduction over the rows this thread sees of one v
  const int total = Tn * V;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {   // V <= blockDim: one pass
    float s = 0.f, q = 0.f;
    for (int t = 0; t < Tn; ++t) {
      const float val = plane[t * V + v];
      s += val;
      q = fmaf(val, val, q);
    }
    red[v] = s;
    red[V + v] = q;
  }
  (void)total;
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    atomicAdd(sum + v * C + c, static_cast<double>(red[v]));
    atomicAdd(sq + v * C + c, static_cast<double>(red[V + v]));
  }
}

// grid = (T chunks, N): thread = flat (t, v, c) output element of its chunk (contiguous, coalesced writes)
template <typename T>
__global__ void databn_apply_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                    T* __restrict__ y, int C, int Tn, int V, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk, t1 = min(t0 + tchunk, Tn);
  const int per_t = V * C;
  const float* xin = x + static_cast<size_t>(n) * C * Tn * V;
  T* yo = y + static_cast<size_t>(n) * Tn * per_t;
  for (int i = t0 * per_t + threadIdx.x; i < t1 * per_t; i += blockDim.x) {
    const int t = i / per_t, vc = i - t * per_t;
    const int v = vc / C, c = vc - v * C;
    yo[i] = from_f32<T>(fmaf(xin[(static_cast<size_t>(c) * Tn + t) * V + v], a[vc], b[vc]));
  }
}

// grid = (T chunks, N), block = 128 >= V*C: thread = (v,c) pair, walks the frames of its chunk
template <typename T>
__global__ void databn_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, double* __restrict__ dgamma, double* __restrict__ dbeta, int C,
                                  int Tn, int V, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk, t1 = min(t0 + tchunk, Tn);
  const int per_t = V * C;
  for (int vc = threadIdx.x; vc < per_t; vc += blockDim.x) {
    const int v = vc / C, c = vc - v * C;
    const float m = mean[vc], r = rstd[vc];
    float g = 0.f, s = 0.f;
    for (int t = t0; t < t1; ++t) {
      const float d = to_f32(dy[(static_cast<size_t>(n) * Tn + t) * per_t + vc]);
      const float xv = x[((static_cast<size_t>(n) * C + c) * Tn + t) * V + v];
      g = fmaf(d, (xv - m) * r, g);
      s += d;
    }
    atomicAdd(dgamma + vc, static_cast<double>(g));
    atomicAdd(dbeta + vc, static_cast<double>(s));
  }
}

}  // namespace fmm

using namespace fmm;

extern "C" {

// sum / sq: fp64 [V*C], zeroed by the caller, accumulated into; feed them to fmm_bn_finalize(count = N*T, C = V*C)
int fmm_databn_stats(const float* x, double* sum, double* sq, int N, int C, int T, int V, cudaStream_t stream) {
  FMM_CHECK_ARG(x && sum && sq && N > 0 && C > 0 && T > 0 && V > 0 && V <= 128, "databn_stats: bad arguments");
  databn_stats_kernel<<<N * C, 128, 2 * V * sizeof(float), stream>>>(x, sum, sq, C, T, V);
  FMM_CHECK_LAUNCH("databn_stats");
  return FMM_OK;
}

int fmm_databn_apply(const float* x, const float* a, const float* b, void* y, int N, int C, int T, int V, int dtype,
                     cudaStream_t stream) {
  FMM_CHECK_ARG(x && a && b && y && N > 0 && C > 0 && T > 0 && V > 0, "databn_apply: bad arguments");
  FMM_CHECK_ARG(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "databn_apply: bad dtype");
  const int tchunk = 8;
  dim3 grid((T + tchunk - 1) / tchunk, N);
  if (dtype == FMM_DT_BF16)
    databn_apply_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x, a, b, reinterpret_cast<__nv_bfloat16*>(y), C, T, V, tchunk);
  else
    databn_apply_kernel<float><<<grid, 256, 0, stream>>>(x, a, b, reinterpret_cast<float*>(y), C, T, V, tchunk);
  FMM_CHECK_LAUNCH("databn_apply");
  return FMM_OK;
}

// dgamma / dbeta: fp64 [V*C], zeroed by the caller
int fmm_databn_bwd(const void* dy, const float* x, const float* mean, const float* rstd, double* dgamma, double* dbeta, int N,
                   int C, int T, int V, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(dy && x && mean && rstd && dgamma && dbeta && N > 0 && C > 0 && T > 0 && V > 0, "databn_bwd: bad arguments");
  FMM_CHECK_ARG(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "databn_bwd: bad dtype");
  const int tchunk = 16;
  dim3 grid((T + tchunk - 1) / tchunk, N);
  if (dtype == FMM_DT_BF16)
    databn_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), x, mean, rstd, dgamma, dbeta, C, T, V, tchunk);
  else
    databn_bwd_kernel<float><<<grid, 128, 0, stream>>>(reinterpret_cast<const float*>(dy), x, mean, rstd, dgamma, dbeta, C, T, V, tchunk);
  FMM_CHECK_LAUNCH("databn_bwd");
  return FMM_OK;
}

}  // extern "C"
