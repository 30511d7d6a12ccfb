// Parameter-side algebra of the graph convolution (reference stgcan.py:222 `self.A * importance`, :50-56): everything that
// depends only on the (K,V,V) adjacency, the learned edge importance and the conv bias - no activations. One launch per
// block and direction instead of ~8 tiny torch kernels each (index / reduce / matmul glue: ~450 of the ~950 launches of a
// round-1 train step).
//
//   forward   coef_f[e] = (A*imp)[dense_idx[e]]            edge coefficients in forward CSR order
//             coef_b[j] = coef_f[bwd_perm[j]]              ... and in out-edge order
//             colsum[k][w] = sum_v (A*imp)[k][v][w]
//             bias_eff[w][co] = sum_k colsum[k][w] * bg[k][co]        (conv bias folded through the aggregation)
//   backward  Tbl[w][co] = sum_rep TblR[rep][w][co]                    (per-joint sums of dG from bn1_bwd_apply)
//             dbg[k][co] = sum_w colsum[k][w] * Tbl[w][co]
//             dimp[k][v][w] = A[k][v][w] * (dA[k][v][w] + sum_co bg[k][co] * Tbl[w][co]),  dA = scatter(dcoef, dense_idx)
#include "common.cuh"

namespace fmm {

__global__ void gcn_prep_fwd_kernel(const float* __restrict__ A, const float* __restrict__ imp, const float* __restrict__ bg,
                                    const long long* __restrict__ dense_idx, const long long* __restrict__ bwd_perm,
                                    float* __restrict__ coef_f, float* __restrict__ coef_b, float* __restrict__ colsum,
                                    float* __restrict__ bias_eff, int K, int V, int Cout, int E) {
  extern __shared__ float cs[];   // [K][V]
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const long long d = dense_idx[e];
    coef_f[e] = A[d] * imp[d];
    const long long db = dense_idx[bwd_perm[e]];
    coef_b[e] = A[db] * imp[db];
  }
  for (int i = threadIdx.x; i < K * V; i += blockDim.x) {
    const int k = i / V, w = i - k * V;
    float s = 0.f;
    for (int v = 0; v < V; ++v) {
      const int d = (k * V + v) * V + w;
      s = fmaf(A[d], imp[d], s);
    }
    cs[i] = s;
    colsum[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * Cout; i += blockDim.x) {
    const int w = i / Cout, co = i - w * Cout;
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(cs[k * V + w], bg[k * Cout + co], s);
    bias_eff[i] = s;
  }
}

// grid = K blocks (one per partition), dynamic smem: Tbl [V][Cout] is recomputed per block (tiny)
__global__ void gcn_prep_bwd_kernel(const float* __restrict__ A, const float* __restrict__ bg, const float* __restrict__ colsum,
                                    const float* __restrict__ TblR, int nrep, const float* __restrict__ dcoef,
                                    const long long* __restrict__ dense_idx, float* __restrict__ dbg, float* __restrict__ dimp, int K,
                                    int V, int Cout, int E) {
  extern __shared__ float sm[];   // Tbl [V][Cout], then t[w] = sum_co bg[k][co] Tbl[w][co]  ([V]), then dA [V][V]
  float* Tbl = sm;
  float* tw = sm + V * Cout;
  float* dA = tw + V;
  const int k = blockIdx.x;
  for (int i = threadIdx.x; i < V * Cout; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < nrep; ++r) s += TblR[static_cast<size_t>(r) * V * Cout + i];
    Tbl[i] = s;
  }
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) dA[i] = 0.f;
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const long long d = dense_idx[e];
    if (d / (V * V) == k) dA[d - static_cast<long long>(k) * V * V] = dcoef[e];
  }
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < V; ++w) s = fmaf(colsum[k * V + w], Tbl[w * Cout + co], s);
    dbg[k * Cout + co] = s;
  }
  for (int w = threadIdx.x; w < V; w += blockDim.x) {
    float s = 0.f;
    for (int co = 0; co < Cout; ++co) s = fmaf(bg[k * Cout + co], Tbl[w * Cout + co], s);
    tw[w] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) {
    const int w = i % V;
    const int d = k * V * V + i;
    dimp[d] = A[d] * (dA[i] + tw[w]);
  }
}

}  // namespace fmm

using namespace fmm;

extern "C" {

int fmm_gcn_prep_fwd(const float* A, const float* imp, const float* bg, const long long* dense_idx, const long long* bwd_perm,
                     float* coef_f, float* coef_b, float* colsum, float* bias_eff, int K, int V, int Cout, int E,
                     cudaStream_t stream) {
  FMM_CHECK_ARG(A && imp && bg && dense_idx && bwd_perm && coef_f && coef_b && colsum && bias_eff && K > 0 && V > 0 && Cout > 0 && E > 0,
                "gcn_prep_fwd: bad arguments");
  gcn_prep_fwd_kernel<<<1, 512, K * V * sizeof(float), stream>>>(A, imp, bg, dense_idx, bwd_perm, coef_f, coef_b, colsum, bias_eff, K, V,
                                                                  Cout, E);
  FMM_CHECK_LAUNCH("gcn_prep_fwd");
  return FMM_OK;
}

int fmm_gcn_prep_bwd(const float* A, const float* bg, const float* colsum, const float* TblR, int nrep, const float* dcoef,
                     const long long* dense_idx, float* dbg, float* dimp, int K, int V, int Cout, int E, cudaStream_t stream) {
  FMM_CHECK_ARG(A && bg && colsum && TblR && dcoef && dense_idx && dbg && dimp && nrep > 0 && K > 0 && V > 0 && Cout > 0 && E > 0,
                "gcn_prep_bwd: bad arguments");
  const size_t smem = (static_cast<size_t>(V) * Cout + V + static_cast<size_t>(V) * V) * sizeof(float);
  FMM_CHECK_ARG(smem <= 48 * 1024, "gcn_prep_bwd: V=%d Cout=%d needs %zu bytes of shared memory", V, Cout, smem);
  gcn_prep_bwd_kernel<<<K, 512, smem, stream>>>(A, bg, colsum, TblR, nrep, dcoef, dense_idx, dbg, dimp, K, V, Cout, E);
  FMM_CHECK_LAUNCH("gcn_prep_bwd");
  return FMM_OK;
}

}  // extern "C"
