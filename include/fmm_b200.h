/* libfmm_b200 — C ABI of the B200-native (sm_100a) kernels behind the GSTCAN fall/HAR train step.
 *
 * The reference (musaru/Fall_Multimodal) is pure PyTorch: it has no operator/FFI layer, its hot path
 * is whatever ATen dispatches for the call sites below. This header is therefore the boundary a
 * maintainer binds INSTEAD of those torch calls (ctypes stub: fall_multimodal_b200/_lib.py; see
 * INTEGRATION.md). Reference file:line citations are relative to
 * /root/reference/Fall_2_Spatial_Temporal_SR/Model/ unless stated otherwise.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (allocated by PyTorch's caching
 *     allocator); the library never allocates, frees or keeps device memory;
 *   - activations are channels-last [N][T][V][C] contiguous, dtype FMM_DT_BF16 or FMM_DT_F32
 *     (fp32 activations run the tensor cores in a bf16x3 split with fp32-grade products);
 *     statistics, coefficients and parameter gradients are fp32 / fp64 as noted;
 *   - all launches go to the given stream and return immediately; return 0 on success, a negative
 *     status otherwise (fmm_last_error() has the text). There is NO CPU fallback;
 *   - `err` (may be NULL) is a device word the GEMM kernels write before trapping if an mbarrier
 *     wait exceeds ~2 s (protocol watchdog: a bug becomes a CUDA error, not a hung GPU).
 */
#ifndef FMM_B200_H
#define FMM_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMM_DT_BF16 0
#define FMM_DT_F32 1

const char* fmm_last_error(void);
int fmm_version(void);
int fmm_device_supported(void); /* 1 iff the current device is sm_100 */
/* grid budget of the persistent kernels: 0 = all SMs (default), n = at most n (even) CTAs per launch, e.g. SMs / 2 while the two
 * trunks of a fusion model run on concurrent streams; process-wide, read at launch time; returns the previous value */
int fmm_set_sm_limit(int n);
/* (the two measurement aids fmm_debug_* are declared in fmm_b200_debug.h: not part of the product interface) */

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM engines
 * ------------------------------------------------------------------------------------------- */

/* out[n, j*ostride+ooff, v, co] = bias + sum_{m<ntaps} sum_ci f(x[n, j*istride+shifts[m], v, ci]) * W[m][co][ci]
 * f(x) = relu?(x*in_scale[ci]+in_shift[ci]); zero outside [0,Tin) (padding after the prologue).
 * Replaces: nn.Conv2d (9,1) stride (s,1) pad (4,0) with the preceding BatchNorm2d+ReLU (stgcan.py:112-118),
 * the 1x1 conv of GraphConvolution on the aggregated input (stgcan.py:51-54, reassociated), the strided
 * residual 1x1 conv (stgcan.py:128-131), and the dgrads of all three (autograd of the same lines).
 * `wpk` comes from fmm_tapconv_pack; bias_per_joint: bias is [V][Cout] (the graph-conv bias folded through
 * the adjacency column sums). */
int fmm_tapconv_bn(int cout, int dtype);
long long fmm_tapconv_packed_bytes(int cin, int cout, int ntaps, int dtype);
int fmm_tapconv_pack(const float* w, void* out, int cout, int cin, int n2, int k2, long long sn1, long long sn2,
                     long long sk1, long long sk2, long long sm, int ntaps, const int* tapmap, int dtype,
                     cudaStream_t stream);
int fmm_tapconv(const void* x, void* out, const void* wpk, const float* in_scale, const float* in_shift, int in_relu,
                const float* bias, int bias_per_joint, int N, int V, int Tin, int Tout, int Cin, int Cout, int Tj,
                int istride, int ostride, int ooff, int ntaps, const int* shifts, int dtype, unsigned* err,
                cudaStream_t stream);

/* dw[m*s_m + (ci/c2)*s_c1 + (ci%c2)*s_c2 + co*s_co] += sum_{n,v,j} f(x[n, j*istride+shifts[m], v, ci]) * dy[n,j,v,co]
 * (fp32 atomics: zero dw first). Replaces the weight gradients autograd computes for the three convs above. */
int fmm_wgrad(const void* x, const void* dy, float* dw, const float* in_scale, const float* in_shift, int in_relu,
              int N, int V, int Tin, int Tj, int Cin, int Cout, int istride, int ntaps, const int* shifts, int c2,
              long long s_m, long long s_c1, long long s_c2, long long s_co, int dtype, unsigned* err,
              cudaStream_t stream);

/* Fused spatial graph convolution (stgcan.py:50-56 + the A*edge_importance product of :222), bf16 activations:
 *   g[r][co] = bias[v(r)][co] + sum_k sum_ci (sum_{e in in(k,v(r))} coef[e] * x[frame(r)*V + src[e]][ci]) * W[k*Cout+co][ci]
 * over the flat rows r = (n,t,v) of channels-last x[rows][Cin] -> g[rows][Cout]. The adjacency aggregation runs in the GEMM
 * prologue (shared memory), the K-times wider intermediate never reaches HBM; x is fetched and g written with tensor-map
 * TMA. ch_sum / ch_sq (nullable, fp64 [nrep][Cout], accumulated into) receive the per-channel sum / sum of squares of the
 * stored g: the BatchNorm2d batch statistics of stgcan.py:112. `xa` (nullable) additionally receives the aggregated operand
 * [rows][K*Cin] (dev / cross-check output). CSR arrays (device) as fmm_agg_fwd; kdeg = HOST array of K ints, the maximum
 * in-degree of each partition of the (static) graph. Cin, Cout multiples of 64, Cout <= 256, V <= 33.
 * `wpk` from fmm_gcn_pack(W fp32 [K*Cout][Cin]). */
long long fmm_gcn_packed_bytes(int K, int Cin, int Cout);
int fmm_gcn_pack(const float* w, void* out, int K, int Cin, int Cout, cudaStream_t stream);
int fmm_gcn_fwd(const void* x, void* g, void* xa, const void* wpk, const float* bias, const int* rowptr, const int* src,
                const float* coef, const int* kdeg, double* ch_sum, double* ch_sq, int nrep, long long rows, int V, int K,
                int Cin, int Cout, int E, unsigned* err, cudaStream_t stream);

/* Weight gradient of the fused graph convolution: dw[(k*Cout + co)*Cin + ci] += sum_r dg[r][co] * A_k[r][ci], where A_k is the
 * aggregated input of fmm_gcn_fwd, re-derived in the prologue (nothing saved by the forward pass). fp32 atomics: zero dw
 * first. Replaces autograd's weight gradient of the 1x1 conv of stgcan.py:42-51 through the einsum of :54. */
int fmm_gcn_wgrad(const void* x, const void* dg, float* dw, const int* rowptr, const int* src, const float* coef,
                  const int* kdeg, long long rows, int V, int K, int Cin, int Cout, int E, unsigned* err, cudaStream_t stream);

/* Backward data of the fused graph convolution, bf16:
 *   dx[(f,v)][ci] = addend[(f,v)][ci] + sum_{e in out(v)} coef[e] * P[(f, dst e)][kk e][ci],   P = dg . W_k^T  (never in HBM)
 *   dcoef[eid[e]] += sum_{f,ci} x[(f,v)][ci] * P[(f, dst e)][kk e][ci]     (gradient of A*edge_importance, stgcan.py:222)
 * GEMM on tcgen05 (N = K*64 per 64-channel slab), transposed adjacency aggregation in the epilogue. CSR over v (out-edges):
 * rowptr[V+1], dst / kk / coef / eid [E] (device); max_out_degree = largest number of out-edges of a joint (host knowledge
 * of the static graph, <= 8). addend, x/eid/dcoef nullable. `wpk` from fmm_gcn_pack_bwd. K <= 3, rows % V == 0.
 * relu_mask = 1 stores dx * (x > 0): x is the ReLU output of the previous block (stgcan.py:135), whose backward then reads an
 * already masked gradient (fmm_blockout_bwd_reduce / fmm_bn2_bwd_apply with Y = NULL) and never re-reads its output. */
long long fmm_gcn_packed_bwd_bytes(int K, int Cin, int Cout);
int fmm_gcn_pack_bwd(const float* w, void* out, int K, int Cin, int Cout, cudaStream_t stream);
int fmm_gcn_bwd(const void* dg, const void* x, const void* addend, void* dx, const void* wpk, const int* rowptr, const int* dst,
                const int* kk, const float* coef, const int* eid, float* dcoef, int relu_mask, int max_out_degree, long long rows,
                int V, int K, int Cin, int Cout, unsigned* err, cudaStream_t stream);

/* data_bn (stgcan.py:212-218): BatchNorm1d(V*C) over (N,T) of the clip x (N,C,T,V) fp32, channel index v*C + c, written
 * once in the channels-last (N,T,V,C) activation layout. stats -> fmm_bn_finalize(count = N*T, C = V*C) -> apply; bwd gives
 * the affine parameter gradients (fp64 accumulators, zeroed by the caller; the clip itself needs no gradient). */
int fmm_databn_stats(const float* x, double* sum, double* sq, int N, int C, int T, int V, cudaStream_t stream);
int fmm_databn_apply(const float* x, const float* a, const float* b, void* y, int N, int C, int T, int V, int dtype,
                     cudaStream_t stream);
int fmm_databn_bwd(const void* dy, const float* x, const float* mean, const float* rstd, double* dgamma, double* dbeta, int N,
                   int C, int T, int V, int dtype, cudaStream_t stream);

/* Fused late-fusion head + cross-entropy (combination.py:37-46: cat -> Linear; F2/main.py:113,280: CrossEntropyLoss with
 * label smoothing on PROBABILITY targets, mean over the batch; pre_softmax: the notebooks' softmax-before-the-loss).
 * Up to 4 feature segments (no concat copy), at most 32 classes, everything fp32. fwd: out = logits (or softmax(logits)),
 * prob / prob2 saved, loss accumulated into a zeroed scalar. bwd: dfeat[s] (nullable per segment), dW [C][F], dbias [C],
 * dz = [N][C] workspace, gloss = device scalar d L / d loss. */
typedef struct fmm_head_args {
  const float* feat[4];
  float* dfeat[4];
  int width[4];
  int nseg;
  const float* W;
  const float* bias;
  const float* target;
  float* out;
  float* prob;
  float* prob2;
  float* loss;
  const float* gloss;
  float* dz;
  float* dW;
  float* dbias;
  int N, C, F;
  int pre_softmax;
  float smoothing;
} fmm_head_args;
int fmm_head_ce_fwd(const fmm_head_args* a, cudaStream_t stream);
int fmm_head_ce_bwd(const fmm_head_args* a, cudaStream_t stream);

/* Multi-tensor RMSprop (F2/optimizer.py:20-21; MF3/main.py:103-113 for unscale + clip_grad_norm_): ONE launch for all
 * parameter tensors. `tensors`: device array of {float* param; const float* grad; float* square_avg; long long numel};
 * chunk tables: fmm_opt_chunk() elements per chunk. lr / norm_sq / inv_scale are DEVICE scalars (graph-replay safe);
 * norm_sq (nullable) = sum of squared gradients from fmm_grad_norm_sq (zero it first): enables clipping to max_norm (> 0)
 * and skips the step when it is not finite. */
int fmm_opt_chunk(void);
int fmm_rmsprop_step(const void* tensors, const int* chunk_tensor, const long long* chunk_off, int nchunks, const float* lr,
                     float alpha, float eps, float weight_decay, const float* norm_sq, float max_norm, const float* inv_scale,
                     cudaStream_t stream);
int fmm_grad_norm_sq(const void* tensors, const int* chunk_tensor, const long long* chunk_off, int nchunks, float* norm_sq,
                     cudaStream_t stream);

/* Parameter-side algebra of the graph conv (stgcan.py:222 `A * importance`, bias folded through the aggregation) and its
 * backward: edge coefficients in forward / out-edge CSR order, column sums, per-joint bias table; gradients of the conv bias
 * and the edge importance from the kernels' per-edge / per-joint sums. One launch each. dense_idx / bwd_perm: int64 [E]. */
int fmm_gcn_prep_fwd(const float* A, const float* imp, const float* bg, const long long* dense_idx, const long long* bwd_perm,
                     float* coef_f, float* coef_b, float* colsum, float* bias_eff, int K, int V, int Cout, int E,
                     cudaStream_t stream);
int fmm_gcn_prep_bwd(const float* A, const float* bg, const float* colsum, const float* TblR, int nrep, const float* dcoef,
                     const long long* dense_idx, float* dbg, float* dimp, int K, int V, int Cout, int E, cudaStream_t stream);

/* Inference path of the bidirectional LSTM (bilstm.py:29,48-55) on tensor cores: feat [N][ndir*64] = mean over T (mean_feature
 * != 0) or the t = T-1 output, from zero state. 64 windows per CTA for the whole sequence, recurrent + input weights resident
 * in registers as mma.sync fragments, 3xTF32 split (fp32-grade), state double buffered in shared memory. Weight layout as
 * fmm_lstm_fwd. H = 64, I <= 39. */
int fmm_lstm_infer(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* feat, int N,
                   int T, int I, int H, int ndir, int mean_feature, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * memory-bound kernels (one pass over an activation each)
 * ------------------------------------------------------------------------------------------- */

/* xa[(n,t,w), k*Cin+ci] = sum_{e in in(k,w)} coef[e] * x[(n,t,src[e]), ci]   — einsum 'nkctv,kvw->nctw'
 * (stgcan.py:54) with A*importance (stgcan.py:222) in CSR form, applied to the INPUT channels. */
int fmm_agg_fwd(const void* x, void* xa, const int* rowptr, const int* src, const float* coef, int N, int T, int V,
                int Cin, int K, int dtype, cudaStream_t stream);
/* dx = addend + A^T-aggregate(P); with x != NULL also dcoef[eid[e]] += <x[src e], P[dst e, k e]> (edge-importance grad) */
int fmm_agg_bwd(const void* P, const void* addend, void* dx, const int* rowptr, const int* dst, const int* kk,
                const float* coef, const void* x, const int* eid, float* dcoef, int N, int T, int V, int Cin, int K,
                int dtype, cudaStream_t stream);
int fmm_agg_dcoef(const void* x, const void* P, float* dcoef, const int* src, const int* dst, const int* kk, int E,
                  int N, int T, int V, int Cin, int K, int dtype, cudaStream_t stream);

/* Cross-block accumulators (ch_sum, ch_sq, T1, T2, sum_dU, sum_dR: [nrep][C] doubles; Tbl: [nrep][V][C] floats) are
 * replicated nrep times to spread the atomics; consumers sum the replicas. */
/* per-channel sum / sum of squares (fp64 accumulators) and per-(n,c) sums: BatchNorm2d batch statistics
 * (stgcan.py:112,119,132) and the AdaptiveAvgPool2d of Channel_Attention (stgcan.py:64) in one pass. */
int fmm_colstats(const void* X, double* ch_sum, double* ch_sq, float* nc_sum, int nrep, int N, int T, int V, int C,
                 int dtype, cudaStream_t stream);
/* Y = relu(k1[n,c]*U + k0[n,c] + res): BN2 + SE scale + residual + ReLU of st_gcan.forward (stgcan.py:138-144) */
int fmm_block_out(const void* U, const float* k1, const float* k0, const void* res, const float* ar, const float* br,
                  void* Y, int N, int T, int V, int C, int dtype, cudaStream_t stream);
/* H = relu(a[c]*X + b[c]): BatchNorm2d + ReLU of tcn[0..1] (stgcan.py:112-113), materialised once per block */
int fmm_affine_relu(const void* X, const float* a, const float* b, void* H, int N, int T, int V, int C, int dtype,
                    cudaStream_t stream);
int fmm_blockout_bwd_reduce(const void* dY, const void* Y, const void* U, const void* R, float* S1, float* S2,
                            float* S3, int N, int T, int V, int C, int dtype, cudaStream_t stream);
int fmm_bn2_bwd_apply(const void* dY, const void* Y, const void* U, const void* R, const float* k1, const float* k2,
                      const float* k3, const float* r1, const float* r2, const float* r3, void* dU, void* dR,
                      void* dPre, double* sum_dU, double* sum_dR, int nrep, int N, int T, int V, int C, int dtype,
                      cudaStream_t stream);
int fmm_bn1_bwd_reduce(const void* dH, const void* G, const float* a1, const float* b1, double* T1, double* T2,
                       int nrep, int N, int T, int V, int C, int dtype, cudaStream_t stream);
int fmm_bn1_bwd_apply(const void* dH, const void* G, const float* a1, const float* b1, const float* c1,
                      const float* c2, const float* c3, void* dG, float* Tbl, int nrep, int N, int T, int V, int C,
                      int dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * per-channel / per-clip kernels
 * ------------------------------------------------------------------------------------------- */
/* nn.BatchNorm*: statistics -> scale/shift (+ running stats, momentum/eps as torch) */
int fmm_bn_finalize(const double* ch_sum, const double* ch_sq, int nrep, double count, const float* gamma, const float* beta,
                    float* rmean, float* rvar, float momentum, float eps, int training, float* a, float* b,
                    float* mean_out, float* rstd_out, int C, cudaStream_t stream);
/* Channel_Attention MLP (stgcan.py:63-70): pooled sums -> s[n,c] and the fused output coefficients k1, k0 */
int fmm_se_fwd(const float* pool, const float* a2, const float* b2, float invM, const float* W1, const float* b1,
               const float* gamma, const float* beta, float* rmean, float* rvar, float momentum, float eps,
               int training, const float* W2, const float* b2se, float* p, float* h, float* ah, float* bh,
               float* hmean, float* hrstd, float* s, float* k1, float* k0, int N, int C, int C4, cudaStream_t stream);
int fmm_se_bwd(const float* S1, const float* S2, const float* a2, const float* b2, const float* s, const float* p,
               const float* h, const float* ah, const float* bh, const float* hmean, const float* hrstd,
               const float* W1, const float* W2, int training, float* dq, float* dhr, float* r, float* dh, float* dp,
               float* dW1, float* db1, float* dgamma, float* dbeta, float* dW2, float* db2se, int N, int C, int C4,
               cudaStream_t stream);
/* The two weight-gradient products of fmm_se_bwd on their own (call fmm_se_bwd with dW1 = db1 = dW2 = db2se = NULL first):
 * dW2[c][j] += sum_n dq[n,c] r[n,j], db2se[c] += sum_n dq[n,c]; dW1[j][c] += sum_n dh[n,j] p[n,c], db1[j] += sum_n dh[n,j]
 * (channel_attention_module.atten.4 / .1 of stgcan.py:63-70). Nothing on the activation-gradient chain depends on them. */
int fmm_se_bwd_params(const float* dq, const float* r, const float* dh, const float* p, float* dW1, float* db1, float* dW2,
                      float* db2se, int N, int C, int C4, cudaStream_t stream);
int fmm_bn2_bwd_coef(const float* S1, const float* S2, const float* S3, const float* pool, const float* dp,
                     const float* s, const float* a2, const float* mean2, const float* rstd2, const float* ar,
                     const float* meanr, const float* rstdr, float M, double count, int training, float* k1, float* k2,
                     float* k3, float* r1, float* r2, float* r3, float* dgamma2, float* dbeta2, float* dgammar,
                     float* dbetar, int N, int C, cudaStream_t stream);
int fmm_bn1_bwd_coef(const double* T1, const double* T2, int nrep, const float* a1, const float* mean1, const float* rstd1,
                     double count, int training, float* c1, float* c2, float* c3, float* dgamma, float* dbeta, int C,
                     cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * accelerometer branch: notebook CNN1D (GSTCAN_HAR_conv_10kfold.ipynb#cell2:L6-27), fp32 [N][L][C]
 * ------------------------------------------------------------------------------------------- */
int fmm_conv1d_k5_fwd(const float* x, const float* w, const float* b, float* y, int N, int L, int Ci, int Co,
                      cudaStream_t stream);
int fmm_bn_relu_pool2_fwd(const float* y, const float* a, const float* b, float* out, int N, int L, int C,
                          cudaStream_t stream);
int fmm_pool2_bwd(const float* y, const float* a, const float* b, const float* dout, float* dh, int N, int L, int C,
                  cudaStream_t stream);
int fmm_conv1d_k5_bwd(const float* x, const float* dy, const float* w, float* dx, float* dw, float* db, int N, int L,
                      int Ci, int Co, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * accelerometer branch: nn.LSTM(I, H=64, 1 layer, batch_first, bidirectional) of BiLSTM (bilstm.py:29,48),
 * persistent-CTA recurrence, fp32. Weights stacked per direction ([ndir][4H][I], [ndir][4H][H], [ndir][4H]);
 * out [N][T][ndir*H]; gates [ndir][N][T][4H] / cseq [ndir][N][T][H] are saved for BPTT when non-NULL.
 * lstm_bwd ACCUMULATES dw_ih, dw_hh, db (zero first); dx (optional, zero first) is the input gradient.
 * ------------------------------------------------------------------------------------------- */
int fmm_lstm_fwd(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                 float* out, float* gates, float* cseq, int N, int T, int I, int H, int ndir, cudaStream_t stream);
int fmm_lstm_bwd(const float* x, const float* w_ih, const float* w_hh, const float* out, const float* gates,
                 const float* cseq, const float* dout, float* dw_ih, float* dw_hh, float* db, float* dx, int N,
                 int T, int I, int H, int ndir, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * TRAGCN family (EmbGCN.py:59-89, GRU.py:17-26, TRAGCN.py:150-224, TA.py:40-108).
 *
 * fmm_bgemm: strided batched GEMM  C[g1,g2][m][n] = act(alpha * sum_k A[g1,g2][m][k] B[g1,g2][k][n]
 * + bias_m[m] + bias_n[n]) (+ C when beta).  All strides are in ELEMENTS, a stride of 0 broadcasts,
 * the contraction index is k = (k1*K2 + k2)*K3 + k3 with one stride per level. It replaces torch.einsum /
 * torch.matmul / nn.Linear / nn.Conv2d(T,T,(1,3)) at every call site of those reference files.
 * dtype = operand type (FMM_DT_*), c_dtype = type of C; splitk > 1 accumulates partial sums into an
 * fp32 C with atomics (zero it first; no bias/act/beta).
 * ------------------------------------------------------------------------------------------- */
typedef struct fmm_bgemm_desc {
  const void* A;
  const void* B;
  void* C;
  const float* bias_m;
  const float* bias_n;
  long long a_g1, a_g2, a_m, a_k1, a_k2, a_k3;
  long long b_g1, b_g2, b_n, b_k1, b_k2, b_k3;
  long long c_g1, c_g2, c_m, c_n;
  int G1, G2, M, N, K1, K2, K3;
  float alpha;
  int beta;
  int act; /* 0 none, 1 relu */
  int splitk;
  int dtype;
  int c_dtype;
} fmm_bgemm_desc;
int fmm_bgemm(const fmm_bgemm_desc* desc, cudaStream_t stream);

/* Graph-GRU cell glue (GRU.py:17-26 around EmbGCN.py:73-88; csrc/gru_cell.cu): one launch between consecutive
 * per-node GEMMs, one block per clip, activations are [B][V][*] slices with a clip and a joint stride. The cell input
 * is laid out [h (H) | x (Din) | 1 | 0-pad] (Cp wide, per-node weight rows permuted to match).
 *  cell_fwd mode 0: cat=[h_{t-1}|x_t] -> xc0 = S.cat, xc1 = cat
 *           mode 1: zr = sigmoid(pre + silu(lin)) -> zr, lg ; cat=[r*h_{t-1}|x_t] -> xc0, xc1
 *           mode 2: hc = tanh(pre + silu(lin)), h_t = z*h_{t-1} + (1-z)*hc -> hc, lu, hout ; cat=[h_t|x] -> xc0, xc1 (if xc0)
 *  cell_bwd mode 0: update/candidate backward of one step from carry + dH -> dz, carry, dpre_u, dlin_u
 *           mode 1: candidate-stage mix backward -> dx, carry, dpre_g, dlin_g
 *           mode 2: gate-stage mix backward -> dx (+=), carry, then mode 0 of the previous step when do_bwd1 */
typedef struct fmm_cell_fwd_args {
  const void* x; long long xb, xv;
  const void* hprev; long long hb, hv;
  const float* S;
  const void* pre; const void* lin;   /* stage GEMM outputs, activation dtype */
  void* zr; void* lg;
  void* hc; void* lu;
  void* hout; long long ob, ov;
  void* xc0; void* xc1;
  int mode, B, V, Din, H, Cp;
} fmm_cell_fwd_args;
typedef struct fmm_cell_bwd_args {
  const float* S;
  float* carry; float* dz;
  const void* dxc0; const void* dxc1;
  void* dx; long long dxb, dxv;
  const void* hprev; long long hb, hv;
  const void* zr; const void* lg;
  void* dpre_g; void* dlin_g;
  const void* dH; long long db, dv;
  const void* z1;
  const void* hprev1; long long hb1, hv1;
  const void* hc1; const void* lu1;
  void* dpre_u; void* dlin_u;
  int mode, dx_accum, do_bwd1, B, V, Din, H, Cp;
} fmm_cell_bwd_args;
int fmm_tg_cell_fwd(const fmm_cell_fwd_args* args, int dtype, cudaStream_t stream);
int fmm_tg_cell_bwd(const fmm_cell_bwd_args* args, int dtype, cudaStream_t stream);

/* Persistent graph-GRU scan (csrc/gruscan.cu; GRU.py:17-27 around EmbGCN.py:69-89, the time loop of
 * TRAGCN.py:158-166 as ONE launch per layer and direction). bf16, 64 hidden channels, V <= 32 joints. Clusters of 8 CTAs
 * own BC clips each (fmm_gruscan_geometry); CTA j of a cluster owns hidden channels 8j..8j+7 of every joint.
 *   blocked state   XC[slot][cluster][8 slices][pm: 0 plain, 1 mixed][V][BC][8]
 *   fragment order  PX[t][cluster][8][item][3][32 lanes][8], FS[...][4][32][8]; item = (warp * NPW + q) * (BC/16) + mt
 * mode 0: input half of both EmbGCN products for every step (parallel over t, `tsplit` clusters per clip group);
 * mode 1: forward scan (xcg slot t+1 / xcu slot t / fs / hout written);
 * mode 2: backward scan (dxu / dxgz / dxgr slot t written: pre-activation gradients of the candidate, z and r products). */
typedef struct fmm_gruscan_args {
  const void* xb;
  void* px;
  void* xcg;
  void* xcu;
  void* fs;
  void* hout;
  const void* W;      /* [V][K][192] bf16, columns (z 64 | r 64 | candidate 64) */
  const void* Lw;     /* [K][192] bf16 */
  const float* cs;    /* [V] */
  const float* S;     /* [V][V] */
  const float* bg;    /* [V][192] */
  const float* bl;    /* [192] */
  const void* dhout; long long dh_b, dh_t, dh_v;
  void* dxu; void* dxgz; void* dxgr;
  const void* WT;     /* [V][192][64] bf16 */
  const void* LT;     /* [192][64] bf16 */
  unsigned* err;
  unsigned long long* prof; /* dev aid: 16 cycle counters (null: off) */
  int B, T, V, KS, xb_slices, xb_slot0, NC, tsplit;
} fmm_gruscan_args;
int fmm_gruscan_geometry(int V, int* BC, int* NPW);
int fmm_gruscan_max_clusters(int V); /* resident clusters of the scan on the current device */
int fmm_gruscan(const fmm_gruscan_args* args, int mode, cudaStream_t stream);
/* blocked state (+ blocked input) -> row-major [2: mixed, plain][T][B][V][Cp] with columns [h 64 | x Din | 1 | 0] */
int fmm_gruscan_export_xc(const void* xc, const void* xb, void* out, int T, int B, int V, int KS, int xb_slices, int xb_slot0,
                          int Din, int Cp, cudaStream_t stream);
/* blocked pre-activation gradients of the backward scan -> dPLu [2: graph, Linear][T][B][V][64], dPLg [2][T][B][V][128] */
int fmm_gruscan_export_dg(const void* dxu, const void* dxgz, const void* dxgr, void* dPLu, void* dPLg, int T, int B, int V,
                          cudaStream_t stream);
/* dX (B,T,V,Din) = S^T (dXg[0] + dXu[0])[.., c0:c0+Din] + (dXg[1] + dXu[1])[.., c0:c0+Din] from the stage input gradients (2,T,B,V,Cp) */
int fmm_gruscan_mix_dx(const void* dXg, const void* dXu, const float* S, void* dX, int T, int B, int V, int Cp, int c0, int Din,
                       cudaStream_t stream);
/* fragment-order gate values -> ZR, LG [T][B][V][128], HC, LU [T][B][V][64] */
int fmm_gruscan_export_fs(const void* fs, void* ZR, void* LG, void* HC, void* LU, int T, int B, int V, cudaStream_t stream);

/* Row-streaming per-node GEMMs for the batched work after the graph-GRU sweep (csrc/pnode.cu; EmbGCN.py:80-86 has one weight
 * matrix per joint): bf16 operands, rows = T*B.
 *   pn_dgrad: OUT[p][row][n][c0 + c] = sum_k IN[p][row][n][k] W[p][n][c0 + c][k], c < ncols (K in {64,128}, Cp <= 144)
 *   pn_wgrad: part[chunk][p][n][c][o] = sum over the chunk's rows of XC[p][row][n][c] DY[p][row][n][o] (Co in {64,128}); the caller
 *             sums over fmm_pn_wgrad_chunks(P, rows, V) chunks */
int fmm_pn_dgrad(const void* in, const void* W, void* out, int P, long long rows, int V, int K, int Cp, int c0, int ncols, const float* bias,
                 int relu, cudaStream_t stream); /* + bias[c0 + c] (fp32, may be null), optional ReLU: V = 1 is a Linear layer (TA.py:33-37) */
int fmm_pn_wgrad_chunks(int P, long long rows, int V);
/* supports gradient: part[i][n][m] (i < fmm_pn_ds_parts(), 32 x 32 fp32 each, summed by the caller) of sum_{row,c} G[row][n][c] X[row][m][c] */
int fmm_pn_ds_parts(void);
int fmm_pn_ds(const void* G, const void* X, float* part, long long rows, int V, int Cp, cudaStream_t stream);
int fmm_pn_wgrad(const void* xc, const void* dy, float* part, int P, long long rows, int V, int Cp, int Co, cudaStream_t stream);

/* Flash-style time-axis attention (csrc/tattn.cu; TA.py:55-62 softmax(q k^T / sqrt(c), -1) v per (clip, joint)): one CTA per
 * head, the (B,V,T,T) scores are never written. q, k feature-major (B,F,V,Tp) as the time-as-channel convolutions leave them,
 * v (B,V,T,64), out / dout (B,T,V,64), lse (B*V,Tp) fp32 (log2 domain). bf16, F <= 64, Tp % 64 == 0, Tp <= 320.
 * mode 0: forward (out, lse); mode 1: backward (dq, dk, dv; every column of dq / dk up to Tp is written). */
typedef struct fmm_tattn_args {
  const void* q; const void* k; const void* v;
  void* out;
  float* lse;
  const void* dout;
  void* dq; void* dk; void* dv;
  int B, V, T, Tp, F;
  float scale;
  int v_btvc; /* 1: v and dv are (B,T,V,64) */
} fmm_tattn_args;
int fmm_tattn(const fmm_tattn_args* args, int mode, cudaStream_t stream);

/* Time-axis attention pieces (TA.py:55-68): in-place row softmax over the first L of Lp entries (+ backward,
 * written over dp), LayerNorm over C of (a + b) with saved mean/rstd (+ backward; dgamma/dbeta accumulate),
 * positional encoding add, ReLU mask. */
int fmm_tg_softmax_fwd(void* x, long long rows, int L, int Lp, int dtype, cudaStream_t stream);
int fmm_tg_softmax_bwd(const void* p, void* dp, long long rows, int L, int Lp, int dtype, cudaStream_t stream);
int fmm_tg_ln_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                  long long rows, int C, float eps, int dtype, cudaStream_t stream);
int fmm_tg_ln_bwd(const void* dy, const void* a, const void* b, const float* gamma, const float* mean, const float* rstd,
                  void* dx, float* dgamma, float* dbeta, long long rows, int C, int dtype, cudaStream_t stream);
int fmm_tg_add_pe(const void* x, const float* pe, void* y, int B, int T, int V, int C, int dtype, cudaStream_t stream);
int fmm_tg_relu_mask(void* dx, const void* y, long long total, int dtype, cudaStream_t stream);
/* batched 2-D transpose in[g1,g2][r][c] -> out[g1,g2][c][r], rows r in [R, Rp) written as zero (the time <-> feature
 * swap that turns TA.py's time-as-channel (1,3) convolutions into tap convolutions for fmm_tapconv / fmm_wgrad) */
int fmm_tg_transpose(const void* in, void* out, int R, int C, int Rp, long long in_g1, long long in_g2, long long in_rs,
                     long long out_g1, long long out_g2, long long out_rs, int G1, int G2, int dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * musa `Model` (Multimodal_Fall3/model/musa_model.py), channels-last [N][T][V][C], C % 8 == 0.
 *  dwconv_*          depthwise (k x 1) temporal conv, groups = C (:165-168, :426, :446): w [C][k] fp32, bias [C]
 *  affine_act        y = act(a[c]*x + b[c] (+ res)), act 0 none / 1 relu / 2 tanh / 3 leaky-relu(0.01): a folded
 *                    BatchNorm2d + residual add + activation (:144-146, :195-199, :430-436)
 *  bn_act_bwd_reduce S1[c] += sum dz, S2[c] += sum dz*xhat with dz = dy*act'(y), xhat = (x-mean)*rstd (fp64, zero first)
 *  bn_act_bwd_apply  dx = a*(dz - S1/cnt - xhat*S2/cnt) (training) | a*dz (eval); dres = dz (optional)
 * ------------------------------------------------------------------------------------------- */
int fmm_dwconv_fwd(const void* x, const float* w, const float* b, void* out, int N, int Tin, int Tout, int V, int C, int k,
                   int stride, int pad, int dtype, cudaStream_t stream);
int fmm_dwconv_bwd_data(const void* dy, const float* w, void* dx, int N, int Tin, int Tout, int V, int C, int k, int stride,
                        int pad, int dtype, cudaStream_t stream);
int fmm_dwconv_bwd_weight(const void* x, const void* dy, float* dw, float* db, int N, int Tin, int Tout, int V, int C, int k,
                          int stride, int pad, int dtype, cudaStream_t stream);
int fmm_affine_act(const void* x, const float* a, const float* b, const void* res, void* y, long long rows, int C, int act,
                   int dtype, cudaStream_t stream);
int fmm_bn_act_bwd_reduce(const void* dy, const void* y, const void* x, const float* mean, const float* rstd, double* S1,
                          double* S2, long long rows, int C, int act, int dtype, cudaStream_t stream);
int fmm_bn_act_bwd_apply(const void* dy, const void* y, const void* x, const float* a, const float* mean, const float* rstd,
                         const double* S1, const double* S2, double inv_count, int training, void* dx, void* dres,
                         long long rows, int C, int act, int dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * On-device input preparation (3_stream/har_create4_sensor.py:36-47,113-132; Multimodal_Fall3/dataset.py:28-41;
 * Fall_2_Spatial_Temporal_SR/dataset.py:27; Model/combination.py:39): a resident recording is normalised once
 * (prep_frames: xys [L][J][3] fp64 (x, y, score) -> frames [L][J+1][3] fp32 with the centre joint, per-frame
 * score scr [L], score-weighted targets lbw [L][C]; main_mask = joints whose score is weighted x1.5) and
 * T-frame windows are gathered at arbitrary start frames into skeleton [N][3][T][V], motion [N][2][T-1][V]
 * (optional), sensor [N][T][S] (optional) and window-mean labels [N][C] (optional).
 * ------------------------------------------------------------------------------------------- */
int fmm_prep_frames(const double* xys, const double* labels, float* frames, float* scr, float* lbw, int L, int J, int C,
                    unsigned main_mask, int center_main, int nan_to_num, cudaStream_t stream);
int fmm_prep_windows(const float* frames, const float* lbw, const float* sensors, const int* starts, float* skel, float* mot,
                     float* sensor_out, float* label_out, int N, int T, int V, int C, int S, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FMM_B200_H */
