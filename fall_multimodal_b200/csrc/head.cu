// Fused late-fusion head + cross-entropy (BASELINE north star piece 4; reference combination.py:37-46 for
// cat -> Linear, F2/main.py:111-113 + :280 for CrossEntropyLoss(label_smoothing) on probability targets, and the notebooks'
// softmax-before-the-loss variant, GSTCAN_HAR_conv_10kfold.ipynb#cell1:L416 / #cell7:L129).
//
//   z[n][c]  = bias[c] + sum_s sum_f feat_s[n][f] * W[c][off_s + f]            (up to 4 feature segments, no concat copy)
//   o        = pre_softmax ? softmax(z) : z                                    (what the model returns)
//   t'       = t * (1 - eps) + eps / C                                         (label smoothing on probability targets)
//   loss     = -(1/N) sum_n sum_c t'[n][c] * log_softmax(o[n])[c]
//
// forward: one block per row n (logits by warp-shuffle dot products, softmax / loss by the first warp);
// backward: dz from the saved probabilities, then d feat_s = dz W_s (one block per row) and dW = dz^T feat, db = sum_n dz
// (one block per 128 input features). Everything fp32; N*C*F is ~2 MFLOP: the point is two launches instead of ~25.
#include "common.cuh"

namespace fmm {

constexpr int kHeadMaxC = 32;
constexpr int kHeadMaxSeg = 4;

struct HeadArgs {
  const float* feat[kHeadMaxSeg];
  float* dfeat[kHeadMaxSeg];
  int width[kHeadMaxSeg];
  int nseg;
  const float* W;      // [C][F]
  const float* bias;   // [C]
  const float* target; // [N][C] probabilities
  float* out;          // [N][C] logits (or probabilities when pre_softmax)
  float* prob;         // [N][C] softmax(z)                      (saved for backward)
  float* prob2;        // [N][C] softmax(softmax(z)) when pre_softmax (saved for backward)
  float* loss;         // scalar, accumulated into (zero it first)
  const float* gloss;  // scalar: d L / d loss (backward)
  float* dz;           // [N][C] workspace (backward)
  float* dW;           // [C][F]
  float* dbias;        // [C]
  int N, C, F;
  int pre_softmax;
  float smoothing;
};

__global__ void head_ce_fwd_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ float z[kHeadMaxC];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < a.C; c += nwarps) {
    const float* w = a.W + static_cast<size_t>(c) * a.F;
    float acc = 0.f;
    int off = 0;
    for (int s = 0; s < a.nseg; ++s) {
      const float* f = a.feat[s] + static_cast<size_t>(n) * a.width[s];
      for (int i = lane; i < a.width[s]; i += 32) acc = fmaf(f[i], w[off + i], acc);
      off += a.width[s];
    }
    acc = warp_sum(acc);
    if (lane == 0) z[c] = acc + (a.bias ? a.bias[c] : 0.f);
  }
  __syncthreads();
  if (warp == 0) {
    const bool ok = lane < a.C;
    const float zc = ok ? z[lane] : -INFINITY;
    float m = zc;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = ok ? expf(zc - m) : 0.f;
    const float se = warp_sum(e);
    const float p = e / se;
    float logq = zc - m - logf(se);   // log_softmax(z)
    float outv = zc;
    if (a.pre_softmax) {
      // the notebooks return softmax(z) and feed it to CrossEntropyLoss: softmax applied twice
      float m2 = ok ? p : -INFINITY;
      for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      const float e2 = ok ? expf(p - m2) : 0.f;
      const float se2 = warp_sum(e2);
      logq = p - m2 - logf(se2);
      outv = p;
      if (ok && a.prob2) a.prob2[static_cast<size_t>(n) * a.C + lane] = e2 / se2;
    }
    const float t = ok ? a.target[static_cast<size_t>(n) * a.C + lane] * (1.f - a.smoothing) + a.smoothing / a.C : 0.f;
    const float l = warp_sum(ok ? -t * logq : 0.f);
    if (ok) {
      a.out[static_cast<size_t>(n) * a.C + lane] = outv;
      a.prob[static_cast<size_t>(n) * a.C + lane] = p;
    }
    if (lane == 0) atomicAdd(a.loss, l / a.N);
  }
}

// one block per row: dz, then d feat_s[n][f] = sum_c dz[c] W[c][off_s + f]
__global__ void head_ce_bwd_rows_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ float dzs[kHeadMaxC];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const bool ok = lane < a.C;
    const size_t i = static_cast<size_t>(n) * a.C + lane;
    const float g = a.gloss[0] / a.N;
    const float t = ok ? a.target[i] * (1.f - a.smoothing) + a.smoothing / a.C : 0.f;
    const float st = warp_sum(t);
    float d;
    if (!a.pre_softmax) {
      d = ok ? g * (a.prob[i] * st - t) : 0.f;
    } else {
      const float p = ok ? a.prob[i] : 0.f;
      const float dp = ok ? g * (a.prob2[i] * st - t) : 0.f;       // d loss / d p (p = softmax(z) is the loss input)
      const float dot = warp_sum(dp * p);
      d = p * (dp - dot);                                          // back through the first softmax
    }
    if (ok) {
      dzs[lane] = d;
      a.dz[i] = d;
    }
  }
  __syncthreads();
  int off = 0;
  for (int s = 0; s < a.nseg; ++s) {
    if (a.dfeat[s]) {
      float* df = a.dfeat[s] + static_cast<size_t>(n) * a.width[s];
      for (int f = threadIdx.x; f < a.width[s]; f += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < a.C; ++c) acc = fmaf(dzs[c], a.W[static_cast<size_t>(c) * a.F + off + f], acc);
        df[f] = acc;
      }
    }
    off += a.width[s];
  }
}

// grid = (F / 128, batch slices): dW[c][f] += sum_{n in slice} dz[n][c] feat[n][f] (dW zeroed by the launcher: one block
// looping over the whole batch took 180 us at N = 256, a latency-bound serial chain); slice 0 / block 0 also adds dbias
constexpr int kHeadSlices = 16;
__global__ void head_ce_bwd_w_kernel(const __grid_constant__ HeadArgs a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = static_cast<int>(static_cast<long long>(a.N) * blockIdx.y / gridDim.y);
  const int n1 = static_cast<int>(static_cast<long long>(a.N) * (blockIdx.y + 1) / gridDim.y);
  if (blockIdx.x == 0 && threadIdx.x < a.C && a.dbias) {
    float s = 0.f;
    for (int n = n0; n < n1; ++n) s += a.dz[static_cast<size_t>(n) * a.C + threadIdx.x];
    atomicAdd(a.dbias + threadIdx.x, s);
  }
  if (f >= a.F) return;
  int s = 0, off = 0;
  while (s < a.nseg - 1 && f >= off + a.width[s]) off += a.width[s++];
  const float* col = a.feat[s] + (f - off);
  const int wd = a.width[s];
  float acc[kHeadMaxC];
#pragma unroll
  for (int c = 0; c < kHeadMaxC; ++c) acc[c] = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float x = col[static_cast<size_t>(n) * wd];
    const float* d = a.dz + static_cast<size_t>(n) * a.C;
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c)
      if (c < a.C) acc[c] = fmaf(d[c], x, acc[c]);
  }
#pragma unroll
  for (int c = 0; c < kHeadMaxC; ++c)
    if (c < a.C) atomicAdd(a.dW + static_cast<size_t>(c) * a.F + f, acc[c]);
}

}  // namespace fmm

using namespace fmm;

extern "C" {

struct fmm_head_args {   // mirror of fmm::HeadArgs (include/fmm_b200.h)
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
};
static_assert(sizeof(fmm_head_args) == sizeof(HeadArgs), "fmm_head_args must mirror fmm::HeadArgs");

static int head_check(const fmm_head_args* a, const char* what) {
  FMM_CHECK_ARG(a && a->nseg >= 1 && a->nseg <= kHeadMaxSeg && a->N > 0 && a->C > 0 && a->C <= kHeadMaxC && a->F > 0,
                "%s: bad shape (at most %d classes, %d feature segments)", what, kHeadMaxC, kHeadMaxSeg);
  int f = 0;
  for (int s = 0; s < a->nseg; ++s) {
    FMM_CHECK_ARG(a->feat[s] && a->width[s] > 0, "%s: segment %d is empty", what, s);
    f += a->width[s];
  }
  FMM_CHECK_ARG(f == a->F && a->W && a->target && a->prob, "%s: segment widths do not add up to F=%d", what, a->F);
  FMM_CHECK_ARG(!a->pre_softmax || a->prob2, "%s: pre_softmax needs prob2", what);
  return FMM_OK;
}

int fmm_head_ce_fwd(const fmm_head_args* a, cudaStream_t stream) {
  int st = head_check(a, "head_ce_fwd");
  if (st != FMM_OK) return st;
  FMM_CHECK_ARG(a->out && a->loss, "head_ce_fwd: null output");
  head_ce_fwd_kernel<<<a->N, 128, 0, stream>>>(*reinterpret_cast<const HeadArgs*>(a));
  FMM_CHECK_LAUNCH("head_ce_fwd");
  return FMM_OK;
}

int fmm_head_ce_bwd(const fmm_head_args* a, cudaStream_t stream) {
  int st = head_check(a, "head_ce_bwd");
  if (st != FMM_OK) return st;
  FMM_CHECK_ARG(a->gloss && a->dz && a->dW, "head_ce_bwd: null pointer");
  const HeadArgs& h = *reinterpret_cast<const HeadArgs*>(a);
  head_ce_bwd_rows_kernel<<<a->N, 256, 0, stream>>>(h);
  cudaMemsetAsync(a->dW, 0, sizeof(float) * a->C * a->F, stream);
  if (a->dbias) cudaMemsetAsync(a->dbias, 0, sizeof(float) * a->C, stream);
  const int slices = a->N < kHeadSlices ? a->N : kHeadSlices;
  head_ce_bwd_w_kernel<<<dim3((a->F + 127) / 128, slices), 128, 0, stream>>>(h);
  FMM_CHECK_LAUNCH("head_ce_bwd");
  return FMM_OK;
}

}  // extern "C"
