// Persistent-CTA LSTM recurrence for the accelerometer branch (reference: nn.LSTM(I, H, 1 layer,
// batch_first, bidirectional) in /root/reference/Fall_2_Spatial_Temporal_SR/Model/bilstm.py:29,48;
// cuDNN semantics: gate order i,f,g,o, biases b_ih + b_hh, zero initial state).
//
// One CTA owns kBT samples of ONE direction for the whole sequence: thread j = gate row j keeps its
// rows of W_hh and W_ih in registers, h_{t-1} of the kBT samples lives in shared memory and is read
// as a broadcast, the cell state stays in registers of the threads that own (sample, unit) pairs.
// Nothing but x_t comes from global memory inside the time loop. fp32 throughout (the recurrence is
// latency-bound: ~80 FMAs per gate row per sample per step).
// Backward (BPTT) is the mirror: dW rows accumulate in registers over all steps and samples of the
// CTA, dh_{t-1} = dgates . W_hh reads W_hh from shared memory.
#include "common.cuh"

namespace fmm {

constexpr int kBT = 16;      // samples per CTA
constexpr int kMaxI = 32;    // input features kept in registers per gate row
constexpr int kMaxH = 64;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// grid (ceil(N/kBT), ndir), block 4H threads
template <int H>
__global__ void __launch_bounds__(4 * H) lstm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_ih,
                                                         const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                                         const float* __restrict__ b_hh, float* __restrict__ out,
                                                         float* __restrict__ gates, float* __restrict__ cseq, int N,
                                                         int T, int I, int ndir) {
  constexpr int G = 4 * H;
  __shared__ float h_s[kBT][H];
  __shared__ float g_s[kBT][G];
  __shared__ float x_s[kBT][kMaxI];
  const int dir = blockIdx.y;
  const int n0 = blockIdx.x * kBT;
  const int j = threadIdx.x;
  const size_t wo = static_cast<size_t>(dir);
  float whh[H], wih[kMaxI];
#pragma unroll
  for (int k = 0; k < H; ++k) whh[k] = w_hh[(wo * G + j) * H + k];
#pragma unroll
  for (int i = 0; i < kMaxI; ++i) wih[i] = i < I ? w_ih[(wo * G + j) * I + i] : 0.f;
  const float bias = b_ih[wo * G + j] + b_hh[wo * G + j];
  constexpr int PAIRS = kBT * H / G;  // (sample, unit) pairs per thread
  float c_reg[PAIRS];
#pragma unroll
  for (int p = 0; p < PAIRS; ++p) c_reg[p] = 0.f;
  for (int i = j; i < kBT * H; i += G) (&h_s[0][0])[i] = 0.f;
  const int OW = ndir * H;
  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    for (int i = j; i < kBT * I; i += G) {
      const int b = i / I, f = i % I;
      x_s[b][f] = (n0 + b) < N ? x[(static_cast<size_t>(n0 + b) * T + t) * I + f] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int b = 0; b < kBT; ++b) {
      float acc = bias;
      const float4* hv = reinterpret_cast<const float4*>(&h_s[b][0]);
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 h4 = hv[k4];
        acc = fmaf(whh[4 * k4], h4.x, acc);
        acc = fmaf(whh[4 * k4 + 1], h4.y, acc);
        acc = fmaf(whh[4 * k4 + 2], h4.z, acc);
        acc = fmaf(whh[4 * k4 + 3], h4.w, acc);
      }
#pragma unroll
      for (int i = 0; i < kMaxI; ++i)
        if (i < I) acc = fmaf(wih[i], x_s[b][i], acc);
      g_s[b][j] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
      const int pi = j + p * G;
      const int b = pi / H, u = pi % H;
      const float ig = sigmoidf_(g_s[b][u]);
      const float fg = sigmoidf_(g_s[b][H + u]);
      const float gg = tanhf(g_s[b][2 * H + u]);
      const float og = sigmoidf_(g_s[b][3 * H + u]);
      const float c = fmaf(fg, c_reg[p], ig * gg);
      c_reg[p] = c;
      const float h = og * tanhf(c);
      h_s[b][u] = h;
      if (n0 + b < N) {
        const size_t row = static_cast<size_t>(n0 + b) * T + t;
        out[row * OW + dir * H + u] = h;
        if (gates) {
          float* gp = gates + ((wo * N + n0 + b) * T + t) * G;
          gp[u] = ig;
          gp[H + u] = fg;
          gp[2 * H + u] = gg;
          gp[3 * H + u] = og;
          cseq[((wo * N + n0 + b) * T + t) * H + u] = c;
        }
      }
    }
    __syncthreads();
  }
}

// BPTT. dout: [N][T][ndir*H]; gates/cseq saved by the forward; out = forward hidden states.
// Accumulates dw_ih [ndir][4H][I], dw_hh [ndir][4H][H], db [ndir][4H] (atomics; zero first); dx optional.
template <int H>
__global__ void __launch_bounds__(4 * H) lstm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w_ih,
                                                         const float* __restrict__ w_hh, const float* __restrict__ out,
                                                         const float* __restrict__ gates, const float* __restrict__ cseq,
                                                         const float* __restrict__ dout, float* __restrict__ dw_ih,
                                                         float* __restrict__ dw_hh, float* __restrict__ db,
                                                         float* __restrict__ dx, int N, int T, int I, int ndir) {
  constexpr int G = 4 * H;
  extern __shared__ float sm[];
  float* w_s = sm;                       // [G][H]   W_hh of this direction
  float* wi_s = w_s + G * H;             // [G][kMaxI]
  float* dg_s = wi_s + G * kMaxI;        // [kBT][G]
  float* hp_s = dg_s + kBT * G;          // [kBT][H]  h_{prev}
  float* dh_s = hp_s + kBT * H;          // [kBT][H]  recurrent dh
  float* x_s = dh_s + kBT * H;           // [kBT][kMaxI]
  const int dir = blockIdx.y;
  const int n0 = blockIdx.x * kBT;
  const int j = threadIdx.x;
  const size_t wo = static_cast<size_t>(dir);
  for (int i = j; i < G * H; i += G) w_s[i] = w_hh[wo * G * H + i];
  for (int i = j; i < G * kMaxI; i += G) {
    const int r = i / kMaxI, f = i % kMaxI;
    wi_s[i] = f < I ? w_ih[(wo * G + r) * I + f] : 0.f;
  }
  for (int i = j; i < kBT * H; i += G) dh_s[i] = 0.f;
  float dwhh[H], dwih[kMaxI], dbias = 0.f;
#pragma unroll
  for (int k = 0; k < H; ++k) dwhh[k] = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxI; ++i) dwih[i] = 0.f;
  constexpr int PAIRS = kBT * H / G;
  float dc_reg[PAIRS];
#pragma unroll
  for (int p = 0; p < PAIRS; ++p) dc_reg[p] = 0.f;
  const int OW = ndir * H;
  __syncthreads();
  for (int s = T - 1; s >= 0; --s) {
    const int t = dir == 0 ? s : T - 1 - s;            // time of this step
    const int tp = dir == 0 ? t - 1 : t + 1;           // time of the previous step in forward order
    const bool has_prev = s > 0;
    // stage x_t and h_prev
    for (int i = j; i < kBT * I; i += G) {
      const int b = i / I, f = i % I;
      x_s[b * kMaxI + f] = (n0 + b) < N ? x[(static_cast<size_t>(n0 + b) * T + t) * I + f] : 0.f;
    }
    for (int i = j; i < kBT * H; i += G) {
      const int b = i / H, u = i % H;
      hp_s[i] = (has_prev && n0 + b < N) ? out[(static_cast<size_t>(n0 + b) * T + tp) * OW + dir * H + u] : 0.f;
    }
    // A. gate gradients for the (sample, unit) pairs of this thread
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
      const int pi = j + p * G;
      const int b = pi / H, u = pi % H;
      float di = 0.f, df = 0.f, dg = 0.f, dov = 0.f;
      if (n0 + b < N) {
        const size_t row = (wo * N + n0 + b) * T + t;
        const float* gp = gates + row * G;
        const float ig = gp[u], fg = gp[H + u], gg = gp[2 * H + u], og = gp[3 * H + u];
        const float c = cseq[row * H + u];
        const float cp = has_prev ? cseq[((wo * N + n0 + b) * T + tp) * H + u] : 0.f;
        const float dh = dout[(static_cast<size_t>(n0 + b) * T + t) * OW + dir * H + u] + dh_s[b * H + u];
        const float tc = tanhf(c);
        dov = dh * tc * og * (1.f - og);
        const float dct = dc_reg[p] + dh * og * (1.f - tc * tc);
        di = dct * gg * ig * (1.f - ig);
        df = dct * cp * fg * (1.f - fg);
        dg = dct * ig * (1.f - gg * gg);
        dc_reg[p] = dct * fg;
      }
      dg_s[b * G + u] = di;
      dg_s[b * G + H + u] = df;
      dg_s[b * G + 2 * H + u] = dg;
      dg_s[b * G + 3 * H + u] = dov;
    }
    __syncthreads();
    // B1. weight-gradient rows of gate row j
    for (int b = 0; b < kBT; ++b) {
      const float d = dg_s[b * G + j];
      dbias += d;
      const float4* hv = reinterpret_cast<const float4*>(hp_s + b * H);
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 h4 = hv[k4];
        dwhh[4 * k4] = fmaf(d, h4.x, dwhh[4 * k4]);
        dwhh[4 * k4 + 1] = fmaf(d, h4.y, dwhh[4 * k4 + 1]);
        dwhh[4 * k4 + 2] = fmaf(d, h4.z, dwhh[4 * k4 + 2]);
        dwhh[4 * k4 + 3] = fmaf(d, h4.w, dwhh[4 * k4 + 3]);
      }
#pragma unroll
      for (int i = 0; i < kMaxI; ++i)
        if (i < I) dwih[i] = fmaf(d, x_s[b * kMaxI + i], dwih[i]);
    }
    __syncthreads();  // dh_s of this step fully consumed (A) before it is overwritten
    // B2. dh_{prev}[b][k] = sum_r dgates[b][r] * W_hh[r][k]
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
      const int pi = j + p * G;
      const int b = pi / H, k = pi % H;
      float acc = 0.f;
#pragma unroll 8
      for (int r = 0; r < G; ++r) acc = fmaf(dg_s[b * G + r], w_s[r * H + k], acc);
      dh_s[b * H + k] = acc;
    }
    // B3. dx_t[b][i] = sum_r dgates[b][r] * W_ih[r][i]
    if (dx) {
      for (int pi = j; pi < kBT * I; pi += G) {
        const int b = pi / I, f = pi % I;
        float acc = 0.f;
        for (int r = 0; r < G; ++r) acc = fmaf(dg_s[b * G + r], wi_s[r * kMaxI + f], acc);
        if (n0 + b < N) atomicAdd(dx + (static_cast<size_t>(n0 + b) * T + t) * I + f, acc);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < H; ++k) atomicAdd(dw_hh + (wo * G + j) * H + k, dwhh[k]);
#pragma unroll
  for (int i = 0; i < kMaxI; ++i)
    if (i < I) atomicAdd(dw_ih + (wo * G + j) * I + i, dwih[i]);
  atomicAdd(db + wo * G + j, dbias);
}

}  // namespace fmm

using namespace fmm;

extern "C" {

// x [N][T][I] fp32; weights stacked per direction: w_ih [ndir][4H][I], w_hh [ndir][4H][H], b_* [ndir][4H].
// out [N][T][ndir*H]; gates [ndir][N][T][4H] and cseq [ndir][N][T][H] are written when non-NULL (training).
int fmm_lstm_fwd(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                 float* out, float* gates, float* cseq, int N, int T, int I, int H, int ndir, cudaStream_t stream) {
  FMM_CHECK_ARG(x && w_ih && w_hh && b_ih && b_hh && out && N > 0 && T > 0, "lstm_fwd: bad args");
  FMM_CHECK_ARG(H == 64 && I >= 1 && I <= kMaxI && (ndir == 1 || ndir == 2), "lstm_fwd: H must be 64, I <= 32");
  FMM_CHECK_ARG((gates == nullptr) == (cseq == nullptr), "lstm_fwd: gates and cseq go together");
  dim3 grid((N + kBT - 1) / kBT, ndir);
  lstm_fwd_kernel<64><<<grid, 256, 0, stream>>>(x, w_ih, w_hh, b_ih, b_hh, out, gates, cseq, N, T, I, ndir);
  FMM_CHECK_LAUNCH("lstm_fwd");
  return FMM_OK;
}

int fmm_lstm_bwd(const float* x, const float* w_ih, const float* w_hh, const float* out, const float* gates,
                 const float* cseq, const float* dout, float* dw_ih, float* dw_hh, float* db, float* dx, int N,
                 int T, int I, int H, int ndir, cudaStream_t stream) {
  FMM_CHECK_ARG(x && w_ih && w_hh && out && gates && cseq && dout && dw_ih && dw_hh && db, "lstm_bwd: bad args");
  FMM_CHECK_ARG(H == 64 && I >= 1 && I <= kMaxI && (ndir == 1 || ndir == 2), "lstm_bwd: H must be 64, I <= 32");
  const size_t smem = sizeof(float) * (256 * 64 + 256 * kMaxI + kBT * 256 + 2 * kBT * 64 + kBT * kMaxI);
  cudaError_t e = cudaFuncSetAttribute(lstm_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("lstm_bwd: smem attribute: %s", cudaGetErrorString(e));
    return FMM_ERR_SMEM;
  }
  dim3 grid((N + kBT - 1) / kBT, ndir);
  lstm_bwd_kernel<64><<<grid, 256, smem, stream>>>(x, w_ih, w_hh, out, gates, cseq, dout, dw_ih, dw_hh, db, dx, N, T, I, ndir);
  FMM_CHECK_LAUNCH("lstm_bwd");
  return FMM_OK;
}

}  // extern "C"
