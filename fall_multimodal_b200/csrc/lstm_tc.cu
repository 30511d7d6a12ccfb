// Bidirectional LSTM recurrence on tensor cores, inference path (reference F2/Model/bilstm.py:29,48: nn.LSTM(I, 64,
// bidirectional) from zero state, then mean over T or the last step; BASELINE config 5: 8192 windows of 128 x 6).
//
// Per time step the cell is one small GEMM  gates[256][windows] = W[256][64 + I + 1] . [h ; x_t ; 1]  followed by the gate
// non-linearities. A CTA owns 64 windows of one direction for the whole sequence:
//   * warp w owns hidden units 8w .. 8w+7 and keeps THEIR 32 weight rows (i, f, g, o) as mma.sync A fragments in registers
//     for all T steps (rows permuted so that one lane's accumulators are the four gates of one unit for two windows: the
//     cell update needs no exchange between lanes),
//   * the state [h ; x_t ; 1] is the B operand, double buffered in shared memory (one __syncthreads per step),
//   * fp32 accuracy comes from the 3xTF32 split (hi*hi + hi*lo + lo*hi, error ~2^-21) - the fp32 parity gate of the
//     recurrent branch is 1e-4 over 128 dependent steps,
//   * c stays in registers; the mean / last-step feature is accumulated in registers and written once.
// The round-1 kernel (csrc/lstm.cu: scalar FMA, 16 windows per CTA) remains the training path (it saves gates / cell states
// for BPTT); it ran this inference in 8.6 ms.
#include "common.cuh"

namespace fmm {

constexpr int kLtWin = 64;          // windows per CTA
constexpr int kLtPitch = 2 * 64 + 8; // floats per state row: 64 windows x (hi, lo) + 8 (conflict-free 8-byte B-fragment loads)
constexpr int kLtH = 64;

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// MUFU-based exp / reciprocal (2 ulp each): the accurate expf / tanhf / division sequences made the gate math 60 % of the
// kernel's instructions. tanh(x) = 2*sigmoid(2x) - 1 (absolute error ~2e-7).
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, sigm(2.f * x), -1.f); }
// the state is stored split for the 3xTF32 product: hi = tf32(v), lo = tf32(v - hi), once by the writer instead of by all 8 warps
__device__ __forceinline__ float2 split_tf32(float v) {
  const uint32_t h = to_tf32(v);
  return make_float2(__uint_as_float(h), __uint_as_float(to_tf32(v - __uint_as_float(h))));
}

// KS = k-steps of 8 covering K = 64 + I + 1 (zero padded)
template <int KS>
__global__ void __launch_bounds__(256, 1)
lstm_infer_kernel(const float* __restrict__ x, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                  const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ feat, int N, int T, int I,
                  int ndir, int mean_feature) {
  extern __shared__ float S[];                  // [2][KS*8][kLtPitch]: per row 64 x (hi, lo)
  constexpr int XR = (kLtWin * (8 * KS - 65) + 255) / 256;   // input elements per thread and step
  const int dir = blockIdx.y;
  const int n0 = blockIdx.x * kLtWin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int KP = KS * 8;
  const int G = 4 * kLtH;
  const float* Wih = w_ih + static_cast<size_t>(dir) * G * I;
  const float* Whh = w_hh + static_cast<size_t>(dir) * G * kLtH;
  const float* Bi = b_ih + static_cast<size_t>(dir) * G;
  const float* Bh = b_hh + static_cast<size_t>(dir) * G;

  // weight fragments: m-tile 0 rows = (i | f) of units 8w..8w+7, m-tile 1 rows = (g | o)
  auto wval = [&](int gate, int k) -> float {
    const int row = gate * kLtH + 8 * warp + g;
    if (k < kLtH) return Whh[static_cast<size_t>(row) * kLtH + k];
    if (k < kLtH + I) return Wih[static_cast<size_t>(row) * I + (k - kLtH)];
    if (k == kLtH + I) return Bi[row] + Bh[row];
    return 0.f;
  };
  float wf[2][KS][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      wf[mt][ks][0] = wval(2 * mt, ks * 8 + t4);
      wf[mt][ks][1] = wval(2 * mt + 1, ks * 8 + t4);
      wf[mt][ks][2] = wval(2 * mt, ks * 8 + t4 + 4);
      wf[mt][ks][3] = wval(2 * mt + 1, ks * 8 + t4 + 4);
    }

  // state init: h = 0, x rows of the first step, ones row, zero padding
  const int tfirst = dir == 0 ? 0 : T - 1;
  for (int i = threadIdx.x; i < 2 * KP * kLtPitch; i += blockDim.x) S[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < kLtWin * I; i += blockDim.x) {
    const int w = i / I, ci = i - w * I;
    if (n0 + w < N)
      *reinterpret_cast<float2*>(S + (kLtH + ci) * kLtPitch + 2 * w) = split_tf32(x[(static_cast<size_t>(n0 + w) * T + tfirst) * I + ci]);
  }
  for (int i = threadIdx.x; i < 2 * kLtWin; i += blockDim.x)
    S[(i / kLtWin) * KP * kLtPitch + (kLtH + I) * kLtPitch + 2 * (i % kLtWin)] = 1.f;     // ones row: hi = 1, lo = 0
  __syncthreads();

  float c[16], fsum[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = fsum[i] = 0.f;

  int buf = 0;
  for (int step = 0; step < T; ++step) {
    const int tcur = dir == 0 ? step : T - 1 - step;
    const float* Sc = S + buf * KP * kLtPitch;
    float* Sn = S + (buf ^ 1) * KP * kLtPitch;
    // prefetch the next step's inputs (global latency hides behind the MMAs)
    const int tnext = dir == 0 ? step + 1 : T - 2 - step;
    float xn[XR];
    const bool more = step + 1 < T;
#pragma unroll
    for (int r = 0; r < XR; ++r) {
      const int i = threadIdx.x + r * 256;
      xn[r] = 0.f;
      if (more && i < kLtWin * I) {
        const int w = i / I, ci = i - w * I;
        if (n0 + w < N) xn[r] = x[(static_cast<size_t>(n0 + w) * T + tnext) * I + ci];
      }
    }
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ah[mt][e] = to_tf32(wf[mt][ks][e]);
          al[mt][e] = to_tf32(wf[mt][ks][e] - __uint_as_float(ah[mt][e]));
        }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 v0 = *reinterpret_cast<const float2*>(Sc + (ks * 8 + t4) * kLtPitch + 2 * (nt * 8 + g));
        const float2 v1 = *reinterpret_cast<const float2*>(Sc + (ks * 8 + t4 + 4) * kLtPitch + 2 * (nt * 8 + g));
        const uint32_t bh0 = __float_as_uint(v0.x), bh1 = __float_as_uint(v1.x);
        const uint32_t bl0 = __float_as_uint(v0.y), bl1 = __float_as_uint(v1.y);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], al[mt], bh0, bh1);
          mma_tf32(acc[mt][nt], ah[mt], bl0, bl1);
          mma_tf32(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
    // cell update: lane holds (i, f) in acc[0][nt] = {i(w0), i(w1), f(w0), f(w1)} and (g, o) in acc[1][nt] for windows nt*8+2*t4 (+1)
    const int j = 8 * warp + g;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float hv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ig = sigm(acc[0][nt][e]), fg = sigm(acc[0][nt][2 + e]);
        const float gg = tanh_fast(acc[1][nt][e]), og = sigm(acc[1][nt][2 + e]);
        const float cn = fmaf(fg, c[nt * 2 + e], ig * gg);
        c[nt * 2 + e] = cn;
        hv[e] = og * tanh_fast(cn);
        if (mean_feature) fsum[nt * 2 + e] += hv[e];
        else if (tcur == T - 1) fsum[nt * 2 + e] = hv[e];       // out[:, -1, :] of this direction
      }
      const float2 s0 = split_tf32(hv[0]), s1 = split_tf32(hv[1]);
      *reinterpret_cast<float4*>(Sn + j * kLtPitch + 2 * (nt * 8 + 2 * t4)) = make_float4(s0.x, s0.y, s1.x, s1.y);
    }
#pragma unroll
    for (int r = 0; r < XR; ++r) {
      const int i = threadIdx.x + r * 256;
      if (more && i < kLtWin * I) {
        const int w = i / I, ci = i - w * I;
        *reinterpret_cast<float2*>(Sn + (kLtH + ci) * kLtPitch + 2 * w) = split_tf32(xn[r]);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  const float scale = mean_feature ? 1.f / T : 1.f;
  const int j = 8 * warp + g;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = n0 + nt * 8 + 2 * t4 + e;
      if (n < N) feat[static_cast<size_t>(n) * (ndir * kLtH) + dir * kLtH + j] = fsum[nt * 2 + e] * scale;
    }
}

}  // namespace fmm

using namespace fmm;

extern "C" {

// feat [N][ndir*64] = mean over T (mean_feature != 0) or the t = T-1 output of nn.LSTM(I, 64, bidirectional) from zero state.
int fmm_lstm_infer(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* feat, int N,
                   int T, int I, int H, int ndir, int mean_feature, cudaStream_t stream) {
  FMM_CHECK_ARG(x && w_ih && w_hh && b_ih && b_hh && feat && N > 0 && T > 0, "lstm_infer: bad args");
  FMM_CHECK_ARG(H == 64 && I >= 1 && I <= 39 && (ndir == 1 || ndir == 2), "lstm_infer: H must be 64, I <= 39");
  const int KS = (64 + I + 1 + 7) / 8;
  dim3 grid((N + kLtWin - 1) / kLtWin, ndir);
  const size_t smem = sizeof(float) * 2 * KS * 8 * kLtPitch;
#define FMM_LT(K_)                                                                                                  \
  case K_: {                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(lstm_infer_kernel<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) {                                                                                         \
      set_last_error("lstm_infer: smem attribute: %s", cudaGetErrorString(e));                                      \
      return FMM_ERR_SMEM;                                                                                          \
    }                                                                                                               \
    lstm_infer_kernel<K_><<<grid, 256, smem, stream>>>(x, w_ih, w_hh, b_ih, b_hh, feat, N, T, I, ndir, mean_feature); \
  } break;
  switch (KS) {
    FMM_LT(9) FMM_LT(10) FMM_LT(11) FMM_LT(12) FMM_LT(13)
    default:
      set_last_error("lstm_infer: no instantiation for I=%d", I);
      return FMM_ERR_ARG;
  }
#undef FMM_LT
  FMM_CHECK_LAUNCH("lstm_infer");
  return FMM_OK;
}

}  // extern "C"
