// Memory-bound kernels of the TRAGCN family (reference: EmbGCN.py:59-89, GRU.py:17-26, TA.py:40-69):
// the concat + adaptive-adjacency mix feeding the per-node products, the gate / candidate / state
// update of the graph GRU and their backward counterparts, row softmax, LayerNorm(a+b), positional
// encoding, ReLU masking. The matrix products in between run in bgemm.cu.
#include "common.cuh"

namespace fmm {

constexpr int kMaxV = 32;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dsilu_(float x) {
  float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}

// ---------------------------------------------------------------------------------------------
// catmix: cat = [x_t | h (*r) | 1 | 0-pad];  XC1 = cat;  XC0 = S . cat (the bias column stays 1)
// one block per clip
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) catmix_kernel(const T* __restrict__ x, long long xb, long long xv,
                                                     const T* __restrict__ h, long long hb, long long hv,
                                                     const T* __restrict__ r, long long rb, long long rv,
                                                     const float* __restrict__ S, T* __restrict__ xc0,
                                                     T* __restrict__ xc1, int V, int Din, int H, int Cp) {
  extern __shared__ float sm[];
  float* cat = sm;             // [V][Cp]
  float* Ss = sm + V * Cp;     // [V][V]
  const int b = blockIdx.x, Cin = Din + H;
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) Ss[i] = S[i];
  for (int i = threadIdx.x; i < V * Cp; i += blockDim.x) {
    int m = i / Cp, c = i % Cp;
    float v = 0.f;
    if (c < Din) {
      v = to_f32(x[b * xb + m * xv + c]);
    } else if (c < Cin) {
      if (h) {
        v = to_f32(h[b * hb + m * hv + (c - Din)]);
        if (r) v = to_f32(from_f32<T>(v * to_f32(r[b * rb + m * rv + (c - Din)])));  // r*state is a rounded tensor in the reference
      }
    } else if (c == Cin) {
      v = 1.f;
    }
    cat[i] = v;
  }
  __syncthreads();
  const long long ob = (long long)b * V * Cp;
  for (int i = threadIdx.x; i < V * Cp; i += blockDim.x) {
    int n = i / Cp, c = i % Cp;
    float acc;
    if (c < Cin) {
      acc = 0.f;
      for (int m = 0; m < V; ++m) acc += Ss[n * V + m] * cat[m * Cp + c];
    } else {
      acc = cat[i];
    }
    xc0[ob + i] = from_f32<T>(acc);
    xc1[ob + i] = from_f32<T>(cat[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// gate: out = act(pre + silu(lin)); mode 0: sigmoid -> zr.  mode 1: tanh -> hc, then h' = z*h + (1-z)*hc
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void gate_kernel(const float* __restrict__ pre, const float* __restrict__ lin, T* __restrict__ out,
                            T* __restrict__ lin_save, int mode, const T* __restrict__ z, long long zs,
                            const T* __restrict__ hprev, long long hb, long long hv, T* __restrict__ hout,
                            long long ob, long long ov, int B, int V, int C) {
  long long total = (long long)B * V * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long row = i / C;
    float l = lin[i];
    float a = pre[i] + silu_(l);
    lin_save[i] = from_f32<T>(l);
    if (mode == 0) {
      out[i] = from_f32<T>(sigmoidf_(a));
    } else {
      int v = (int)(row % V);
      long long b = row / V;
      float hc = to_f32(from_f32<T>(tanhf(a)));
      out[i] = from_f32<T>(hc);
      float zz = to_f32(z[row * zs + c]);
      float hp = hprev ? to_f32(hprev[b * hb + v * hv + c]) : 0.f;
      hout[b * ob + v * ov + c] = from_f32<T>(zz * hp + (1.f - zz) * hc);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, part 1 (state update + candidate): dh_tot = carry + dH_t
//   dz = dh_tot*(h - hc); carry = dh_tot*z; dpre_u = dh_tot*(1-z)*(1-hc^2); dlin_u = dpre_u*silu'(lin_u)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cell_bwd1_kernel(float* __restrict__ carry, const T* __restrict__ dH, long long db, long long dv,
                                 const T* __restrict__ z, long long zs, const T* __restrict__ hprev, long long hb,
                                 long long hv, const T* __restrict__ hc, const T* __restrict__ lu,
                                 float* __restrict__ dz, T* __restrict__ dpre, T* __restrict__ dlin, int B, int V, int H) {
  long long total = (long long)B * V * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    long long row = i / H;
    int v = (int)(row % V);
    long long b = row / V;
    float d = carry[i] + (dH ? to_f32(dH[b * db + v * dv + c]) : 0.f);
    float zz = to_f32(z[row * zs + c]);
    float hp = hprev ? to_f32(hprev[b * hb + v * hv + c]) : 0.f;
    float hcc = to_f32(hc[i]);
    dz[i] = d * (hp - hcc);
    carry[i] = d * zz;
    float dp = d * (1.f - zz) * (1.f - hcc * hcc);
    dpre[i] = from_f32<T>(dp);
    dlin[i] = from_f32<T>(dp * dsilu_(to_f32(lu[i])));
  }
}

// ---------------------------------------------------------------------------------------------
// backward, part 2: undo concat + mix.  dcat[m] = sum_n S[n][m]*dXC0[n] + dXC1[m];
//   dS[n][m] += sum_{c<Cin} dXC0[n][c]*cat[m][c]
//   mode 1 (candidate stage): dx += dcat[:Din]; drh = dcat[Din:]; dr = drh*h; carry += drh*r;
//                             dpre_g = [dz*z(1-z) | dr*r(1-r)]; dlin_g = dpre_g*silu'(lin_g)
//   mode 0 (gate stage):      dx += dcat[:Din]; carry += dcat[Din:]
// one block per clip
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) mix_bwd_kernel(const float* __restrict__ dxc0, const float* __restrict__ dxc1,
                                                      const T* __restrict__ cat, const float* __restrict__ S,
                                                      float* __restrict__ dS, int nrep, int mode, T* __restrict__ dx,
                                                      long long dxb, long long dxv, int dx_accum,
                                                      float* __restrict__ carry, const T* __restrict__ hprev,
                                                      long long hb, long long hv, const T* __restrict__ zr,
                                                      const float* __restrict__ dz, const T* __restrict__ lg,
                                                      T* __restrict__ dpre, T* __restrict__ dlin, int V, int Din,
                                                      int H, int Cp) {
  extern __shared__ float sm[];
  float* d0 = sm;                 // [V][Cp]
  float* ct = d0 + V * Cp;        // [V][Cp]
  float* Ss = ct + V * Cp;        // [V][V]
  const int b = blockIdx.x, Cin = Din + H;
  const long long ob = (long long)b * V * Cp;
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) Ss[i] = S[i];
  for (int i = threadIdx.x; i < V * Cp; i += blockDim.x) {
    d0[i] = dxc0[ob + i];
    ct[i] = to_f32(cat[ob + i]);
  }
  __syncthreads();
  // dS partial: V*V dot products of length Cin
  float* dSr = dS + (size_t)(b % nrep) * V * V;
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) {
    int n = i / V, m = i % V;
    float acc = 0.f;
    for (int c = 0; c < Cin; ++c) acc += d0[n * Cp + c] * ct[m * Cp + c];
    atomicAdd(&dSr[i], acc);
  }
  for (int i = threadIdx.x; i < V * Cin; i += blockDim.x) {
    int m = i / Cin, c = i % Cin;
    float acc = dxc1[ob + m * Cp + c];
    for (int n = 0; n < V; ++n) acc += Ss[n * V + m] * d0[n * Cp + c];
    if (c < Din) {
      if (dx) {
        T* p = dx + b * dxb + m * dxv + c;
        *p = from_f32<T>(dx_accum ? to_f32(*p) + acc : acc);
      }
      continue;
    }
    int j = c - Din;
    long long row = (long long)b * V + m;
    long long hi = row * H + j;
    if (mode == 0) {
      carry[hi] += acc;
    } else {
      float hp = hprev ? to_f32(hprev[b * hb + m * hv + j]) : 0.f;
      float zz = to_f32(zr[row * 2 * H + j]), rr = to_f32(zr[row * 2 * H + H + j]);
      carry[hi] += acc * rr;
      float dr = acc * hp;
      float gz = dz[hi] * zz * (1.f - zz), gr = dr * rr * (1.f - rr);
      long long o = row * 2 * H;
      dpre[o + j] = from_f32<T>(gz);
      dpre[o + H + j] = from_f32<T>(gr);
      dlin[o + j] = from_f32<T>(gz * dsilu_(to_f32(lg[o + j])));
      dlin[o + H + j] = from_f32<T>(gr * dsilu_(to_f32(lg[o + H + j])));
    }
  }
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

int fmm_tg_catmix(const void* x, long long xb, long long xv, const void* h, long long hb, long long hv, const void* r,
                  long long rb, long long rv, const float* S, void* xc0, void* xc1, int B, int V, int Din, int H, int Cp,
                  int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_catmix");
  FMM_CHECK_ARG(B > 0 && V > 0 && V <= kMaxV && Din > 0 && H > 0 && Cp >= Din + H + 1, "tg_catmix: bad shape (B %d V %d Din %d H %d Cp %d)",
                B, V, Din, H, Cp);
  size_t smem = sizeof(float) * ((size_t)V * Cp + (size_t)V * V);
  FMM_CHECK_ARG(smem <= 48 * 1024, "tg_catmix: V*Cp too large for shared memory");
  TG_DISPATCH(dtype, catmix_kernel<T><<<B, 256, smem, (cudaStream_t)stream>>>(
      (const T*)x, xb, xv, (const T*)h, hb, hv, (const T*)r, rb, rv, S, (T*)xc0, (T*)xc1, V, Din, H, Cp);)
  FMM_CHECK_LAUNCH("tg_catmix");
  return FMM_OK;
}

int fmm_tg_gate(const float* pre, const float* lin, void* out, void* lin_save, int mode, const void* z, long long zs,
                const void* hprev, long long hb, long long hv, void* hout, long long ob, long long ov, int B, int V, int C,
                int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_gate");
  FMM_CHECK_ARG(B > 0 && V > 0 && C > 0 && (mode == 0 || (z && hout)), "tg_gate: bad arguments");
  long long total = (long long)B * V * C;
  TG_DISPATCH(dtype, gate_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
      pre, lin, (T*)out, (T*)lin_save, mode, (const T*)z, zs, (const T*)hprev, hb, hv, (T*)hout, ob, ov, B, V, C);)
  FMM_CHECK_LAUNCH("tg_gate");
  return FMM_OK;
}

int fmm_tg_cell_bwd1(float* carry, const void* dH, long long db, long long dv, const void* z, long long zs,
                     const void* hprev, long long hb, long long hv, const void* hc, const void* lu, float* dz, void* dpre,
                     void* dlin, int B, int V, int H, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_cell_bwd1");
  long long total = (long long)B * V * H;
  FMM_CHECK_ARG(total > 0, "tg_cell_bwd1: empty");
  TG_DISPATCH(dtype, cell_bwd1_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
      carry, (const T*)dH, db, dv, (const T*)z, zs, (const T*)hprev, hb, hv, (const T*)hc, (const T*)lu, dz, (T*)dpre,
      (T*)dlin, B, V, H);)
  FMM_CHECK_LAUNCH("tg_cell_bwd1");
  return FMM_OK;
}

int fmm_tg_mix_bwd(const float* dxc0, const float* dxc1, const void* cat, const float* S, float* dS, int nrep, int mode,
                   void* dx, long long dxb, long long dxv, int dx_accum, float* carry, const void* hprev, long long hb,
                   long long hv, const void* zr, const float* dz, const void* lg, void* dpre, void* dlin, int B, int V,
                   int Din, int H, int Cp, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_mix_bwd");
  FMM_CHECK_ARG(B > 0 && V > 0 && V <= kMaxV && nrep > 0 && Cp >= Din + H + 1, "tg_mix_bwd: bad shape");
  FMM_CHECK_ARG(mode == 0 || (zr && dz && lg && dpre && dlin), "tg_mix_bwd: candidate stage needs the gate tensors");
  size_t smem = sizeof(float) * (2 * (size_t)V * Cp + (size_t)V * V);
  FMM_CHECK_ARG(smem <= 48 * 1024, "tg_mix_bwd: V*Cp too large for shared memory");
  TG_DISPATCH(dtype, mix_bwd_kernel<T><<<B, 256, smem, (cudaStream_t)stream>>>(
      dxc0, dxc1, (const T*)cat, S, dS, nrep, mode, (T*)dx, dxb, dxv, dx_accum, carry, (const T*)hprev, hb, hv,
      (const T*)zr, dz, (const T*)lg, (T*)dpre, (T*)dlin, V, Din, H, Cp);)
  FMM_CHECK_LAUNCH("tg_mix_bwd");
  return FMM_OK;
}

int fmm_tg_softmax_fwd(void* x, long long rows, int L, int Lp, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_softmax_fwd");
  FMM_CHECK_ARG(rows > 0 && L > 0 && Lp >= L, "tg_softmax_fwd: bad shape");
  TG_DISPATCH(dtype, softmax_fwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((T*)x, rows, L, Lp);)
  FMM_CHECK_LAUNCH("tg_softmax_fwd");
  return FMM_OK;
}

int fmm_tg_softmax_bwd(const void* p, void* dp, long long rows, int L, int Lp, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_softmax_bwd");
  FMM_CHECK_ARG(rows > 0 && L > 0 && Lp >= L, "tg_softmax_bwd: bad shape");
  TG_DISPATCH(dtype, softmax_bwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const T*)p, (T*)dp, rows, L, Lp);)
  FMM_CHECK_LAUNCH("tg_softmax_bwd");
  return FMM_OK;
}

int fmm_tg_ln_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                  long long rows, int C, float eps, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_ln_fwd");
  FMM_CHECK_ARG(rows > 0 && C >= 32 && C <= 256 && (C % 32) == 0, "tg_ln_fwd: C must be a multiple of 32 up to 256, got %d", C);
  TG_DISPATCH(dtype, ln_fwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const T*)a, (const T*)b, gamma, beta, (T*)y, mean, rstd, rows, C, eps);)
  FMM_CHECK_LAUNCH("tg_ln_fwd");
  return FMM_OK;
}

int fmm_tg_ln_bwd(const void* dy, const void* a, const void* b, const float* gamma, const float* mean, const float* rstd,
                  void* dx, float* dgamma, float* dbeta, long long rows, int C, int dtype, void* stream) {
  TG_CHECK_DT(dtype, "tg_ln_bwd");
  FMM_CHECK_ARG(rows > 0 && C >= 32 && C <= 256 && (C % 32) == 0, "tg_ln_bwd: C must be a multiple of 32 up to 256, got %d", C);
  long long warps = (long long)num_sms() * 8 * 4;
  int rpw = (int)((rows + warps - 1) / warps);
  if (rpw < 1) rpw = 1;
  long long blocks = (rows + 8ll * rpw - 1) / (8ll * rpw);
  TG_DISPATCH(dtype, ln_bwd_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const T*)dy, (const T*)a, (const T*)b, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows, C, rpw);)
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
