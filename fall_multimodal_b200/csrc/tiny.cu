// Small per-channel / per-clip kernels between the big passes: BatchNorm finalisation, the
// squeeze-excite MLP (forward and backward) and the coefficient algebra that lets the BatchNorm /
// SE / ReLU backward be applied as ONE affine elementwise pass (csrc/elementwise.cu).
//
// Reference semantics: nn.BatchNorm{1,2}d (train: biased batch variance for normalisation,
// unbiased into running_var, momentum 0.1, eps 1e-5) and Channel_Attention
// (/root/reference/Fall_2_Spatial_Temporal_SR/Model/stgcan.py:59-74): s = sigmoid(W2 relu(BN(W1 p + b1)) + b2),
// p = mean_{t,v} of the tcn output; its BatchNorm2d sees a 1x1 map, i.e. normalises over N only.
#include "common.cuh"

namespace fmm {

// ---------------------------------------------------------------------------------------------
// bn_finalize: (sum, sumsq, count) -> mean, rstd, scale a = gamma*rstd, shift b = beta - mean*a
// ---------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ ch_sum, const double* __restrict__ ch_sq, int nrep, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                   int training, float* __restrict__ a, float* __restrict__ b,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (training) {
    double s = 0, q = 0;
    for (int r = 0; r < nrep; ++r) {
      s += ch_sum[static_cast<size_t>(r) * C + c];
      q += ch_sq[static_cast<size_t>(r) * C + c];
    }
    mean = s / count;
    var = q / count - mean * mean;
    if (var < 0) var = 0;
    if (rmean) {
      const double unb = count > 1 ? var * count / (count - 1) : var;
      rmean[c] = static_cast<float>((1.0 - momentum) * rmean[c] + momentum * mean);
      rvar[c] = static_cast<float>((1.0 - momentum) * rvar[c] + momentum * unb);
    }
  } else {
    mean = rmean[c];
    var = rvar[c];
  }
  const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
  const double g = gamma ? gamma[c] : 1.0;
  const double sc = g * rstd;
  a[c] = static_cast<float>(sc);
  b[c] = static_cast<float>((beta ? beta[c] : 0.0) - mean * sc);
  if (mean_out) mean_out[c] = static_cast<float>(mean);
  if (rstd_out) rstd_out[c] = static_cast<float>(rstd);
}

// ---------------------------------------------------------------------------------------------
// SE forward
// ---------------------------------------------------------------------------------------------
// a) p[n,c] = a2[c]*pool[n,c]*invM + b2[c];  h[n,j] = b1[j] + sum_c W1[j,c] p[n,c]     (grid = N)
__global__ void se_fwd_a_kernel(const float* __restrict__ pool, const float* __restrict__ a2,
                                const float* __restrict__ b2, float invM, const float* __restrict__ W1,
                                const float* __restrict__ b1, float* __restrict__ p, float* __restrict__ h, int C,
                                int C4) {
  extern __shared__ float sp[];  // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = fmaf(a2[c], pool[static_cast<size_t>(n) * C + c] * invM, b2[c]);
    sp[c] = v;
    p[static_cast<size_t>(n) * C + c] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < C4; j += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(W1[static_cast<size_t>(j) * C + c], sp[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) h[static_cast<size_t>(n) * C4 + j] = acc + b1[j];
  }
}

// b) BatchNorm over N of h[:, j]  (grid = C4, one block per hidden channel)
__global__ void se_bn_kernel(const float* __restrict__ h, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                             float momentum, float eps, int training, float* __restrict__ ah, float* __restrict__ bh,
                             float* __restrict__ hmean, float* __restrict__ hrstd, int N, int C4) {
  __shared__ double ss[32], sq[32];
  const int j = blockIdx.x;
  double s = 0, q = 0;
  if (training) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const double v = h[static_cast<size_t>(n) * C4 + j];
      s += v;
      q += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
      ss[threadIdx.x >> 5] = s;
      sq[threadIdx.x >> 5] = q;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double mean, var;
    if (training) {
      s = 0;
      q = 0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) {
        s += ss[w];
        q += sq[w];
      }
      mean = s / N;
      var = q / N - mean * mean;
      if (var < 0) var = 0;
      const double unb = N > 1 ? var * N / (N - 1.0) : var;
      rmean[j] = static_cast<float>((1.0 - momentum) * rmean[j] + momentum * mean);
      rvar[j] = static_cast<float>((1.0 - momentum) * rvar[j] + momentum * unb);
    } else {
      mean = rmean[j];
      var = rvar[j];
    }
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double sc = gamma[j] * rstd;
    ah[j] = static_cast<float>(sc);
    bh[j] = static_cast<float>(beta[j] - mean * sc);
    hmean[j] = static_cast<float>(mean);
    hrstd[j] = static_cast<float>(rstd);
  }
}

// c) r = relu(ah*h+bh); s = sigmoid(b2se + W2 r); k1 = s*a2; k0 = s*b2          (grid = N)
__global__ void se_fwd_b_kernel(const float* __restrict__ h, const float* __restrict__ ah,
                                const float* __restrict__ bh, const float* __restrict__ W2,
                                const float* __restrict__ b2se, const float* __restrict__ a2,
                                const float* __restrict__ b2, float* __restrict__ s_out, float* __restrict__ k1,
                                float* __restrict__ k0, int C, int C4) {
  extern __shared__ float sr[];  // [C4]
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < C4; j += blockDim.x)
    sr[j] = fmaxf(fmaf(ah[j], h[static_cast<size_t>(n) * C4 + j], bh[j]), 0.f);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b2se[c];
    for (int j = 0; j < C4; ++j) acc = fmaf(W2[static_cast<size_t>(c) * C4 + j], sr[j], acc);
    const float s = 1.f / (1.f + expf(-acc));
    s_out[static_cast<size_t>(n) * C + c] = s;
    k1[static_cast<size_t>(n) * C + c] = s * a2[c];
    k0[static_cast<size_t>(n) * C + c] = s * b2[c];
  }
}

// ---------------------------------------------------------------------------------------------
// SE backward
// ---------------------------------------------------------------------------------------------
// a) ds = a2*S2 + b2*S1; dq = ds*s*(1-s); dhr[n,j] = (sum_c dq[c] W2[c,j]) * (ah*h+bh > 0)    (grid = N)
__global__ void se_bwd_a_kernel(const float* __restrict__ S1, const float* __restrict__ S2,
                                const float* __restrict__ a2, const float* __restrict__ b2,
                                const float* __restrict__ s, const float* __restrict__ h,
                                const float* __restrict__ ah, const float* __restrict__ bh,
                                const float* __restrict__ W2, float* __restrict__ dq, float* __restrict__ dhr,
                                float* __restrict__ r_out, int C, int C4) {
  extern __shared__ float sdq[];  // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t i = static_cast<size_t>(n) * C + c;
    const float ds = fmaf(a2[c], S2[i], b2[c] * S1[i]);
    const float sv = s[i];
    const float v = ds * sv * (1.f - sv);
    sdq[c] = v;
    dq[i] = v;
  }
  __syncthreads();
  // dhr[j] = sum_c dq[c] W2[c][j]: all threads take part - (blockDim / C4) partial sums per hidden channel j, combined
  // through shared memory (the C4 <= 64 threads x C serial steps of the first version made this a 26 us kernel)
  float* part = sdq + C;                       // [blockDim.x]
  const int G = C4 <= static_cast<int>(blockDim.x) ? blockDim.x / C4 : 1;
  if (C4 <= static_cast<int>(blockDim.x)) {
    const int j = threadIdx.x % C4, g = threadIdx.x / C4;
    float acc = 0.f;
    if (g < G)
      for (int c = g; c < C; c += G) acc = fmaf(sdq[c], W2[static_cast<size_t>(c) * C4 + j], acc);
    part[threadIdx.x] = acc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C4; j += blockDim.x) {
    float acc = 0.f;
    if (C4 <= static_cast<int>(blockDim.x)) {
      for (int g = 0; g < G; ++g) acc += part[g * C4 + j];
    } else {
      for (int c = 0; c < C; ++c) acc = fmaf(sdq[c], W2[static_cast<size_t>(c) * C4 + j], acc);
    }
    const float pre = fmaf(ah[j], h[static_cast<size_t>(n) * C4 + j], bh[j]);
    dhr[static_cast<size_t>(n) * C4 + j] = pre > 0.f ? acc : 0.f;
    r_out[static_cast<size_t>(n) * C4 + j] = fmaxf(pre, 0.f);
  }
}

// b) BatchNorm-over-N backward on hidden channel j: dgamma, dbeta, dh   (grid = C4)
__global__ void se_bwd_bn_kernel(const float* __restrict__ dhr, const float* __restrict__ h,
                                 const float* __restrict__ ah, const float* __restrict__ hmean,
                                 const float* __restrict__ hrstd, int training, float* __restrict__ dh,
                                 float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int C4) {
  __shared__ double s1[32], s2[32];
  __shared__ double m1, m2;
  const int j = blockIdx.x;
  const double mu = hmean[j], rs = hrstd[j];
  double a = 0, b = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double d = dhr[static_cast<size_t>(n) * C4 + j];
    const double hh = (h[static_cast<size_t>(n) * C4 + j] - mu) * rs;
    a += d;
    b += d * hh;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s1[threadIdx.x >> 5] = a;
    s2[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0;
    b = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) {
      a += s1[w];
      b += s2[w];
    }
    dbeta[j] += static_cast<float>(a);
    dgamma[j] += static_cast<float>(b);
    m1 = a / N;
    m2 = b / N;
  }
  __syncthreads();
  const double sc = ah[j];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double d = dhr[static_cast<size_t>(n) * C4 + j];
    const double hh = (h[static_cast<size_t>(n) * C4 + j] - mu) * rs;
    dh[static_cast<size_t>(n) * C4 + j] = static_cast<float>(training ? sc * (d - m1 - hh * m2) : sc * d);
  }
}

// c) dp[n,c] = sum_j dh[n,j] W1[j,c]      (grid = N)
__global__ void se_bwd_dp_kernel(const float* __restrict__ dh, const float* __restrict__ W1, float* __restrict__ dp,
                                 int C, int C4) {
  extern __shared__ float sdh[];  // [C4]
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < C4; j += blockDim.x) sdh[j] = dh[static_cast<size_t>(n) * C4 + j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < C4; ++j) acc = fmaf(sdh[j], W1[static_cast<size_t>(j) * C + c], acc);
    dp[static_cast<size_t>(n) * C + c] = acc;
  }
}

// out[p][q] += sum_n A[n][p] * B[n][q]  and (optional) colsum[p] += sum_n A[n][p]   (grid = P, 256 threads:
// q lanes x n-splits, reduced through shared memory)
__global__ void small_tn_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out,
                                     float* __restrict__ colsum, int N, int P, int Q) {
  __shared__ float red[256];
  __shared__ float cs[256];
  const int pp = blockIdx.x;
  const int qn = Q < 256 ? Q : 256;
  const int ns = 256 / qn;
  const int ql = threadIdx.x % qn, sl = threadIdx.x / qn;
  float csum = 0.f;
  for (int q0 = 0; q0 < Q; q0 += qn) {
    const int q = q0 + ql;
    float acc = 0.f;
    if (sl < ns && q < Q)
      for (int n = sl; n < N; n += ns) acc = fmaf(A[static_cast<size_t>(n) * P + pp], B[static_cast<size_t>(n) * Q + q], acc);
    red[threadIdx.x] = acc;
    __syncthreads();
    if (sl == 0 && q < Q) {
      for (int k = 1; k < ns; ++k) acc += red[k * qn + ql];
      out[static_cast<size_t>(pp) * Q + q] += acc;
    }
    __syncthreads();
  }
  if (colsum) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) csum += A[static_cast<size_t>(n) * P + pp];
    cs[threadIdx.x] = csum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < 256; ++k) t += cs[k];
      colsum[pp] += t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bn2_bwd_coef (grid = C, one block per channel; threads over n):
//   dz = s*dpre + dp/M ;  sum dz = sum_n (s*S1 + dp) ; sum dz*u^ = rstd * sum_n ( s*(S2 - mu*S1) + dp/M*(pool - M*mu) )
//   dU = k1[n,c]*dpre + k2[c]*U + k3[n,c]
//   residual BN:  dR = r1*dpre + r2*R + r3  from  sum dpre (=sum_n S1)  and  sum dpre*R (= sum_n S3)
// ---------------------------------------------------------------------------------------------
__global__ void bn2_bwd_coef_kernel(const float* __restrict__ S1, const float* __restrict__ S2,
                                    const float* __restrict__ S3, const float* __restrict__ pool,
                                    const float* __restrict__ dp, const float* __restrict__ s,
                                    const float* __restrict__ a2, const float* __restrict__ mean2,
                                    const float* __restrict__ rstd2, const float* __restrict__ ar,
                                    const float* __restrict__ meanr, const float* __restrict__ rstdr, float M,
                                    double count, int training, float* __restrict__ k1, float* __restrict__ k2,
                                    float* __restrict__ k3, float* __restrict__ r1, float* __restrict__ r2,
                                    float* __restrict__ r3, float* __restrict__ dgamma2, float* __restrict__ dbeta2,
                                    float* __restrict__ dgammar, float* __restrict__ dbetar, int N, int C) {
  __shared__ double sh[4][32];
  __shared__ double tot[4];
  const int c = blockIdx.x;
  const double mu = mean2[c], rs = rstd2[c];
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = static_cast<size_t>(n) * C + c;
    const double sv = s[i], d = dp[i], x1 = S1[i], x2 = S2[i];
    v0 += sv * x1 + d;
    v1 += sv * (x2 - mu * x1) + d / M * (pool[i] - M * mu);
    v2 += x1;
    if (S3) v3 += S3[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    v3 += __shfl_xor_sync(0xffffffffu, v3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = v0;
    sh[1][threadIdx.x >> 5] = v1;
    sh[2][threadIdx.x >> 5] = v2;
    sh[3][threadIdx.x >> 5] = v3;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  const double sum_dz = tot[0];
  const double sum_dzu = tot[1] * rs;  // sum dz * u^
  const double A2 = a2[c];
  const double mdz = training ? sum_dz / count : 0.0;
  const double mdzu = training ? sum_dzu / count : 0.0;
  if (threadIdx.x == 0) {
    dbeta2[c] += static_cast<float>(sum_dz);
    dgamma2[c] += static_cast<float>(sum_dzu);
    k2[c] = static_cast<float>(-A2 * rs * mdzu);
    if (S3) {
      const double mur = meanr[c], rsr = rstdr[c], Ar = ar[c];
      const double sum_d = tot[2];
      const double sum_dr = (tot[3] - mur * tot[2]) * rsr;  // sum dpre * r^
      dbetar[c] += static_cast<float>(sum_d);
      dgammar[c] += static_cast<float>(sum_dr);
      const double md = training ? sum_d / count : 0.0, mdr = training ? sum_dr / count : 0.0;
      r1[c] = static_cast<float>(Ar);
      r2[c] = static_cast<float>(-Ar * rsr * mdr);
      r3[c] = static_cast<float>(Ar * (-md + mur * rsr * mdr));
    }
  }
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = static_cast<size_t>(n) * C + c;
    k1[i] = static_cast<float>(A2 * s[i]);
    k3[i] = static_cast<float>(A2 * (dp[i] / M - mdz + mu * rs * mdzu));
  }
}

// bn1_bwd_coef: per channel, from T1 = sum dy1, T2 = sum dy1*G:
//   dG = c1*dy1 + c2*G + c3 ; dgamma1 = rstd*(T2 - mu*T1) ; dbeta1 = T1
__global__ void bn1_bwd_coef_kernel(const double* __restrict__ T1r, const double* __restrict__ T2r, int nrep,
                                    const float* __restrict__ a1, const float* __restrict__ mean1,
                                    const float* __restrict__ rstd1, double count, int training,
                                    float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ c3,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = mean1[c], rs = rstd1[c], A1 = a1[c];
  double t1 = 0, t2raw = 0;
  for (int r = 0; r < nrep; ++r) {
    t1 += T1r[static_cast<size_t>(r) * C + c];
    t2raw += T2r[static_cast<size_t>(r) * C + c];
  }
  const double t2 = (t2raw - mu * t1) * rs;
  dbeta[c] += static_cast<float>(t1);
  dgamma[c] += static_cast<float>(t2);
  const double m1 = training ? t1 / count : 0.0, m2 = training ? t2 / count : 0.0;
  c1[c] = static_cast<float>(A1);
  c2[c] = static_cast<float>(-A1 * rs * m2);
  c3[c] = static_cast<float>(A1 * (-m1 + mu * rs * m2));
}

}  // namespace fmm

using namespace fmm;

extern "C" {

int fmm_bn_finalize(const double* ch_sum, const double* ch_sq, int nrep, double count, const float* gamma, const float* beta,
                    float* rmean, float* rvar, float momentum, float eps, int training, float* a, float* b,
                    float* mean_out, float* rstd_out, int C, cudaStream_t stream) {
  FMM_CHECK_ARG(a && b && C > 0 && (training ? (ch_sum && ch_sq) : (rmean && rvar)), "bn_finalize: bad args");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(ch_sum, ch_sq, nrep, count, gamma, beta, rmean, rvar, momentum, eps,
                                                          training, a, b, mean_out, rstd_out, C);
  FMM_CHECK_LAUNCH("bn_finalize");
  return FMM_OK;
}

int fmm_se_fwd(const float* pool, const float* a2, const float* b2, float invM, const float* W1, const float* b1,
               const float* gamma, const float* beta, float* rmean, float* rvar, float momentum, float eps,
               int training, const float* W2, const float* b2se, float* p, float* h, float* ah, float* bh,
               float* hmean, float* hrstd, float* s, float* k1, float* k0, int N, int C, int C4,
               cudaStream_t stream) {
  FMM_CHECK_ARG(pool && a2 && b2 && W1 && b1 && gamma && beta && rmean && rvar && W2 && b2se && p && h && ah && bh &&
                    hmean && hrstd && s && k1 && k0 && N > 0 && C > 0 && C4 > 0,
                "se_fwd: bad args");
  se_fwd_a_kernel<<<N, 256, C * sizeof(float), stream>>>(pool, a2, b2, invM, W1, b1, p, h, C, C4);
  se_bn_kernel<<<C4, 128, 0, stream>>>(h, gamma, beta, rmean, rvar, momentum, eps, training, ah, bh, hmean, hrstd, N, C4);
  se_fwd_b_kernel<<<N, 256, C4 * sizeof(float), stream>>>(h, ah, bh, W2, b2se, a2, b2, s, k1, k0, C, C4);
  FMM_CHECK_LAUNCH("se_fwd");
  return FMM_OK;
}

// SE backward: fills dq, dhr, r, dh, dp (workspaces, [N][C] / [N][C4]) and ACCUMULATES the parameter
// gradients dW1[C4][C], db1[C4], dgamma[C4], dbeta[C4], dW2[C][C4], db2se[C].
int fmm_se_bwd(const float* S1, const float* S2, const float* a2, const float* b2, const float* s, const float* p,
               const float* h, const float* ah, const float* bh, const float* hmean, const float* hrstd,
               const float* W1, const float* W2, int training, float* dq, float* dhr, float* r, float* dh, float* dp,
               float* dW1, float* db1, float* dgamma, float* dbeta, float* dW2, float* db2se, int N, int C, int C4,
               cudaStream_t stream) {
  FMM_CHECK_ARG(S1 && S2 && a2 && b2 && s && p && h && ah && bh && hmean && hrstd && W1 && W2 && dq && dhr && r && dh &&
                    dp && dgamma && dbeta,
                "se_bwd: bad args");
  FMM_CHECK_ARG((dW1 && db1 && dW2 && db2se) || (!dW1 && !db1 && !dW2 && !db2se), "se_bwd: dW1/db1/dW2/db2se all or none");
  se_bwd_a_kernel<<<N, 256, (C + 256) * sizeof(float), stream>>>(S1, S2, a2, b2, s, h, ah, bh, W2, dq, dhr, r, C, C4);
  se_bwd_bn_kernel<<<C4, 128, 0, stream>>>(dhr, h, ah, hmean, hrstd, training, dh, dgamma, dbeta, N, C4);
  se_bwd_dp_kernel<<<N, 256, C4 * sizeof(float), stream>>>(dh, W1, dp, C, C4);
  if (dW2) {
    small_tn_gemm_kernel<<<C, 256, 0, stream>>>(dq, r, dW2, db2se, N, C, C4);   // dW2[c][j] = sum_n dq[n,c] r[n,j]
    small_tn_gemm_kernel<<<C4, 256, 0, stream>>>(dh, p, dW1, db1, N, C4, C);   // dW1[j][c] = sum_n dh[n,j] p[n,c]
  }
  FMM_CHECK_LAUNCH("se_bwd");
  return FMM_OK;
}

// The two weight-gradient products of fmm_se_bwd on their own (fmm_se_bwd called with dW1 = db1 = dW2 = db2se = NULL):
// nothing on the activation-gradient chain depends on them, so the caller can run them on a side stream.
int fmm_se_bwd_params(const float* dq, const float* r, const float* dh, const float* p, float* dW1, float* db1, float* dW2,
                      float* db2se, int N, int C, int C4, cudaStream_t stream) {
  FMM_CHECK_ARG(dq && r && dh && p && dW1 && db1 && dW2 && db2se && N > 0 && C > 0 && C4 > 0, "se_bwd_params: bad args");
  small_tn_gemm_kernel<<<C, 256, 0, stream>>>(dq, r, dW2, db2se, N, C, C4);
  small_tn_gemm_kernel<<<C4, 256, 0, stream>>>(dh, p, dW1, db1, N, C4, C);
  FMM_CHECK_LAUNCH("se_bwd_params");
  return FMM_OK;
}

int fmm_bn2_bwd_coef(const float* S1, const float* S2, const float* S3, const float* pool, const float* dp,
                     const float* s, const float* a2, const float* mean2, const float* rstd2, const float* ar,
                     const float* meanr, const float* rstdr, float M, double count, int training, float* k1, float* k2,
                     float* k3, float* r1, float* r2, float* r3, float* dgamma2, float* dbeta2, float* dgammar,
                     float* dbetar, int N, int C, cudaStream_t stream) {
  FMM_CHECK_ARG(S1 && S2 && pool && dp && s && a2 && mean2 && rstd2 && k1 && k2 && k3 && dgamma2 && dbeta2,
                "bn2_bwd_coef: bad args");
  FMM_CHECK_ARG(!S3 || (ar && meanr && rstdr && r1 && r2 && r3 && dgammar && dbetar), "bn2_bwd_coef: residual args");
  bn2_bwd_coef_kernel<<<C, 128, 0, stream>>>(S1, S2, S3, pool, dp, s, a2, mean2, rstd2, ar, meanr, rstdr, M, count,
                                             training, k1, k2, k3, r1, r2, r3, dgamma2, dbeta2, dgammar, dbetar, N, C);
  FMM_CHECK_LAUNCH("bn2_bwd_coef");
  return FMM_OK;
}

int fmm_bn1_bwd_coef(const double* T1, const double* T2, int nrep, const float* a1, const float* mean1, const float* rstd1,
                     double count, int training, float* c1, float* c2, float* c3, float* dgamma, float* dbeta, int C,
                     cudaStream_t stream) {
  FMM_CHECK_ARG(T1 && T2 && a1 && mean1 && rstd1 && c1 && c2 && c3 && dgamma && dbeta, "bn1_bwd_coef: bad args");
  bn1_bwd_coef_kernel<<<(C + 127) / 128, 128, 0, stream>>>(T1, T2, nrep, a1, mean1, rstd1, count, training, c1, c2, c3, dgamma,
                                                           dbeta, C);
  FMM_CHECK_LAUNCH("bn1_bwd_coef");
  return FMM_OK;
}

}  // extern "C"
