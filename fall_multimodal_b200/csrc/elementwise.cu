// Memory-bound kernels of the ST-GCN block (everything that touches a full activation but is not
// a GEMM): adjacency aggregation, BatchNorm/SE statistics, the fused block output, and the
// elementwise halves of the BatchNorm / ReLU / SE / residual backward.
//
// Layout: activations are channels-last [N][T][V][C]; a frame (one n,t) is V*C contiguous
// elements. All kernels use the same mapping: grid = (time chunks, N); inside a block a thread
// owns a fixed (v, 8-channel group) "pair" and walks over the frames of its chunk, so global
// accesses are perfectly coalesced 16/32-byte vectors and per-channel coefficients stay in
// registers. Reductions over frames are kept in registers, reduced across the block in shared
// memory and committed with one atomic per (block, output element).
//
// Reference call sites: /root/reference/Fall_2_Spatial_Temporal_SR/Model/stgcan.py
//   :54 (einsum nkctv,kvw->nctw, reassociated onto the input), :63-73 (SE pooling and scale),
//   :112-119 (BatchNorm2d/ReLU around the temporal conv), :138-144 (attention, residual, ReLU).
#include <stdlib.h>

#include "common.cuh"
#include "stream.cuh"

#ifndef FMM_EW_MINB
#define FMM_EW_MINB 3   // resident 256-thread blocks per SM the streaming kernels are compiled for (<= 85 registers)
#endif

namespace fmm {

constexpr int kEwThreads = 256;

__device__ __forceinline__ void atomic_add_f64(double* p, double v) { atomicAdd(p, v); }
// Cross-block accumulators ([C] doubles, [V][C] floats) are replicated `nrep` times ([nrep][...]);
// a block adds into replica (linear block id % nrep) and the consumer sums the replicas. Thousands
// of blocks hitting the same few hundred addresses otherwise serialise in the L2 atomic units.
__device__ __forceinline__ int replica_of_block(int nrep) { return (blockIdx.x + blockIdx.y * gridDim.x) % nrep; }

// ------------------------------------------------------------------------------------------
// agg_fwd:  Xa[(n,t,w), k*Cin+ci] = sum_{e in in(k,w)} coef[e] * x[(n,t,src[e]), ci]
// CSR over (k,w): rowptr[K*V+1], src[E], coef[E].
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void agg_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ xa, const int* __restrict__ rowptr,
                                   const int* __restrict__ src, const float* __restrict__ coef, int Tn, int V,
                                   int Cin, int K, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk;
  const int t1 = min(t0 + tchunk, Tn);
  const int c8n = Cin / 8;
  const int pairs = V * K * c8n;
  for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
    const int c8 = i % c8n;
    const int k = (i / c8n) % K;
    const int w = i / (c8n * K);
    const int e0 = rowptr[k * V + w], e1 = rowptr[k * V + w + 1];
    for (int t = t0; t < t1; ++t) {
      const size_t frame = static_cast<size_t>(n) * Tn + t;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int e = e0; e < e1; ++e) {
        float f[8];
        load8(x + (frame * V + src[e]) * Cin + c8 * 8, f);
        const float cf = coef[e];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(cf, f[j], acc[j]);
      }
      store8(xa + (frame * V + w) * (static_cast<size_t>(K) * Cin) + k * Cin + c8 * 8, acc);
    }
  }
}

template <typename T>
__global__ void agg_fwd_scalar_kernel(const T* __restrict__ x, T* __restrict__ xa, const int* __restrict__ rowptr,
                                      const int* __restrict__ src, const float* __restrict__ coef, int Tn, int V,
                                      int Cin, int K, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk;
  const int t1 = min(t0 + tchunk, Tn);
  const int per_frame = V * K * Cin;
  const int total = (t1 - t0) * per_frame;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int t = t0 + i / per_frame;
    const int r = i % per_frame;
    const int ci = r % Cin;
    const int k = (r / Cin) % K;
    const int w = r / (Cin * K);
    const size_t frame = static_cast<size_t>(n) * Tn + t;
    float acc = 0.f;
    for (int e = rowptr[k * V + w]; e < rowptr[k * V + w + 1]; ++e)
      acc = fmaf(coef[e], to_f32(x[(frame * V + src[e]) * Cin + ci]), acc);
    xa[(frame * V + w) * (static_cast<size_t>(K) * Cin) + k * Cin + ci] = from_f32<T>(acc);
  }
}

// ------------------------------------------------------------------------------------------
// agg_bwd: dx[(n,t,v), ci] = addend + sum_{e in out(v)} coef[e] * P[(n,t,dst[e]), kk[e]*Cin+ci]
// CSR over v: rowptr[V+1], dst[E], kk[E], coef[E].
// ------------------------------------------------------------------------------------------
// With `x` given, the same pass also accumulates the edge-coefficient gradient
//   dcoef[eid[e]] += sum_{n,t,ci} x[(n,t,v),ci] * P[(n,t,dst[e]), kk[e]*Cin+ci]
// (every P piece is read exactly once by the thread that needs it for dx, so this costs no traffic).
constexpr int kMaxOutEdges = 8;
template <typename T>
__global__ void agg_bwd_vec_kernel(const T* __restrict__ P, const T* __restrict__ addend, T* __restrict__ dx,
                                   const int* __restrict__ rowptr, const int* __restrict__ dst,
                                   const int* __restrict__ kk, const float* __restrict__ coef,
                                   const T* __restrict__ x, const int* __restrict__ eid, float* __restrict__ dcoef,
                                   int Tn, int V, int Cin, int K, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk;
  const int t1 = min(t0 + tchunk, Tn);
  const int c8n = Cin / 8;
  const int pairs = V * c8n;
  const size_t prow = static_cast<size_t>(K) * Cin;
  const bool pow2 = (c8n & (c8n - 1)) == 0 && c8n <= 32;
  for (int i0 = 0; i0 < pairs; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool active = i < pairs;
    const int c8 = active ? i % c8n : 0;
    const int v = active ? i / c8n : 0;
    const int e0 = active ? rowptr[v] : 0, e1 = active ? rowptr[v + 1] : 0;
    float dc[kMaxOutEdges];
#pragma unroll
    for (int j = 0; j < kMaxOutEdges; ++j) dc[j] = 0.f;
    for (int t = t0; t < t1 && active; ++t) {
      const size_t frame = static_cast<size_t>(n) * Tn + t;
      float acc[8], xv[8];
      if (addend) {
        load8(addend + (frame * V + v) * Cin + c8 * 8, acc);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      }
      if (x) load8(x + (frame * V + v) * Cin + c8 * 8, xv);
#pragma unroll
      for (int j = 0; j < kMaxOutEdges; ++j) {
        const int e = e0 + j;
        if (e < e1) {
          float f[8];
          load8(P + (frame * V + dst[e]) * prow + kk[e] * Cin + c8 * 8, f);
          const float cf = coef[e];
          float d = 0.f;
#pragma unroll
          for (int l = 0; l < 8; ++l) {
            acc[l] = fmaf(cf, f[l], acc[l]);
            if (x) d = fmaf(xv[l], f[l], d);
          }
          dc[j] += d;
        }
      }
      for (int e = e0 + kMaxOutEdges; e < e1; ++e) {  // rare: joints with more than 8 out-edges
        float f[8];
        load8(P + (frame * V + dst[e]) * prow + kk[e] * Cin + c8 * 8, f);
        const float cf = coef[e];
        float d = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          acc[l] = fmaf(cf, f[l], acc[l]);
          if (x) d = fmaf(xv[l], f[l], d);
        }
        if (x) atomicAdd(dcoef + eid[e], d);
      }
      store8(dx + (frame * V + v) * Cin + c8 * 8, acc);
    }
    if (x) {
      // reduce over the channel groups of one joint (consecutive lanes), one atomic per (block, edge)
#pragma unroll
      for (int j = 0; j < kMaxOutEdges; ++j) {
        float d = dc[j];
        if (pow2) {
          for (int off = c8n >> 1; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
          if (active && c8 == 0 && e0 + j < e1) atomicAdd(dcoef + eid[e0 + j], d);
        } else if (active && e0 + j < e1) {
          atomicAdd(dcoef + eid[e0 + j], d);
        }
      }
    }
  }
}

template <typename T>
__global__ void agg_bwd_scalar_kernel(const T* __restrict__ P, const T* __restrict__ addend, T* __restrict__ dx,
                                      const int* __restrict__ rowptr, const int* __restrict__ dst,
                                      const int* __restrict__ kk, const float* __restrict__ coef, int Tn, int V,
                                      int Cin, int K, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk;
  const int t1 = min(t0 + tchunk, Tn);
  const int per_frame = V * Cin;
  const int total = (t1 - t0) * per_frame;
  const size_t prow = static_cast<size_t>(K) * Cin;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int t = t0 + i / per_frame;
    const int r = i % per_frame;
    const int ci = r % Cin;
    const int v = r / Cin;
    const size_t frame = static_cast<size_t>(n) * Tn + t;
    float acc = addend ? to_f32(addend[(frame * V + v) * Cin + ci]) : 0.f;
    for (int e = rowptr[v]; e < rowptr[v + 1]; ++e)
      acc = fmaf(coef[e], to_f32(P[(frame * V + dst[e]) * prow + kk[e] * Cin + ci]), acc);
    dx[(frame * V + v) * Cin + ci] = from_f32<T>(acc);
  }
}

// ------------------------------------------------------------------------------------------
// agg_dcoef: dcoef[e] += sum_{n,t,ci} x[(n,t,src[e]),ci] * P[(n,t,dst[e]), kk[e]*Cin+ci]
// (edge lists in any order). One warp per (edge, frame range); lanes stride the channels.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void agg_dcoef_kernel(const T* __restrict__ x, const T* __restrict__ P, float* __restrict__ dcoef,
                                 const int* __restrict__ src, const int* __restrict__ dst,
                                 const int* __restrict__ kk, int E, int Tn, int V, int Cin, int K, int tchunk) {
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tchunk;
  const int t1 = min(t0 + tchunk, Tn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const size_t prow = static_cast<size_t>(K) * Cin;
  for (int e = warp; e < E; e += nwarps) {
    const int v = src[e], w = dst[e], k = kk[e];
    float acc = 0.f;
    for (int t = t0; t < t1; ++t) {
      const size_t frame = static_cast<size_t>(n) * Tn + t;
      const T* xr = x + (frame * V + v) * Cin;
      const T* pr = P + (frame * V + w) * prow + static_cast<size_t>(k) * Cin;
      for (int ci = lane; ci < Cin; ci += 32) acc = fmaf(to_f32(xr[ci]), to_f32(pr[ci]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(dcoef + e, acc);
  }
}

// Narrow inputs (the first block: 3 or 2 channels): lanes stride the FRAMES of a clip instead of the channels, one block per
// clip, warps stride the edges (with lanes on channels 29 of 32 lanes idled: 188 us per launch at the bench shape).
template <typename T>
__global__ void agg_dcoef_narrow_kernel(const T* __restrict__ x, const T* __restrict__ P, float* __restrict__ dcoef,
                                        const int* __restrict__ src, const int* __restrict__ dst,
                                        const int* __restrict__ kk, int E, int Tn, int V, int Cin, int K) {
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const size_t prow = static_cast<size_t>(K) * Cin;
  for (int e = warp; e < E; e += nwarps) {
    const int v = src[e], w = dst[e], k = kk[e];
    float acc = 0.f;
    for (int t = lane; t < Tn; t += 32) {
      const size_t frame = static_cast<size_t>(n) * Tn + t;
      const T* xr = x + (frame * V + v) * Cin;
      const T* pr = P + (frame * V + w) * prow + static_cast<size_t>(k) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(to_f32(xr[ci]), to_f32(pr[ci]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(dcoef + e, acc);
  }
}

// ------------------------------------------------------------------------------------------
// Row-walk mapping for everything that needs no per-joint output: a block owns the contiguous rows
// [t0*V, t1*V) of clip n; thread = (8-channel group c8 = tid % c8n, row lane = tid / c8n) walks
// the rows with stride blockDim/c8n. Perfectly balanced for any V, coalesced, per-channel
// coefficients in registers. Per-channel partial sums are reduced with warp shuffles over the lanes
// that share c8, then shared-memory atomics across warps, then ONE global atomic per channel.
// ------------------------------------------------------------------------------------------
// Every thread holds NV x 8 partial sums for channel group c8 (row lane rl). They are staged as
// scratch[rl][c8n*8][NV] with plain conflict-free stores and summed over rl by C*NV threads:
// shared-memory atomics cost ~2 cycles per lane and serialise the whole SM at the end of each block.
// The staging area is laid out [j*NV+v][rl][c8] (the thread index fastest: consecutive lanes hit consecutive banks; the
// [rl][c][v] layout it replaces cost a 15-way bank conflict per store, ncu r02).  red[c*NV+v] as before.
template <int NV>
__device__ __forceinline__ void block_channel_reduce(const float (&val)[NV][8], float* scratch, float* red, int c8,
                                                     int c8n, int rl, int RL) {
  const int C = c8n * 8;
  const int nt = RL * c8n;   // threads that hold partial sums (= blockDim.x for the row-walk launches)
  __syncthreads();           // the previous segment's readers are done with scratch / red
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int v = 0; v < NV; ++v) scratch[static_cast<size_t>(j * NV + v) * nt + rl * c8n + c8] = val[v][j];
  __syncthreads();
  for (int i = threadIdx.x; i < C * NV; i += blockDim.x) {
    const int c = i / NV, v = i - c * NV;
    const float* col = scratch + static_cast<size_t>((c & 7) * NV + v) * nt + (c >> 3);
    float acc = 0.f;
    for (int r = 0; r < RL; ++r) acc += col[r * c8n];
    red[i] = acc;
  }
}

// Balanced persistent partition for the streaming kernels: the grid is ONE resident wave (SMs x blocks per SM) and block b
// owns the rows [R*b/G, R*(b+1)/G) of the flattened (clip, frame*joint) row space - every block the same number of rows
// (grids of N x T/chunk blocks ran 1.4-3.5 waves: up to half of the kernel was the partial last wave, ncu r02). A block's
// range is walked clip by clip (per-clip coefficients / per-clip sums): `next` yields (n, r0, r1) with rows local to clip n.
struct RowSegs {
  long long cur, end;
  int rows_per_n;
  __device__ RowSegs(int N, int rows_per_n_) : rows_per_n(rows_per_n_) {
    const long long total = static_cast<long long>(N) * rows_per_n_;
    cur = total * blockIdx.x / gridDim.x;
    end = total * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ bool next(int& n, int& r0, int& r1) {
    if (cur >= end) return false;
    n = static_cast<int>(cur / rows_per_n);
    const long long base = static_cast<long long>(n) * rows_per_n;
    const long long e = end < base + rows_per_n ? end : base + rows_per_n;
    r0 = static_cast<int>(cur - base);
    r1 = static_cast<int>(e - base);
    cur = e;
    return true;
  }
};

// colstats: per-channel sum / sum of squares (double) and optional per-(n,c) sums (float) of X.
template <typename T>
__global__ void __launch_bounds__(256, FMM_EW_MINB) colstats_kernel(const T* __restrict__ X, double* __restrict__ ch_sum, double* __restrict__ ch_sq,
                                float* __restrict__ nc_sum, int N, int Tn, int V, int C, int nrep) {
  extern __shared__ float red[];  // [C][2] followed by the [RL][C][2] staging area
  float* scratch = red + 2 * C;
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  RowSegs segs(N, Tn * V);
  int n, r0, r1;
  while (segs.next(n, r0, r1)) {
  const T* base = X + static_cast<size_t>(n) * Tn * V * C + c8 * 8;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  constexpr int U = 8;   // rows in flight per thread
  for (int r = r0 + rl; r < r1; r += U * RL) {
    Raw8<T> raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (r + u * RL < r1) ldraw8(base + static_cast<size_t>(r + u * RL) * C, raw[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * RL < r1) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += f[j];
          acc[1][j] = fmaf(f[j], f[j], acc[1][j]);
        }
      }
    }
  }
  block_channel_reduce<2>(acc, scratch, red, c8, c8n, rl, RL);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s = red[2 * c], q = red[2 * c + 1];
    const size_t ro = static_cast<size_t>(replica_of_block(nrep)) * C;
    if (ch_sum) atomic_add_f64(ch_sum + ro + c, static_cast<double>(s));
    if (ch_sq) atomic_add_f64(ch_sq + ro + c, static_cast<double>(q));
    if (nc_sum) atomicAdd(nc_sum + static_cast<size_t>(n) * C + c, s);
  }
  }
}

// block_out: y = relu(k1[n,c]*U + k0[n,c] + res), res = 0 | X | ar[c]*R + br[c]
template <typename T>
__global__ void block_out_kernel(const T* __restrict__ U, const float* __restrict__ k1, const float* __restrict__ k0,
                                 const T* __restrict__ res, const float* __restrict__ ar,
                                 const float* __restrict__ br, T* __restrict__ Y, int N, int Tn, int V, int C) {
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  const int c0 = c8 * 8;
  RowSegs segs(N, Tn * V);
  int n, r0, r1;
  while (segs.next(n, r0, r1)) {
  float a[8], b[8], ra[8], rb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = k1[static_cast<size_t>(n) * C + c0 + j];
    b[j] = k0[static_cast<size_t>(n) * C + c0 + j];
    ra[j] = ar ? ar[c0 + j] : 1.f;
    rb[j] = br ? br[c0 + j] : 0.f;
  }
  const size_t base = static_cast<size_t>(n) * Tn * V * C + c0;
  constexpr int NB = 4;   // rows in flight per thread
  for (int r = r0 + rl; r < r1; r += NB * RL) {
    Raw8<T> ru[NB], rres[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        const size_t off = base + static_cast<size_t>(r + q * RL) * C;
        ldraw8(U + off, ru[q]);
        if (res) ldraw8(res + off, rres[q]);
      }
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        const size_t off = base + static_cast<size_t>(r + q * RL) * C;
        float u[8], rr[8], y[8];
        unpack8(ru[q], u);
        if (res) unpack8(rres[q], rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = fmaf(a[j], u[j], b[j]);
          if (res) v += fmaf(ra[j], rr[j], rb[j]);
          y[j] = fmaxf(v, 0.f);
        }
        store8(Y + off, y);
      }
  }
  }
}

// affine_relu: H = relu(a[c]*X + b[c])  (BatchNorm + ReLU of tcn[0..1], stgcan.py:112-113, materialised once per
// block so that the temporal conv and its weight gradient stream H with plain async copies)
template <typename T>
__global__ void affine_relu_kernel(const T* __restrict__ X, const float* __restrict__ a, const float* __restrict__ b,
                                   T* __restrict__ H, int N, int Tn, int V, int C) {
  const long long total = static_cast<long long>(N) * Tn * V;
  const long long r0 = total * blockIdx.x / gridDim.x, r1 = total * (blockIdx.x + 1) / gridDim.x;
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sa[j] = a[c8 * 8 + j];
    sb[j] = b[c8 * 8 + j];
  }
  const size_t base = static_cast<size_t>(c8) * 8;
  constexpr int NB = 8;   // rows in flight per thread
  for (long long r = r0 + rl; r < r1; r += NB * RL) {
    Raw8<T> raw[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) ldraw8(X + base + static_cast<size_t>(r + q * RL) * C, raw[q]);
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        float f[8];
        unpack8(raw[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(sa[j], f[j], sb[j]), 0.f);
        store8(H + base + static_cast<size_t>(r + q * RL) * C, f);
      }
  }
}

// blockout_bwd_reduce: dpre = dY*(Y>0); S1[n,c]=sum dpre; S2[n,c]=sum dpre*U; S3[n,c]=sum dpre*R
template <typename T, bool kRes, bool kMask>   // residual-conv branch present / dY still needs the ReLU mask
__global__ void __launch_bounds__(256, FMM_EW_MINB) blockout_bwd_reduce_kernel(const T* __restrict__ dY, const T* __restrict__ Y, const T* __restrict__ U,
                                           const T* __restrict__ R, float* __restrict__ S1, float* __restrict__ S2,
                                           float* __restrict__ S3, int N, int Tn, int V, int C) {
  extern __shared__ float red[];  // [C][3] followed by the [RL][C][3] staging area
  float* scratch = red + 3 * C;
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  RowSegs segs(N, Tn * V);
  int n, r0, r1;
  while (segs.next(n, r0, r1)) {
  const size_t base = static_cast<size_t>(n) * Tn * V * C + c8 * 8;
  float acc[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
  constexpr int NB = (kRes || kMask) ? 2 : 4;   // rows in flight per thread (2-4 tensors each)
  for (int r = r0 + rl; r < r1; r += NB * RL) {
    Raw8<T> rg[NB], ry[kMask ? NB : 1], ru[NB], rres[kRes ? NB : 1];
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        const size_t off = base + static_cast<size_t>(r + q * RL) * C;
        ldraw8(dY + off, rg[q]);
        if (kMask) ldraw8(Y + off, ry[q]);
        ldraw8(U + off, ru[q]);
        if (kRes) ldraw8(R + off, rres[q]);
      }
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        float g[8], y[8], u[8], rr[8];
        unpack8(rg[q], g);
        if (kMask) unpack8(ry[q], y);
        unpack8(ru[q], u);
        if (kRes) unpack8(rres[q], rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = (!kMask || y[j] > 0.f) ? g[j] : 0.f;   // Y == null: dY arrives already masked (fmm_gcn_bwd relu_mask)
          acc[0][j] += d;
          acc[1][j] = fmaf(d, u[j], acc[1][j]);
          if (kRes) acc[2][j] = fmaf(d, rr[j], acc[2][j]);
        }
      }
  }
  block_channel_reduce<3>(acc, scratch, red, c8, c8n, rl, RL);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(S1 + static_cast<size_t>(n) * C + c, red[3 * c]);
    atomicAdd(S2 + static_cast<size_t>(n) * C + c, red[3 * c + 1]);
    if (kRes) atomicAdd(S3 + static_cast<size_t>(n) * C + c, red[3 * c + 2]);
  }
  }
}

// bn2_bwd_apply: dpre = dY*(Y>0)
//   dU = k1[n,c]*dpre + k2[c]*U + k3[n,c]        (always)
//   dR = r1[c]*dpre + r2[c]*R + r3[c]            (if R)
//   dPre = dpre                                   (if dPre: identity residual)
//   colsum_dU[c] += sum dU (as stored), colsum_dR[c] += sum dR   (optional, double)
// kRes / kMask: residual-conv branch present / dY still needs the ReLU mask (compile-time: the unused coefficient and raw
// registers of the common pre-masked identity block are what kept this kernel at one block per SM)
template <typename T, bool kRes, bool kMask>
__global__ void __launch_bounds__(256, kRes ? 2 : 3) bn2_bwd_apply_kernel(const T* __restrict__ dY, const T* __restrict__ Y, const T* __restrict__ U,
                                     const T* __restrict__ R, const float* __restrict__ k1,
                                     const float* __restrict__ k2, const float* __restrict__ k3,
                                     const float* __restrict__ r1, const float* __restrict__ r2,
                                     const float* __restrict__ r3, T* __restrict__ dU, T* __restrict__ dR,
                                     T* __restrict__ dPre, double* __restrict__ sum_dU,
                                     double* __restrict__ sum_dR, int N, int Tn, int V, int C, int nrep) {
  extern __shared__ float red[];  // [C][2] followed by the [RL][C][2] staging area
  float* scratch = red + 2 * C;
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  const int c0 = c8 * 8;
  float acc[2][8];   // per-channel sums of the stored dU / dR over ALL of this block's rows: one flush at the end
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  RowSegs segs(N, Tn * V);
  int n, r0, r1r;
  while (segs.next(n, r0, r1r)) {
  float a1[8], a2[8], a3[8], b1[kRes ? 8 : 1], b2[kRes ? 8 : 1], b3[kRes ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = k1[static_cast<size_t>(n) * C + c0 + j];
    a2[j] = k2[c0 + j];
    a3[j] = k3[static_cast<size_t>(n) * C + c0 + j];
    if (kRes) {
      b1[j] = r1[c0 + j];
      b2[j] = r2[c0 + j];
      b3[j] = r3[c0 + j];
    }
  }
  const size_t base = static_cast<size_t>(n) * Tn * V * C + c0;
  constexpr int NB = 2;   // rows in flight per thread (3-4 tensors each)
  for (int rb = r0 + rl; rb < r1r; rb += NB * RL) {
    Raw8<T> rg[NB], ry[kMask ? NB : 1], ru[NB], rres[kRes ? NB : 1];
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (rb + q * RL < r1r) {
        const size_t off = base + static_cast<size_t>(rb + q * RL) * C;
        ldraw8(dY + off, rg[q]);
        if (kMask) ldraw8(Y + off, ry[q]);
        ldraw8(U + off, ru[q]);
        if (kRes) ldraw8(R + off, rres[q]);
      }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
    if (rb + q * RL >= r1r) continue;
    const size_t off = base + static_cast<size_t>(rb + q * RL) * C;
    float g[8], y[8], u[8], rr[8], ou[8], orr[8], d[8];
    unpack8(rg[q], g);
    if (kMask) unpack8(ry[q], y);
    unpack8(ru[q], u);
    if (kRes) unpack8(rres[q], rr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d[j] = (!kMask || y[j] > 0.f) ? g[j] : 0.f;   // !kMask (Y == null): dY arrives already masked
      ou[j] = fmaf(a1[j], d[j], fmaf(a2[j], u[j], a3[j]));
      acc[0][j] += to_f32(from_f32<T>(ou[j]));
      if (kRes) {
        orr[j] = fmaf(b1[j], d[j], fmaf(b2[j], rr[j], b3[j]));
        acc[1][j] += to_f32(from_f32<T>(orr[j]));
      }
    }
    store8(dU + off, ou);
    if (kRes) store8(dR + off, orr);
    if (dPre) store8(dPre + off, d);
    }
  }
  }
  if (!sum_dU && !sum_dR) return;
  block_channel_reduce<2>(acc, scratch, red, c8, c8n, rl, RL);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t ro = static_cast<size_t>(replica_of_block(nrep)) * C;
    if (sum_dU) atomic_add_f64(sum_dU + ro + c, static_cast<double>(red[2 * c]));
    if (sum_dR) atomic_add_f64(sum_dR + ro + c, static_cast<double>(red[2 * c + 1]));
  }
}

// bn1_bwd_reduce: y1 = a1*G+b1; dy1 = dH*(y1>0); T1[c] += sum dy1; T2[c] += sum dy1*G
template <typename T>
__global__ void bn1_bwd_reduce_kernel(const T* __restrict__ dH, const T* __restrict__ G, const float* __restrict__ a1,
                                      const float* __restrict__ b1, double* __restrict__ T1,
                                      double* __restrict__ T2, int N, int Tn, int V, int C, int nrep) {
  extern __shared__ float red[];  // [C][2] followed by the [RL][C][2] staging area
  float* scratch = red + 2 * C;
  const long long total = static_cast<long long>(N) * Tn * V;
  const long long r0 = total * blockIdx.x / gridDim.x, r1 = total * (blockIdx.x + 1) / gridDim.x;
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = blockDim.x / c8n;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = a1[c8 * 8 + j];
    b[j] = b1[c8 * 8 + j];
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const size_t base = static_cast<size_t>(c8) * 8;
  constexpr int NB = 4;   // rows in flight per thread (2 tensors each)
  for (long long r = r0 + rl; r < r1; r += NB * RL) {
    Raw8<T> rh[NB], rg[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        const size_t off = base + static_cast<size_t>(r + q * RL) * C;
        ldraw8(dH + off, rh[q]);
        ldraw8(G + off, rg[q]);
      }
#pragma unroll
    for (int q = 0; q < NB; ++q)
      if (r + q * RL < r1) {
        float g[8], h[8];
        unpack8(rh[q], h);
        unpack8(rg[q], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = fmaf(a[j], g[j], b[j]) > 0.f ? h[j] : 0.f;
          acc[0][j] += d;
          acc[1][j] = fmaf(d, g[j], acc[1][j]);
        }
      }
  }
  block_channel_reduce<2>(acc, scratch, red, c8, c8n, rl, RL);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t ro = static_cast<size_t>(replica_of_block(nrep)) * C;
    atomic_add_f64(T1 + ro + c, static_cast<double>(red[2 * c]));
    atomic_add_f64(T2 + ro + c, static_cast<double>(red[2 * c + 1]));
  }
}

// The same reduction, inputs staged through the TMA ring of stream.cuh (the launcher's default).
template <typename T>
__global__ void __launch_bounds__(kStThreads, 1)
    bn1_bwd_reduce_tma_kernel(const T* __restrict__ dH, const T* __restrict__ G, const float* __restrict__ a1,
                              const float* __restrict__ b1, double* __restrict__ T1, double* __restrict__ T2, int N, int Tn,
                              int V, int C, int nrep, int chunk_rows, unsigned* err) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  StreamPipe<2> pipe;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(T);
  pipe.init(st_smem, static_cast<uint32_t>(chunk_rows) * row_bytes);
  float* red = reinterpret_cast<float*>(st_smem + (pipe.empty0 + 8u * kStStages + 64u - smem_u32(st_smem)));
  float* scratch = red + 2 * C;
  ChunkIter it(static_cast<long long>(N) * Tn * V, 1, Tn * V, chunk_rows, false);
  long long row0;
  int nrows, n, i = 0;
  if (threadIdx.x >= kStConsumers) {
    if (threadIdx.x == kStConsumers) {
      const void* const src[2] = {dH, G};
      while (it.next(row0, nrows, n)) pipe.produce(i++, src, row0, nrows, row_bytes, err);
    }
    return;
  }
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = kStConsumers / c8n;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = a1[c8 * 8 + j];
    b[j] = b1[c8 * 8 + j];
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const uint32_t toff = static_cast<uint32_t>(threadIdx.x) * 8u * sizeof(T);   // (rl * C + c8 * 8) elements
  const uint32_t rstep = static_cast<uint32_t>(RL) * row_bytes;
  while (it.next(row0, nrows, n)) {
    const int s = pipe.acquire(i++, err);
    uint32_t ah = pipe.tensor(s, 0) + toff, ag = pipe.tensor(s, 1) + toff;
#pragma unroll 2
    for (int r = rl; r < nrows; r += RL, ah += rstep, ag += rstep) {
      float g[8], h[8];
      lds8(ah, h, static_cast<const T*>(nullptr));
      lds8(ag, g, static_cast<const T*>(nullptr));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = fmaf(a[j], g[j], b[j]) > 0.f ? h[j] : 0.f;
        acc[0][j] += d;
        acc[1][j] = fmaf(d, g[j], acc[1][j]);
      }
    }
    pipe.release(s);
  }
  // block reduce among the consumer threads only (the producer warp has left): named barrier 1
  {
    const int nt = RL * c8n;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int v = 0; v < 2; ++v) scratch[static_cast<size_t>(j * 2 + v) * nt + rl * c8n + c8] = acc[v][j];
    asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory");
    for (int k = threadIdx.x; k < C * 2; k += kStConsumers) {
      const int c = k >> 1, v = k & 1;
      const float* col = scratch + static_cast<size_t>((c & 7) * 2 + v) * nt + (c >> 3);
      float sum = 0.f;
      for (int r = 0; r < RL; ++r) sum += col[r * c8n];
      const size_t ro = static_cast<size_t>(replica_of_block(nrep)) * C;
      atomic_add_f64((v ? T2 : T1) + ro + c, static_cast<double>(sum));
    }
  }
}

// Block reduce among the kStConsumers consumer threads of a TMA-staged kernel (the producer warp does not take part: named
// barrier 1): NV x 8 partial sums per thread -> red[c * NV + v].
template <int NV>
__device__ __forceinline__ void consumer_channel_reduce(const float (&val)[NV][8], float* scratch, float* red, int C, int c8,
                                                        int c8n, int rl, int RL) {
  const int nt = RL * c8n;
  asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory");   // readers of the previous flush are done
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int v = 0; v < NV; ++v) scratch[static_cast<size_t>(j * NV + v) * nt + rl * c8n + c8] = val[v][j];
  asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory");
  for (int i = threadIdx.x; i < C * NV; i += kStConsumers) {
    const int c = i / NV, v = i - c * NV;
    const float* col = scratch + static_cast<size_t>((c & 7) * NV + v) * nt + (c >> 3);
    float acc = 0.f;
    for (int r = 0; r < RL; ++r) acc += col[r * c8n];
    red[i] = acc;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory");
}

// blockout_bwd_reduce with its 2-4 inputs staged through the TMA ring (chunks never straddle a clip: sums are per clip)
template <typename T, bool kRes, bool kMask>
__global__ void __launch_bounds__(kStThreads, 1)
    blockout_bwd_reduce_tma_kernel(const T* __restrict__ dY, const T* __restrict__ Y, const T* __restrict__ U, const T* __restrict__ R,
                                   float* __restrict__ S1, float* __restrict__ S2, float* __restrict__ S3, int N, int Tn, int V,
                                   int C, int chunk_rows, unsigned* err) {
  constexpr int NT = 2 + (kRes ? 1 : 0) + (kMask ? 1 : 0);
  constexpr int iU = 1, iY = 2, iR = kMask ? 3 : 2;
  extern __shared__ __align__(128) uint8_t st_smem[];
  StreamPipe<NT> pipe;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(T);
  pipe.init(st_smem, static_cast<uint32_t>(chunk_rows) * row_bytes);
  float* red = reinterpret_cast<float*>(st_smem + (pipe.empty0 + 8u * kStStages + 64u - smem_u32(st_smem)));
  float* scratch = red + 3 * C;
  ChunkIter it(static_cast<long long>(N) * Tn * V, 1, Tn * V, chunk_rows, true);
  long long row0;
  int nrows, n, i = 0;
  if (threadIdx.x >= kStConsumers) {
    if (threadIdx.x == kStConsumers) {
      const void* src[NT];
      src[0] = dY;
      src[iU] = U;
      if (kMask) src[iY] = Y;
      if (kRes) src[iR] = R;
      const void* const(&csrc)[NT] = src;
      while (it.next(row0, nrows, n)) pipe.produce(i++, csrc, row0, nrows, row_bytes, err);
    }
    return;
  }
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = kStConsumers / c8n;
  float acc[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
  const uint32_t toff = static_cast<uint32_t>(threadIdx.x) * 8u * sizeof(T);
  const uint32_t rstep = static_cast<uint32_t>(RL) * row_bytes;
  int cur_n = -1;
  auto flush = [&](int nn) {
    consumer_channel_reduce<3>(acc, scratch, red, C, c8, c8n, rl, RL);
    for (int c = threadIdx.x; c < C; c += kStConsumers) {
      atomicAdd(S1 + static_cast<size_t>(nn) * C + c, red[3 * c]);
      atomicAdd(S2 + static_cast<size_t>(nn) * C + c, red[3 * c + 1]);
      if (kRes) atomicAdd(S3 + static_cast<size_t>(nn) * C + c, red[3 * c + 2]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
  };
  while (it.next(row0, nrows, n)) {
    if (n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
      cur_n = n;
    }
    const int s = pipe.acquire(i++, err);
    uint32_t off = toff;
#pragma unroll 2
    for (int r = rl; r < nrows; r += RL, off += rstep) {
      float g[8], y[8], u[8], rr[8];
      lds8(pipe.tensor(s, 0) + off, g, static_cast<const T*>(nullptr));
      lds8(pipe.tensor(s, iU) + off, u, static_cast<const T*>(nullptr));
      if (kMask) lds8(pipe.tensor(s, iY) + off, y, static_cast<const T*>(nullptr));
      if (kRes) lds8(pipe.tensor(s, iR) + off, rr, static_cast<const T*>(nullptr));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = (!kMask || y[j] > 0.f) ? g[j] : 0.f;
        acc[0][j] += d;
        acc[1][j] = fmaf(d, u[j], acc[1][j]);
        if (kRes) acc[2][j] = fmaf(d, rr[j], acc[2][j]);
      }
    }
    pipe.release(s);
  }
  if (cur_n >= 0) flush(cur_n);
}

// colstats with the input staged through the TMA ring
template <typename T>
__global__ void __launch_bounds__(kStThreads, 1)
    colstats_tma_kernel(const T* __restrict__ X, double* __restrict__ ch_sum, double* __restrict__ ch_sq, float* __restrict__ nc_sum,
                        int N, int Tn, int V, int C, int nrep, int chunk_rows, unsigned* err) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  StreamPipe<1> pipe;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(T);
  pipe.init(st_smem, static_cast<uint32_t>(chunk_rows) * row_bytes);
  float* red = reinterpret_cast<float*>(st_smem + (pipe.empty0 + 8u * kStStages + 64u - smem_u32(st_smem)));
  float* scratch = red + 2 * C;
  ChunkIter it(static_cast<long long>(N) * Tn * V, 1, Tn * V, chunk_rows, nc_sum != nullptr);
  long long row0;
  int nrows, n, i = 0;
  if (threadIdx.x >= kStConsumers) {
    if (threadIdx.x == kStConsumers) {
      const void* const src[1] = {X};
      while (it.next(row0, nrows, n)) pipe.produce(i++, src, row0, nrows, row_bytes, err);
    }
    return;
  }
  const int c8n = C / 8;
  const int c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, RL = kStConsumers / c8n;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const uint32_t toff = static_cast<uint32_t>(threadIdx.x) * 8u * sizeof(T);
  const uint32_t rstep = static_cast<uint32_t>(RL) * row_bytes;
  int cur_n = -1;
  auto flush = [&](int nn) {
    consumer_channel_reduce<2>(acc, scratch, red, C, c8, c8n, rl, RL);
    for (int c = threadIdx.x; c < C; c += kStConsumers) {
      const size_t ro = static_cast<size_t>(replica_of_block(nrep)) * C;
      if (ch_sum) atomic_add_f64(ch_sum + ro + c, static_cast<double>(red[2 * c]));
      if (ch_sq) atomic_add_f64(ch_sq + ro + c, static_cast<double>(red[2 * c + 1]));
      if (nc_sum) atomicAdd(nc_sum + static_cast<size_t>(nn) * C + c, red[2 * c]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  };
  while (it.next(row0, nrows, n)) {
    if (nc_sum && n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
    }
    cur_n = n;
    const int s = pipe.acquire(i++, err);
    uint32_t a = pipe.tensor(s, 0) + toff;
#pragma unroll 4
    for (int r = rl; r < nrows; r += RL, a += rstep) {
      float f[8];
      lds8(a, f, static_cast<const T*>(nullptr));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += f[j];
        acc[1][j] = fmaf(f[j], f[j], acc[1][j]);
      }
    }
    pipe.release(s);
  }
  if (cur_n >= 0) flush(cur_n);
}

// bn1_bwd_apply: dG = c1[c]*dy1 + c2[c]*G + c3[c]  (dy1 as above); Tbl[v][c] += sum_{n,t} dG
template <typename T>
__global__ void __launch_bounds__(512) bn1_bwd_apply_kernel(const T* __restrict__ dH, const T* __restrict__ G, const float* __restrict__ a1,
                                     const float* __restrict__ b1, const float* __restrict__ c1,
                                     const float* __restrict__ c2, const float* __restrict__ c3, T* __restrict__ dG,
                                     float* __restrict__ Tbl, int N, int Tn, int V, int C, int nrep) {
  // balanced persistent partition over the N*Tn frames (nothing here depends on the clip); fewer, longer blocks also
  // mean fewer [V][C] table flushes: 8.6 M float atomics per launch at C = 256 with the old N x T/chunk grid
  const long long frames = static_cast<long long>(N) * Tn;
  const long long t0 = frames * blockIdx.x / gridDim.x, t1 = frames * (blockIdx.x + 1) / gridDim.x;
  const int c8n = C / 8;
  const int pairs = V * c8n;
  for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
    const int c0 = (i % c8n) * 8;
    float a[8], b[8], k1[8], k2[8], k3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = a1[c0 + j];
      b[j] = b1[c0 + j];
      k1[j] = c1[c0 + j];
      k2[j] = c2[c0 + j];
      k3[j] = c3[c0 + j];
    }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    constexpr int NB = sizeof(T) == 2 ? 4 : 2;   // frames in flight per thread (2 tensors each)
    for (long long tb = t0; tb < t1; tb += NB) {
      Raw8<T> rh[NB], rg[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q)
        if (tb + q < t1) {
          const size_t off = static_cast<size_t>(tb + q) * V * C + static_cast<size_t>(i) * 8;
          ldraw8(dH + off, rh[q]);
          ldraw8(G + off, rg[q]);
        }
#pragma unroll
      for (int q = 0; q < NB; ++q)
        if (tb + q < t1) {
          const size_t off = static_cast<size_t>(tb + q) * V * C + static_cast<size_t>(i) * 8;
          float g[8], h[8], o[8];
          unpack8(rh[q], h);
          unpack8(rg[q], g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = fmaf(a[j], g[j], b[j]) > 0.f ? h[j] : 0.f;
            o[j] = fmaf(k1[j], d, fmaf(k2[j], g[j], k3[j]));
            acc[j] += to_f32(from_f32<T>(o[j]));
          }
          store8(dG + off, o);
        }
    }
    if (Tbl) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        atomicAdd(Tbl + static_cast<size_t>(replica_of_block(nrep)) * V * C + static_cast<size_t>(i) * 8 + j, acc[j]);
    }
  }
}

// bn1_bwd_apply with the two inputs staged through the TMA ring of stream.cuh. Chunks are whole frames; consumer thread t owns the
// (joint, channel-group) pairs t, t + 256, ... of every frame (the channel group is the same for all of them: 256 is a multiple
// of C/8), keeps their [v][c] table sums in registers and flushes them once.
template <typename T, int MP>   // MP = pairs per thread = ceil(V * C/8 / 256)
__global__ void __launch_bounds__(kStThreads, 1)
    bn1_bwd_apply_tma_kernel(const T* __restrict__ dH, const T* __restrict__ G, const float* __restrict__ a1,
                             const float* __restrict__ b1, const float* __restrict__ c1, const float* __restrict__ c2,
                             const float* __restrict__ c3, T* __restrict__ dG, float* __restrict__ Tbl, int N, int Tn, int V, int C,
                             int nrep, int chunk_rows, unsigned* err) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  StreamPipe<2> pipe;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(T);
  pipe.init(st_smem, static_cast<uint32_t>(chunk_rows) * row_bytes);
  ChunkIter it(static_cast<long long>(N) * Tn * V, V, Tn * V, chunk_rows, false);
  long long row0;
  int nrows, n, i = 0;
  if (threadIdx.x >= kStConsumers) {
    if (threadIdx.x == kStConsumers) {
      const void* const src[2] = {dH, G};
      while (it.next(row0, nrows, n)) pipe.produce(i++, src, row0, nrows, row_bytes, err);
    }
    return;
  }
  const int c8n = C / 8, pairs = V * c8n;
  const int c0 = (threadIdx.x % c8n) * 8;
  float a[8], b[8], k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = a1[c0 + j];
    b[j] = b1[c0 + j];
    k1[j] = c1[c0 + j];
    k2[j] = c2[c0 + j];
    k3[j] = c3[c0 + j];
  }
  float acc[MP][8];
#pragma unroll
  for (int m = 0; m < MP; ++m)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[m][j] = 0.f;
  const uint32_t frame_bytes = static_cast<uint32_t>(V) * row_bytes;
  while (it.next(row0, nrows, n)) {
    const int s = pipe.acquire(i++, err);
    const uint32_t sh = pipe.tensor(s, 0), sg = pipe.tensor(s, 1);
    const int nfr = nrows / V;
    for (int f = 0; f < nfr; ++f) {
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        const int pi = static_cast<int>(threadIdx.x) + m * kStConsumers;
        if (pi < pairs) {
          const uint32_t so = static_cast<uint32_t>(f) * frame_bytes + static_cast<uint32_t>(pi) * 8u * sizeof(T);
          float g[8], h[8], o[8];
          lds8(sh + so, h, static_cast<const T*>(nullptr));
          lds8(sg + so, g, static_cast<const T*>(nullptr));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = fmaf(a[j], g[j], b[j]) > 0.f ? h[j] : 0.f;
            o[j] = fmaf(k1[j], d, fmaf(k2[j], g[j], k3[j]));
            acc[m][j] += to_f32(from_f32<T>(o[j]));
          }
          store8(dG + (static_cast<size_t>(row0) + static_cast<size_t>(f) * V) * C + static_cast<size_t>(pi) * 8, o);
        }
      }
    }
    pipe.release(s);
  }
  if (Tbl) {
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      const int pi = static_cast<int>(threadIdx.x) + m * kStConsumers;
      if (pi < pairs) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          atomicAdd(Tbl + static_cast<size_t>(replica_of_block(nrep)) * V * C + static_cast<size_t>(pi) * 8 + j, acc[m][j]);
      }
    }
  }
}

// threads for the row-walk kernels: a multiple of the channel groups (fixed c8 per thread)
static inline int rowwalk_threads(int C) {
  const int c8n = C / 8;
  int t = (kEwThreads / c8n) * c8n;
  return t < c8n ? c8n : t;
}
// threads for the (joint, channel-group) pair kernels: balance the strided loop over the pairs
static inline int pair_threads(int pairs) {
  const int iters = (pairs + 511) / 512;
  int t = ((pairs + iters - 1) / iters + 31) / 32 * 32;
  return t > 1024 ? 1024 : t;
}

// one resident wave of blocks for the balanced persistent kernels (capped for tiny inputs: >= 32 rows per block)
template <typename Kern>
static inline int resident_grid(Kern kernel, int threads, size_t smem, long long rows) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 1;
  long long g = static_cast<long long>(nb) * num_sms();
  const long long cap = (rows + 31) / 32;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

static inline int pick_tchunk(int N, int Tn) {
  // enough blocks to fill the machine a few times over, but long enough frame walks per thread
  int chunk = Tn;
  while (chunk > 4 && static_cast<long long>(N) * ((Tn + chunk - 1) / chunk) < 4LL * num_sms()) chunk = (chunk + 1) / 2;
  return chunk;
}

}  // namespace fmm

using namespace fmm;

#define FMM_DISPATCH(dtype, ...)                          \
  if ((dtype) == FMM_DT_BF16) {                           \
    using T = __nv_bfloat16;                              \
    __VA_ARGS__                                           \
  } else if ((dtype) == FMM_DT_F32) {                     \
    using T = float;                                      \
    __VA_ARGS__                                           \
  } else {                                                \
    fmm::set_last_error("bad dtype %d", (int)(dtype));    \
    return FMM_ERR_ARG;                                   \
  }

extern "C" {

int fmm_agg_fwd(const void* x, void* xa, const int* rowptr, const int* src, const float* coef, int N, int Tn,
                int V, int Cin, int K, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(x && xa && rowptr && src && coef && N > 0 && Tn > 0 && V > 0 && Cin > 0 && K > 0, "agg_fwd: bad args");
  const int tchunk = pick_tchunk(N, Tn);
  dim3 grid((Tn + tchunk - 1) / tchunk, N);
  FMM_DISPATCH(dtype, {
    if (Cin % 8 == 0)
      agg_fwd_vec_kernel<T><<<grid, pair_threads(V * K * (Cin / 8)), 0, stream>>>((const T*)x, (T*)xa, rowptr, src, coef, Tn, V, Cin, K, tchunk);
    else
      agg_fwd_scalar_kernel<T><<<grid, kEwThreads, 0, stream>>>((const T*)x, (T*)xa, rowptr, src, coef, Tn, V, Cin, K, tchunk);
  })
  FMM_CHECK_LAUNCH("agg_fwd");
  return FMM_OK;
}

int fmm_agg_bwd(const void* P, const void* addend, void* dx, const int* rowptr, const int* dst, const int* kk,
                const float* coef, const void* x, const int* eid, float* dcoef, int N, int Tn, int V, int Cin, int K,
                int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(!x || (eid && dcoef), "agg_bwd: dcoef accumulation needs eid and dcoef");
  FMM_CHECK_ARG(!x || Cin % 8 == 0, "agg_bwd: fused dcoef needs Cin %% 8 == 0 (use agg_dcoef)");
  FMM_CHECK_ARG(P && dx && rowptr && dst && kk && coef && N > 0 && Tn > 0 && V > 0 && Cin > 0 && K > 0, "agg_bwd: bad args");
  const int tchunk = pick_tchunk(N, Tn);
  dim3 grid((Tn + tchunk - 1) / tchunk, N);
  FMM_DISPATCH(dtype, {
    if (Cin % 8 == 0)
      agg_bwd_vec_kernel<T><<<grid, pair_threads(V * (Cin / 8)), 0, stream>>>((const T*)P, (const T*)addend, (T*)dx, rowptr, dst, kk, coef, (const T*)x, eid, dcoef, Tn, V, Cin, K, tchunk);
    else
      agg_bwd_scalar_kernel<T><<<grid, kEwThreads, 0, stream>>>((const T*)P, (const T*)addend, (T*)dx, rowptr, dst, kk, coef, Tn, V, Cin, K, tchunk);
  })
  FMM_CHECK_LAUNCH("agg_bwd");
  return FMM_OK;
}

int fmm_agg_dcoef(const void* x, const void* P, float* dcoef, const int* src, const int* dst, const int* kk, int E,
                  int N, int Tn, int V, int Cin, int K, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(x && P && dcoef && src && dst && kk && E > 0 && N > 0 && Tn > 0, "agg_dcoef: bad args");
  if (Cin <= 8) {
    FMM_DISPATCH(dtype, {
      agg_dcoef_narrow_kernel<T><<<N, 512, 0, stream>>>((const T*)x, (const T*)P, dcoef, src, dst, kk, E, Tn, V, Cin, K);
    })
    FMM_CHECK_LAUNCH("agg_dcoef");
    return FMM_OK;
  }
  const int tchunk = pick_tchunk(N, Tn);
  dim3 grid((Tn + tchunk - 1) / tchunk, N);
  FMM_DISPATCH(dtype, {
    agg_dcoef_kernel<T><<<grid, kEwThreads, 0, stream>>>((const T*)x, (const T*)P, dcoef, src, dst, kk, E, Tn, V, Cin, K, tchunk);
  })
  FMM_CHECK_LAUNCH("agg_dcoef");
  return FMM_OK;
}

int fmm_colstats(const void* X, double* ch_sum, double* ch_sq, float* nc_sum, int nrep, int N, int Tn, int V, int C,
                 int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(nrep >= 1, "colstats: nrep");
  FMM_CHECK_ARG(X && N > 0 && Tn > 0 && V > 0 && C > 0 && C % 8 == 0, "colstats: bad args (C must be a multiple of 8)");
  const int th = rowwalk_threads(C);
  const size_t sm = (2 * C + th * 16) * sizeof(float);
  const long long rows = static_cast<long long>(N) * Tn * V;
  {
    const char* tma_str = getenv("FMM_EW_TMA");
    if ((!tma_str || atoi(tma_str)) && kStConsumers % (C / 8) == 0 && rows >= 64) {
      FMM_DISPATCH(dtype, {
        const int cr = stream_chunk_rows(32 * 1024, C * sizeof(T), 1);
        const size_t smem = StreamPipe<1>::bytes(cr * C * sizeof(T)) + 64 + (2 * C + kStConsumers * 16) * sizeof(float);
        cudaFuncSetAttribute(colstats_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        colstats_tma_kernel<T><<<num_sms(), kStThreads, smem, stream>>>((const T*)X, ch_sum, ch_sq, nc_sum, N, Tn, V, C, nrep, cr, nullptr);
      })
      FMM_CHECK_LAUNCH("colstats");
      return FMM_OK;
    }
  }
  FMM_DISPATCH(dtype, {
    colstats_kernel<T><<<resident_grid(colstats_kernel<T>, th, sm, rows), th, sm, stream>>>((const T*)X, ch_sum, ch_sq, nc_sum, N, Tn, V, C, nrep);
  })
  FMM_CHECK_LAUNCH("colstats");
  return FMM_OK;
}

int fmm_block_out(const void* U, const float* k1, const float* k0, const void* res, const float* ar, const float* br,
                  void* Y, int N, int Tn, int V, int C, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(U && k1 && k0 && Y && C % 8 == 0, "block_out: bad args");
  const int th = rowwalk_threads(C);
  const long long rows = static_cast<long long>(N) * Tn * V;
  FMM_DISPATCH(dtype, {
    block_out_kernel<T><<<resident_grid(block_out_kernel<T>, th, 0, rows), th, 0, stream>>>((const T*)U, k1, k0, (const T*)res, ar, br, (T*)Y, N, Tn, V, C);
  })
  FMM_CHECK_LAUNCH("block_out");
  return FMM_OK;
}

int fmm_affine_relu(const void* X, const float* a, const float* b, void* H, int N, int Tn, int V, int C, int dtype,
                    cudaStream_t stream) {
  FMM_CHECK_ARG(X && a && b && H && C % 8 == 0, "affine_relu: bad args");
  const int th = rowwalk_threads(C);
  const long long rows = static_cast<long long>(N) * Tn * V;
  FMM_DISPATCH(dtype, {
    affine_relu_kernel<T><<<resident_grid(affine_relu_kernel<T>, th, 0, rows), th, 0, stream>>>((const T*)X, a, b, (T*)H, N, Tn, V, C);
  })
  FMM_CHECK_LAUNCH("affine_relu");
  return FMM_OK;
}

int fmm_blockout_bwd_reduce(const void* dY, const void* Y, const void* U, const void* R, float* S1, float* S2,
                            float* S3, int N, int Tn, int V, int C, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(dY && U && S1 && S2 && C % 8 == 0 && (!R || S3), "blockout_bwd_reduce: bad args");
  const int th = rowwalk_threads(C);
  const size_t sm = (3 * C + th * 24) * sizeof(float);
  const long long rows = static_cast<long long>(N) * Tn * V;
  {
    const char* tma_str = getenv("FMM_EW_TMA");
    if ((!tma_str || atoi(tma_str)) && kStConsumers % (C / 8) == 0 && rows >= 64) {
      const int live = 2 + (R ? 1 : 0) + (Y ? 1 : 0);
#define FMM_BOUT_REDUCE_TMA(RES, MASK)                                                                                        \
  do {                                                                                                                        \
    const int cr = stream_chunk_rows(32 * 1024 / live, C * sizeof(T), 1);                                                     \
    const size_t smem = StreamPipe<2 + (RES ? 1 : 0) + (MASK ? 1 : 0)>::bytes(cr * C * sizeof(T)) + 64 +                       \
                        (3 * C + kStConsumers * 24) * sizeof(float);                                                           \
    cudaFuncSetAttribute(blockout_bwd_reduce_tma_kernel<T, RES, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    blockout_bwd_reduce_tma_kernel<T, RES, MASK><<<num_sms(), kStThreads, smem, stream>>>(                                     \
        (const T*)dY, (const T*)Y, (const T*)U, (const T*)R, S1, S2, S3, N, Tn, V, C, cr, nullptr);                            \
  } while (0)
      FMM_DISPATCH(dtype, {
        if (R) {
          if (Y) FMM_BOUT_REDUCE_TMA(true, true); else FMM_BOUT_REDUCE_TMA(true, false);
        } else {
          if (Y) FMM_BOUT_REDUCE_TMA(false, true); else FMM_BOUT_REDUCE_TMA(false, false);
        }
      })
#undef FMM_BOUT_REDUCE_TMA
      FMM_CHECK_LAUNCH("blockout_bwd_reduce");
      return FMM_OK;
    }
  }
#define FMM_BOUT_REDUCE(RES, MASK)                                                                                     \
  blockout_bwd_reduce_kernel<T, RES, MASK><<<resident_grid(blockout_bwd_reduce_kernel<T, RES, MASK>, th, sm, rows), th, sm, stream>>>( \
      (const T*)dY, (const T*)Y, (const T*)U, (const T*)R, S1, S2, S3, N, Tn, V, C)
  FMM_DISPATCH(dtype, {
    if (R) {
      if (Y) FMM_BOUT_REDUCE(true, true); else FMM_BOUT_REDUCE(true, false);
    } else {
      if (Y) FMM_BOUT_REDUCE(false, true); else FMM_BOUT_REDUCE(false, false);
    }
  })
#undef FMM_BOUT_REDUCE
  FMM_CHECK_LAUNCH("blockout_bwd_reduce");
  return FMM_OK;
}

int fmm_bn2_bwd_apply(const void* dY, const void* Y, const void* U, const void* R, const float* k1, const float* k2,
                      const float* k3, const float* r1, const float* r2, const float* r3, void* dU, void* dR,
                      void* dPre, double* sum_dU, double* sum_dR, int nrep, int N, int Tn, int V, int C, int dtype,
                      cudaStream_t stream) {
  FMM_CHECK_ARG(nrep >= 1, "bn2_bwd_apply: nrep");
  FMM_CHECK_ARG(dY && U && k1 && k2 && k3 && dU && C % 8 == 0, "bn2_bwd_apply: bad args");
  FMM_CHECK_ARG(!R || (r1 && r2 && r3 && dR), "bn2_bwd_apply: residual branch needs r1..r3 and dR");
  const int th = rowwalk_threads(C);
  const size_t sm = (2 * C + th * 16) * sizeof(float);
  const long long rows = static_cast<long long>(N) * Tn * V;
#define FMM_BN2_APPLY(RES, MASK)                                                                                       \
  bn2_bwd_apply_kernel<T, RES, MASK><<<resident_grid(bn2_bwd_apply_kernel<T, RES, MASK>, th, sm, rows), th, sm, stream>>>( \
      (const T*)dY, (const T*)Y, (const T*)U, (const T*)R, k1, k2, k3, r1, r2, r3, (T*)dU, (T*)dR, (T*)dPre, sum_dU, sum_dR, N, Tn, V, C, nrep)
  FMM_DISPATCH(dtype, {
    if (R) {
      if (Y) FMM_BN2_APPLY(true, true); else FMM_BN2_APPLY(true, false);
    } else {
      if (Y) FMM_BN2_APPLY(false, true); else FMM_BN2_APPLY(false, false);
    }
  })
#undef FMM_BN2_APPLY
  FMM_CHECK_LAUNCH("bn2_bwd_apply");
  return FMM_OK;
}

int fmm_bn1_bwd_reduce(const void* dH, const void* G, const float* a1, const float* b1, double* T1, double* T2,
                       int nrep, int N, int Tn, int V, int C, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(nrep >= 1, "bn1_bwd_reduce: nrep");
  FMM_CHECK_ARG(dH && G && a1 && b1 && T1 && T2 && C % 8 == 0, "bn1_bwd_reduce: bad args");
  const int th = rowwalk_threads(C);
  const size_t sm = (2 * C + th * 16) * sizeof(float);
  const long long rows = static_cast<long long>(N) * Tn * V;
  static const int tma_env = getenv("FMM_EW_TMA") ? atoi(getenv("FMM_EW_TMA")) : 1;
  if (tma_env && kStConsumers % (C / 8) == 0 && rows >= 64) {
    FMM_DISPATCH(dtype, {
      const int cr = stream_chunk_rows(16 * 1024, C * sizeof(T), 1);
      const size_t smem = StreamPipe<2>::bytes(cr * C * sizeof(T)) + 64 + (2 * C + kStConsumers * 16) * sizeof(float);
      cudaFuncSetAttribute(bn1_bwd_reduce_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      bn1_bwd_reduce_tma_kernel<T><<<num_sms(), kStThreads, smem, stream>>>((const T*)dH, (const T*)G, a1, b1, T1, T2, N, Tn, V, C, nrep, cr, nullptr);
    })
    FMM_CHECK_LAUNCH("bn1_bwd_reduce");
    return FMM_OK;
  }
  FMM_DISPATCH(dtype, {
    bn1_bwd_reduce_kernel<T><<<resident_grid(bn1_bwd_reduce_kernel<T>, th, sm, rows), th, sm, stream>>>((const T*)dH, (const T*)G, a1, b1, T1, T2, N, Tn, V, C, nrep);
  })
  FMM_CHECK_LAUNCH("bn1_bwd_reduce");
  return FMM_OK;
}

int fmm_bn1_bwd_apply(const void* dH, const void* G, const float* a1, const float* b1, const float* c1,
                      const float* c2, const float* c3, void* dG, float* Tbl, int nrep, int N, int Tn, int V, int C,
                      int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(nrep >= 1, "bn1_bwd_apply: nrep");
  FMM_CHECK_ARG(dH && G && a1 && b1 && c1 && c2 && c3 && dG && C % 8 == 0, "bn1_bwd_apply: bad args");
  const int th = pair_threads(V * (C / 8));   // <= 512 (the kernel's launch bound)
  const long long frames = static_cast<long long>(N) * Tn;
  {
    const char* tma_str = getenv("FMM_EW_TMA");
    const int c8n = C / 8, mp = (V * c8n + kStConsumers - 1) / kStConsumers;
    const size_t es = dtype == FMM_DT_BF16 ? 2 : 4;
    const int cr = stream_chunk_rows(16 * 1024, C * es, V);
    const size_t smem = StreamPipe<2>::bytes(static_cast<uint32_t>(cr * C * es));
    if ((!tma_str || atoi(tma_str)) && kStConsumers % c8n == 0 && mp >= 1 && mp <= 5 && frames >= 2 && smem <= 200 * 1024) {
#define FMM_BN1_APPLY_TMA(MPV)                                                                                              \
  do {                                                                                                                     \
    cudaFuncSetAttribute(bn1_bwd_apply_tma_kernel<T, MPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    int g = num_sms() < frames ? num_sms() : static_cast<int>(frames);                                                     \
    bn1_bwd_apply_tma_kernel<T, MPV><<<g, kStThreads, smem, stream>>>((const T*)dH, (const T*)G, a1, b1, c1, c2, c3, (T*)dG, Tbl, N, \
                                                                      Tn, V, C, nrep, cr, nullptr);                       \
  } while (0)
      FMM_DISPATCH(dtype, {
        switch (mp) {
          case 1: FMM_BN1_APPLY_TMA(1); break;
          case 2: FMM_BN1_APPLY_TMA(2); break;
          case 3: FMM_BN1_APPLY_TMA(3); break;
          case 4: FMM_BN1_APPLY_TMA(4); break;
          default: FMM_BN1_APPLY_TMA(5); break;
        }
      })
#undef FMM_BN1_APPLY_TMA
      FMM_CHECK_LAUNCH("bn1_bwd_apply");
      return FMM_OK;
    }
  }
  FMM_DISPATCH(dtype, {
    int g = resident_grid(bn1_bwd_apply_kernel<T>, th, 0, frames * 32);
    if (g > frames) g = static_cast<int>(frames);
    bn1_bwd_apply_kernel<T><<<g, th, 0, stream>>>((const T*)dH, (const T*)G, a1, b1, c1, c2, c3, (T*)dG, Tbl, N, Tn, V, C, nrep);
  })
  FMM_CHECK_LAUNCH("bn1_bwd_apply");
  return FMM_OK;
}

}  // extern "C"
