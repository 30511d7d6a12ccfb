// Strided, batched GEMM for the TRAGCN family (graph-GRU cells, the time-axis attention, the head):
//
//   C[g1,g2][m][n] = act( alpha * sum_k A[g1,g2][m][k] * B[g1,g2][k][n] + bias_m[m] + bias_n[n] ) (+ C)
//
// Every operand is addressed through element strides, the contraction index may be a product of up
// to three strided levels (k = (k1*K2 + k2)*K3 + k3), and a stride of 0 broadcasts.  That lets one
// kernel express the per-node weight products of EmbGCN (batch = node), the adjacency mixes, the
// (1,3) convolutions whose channel axis is TIME, QK^T / PV of the attention, every Linear, and all
// of their gradients (transposes are stride swaps; weight gradients contract over batch levels)
// without materialising a single permuted copy.
//
// The shapes here are small per batch entry (V = 25 rows, 64..300 wide) so tiles are 64x64x32 on
// warp-level bf16 tensor-core MMAs (mma.sync.m16n8k16, fp32 accumulate) rather than the 128-row
// tcgen05 tiles of tapconv.cu.  fp32 operands run as three bf16 parts (6 product terms, the
// corrections in their own accumulator) like the fp32 mode of the other engines.
#include "common.cuh"

namespace fmm {

struct BgemmDesc {
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
  int beta;     // 1: add the previous contents of C
  int act;      // 0 none, 1 relu
  int splitk;   // >1: partial sums are atomically added to an fp32 C (no bias / act / beta)
  int dtype;    // operand type of A and B
  int c_dtype;  // type of C
};

constexpr int kBM = 64, kBN = 64, kBK = 32, kPitch = kBK + 8, kThreads = 128;

// operand tile loaders: 64 rows x 32 k, 16 elements per thread
// mode 0: scalar, k fastest   mode 1: scalar, row fastest   mode 2: 8-vectors along k   mode 3: 8-vectors along rows
struct Operand {
  const void* base;
  long long row_stride, k1s, k2s, k3s;
  int rows;  // valid rows of this matrix (M or N)
  int mode;
};

__device__ __forceinline__ long long koff(int k, int K2, int K3, long long s1, long long s2, long long s3) {
  if (K2 == 1 && K3 == 1) return (long long)k * s1;
  int k3 = k % K3;
  int q = k / K3;
  int k2 = q % K2;
  int k1 = q / K2;
  return k1 * s1 + k2 * s2 + k3 * s3;
}

template <typename T>
__device__ __forceinline__ void load_operand(const Operand& op, int row0, int k0, int K, int K2, int K3, int tid,
                                             T (&r)[16]) {
  const T* base = reinterpret_cast<const T*>(op.base);
  if (op.mode == 0) {
    int kc = tid & 31;
    int k = k0 + kc;
    bool kok = k < K;
    long long ko = kok ? koff(k, K2, K3, op.k1s, op.k2s, op.k3s) : 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int row = row0 + i * 4 + (tid >> 5);
      r[i] = (kok && row < op.rows) ? base[ko + row * op.row_stride] : from_f32<T>(0.f);
    }
  } else if (op.mode == 1) {
    int row = row0 + (tid & 63);
    bool rok = row < op.rows;
    long long ro = row * op.row_stride;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int k = k0 + i * 2 + (tid >> 6);
      r[i] = (rok && k < K) ? base[ro + koff(k, K2, K3, op.k1s, op.k2s, op.k3s)] : from_f32<T>(0.f);
    }
  } else if (op.mode == 2) {
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int v = i * kThreads + tid;
        int row = row0 + (v >> 2);
        int k = k0 + (v & 3) * 8;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (row < op.rows && k < K) {
          const T* src = base + row * op.row_stride + koff(k, K2, K3, op.k1s, op.k2s, op.k3s);
          if (k + 8 <= K) {
            u = *reinterpret_cast<const uint4*>(src);
          } else {  // ragged end of a single-level contraction
            T* e = reinterpret_cast<T*>(&u);
            for (int j = 0; j < K - k; ++j) e[j] = src[j];
          }
        }
        *reinterpret_cast<uint4*>(&r[i * 8]) = u;
      }
    }
  } else {
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int v = i * kThreads + tid;
        int row = row0 + (v & 7) * 8;
        int k = k0 + (v >> 3);
        uint4 u = make_uint4(0, 0, 0, 0);
        if (row < op.rows && k < K) {
          const T* src = base + row + koff(k, K2, K3, op.k1s, op.k2s, op.k3s);
          if (row + 8 <= op.rows) {
            u = *reinterpret_cast<const uint4*>(src);
          } else {  // ragged last rows
            T* e = reinterpret_cast<T*>(&u);
            for (int j = 0; j < op.rows - row; ++j) e[j] = src[j];
          }
        }
        *reinterpret_cast<uint4*>(&r[i * 8]) = u;
      }
    }
  }
}

// smem tile [part][row][kPitch]; position of staged element i for each mode
template <typename T, int kParts>
__device__ __forceinline__ void store_operand(__nv_bfloat16 (*S)[kBM][kPitch], int mode, int tid, const T (&r)[16]) {
  auto put = [&](int row, int kc, T v) {
    if constexpr (kParts == 1) {
      S[0][row][kc] = __float2bfloat16_rn(to_f32(v));
    } else {
      float f = to_f32(v);
#pragma unroll
      for (int p = 0; p < kParts; ++p) {
        __nv_bfloat16 h = __float2bfloat16_rn(f);
        S[p][row][kc] = h;
        f -= __bfloat162float(h);
      }
    }
  };
  if (mode == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) put(i * 4 + (tid >> 5), tid & 31, r[i]);
  } else if (mode == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) put(tid & 63, i * 2 + (tid >> 6), r[i]);
  } else if (mode == 2) {
    if constexpr (sizeof(T) == 2 && kParts == 1) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int v = i * kThreads + tid;
        *reinterpret_cast<uint4*>(&S[0][v >> 2][(v & 3) * 8]) = *reinterpret_cast<const uint4*>(&r[i * 8]);
      }
    }
  } else {
    if constexpr (sizeof(T) == 2 && kParts == 1) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int v = i * kThreads + tid;
#pragma unroll
        for (int j = 0; j < 8; ++j) S[0][(v & 7) * 8 + j][v >> 3] = r[i * 8 + j];
      }
    }
  }
}

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// epilogue shared by both kernels: a warp's 32x32 accumulator block at (mw, nw)
__device__ __forceinline__ void store_tile(const BgemmDesc& p, const float (&acc)[2][4][4], int mw, int nw, int lane,
                                           long long cbase) {
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int m = mw + mi * 16 + gq + hh * 8;
      if (m >= p.M) continue;
      const long long rowoff = cbase + m * p.c_m;
      const float bm = (p.bias_m && p.splitk == 1) ? p.bias_m[m] : 0.f;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = nw + ni * 8 + 2 * tq + e;
          if (n >= p.N) continue;
          float v = acc[mi][ni][hh * 2 + e] * p.alpha;
          const long long off = rowoff + n * p.c_n;
          if (p.splitk > 1) {
            atomicAdd(reinterpret_cast<float*>(p.C) + off, v);
            continue;
          }
          v += bm;
          if (p.bias_n) v += p.bias_n[n];
          if (p.act == 1) v = fmaxf(v, 0.f);
          if (p.c_dtype == FMM_DT_F32) {
            float* c = reinterpret_cast<float*>(p.C) + off;
            *c = p.beta ? *c + v : v;
          } else {
            __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + off;
            *c = __float2bfloat16_rn(p.beta ? __bfloat162float(*c) + v : v);
          }
        }
    }
}

template <typename T, int kParts>
__global__ void __launch_bounds__(kThreads) bgemm_kernel(const BgemmDesc p, int a_mode, int b_mode) {
  __shared__ __align__(16) __nv_bfloat16 As[kParts][kBM][kPitch];
  __shared__ __align__(16) __nv_bfloat16 Bs[kParts][kBN][kPitch];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_m = (p.M + kBM - 1) / kBM, tiles_n = (p.N + kBN - 1) / kBN;
  const int per_g = tiles_m * tiles_n;
  const int g = blockIdx.x / per_g, tile = blockIdx.x % per_g;
  const int g1 = g / p.G2, g2 = g % p.G2;
  const int m0 = (tile % tiles_m) * kBM, n0 = (tile / tiles_m) * kBN;
  const int K = p.K1 * p.K2 * p.K3;
  const int ktiles = (K + kBK - 1) / kBK;
  const int per_split = (ktiles + p.splitk - 1) / p.splitk;
  const int kt0 = blockIdx.y * per_split, kt1 = min(ktiles, kt0 + per_split);
  if (kt0 >= kt1) return;

  Operand oa{reinterpret_cast<const T*>(p.A) + g1 * p.a_g1 + g2 * p.a_g2, p.a_m, p.a_k1, p.a_k2, p.a_k3, p.M, a_mode};
  Operand ob{reinterpret_cast<const T*>(p.B) + g1 * p.b_g1 + g2 * p.b_g2, p.b_n, p.b_k1, p.b_k2, p.b_k3, p.N, b_mode};

  float acc[2][4][4];
  float cor[kParts > 1 ? 2 : 1][kParts > 1 ? 4 : 1][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[i][j][q] = 0.f;
        if constexpr (kParts > 1) cor[i][j][q] = 0.f;
      }

  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int gq = lane >> 2, tq = lane & 3;
  T ra[16], rb[16];
  load_operand<T>(oa, m0, kt0 * kBK, K, p.K2, p.K3, tid, ra);
  load_operand<T>(ob, n0, kt0 * kBK, K, p.K2, p.K3, tid, rb);
  for (int kt = kt0; kt < kt1; ++kt) {
    __syncthreads();  // previous tile fully consumed
    store_operand<T, kParts>(As, a_mode, tid, ra);
    store_operand<T, kParts>(Bs, b_mode, tid, rb);
    __syncthreads();
    if (kt + 1 < kt1) {
      load_operand<T>(oa, m0, (kt + 1) * kBK, K, p.K2, p.K3, tid, ra);
      load_operand<T>(ob, n0, (kt + 1) * kBK, K, p.K2, p.K3, tid, rb);
    }
#pragma unroll
    for (int ks = 0; ks < kBK; ks += 16) {
      uint32_t af[kParts][2][4], bf[kParts][4][2];
#pragma unroll
      for (int pp = 0; pp < kParts; ++pp) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const __nv_bfloat16* r0 = &As[pp][wm + mi * 16 + gq][ks + 2 * tq];
          const __nv_bfloat16* r1 = &As[pp][wm + mi * 16 + gq + 8][ks + 2 * tq];
          af[pp][mi][0] = *reinterpret_cast<const uint32_t*>(r0);
          af[pp][mi][1] = *reinterpret_cast<const uint32_t*>(r1);
          af[pp][mi][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
          af[pp][mi][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const __nv_bfloat16* c0 = &Bs[pp][wn + ni * 8 + gq][ks + 2 * tq];
          bf[pp][ni][0] = *reinterpret_cast<const uint32_t*>(c0);
          bf[pp][ni][1] = *reinterpret_cast<const uint32_t*>(c0 + 8);
        }
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          mma16816(acc[mi][ni], af[0][mi], bf[0][ni]);
          if constexpr (kParts == 3) {
            // smallest terms first inside the correction accumulator
            mma16816(cor[mi][ni], af[0][mi], bf[2][ni]);
            mma16816(cor[mi][ni], af[2][mi], bf[0][ni]);
            mma16816(cor[mi][ni], af[1][mi], bf[1][ni]);
            mma16816(cor[mi][ni], af[0][mi], bf[1][ni]);
            mma16816(cor[mi][ni], af[1][mi], bf[0][ni]);
          }
        }
    }
  }

  if constexpr (kParts > 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] += cor[i][j][q];
  }
  store_tile(p, acc, m0 + wm, n0 + wn, lane, g1 * p.c_g1 + g2 * p.c_g2);
}

// ---------------------------------------------------------------------------------------------
// Pipelined variant for bf16 operands whose tiles can be fetched with 16-byte vectors (loader modes 2 / 3):
// 128x64x32 tiles, 3-stage cp.async ring straight into shared memory, ldmatrix(.trans) fragments.
// An operand that is contiguous along its rows (mode 3) is staged as [k][row] and read transposed.
// ---------------------------------------------------------------------------------------------
constexpr int pBM = 128, pBN = 64, pThreads = 256, pStages = 3;
constexpr int pAStage = pBM * kPitch, pBStage = pBN * kPitch;  // halves; also cover the [k][rows+8] layouts

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc, int bytes) {
  uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* ptr) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* ptr) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <int MODE, int ROWS>
__device__ __forceinline__ void pipe_load(__nv_bfloat16* S, const Operand& op, int row0, int k0, int K, int K2, int K3,
                                          int tid) {
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(op.base);
  constexpr int NV = ROWS * 4 / pThreads;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = i * pThreads + tid;
    const __nv_bfloat16* src = base;
    int bytes = 0;
    __nv_bfloat16* dst;
    if constexpr (MODE == 2) {
      const int r = v >> 2, kc = (v & 3) * 8;
      dst = S + r * kPitch + kc;
      const int row = row0 + r, k = k0 + kc;
      if (row < op.rows && k < K) {
        src = base + row * op.row_stride + koff(k, K2, K3, op.k1s, op.k2s, op.k3s);
        bytes = min(8, K - k) * 2;
      }
    } else {
      constexpr int RV = ROWS / 8;
      const int r8 = (v % RV) * 8, kc = v / RV;
      dst = S + kc * (ROWS + 8) + r8;
      const int row = row0 + r8, k = k0 + kc;
      if (row < op.rows && k < K) {
        src = base + row + koff(k, K2, K3, op.k1s, op.k2s, op.k3s);
        bytes = min(8, op.rows - row) * 2;
      }
    }
    cpa16(dst, src, bytes);
  }
}

template <int AM, int BMo>
__global__ void __launch_bounds__(pThreads, 3) bgemm_pipe_kernel(const BgemmDesc p, int vec_epi) {
  __shared__ __align__(128) __nv_bfloat16 smem[pStages * (pAStage + pBStage)];
  __nv_bfloat16(*As)[pAStage] = reinterpret_cast<__nv_bfloat16(*)[pAStage]>(smem);
  __nv_bfloat16(*Bs)[pBStage] = reinterpret_cast<__nv_bfloat16(*)[pBStage]>(smem + pStages * pAStage);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_m = (p.M + pBM - 1) / pBM, tiles_n = (p.N + pBN - 1) / pBN;
  const int per_g = tiles_m * tiles_n;
  const int g = blockIdx.x / per_g, tile = blockIdx.x % per_g;
  const int g1 = g / p.G2, g2 = g % p.G2;
  const int m0 = (tile % tiles_m) * pBM, n0 = (tile / tiles_m) * pBN;
  const int K = p.K1 * p.K2 * p.K3;
  const int ktiles = (K + kBK - 1) / kBK;
  const int per_split = (ktiles + p.splitk - 1) / p.splitk;
  const int kt0 = blockIdx.y * per_split, kt1 = min(ktiles, kt0 + per_split);
  if (kt0 >= kt1) return;
  const int nk = kt1 - kt0;

  const Operand oa{reinterpret_cast<const __nv_bfloat16*>(p.A) + g1 * p.a_g1 + g2 * p.a_g2, p.a_m, p.a_k1, p.a_k2, p.a_k3, p.M, AM};
  const Operand ob{reinterpret_cast<const __nv_bfloat16*>(p.B) + g1 * p.b_g1 + g2 * p.b_g2, p.b_n, p.b_k1, p.b_k2, p.b_k3, p.N, BMo};

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  auto issue = [&](int i) {
    if (i < nk) {
      const int st = i % pStages, k0 = (kt0 + i) * kBK;
      pipe_load<AM, pBM>(As[st], oa, m0, k0, K, p.K2, p.K3, tid);
      pipe_load<BMo, pBN>(Bs[st], ob, n0, k0, K, p.K2, p.K3, tid);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
#pragma unroll
  for (int s = 0; s < pStages - 1; ++s) issue(s);
  for (int i = 0; i < nk; ++i) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(pStages - 2) : "memory");
    __syncthreads();
    issue(i + pStages - 1);
    const __nv_bfloat16* Ast = As[i % pStages];
    const __nv_bfloat16* Bst = Bs[i % pStages];
#pragma unroll
    for (int ks = 0; ks < kBK; ks += 16) {
      uint32_t af[2][4], bf[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if constexpr (AM == 2)
          ldsm4(af[mi], Ast + (wm + mi * 16 + (lane & 15)) * kPitch + ks + (lane >> 4) * 8);
        else
          ldsm4t(af[mi], Ast + (ks + (lane & 7) + ((lane >> 4) & 1) * 8) * (pBM + 8) + wm + mi * 16 + ((lane >> 3) & 1) * 8);
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {
        if constexpr (BMo == 2)
          ldsm4(bf[nj], Bst + (wn + nj * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * kPitch + ks + ((lane >> 3) & 1) * 8);
        else
          ldsm4t(bf[nj], Bst + (ks + (lane & 7) + ((lane >> 3) & 1) * 8) * (pBN + 8) + wn + nj * 16 + ((lane >> 4) & 1) * 8);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const uint32_t b2[2] = {bf[ni >> 1][(ni & 1) * 2], bf[ni >> 1][(ni & 1) * 2 + 1]};
          mma16816(acc[mi][ni], af[mi], b2);
        }
    }
  }
  const long long cbase = g1 * p.c_g1 + g2 * p.c_g2;
  if (!vec_epi) {
    store_tile(p, acc, m0 + wm, n0 + wn, lane, cbase);
    return;
  }
  // rows of C are contiguous and 16-byte aligned: stage the 128x64 tile in shared memory, store 16-byte vectors
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  constexpr int cP = pBN + 8;
  float* Cs = reinterpret_cast<float*>(smem);  // [128][72] fp32 = 36 KB of the 45 KB ring
  {
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          *reinterpret_cast<float2*>(&Cs[(wm + mi * 16 + gq + hh * 8) * cP + wn + ni * 8 + 2 * tq]) =
              make_float2(acc[mi][ni][hh * 2], acc[mi][ni][hh * 2 + 1]);
  }
  __syncthreads();
  const bool f32 = p.c_dtype == FMM_DT_F32;
  const int per = f32 ? 4 : 8, vpr = pBN / per;  // elements per 16-byte vector, vectors per tile row
  for (int v = tid; v < pBM * vpr; v += pThreads) {
    const int r = v / vpr, c = (v % vpr) * per;
    const int m = m0 + r, n = n0 + c;
    if (m >= p.M || n >= p.N) continue;
    float vals[8];
    const float bm = p.bias_m ? p.bias_m[m] : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < per) {
        float x = Cs[r * cP + c + e] * p.alpha + bm;
        if (p.bias_n && n + e < p.N) x += p.bias_n[n + e];
        vals[e] = p.act == 1 ? fmaxf(x, 0.f) : x;
      }
    }
    const long long off = cbase + m * p.c_m + n;
    if (n + per <= p.N) {
      if (f32) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + off);
        float4 o = make_float4(vals[0], vals[1], vals[2], vals[3]);
        if (p.beta) {
          const float4 old = *dst;
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *dst = o;
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + off;
        if (p.beta) {
          float old[8];
          load8(dst, old);
#pragma unroll
          for (int e = 0; e < 8; ++e) vals[e] += old[e];
        }
        store8(dst, vals);
      }
    } else {
      for (int e = 0; e < per && n + e < p.N; ++e) {
        if (f32) {
          float* dst = reinterpret_cast<float*>(p.C) + off + e;
          *dst = p.beta ? *dst + vals[e] : vals[e];
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + off + e;
          *dst = __float2bfloat16_rn(p.beta ? __bfloat162float(*dst) + vals[e] : vals[e]);
        }
      }
    }
  }
}

static bool mult8(long long v) { return (v & 7) == 0; }

// pick the loader for one operand (see load_operand)
static int pick_mode(const void* base, int dtype, long long row_stride, long long g1s, long long g2s, long long k1s,
                     long long k2s, long long k3s, int rows, int K1, int K2, int K3) {
  bool aligned = (reinterpret_cast<uintptr_t>(base) & 15) == 0 && mult8(g1s) && mult8(g2s);
  // innermost contraction level actually walked
  long long kin = K3 > 1 ? k3s : (K2 > 1 ? k2s : k1s);
  int Kin = K3 > 1 ? K3 : (K2 > 1 ? K2 : K1);
  if (dtype == FMM_DT_BF16 && aligned) {
    bool outer8 = (K3 > 1 ? (mult8(k2s) || K2 == 1) && (mult8(k1s) || K1 == 1) : (K2 > 1 ? (mult8(k1s) || K1 == 1) : true));
    if (kin == 1 && ((Kin % 8) == 0 || (K2 == 1 && K3 == 1)) && mult8(row_stride) && outer8) {
      // a vector must not straddle contraction levels: with K3 (or K2) the innermost extent this holds as Kin % 8 == 0
      return 2;
    }
    if (row_stride == 1 && mult8(k1s) && (K2 == 1 || mult8(k2s)) && (K3 == 1 || mult8(k3s))) return 3;
  }
  long long rs = row_stride < 0 ? -row_stride : row_stride;
  long long ks = kin < 0 ? -kin : kin;
  return (rs != 0 && (ks == 0 || rs < ks)) ? 1 : 0;
}

}  // namespace fmm

extern "C" int fmm_bgemm(const fmm::BgemmDesc* d, void* stream) {
  using namespace fmm;
  FMM_CHECK_ARG(d && d->A && d->B && d->C, "bgemm: null operand");
  FMM_CHECK_ARG(d->G1 > 0 && d->G2 > 0 && d->M > 0 && d->N > 0 && d->K1 > 0 && d->K2 > 0 && d->K3 > 0,
                "bgemm: empty problem (G %dx%d, M %d, N %d, K %dx%dx%d)", d->G1, d->G2, d->M, d->N, d->K1, d->K2, d->K3);
  FMM_CHECK_ARG(d->dtype == FMM_DT_BF16 || d->dtype == FMM_DT_F32, "bgemm: bad operand dtype %d", d->dtype);
  FMM_CHECK_ARG(d->c_dtype == FMM_DT_BF16 || d->c_dtype == FMM_DT_F32, "bgemm: bad output dtype %d", d->c_dtype);
  FMM_CHECK_ARG(d->splitk >= 1 && d->splitk <= 65535, "bgemm: bad splitk %d", d->splitk);
  FMM_CHECK_ARG(d->splitk == 1 || (d->c_dtype == FMM_DT_F32 && !d->bias_m && !d->bias_n && d->act == 0),
                "bgemm: split-K needs an fp32 C and no bias/activation");
  FMM_CHECK_ARG(d->act == 0 || d->act == 1, "bgemm: bad act %d", d->act);
  long long K = (long long)d->K1 * d->K2 * d->K3;
  FMM_CHECK_ARG(K < (1ll << 31), "bgemm: contraction too long");
  int a_mode = pick_mode(d->A, d->dtype, d->a_m, d->a_g1, d->a_g2, d->a_k1, d->a_k2, d->a_k3, d->M, d->K1, d->K2, d->K3);
  int b_mode = pick_mode(d->B, d->dtype, d->b_n, d->b_g1, d->b_g2, d->b_k1, d->b_k2, d->b_k3, d->N, d->K1, d->K2, d->K3);
  const bool pipe = d->dtype == FMM_DT_BF16 && a_mode >= 2 && b_mode >= 2;
  const int bm = pipe ? pBM : kBM, bn = pipe ? pBN : kBN;
  long long tiles = (long long)((d->M + bm - 1) / bm) * ((d->N + bn - 1) / bn) * d->G1 * d->G2;
  FMM_CHECK_ARG(tiles < (1ll << 31), "bgemm: too many tiles");
  dim3 grid((unsigned)tiles, (unsigned)d->splitk);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pipe) {
    const long long cal = d->c_dtype == FMM_DT_F32 ? 3 : 7;  // elements per 16 bytes, minus one
    const int vec_epi = d->splitk == 1 && d->c_n == 1 && (d->c_m & cal) == 0 && (d->c_g1 & cal) == 0 &&
                        (d->c_g2 & cal) == 0 && (reinterpret_cast<uintptr_t>(d->C) & 15) == 0;
    if (a_mode == 2 && b_mode == 2) bgemm_pipe_kernel<2, 2><<<grid, pThreads, 0, st>>>(*d, vec_epi);
    else if (a_mode == 2) bgemm_pipe_kernel<2, 3><<<grid, pThreads, 0, st>>>(*d, vec_epi);
    else if (b_mode == 2) bgemm_pipe_kernel<3, 2><<<grid, pThreads, 0, st>>>(*d, vec_epi);
    else bgemm_pipe_kernel<3, 3><<<grid, pThreads, 0, st>>>(*d, vec_epi);
  } else if (d->dtype == FMM_DT_BF16) {
    bgemm_kernel<__nv_bfloat16, 1><<<grid, kThreads, 0, st>>>(*d, a_mode, b_mode);
  } else {
    bgemm_kernel<float, 3><<<grid, kThreads, 0, st>>>(*d, a_mode, b_mode);
  }
  FMM_CHECK_LAUNCH("bgemm");
  return FMM_OK;
}
