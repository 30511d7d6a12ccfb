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

  // epilogue
  const long long cbase = g1 * p.c_g1 + g2 * p.c_g2;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int m = m0 + wm + mi * 16 + gq + (q >> 1) * 8;
        int n = n0 + wn + ni * 8 + 2 * tq + (q & 1);
        if (m >= p.M || n >= p.N) continue;
        float v = acc[mi][ni][q];
        if constexpr (kParts > 1) v += cor[mi][ni][q];
        v *= p.alpha;
        long long off = cbase + m * p.c_m + n * p.c_n;
        if (p.splitk > 1) {
          atomicAdd(reinterpret_cast<float*>(p.C) + off, v);
          continue;
        }
        if (p.bias_m) v += p.bias_m[m];
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
  long long tiles = (long long)((d->M + kBM - 1) / kBM) * ((d->N + kBN - 1) / kBN) * d->G1 * d->G2;
  FMM_CHECK_ARG(tiles < (1ll << 31), "bgemm: too many tiles");
  int a_mode = pick_mode(d->A, d->dtype, d->a_m, d->a_g1, d->a_g2, d->a_k1, d->a_k2, d->a_k3, d->M, d->K1, d->K2, d->K3);
  int b_mode = pick_mode(d->B, d->dtype, d->b_n, d->b_g1, d->b_g2, d->b_k1, d->b_k2, d->b_k3, d->N, d->K1, d->K2, d->K3);
  dim3 grid((unsigned)tiles, (unsigned)d->splitk);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->dtype == FMM_DT_BF16)
    bgemm_kernel<__nv_bfloat16, 1><<<grid, kThreads, 0, st>>>(*d, a_mode, b_mode);
  else
    bgemm_kernel<float, 3><<<grid, kThreads, 0, st>>>(*d, a_mode, b_mode);
  FMM_CHECK_LAUNCH("bgemm");
  return FMM_OK;
}
