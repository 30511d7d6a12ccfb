// Row-streaming per-node GEMMs for the batched work after the graph-GRU sweep (EmbGCN.py:80-86 has one weight matrix per
// joint, so every product over all (t, clip) rows is V independent tall-skinny GEMMs whose small operand fits in shared memory):
//   pn_dgrad: OUT[p][row][n][c0 + c] = sum_k IN[p][row][n][k] * W[p][n][c0 + c][k]      (input gradients of a stage)
//   pn_wgrad: dW[p][n][c][o]        = sum_row XC[p][row][n][c] * DY[p][row][n][o]        (weight + bias gradients of a stage)
// rows = T*B (10^5), K / Co in {64, 128}, Cp <= 144. One CTA owns (path, joint, row range): the joint's weights (dgrad) or the
// output tile (wgrad, fp32 registers) stay resident while 64/128-row tiles stream through a cp.async ring; bf16 mma.sync, both
// operands of the weight gradient are read row-major through ldmatrix.trans (no transposed copies). HBM bound by design:
// every activation byte is read once.
#include "common.cuh"
#include "ptx.cuh"
#include <algorithm>

namespace fmm {
namespace pn {

typedef __nv_bfloat16 bf16;
constexpr int NW = 8, NT = NW * 32;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts16z(uint32_t dst) { asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// dgrad: K = 64 * KB. Tile = 128 rows (16 per warp) x NCP columns (NCP = padded column count, multiple of 16, <= 144).
// ------------------------------------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(NT, 1) pn_dgrad_kernel(const bf16* __restrict__ in, const bf16* __restrict__ W, bf16* __restrict__ out,
                                                         long long in_path, long long w_path, long long out_path, long long rows, int V, int Cp,
                                                         int c0, int ncols, int chunks, const float* __restrict__ bias, int relu) {
  constexpr int K = 64 * KB, AS = K + 8;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int ncp = (ncols + 15) & ~15;
  const uint32_t Ws = smem_u32(smem_raw);            // [ncp][AS]: W[n][c0 + c][k]
  const uint32_t As = Ws + ncp * AS * 2;             // [2][128][AS]
  const int n = blockIdx.y, p = blockIdx.z;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  in += (size_t)p * in_path + (size_t)n * K;
  out += (size_t)p * out_path + (size_t)n * Cp + c0;
  const bf16* Wn = W + (size_t)p * w_path + ((size_t)n * Cp + c0) * K;
  for (int i = threadIdx.x; i < ncp * (K / 8); i += NT) {
    const int c = i / (K / 8), k8 = i % (K / 8);
    const uint32_t d = Ws + (c * AS + k8 * 8) * 2;
    if (c < ncols) cp16(d, Wn + (size_t)c * K + k8 * 8);
    else sts16z(d);
  }
  const long long ntiles = (rows + 127) / 128;
  const long long t0 = ntiles * blockIdx.x / chunks, t1 = ntiles * (blockIdx.x + 1) / chunks;
  auto load_tile = [&](long long tile, int stage) {
    const long long r0 = tile * 128;
    for (int i = threadIdx.x; i < 128 * (K / 8); i += NT) {
      const int r = i / (K / 8), k8 = i % (K / 8);
      const uint32_t d = As + ((stage * 128 + r) * AS + k8 * 8) * 2;
      if (r0 + r < rows) cp16(d, in + (size_t)(r0 + r) * V * K + k8 * 8);
      else sts16z(d);
    }
  };
  if (t0 < t1) load_tile(t0, 0);
  cp_commit();
  for (long long tile = t0; tile < t1; ++tile) {
    const int st = (int)((tile - t0) & 1);
    if (tile + 1 < t1) load_tile(tile + 1, st ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    const long long r0 = tile * 128 + w * 16;
    uint32_t af[K / 16][4];
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
      const int mi = lane >> 3, r = lane & 7;
      ldsm_x4(As + ((st * 128 + w * 16 + (mi & 1) * 8 + r) * AS + ks * 16 + (mi >> 1) * 8) * 2, af[ks]);
    }
    for (int np = 0; np < (ncp >> 4); ++np) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < K / 16; ++ks) {
        uint32_t bb[4];
        const int mi = lane >> 3, r = lane & 7;
        ldsm_x4(Ws + ((np * 16 + (mi >> 1) * 8 + r) * AS + ks * 16 + (mi & 1) * 8) * 2, bb);
        mma16816(acc[0], af[ks], bb[0], bb[1]);
        mma16816(acc[1], af[ks], bb[2], bb[3]);
      }
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int c = np * 16 + h2 * 8 + 2 * tq;
        if (c < ncols) {
          const float b0 = bias ? bias[c0 + c] : 0.f, b1 = bias ? bias[c0 + c + 1] : 0.f;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const long long row = r0 + g + 8 * half;
            float v0 = acc[h2][2 * half] + b0, v1 = acc[h2][2 * half + 1] + b1;
            if (relu) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            if (row < rows) *reinterpret_cast<uint32_t*>(out + (size_t)row * V * Cp + c) = pk(v0, v1);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad: output tile Cp (<= 144, 9 m-tiles) x Co (= 64 * OB) in registers, warp w owns columns [w * 8 * OB, (w + 1) * 8 * OB).
// Row tiles of 64 through a 2-stage ring. Partial sums per row chunk -> part[chunk][p][n][Cp][Co] (summed by the caller).
// ------------------------------------------------------------------------------------------------------------------
template <int OB>
__global__ void __launch_bounds__(NT, 1) pn_wgrad_kernel(const bf16* __restrict__ xc, const bf16* __restrict__ dy, float* __restrict__ part,
                                                         long long xc_path, long long dy_path, long long rows, int V, int Cp, int chunks) {
  constexpr int Co = 64 * OB, DS = Co + 8, MT = 9, RT = 64;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int XS = 144 + 8;
  const uint32_t Xs = smem_u32(smem_raw);              // [2][RT][XS]  (columns >= Cp zero)
  const uint32_t Ds = Xs + 2 * RT * XS * 2;            // [2][RT][DS]
  const int n = blockIdx.y, p = blockIdx.z;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  xc += (size_t)p * xc_path + (size_t)n * Cp;
  dy += (size_t)p * dy_path + (size_t)n * Co;
  const int c8 = Cp >> 3;
  // zero the padded columns of both stages once
  for (int i = threadIdx.x; i < 2 * RT * (18 - c8); i += NT) {
    const int r = i / (18 - c8), k8 = c8 + i % (18 - c8);
    sts16z(Xs + (r * XS + k8 * 8) * 2);
  }
  const long long ntiles = (rows + RT - 1) / RT;
  const long long t0 = ntiles * blockIdx.x / chunks, t1 = ntiles * (blockIdx.x + 1) / chunks;
  auto load_tile = [&](long long tile, int stage) {
    const long long r0 = tile * RT;
    for (int i = threadIdx.x; i < RT * c8; i += NT) {
      const int r = i / c8, k8 = i % c8;
      const uint32_t d = Xs + ((stage * RT + r) * XS + k8 * 8) * 2;
      if (r0 + r < rows) cp16(d, xc + (size_t)(r0 + r) * V * Cp + k8 * 8);
      else sts16z(d);
    }
    for (int i = threadIdx.x; i < RT * (Co / 8); i += NT) {
      const int r = i / (Co / 8), k8 = i % (Co / 8);
      const uint32_t d = Ds + ((stage * RT + r) * DS + k8 * 8) * 2;
      if (r0 + r < rows) cp16(d, dy + (size_t)(r0 + r) * V * Co + k8 * 8);
      else sts16z(d);
    }
  };
  float acc[MT][OB][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int nt = 0; nt < OB; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mi][nt][i] = 0.f;
  if (t0 < t1) load_tile(t0, 0);
  cp_commit();
  for (long long tile = t0; tile < t1; ++tile) {
    const int st = (int)((tile - t0) & 1);
    if (tile + 1 < t1) load_tile(tile + 1, st ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < RT / 16; ++ks) {
      // B fragments: DY tile [row = k][o = n] (k-major) -> this warp's 8 * OB columns
      uint32_t bb[OB][2];
      {
        const int mi = lane >> 3, r = lane & 7;
        if (OB == 2) {
          uint32_t t4[4];
          ldsm_x4_trans(Ds + ((st * RT + ks * 16 + (mi & 1) * 8 + r) * DS + w * 16 + (mi >> 1) * 8) * 2, t4);
          bb[0][0] = t4[0]; bb[0][1] = t4[1]; bb[OB - 1][0] = t4[2]; bb[OB - 1][1] = t4[3];
        } else {
          uint32_t t4[4];   // x4 over two 8-column groups; only this warp's group is used
          ldsm_x4_trans(Ds + ((st * RT + ks * 16 + (mi & 1) * 8 + r) * DS + (w >> 1) * 16 + (mi >> 1) * 8) * 2, t4);
          bb[0][0] = (w & 1) ? t4[2] : t4[0];
          bb[0][1] = (w & 1) ? t4[3] : t4[1];
        }
      }
#pragma unroll
      for (int mi9 = 0; mi9 < MT; ++mi9) {
        if (mi9 * 16 >= Cp) break;
        uint32_t a[4];
        const int mi = lane >> 3, r = lane & 7;
        // A fragment: XC tile [row = k][c = m] (k-major)
        ldsm_x4_trans(Xs + ((st * RT + ks * 16 + (mi >> 1) * 8 + r) * XS + mi9 * 16 + (mi & 1) * 8) * 2, a);
#pragma unroll
        for (int nt = 0; nt < OB; ++nt) mma16816(acc[mi9][nt], a, bb[nt][0], bb[nt][1]);
      }
    }
    __syncthreads();
  }
  float* dst = part + ((((size_t)blockIdx.x * gridDim.z + p) * V + n) * Cp) * Co;
#pragma unroll
  for (int mi9 = 0; mi9 < MT; ++mi9)
#pragma unroll
    for (int nt = 0; nt < OB; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = mi9 * 16 + g + 8 * half, o = w * 8 * OB + nt * 8 + 2 * tq;
        if (c < Cp) *reinterpret_cast<float2*>(dst + (size_t)c * Co + o) = make_float2(acc[mi9][nt][2 * half], acc[mi9][nt][2 * half + 1]);
      }
}

// ------------------------------------------------------------------------------------------------------------------
// dS partials: part[cta][n][m] = sum over the CTA's rows and all c of G[row][n][c] * X[row][m][c]   (V <= 32, Cp <= 144).
// Each of 4 warps streams its own (t, clip) rows through a private 2-stage cp.async ring: per row one (V x Cp).(Cp x V) product.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) pn_ds_kernel(const bf16* __restrict__ G, const bf16* __restrict__ X, float* __restrict__ part,
                                                       long long rows, int V, int Cp) {
  constexpr int XS = 144 + 8;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const uint32_t base = smem_u32(smem_raw) + w * (4 * 32 * XS * 2);     // [stage][G | X][32][XS]
  for (int i = lane; i < 4 * 32 * (XS / 8); i += 32) sts16z(base + i * 16);
  __syncwarp();
  const int c8 = Cp >> 3;
  const long long wid = (long long)blockIdx.x * 4 + w, nw = (long long)gridDim.x * 4;
  auto load_row = [&](long long row, int stage) {
    for (int i = lane; i < V * c8; i += 32) {
      const int n = i / c8, k8 = i % c8;
      cp16(base + (((stage * 2 + 0) * 32 + n) * XS + k8 * 8) * 2, G + ((size_t)row * V + n) * Cp + k8 * 8);
      cp16(base + (((stage * 2 + 1) * 32 + n) * XS + k8 * 8) * 2, X + ((size_t)row * V + n) * Cp + k8 * 8);
    }
  };
  float acc[2][4][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mi][nt][i] = 0.f;
  if (wid < rows) load_row(wid, 0);
  cp_commit();
  int st = 0;
  for (long long row = wid; row < rows; row += nw, st ^= 1) {
    if (row + nw < rows) load_row(row + nw, st ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncwarp();
    const uint32_t Gs = base + (st * 2 + 0) * 32 * XS * 2, Xs = base + (st * 2 + 1) * 32 * XS * 2;
    const int mi4 = lane >> 3, r = lane & 7;
    for (int ks = 0; ks < ((Cp + 15) >> 4); ++ks) {
      uint32_t a[2][4], bb[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) ldsm_x4(Gs + ((mi * 16 + (mi4 & 1) * 8 + r) * XS + ks * 16 + (mi4 >> 1) * 8) * 2, a[mi]);
#pragma unroll
      for (int np = 0; np < 2; ++np) ldsm_x4(Xs + ((np * 16 + (mi4 >> 1) * 8 + r) * XS + ks * 16 + (mi4 & 1) * 8) * 2, bb[np]);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          mma16816(acc[mi][2 * np], a[mi], bb[np][0], bb[np][1]);
          mma16816(acc[mi][2 * np + 1], a[mi], bb[np][2], bb[np][3]);
        }
    }
    __syncwarp();
  }
  float* dst = part + ((size_t)blockIdx.x * 4 + w) * 32 * 32;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half)
        *reinterpret_cast<float2*>(dst + (mi * 16 + g + 8 * half) * 32 + nt * 8 + 2 * tq) = make_float2(acc[mi][nt][2 * half], acc[mi][nt][2 * half + 1]);
}

}  // namespace pn
}  // namespace fmm

extern "C" {

// OUT[p][row][n][c0 + c] = act(sum_k IN[p][row][n][k] W[p][n][c0 + c][k] + bias[c0 + c]), c < ncols; IN (P,rows,V,K), W (P,V,Cp,K),
// OUT (P,rows,V,Cp) bf16; bias (Cp) fp32 or null (shared by all joints: the V = 1 case is a plain Linear layer), relu 0/1
int fmm_pn_dgrad(const void* in, const void* W, void* out, int P, long long rows, int V, int K, int Cp, int c0, int ncols, const float* bias,
                 int relu, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG((K == 64 || K == 128) && Cp % 8 == 0 && Cp <= 144 && c0 % 8 == 0 && ncols % 8 == 0 && ncols >= 8 && c0 + ncols <= Cp && P >= 1 && rows >= 1,
                "pn_dgrad: unsupported sizes (K=%d Cp=%d c0=%d ncols=%d)", K, Cp, c0, ncols);
  const int ncp = (ncols + 15) & ~15, AS = K + 8;
  const size_t smem = (size_t)ncp * AS * 2 + (size_t)2 * 128 * AS * 2;
  const long long ntiles = (rows + 127) / 128;
  int chunks = (int)std::min<long long>(ntiles, std::max(1, 2 * num_sms() / (V * P)));
  dim3 grid(chunks, V, P);
#define FMM_PN_DG(KB)                                                                                                          \
  do {                                                                                                                         \
    cudaError_t e = cudaFuncSetAttribute(pn::pn_dgrad_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e != cudaSuccess) {                                                                                                    \
      set_last_error("pn_dgrad: smem attribute: %s", cudaGetErrorString(e));                                                   \
      return FMM_ERR_SMEM;                                                                                                     \
    }                                                                                                                          \
    pn::pn_dgrad_kernel<KB><<<grid, pn::NT, smem, stream>>>(reinterpret_cast<const pn::bf16*>(in), reinterpret_cast<const pn::bf16*>(W), \
                                                            reinterpret_cast<pn::bf16*>(out), rows * V * K, (long long)V * Cp * K,        \
                                                            rows * V * Cp, rows, V, Cp, c0, ncols, chunks, bias, relu);                   \
  } while (0)
  if (K == 64) FMM_PN_DG(1); else FMM_PN_DG(2);
#undef FMM_PN_DG
  FMM_CHECK_LAUNCH("pn_dgrad");
  return FMM_OK;
}

// number of row chunks fmm_pn_wgrad uses (the caller allocates part[chunks][P][V][Cp][Co] fp32 and sums over chunks)
int fmm_pn_wgrad_chunks(int P, long long rows, int V) {
  const long long ntiles = (rows + 63) / 64;
  return (int)std::min<long long>(ntiles, std::max(1, 2 * fmm::num_sms() / (V * P)));
}

// part[chunk][p][n][c][o] = sum over the chunk's rows of XC[p][row][n][c] DY[p][row][n][o]; XC (P,rows,V,Cp), DY (P,rows,V,Co); bf16 in, fp32 out
int fmm_pn_wgrad(const void* xc, const void* dy, float* part, int P, long long rows, int V, int Cp, int Co, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG((Co == 64 || Co == 128) && Cp % 8 == 0 && Cp <= 144 && P >= 1 && rows >= 1, "pn_wgrad: unsupported sizes (Cp=%d Co=%d)", Cp, Co);
  const int chunks = fmm_pn_wgrad_chunks(P, rows, V);
  const size_t smem = (size_t)2 * 64 * (144 + 8) * 2 + (size_t)2 * 64 * (Co + 8) * 2;
  dim3 grid(chunks, V, P);
#define FMM_PN_WG(OB)                                                                                                          \
  do {                                                                                                                         \
    cudaError_t e = cudaFuncSetAttribute(pn::pn_wgrad_kernel<OB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e != cudaSuccess) {                                                                                                    \
      set_last_error("pn_wgrad: smem attribute: %s", cudaGetErrorString(e));                                                   \
      return FMM_ERR_SMEM;                                                                                                     \
    }                                                                                                                          \
    pn::pn_wgrad_kernel<OB><<<grid, pn::NT, smem, stream>>>(reinterpret_cast<const pn::bf16*>(xc), reinterpret_cast<const pn::bf16*>(dy), part, \
                                                            rows * V * Cp, rows * V * Co, rows, V, Cp, chunks);                                 \
  } while (0)
  if (Co == 64) FMM_PN_WG(1); else FMM_PN_WG(2);
#undef FMM_PN_WG
  FMM_CHECK_LAUNCH("pn_wgrad");
  return FMM_OK;
}

// part[i][n][m] (i < fmm_pn_ds_parts(), 32 x 32 fp32 each) = partial sums of sum_{row,c} G[row][n][c] X[row][m][c]; G, X (rows,V,Cp) bf16
int fmm_pn_ds_parts(void) { return 4 * 2 * fmm::num_sms(); }
int fmm_pn_ds(const void* G, const void* X, float* part, long long rows, int V, int Cp, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(V >= 1 && V <= 32 && Cp % 8 == 0 && Cp <= 144 && rows >= 1, "pn_ds: unsupported sizes (V=%d Cp=%d)", V, Cp);
  const size_t smem = (size_t)4 * 4 * 32 * (144 + 8) * 2;
  cudaError_t e = cudaFuncSetAttribute(pn::pn_ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("pn_ds: smem attribute: %s", cudaGetErrorString(e));
    return FMM_ERR_SMEM;
  }
  pn::pn_ds_kernel<<<2 * num_sms(), 128, smem, stream>>>(reinterpret_cast<const pn::bf16*>(G), reinterpret_cast<const pn::bf16*>(X), part, rows, V, Cp);
  FMM_CHECK_LAUNCH("pn_ds");
  return FMM_OK;
}

}  // extern "C"
