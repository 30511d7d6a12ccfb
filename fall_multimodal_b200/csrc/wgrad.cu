// wgrad: weight-gradient GEMMs of the ST-GCN blocks on tcgen05 tensor cores.
//
//   dW[m][ci][co] += sum_{n,v,j}  f(X[n, j*is+shift[m], v, ci]) * dY[n, j, v, co]
//
// i.e. the contraction runs over the (huge) row dimension and the result is a small matrix, the
// mirror image of csrc/tapconv.cu. Both operands are activations, read channels-last, and both are
// staged in the SAME shared-memory images as the forward engine ([time step][8 columns][64 ch],
// SWIZZLE_128B): what is a K-major A tile for the forward GEMM is an MN-major operand here
// (M' = channels, K' = rows), so taps are again whole-atom descriptor shifts and stride is SBO.
//
// Serves: tcn 9x1 conv wgrad (BN+ReLU of the saved pre-activation fused in the prologue, as in
// stgcan.py:112-118), the 1x1 gcn conv wgrad on the aggregated input (:42-54) and the strided
// residual 1x1 conv wgrad (:128-131).
//
// Work split: item = (ci tile, tap group, co tile) keeps TG accumulators [128 x BN] resident in
// TMEM; the row dimension is sliced across CTAs and the partial results are added to global
// memory with fp32 atomics (dW must be zeroed by the caller).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "producer.cuh"
#include "ptx.cuh"

namespace fmm {

struct WgradParams {
  const void* x;
  const void* dy;
  float* dw;
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  int N, V, Tin, Tj, Cin, Cout;
  int istride;
  int ntaps;
  int shift[9];
  int minshift, win_atoms;
  int JT;                 // positions per unit (16 or 8) -> K' = 8*JT rows
  int MCH;                // 64-channel chunks per ci tile (1 or 2)
  int BN, TG;             // co tile width, accumulators (taps, or tap pairs) per item
  int pair;               // 1: Cin <= 64, one M=128 MMA covers taps (m, m+1): rows 64..127 read the window one atom later
  int nacc_total;         // accumulators over all items: ntaps, or ceil(ntaps/2) in pair mode
  int ci_tiles, tap_groups, co_tiles, items, slices;
  int ncols, ngroups, njchunks, total_units;
  int nstages;
  int C2;                 // ci = c1*C2 + c2
  long long s_m, s_c1, s_c2, s_co;
  unsigned* err;
};

constexpr int kWgThreads = 288;  // warps 0-7 producers (0-3 also run the epilogue), 8 = MMA issuer
constexpr int kWgProducers = 256;

// kTaps only names the launch class for profilers (multi-tap: tensor bound; 1x1: HBM bound); the code is identical.
template <typename T, bool kTaps>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int kParts = ActTraits<T>::kParts;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_part = static_cast<uint32_t>(p.MCH) * p.win_atoms * 1024u;
  const uint32_t b_part = static_cast<uint32_t>(p.BN / 64) * p.JT * 1024u;
  const uint32_t a_bytes = a_part * kParts;
  const uint32_t b_bytes = b_part * kParts;
  const uint32_t stage_bytes = a_bytes + b_bytes + 1024u;  // +1 atom of slack: MCH=1 reads one atom past A
  const uint32_t bars0 = smem_base + p.nstages * stage_bytes;
  auto full = [&](int s) { return bars0 + 8u * s; };
  auto empty = [&](int s) { return bars0 + 8u * (p.nstages + s); };
  const uint32_t acc_full = bars0 + 8u * (2 * p.nstages);
  const uint32_t tmem_slot = acc_full + 8u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // fp32 mode: main (hi*hi) and correction terms in separate TMEM tiles, see csrc/tapconv.cu
  const uint32_t acc_stride = (kParts == 1 ? 1u : 2u) * static_cast<uint32_t>(p.BN);
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.TG) * acc_stride) tmem_cols <<= 1;

  const int item = blockIdx.x % p.items;
  const int slice = blockIdx.x / p.items;
  const int co_tile = item % p.co_tiles;
  const int tg = (item / p.co_tiles) % p.tap_groups;
  const int ci_tile = item / (p.co_tiles * p.tap_groups);
  const int acc0 = tg * p.TG;                                                    // first accumulator of this item
  const int mt = (p.nacc_total - acc0) < p.TG ? (p.nacc_total - acc0) : p.TG;   // accumulators of this item
  const int m0 = p.pair ? 2 * acc0 : acc0;                                       // first tap of this item
  const int tstep = p.pair ? 2 : 1;
  const int my_units = slice < p.total_units ? (p.total_units - slice + p.slices - 1) / p.slices : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(full(s), kWgProducers);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (my_units > 0) {
    if (warp < 8) {
      // ------------------------------ producers ------------------------------
      const T* __restrict__ X = reinterpret_cast<const T*>(p.x);
      const T* __restrict__ DY = reinterpret_cast<const T*>(p.dy);
      const int pt = threadIdx.x;
      const int pc = pt & 7;
      const int q = (pt >> 3) & 7;
      const int a0 = pt >> 6;
      const bool xvec = (p.Cin % 8) == 0;
      const bool yvec = (p.Cout % 8) == 0;
      const bool affine = p.in_scale != nullptr;
      const size_t xpitch = static_cast<size_t>(p.V) * p.Cin;
      const size_t ypitch = static_cast<size_t>(p.V) * p.Cout;
      const float one[8] = {1, 1, 1, 1, 1, 1, 1, 1}, zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const bool use_cpasync = (kParts == 1) && xvec && yvec;
      if (use_cpasync) {
        // ---- cp.async software pipeline over the units of this CTA (D = nstages-1 units in flight) ----
        const __nv_bfloat16* __restrict__ Xb = reinterpret_cast<const __nv_bfloat16*>(p.x);
        const __nv_bfloat16* __restrict__ Yb = reinterpret_cast<const __nv_bfloat16*>(p.dy);
        const int D = p.nstages - 1;
        int i_u = slice, i_st = 0;
        uint32_t i_ph = 0;
        auto issue_one = [&]() {
          if (i_u < p.total_units) {
            const int jchunk = i_u % p.njchunks;
            const int group = i_u / p.njchunks;
            const int col = group * 8 + q;
            const bool col_ok = col < p.ncols;
            const int n = col_ok ? col / p.V : 0;
            const int v = col_ok ? col % p.V : 0;
            const int j0 = jchunk * p.JT;
            const int t_lo = j0 * p.istride + p.minshift;
            mbar_wait(empty(i_st), i_ph ^ 1u, p.err, 11);
            const uint32_t a_base = smem_base + i_st * stage_bytes;
            const uint32_t b_base = a_base + a_bytes;
            const __nv_bfloat16* xcol = Xb + static_cast<size_t>(n) * p.Tin * xpitch + static_cast<size_t>(v) * p.Cin;
            const __nv_bfloat16* ycol = Yb + static_cast<size_t>(n) * p.Tj * ypitch + static_cast<size_t>(v) * p.Cout;
            for (int h = 0; h < p.MCH; ++h) {
              const int cb = (ci_tile * p.MCH + h) * 64 + pc * 8;
              const bool ok = col_ok && cb < p.Cin;
              cpasync_issue_chunk(xcol + (ok ? cb : 0), xpitch, ok, p.Tin, t_lo, p.win_atoms, a0, 4,
                                  a_base + h * p.win_atoms * 1024u + q * 128u + ((pc ^ q) << 4));
            }
            for (int h = 0; h < p.BN / 64; ++h) {
              const int cb = co_tile * p.BN + h * 64 + pc * 8;
              const bool ok = col_ok && cb < p.Cout;
              cpasync_issue_chunk(ycol + (ok ? cb : 0), ypitch, ok, p.Tj, j0, p.JT, a0, 4,
                                  b_base + h * p.JT * 1024u + q * 128u + ((pc ^ q) << 4));
            }
            if (++i_st == p.nstages) {
              i_st = 0;
              i_ph ^= 1u;
            }
            i_u += p.slices;
          }
          cp_async_commit();
        };
        for (int d = 0; d < (D > 0 ? D : 1); ++d) issue_one();
        int st = 0;
        for (int u = slice; u < p.total_units; u += p.slices) {
          cp_async_wait_dyn(D > 0 ? D - 1 : 0);  // finish unit u first, queue unit +D afterwards (see tapconv.cu)
          if (affine) {
            const int jchunk = u % p.njchunks;
            const int group = u / p.njchunks;
            const int col = group * 8 + q;
            const int t_lo = jchunk * p.JT * p.istride + p.minshift;
            const uint32_t a_base = smem_base + st * stage_bytes;
            for (int h = 0; h < p.MCH; ++h) {
              const int cb = (ci_tile * p.MCH + h) * 64 + pc * 8;
              float sc[8], sh[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const bool ok = (cb + i) < p.Cin;
                sc[i] = ok ? p.in_scale[cb + i] : 1.f;
                sh[i] = ok ? p.in_shift[cb + i] : 0.f;
              }
              inplace_affine_chunk(col < p.ncols && cb < p.Cin, p.Tin, t_lo, p.win_atoms, a0, 4, sc, sh, p.in_relu != 0,
                                   a_base + h * p.win_atoms * 1024u + q * 128u + ((pc ^ q) << 4));
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(full(st));
          if (++st == p.nstages) st = 0;
          issue_one();
        }
      } else {
      int st = 0;
      uint32_t ph = 0;
      for (int u = slice; u < p.total_units; u += p.slices) {
        const int jchunk = u % p.njchunks;
        const int group = u / p.njchunks;
        const int col = group * 8 + q;
        const bool col_ok = col < p.ncols;
        const int n = col_ok ? col / p.V : 0;
        const int v = col_ok ? col % p.V : 0;
        const int j0 = jchunk * p.JT;
        const int t_lo = j0 * p.istride + p.minshift;
        mbar_wait(empty(st), ph ^ 1u, p.err, 11);
        const uint32_t a_base = smem_base + st * stage_bytes;
        const uint32_t b_base = a_base + a_bytes;
        const T* xcol = X + static_cast<size_t>(n) * p.Tin * xpitch + static_cast<size_t>(v) * p.Cin;
        const T* ycol = DY + static_cast<size_t>(n) * p.Tj * ypitch + static_cast<size_t>(v) * p.Cout;
        // A' : X window, MCH chunks
        for (int h = 0; h < p.MCH; ++h) {
          const int cb = (ci_tile * p.MCH + h) * 64 + pc * 8;
          float sc[8], sh[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = affine && (cb + i) < p.Cin;
            sc[i] = ok ? p.in_scale[cb + i] : 1.f;
            sh[i] = ok ? p.in_shift[cb + i] : 0.f;
          }
          const uint32_t sdst = a_base + h * p.win_atoms * 1024u + q * 128u + ((pc ^ q) << 4);
          if (affine)
            produce_chunk<T, kParts, true>(xcol + cb, xpitch, col_ok, p.Tin, t_lo, p.win_atoms, a0, 4, p.Cin - cb, xvec, sc,
                                           sh, p.in_relu != 0, sdst, a_part);
          else
            produce_chunk<T, kParts, false>(xcol + cb, xpitch, col_ok, p.Tin, t_lo, p.win_atoms, a0, 4, p.Cin - cb, xvec,
                                            sc, sh, false, sdst, a_part);
        }
        // B' : dY tile, BN/64 chunks x JT atoms
        for (int h = 0; h < p.BN / 64; ++h) {
          const int cb = co_tile * p.BN + h * 64 + pc * 8;
          const uint32_t sdst = b_base + h * p.JT * 1024u + q * 128u + ((pc ^ q) << 4);
          produce_chunk<T, kParts, false>(ycol + cb, ypitch, col_ok, p.Tj, j0, p.JT, a0, 4, p.Cout - cb, yvec, one, zero,
                                          false, sdst, b_part);
        }
        fence_proxy_async_smem();
        mbar_arrive(full(st));
        if (++st == p.nstages) {
          st = 0;
          ph ^= 1u;
        }
      }
      }
    }
    if (warp < 4) {
      // ------------------------------ epilogue (same warps) ------------------------------
      mbar_wait(acc_full, 0, p.err, 12);
      tc_fence_after();
      const int row = warp * 32 + lane;  // channel inside the ci tile (pair mode: rows 64..127 = next tap)
      const int second = (p.pair && row >= 64) ? 1 : 0;
      const int ci = ci_tile * p.MCH * 64 + (second ? row - 64 : row);
      const bool ci_ok = (p.pair || row < p.MCH * 64) && ci < p.Cin;
      const long long ci_off = ci_ok ? (ci / p.C2) * p.s_c1 + (ci % p.C2) * p.s_c2 : 0;
      for (int t = 0; t < mt; ++t) {
        for (int cg = 0; cg < p.BN / 32; ++cg) {
          uint32_t acc[32];
          const uint32_t taddr = tmem_base + static_cast<uint32_t>(t) * acc_stride + static_cast<uint32_t>(cg * 32) +
                                 (static_cast<uint32_t>(warp * 32) << 16);
          tmem_ld32(taddr, acc);
          tmem_ld_wait();
          if (kParts != 1) {
            uint32_t corr[32];
            tmem_ld32(taddr + p.BN, corr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(corr[i]));
          }
          if (ci_ok && m0 + t * tstep + second < p.ntaps) {
            float* dst = p.dw + (m0 + t * tstep + second) * p.s_m + ci_off;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = co_tile * p.BN + cg * 32 + i;
              if (co < p.Cout) atomicAdd(dst + co * p.s_co, __uint_as_float(acc[i]));
            }
          }
        }
      }
    } else if (warp == 8) {
      // ---------------------------------- MMA issuer (warp converged, elect_one regions) ---------------
      {
        const uint32_t idesc = make_idesc_bf16(p.BN, 1, 1);
        const uint32_t a_sbo = static_cast<uint32_t>(p.istride) * 1024u;
        const uint32_t a_lbo = p.MCH == 2 ? static_cast<uint32_t>(p.win_atoms) * 1024u : 1024u;
        const uint32_t b_lbo = static_cast<uint32_t>(p.JT) * 1024u;
        const uint32_t a_hi = desc_hi(a_sbo), b_hi = desc_hi(1024);
        const uint32_t a_kstep = (2u * a_sbo) >> 4, b_kstep = 2048u >> 4;
        const uint32_t pa_lo = a_part >> 4, pb_lo = b_part >> 4;
        int st = 0;
        uint32_t ph = 0;
        uint32_t accum = 0;
        for (int it = 0; it < my_units; ++it) {
          mbar_wait(full(st), ph, p.err, 13);
          tc_fence_after();
          const uint32_t a_base = smem_base + st * stage_bytes;
          const uint32_t b_lo0 = desc_lo(a_base + a_bytes, b_lbo);
          if (elect_one()) {
            for (int t = 0; t < mt; ++t) {
              const uint32_t a_lo0 = desc_lo(a_base + static_cast<uint32_t>(p.shift[m0 + t * tstep] - p.minshift) * 1024u, a_lbo);
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(t) * acc_stride;
              for (int kk = 0; kk < p.JT / 2; ++kk) {
                const uint32_t a_lo = a_lo0 + kk * a_kstep, b_lo = b_lo0 + kk * b_kstep;
                const uint32_t acc = accum | static_cast<uint32_t>(kk);
                if (kParts == 1) {
                  umma_bf16_lh(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                } else {
                  const uint32_t pa[5] = {2, 0, 1, 1, 0};
                  const uint32_t pb[5] = {0, 2, 1, 0, 1};
#pragma unroll
                  for (int e = 0; e < 5; ++e)
                    umma_bf16_lh(d_tmem + p.BN, a_lo + pa[e] * pa_lo, a_hi, b_lo + pb[e] * pb_lo, b_hi, idesc,
                                 acc | static_cast<uint32_t>(e));
                  umma_bf16_lh(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                }
              }
            }
            umma_commit(empty(st));
            if (it == my_units - 1) umma_commit(acc_full);
          }
          __syncwarp();
          accum = 1;
          if (++st == p.nstages) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// wgrad with tensor-map TMA operand stages (bf16, no fused prologue: the engine's path).
//
// The cp.async producers above need 5 120 LDGSTS (20 per thread, ~1 300 issue cycles) per 81 KB stage and only two stages fit:
// the MMA warp waits on operand stages ~40 % of the kernel at C >= 128 (wait-site profile, profiles/r02_summary.md). Here the
// column groups are CLIP-ALIGNED (ceil(V/8) groups per clip, the last one padded with zero columns: 33 of 40 rows carry data at
// V = 33), which makes every operand tile one box of a 4-D tensor map over [n][t][v][c]: {64 ch, 8 joints, W frames, 1 clip},
// out-of-range frames (the conv's zero padding) and joints are zero-filled by the TMA engine. One lane issues MCH + BN/64 boxes
// per stage; the window lands in exactly the [frame][8 joints][64 ch] SWIZZLE_128B image the MMA descriptors expect.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

template <bool kTaps>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tma_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = static_cast<uint32_t>(p.MCH) * p.win_atoms * 1024u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN / 64) * p.JT * 1024u;
  const uint32_t stage_bytes = a_bytes + b_bytes + 1024u;  // +1 atom of slack: MCH=1 reads one atom past A
  const uint32_t bars0 = smem_base + p.nstages * stage_bytes;
  auto full = [&](int s) { return bars0 + 8u * s; };
  auto empty = [&](int s) { return bars0 + 8u * (p.nstages + s); };
  const uint32_t acc_full = bars0 + 8u * (2 * p.nstages);
  const uint32_t tmem_slot = acc_full + 8u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t acc_stride = static_cast<uint32_t>(p.BN);
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.TG) * acc_stride) tmem_cols <<= 1;

  const int item = blockIdx.x % p.items;
  const int slice = blockIdx.x / p.items;
  const int co_tile = item % p.co_tiles;
  const int tg = (item / p.co_tiles) % p.tap_groups;
  const int ci_tile = item / (p.co_tiles * p.tap_groups);
  const int acc0 = tg * p.TG;
  const int mt = (p.nacc_total - acc0) < p.TG ? (p.nacc_total - acc0) : p.TG;
  const int m0 = p.pair ? 2 * acc0 : acc0;
  const int tstep = p.pair ? 2 : 1;
  const int my_units = slice < p.total_units ? (p.total_units - slice + p.slices - 1) / p.slices : 0;
  const int gpc = (p.V + 7) / 8;   // column groups per clip

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_dy)) : "memory");
  }
  // the slack atom behind every A image is read by the tap-pair MMAs (rows 64..127 of the last tap pair): keep it finite
  for (int s = threadIdx.x; s < p.nstages * 64; s += blockDim.x) {
    const uint32_t addr = smem_base + (s / 64) * stage_bytes + a_bytes + b_bytes + (s % 64) * 16u;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (my_units > 0) {
    if (warp == 7 && lane == 0) {
      // ------------------------------ TMA producer (one lane) ------------------------------
      int st = 0;
      uint32_t ph = 0;
      for (int u = slice; u < p.total_units; u += p.slices) {
        const int jchunk = u % p.njchunks;
        const int group = u / p.njchunks;
        const int n = group / gpc, v0 = (group - n * gpc) * 8;
        const int j0 = jchunk * p.JT;
        const int t_lo = j0 * p.istride + p.minshift;
        mbar_wait_relaxed(empty(st), ph ^ 1u, p.err, 11, 32);
        const uint32_t a_base = smem_base + st * stage_bytes;
        mbar_arrive_expect_tx(full(st), a_bytes + b_bytes);
        for (int h = 0; h < p.MCH; ++h)
          tma_load_4d(a_base + h * p.win_atoms * 1024u, &tm_x, (ci_tile * p.MCH + h) * 64, v0, t_lo, n, full(st));
        for (int h = 0; h < p.BN / 64; ++h)
          tma_load_4d(a_base + a_bytes + h * p.JT * 1024u, &tm_dy, co_tile * p.BN + h * 64, v0, j0, n, full(st));
        if (++st == p.nstages) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
    if (warp < 4) {
      // ------------------------------ epilogue ------------------------------
      mbar_wait_relaxed(acc_full, 0, p.err, 12, 256);
      tc_fence_after();
      const int row = warp * 32 + lane;  // channel inside the ci tile (pair mode: rows 64..127 = next tap)
      const int second = (p.pair && row >= 64) ? 1 : 0;
      const int ci = ci_tile * p.MCH * 64 + (second ? row - 64 : row);
      const bool ci_ok = (p.pair || row < p.MCH * 64) && ci < p.Cin;
      const long long ci_off = ci_ok ? (ci / p.C2) * p.s_c1 + (ci % p.C2) * p.s_c2 : 0;
      for (int t = 0; t < mt; ++t) {
        for (int cg = 0; cg < p.BN / 32; ++cg) {
          uint32_t acc[32];
          const uint32_t taddr = tmem_base + static_cast<uint32_t>(t) * acc_stride + static_cast<uint32_t>(cg * 32) +
                                 (static_cast<uint32_t>(warp * 32) << 16);
          tmem_ld32(taddr, acc);
          tmem_ld_wait();
          if (ci_ok && m0 + t * tstep + second < p.ntaps) {
            float* dst = p.dw + (m0 + t * tstep + second) * p.s_m + ci_off;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = co_tile * p.BN + cg * 32 + i;
              if (co < p.Cout) atomicAdd(dst + co * p.s_co, __uint_as_float(acc[i]));
            }
          }
        }
      }
    } else if (warp == 8) {
      // ---------------------------------- MMA issuer ----------------------------------
      const uint32_t idesc = make_idesc_bf16(p.BN, 1, 1);
      const uint32_t a_sbo = static_cast<uint32_t>(p.istride) * 1024u;
      const uint32_t a_lbo = p.MCH == 2 ? static_cast<uint32_t>(p.win_atoms) * 1024u : 1024u;
      const uint32_t b_lbo = static_cast<uint32_t>(p.JT) * 1024u;
      const uint32_t a_hi = desc_hi(a_sbo), b_hi = desc_hi(1024);
      const uint32_t a_kstep = (2u * a_sbo) >> 4, b_kstep = 2048u >> 4;
      int st = 0;
      uint32_t ph = 0;
      uint32_t accum = 0;
      for (int it = 0; it < my_units; ++it) {
        mbar_wait(full(st), ph, p.err, 13);
        tc_fence_after();
        const uint32_t a_base = smem_base + st * stage_bytes;
        const uint32_t b_lo0 = desc_lo(a_base + a_bytes, b_lbo);
        if (elect_one()) {
          for (int t = 0; t < mt; ++t) {
            const uint32_t a_lo0 = desc_lo(a_base + static_cast<uint32_t>(p.shift[m0 + t * tstep] - p.minshift) * 1024u, a_lbo);
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(t) * acc_stride;
            for (int kk = 0; kk < p.JT / 2; ++kk)
              umma_bf16_lh(d_tmem, a_lo0 + kk * a_kstep, a_hi, b_lo0 + kk * b_kstep, b_hi, idesc, accum | static_cast<uint32_t>(kk));
          }
          umma_commit(empty(st));
          if (it == my_units - 1) umma_commit(acc_full);
        }
        __syncwarp();
        accum = 1;
        if (++st == p.nstages) {
          st = 0;
          ph ^= 1u;
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace fmm

using namespace fmm;

namespace {
typedef CUresult (*WgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
WgEncodeTiledFn wg_encode_fn() {
  static WgEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WgEncodeTiledFn>(ptr);
  }
  return fn;
}
// bf16 activation [N][T][V][C] (channels-last), box {64 channels, 8 joints, frames, 1 clip}, SWIZZLE_128B, zero fill out of range
bool make_tmap_act(CUtensorMap* map, const void* base, int N, int T, int V, int C, int box_frames) {
  WgEncodeTiledFn fn = wg_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(V), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(V) * C * 2, static_cast<cuuint64_t>(T) * V * C * 2};
  cuuint32_t box[4] = {64, 8, static_cast<cuuint32_t>(box_frames), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" {

int fmm_wgrad(const void* x, const void* dy, float* dw, const float* in_scale, const float* in_shift,
              int in_relu, int N, int V, int Tin, int Tj, int Cin, int Cout, int istride, int ntaps,
              const int* shifts, int c2, long long s_m, long long s_c1, long long s_c2, long long s_co,
              int dtype, unsigned* err, cudaStream_t stream) {
  FMM_CHECK_ARG(x && dy && dw, "wgrad: null pointer");
  FMM_CHECK_ARG(N > 0 && V > 0 && Tin > 0 && Tj > 0 && Cin > 0 && Cout > 0 && c2 > 0, "wgrad: bad shape");
  FMM_CHECK_ARG(ntaps >= 1 && ntaps <= 9 && istride >= 1 && istride <= 2, "wgrad: bad taps/stride");
  FMM_CHECK_ARG(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "wgrad: bad dtype");
  WgradParams p;
  p.x = x;
  p.dy = dy;
  p.dw = dw;
  p.in_scale = in_scale;
  p.in_shift = in_shift;
  p.in_relu = in_relu;
  p.N = N;
  p.V = V;
  p.Tin = Tin;
  p.Tj = Tj;
  p.Cin = Cin;
  p.Cout = Cout;
  p.istride = istride;
  p.ntaps = ntaps;
  int mn = shifts[0], mx = shifts[0];
  for (int i = 0; i < 9; ++i) {
    p.shift[i] = i < ntaps ? shifts[i] : 0;
    if (i < ntaps) {
      mn = shifts[i] < mn ? shifts[i] : mn;
      mx = shifts[i] > mx ? shifts[i] : mx;
    }
  }
  const int nparts = dtype == FMM_DT_F32 ? 3 : 1;
  const int cout64 = (Cout + 63) / 64 * 64;
  if (nparts == 1) {
    static const int kWgJt = getenv("FMM_WG_JT") ? atoi(getenv("FMM_WG_JT")) : 16;
    p.JT = (ntaps > 1 && Cin > 64) ? kWgJt : 16;
    p.MCH = Cin > 64 ? 2 : 1;
    // multi-tap: the accumulators of a tap group share TMEM (TG * BN <= 512). N = 64 MMAs run at ~57 cycles (shared-memory
    // bound, floor 32); N = 128 runs at its 64-cycle floor, so wide layers take 128-column tiles (4 taps per group)
    static const int kWgTapBn = getenv("FMM_WG_TAP_BN") ? atoi(getenv("FMM_WG_TAP_BN")) : 128;
    p.BN = ntaps > 1 ? (cout64 >= 128 && p.MCH == 2 ? kWgTapBn : 64) : (cout64 < 256 ? cout64 : 256);
  } else {
    p.JT = 8;
    p.MCH = 1;
    p.BN = ntaps > 1 ? 64 : (cout64 < 128 ? cout64 : 128);
  }
  bool consecutive = ntaps > 1;
  for (int i = 1; i < ntaps; ++i) consecutive = consecutive && (shifts[i] == shifts[i - 1] + 1);
  p.pair = (nparts == 1 && p.MCH == 1 && consecutive) ? 1 : 0;
  p.nacc_total = p.pair ? (ntaps + 1) / 2 : ntaps;
  const int max_tg = 512 / (p.BN * (nparts == 1 ? 1 : 2));
  const int ngrp = (p.nacc_total + max_tg - 1) / max_tg;
  p.TG = (p.nacc_total + ngrp - 1) / ngrp;
  p.tap_groups = (p.nacc_total + p.TG - 1) / p.TG;
  p.minshift = mn;
  p.win_atoms = (p.JT - 1) * istride + (mx - mn) + 1;
  p.ci_tiles = (Cin + 64 * p.MCH - 1) / (64 * p.MCH);
  p.co_tiles = (Cout + p.BN - 1) / p.BN;
  p.items = p.ci_tiles * p.tap_groups * p.co_tiles;
  p.ncols = N * V;
  p.ngroups = (p.ncols + 7) / 8;
  p.njchunks = (Tj + p.JT - 1) / p.JT;
  p.total_units = p.ngroups * p.njchunks;
  int slices = num_sms() / p.items;
  if (slices < 1) slices = 1;
  if (slices > p.total_units) slices = p.total_units;
  p.slices = slices;
  p.C2 = c2;
  p.s_m = s_m;
  p.s_c1 = s_c1;
  p.s_c2 = s_c2;
  p.s_co = s_co;
  p.err = err;
  const size_t a_bytes = static_cast<size_t>(nparts) * p.MCH * p.win_atoms * 1024;
  const size_t b_bytes = static_cast<size_t>(nparts) * (p.BN / 64) * p.JT * 1024;
  const size_t stage = a_bytes + b_bytes + 1024;
  const size_t budget = 227 * 1024 - 1024 - 256;
  static const int kWgNsMax = getenv("FMM_WG_NS_MAX") ? atoi(getenv("FMM_WG_NS_MAX")) : 3;
  int ns = kWgNsMax;
  while (ns > 1 && ns * stage > budget) --ns;
  FMM_CHECK_ARG(ns * stage <= budget, "wgrad: stage does not fit shared memory (%zu bytes)", stage);
  p.nstages = ns;
  const size_t smem = ns * stage + 1024 + 256;
  // TMA operand stages (clip-aligned column groups) for plain bf16 launches on whole 64-channel chunks
  {
    const char* tma_str = getenv("FMM_WG_TMA");
    const int use_tma = tma_str ? atoi(tma_str) : 1;
    // (64-channel inputs keep the cp.async producers: their tap-pair N = 64 MMAs are shared-memory bound and the 21 % of
    // zero-padded columns cost more than the producers did: 78.8 us against 72.7 us at 64 -> 64, T = 64; FMM_WG_TMA=2 forces TMA)
    if (use_tma && (Cin >= 128 || use_tma >= 2) && dtype == FMM_DT_BF16 && !in_scale && Cin % 64 == 0 && Cout % 64 == 0 && Cout % p.BN == 0 &&
        p.win_atoms <= 256 && p.JT <= 256) {
      WgradParams q = p;
      const int gpc = (V + 7) / 8;
      q.ngroups = N * gpc;
      q.total_units = q.ngroups * q.njchunks;
      int sl = num_sms() / q.items;
      if (sl < 1) sl = 1;
      if (sl > q.total_units) sl = q.total_units;
      q.slices = sl;
      CUtensorMap tm_x, tm_dy;
      if (make_tmap_act(&tm_x, x, N, Tin, V, Cin, q.win_atoms) && make_tmap_act(&tm_dy, dy, N, Tj, V, Cout, q.JT)) {
        cudaError_t e2;
        if (ntaps > 1) {
          e2 = cudaFuncSetAttribute(wgrad_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e2 == cudaSuccess) wgrad_tma_kernel<true><<<q.items * q.slices, kWgThreads, smem, stream>>>(q, tm_x, tm_dy);
        } else {
          e2 = cudaFuncSetAttribute(wgrad_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e2 == cudaSuccess) wgrad_tma_kernel<false><<<q.items * q.slices, kWgThreads, smem, stream>>>(q, tm_x, tm_dy);
        }
        if (e2 != cudaSuccess) {
          set_last_error("wgrad (tma): smem attribute: %s", cudaGetErrorString(e2));
          return FMM_ERR_SMEM;
        }
        FMM_CHECK_LAUNCH("wgrad (tma)");
        return FMM_OK;
      }
    }
  }
  const int grid = p.items * p.slices;
  cudaError_t e;
#define FMM_LAUNCH_WGRAD(TT, TAPS)                                                                               \
  do {                                                                                                           \
    e = cudaFuncSetAttribute(wgrad_kernel<TT, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) {                                                                                      \
      set_last_error("wgrad: smem attribute: %s", cudaGetErrorString(e));                                        \
      return FMM_ERR_SMEM;                                                                                       \
    }                                                                                                            \
    wgrad_kernel<TT, TAPS><<<grid, kWgThreads, smem, stream>>>(p);                                               \
  } while (0)
  if (dtype == FMM_DT_BF16) {
    if (ntaps > 1) FMM_LAUNCH_WGRAD(__nv_bfloat16, true); else FMM_LAUNCH_WGRAD(__nv_bfloat16, false);
  } else {
    if (ntaps > 1) FMM_LAUNCH_WGRAD(float, true); else FMM_LAUNCH_WGRAD(float, false);
  }
#undef FMM_LAUNCH_WGRAD
  FMM_CHECK_LAUNCH("wgrad");
  return FMM_OK;
}

}  // extern "C"
