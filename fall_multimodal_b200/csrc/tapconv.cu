// tapconv: the forward/dgrad GEMM engine of the ST-GCN blocks on tcgen05 tensor cores.
//
// Computes, for channels-last activations X[(n,t,v)][Cin] -> Out[(n,t',v)][Cout]:
//     Out[n, j*os+oo, v, co] = bias[co] + sum_{m<ntaps} sum_ci  f(X[n, j*is+shift[m], v, ci]) * W[m][co][ci]
// with f(x) = relu?(x*in_scale[ci] + in_shift[ci]) and f := 0 outside [0,Tin) (zero padding is
// applied AFTER the BatchNorm+ReLU prologue, as nn.Conv2d pads its already-normalised input).
//
// One kernel serves (reference call sites, /root/reference/Fall_2_Spatial_Temporal_SR/Model/stgcan.py):
//   * the 9x1 temporal conv, stride 1|2, with the BN->ReLU of tcn[0..1] fused in the prologue (:112-118)
//   * its dgrad (transposed conv; stride 2 as two parity phases)
//   * the 1x1 graph-conv channel mix on the adjacency-aggregated input (:42-54, reassociated)
//   * the 1x1 strided residual conv (:128-131) and the dgrads of both 1x1 convs.
//
// Tiling ("column groups"): a column is one (n,v) pair = a time series; 8 consecutive columns form
// a group. An M=128 MMA tile is 16 output positions j x 8 columns. The input window of a tile is
// staged ONCE in shared memory as [time step][8 columns][64 channels] bf16 in the canonical
// SWIZZLE_128B K-major image (one 1024-byte atom per time step), so every tap is the same image
// read through a descriptor whose start address is shifted by whole atoms (no im2col copies),
// and a temporal stride is just SBO = is*1024.
//
// Warp roles (576 threads, persistent over tiles): 0-7 window producers (cp.async global -> smem,
// optional BN/ReLU in place), 8-15 epilogue (two warps per TMEM lane quadrant, alternating 32-column
// groups: TMEM -> regs -> bias -> global), 16 weight loader (bulk async copies of pre-packed weight
// images), 17 MMA issuer (one elected lane) + TMEM allocator.
#include <stdlib.h>

#include "common.cuh"
#include "producer.cuh"
#include "ptx.cuh"

namespace fmm {

struct TapConvParams {
  const void* x;
  void* out;
  const void* wpk;
  const float* in_scale;
  const float* in_shift;
  const float* bias;
  int bias_vstride;  // 0: bias[co]; Cout: bias[v][co] (per-joint bias of the reassociated graph conv)
  int in_relu;
  int N, V, Tin, Tout, Cin, Cout;
  int Tj, istride, ostride, ooff;
  int ntaps;
  int shift[9];
  int minshift, win_atoms;
  int BN, ntiles_n, nchunks;
  int ncols, ngroups, ntchunks, total_tiles;
  int nslots, nbstages;
  int MT;        // M=128 tiles per pipeline item (2: 32 positions x 8 columns share one window and every weight image)
  int tps;       // taps per streamed weight stage (one barrier round trip per `tps` taps)
  int resident;  // 1: the weight images stay in shared memory for the CTA's lifetime
  int bias_smem; // > 0: this many floats of `bias` are staged in shared memory once per CTA (the epilogue's bias loads sit on
                 // the tile's critical path: a per-joint table read from global memory cost +50 % on the 1x1 graph convs)
  int res_local; // resident && 1: only the images of the CTA's own N tile (gridDim.x is a multiple of ntiles_n)
  int mtg;       // 1: the MT M-tiles of an item are MT consecutive column GROUPS at the same 16 positions (each with its own
                 // window), not MT consecutive 16-position chunks of one group: short clips (T <= 16) still share every weight
                 // image between two M tiles
  int nacc;      // TMEM accumulator stages (2; 1 when MT * BN * 2 would exceed the 512 columns)
  int dbg;       // dev experiments (FMM_TAP_DBG): 1 no window copies, 2 no stores, 4 relaxed waits, 8 no TMEM loads
  unsigned* err;
};

constexpr int kTapMaxThreads = 576;  // 8 producer + kEpi epilogue + loader + MMA warps
constexpr int kTapProducers = 256;

// kEpi = 4 | 8 epilogue warps: wide output tiles (BN >= 128) are epilogue-bound and take two warps per
// TMEM lane quadrant; narrow tiles run with 4 (fewer warps competing with the producers for issue slots).
// kTaps only names the launch class (multi-tap temporal conv: tensor bound; 1x1 channel mix: HBM bound) so that profiler
// output can be split per class; the code is identical.
//
// kPair: the kernel runs as CTA pairs (cluster of 2 = one TPC). The pair owns two row tiles of the same N tile; every MMA is
// one `tcgen05.mma.cta_group::2` with M = 256, issued by the leader CTA (cluster rank 0): each CTA stages its own window
// and only HALF of every weight image (BN/2 rows) - half the weight fill traffic from L2 and half the shared-memory read
// bandwidth per FLOP, which is what bounds the single-CTA kernel on the wide layers (profiles/r02_summary.md). Producers and
// epilogue warps of the peer CTA arrive on the LEADER's barriers through shared::cluster addresses; the leader's commits
// are multicast to the barriers of both CTAs.
template <typename T, int kEpi, bool kTaps, bool kPair>
__global__ void __launch_bounds__((10 + kEpi) * 32, 1) tapconv_kernel(const __grid_constant__ TapConvParams p) {
  constexpr int kLoaderWarp = 8 + kEpi, kMmaWarp = 9 + kEpi;
  constexpr int kParts = ActTraits<T>::kParts;
  static_assert(!kPair || kParts == 1, "CTA pairs: bf16 activations only");
  constexpr uint32_t kCtas = kPair ? 2u : 1u;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slot_bytes = static_cast<uint32_t>(kParts) * p.win_atoms * 1024u * (p.mtg ? p.MT : 1);
  const int MTT = p.mtg ? 1 : p.MT;   // 16-position chunks per item
  const int MTG = p.mtg ? p.MT : 1;   // column groups per item
  const uint32_t part_bytes_a = static_cast<uint32_t>(p.win_atoms) * 1024u;
  const uint32_t bstage_bytes = static_cast<uint32_t>(kParts) * p.BN * 128u / kCtas;  // this CTA's rows of one weight image
  const uint32_t part_bytes_b = static_cast<uint32_t>(p.BN) * 128u / kCtas;
  const uint32_t gimg_bytes = bstage_bytes * kCtas;                                    // one whole image in global memory
  const uint32_t slots0 = smem_base;
  const uint32_t bst0 = slots0 + p.nslots * slot_bytes;
  const uint32_t bring_bytes = bstage_bytes * static_cast<uint32_t>(p.tps);  // one ring stage = tps images
  const uint32_t bars0 = bst0 + p.nbstages * bring_bytes;
  // barrier map (8 bytes each)
  auto win_full = [&](int s) { return bars0 + 8u * s; };
  auto win_empty = [&](int s) { return bars0 + 8u * (p.nslots + s); };
  auto b_full = [&](int s) { return bars0 + 8u * (2 * p.nslots + s); };
  auto b_empty = [&](int s) { return bars0 + 8u * (2 * p.nslots + p.nbstages + s); };
  auto acc_full = [&](int s) { return bars0 + 8u * (2 * p.nslots + 2 * p.nbstages + s); };
  auto acc_empty = [&](int s) { return bars0 + 8u * (2 * p.nslots + 2 * p.nbstages + 2 + s); };
  const uint32_t tmem_slot = bars0 + 8u * (2 * p.nslots + 2 * p.nbstages + 4);
  // leader only: "the peer CTA's half of weight stage s has landed" (one remote arrive per fill by the peer's forwarder)
  auto peer_b_full = [&](int s) { return bars0 + 8u * (2 * p.nslots + 2 * p.nbstages + 5 + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // fp32 mode: the dominant hi*hi products and the five small correction terms accumulate in
  // SEPARATE TMEM tiles (summed in the epilogue). The tensor core truncates its fp32 accumulator on
  // every MMA; keeping the small terms out of the big accumulator cuts those truncations 6x.
  const uint32_t acc_stride = (kParts == 1 ? 1u : 2u) * static_cast<uint32_t>(p.BN) * static_cast<uint32_t>(p.MT);
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.nacc) * acc_stride) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(win_full(s), kTapProducers * kCtas);
      mbar_init(win_empty(s), 1);
    }
    for (int s = 0; s < p.nbstages; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), kEpi * 32 * kCtas);
    }
    if (kPair)
      for (int s = 0; s < (p.resident ? 1 : p.nbstages); ++s) mbar_init(peer_b_full(s), 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    if (kPair) {
      tmem_alloc2(tmem_slot, tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, tmem_cols);
      tmem_relinquish();
    }
  }
  const uint32_t bias_s0 = bars0 + 2048u;  // behind the barrier block
  // staged as [Cout/4][rows][4]: the 8 joints of a tile read 8 neighbouring float4 (one conflict-free wavefront per load;
  // the row-major table costs 8 wavefronts per load, in shared memory as in L1)
  const int bias_rows = p.bias_vstride ? p.V : 1;
  for (int i = threadIdx.x; i < p.bias_smem; i += blockDim.x) {
    const int v = i / p.Cout, c = i % p.Cout;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s0 + 4u * (((c >> 2) * bias_rows + v) * 4 + (c & 3))), "f"(p.bias[i]) : "memory");
  }
  const float* __restrict__ bias_tab = reinterpret_cast<const float*>(smem_raw + (bias_s0 - smem_u32(smem_raw)));
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // (pair: the peer's barriers are initialised before any remote arrive)
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // tile index = rows * ntiles_n + ntile; in pair mode `rows` counts PAIRS of row tiles and this CTA takes row tile 2*rows+rank
  // (a row tile past the end decodes to columns >= ncols: zero window, no stores)
  const int first_tile = kPair ? blockIdx.x >> 1 : blockIdx.x;
  const int tile_step = kPair ? gridDim.x >> 1 : gridDim.x;
  auto rest_of = [&](int tile) { return kPair ? (tile / p.ntiles_n) * 2 + static_cast<int>(rank) : tile / p.ntiles_n; };
  // arrive on a barrier of the leader CTA (plain local arrive when the kernel is not paired)
  const uint32_t leader_bars0 = kPair ? mapa_shared(bars0, 0) : bars0;
  auto arrive_leader = [&](uint32_t bar) {
    if (kPair) mbar_arrive_cluster(leader_bars0 + (bar - bars0)); else mbar_arrive(bar);
  };

  if (warp < 8) {
    // ------------------------------ window producers ------------------------------
    const T* __restrict__ X = reinterpret_cast<const T*>(p.x);
    const int pt = threadIdx.x;       // 0..255
    const int pc = pt & 7;            // 16-byte piece = 8 channels
    const int q = (pt >> 3) & 7;      // column inside the group
    // atoms a0, a0+4, ...; a single-tap strided conv (the residual 1x1, stride 2) only ever reads every istride-th frame of its
    // window through the descriptor's SBO, so only those atoms are fetched: a0*is, (a0+4)*is, ...
    const int askip = (p.ntaps == 1) ? p.istride : 1;
    const int a0 = (pt >> 6) * askip;
    const bool vec_ok = (p.Cin % 8) == 0;
    const bool affine = p.in_scale != nullptr;
    const size_t pitch_t = static_cast<size_t>(p.V) * p.Cin;
    int slot = 0;
    uint32_t ph = 0;
    const bool use_cpasync = (kParts == 1) && vec_ok;
    if (use_cpasync) {
      // ---- cp.async software pipeline: copies of up to D = nslots-1 later items are in flight ----
      const __nv_bfloat16* __restrict__ Xb = reinterpret_cast<const __nv_bfloat16*>(p.x);
      const int D = p.nslots - 1;
      // issue-side iterator (runs D items ahead of the completion-side iterator)
      int i_tile = first_tile, i_c = 0, i_slot = 0;
      uint32_t i_ph = 0;
      // per-tile decode of the issue-side iterator, refreshed only when it moves to a new tile (the
      // integer divisions are a real cost for the 1x1 GEMMs, whose items are only 16 KB each)
      bool i_colok[2] = {false, false};
      int i_tlo = 0;
      const __nv_bfloat16* i_colp[2] = {Xb, Xb};
      auto decode_tile = [&]() {
        const int rest = rest_of(i_tile);
        const int tchunk = rest % p.ntchunks;
        const int gidx = rest / p.ntchunks;
        i_tlo = tchunk * 16 * MTT * p.istride + p.minshift;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < MTG) {
            const int col = (gidx * MTG + g) * 8 + q;
            i_colok[g] = col < p.ncols;
            const int n = i_colok[g] ? col / p.V : 0;
            const int v = i_colok[g] ? col - n * p.V : 0;
            i_colp[g] = Xb + static_cast<size_t>(n) * p.Tin * pitch_t + static_cast<size_t>(v) * p.Cin;
          }
        }
      };
      if (i_tile < p.total_tiles) decode_tile();
      auto issue_one = [&]() {
        if (i_tile < p.total_tiles) {
          const int cb = i_c * 64 + pc * 8;
          if (p.dbg & 4) mbar_wait_relaxed(win_empty(i_slot), i_ph ^ 1u, p.err, 1, 32); else
          mbar_wait(win_empty(i_slot), i_ph ^ 1u, p.err, 1);
          if (!(p.dbg & 1)) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (g < MTG) {
                const bool col_ok = i_colok[g] && cb < p.Cin;
                cpasync_issue_chunk(i_colp[g] + (col_ok ? cb : 0), pitch_t, col_ok, p.Tin, i_tlo, p.win_atoms, a0, 4 * askip,
                                    slots0 + i_slot * slot_bytes + static_cast<uint32_t>(g * p.win_atoms) * 1024u + q * 128u + ((pc ^ q) << 4));
              }
            }
          }
          if (++i_slot == p.nslots) {
            i_slot = 0;
            i_ph ^= 1u;
          }
          if (++i_c == p.nchunks) {
            i_c = 0;
            i_tile += tile_step;
            if (i_tile < p.total_tiles) decode_tile();
          }
        }
        cp_async_commit();
      };
      for (int d = 0; d < (D > 0 ? D : 1); ++d) issue_one();
      int slot = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        const int rest = rest_of(tile);
        const int tchunk = rest % p.ntchunks;
        const int group = rest / p.ntchunks;
        const int col = group * 8 + q;                    // (the in-place prologue is never combined with mtg)
        const int t_lo = tchunk * 16 * MTT * p.istride + p.minshift;
        for (int c = 0; c < p.nchunks; ++c) {
          // finish item (tile, c) FIRST and only then queue the copies of item +D: queueing needs the
          // slot the MMAs of the previous item are still reading, and waiting for it before the
          // transform would serialise the transform with the tensor core.
          cp_async_wait_dyn(D > 0 ? D - 1 : 0);  // the copies of (tile, c) have landed
          if (affine) {
            const int cb = c * 64 + pc * 8;
            const bool col_ok = col < p.ncols && cb < p.Cin;
            float sc[8], sh[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool ok = (cb + i) < p.Cin;
              sc[i] = ok ? p.in_scale[cb + i] : 1.f;
              sh[i] = ok ? p.in_shift[cb + i] : 0.f;
            }
            inplace_affine_chunk(col_ok, p.Tin, t_lo, p.win_atoms, a0, 4 * askip, sc, sh, p.in_relu != 0,
                                 slots0 + slot * slot_bytes + q * 128u + ((pc ^ q) << 4));
          }
          fence_proxy_async_smem();
          arrive_leader(win_full(slot));
          if (++slot == p.nslots) slot = 0;
          issue_one();
        }
      }
    } else {
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
      const int rest = rest_of(tile);
      const int tchunk = rest % p.ntchunks;
      const int group = rest / p.ntchunks;
      const int col = group * 8 + q;
      const bool col_ok = col < p.ncols;
      const int n = col_ok ? col / p.V : 0;
      const int v = col_ok ? col % p.V : 0;
      const int t_lo = tchunk * 16 * p.MT * p.istride + p.minshift;
      const T* colp = X + static_cast<size_t>(n) * p.Tin * pitch_t + static_cast<size_t>(v) * p.Cin;
      for (int c = 0; c < p.nchunks; ++c) {
        const int cb = c * 64 + pc * 8;
        float sc[8], sh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = affine && (cb + i) < p.Cin;
          sc[i] = ok ? p.in_scale[cb + i] : 1.f;
          sh[i] = ok ? p.in_shift[cb + i] : 0.f;
        }
        mbar_wait(win_empty(slot), ph ^ 1u, p.err, 1);
        const uint32_t sdst = slots0 + slot * slot_bytes + q * 128u + ((pc ^ q) << 4);
        if (affine)
          produce_chunk<T, kParts, true>(colp + cb, pitch_t, col_ok, p.Tin, t_lo, p.win_atoms, a0, 4 * askip, p.Cin - cb, vec_ok,
                                         sc, sh, p.in_relu != 0, sdst, part_bytes_a);
        else
          produce_chunk<T, kParts, false>(colp + cb, pitch_t, col_ok, p.Tin, t_lo, p.win_atoms, a0, 4 * askip, p.Cin - cb, vec_ok,
                                          sc, sh, false, sdst, part_bytes_a);
        fence_proxy_async_smem();
        arrive_leader(win_full(slot));
        if (++slot == p.nslots) {
          slot = 0;
          ph ^= 1u;
        }
      }
    }
    }
  } else if (warp < 8 + kEpi) {
    // ---------------------------------- epilogue ----------------------------------
    T* __restrict__ O = reinterpret_cast<T*>(p.out);
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = (warp - 8) >> 2;   // which 32-column groups it takes (even / odd when kEpi == 8)
    const int r = quad * 32 + lane;
    const int q = r & 7;
    int as = 0;
    uint32_t aph = 0;
    const bool vec_ok = (p.Cout % 8) == 0;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
      const int ntile = tile % p.ntiles_n;
      const int rest = rest_of(tile);
      const int tchunk = rest % p.ntchunks;
      const int gidx = rest / p.ntchunks;
      if (p.dbg & 4) mbar_wait_relaxed(acc_full(as), aph, p.err, 2, 64); else
      mbar_wait(acc_full(as), aph, p.err, 2);
      tc_fence_after();
      for (int mt = 0; mt < ((p.dbg & 8) ? 0 : p.MT); ++mt) {
      const int col = (p.mtg ? gidx * p.MT + mt : gidx) * 8 + q;
      const bool col_ok = col < p.ncols;
      const int n = col_ok ? col / p.V : 0;
      const int v = col_ok ? col % p.V : 0;
      const int j = (p.mtg ? tchunk : tchunk * p.MT + mt) * 16 + (r >> 3);
      const bool row_ok = col_ok && (j < p.Tj);
      T* orow = O + (static_cast<size_t>(n) * p.Tout + (row_ok ? j * p.ostride + p.ooff : 0)) * p.V * p.Cout +
                static_cast<size_t>(v) * p.Cout;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(as) * acc_stride + static_cast<uint32_t>(mt) * (acc_stride / p.MT) +
                             (static_cast<uint32_t>(quad * 32) << 16);
      for (int cg = half; cg < p.BN / 32; cg += kEpi / 4) {
        uint32_t acc[32];
        tmem_ld32(taddr + cg * 32, acc);
        tmem_ld_wait();
        if (kParts != 1) {
          uint32_t corr[32];
          tmem_ld32(taddr + p.BN + cg * 32, corr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(corr[i]));
        }
        const int co0 = ntile * p.BN + cg * 32;
        if (p.dbg & 2) {
          if (acc[0] == 0x7fc12345u) orow[0] = from_f32<T>(0.f);
        } else if (row_ok && vec_ok && co0 + 32 <= p.Cout) {
          // fast path: whole 32-column group in range
          if (p.bias) {
            const int vrow = p.bias_vstride ? v : 0;
            const float4* bp = p.bias_smem > 0
                                   ? reinterpret_cast<const float4*>(bias_tab) + (co0 >> 2) * bias_rows + vrow
                                   : reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(v) * p.bias_vstride + co0);
            const int bstep = p.bias_smem > 0 ? bias_rows : 1;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b = bp[g * bstep];
              acc[4 * g + 0] = __float_as_uint(__uint_as_float(acc[4 * g + 0]) + b.x);
              acc[4 * g + 1] = __float_as_uint(__uint_as_float(acc[4 * g + 1]) + b.y);
              acc[4 * g + 2] = __float_as_uint(__uint_as_float(acc[4 * g + 2]) + b.z);
              acc[4 * g + 3] = __float_as_uint(__uint_as_float(acc[4 * g + 3]) + b.w);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(acc[g * 8 + i]);
            store8(orow + co0 + g * 8, f);
          }
        } else if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int co = co0 + g * 8 + i;
              float b = 0.f;
              if (p.bias && co < p.Cout)
                b = p.bias_smem > 0 ? bias_tab[((co >> 2) * bias_rows + (p.bias_vstride ? v : 0)) * 4 + (co & 3)]
                                    : p.bias[v * p.bias_vstride + co];
              f[i] = __uint_as_float(acc[g * 8 + i]) + b;
            }
            const int co = co0 + g * 8;
            if (vec_ok && co + 8 <= p.Cout) {
              store8(orow + co, f);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (co + i < p.Cout) orow[co + i] = from_f32<T>(f[i]);
            }
          }
        }
      }
      }
      tc_fence_before();
      arrive_leader(acc_empty(as));
      if (++as == p.nacc) {
        as = 0;
        aph ^= 1u;
      }
    }
  } else if (warp == kLoaderWarp) {
    // -------------------------------- weight loader --------------------------------
    if (lane == 0) {
      const uint8_t* W = reinterpret_cast<const uint8_t*>(p.wpk);
      int bs = 0;
      uint32_t bph = 0;
      if (p.resident) {
        // all images once: one transaction barrier, one bulk copy per image
        if (first_tile < p.total_tiles) {
          const size_t base = (p.res_local ? static_cast<size_t>(first_tile % p.ntiles_n) * p.nbstages * gimg_bytes : 0) +
                              static_cast<size_t>(rank) * bstage_bytes;
          mbar_arrive_expect_tx(b_full(0), static_cast<uint32_t>(p.nbstages) * bstage_bytes);
          for (int i = 0; i < p.nbstages; ++i)
            bulk_g2s(bst0 + i * bstage_bytes, W + base + static_cast<size_t>(i) * gimg_bytes, bstage_bytes, b_full(0));
        }
      } else
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        const int ntile = tile % p.ntiles_n;
        for (int c = 0; c < p.nchunks; ++c) {
          for (int m = 0; m < p.ntaps; m += p.tps) {
            const uint32_t ntp = static_cast<uint32_t>(p.ntaps - m < p.tps ? p.ntaps - m : p.tps);
            if (p.dbg & 4) mbar_wait_relaxed(b_empty(bs), bph ^ 1u, p.err, 3, 32); else
            mbar_wait(b_empty(bs), bph ^ 1u, p.err, 3);
            mbar_arrive_expect_tx(b_full(bs), ntp * bstage_bytes);
            const size_t off = ((static_cast<size_t>(ntile) * p.nchunks + c) * p.ntaps + m) * gimg_bytes +
                               static_cast<size_t>(rank) * bstage_bytes;
            if (kPair) {  // this CTA's rows of each image: one copy per image
              for (uint32_t i = 0; i < ntp; ++i)
                bulk_g2s(bst0 + bs * bring_bytes + i * bstage_bytes, W + off + static_cast<size_t>(i) * gimg_bytes, bstage_bytes, b_full(bs));
            } else {
              bulk_g2s(bst0 + bs * bring_bytes, W + off, ntp * bstage_bytes, b_full(bs));
            }
            if (++bs == p.nbstages) {
              bs = 0;
              bph ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------- MMA issuer ----------------------------------
    // The warp walks the loops converged; each group of MMAs (+ its commit) is issued inside ONE
    // elect_one() region with (lo, hi) descriptor halves: ~15 instructions per MMA instead of the
    // ~40 (and a divergence waterfall) a per-thread branch costs, which matters for narrow tiles.
    auto wait_x = [&](uint32_t bar, uint32_t parity, unsigned tag) {  // barriers the peer CTA also arrives on
      if (kPair) mbar_wait_cluster(bar, parity, p.err, tag); else mbar_wait(bar, parity, p.err, tag);
    };
    auto mma = [&](uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t id, uint32_t acc) {
      if (kPair) umma2_bf16_lh(d, alo, ahi, blo, bhi, id, acc); else umma_bf16_lh(d, alo, ahi, blo, bhi, id, acc);
    };
    auto commit = [&](uint32_t bar) {
      if (kPair) umma2_commit(bar); else umma_commit(bar);
    };
    if (kPair && rank != 0) {
      // peer CTA of a pair: no MMAs to issue; forward "my half of weight stage s has landed" to the leader
      if (p.resident) {
        if (first_tile < p.total_tiles) {
          mbar_wait(b_full(0), 0, p.err, 6);
          if (lane == 0) arrive_leader(peer_b_full(0));
        }
      } else {
        int bs = 0;
        uint32_t bph = 0;
        for (int tile = first_tile; tile < p.total_tiles; tile += tile_step)
          for (int c = 0; c < p.nchunks; ++c)
            for (int m0 = 0; m0 < p.ntaps; m0 += p.tps) {
              mbar_wait_relaxed(b_full(bs), bph, p.err, 6, 32);
              if (lane == 0) arrive_leader(peer_b_full(bs));
              if (++bs == p.nbstages) {
                bs = 0;
                bph ^= 1u;
              }
            }
      }
    } else {
      const uint32_t idesc = make_idesc_bf16(p.BN, 0, 0, kPair ? 256 : 128);
      const uint32_t a_sbo = static_cast<uint32_t>(p.istride) * 1024u;
      const uint32_t a_hi = desc_hi(a_sbo), b_hi = desc_hi(1024);
      const uint32_t pa_lo = part_bytes_a >> 4, pb_lo = part_bytes_b >> 4;
      int slot = 0, bs = 0, as = 0;
      uint32_t wph = 0, bph = 0, aph = 0;
      if (p.resident && first_tile < p.total_tiles) {
        mbar_wait(b_full(0), 0, p.err, 6);
        if (kPair) mbar_wait_cluster(peer_b_full(0), 0, p.err, 7);
      }
      if (p.resident && kParts == 1) {
        // Resident weights: nothing to wait for between taps, so ALL MMAs of a (tile, chunk) item are
        // issued back to back from one elected lane (2 adds per MMA); narrow tiles are otherwise bound
        // by the issue loop, not by the tensor pipe.
        uint32_t tap_lo[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) tap_lo[m] = static_cast<uint32_t>(p.shift[m] - p.minshift) * 64u;  // atoms, 16-byte units
        const uint32_t img_lo = bstage_bytes >> 4;
        // second M tile: 16 positions later in the same window, or the next column group's window
        const uint32_t mt_lo = p.mtg ? static_cast<uint32_t>(p.win_atoms) * 64u : 16u * static_cast<uint32_t>(p.istride) * 64u;
        for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
          const int ntile = p.res_local ? 0 : tile % p.ntiles_n;  // index into the resident images
          wait_x(acc_empty(as), aph ^ 1u, 4);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * acc_stride;
          for (int c = 0; c < p.nchunks; ++c) {
            wait_x(win_full(slot), wph, 5);
            tc_fence_after();
            const uint32_t a_lo0 = desc_lo(slots0 + slot * slot_bytes, 16);
            const uint32_t b_lo0 = desc_lo(bst0, 16) + static_cast<uint32_t>((ntile * p.nchunks + c) * p.ntaps) * img_lo;
            const long long tq0 = (p.dbg & 16) ? clock64() : 0;
            if (elect_one()) {
#pragma unroll
              for (int m = 0; m < 9; ++m) {
                if (m < p.ntaps) {
                  for (uint32_t mt = 0; mt < static_cast<uint32_t>(p.MT); ++mt) {
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk)
                      mma(d_tmem + mt * p.BN, a_lo0 + tap_lo[m] + mt * mt_lo + kk * 2u, a_hi,
                                   b_lo0 + m * img_lo + kk * 2u, b_hi, idesc,
                                   static_cast<uint32_t>(c) | static_cast<uint32_t>(m) | kk);
                  }
                }
              }
              commit(win_empty(slot));
              if (c == p.nchunks - 1) commit(acc_full(as));
            }
            __syncwarp();
            if ((p.dbg & 16) && lane == 0) atomicAdd(&g_wait_prof[8], static_cast<unsigned long long>(clock64() - tq0));
            if (++slot == p.nslots) {
              slot = 0;
              wph ^= 1u;
            }
          }
          if (++as == p.nacc) {
            as = 0;
            aph ^= 1u;
          }
        }
      } else
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        const int ntile = p.res_local ? 0 : tile % p.ntiles_n;  // only used to index resident images
        wait_x(acc_empty(as), aph ^ 1u, 4);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * acc_stride;
        uint32_t accum = 0;
        for (int c = 0; c < p.nchunks; ++c) {
          wait_x(win_full(slot), wph, 5);
          const uint32_t a_slot = slots0 + slot * slot_bytes;
          for (int m0 = 0; m0 < p.ntaps; m0 += p.tps) {
            const int ntp = p.ntaps - m0 < p.tps ? p.ntaps - m0 : p.tps;
            if (!p.resident) {
              mbar_wait(b_full(bs), bph, p.err, 6);
              if (kPair) mbar_wait_cluster(peer_b_full(bs), bph, p.err, 7);
            }
            tc_fence_after();
            const uint32_t b_img = p.resident ? static_cast<uint32_t>((ntile * p.nchunks + c) * p.ntaps + m0) : 0u;
            const uint32_t b_lo0 = desc_lo(p.resident ? bst0 + b_img * bstage_bytes : bst0 + bs * bring_bytes, 16);
            const bool last = m0 + ntp == p.ntaps;
            if (elect_one()) {
              for (int mm = 0; mm < ntp; ++mm) {
                const uint32_t a_lo = desc_lo(a_slot + static_cast<uint32_t>(p.shift[m0 + mm] - p.minshift) * 1024u, 16);
                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(mm) * (bstage_bytes >> 4);
                if (kParts == 1) {
                  for (uint32_t mt = 0; mt < static_cast<uint32_t>(p.MT); ++mt) {
                    // 16 positions = 16*a_sbo bytes = a_sbo 16-byte units; mtg: the next column group's window
                    const uint32_t a_mt = a_lo + mt * (p.mtg ? static_cast<uint32_t>(p.win_atoms) * 64u : a_sbo), d_mt = d_tmem + mt * p.BN;
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk)
                      mma(d_mt, a_mt + kk * 2u, a_hi, b_lo + kk * 2u, b_hi, idesc, accum | static_cast<uint32_t>(mm) | kk);
                  }
                } else {
#pragma unroll
                  for (uint32_t kk = 0; kk < 4; ++kk) {
                    // (a0+a1+a2)(b0+b1+b2) ~ a0b0 + [a0b1 + a1b0 + a1b1 + a0b2 + a2b0] (rel. err ~2^-24)
                    const uint32_t pa[5] = {2, 0, 1, 1, 0};
                    const uint32_t pb[5] = {0, 2, 1, 0, 1};
#pragma unroll
                    for (int e = 0; e < 5; ++e)
                      mma(d_tmem + p.BN, a_lo + pa[e] * pa_lo + kk * 2u, a_hi, b_lo + pb[e] * pb_lo + kk * 2u, b_hi,
                                   idesc, accum | static_cast<uint32_t>(mm) | kk | static_cast<uint32_t>(e));
                    mma(d_tmem, a_lo + kk * 2u, a_hi, b_lo + kk * 2u, b_hi, idesc, accum | static_cast<uint32_t>(mm) | kk);
                  }
                }
              }
              if (!p.resident) commit(b_empty(bs));
              if (last) commit(win_empty(slot));
              if (last && c == p.nchunks - 1) commit(acc_full(as));
            }
            __syncwarp();
            accum = 1;
            if (!p.resident && ++bs == p.nbstages) {
              bs = 0;
              bph ^= 1u;
            }
          }
          if (++slot == p.nslots) {
            slot = 0;
            wph ^= 1u;
          }
        }
        if (++as == p.nacc) {
          as = 0;
          aph ^= 1u;
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // (pair: the leader's MMAs read the peer's shared memory until here)
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: W (fp32, arbitrary 2-level strides) -> per-(ntile, chunk, tap, part) images
// [BN rows (co)][64 k (ci)] bf16, SWIZZLE_128B, zero padded.
//   n = n1*N2 + n2  -> offset n1*sn1 + n2*sn2 ;  k = k1*K2 + k2 -> offset k1*sk1 + k2*sk2
//   tap m -> offset tapmap[m]*sm
// ------------------------------------------------------------------------------------------
struct PackParams {
  const float* w;
  void* out;
  int Nout, Kin, N2, K2;
  long long sn1, sn2, sk1, sk2, sm;
  int ntaps;
  int tapmap[9];
  int BN, ntiles_n, nchunks, nparts;
};

__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
  // one thread per (image, row, 16-byte chunk)
  const long long total = static_cast<long long>(p.ntiles_n) * p.nchunks * p.ntaps * p.BN * 8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pc = static_cast<int>(idx & 7);
    long long r0 = idx >> 3;
    const int row = static_cast<int>(r0 % p.BN);
    r0 /= p.BN;
    const int m = static_cast<int>(r0 % p.ntaps);
    r0 /= p.ntaps;
    const int c = static_cast<int>(r0 % p.nchunks);
    const int ntile = static_cast<int>(r0 / p.nchunks);
    const int n = ntile * p.BN + row;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = c * 64 + pc * 8 + i;
      float val = 0.f;
      if (n < p.Nout && k < p.Kin) {
        const long long off = (n / p.N2) * p.sn1 + (n % p.N2) * p.sn2 + (k / p.K2) * p.sk1 +
                              (k % p.K2) * p.sk2 + p.tapmap[m] * p.sm;
        val = p.w[off];
      }
      f[i] = val;
    }
    const size_t img = ((static_cast<size_t>(ntile) * p.nchunks + c) * p.ntaps + m) * p.nparts;
    uint8_t* base = reinterpret_cast<uint8_t*>(p.out) + img * (static_cast<size_t>(p.BN) * 128u);
    for (int part = 0; part < p.nparts; ++part) {
      uint4 u = split8_bf16(f);
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(part) * p.BN * 128u + sw128_off(row, pc)) = u;
    }
  }
}

static int pick_bn(int cout, int dtype) {
  // largest tile <= cap that is a multiple of 32 and divides the 32-padded channel count evenly
  // (fp32 activations carry 3 bf16 parts per operand, so their tiles are capped at 128 columns)
  const int cap = dtype == FMM_DT_F32 ? 128 : 256;
  const int c32 = (cout + 31) / 32 * 32;
  if (c32 <= cap) return c32;
  for (int bn = cap; bn >= 32; bn -= 32)
    if (c32 % bn == 0) return bn;
  return 32;
}

}  // namespace fmm

using namespace fmm;

extern "C" {

// Geometry shared by pack + launch so both sides agree on the packed layout.
int fmm_tapconv_bn(int cout, int dtype) { return pick_bn(cout, dtype); }

long long fmm_tapconv_packed_bytes(int cin, int cout, int ntaps, int dtype) {
  const int bn = pick_bn(cout, dtype);
  const int ntiles = ((cout + 31) / 32 * 32 + bn - 1) / bn;
  const int nchunks = (cin + 63) / 64;
  const int nparts = dtype == FMM_DT_F32 ? 3 : 1;
  return static_cast<long long>(ntiles) * nchunks * ntaps * nparts * bn * 128;
}

int fmm_tapconv_pack(const float* w, void* out, int cout, int cin, int n2, int k2, long long sn1,
                     long long sn2, long long sk1, long long sk2, long long sm, int ntaps,
                     const int* tapmap, int dtype, cudaStream_t stream) {
  FMM_CHECK_ARG(w && out && cout > 0 && cin > 0 && ntaps >= 1 && ntaps <= 9 && n2 > 0 && k2 > 0,
                "tapconv_pack: bad arguments");
  PackParams p;
  p.w = w;
  p.out = out;
  p.Nout = cout;
  p.Kin = cin;
  p.N2 = n2;
  p.K2 = k2;
  p.sn1 = sn1;
  p.sn2 = sn2;
  p.sk1 = sk1;
  p.sk2 = sk2;
  p.sm = sm;
  p.ntaps = ntaps;
  for (int i = 0; i < 9; ++i) p.tapmap[i] = i < ntaps ? tapmap[i] : 0;
  p.BN = pick_bn(cout, dtype);
  p.ntiles_n = ((cout + 31) / 32 * 32 + p.BN - 1) / p.BN;
  p.nchunks = (cin + 63) / 64;
  p.nparts = dtype == FMM_DT_F32 ? 3 : 1;
  const long long total = static_cast<long long>(p.ntiles_n) * p.nchunks * ntaps * p.BN * 8;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads < 1184 ? (total + threads - 1) / threads : 1184);
  pack_weights_kernel<<<blocks, threads, 0, stream>>>(p);
  FMM_CHECK_LAUNCH("tapconv_pack");
  return FMM_OK;
}

int fmm_tapconv(const void* x, void* out, const void* wpk, const float* in_scale,
                const float* in_shift, int in_relu, const float* bias, int bias_per_joint, int N, int V, int Tin,
                int Tout, int Cin, int Cout, int Tj, int istride, int ostride, int ooff, int ntaps,
                const int* shifts, int dtype, unsigned* err, cudaStream_t stream) {
  FMM_CHECK_ARG(x && out && wpk, "tapconv: null pointer");
  FMM_CHECK_ARG(N > 0 && V > 0 && Tin > 0 && Tout > 0 && Cin > 0 && Cout > 0 && Tj > 0, "tapconv: bad shape");
  FMM_CHECK_ARG(ntaps >= 1 && ntaps <= 9 && istride >= 1 && istride <= 2 && ostride >= 1,
                "tapconv: bad taps/stride");
  FMM_CHECK_ARG(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "tapconv: bad dtype");
  FMM_CHECK_ARG((Tj - 1) * ostride + ooff < Tout, "tapconv: output positions exceed Tout");
  TapConvParams p;
  p.x = x;
  p.out = out;
  p.wpk = wpk;
  p.in_scale = in_scale;
  p.in_shift = in_shift;
  p.bias = bias;
  p.bias_vstride = bias_per_joint ? Cout : 0;
  p.in_relu = in_relu;
  p.N = N;
  p.V = V;
  p.Tin = Tin;
  p.Tout = Tout;
  p.Cin = Cin;
  p.Cout = Cout;
  p.Tj = Tj;
  p.istride = istride;
  p.ostride = ostride;
  p.ooff = ooff;
  p.ntaps = ntaps;
  int mn = shifts[0], mx = shifts[0];
  for (int i = 0; i < 9; ++i) {
    p.shift[i] = i < ntaps ? shifts[i] : 0;
    if (i < ntaps) {
      mn = shifts[i] < mn ? shifts[i] : mn;
      mx = shifts[i] > mx ? shifts[i] : mx;
    }
  }
  p.minshift = mn;
  // two M tiles per item when the accumulators fit (2 stages x 2 tiles x BN columns) and the window stays small
  const int bn_probe = pick_bn(Cout, dtype);
  p.MT = (dtype == FMM_DT_BF16 && istride == 1 && bn_probe <= 128 && Tj > 16) ? 2 : 1;
  // CTA pairs (cta_group::2) for the multi-tap bf16 convs whose weights do not fit shared memory (256 channels: 1.18 MB of
  // images streamed per row tile, L2 -> SM fill bound): half of every weight image per CTA. Measured (B200, N=256 clips):
  // 256 ch 180 -> 150 us; 128 ch 90 -> 88 us (weights resident as halves, one M tile per item), 64 ch 59 -> 68 us (already
  // resident, N = 64 MMAs are shared-memory-read bound either way) - so only BN = 256 runs paired by default (FMM_TAP_PAIR=2:
  // every multi-tap launch, =0: none).
  static const int pair_env = getenv("FMM_TAP_PAIR") ? atoi(getenv("FMM_TAP_PAIR")) : 1;
  const bool pair = pair_env && (pair_env >= 2 || bn_probe >= 256) && dtype == FMM_DT_BF16 && ntaps > 1 && (Cin % 64) == 0 &&
                    (Cout % 32) == 0 && (num_sms() % 2) == 0;
  if (pair && bn_probe == 128) p.MT = 1;
  // 256-column pairs: two column groups per item (every weight half-image feeds 2 x 4 MMAs: half the L2 -> SM weight
  // stream, which bounds this shape) on a single 512-column accumulator stage
  static const int mtg_env = getenv("FMM_TAP_MTG") ? atoi(getenv("FMM_TAP_MTG")) : 1;
  p.mtg = (pair && mtg_env && bn_probe >= 256 && in_scale == nullptr) ? 1 : 0;
  if (p.mtg) p.MT = 2;
  p.nacc = (dtype == FMM_DT_BF16 && p.MT * bn_probe * 2 > 512) ? 1 : 2;
  p.win_atoms = (16 * (p.mtg ? 1 : p.MT) - 1) * istride + (mx - mn) + 1;
  p.BN = pick_bn(Cout, dtype);
  p.ntiles_n = ((Cout + 31) / 32 * 32 + p.BN - 1) / p.BN;
  p.nchunks = (Cin + 63) / 64;
  p.ncols = N * V;
  p.ngroups = (p.ncols + 7) / 8;
  const int mtt = p.mtg ? 1 : p.MT;
  p.ntchunks = (Tj + 16 * mtt - 1) / (16 * mtt);
  const int gunits = p.mtg ? (p.ngroups + p.MT - 1) / p.MT : p.ngroups;   // column-group units of one item
  p.total_tiles = pair ? (gunits * p.ntchunks + 1) / 2 * p.ntiles_n : gunits * p.ntchunks * p.ntiles_n;
  p.err = err;
  static const int dbg_env = getenv("FMM_TAP_DBG") ? atoi(getenv("FMM_TAP_DBG")) : 0;
  p.dbg = dbg_env;
  const int nparts = dtype == FMM_DT_F32 ? 3 : 1;
  const size_t slot_bytes = static_cast<size_t>(nparts) * p.win_atoms * 1024 * (p.mtg ? p.MT : 1);
  const size_t bstage_bytes = static_cast<size_t>(nparts) * p.BN * 128 / (pair ? 2 : 1);  // per CTA
  // the bias table ([V][Cout] for the per-joint bias of the graph conv, else [Cout]) rides along in shared memory when small
  const size_t bias_floats = bias ? static_cast<size_t>(bias_per_joint ? V : 1) * Cout : 0;
  p.bias_smem = (bias_floats > 0 && bias_floats * 4 <= 36 * 1024 && (Cout % 4) == 0) ? static_cast<int>(bias_floats) : 0;
  const size_t bias_bytes = (static_cast<size_t>(p.bias_smem) * 4 + 15) / 16 * 16;
  // shared-memory budget (FMM_TAP_SMEM_KB caps it: leaving a few KB lets small kernels of another stream share the SM)
  static const int smem_kb = getenv("FMM_TAP_SMEM_KB") ? atoi(getenv("FMM_TAP_SMEM_KB")) : 227;
  const size_t budget = static_cast<size_t>(smem_kb) * 1024 - 1024 /*align*/ - 2048 /*barriers*/ - bias_bytes;
  // Weights: keep every image resident when they fit next to >= 3 window slots (no per-tap barrier
  // round trips at all); otherwise stream them through as deep a ring as fits beside 4 slots.
  const int nimg = p.ntiles_n * p.nchunks * ntaps;
  int nb, ns;
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (pair) grid = 2 * (p.total_tiles < num_sms() / 2 ? p.total_tiles : num_sms() / 2);
  p.resident = (nimg <= 96 && nimg * bstage_bytes + 3 * slot_bytes <= budget) ? 1 : 0;
  p.res_local = 0;
  // several N tiles whose images do not all fit: a CTA whose tiles all share one N tile (tile index = rows * ntiles_n
  // + ntile, stride gridDim.x) keeps just that tile's images
  const int nimg_tile = p.nchunks * ntaps;
  if (!pair && !p.resident && p.ntiles_n > 1 && grid >= p.ntiles_n && nimg_tile <= 96 &&
      nimg_tile * bstage_bytes + 3 * slot_bytes <= budget) {
    p.resident = 1;
    p.res_local = 1;
    grid -= grid % p.ntiles_n;
  }
  p.tps = 1;
  if (p.resident) {
    nb = p.res_local ? nimg_tile : nimg;
    ns = static_cast<int>((budget - nb * bstage_bytes) / slot_bytes);
    if (ns > 8) ns = 8;  // more window slots = more cp.async bytes in flight (the 1x1 GEMMs are streaming kernels)
  } else {
    // several taps per ring stage (<= 48 KB) when the images are small: one barrier round trip per stage
    if (nparts == 1) {
      p.tps = static_cast<int>((48 * 1024) / bstage_bytes);
      if (p.tps > ntaps) p.tps = ntaps;
      if (p.tps < 1) p.tps = 1;
    }
    const size_t ring = bstage_bytes * p.tps;
    ns = 4;
    while (ns > 1 && 2 * ring + ns * slot_bytes > budget) --ns;
    nb = static_cast<int>((budget - ns * slot_bytes) / ring);
    if (nb > 8) nb = 8;
  }
  FMM_CHECK_ARG(nb >= 1 && ns >= 1 && nb * bstage_bytes * p.tps + ns * slot_bytes <= budget,
                "tapconv: tile does not fit shared memory (win_atoms=%d BN=%d parts=%d)", p.win_atoms, p.BN, nparts);
  p.nslots = ns;
  p.nbstages = nb;
  const size_t smem = nb * bstage_bytes * p.tps + ns * slot_bytes + 1024 + 2048 + bias_bytes;
  const bool wide = p.BN >= 128;
#define FMM_LAUNCH_TAPCONV2(TT, EPI, TAPS)                                                                       \
  do {                                                                                                           \
    cudaError_t e = cudaFuncSetAttribute(tapconv_kernel<TT, EPI, TAPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                             \
    if (e != cudaSuccess) {                                                                                      \
      set_last_error("tapconv: smem attribute: %s", cudaGetErrorString(e));                                      \
      return FMM_ERR_SMEM;                                                                                       \
    }                                                                                                            \
    tapconv_kernel<TT, EPI, TAPS, false><<<grid, (10 + EPI) * 32, smem, stream>>>(p);                            \
  } while (0)
#define FMM_LAUNCH_TAPCONV_PAIR(EPI)                                                                             \
  do {                                                                                                           \
    auto kern = tapconv_kernel<__nv_bfloat16, EPI, true, true>;                                                  \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e != cudaSuccess) {                                                                                      \
      set_last_error("tapconv: smem attribute: %s", cudaGetErrorString(e));                                      \
      return FMM_ERR_SMEM;                                                                                       \
    }                                                                                                            \
    cudaLaunchConfig_t cfg = {};                                                                                 \
    cfg.gridDim = dim3(grid);                                                                                    \
    cfg.blockDim = dim3((10 + EPI) * 32);                                                                        \
    cfg.dynamicSmemBytes = smem;                                                                                 \
    cfg.stream = stream;                                                                                         \
    cudaLaunchAttribute at[1];                                                                                   \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                              \
    at[0].val.clusterDim.x = 2;                                                                                  \
    at[0].val.clusterDim.y = 1;                                                                                  \
    at[0].val.clusterDim.z = 1;                                                                                  \
    cfg.attrs = at;                                                                                              \
    cfg.numAttrs = 1;                                                                                            \
    e = cudaLaunchKernelEx(&cfg, kern, p);                                                                       \
    if (e != cudaSuccess) {                                                                                      \
      set_last_error("tapconv: pair launch: %s", cudaGetErrorString(e));                                         \
      return FMM_ERR_CUDA;                                                                                       \
    }                                                                                                            \
  } while (0)
#define FMM_LAUNCH_TAPCONV(TT, EPI)                                                                              \
  do {                                                                                                           \
    if (ntaps > 1) FMM_LAUNCH_TAPCONV2(TT, EPI, true); else FMM_LAUNCH_TAPCONV2(TT, EPI, false);                 \
  } while (0)
  if (pair) {
    if (wide) FMM_LAUNCH_TAPCONV_PAIR(8); else FMM_LAUNCH_TAPCONV_PAIR(4);
  } else if (dtype == FMM_DT_BF16) {
    if (wide) FMM_LAUNCH_TAPCONV(__nv_bfloat16, 8); else FMM_LAUNCH_TAPCONV(__nv_bfloat16, 4);
  } else {
    if (wide) FMM_LAUNCH_TAPCONV(float, 8); else FMM_LAUNCH_TAPCONV(float, 4);
  }
#undef FMM_LAUNCH_TAPCONV2
#undef FMM_LAUNCH_TAPCONV_PAIR
#undef FMM_LAUNCH_TAPCONV
  FMM_CHECK_LAUNCH("tapconv");
  return FMM_OK;
}

}  // extern "C"
