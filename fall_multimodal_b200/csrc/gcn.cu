// gcn: the spatial graph convolution of an ST-GCN block as ONE tcgen05 GEMM with the adjacency aggregation applied in
// the prologue (BASELINE north star, piece 1).
//
//   G[r][co] = bias[w(r)][co] + sum_k sum_ci ( sum_{e in in(k,w(r))} coef[e] * x[frame(r)*V + src[e]][ci] ) * W[k*Cout+co][ci]
//
// i.e. einsum('nkctv,kvw->nctw') of the 1x1 conv output (reference stgcan.py:50-56, with A*edge_importance of :222),
// reassociated onto the input channels so that the K-times wider intermediate never exists: not in HBM (the round-1 path
// wrote and re-read it) and not in the reference's K*Cout form either. Rows r = (n,t,v) of the channels-last activations
// are a flat [R][C] matrix; a tile is 128 consecutive rows. Per tile and 64-channel slab of the input:
//
//   TMA (cp.async.bulk.tensor.2d)  x rows of all frames the tile touches -> shared memory "raw slab" (<= 192 rows x 128 B)
//   8 producer warps               A_k[row][64] = sum_e coef[e] * raw[frame(row)*V + src[e]] for k = 0..K-1, written as the
//                                  canonical SWIZZLE_128B K-major operand images (the (V,V) adjacency lives in shared
//                                  memory as a per-joint edge list, staged once per CTA)
//   1 MMA lane                     tcgen05.mma M=128 x N=Cout x K=16, accumulators in TMEM (double buffered)
//   4 epilogue warps               TMEM -> registers -> + per-joint bias -> bf16 -> swizzled staging tile -> TMA store
//                                  (cp.async.bulk.tensor.2d, full 128-byte lines); the per-channel sum / sum of squares
//                                  of what was stored (the BatchNorm statistics of stgcan.py:112) come from the same
//                                  staging tile, so no separate statistics pass reads G again.
//
// The same file holds the weight-gradient twin (gcn_wgrad): dW[k*Cout+co][ci] = sum_r dG[r][co] * A_k[r][ci] with the
// aggregated operand re-derived in the prologue by the same producer code, so the backward pass needs no saved copy of it.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmm {

// ------------------------------------------------------------------------------------------
// tensor-map TMA (2-D tiles). The map is a __grid_constant__ kernel parameter.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src_smem) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(src_smem)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
  return u;
}
__device__ __forceinline__ void sts128g(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(addr) : "memory");
  return u;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 u;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "r"(addr) : "memory");
  return u;
}

constexpr int kGcnTileRows = 128;
constexpr int kGcnRawBox = 64;                       // rows per TMA box of the raw slab
constexpr int kGcnRawBoxes = 3;                      // 192 rows >= 128 + 2*(V-1) for V <= 33
constexpr int kGcnRawRows = kGcnRawBox * kGcnRawBoxes;
constexpr uint32_t kGcnRawBytes = kGcnRawRows * 128u;
constexpr uint32_t kGcnChunkBytes = kGcnTileRows * 128u;   // one [128 rows][64 ch] bf16 operand image
constexpr int kGcnProducers = 256;
constexpr int kGcnMaxV = 33;

struct GcnEdges {
  // shared-memory adjacency: ptr[(w*K + k)] = (first, last) into tab[]; tab[e] = (src*128 bytes, coef bits)
  uint32_t ptr;
  uint32_t tab;
};

// Stage the CSR adjacency (rowptr over k*V+w, src, coef: csrc/elementwise.cu agg_fwd layout) in shared memory, per joint.
__device__ __forceinline__ void gcn_stage_edges(const GcnEdges& ed, const int* __restrict__ rowptr, const int* __restrict__ src,
                                                const float* __restrict__ coef, int V, int K, int E) {
  for (int i = threadIdx.x; i < V * K; i += blockDim.x) {
    const int k = i / V, w = i - k * V;
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(ed.ptr + 8u * (w * K + k)), "r"(rowptr[i]), "r"(rowptr[i + 1]) : "memory");
  }
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(ed.tab + 8u * e), "r"(src[e] * 128), "r"(__float_as_uint(coef[e])) : "memory");
}

// One producer thread's share of an aggregated operand image: rows (tid>>3) + 32*i, 16-byte piece tid&7.
//   raw    : shared address of the raw slab (row j at j*128, pieces unswizzled)
//   fb[i]  : byte offset of row i's frame inside the slab (negative never happens); valid[i]: row exists
//   dst    : shared address of the [128][64] SWIZZLE_128B image
struct GcnRows {
  int w[4];
  int fb[4];
  bool valid[4];
};
__device__ __forceinline__ void gcn_produce_chunk(const GcnEdges& ed, const GcnRows& rw, uint32_t raw, uint32_t dst, int k, int K) {
  const uint32_t p = threadIdx.x & 7u;
  const uint32_t rl = threadIdx.x >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (rw.valid[i]) {
      const uint2 pe = lds64(ed.ptr + 8u * static_cast<uint32_t>(rw.w[i] * K + k));
      const uint32_t base = raw + static_cast<uint32_t>(rw.fb[i]) + p * 16u;
      for (uint32_t e = pe.x; e < pe.y; ++e) {
        const uint2 en = lds64(ed.tab + 8u * e);
        const uint4 u = lds128(base + en.x);
        const float cf = __uint_as_float(en.y);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __bfloat1622float2(h[j]);
          acc[2 * j] = fmaf(cf, t.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(cf, t.y, acc[2 * j + 1]);
        }
      }
    }
    const uint32_t row = rl + 32u * i;
    sts128g(dst + row * 128u + ((p ^ (row & 7u)) << 4), pack8_bf16(acc));
  }
}
__device__ __forceinline__ void gcn_rows_of_tile(GcnRows& rw, long long r0, int raw_start, long long R, int V) {
  const int rl = threadIdx.x >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + rl + 32 * i;
    rw.valid[i] = r < R;
    const int f = static_cast<int>((rw.valid[i] ? r : r0) / V);
    rw.w[i] = static_cast<int>((rw.valid[i] ? r : r0) - static_cast<long long>(f) * V);
    rw.fb[i] = (f * V - raw_start) * 128;
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct GcnFwdParams {
  const float* bias;   // [V][Cout] or null
  const void* wpk;     // images [(cc*K + k)][BN rows][64] bf16, SWIZZLE_128B (fmm_gcn_pack)
  const int* rowptr;
  const int* src;
  const float* coef;
  double* ch_sum;      // [nrep][Cout] (nullable)
  double* ch_sq;
  int nrep;
  long long R;
  int V, K, Cin, Cout, E;
  int BN, NCC, ntiles;
  int n_raw, n_a, n_b, resident;
  int write_xa;
  unsigned* err;
};

constexpr int kGcnFwdThreads = 480;  // 8 producer, 4 epilogue, raw loader, weight loader, MMA warps

__global__ void __launch_bounds__(kGcnFwdThreads, 1)
gcn_fwd_kernel(const __grid_constant__ GcnFwdParams p, const __grid_constant__ CUtensorMap tm_x,
               const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_xa) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const uint32_t raw0 = base;
  const uint32_t a0 = raw0 + p.n_raw * kGcnRawBytes;
  const uint32_t b0 = a0 + p.n_a * kGcnChunkBytes;
  const uint32_t stg0 = b0 + p.n_b * b_bytes;                     // 4 x 4096: per-epilogue-warp staging tile [32][64] bf16
  const uint32_t bias0 = stg0 + 4u * 4096u;                        // [Cout/4][V][4] fp32
  const uint32_t bias_bytes = (p.bias ? static_cast<uint32_t>(p.V * p.Cout) * 4u : 0u);
  const uint32_t stat0 = bias0 + ((bias_bytes + 15u) & ~15u);      // [4 warps][BN][2] fp32
  const uint32_t stat_bytes = 4u * static_cast<uint32_t>(p.BN) * 8u;
  GcnEdges ed;
  ed.ptr = stat0 + stat_bytes;
  ed.tab = ed.ptr + 8u * static_cast<uint32_t>(p.V * p.K);
  const uint32_t bars0 = (ed.tab + 8u * static_cast<uint32_t>(p.E) + 15u) & ~15u;
  auto raw_full = [&](int s) { return bars0 + 8u * s; };
  auto raw_empty = [&](int s) { return bars0 + 8u * (p.n_raw + s); };
  auto a_full = [&](int s) { return bars0 + 8u * (2 * p.n_raw + s); };
  auto a_empty = [&](int s) { return bars0 + 8u * (2 * p.n_raw + p.n_a + s); };
  auto b_full = [&](int s) { return bars0 + 8u * (2 * p.n_raw + 2 * p.n_a + s); };
  auto b_empty = [&](int s) { return bars0 + 8u * (2 * p.n_raw + 2 * p.n_a + p.n_b + s); };
  auto acc_full = [&](int s) { return bars0 + 8u * (2 * p.n_raw + 2 * p.n_a + 2 * p.n_b + s); };
  auto acc_empty = [&](int s) { return bars0 + 8u * (2 * p.n_raw + 2 * p.n_a + 2 * p.n_b + 2 + s); };
  const uint32_t tmem_slot = bars0 + 8u * (2 * p.n_raw + 2 * p.n_a + 2 * p.n_b + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * static_cast<uint32_t>(p.BN)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_raw; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(raw_empty(s), kGcnProducers);
    }
    for (int s = 0; s < p.n_a; ++s) {
      mbar_init(a_full(s), kGcnProducers);
      mbar_init(a_empty(s), p.write_xa ? 2 : 1);
    }
    for (int s = 0; s < p.n_b; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 128);
    }
    mbar_fence_init();
  }
  if (warp == 14) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
    if (p.write_xa) tma_prefetch_desc(&tm_xa);
  }
  gcn_stage_edges(ed, p.rowptr, p.src, p.coef, p.V, p.K, p.E);
  // bias table [Cout/4][V][4]: the 32 rows of an epilogue warp are consecutive joints -> consecutive float4
  for (int i = threadIdx.x; i < (p.bias ? p.V * p.Cout : 0); i += blockDim.x) {
    const int v = i / p.Cout, c = i - v * p.Cout;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias0 + 4u * (((c >> 2) * p.V + v) * 4 + (c & 3))), "f"(p.bias[i]) : "memory");
  }
  for (int i = threadIdx.x; i < 4 * p.BN * 2; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(stat0 + 4u * i), "f"(0.f) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  const int nchunk = p.NCC * p.K;  // operand chunks (MMA k-blocks of 64) per tile

  if (warp < 8) {
    // ------------------------------ aggregation producers ------------------------------
    int rs = 0, as = 0;
    uint32_t rph = 0, aph = 0;
    for (int tile = first_tile; tile < p.ntiles; tile += tile_step) {
      const long long r0 = static_cast<long long>(tile) * kGcnTileRows;
      const int raw_start = static_cast<int>(r0 / p.V) * p.V;
      GcnRows rw;
      gcn_rows_of_tile(rw, r0, raw_start, p.R, p.V);
      for (int cc = 0; cc < p.NCC; ++cc) {
        mbar_wait(raw_full(rs), rph, p.err, 1);
        const uint32_t raw = raw0 + rs * kGcnRawBytes;
        for (int k = 0; k < p.K; ++k) {
          mbar_wait(a_empty(as), aph ^ 1u, p.err, 2);
          gcn_produce_chunk(ed, rw, raw, a0 + as * kGcnChunkBytes, k, p.K);
          fence_proxy_async_smem();
          mbar_arrive(a_full(as));
          if (++as == p.n_a) {
            as = 0;
            aph ^= 1u;
          }
        }
        mbar_arrive(raw_empty(rs));
        if (++rs == p.n_raw) {
          rs = 0;
          rph ^= 1u;
        }
      }
    }
  } else if (warp < 12) {
    // ---------------------------------- epilogue ----------------------------------
    const int quad = warp & 3;
    const uint32_t stg = stg0 + static_cast<uint32_t>(quad) * 4096u;
    const uint32_t my_stat = stat0 + static_cast<uint32_t>(quad) * static_cast<uint32_t>(p.BN) * 8u;
    const bool stats = p.ch_sum != nullptr;
    int acs = 0;
    uint32_t acph = 0;
    for (int tile = first_tile; tile < p.ntiles; tile += tile_step) {
      const long long r0 = static_cast<long long>(tile) * kGcnTileRows;
      const long long r = r0 + quad * 32 + lane;
      const bool row_ok = r < p.R;
      const int w = static_cast<int>((row_ok ? r : 0) % p.V);
      mbar_wait(acc_full(acs), acph, p.err, 3);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acs) * static_cast<uint32_t>(p.BN) + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c64 = 0; c64 < p.BN / 64; ++c64) {
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + c64 * 64, v0);
        tmem_ld32(taddr + c64 * 64 + 32, v1);
        tmem_ld_wait();
        if (c64 == p.BN / 64 - 1) {
          // the accumulator is in registers: hand the TMEM stage back before the stores
          tc_fence_before();
          mbar_arrive(acc_empty(acs));
        }
        // the previous TMA store of this warp must have read the staging tile before it is overwritten
        if (lane == 0) tma_wait_read<0>();
        __syncwarp();
        const uint32_t srow = stg + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float f[8];
          const uint32_t* vv = g < 4 ? v0 : v1;
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(vv[(g & 3) * 8 + i]);
          if (p.bias) {
            const int co = c64 * 64 + g * 8;
            const uint4 ba = lds128(bias0 + 16u * static_cast<uint32_t>((co >> 2) * p.V + w));
            const uint4 bb = lds128(bias0 + 16u * static_cast<uint32_t>(((co >> 2) + 1) * p.V + w));
            f[0] += __uint_as_float(ba.x); f[1] += __uint_as_float(ba.y); f[2] += __uint_as_float(ba.z); f[3] += __uint_as_float(ba.w);
            f[4] += __uint_as_float(bb.x); f[5] += __uint_as_float(bb.y); f[6] += __uint_as_float(bb.z); f[7] += __uint_as_float(bb.w);
          }
          if (!row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = 0.f;
          }
          sts128g(srow + ((static_cast<uint32_t>(g) ^ (static_cast<uint32_t>(lane) & 7u)) << 4), pack8_bf16(f));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_g, c64 * 64, static_cast<int>(r0) + quad * 32, stg);
          tma_commit();
        }
        if (stats) {
          // column sums of the bf16 values just staged: lane l owns channels 2l, 2l+1 (one 32-bit word per row)
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint32_t piece = static_cast<uint32_t>(lane) >> 2, word = (static_cast<uint32_t>(lane) & 3u) * 4u;
#pragma unroll 8
          for (uint32_t rr = 0; rr < 32; ++rr) {
            const uint32_t u = lds32(stg + rr * 128u + ((piece ^ (rr & 7u)) << 4) + word);
            const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
            s0 += t.x; s1 += t.y;
            q0 = fmaf(t.x, t.x, q0); q1 = fmaf(t.y, t.y, q1);
          }
          const uint32_t sa = my_stat + 8u * static_cast<uint32_t>(c64 * 64 + 2 * lane);
          uint4 old = lds128(sa);
          old.x = __float_as_uint(__uint_as_float(old.x) + s0);
          old.y = __float_as_uint(__uint_as_float(old.y) + q0);
          old.z = __float_as_uint(__uint_as_float(old.z) + s1);
          old.w = __float_as_uint(__uint_as_float(old.w) + q1);
          sts128g(sa, old);
        }
      }
      if (++acs == 2) {
        acs = 0;
        acph ^= 1u;
      }
    }
    if (lane == 0) tma_wait_read<0>();
    __syncwarp();
    if (stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int t = threadIdx.x - 256;  // 0..127
      const int rep = blockIdx.x % p.nrep;
      for (int c = t; c < p.BN && c < p.Cout; c += 128) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int wq = 0; wq < 4; ++wq) {
          const uint2 u = lds64(stat0 + static_cast<uint32_t>(wq) * static_cast<uint32_t>(p.BN) * 8u + 8u * c);
          s += __uint_as_float(u.x);
          q += __uint_as_float(u.y);
        }
        atomicAdd(p.ch_sum + static_cast<size_t>(rep) * p.Cout + c, static_cast<double>(s));
        atomicAdd(p.ch_sq + static_cast<size_t>(rep) * p.Cout + c, static_cast<double>(q));
      }
    }
  } else if (warp == 12) {
    // ------------------------------ raw-slab loader (TMA) ------------------------------
    if (lane == 0) {
      int rs = 0;
      uint32_t rph = 0;
      for (int tile = first_tile; tile < p.ntiles; tile += tile_step) {
        const long long r0 = static_cast<long long>(tile) * kGcnTileRows;
        const int raw_start = static_cast<int>(r0 / p.V) * p.V;
        for (int cc = 0; cc < p.NCC; ++cc) {
          mbar_wait(raw_empty(rs), rph ^ 1u, p.err, 4);
          mbar_arrive_expect_tx(raw_full(rs), kGcnRawBytes);
#pragma unroll
          for (int b = 0; b < kGcnRawBoxes; ++b)
            tma_load_2d(raw0 + rs * kGcnRawBytes + b * (kGcnRawBox * 128u), &tm_x, cc * 64, raw_start + b * kGcnRawBox, raw_full(rs));
          if (++rs == p.n_raw) {
            rs = 0;
            rph ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // -------------------------------- weight loader --------------------------------
    if (lane == 0 && first_tile < p.ntiles) {
      const uint8_t* W = reinterpret_cast<const uint8_t*>(p.wpk);
      if (p.resident) {
        mbar_arrive_expect_tx(b_full(0), static_cast<uint32_t>(nchunk) * b_bytes);
        for (int i = 0; i < nchunk; ++i) bulk_g2s(b0 + i * b_bytes, W + static_cast<size_t>(i) * b_bytes, b_bytes, b_full(0));
      } else {
        int bs = 0;
        uint32_t bph = 0;
        for (int tile = first_tile; tile < p.ntiles; tile += tile_step)
          for (int i = 0; i < nchunk; ++i) {
            mbar_wait(b_empty(bs), bph ^ 1u, p.err, 5);
            mbar_arrive_expect_tx(b_full(bs), b_bytes);
            bulk_g2s(b0 + bs * b_bytes, W + static_cast<size_t>(i) * b_bytes, b_bytes, b_full(bs));
            if (++bs == p.n_b) {
              bs = 0;
              bph ^= 1u;
            }
          }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------- MMA issuer ----------------------------------
    const uint32_t idesc = make_idesc_bf16(p.BN, 0, 0);
    const uint32_t hi = desc_hi(1024);
    int as = 0, bs = 0, acs = 0, prev_as = -1;
    uint32_t aph = 0, bph = 0, acph = 0;
    if (p.resident && first_tile < p.ntiles) mbar_wait(b_full(0), 0, p.err, 6);
    for (int tile = first_tile; tile < p.ntiles; tile += tile_step) {
      mbar_wait(acc_empty(acs), acph ^ 1u, p.err, 7);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acs) * static_cast<uint32_t>(p.BN);
      const int r0 = tile * kGcnTileRows;
      for (int i = 0; i < nchunk; ++i) {
        mbar_wait(a_full(as), aph, p.err, 8);
        if (!p.resident) mbar_wait(b_full(bs), bph, p.err, 9);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(a0 + as * kGcnChunkBytes, 16);
        const uint32_t b_lo = desc_lo(p.resident ? b0 + i * b_bytes : b0 + bs * b_bytes, 16);
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < 4; ++kk)
            umma_bf16_lh(d_tmem, a_lo + kk * 2u, hi, b_lo + kk * 2u, hi, idesc, static_cast<uint32_t>(i) | kk);
          umma_commit(a_empty(as));
          if (!p.resident) umma_commit(b_empty(bs));
          if (i == nchunk - 1) umma_commit(acc_full(acs));
          if (p.write_xa) {
            // training-mode side output (dev / fallback path): the aggregated operand image as rows of Xa[R][K*Cin]
            const int cc = i / p.K, k = i - cc * p.K;
            tma_store_2d(&tm_xa, k * p.Cin + cc * 64, r0, a0 + as * kGcnChunkBytes);
            tma_commit();
            if (prev_as >= 0) {
              tma_wait_read<1>();
              mbar_arrive(a_empty(prev_as));
            }
          }
        }
        __syncwarp();
        prev_as = as;
        if (++as == p.n_a) {
          as = 0;
          aph ^= 1u;
        }
        if (!p.resident && ++bs == p.n_b) {
          bs = 0;
          bph ^= 1u;
        }
      }
      if (++acs == 2) {
        acs = 0;
        acph ^= 1u;
      }
    }
    if (p.write_xa && prev_as >= 0) {
      if (elect_one()) {
        tma_wait_read<0>();
        mbar_arrive(a_empty(prev_as));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 14) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// weights W[(k*Cout + co)][ci] fp32 -> images [(cc*K + k)][BN rows = co][64 = ci - cc*64] bf16 SWIZZLE_128B (zero padded)
__global__ void gcn_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, int K, int Cin, int Cout, int BN, int NCC) {
  const int total = NCC * K * BN * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int pc = idx & 7;
    int r = idx >> 3;
    const int row = r % BN;
    r /= BN;
    const int k = r % K;
    const int cc = r / K;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ci = cc * 64 + pc * 8 + i;
      f[i] = (row < Cout && ci < Cin) ? w[(static_cast<size_t>(k) * Cout + row) * Cin + ci] : 0.f;
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(cc * K + k) * BN * 128 + sw128_off(row, pc)) = pack8_bf16(f);
  }
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 matrix [rows][cols] (row pitch = cols), box [box_rows][box_cols]
int make_tmap_2d(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows, int box_cols, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return FMM_ERR_CUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) for a [%lld][%d] bf16 matrix, box [%d][%d]", (int)r, rows, cols, box_rows, box_cols);
    return FMM_ERR_CUDA;
  }
  return FMM_OK;
}

}  // namespace fmm

using namespace fmm;

extern "C" {

long long fmm_gcn_packed_bytes(int K, int Cin, int Cout) {
  const int BN = (Cout + 63) / 64 * 64;
  const int NCC = (Cin + 63) / 64;
  return static_cast<long long>(NCC) * K * BN * 128;
}

int fmm_gcn_pack(const float* w, void* out, int K, int Cin, int Cout, cudaStream_t stream) {
  FMM_CHECK_ARG(w && out && K > 0 && Cin > 0 && Cout > 0, "gcn_pack: bad arguments");
  const int BN = (Cout + 63) / 64 * 64, NCC = (Cin + 63) / 64;
  const int total = NCC * K * BN * 8;
  gcn_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<uint8_t*>(out), K, Cin, Cout, BN, NCC);
  FMM_CHECK_LAUNCH("gcn_pack");
  return FMM_OK;
}

int fmm_gcn_fwd(const void* x, void* g, void* xa, const void* wpk, const float* bias, const int* rowptr, const int* src,
                const float* coef, double* ch_sum, double* ch_sq, int nrep, long long rows, int V, int K, int Cin, int Cout,
                int E, unsigned* err, cudaStream_t stream) {
  FMM_CHECK_ARG(x && g && wpk && rowptr && src && coef, "gcn_fwd: null pointer");
  FMM_CHECK_ARG(rows > 0 && rows < (1ll << 31) && V > 0 && V <= kGcnMaxV && K > 0 && K <= 8, "gcn_fwd: bad shape (rows=%lld V=%d K=%d)", rows, V, K);
  FMM_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0 && Cout <= 256 && Cin <= 512, "gcn_fwd: channels must be multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  FMM_CHECK_ARG((ch_sum == nullptr) == (ch_sq == nullptr) && (ch_sum == nullptr || nrep > 0), "gcn_fwd: statistics buffers");
  FMM_CHECK_ARG(E > 0 && E <= 4096, "gcn_fwd: edge count %d", E);
  GcnFwdParams p;
  p.bias = bias;
  p.wpk = wpk;
  p.rowptr = rowptr;
  p.src = src;
  p.coef = coef;
  p.ch_sum = ch_sum;
  p.ch_sq = ch_sq;
  p.nrep = nrep > 0 ? nrep : 1;
  p.R = rows;
  p.V = V;
  p.K = K;
  p.Cin = Cin;
  p.Cout = Cout;
  p.E = E;
  p.BN = Cout;
  p.NCC = Cin / 64;
  p.ntiles = static_cast<int>((rows + kGcnTileRows - 1) / kGcnTileRows);
  p.write_xa = xa != nullptr;
  p.err = err;
  const size_t b_bytes = static_cast<size_t>(p.BN) * 128;
  const size_t fixed = 4 * 4096 + ((static_cast<size_t>(bias ? V * Cout : 0) * 4 + 15) & ~15ull) + 4 * static_cast<size_t>(p.BN) * 8 +
                       8 * static_cast<size_t>(V * K) + 8 * static_cast<size_t>(E) + 16 + 512 /*barriers*/ + 1024 /*align*/;
  const size_t budget = 227 * 1024;
  const int nchunk = p.NCC * K;
  p.n_raw = 2;
  p.n_a = 3;
  size_t used = fixed + p.n_raw * kGcnRawBytes + p.n_a * kGcnChunkBytes;
  FMM_CHECK_ARG(used + b_bytes <= budget, "gcn_fwd: tile does not fit shared memory");
  if (used + nchunk * b_bytes <= budget) {
    p.resident = 1;
    p.n_b = nchunk;
  } else {
    p.resident = 0;
    p.n_b = static_cast<int>((budget - used) / b_bytes);
    if (p.n_b > 4) p.n_b = 4;
  }
  used += p.n_b * b_bytes;
  // spend what is left on deeper pipelines (more TMA bytes in flight / more aggregation ahead of the tensor core)
  while (p.n_a < 6 && used + kGcnChunkBytes <= budget) {
    ++p.n_a;
    used += kGcnChunkBytes;
  }
  while (p.n_raw < 3 && used + kGcnRawBytes <= budget) {
    ++p.n_raw;
    used += kGcnRawBytes;
  }
  CUtensorMap tm_x, tm_g, tm_xa;
  int st = make_tmap_2d(&tm_x, x, rows, Cin, kGcnRawBox, 64, false);
  if (st != FMM_OK) return st;
  st = make_tmap_2d(&tm_g, g, rows, Cout, 32, 64, true);
  if (st != FMM_OK) return st;
  if (xa) {
    st = make_tmap_2d(&tm_xa, xa, rows, K * Cin, kGcnTileRows, 64, true);
    if (st != FMM_OK) return st;
  } else {
    tm_xa = tm_g;
  }
  cudaError_t e = cudaFuncSetAttribute(gcn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(used));
  if (e != cudaSuccess) {
    set_last_error("gcn_fwd: smem attribute (%zu bytes): %s", used, cudaGetErrorString(e));
    return FMM_ERR_SMEM;
  }
  const int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
  gcn_fwd_kernel<<<grid, kGcnFwdThreads, used, stream>>>(p, tm_x, tm_g, tm_xa);
  FMM_CHECK_LAUNCH("gcn_fwd");
  return FMM_OK;
}

}  // extern "C"
