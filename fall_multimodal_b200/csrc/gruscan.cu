// Persistent graph-GRU scan of the TRAGCN family (reference: GRU.py:17-27 around EmbGCN.py:69-89, scanned over
// time by TRAGCN.py:158-166): ONE launch runs all T steps of a layer.
//
// Decomposition. A thread-block cluster of 8 CTAs owns 16*MT clips for the whole sequence. CTA j of the cluster
// owns the hidden channels [8j, 8j+8) of EVERY joint: the z, r and candidate columns 8j..8j+7 of both EmbGCN
// products, i.e. a 24-column slice of the per-node weights W_n (64 x 192 for the recurrent half), which its warps
// keep as mma.sync B fragments IN REGISTERS for the whole launch (warp w owns joints w, w+8, ...). The recurrent
// state never leaves the chip's caches on the critical path: after every half step each CTA stores its 8-channel
// slice of (plain, adjacency-mixed) state to the blocked global buffers - which double as the tensors the weight
// gradients need - and ONE multicast bulk copy (cp.async.bulk ... .multicast::cluster) lands that slice in the
// shared memory of all 8 CTAs, where the next half step reads its full-K A operands with ldmatrix. The adjacency
// mix (S = I + softmax(relu(E E^T)), EmbGCN.py:73-74,83) is channel-wise, so each CTA mixes only its own slice:
// one small mma.sync product (S - I in bf16, the identity added exactly).
//
// The input-dependent half of both products does not depend on the state: the same kernel in MODE 0 ("xpart")
// evaluates it for all T steps in parallel and leaves it in FRAGMENT ORDER (the value a lane needs sits at
// [item][chunk][lane]), so the scan reads it with fully coalesced 16-byte loads.
//
//   blocked state layout  XC[slot][cluster][j = 8 slices][pm: 0 plain, 1 mixed][V][BC clips][8 ch]   (bf16)
//   fragment order        PX[t][cluster][j][item][3 chunks][32 lanes][8]  chunks: (PXz,PXr) (LXz,LXr) (PXu,LXu)
//                         FS[t][cluster][j][item][4 chunks][32 lanes][8]  chunks: (z,r) (hc,hprev) (LGz,LGr) (LU,0)
//   item = (warp * NPW + q) * MT + mt  <->  joint n = warp + 8 q, clips 16 mt .. 16 mt + 15
//   a lane's 4 values of a chunk half = the m16n8 accumulator fragment: (clip g, ch 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)
#include "common.cuh"
#include "ptx.cuh"

namespace fmm {

struct GruScanArgs {
  const void* xb;    // xpart input, blocked with `xb_slices` slices per block; block of step t = slot t + xb_slot0
  void* px;          // fragment-order input half (written by MODE 0, read by MODE 1)
  void* xcg;         // [T+1] slots: state BEFORE step t (slot 0 = zeros, never read)
  void* xcu;         // [T] slots: r * h_{t-1}
  void* fs;          // saved gate values (null: inference)
  void* hout;        // (B,T,V,64) bf16
  const void* W;     // [V][K][192] bf16 per-node weights, columns (z 64 | r 64 | candidate 64)
  const void* Lw;    // [K][192] bf16 shared Linear weights (unscaled)
  const float* cs;   // [V] column scale of the static-adjacency path (EmbGCN.py:77)
  const float* S;    // [V][V] fp32 supports
  const float* bg;   // [V][192] fp32 per-node bias (MODE 0)
  const float* bl;   // [192] fp32 Linear bias (MODE 0)
  // backward (MODE 2)
  const void* dhout; long long dh_b, dh_t, dh_v;   // gradient of hout, element strides
  void* dxu; void* dxgz; void* dxgr;                // blocked [T] slots: pm 0 = Linear-path, 1 = graph-path pre-activation gradients
  const void* WT;    // [V][192][64] bf16: WT[n][col][k] = W[n][k][col]
  const void* LT;    // [192][64]
  unsigned* err;
  unsigned long long* prof;   // dev aid: 16 cycle counters of cluster 0 / CTA 0 / thread 0 (null: off)
  int B, T, V, KS, xb_slices, xb_slot0, NC, tsplit;
};

namespace gs {

typedef __nv_bfloat16 bf16;
constexpr int CL = 8;     // CTAs per cluster = 8-channel slices of the 64 hidden channels
constexpr int NW = 8;     // warps per CTA
constexpr int NT = NW * 32;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// global -> the same shared-memory offset of every CTA in `mask`, completion on the barrier at the same offset of each
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float rb(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// one MUFU each (tanh.approx: relative error 2^-11, below the bf16 rounding every gate value gets)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float dsilu(float x) {
  const float s = sigm(x);
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ void unpack4(uint32_t a, uint32_t b, float (&v)[4]) {
  v[0] = lo(a); v[1] = hi(a); v[2] = lo(b); v[3] = hi(b);
}
// generic-proxy global stores of this thread -> ordered before the async-proxy (bulk copy) reads issued after the next barrier
__device__ __forceinline__ void publish_global() {
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
}

// shared-memory address of lane's ldmatrix row for the A fragment (16 clips x 16 channels) of joint n, clip tile mt, k-step ks
__device__ __forceinline__ uint32_t a_addr(uint32_t buf, uint32_t slice_bytes, int V, int BC, int pm, int n, int mt, int ks, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  return buf + (2 * ks + (mi >> 1)) * slice_bytes + ((pm * V + n) * BC + mt * 16 + (mi & 1) * 8 + r) * 16;
}

// B fragments of the per-node weights (K x 192, this CTA's z / r / candidate columns 8j+g) for the joints of warp w
template <int NPW>
__device__ __forceinline__ void load_wregs(uint32_t (&wreg)[NPW][24], const bf16* W, int K, int V, int w, int j, int g, int tq) {
  const uint16_t* Wu = reinterpret_cast<const uint16_t*>(W);
#pragma unroll
  for (int q = 0; q < NPW; ++q) {
    const int n = w + 8 * q;
#pragma unroll
    for (int i = 0; i < 24; ++i) wreg[q][i] = 0u;
    if (n < V) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int k0 = ks * 16 + 2 * tq + 8 * r;
          if (k0 < K) {
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
              const int col = nt * 64 + 8 * j + g;
              const uint32_t v0 = Wu[((size_t)n * K + k0) * 192 + col], v1 = Wu[((size_t)n * K + k0 + 1) * 192 + col];
              wreg[q][nt < 2 ? ks * 4 + nt * 2 + r : 16 + ks * 2 + r] = v0 | (v1 << 16);
            }
          }
        }
    }
  }
}

// dst[n][clip][2t..] (global, this CTA's mixed slot) = stg[n] + sum_m (S - I)[n][m] stg[m]   for this CTA's 8 channels
template <int MT>
__device__ __forceinline__ void mix_slice(uint32_t stg, uint32_t ssm, int V, int VP, int w, int lane, bf16* dst) {
  constexpr int BC = 16 * MT, CPW = BC / 8;
  const int g = lane >> 2, tq = lane & 3;
  const int nmt = VP >> 4;
  uint32_t af[2][2][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      if (mi < nmt && kk < nmt)
        ldsm_x4(ssm + ((mi * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * (VP + 8) + kk * 16 + (lane >> 4) * 8) * 2, af[mi][kk]);
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int cl = w * CPW + c;
    float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      if (kk < nmt) {
        uint32_t b0, b1;
        ldsm_x2_trans(stg + ((kk * 16 + (lane & 15)) * BC + cl) * 16, b0, b1);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
          if (mi < nmt) mma16816(d[mi], af[mi][kk], b0, b1);
      }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int n = mi * 16 + g + 8 * half;
        if (mi < nmt && n < V) {
          const uint32_t pv = lds32(stg + (n * BC + cl) * 16 + tq * 4);
          *reinterpret_cast<uint32_t*>(dst + ((size_t)n * BC + cl) * 8 + 2 * tq) = pk(d[mi][2 * half] + lo(pv), d[mi][2 * half + 1] + hi(pv));
        }
      }
  }
}

struct Smem {
  uint32_t buf, stg, ls, ssm, csm, bar;
  uint32_t slice_bytes;
};
template <int MT>
__host__ __device__ inline size_t smem_bytes(int V) {
  const int BC = 16 * MT, VP = (V + 15) & ~15;
  return (size_t)CL * 2 * V * BC * 16 + (size_t)VP * BC * 16 + 24 * 72 * 2 + 32 * 40 * 2 + 32 * 4 + 16;
}

// MODE 0: input half for all steps (no recurrence); MODE 1: forward scan
template <int NPW, int MT, int MODE>
__global__ void __launch_bounds__(NT, 1) gruscan_kernel(const GruScanArgs p) {
  constexpr int BC = 16 * MT, ITEMS = NW * NPW * MT;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int V = p.V, VP = (V + 15) & ~15, T = p.T, NC = p.NC;
  const uint32_t slice_bytes = 2u * V * BC * 16;
  const size_t slice_el = (size_t)V * BC * 16;   // bf16 elements of one CTA's (plain, mixed) slice
  const uint32_t buf = smem_u32(smem_raw);
  const uint32_t stg = buf + CL * slice_bytes;
  const uint32_t ls = stg + VP * BC * 16;
  const uint32_t ssm = ls + 24 * 72 * 2;
  const uint32_t csm_a = ssm + 32 * 40 * 2;
  const uint32_t bar = csm_a + 32 * 4;
  bf16* Ls = reinterpret_cast<bf16*>(smem_raw + (ls - buf));
  bf16* Ssm = reinterpret_cast<bf16*>(smem_raw + (ssm - buf));
  float* csm = reinterpret_cast<float*>(smem_raw + (csm_a - buf));

  const int j = (int)cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int nc = (MODE == 0) ? cid % NC : cid;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int K = (MODE == 0) ? p.KS * 8 : 64;
  const int ksteps = (MODE == 0) ? p.KS / 2 : 4;

  uint32_t wreg[NPW][24];
  load_wregs<NPW>(wreg, reinterpret_cast<const bf16*>(p.W), K, V, w, j, g, tq);
  {
    const uint16_t* Lu = reinterpret_cast<const uint16_t*>(p.Lw);
    uint16_t* Lsu = reinterpret_cast<uint16_t*>(Ls);
    for (int i = threadIdx.x; i < 24 * 64; i += NT) {
      const int c = i >> 6, k = i & 63;
      const int col = (c >> 3) * 64 + 8 * j + (c & 7);
      Lsu[c * 72 + k] = (k < K) ? Lu[(size_t)k * 192 + col] : (uint16_t)0;
    }
    for (int i = threadIdx.x; i < 32 * 40; i += NT) {
      const int n = i / (VP + 8), m = i % (VP + 8);
      float v = 0.f;
      if (n < V && m < V) v = p.S[n * V + m] - (n == m ? 1.f : 0.f);
      if (i < VP * (VP + 8)) Ssm[i] = __float2bfloat16_rn(v);
    }
    if (threadIdx.x < 32) csm[threadIdx.x] = (threadIdx.x < V) ? p.cs[threadIdx.x] : 0.f;
    if (MODE == 1)
      for (int i = threadIdx.x; i < VP * BC * 4; i += NT) sts32(stg + i * 4, 0u);   // rows >= V stay zero
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
  }
  __syncthreads();
  cluster_sync_all();

  uint32_t ph = 0;
  bf16* px = reinterpret_cast<bf16*>(p.px);
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long pt = prof_on ? clock64() : 0;
#define GS_PROF(slot)                                  \
  if (prof_on) {                                       \
    const long long now = clock64();                   \
    atomicAdd(&p.prof[slot], (unsigned long long)(now - pt)); \
    pt = now;                                          \
  }

  if constexpr (MODE == 0) {
    const bf16* xb = reinterpret_cast<const bf16*>(p.xb);
    for (int t = cid / NC; t < T; t += p.tsplit) {
      if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, p.KS * slice_bytes);
        if (j < p.KS)
          bulk_g2s_mc(buf + j * slice_bytes, xb + (((size_t)(t + p.xb_slot0) * NC + nc) * p.xb_slices + j) * slice_el, slice_bytes, bar,
                      (uint16_t)0xff);
      }
      mbar_wait(bar, ph & 1, p.err, 1);
      ++ph;
#pragma unroll
      for (int q = 0; q < NPW; ++q) {
        const int n = w + 8 * q;
        if (n >= V) continue;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float a1[3][4], a2[3][4];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) a1[nt][i] = a2[nt][i] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < ksteps) {
              uint32_t am[4], ap[4];
              ldsm_x4(a_addr(buf, slice_bytes, V, BC, 1, n, mt, ks, lane), am);
              ldsm_x4(a_addr(buf, slice_bytes, V, BC, 0, n, mt, ks, lane), ap);
#pragma unroll
              for (int nt = 0; nt < 3; ++nt) {
                const int wi = nt < 2 ? ks * 4 + nt * 2 : 16 + ks * 2;
                mma16816(a1[nt], am, wreg[q][wi], wreg[q][wi + 1]);
                const uint32_t la = ls + ((nt * 8 + g) * 72 + ks * 16 + 2 * tq) * 2;
                mma16816(a2[nt], ap, lds32(la), lds32(la + 16));
              }
            }
          const float csn = csm[n];
          const float* bgn = p.bg + (size_t)n * 192 + 8 * j + 2 * tq;
          const float* bln = p.bl + 8 * j + 2 * tq;
          float o1[3][4], o2[3][4];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o1[nt][i] = a1[nt][i] + bgn[nt * 64 + (i & 1)];
              o2[nt][i] = csn * a2[nt][i] + bln[nt * 64 + (i & 1)];
            }
          const int item = (w * NPW + q) * MT + mt;
          bf16* dst = px + (((((size_t)t * NC + nc) * CL + j) * ITEMS + item) * 3) * 256 + lane * 8;
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk(o1[0][0], o1[0][1]), pk(o1[0][2], o1[0][3]), pk(o1[1][0], o1[1][1]), pk(o1[1][2], o1[1][3]));
          *reinterpret_cast<uint4*>(dst + 256) = make_uint4(pk(o2[0][0], o2[0][1]), pk(o2[0][2], o2[0][3]), pk(o2[1][0], o2[1][1]), pk(o2[1][2], o2[1][3]));
          *reinterpret_cast<uint4*>(dst + 512) = make_uint4(pk(o1[2][0], o1[2][1]), pk(o1[2][2], o1[2][3]), pk(o2[2][0], o2[2][1]), pk(o2[2][2], o2[2][3]));
        }
      }
      cluster_sync_all();   // every CTA is done reading the block before the next one lands
    }
    return;
  } else {
    bf16* xcg = reinterpret_cast<bf16*>(p.xcg);
    bf16* xcu = reinterpret_cast<bf16*>(p.xcu);
    bf16* fs = reinterpret_cast<bf16*>(p.fs);
    bf16* hout = reinterpret_cast<bf16*>(p.hout);
    uint32_t hprev[NPW][MT][2], zst[NPW][MT][2];
#pragma unroll
    for (int q = 0; q < NPW; ++q)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) hprev[q][mt][0] = hprev[q][mt][1] = zst[q][mt][0] = zst[q][mt][1] = 0u;

    for (int t = 0; t < T; ++t) {
      const size_t blk = ((size_t)t * NC + nc) * CL + j;
      // next step's input half -> L2 while this step runs (the loads below then see L2 latency, not DRAM latency)
      if (t + 1 < T) {
        const char* nx = reinterpret_cast<const char*>(px + (((blk + (size_t)NC * CL) * ITEMS + (size_t)w * NPW * MT) * 3) * 256);
#pragma unroll
        for (int i = 0; i < (NPW * MT * 3 * 4 + 31) / 32; ++i)
          if (lane + 32 * i < NPW * MT * 3 * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)(lane + 32 * i) * 128));
      }
      // ------------------------------------------------ gate half step: z, r (GRU.py:21-22)
      if (t > 0) {
        mbar_wait(bar, ph & 1, p.err, 2);
        ++ph;
      }
      GS_PROF(0)
      {
        bf16* xu = xcu + blk * slice_el;   // this CTA's slice: [pm][V][BC][8]
        const bf16* pxw = px + ((blk * ITEMS + (size_t)w * NPW * MT) * 3) * 256 + lane * 8;   // this warp's items of step t
        uint4 n0 = *reinterpret_cast<const uint4*>(pxw), n1 = *reinterpret_cast<const uint4*>(pxw + 256);
#pragma unroll
        for (int it = 0; it < NPW * MT; ++it) {
          const int q = it / MT, mt = it % MT;
          const int n = w + 8 * q;
          if (n >= V) break;
          {
            const int item = (w * NPW + q) * MT + mt;
            const uint4 c0 = n0, c1 = n1;
            if (it + 1 < NPW * MT && w + 8 * ((it + 1) / MT) < V) {
              n0 = *reinterpret_cast<const uint4*>(pxw + (it + 1) * 768);
              n1 = *reinterpret_cast<const uint4*>(pxw + (it + 1) * 768 + 256);
            }
            float a1[2][4], a2[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int i = 0; i < 4; ++i) a1[nt][i] = a2[nt][i] = 0.f;
            if (t > 0) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                uint32_t am[4], ap[4];
                ldsm_x4(a_addr(buf, slice_bytes, V, BC, 1, n, mt, ks, lane), am);
                ldsm_x4(a_addr(buf, slice_bytes, V, BC, 0, n, mt, ks, lane), ap);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                  mma16816(a1[nt], am, wreg[q][ks * 4 + nt * 2], wreg[q][ks * 4 + nt * 2 + 1]);
                  const uint32_t la = ls + ((nt * 8 + g) * 72 + ks * 16 + 2 * tq) * 2;
                  mma16816(a2[nt], ap, lds32(la), lds32(la + 16));
                }
              }
            }
            const float csn = csm[n];
            float pxz[4], pxr[4], lxz[4], lxr[4], hp[4], zv[4], rv[4], lgz[4], lgr[4];
            unpack4(c0.x, c0.y, pxz); unpack4(c0.z, c0.w, pxr);
            unpack4(c1.x, c1.y, lxz); unpack4(c1.z, c1.w, lxr);
            unpack4(hprev[q][mt][0], hprev[q][mt][1], hp);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              lgz[i] = csn * a2[0][i] + lxz[i];
              lgr[i] = csn * a2[1][i] + lxr[i];
              zv[i] = rb(sigm(a1[0][i] + pxz[i] + lgz[i] * sigm(lgz[i])));
              rv[i] = rb(sigm(a1[1][i] + pxr[i] + lgr[i] * sigm(lgr[i])));
            }
            zst[q][mt][0] = pk(zv[0], zv[1]);
            zst[q][mt][1] = pk(zv[2], zv[3]);
            if (fs) {
              bf16* fsi = fs + ((blk * ITEMS + item) * 4) * 256 + lane * 8;
              *reinterpret_cast<uint4*>(fsi) = make_uint4(zst[q][mt][0], zst[q][mt][1], pk(rv[0], rv[1]), pk(rv[2], rv[3]));
              *reinterpret_cast<uint4*>(fsi + 512) = make_uint4(pk(lgz[0], lgz[1]), pk(lgz[2], lgz[3]), pk(lgr[0], lgr[1]), pk(lgr[2], lgr[3]));
            }
            // r * h_{t-1}: plain slot (shared staging for the mix + global)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t v = pk(rv[2 * half] * hp[2 * half], rv[2 * half + 1] * hp[2 * half + 1]);
              const int row = n * BC + mt * 16 + g + 8 * half;
              sts32(stg + row * 16 + tq * 4, v);
              *reinterpret_cast<uint32_t*>(xu + (size_t)row * 8 + 2 * tq) = v;
            }
          }
        }
        GS_PROF(1)
        __syncthreads();
        mix_slice<MT>(stg, ssm, V, VP, w, lane, xu + (size_t)V * BC * 8);
        GS_PROF(2)
        publish_global();
        GS_PROF(3)
        cluster_sync_all();
        GS_PROF(4)
        if (threadIdx.x == 0) {
          mbar_arrive_expect_tx(bar, CL * slice_bytes);
          bulk_g2s_mc(buf + j * slice_bytes, xu, slice_bytes, bar, (uint16_t)0xff);
        }
      }
      // ------------------------------------------------ candidate half step + state update (GRU.py:23-26)
      mbar_wait(bar, ph & 1, p.err, 3);
      ++ph;
      GS_PROF(5)
      {
        bf16* xg = xcg + (((size_t)(t + 1) * NC + nc) * CL + j) * slice_el;
        const bf16* pxw = px + ((blk * ITEMS + (size_t)w * NPW * MT) * 3) * 256 + lane * 8 + 512;
        uint4 n2 = *reinterpret_cast<const uint4*>(pxw);
#pragma unroll
        for (int it = 0; it < NPW * MT; ++it) {
          const int q = it / MT, mt = it % MT;
          const int n = w + 8 * q;
          if (n >= V) break;
          {
            const int item = (w * NPW + q) * MT + mt;
            const uint4 c2 = n2;
            if (it + 1 < NPW * MT && w + 8 * ((it + 1) / MT) < V) n2 = *reinterpret_cast<const uint4*>(pxw + (it + 1) * 768);
            float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              uint32_t am[4], ap[4];
              ldsm_x4(a_addr(buf, slice_bytes, V, BC, 1, n, mt, ks, lane), am);
              ldsm_x4(a_addr(buf, slice_bytes, V, BC, 0, n, mt, ks, lane), ap);
              mma16816(a1, am, wreg[q][16 + ks * 2], wreg[q][16 + ks * 2 + 1]);
              const uint32_t la = ls + ((16 + g) * 72 + ks * 16 + 2 * tq) * 2;
              mma16816(a2, ap, lds32(la), lds32(la + 16));
            }
            const float csn = csm[n];
            float pxu[4], lxu[4], hp[4], zv[4], hc[4], lu[4], hn[4];
            unpack4(c2.x, c2.y, pxu); unpack4(c2.z, c2.w, lxu);
            unpack4(hprev[q][mt][0], hprev[q][mt][1], hp);
            unpack4(zst[q][mt][0], zst[q][mt][1], zv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              lu[i] = csn * a2[i] + lxu[i];
              hc[i] = rb(tanh_fast(a1[i] + pxu[i] + lu[i] * sigm(lu[i])));
              hn[i] = zv[i] * hp[i] + (1.f - zv[i]) * hc[i];
            }
            if (fs) {
              bf16* fsi = fs + ((blk * ITEMS + item) * 4) * 256 + lane * 8;
              *reinterpret_cast<uint4*>(fsi + 256) = make_uint4(pk(hc[0], hc[1]), pk(hc[2], hc[3]), hprev[q][mt][0], hprev[q][mt][1]);
              *reinterpret_cast<uint4*>(fsi + 768) = make_uint4(pk(lu[0], lu[1]), pk(lu[2], lu[3]), 0u, 0u);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t v = pk(hn[2 * half], hn[2 * half + 1]);
              hprev[q][mt][half] = v;
              const int clip = mt * 16 + g + 8 * half, row = n * BC + clip;
              sts32(stg + row * 16 + tq * 4, v);
              *reinterpret_cast<uint32_t*>(xg + (size_t)row * 8 + 2 * tq) = v;
              const int b = nc * BC + clip;
              if (b < p.B) *reinterpret_cast<uint32_t*>(hout + (((size_t)b * T + t) * V + n) * 64 + 8 * j + 2 * tq) = v;
            }
          }
        }
        GS_PROF(6)
        __syncthreads();
        mix_slice<MT>(stg, ssm, V, VP, w, lane, xg + (size_t)V * BC * 8);
        GS_PROF(7)
        publish_global();
        GS_PROF(8)
        cluster_sync_all();
        GS_PROF(9)
        if (t + 1 < T && threadIdx.x == 0) {
          mbar_arrive_expect_tx(bar, CL * slice_bytes);
          bulk_g2s_mc(buf + j * slice_bytes, xg, slice_bytes, bar, (uint16_t)0xff);
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Backward scan (BPTT of GRU.py:17-26 through both EmbGCN products, one launch per layer). Same ownership as the forward:
// CTA j owns hidden channels 8j..8j+7 of every joint. Per step three exchanges of pre-activation gradients (candidate, z, r -
// 64 channels each, so the two forward-sized state buffers suffice): every CTA stores its slice of (Linear-path, graph-path)
// gradients to the blocked global tensors - which are also what the weight / input / supports gradients are computed from after
// the sweep - and multicasts it to the cluster; the products dG . W_n^T (per-node weights transposed, this CTA's 8 output
// channels, B fragments in registers) give the gradient of the mixed and of the plain stage input; the transposed adjacency mix
// runs on the own slice in shared memory (S^T - I on the tensor core, in place), the identity and Linear-path terms stay fp32.
// ---------------------------------------------------------------------------------------------------------
// stg[m][clip][c] <- sum_n A[m][n] stg[n][clip][c] (A = the bf16 matrix at ssm), in place: each clip column belongs to one warp
template <int MT>
__device__ __forceinline__ void mix_inplace(uint32_t stg, uint32_t ssm, int V, int VP, int w, int lane) {
  constexpr int BC = 16 * MT, CPW = BC / 8;
  const int g = lane >> 2, tq = lane & 3;
  const int nmt = VP >> 4;
  uint32_t af[2][2][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      if (mi < nmt && kk < nmt)
        ldsm_x4(ssm + ((mi * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * (VP + 8) + kk * 16 + (lane >> 4) * 8) * 2, af[mi][kk]);
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int cl = w * CPW + c;
    float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      if (kk < nmt) {
        uint32_t b0, b1;
        ldsm_x2_trans(stg + ((kk * 16 + (lane & 15)) * BC + cl) * 16, b0, b1);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
          if (mi < nmt) mma16816(d[mi], af[mi][kk], b0, b1);
      }
    __syncwarp();
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int n = mi * 16 + g + 8 * half;
        if (mi < nmt && n < V) sts32(stg + (n * BC + cl) * 16 + tq * 4, pk(d[mi][2 * half], d[mi][2 * half + 1]));
      }
  }
}

template <int NPW, int MT>
__global__ void __launch_bounds__(NT, 1) gruscan_bwd_kernel(const GruScanArgs p) {
  constexpr int BC = 16 * MT, ITEMS = NW * NPW * MT;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int V = p.V, VP = (V + 15) & ~15, T = p.T, NC = p.NC;
  const uint32_t slice_bytes = 2u * V * BC * 16;
  const size_t slice_el = (size_t)V * BC * 16;
  const uint32_t buf = smem_u32(smem_raw);
  const uint32_t stg = buf + CL * slice_bytes;
  const uint32_t ls = stg + VP * BC * 16;
  const uint32_t ssm = ls + 24 * 72 * 2;
  const uint32_t csm_a = ssm + 32 * 40 * 2;
  const uint32_t bar = csm_a + 32 * 4;
  bf16* Ssm = reinterpret_cast<bf16*>(smem_raw + (ssm - buf));
  float* csm = reinterpret_cast<float*>(smem_raw + (csm_a - buf));
  const int j = (int)cluster_ctarank();
  const int nc = blockIdx.x / CL;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;

  // B fragments of W_n^T: product X (0 candidate, 1 z, 2 r) reads columns cb(X)..cb(X)+63 of row 8j+g of W[n] (64 x 192)
  uint32_t wreg[NPW][24];
  {
    const uint32_t* W32 = reinterpret_cast<const uint32_t*>(p.W);
#pragma unroll
    for (int q = 0; q < NPW; ++q) {
      const int n = w + 8 * q;
#pragma unroll
      for (int X = 0; X < 3; ++X)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int col = (X == 0 ? 128 : X == 1 ? 0 : 64) + ks * 16 + 2 * tq + 8 * r;
            wreg[q][X * 8 + ks * 2 + r] = n < V ? W32[(((size_t)n * 64 + 8 * j + g) * 192 + col) >> 1] : 0u;
          }
    }
    const uint16_t* Lu = reinterpret_cast<const uint16_t*>(p.Lw);
    uint16_t* Lsu = reinterpret_cast<uint16_t*>(smem_raw + (ls - buf));
    for (int i = threadIdx.x; i < 24 * 64; i += NT) {
      const int c = i >> 6, k = i & 63, X = c >> 3;
      Lsu[c * 72 + k] = Lu[(size_t)(8 * j + (c & 7)) * 192 + (X == 0 ? 128 : X == 1 ? 0 : 64) + k];
    }
    for (int i = threadIdx.x; i < VP * (VP + 8); i += NT) {
      const int m = i / (VP + 8), n = i % (VP + 8);
      float v = 0.f;
      if (n < V && m < V) v = p.S[n * V + m] - (n == m ? 1.f : 0.f);   // transposed: row m of the A operand holds S[:, m]
      Ssm[i] = __float2bfloat16_rn(v);
    }
    if (threadIdx.x < 32) csm[threadIdx.x] = (threadIdx.x < V) ? p.cs[threadIdx.x] : 0.f;
    for (int i = threadIdx.x; i < VP * BC * 4; i += NT) sts32(stg + i * 4, 0u);
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
  }
  __syncthreads();
  cluster_sync_all();

  const bf16* fs = reinterpret_cast<const bf16*>(p.fs);
  const bf16* dho = reinterpret_cast<const bf16*>(p.dhout);
  bf16* dxu = reinterpret_cast<bf16*>(p.dxu);
  bf16* dxgz = reinterpret_cast<bf16*>(p.dxgz);
  bf16* dxgr = reinterpret_cast<bf16*>(p.dxgr);
  float carry[NPW][MT][4];
  uint32_t dpart[NPW][MT][2];
#pragma unroll
  for (int q = 0; q < NPW; ++q)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) carry[q][mt][i] = 0.f;
      dpart[q][mt][0] = dpart[q][mt][1] = 0u;
    }
  uint32_t ph = 0;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long pt = prof_on ? clock64() : 0;

  // products of one phase: graph path (mixed-input gradient before the mix) -> stg, identity + Linear path -> `direct`
  auto send = [&](const bf16* src) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, CL * slice_bytes);
      bulk_g2s_mc(buf + j * slice_bytes, src, slice_bytes, bar, (uint16_t)0xff);
    }
  };

  for (int t = T - 1; t >= 0; --t) {
    const size_t blk = ((size_t)t * NC + nc) * CL + j;
    const bf16* fsw = fs + ((blk * ITEMS + (size_t)w * NPW * MT) * 4) * 256 + lane * 8;
    bf16* su = dxu + blk * slice_el;
    bf16* sz = dxgz + blk * slice_el;
    bf16* sr = dxgr + blk * slice_el;
    if (t > 0) {   // previous step's saved gate values -> L2
      const char* nx = reinterpret_cast<const char*>(fs + (((blk - (size_t)NC * CL) * ITEMS + (size_t)w * NPW * MT) * 4) * 256);
#pragma unroll
      for (int i = 0; i < (NPW * MT * 4 * 4 + 31) / 32; ++i)
        if (lane + 32 * i < NPW * MT * 4 * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)(lane + 32 * i) * 128));
    }
    // ------------------------------------------------ A: state update and candidate backward (elementwise)
    // (loads of item it+1 are issued before the maths of item it; the previous step's dH rows were prefetched to L2)
    auto load_dh = [&](int it, int tt, uint32_t (&d)[2]) {
      const int q = it / MT, mt = it % MT, n = w + 8 * q;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int b = nc * BC + mt * 16 + g + 8 * half;
        d[half] = (b < p.B && n < V)
                      ? *reinterpret_cast<const uint32_t*>(dho + (size_t)b * p.dh_b + (size_t)tt * p.dh_t + (size_t)n * p.dh_v + 8 * j + 2 * tq)
                      : 0u;
      }
    };
    if (t > 0 && tq == 0) {
#pragma unroll
      for (int it = 0; it < NPW * MT; ++it) {
        const int q = it / MT, mt = it % MT, n = w + 8 * q;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int b = nc * BC + mt * 16 + g + 8 * half;
          if (b < p.B && n < V)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(dho + (size_t)b * p.dh_b + (size_t)(t - 1) * p.dh_t + (size_t)n * p.dh_v + 8 * j));
        }
      }
    }
    uint32_t nd[2];
    uint2 nz = *reinterpret_cast<const uint2*>(fsw), nlz = *reinterpret_cast<const uint2*>(fsw + 512), nlu = *reinterpret_cast<const uint2*>(fsw + 768);
    uint4 nh = *reinterpret_cast<const uint4*>(fsw + 256);
    load_dh(0, t, nd);
#pragma unroll
    for (int it = 0; it < NPW * MT; ++it) {
      const int q = it / MT, mt = it % MT;
      const int n = w + 8 * q;
      if (n >= V) break;
      const uint2 cz = nz, clz = nlz, clu = nlu;
      const uint4 c1 = nh;
      const uint32_t cd[2] = {nd[0], nd[1]};
      if (it + 1 < NPW * MT && w + 8 * ((it + 1) / MT) < V) {
        nz = *reinterpret_cast<const uint2*>(fsw + (it + 1) * 1024);
        nh = *reinterpret_cast<const uint4*>(fsw + (it + 1) * 1024 + 256);
        nlz = *reinterpret_cast<const uint2*>(fsw + (it + 1) * 1024 + 512);
        nlu = *reinterpret_cast<const uint2*>(fsw + (it + 1) * 1024 + 768);
        load_dh(it + 1, t, nd);
      }
      float z[4], hc[4], hp[4], lgz[4], lu[4], g1u[4], g2u[4], g1z[4], g2z[4];
      unpack4(cz.x, cz.y, z); unpack4(c1.x, c1.y, hc); unpack4(c1.z, c1.w, hp); unpack4(clz.x, clz.y, lgz); unpack4(clu.x, clu.y, lu);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        carry[q][mt][2 * half] += lo(cd[half]);
        carry[q][mt][2 * half + 1] += hi(cd[half]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dh = carry[q][mt][i];
        const float dz = dh * (hp[i] - hc[i]);
        const float dpu = dh * (1.f - z[i]) * (1.f - hc[i] * hc[i]);
        carry[q][mt][i] = dh * z[i];
        g1u[i] = dpu;
        g2u[i] = dpu * dsilu(lu[i]);
        g1z[i] = dz * z[i] * (1.f - z[i]);
        g2z[i] = g1z[i] * dsilu(lgz[i]);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const size_t row = (size_t)n * BC + mt * 16 + g + 8 * half;
        *reinterpret_cast<uint32_t*>(su + row * 8 + 2 * tq) = pk(g2u[2 * half], g2u[2 * half + 1]);                        // pm 0: Linear path
        *reinterpret_cast<uint32_t*>(su + (size_t)V * BC * 8 + row * 8 + 2 * tq) = pk(g1u[2 * half], g1u[2 * half + 1]);   // pm 1: graph path
        *reinterpret_cast<uint32_t*>(sz + row * 8 + 2 * tq) = pk(g2z[2 * half], g2z[2 * half + 1]);
        *reinterpret_cast<uint32_t*>(sz + (size_t)V * BC * 8 + row * 8 + 2 * tq) = pk(g1z[2 * half], g1z[2 * half + 1]);
      }
    }
    GS_PROF(0)
    publish_global();
    cluster_sync_all();
    send(su);
    GS_PROF(1)
    // ------------------------------------------------ B, C, D: products of the candidate, z and r gradients
#pragma unroll 1
    for (int X = 0; X < 3; ++X) {
      mbar_wait(bar, ph & 1, p.err, 4 + X);
      ++ph;
      GS_PROF(2)
#pragma unroll
      for (int it = 0; it < NPW * MT; ++it) {
        const int q = it / MT, mt = it % MT;
        const int n = w + 8 * q;
        if (n >= V) break;
        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t am[4], ap[4];
          ldsm_x4(a_addr(buf, slice_bytes, V, BC, 1, n, mt, ks, lane), am);
          ldsm_x4(a_addr(buf, slice_bytes, V, BC, 0, n, mt, ks, lane), ap);
          const uint32_t w0 = X == 0 ? wreg[q][ks * 2] : X == 1 ? wreg[q][8 + ks * 2] : wreg[q][16 + ks * 2];
          const uint32_t w1 = X == 0 ? wreg[q][ks * 2 + 1] : X == 1 ? wreg[q][8 + ks * 2 + 1] : wreg[q][16 + ks * 2 + 1];
          mma16816(a1, am, w0, w1);
          const uint32_t la = ls + ((X * 8 + g) * 72 + ks * 16 + 2 * tq) * 2;
          mma16816(a2, ap, lds32(la), lds32(la + 16));
        }
        const float csn = csm[n];
        float dir[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dir[i] = a1[i] + csn * a2[i];
        if (X == 0) {
          const uint4 c0 = *reinterpret_cast<const uint4*>(fsw + it * 1024), c1 = *reinterpret_cast<const uint4*>(fsw + it * 1024 + 256);
          float r[4], hp[4], part[4];
          unpack4(c0.z, c0.w, r); unpack4(c1.z, c1.w, hp);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            carry[q][mt][i] += dir[i] * r[i];
            part[i] = dir[i] * hp[i] * r[i] * (1.f - r[i]);
          }
          dpart[q][mt][0] = pk(part[0], part[1]);
          dpart[q][mt][1] = pk(part[2], part[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) carry[q][mt][i] += dir[i];
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) sts32(stg + ((n * BC + mt * 16 + g + 8 * half) * 16) + tq * 4, pk(a1[2 * half], a1[2 * half + 1]));
      }
      GS_PROF(3)
      __syncthreads();
      mix_inplace<MT>(stg, ssm, V, VP, w, lane);
      __syncthreads();
      GS_PROF(4)
#pragma unroll
      for (int it = 0; it < NPW * MT; ++it) {
        const int q = it / MT, mt = it % MT;
        const int n = w + 8 * q;
        if (n >= V) break;
        float mx[4];
        unpack4(lds32(stg + ((n * BC + mt * 16 + g) * 16) + tq * 4), lds32(stg + ((n * BC + mt * 16 + g + 8) * 16) + tq * 4), mx);
        if (X == 0) {
          const uint4 c0 = *reinterpret_cast<const uint4*>(fsw + it * 1024), c1 = *reinterpret_cast<const uint4*>(fsw + it * 1024 + 256),
                      c2 = *reinterpret_cast<const uint4*>(fsw + it * 1024 + 512);
          float r[4], hp[4], lgr[4], part[4], g1r[4], g2r[4];
          unpack4(c0.z, c0.w, r); unpack4(c1.z, c1.w, hp); unpack4(c2.z, c2.w, lgr);
          unpack4(dpart[q][mt][0], dpart[q][mt][1], part);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            carry[q][mt][i] += mx[i] * r[i];
            g1r[i] = part[i] + mx[i] * hp[i] * r[i] * (1.f - r[i]);
            g2r[i] = g1r[i] * dsilu(lgr[i]);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const size_t row = (size_t)n * BC + mt * 16 + g + 8 * half;
            *reinterpret_cast<uint32_t*>(sr + row * 8 + 2 * tq) = pk(g2r[2 * half], g2r[2 * half + 1]);
            *reinterpret_cast<uint32_t*>(sr + (size_t)V * BC * 8 + row * 8 + 2 * tq) = pk(g1r[2 * half], g1r[2 * half + 1]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) carry[q][mt][i] += mx[i];
        }
      }
      GS_PROF(5)
      if (X < 2) {
        if (X == 0) publish_global();
        cluster_sync_all();
        send(X == 0 ? sz : sr);
        GS_PROF(6)
      }
      // after X == 2 the next step's phase A ends with the cluster barrier that orders the reuse of buf and stg
    }
  }
  cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------
// Export of the blocked / fragment-order tensors to the row-major tensors the batched GEMMs (weight gradients,
// input gradients, dS) read: XC std = [2: mixed, plain][T][B][V][Cp] with the cell-input layout [h 64 | x Din | 1 | 0].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) export_xc_kernel(const bf16* __restrict__ xc, const bf16* __restrict__ xb, bf16* __restrict__ out,
                                                        int T, int B, int V, int NC, int BC, int KS, int xb_slices, int xb_slot0,
                                                        int Din, int Cp) {
  const int t = blockIdx.x / NC, nc = blockIdx.x % NC;
  const int C8 = Cp / 8;
  const size_t slice_el = (size_t)V * BC * 16;
  const int total = 2 * V * BC * C8;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int k = i % C8;
    int r = i / C8;
    const int clip = r % BC;
    r /= BC;
    const int n = r % V, pmstd = r / V;
    const int b = nc * BC + clip;
    if (b >= B) continue;
    const int pm = 1 - pmstd;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k < 8) {
      v = *reinterpret_cast<const uint4*>(xc + (((size_t)t * NC + nc) * CL + k) * slice_el + ((size_t)(pm * V + n) * BC + clip) * 8);
    } else {
      const int kx = k - 8;
      if (kx < KS)
        v = *reinterpret_cast<const uint4*>(xb + (((size_t)(t + xb_slot0) * NC + nc) * xb_slices + kx) * slice_el +
                                            ((size_t)(pm * V + n) * BC + clip) * 8);
      uint16_t* e = reinterpret_cast<uint16_t*>(&v);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int col = kx * 8 + q;
        if (col == Din) e[q] = 0x3f80;   // the constant-1 column that carries the bias row (never mixed)
        else if (col > Din) e[q] = 0;
      }
    }
    *reinterpret_cast<uint4*>(out + ((((size_t)pmstd * T + t) * B + b) * V + n) * Cp + k * 8) = v;
  }
}

// blocked pre-activation gradients -> dPLu [2: graph, Linear][T][B][V][64], dPLg [2][T][B][V][128] (z | r)
__global__ void __launch_bounds__(256) export_dg_kernel(const bf16* __restrict__ dxu, const bf16* __restrict__ dxgz, const bf16* __restrict__ dxgr,
                                                        bf16* __restrict__ dPLu, bf16* __restrict__ dPLg, int T, int B, int V, int NC, int BC) {
  const int t = blockIdx.x / NC, nc = blockIdx.x % NC;
  const size_t slice_el = (size_t)V * BC * 16;
  const int total = 2 * V * BC * 24;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int k = i % 24;
    int r = i / 24;
    const int clip = r % BC;
    r /= BC;
    const int n = r % V, pstd = r / V;
    const int b = nc * BC + clip;
    if (b >= B) continue;
    const int pm = 1 - pstd;
    const bf16* src = k < 8 ? dxu : k < 16 ? dxgz : dxgr;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (((size_t)t * NC + nc) * CL + (k & 7)) * slice_el + ((size_t)(pm * V + n) * BC + clip) * 8);
    const size_t row = (((size_t)pstd * T + t) * B + b) * V + n;
    if (k < 8) *reinterpret_cast<uint4*>(dPLu + row * 64 + k * 8) = v;
    else *reinterpret_cast<uint4*>(dPLg + row * 128 + (k - 8) * 8) = v;
  }
}

// dX[b][t][n][c] = sum_m S[m][n] (dXg[0] + dXu[0])[t][b][m][c0 + c] + (dXg[1] + dXu[1])[t][b][n][c0 + c]: the transposed adjacency
// mix of the graph-path input gradients plus the Linear-path ones, both stages, one block per (t, clip)
__global__ void __launch_bounds__(256) mix_dx_kernel(const bf16* __restrict__ dXg, const bf16* __restrict__ dXu, const float* __restrict__ S,
                                                     bf16* __restrict__ dX, int T, int B, int V, int Cp, int c0, int Din) {
  extern __shared__ float sm_dx[];
  float* Ss = sm_dx;               // [V][V]
  float* gs_ = sm_dx + V * V;      // [V][Din]
  const int t = blockIdx.x / B, b = blockIdx.x % B;
  const size_t path = (size_t)T * B * V * Cp, row0 = ((size_t)t * B + b) * V;
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) Ss[i] = S[i];
  const int D8 = Din >> 3;
  for (int i = threadIdx.x; i < V * D8; i += blockDim.x) {
    const int m = i / D8, c = (i % D8) * 8;
    float a[8], e[8];
    load8(dXg + (row0 + m) * Cp + c0 + c, a);
    load8(dXu + (row0 + m) * Cp + c0 + c, e);
#pragma unroll
    for (int q = 0; q < 8; ++q) gs_[m * Din + c + q] = a[q] + e[q];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * D8; i += blockDim.x) {
    const int n = i / D8, c = (i % D8) * 8;
    float a[8], e[8], o[8];
    load8(dXg + path + (row0 + n) * Cp + c0 + c, a);
    load8(dXu + path + (row0 + n) * Cp + c0 + c, e);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = a[q] + e[q];
    for (int m = 0; m < V; ++m) {
      const float sv = Ss[m * V + n];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] += sv * gs_[m * Din + c + q];
    }
    store8(dX + (((size_t)b * T + t) * V + n) * Din + c, o);
  }
}

// FS (fragment order) -> ZR, LG [T][B][V][128], HC, LU [T][B][V][64]
__global__ void __launch_bounds__(256) export_fs_kernel(const bf16* __restrict__ fs, bf16* __restrict__ ZR, bf16* __restrict__ LG,
                                                        bf16* __restrict__ HC, bf16* __restrict__ LU, int T, int B, int V, int NC, int NPW,
                                                        int MT) {
  const int blk = blockIdx.x, j = blk % CL, nc = (blk / CL) % NC, t = blk / (CL * NC);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int BC = 16 * MT, ITEMS = NW * NPW * MT;
  for (int q = 0; q < NPW; ++q) {
    const int n = w + 8 * q;
    if (n >= V) continue;
    for (int mt = 0; mt < MT; ++mt) {
      const int item = (w * NPW + q) * MT + mt;
      const bf16* f = fs + (((size_t)blk * ITEMS + item) * 4) * 256 + lane * 8;
      const uint4 c0 = *reinterpret_cast<const uint4*>(f), c1 = *reinterpret_cast<const uint4*>(f + 256),
                  c2 = *reinterpret_cast<const uint4*>(f + 512), c3 = *reinterpret_cast<const uint4*>(f + 768);
      for (int half = 0; half < 2; ++half) {
        const int b = nc * BC + mt * 16 + g + 8 * half;
        if (b >= B) continue;
        const size_t row = ((size_t)t * B + b) * V + n;
        const int c = 8 * j + 2 * tq;
        *reinterpret_cast<uint32_t*>(ZR + row * 128 + c) = half ? c0.y : c0.x;
        *reinterpret_cast<uint32_t*>(ZR + row * 128 + 64 + c) = half ? c0.w : c0.z;
        *reinterpret_cast<uint32_t*>(LG + row * 128 + c) = half ? c2.y : c2.x;
        *reinterpret_cast<uint32_t*>(LG + row * 128 + 64 + c) = half ? c2.w : c2.z;
        *reinterpret_cast<uint32_t*>(HC + row * 64 + c) = half ? c1.y : c1.x;
        *reinterpret_cast<uint32_t*>(LU + row * 64 + c) = half ? c3.y : c3.x;
      }
    }
  }
}

template <int NPW, int MT, int MODE>
struct ScanKernel {
  static constexpr auto fn = gruscan_kernel<NPW, MT, MODE>;
};
template <int NPW, int MT>
struct ScanKernel<NPW, MT, 2> {
  static constexpr auto fn = gruscan_bwd_kernel<NPW, MT>;
};

template <int NPW, int MT, int MODE>
int launch_scan(const GruScanArgs& a, cudaStream_t stream) {
  auto kern = ScanKernel<NPW, MT, MODE>::fn;
  const size_t smem = smem_bytes<MT>(a.V);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("gruscan: smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
    return FMM_ERR_SMEM;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.NC * (MODE == 0 ? a.tsplit : 1) * CL));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    set_last_error("gruscan: launch (mode %d, %zu bytes smem): %s", MODE, smem, cudaGetErrorString(e));
    return FMM_ERR_CUDA;
  }
  return FMM_OK;
}

}  // namespace gs
}  // namespace fmm

extern "C" {

// clips per cluster (BC) and joint slots per warp (NPW) the kernels use for V joints; 0 when V is not supported
int fmm_gruscan_geometry(int V, int* BC, int* NPW) {
  if (V < 1 || V > 32) return 0;
  if (BC) *BC = V <= 25 ? 32 : 16;
  if (NPW) *NPW = V <= 16 ? 2 : 4;
  return 1;
}

// how many clusters of the forward scan can be resident at once on the current device (cudaOccupancyMaxActiveClusters)
int fmm_gruscan_max_clusters(int V) {
  using namespace fmm;
  if (V < 1 || V > 32) return 0;
  int n = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gs::CL * 64);
  cfg.blockDim = dim3(gs::NT);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = gs::CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e;
#define FMM_GS_OCC(NPW, MT)                                                                                   \
  do {                                                                                                        \
    cfg.dynamicSmemBytes = gs::smem_bytes<MT>(V);                                                             \
    cudaFuncSetAttribute(gs::gruscan_kernel<NPW, MT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes); \
    e = cudaOccupancyMaxActiveClusters(&n, gs::gruscan_kernel<NPW, MT, 1>, &cfg);                            \
  } while (0)
  if (V <= 16) FMM_GS_OCC(2, 2);
  else if (V <= 25) FMM_GS_OCC(4, 2);
  else FMM_GS_OCC(4, 1);
#undef FMM_GS_OCC
  if (e != cudaSuccess) {
    set_last_error("gruscan_max_clusters: %s", cudaGetErrorString(e));
    return -1;
  }
  return n;
}

// mode 0: input half of both EmbGCN products for all steps; mode 1: forward scan; mode 2: backward scan
int fmm_gruscan(const fmm::GruScanArgs* a, int mode, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(a && a->V >= 1 && a->V <= 32 && a->T >= 1 && a->B >= 1, "gruscan: bad sizes");
  const int BC = a->V <= 25 ? 32 : 16;
  FMM_CHECK_ARG(a->NC == (a->B + BC - 1) / BC, "gruscan: NC must be ceil(B / %d)", BC);
  FMM_CHECK_ARG(mode >= 0 && mode <= 2, "gruscan: mode %d", mode);
  if (mode == 2) FMM_CHECK_ARG(a->fs && a->dhout && a->dxu && a->dxgz && a->dxgr, "gruscan: backward needs fs, dhout, dxu, dxgz, dxgr");
  if (mode == 0) FMM_CHECK_ARG(a->KS >= 2 && a->KS <= 8 && a->KS % 2 == 0 && a->tsplit >= 1 && a->xb_slices >= a->KS, "gruscan: bad xpart geometry");
  int rc;
#define FMM_GS(NPW, MT) \
  (mode == 0 ? gs::launch_scan<NPW, MT, 0>(*a, stream) : mode == 1 ? gs::launch_scan<NPW, MT, 1>(*a, stream) : gs::launch_scan<NPW, MT, 2>(*a, stream))
  if (a->V <= 16) rc = FMM_GS(2, 2);
  else if (a->V <= 25) rc = FMM_GS(4, 2);
  else rc = FMM_GS(4, 1);
#undef FMM_GS
  if (rc != FMM_OK) return rc;
  FMM_CHECK_LAUNCH("gruscan");
  return FMM_OK;
}

int fmm_gruscan_export_xc(const void* xc, const void* xb, void* out, int T, int B, int V, int KS, int xb_slices, int xb_slot0, int Din,
                          int Cp, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(V >= 1 && V <= 32 && Cp % 8 == 0 && Cp >= 64 + Din + 1, "gruscan_export_xc: bad sizes");
  const int BC = V <= 25 ? 32 : 16, NC = (B + BC - 1) / BC;
  gs::export_xc_kernel<<<T * NC, 256, 0, stream>>>(reinterpret_cast<const gs::bf16*>(xc), reinterpret_cast<const gs::bf16*>(xb),
                                                    reinterpret_cast<gs::bf16*>(out), T, B, V, NC, BC, KS, xb_slices, xb_slot0, Din, Cp);
  FMM_CHECK_LAUNCH("gruscan_export_xc");
  return FMM_OK;
}

int fmm_gruscan_export_dg(const void* dxu, const void* dxgz, const void* dxgr, void* dPLu, void* dPLg, int T, int B, int V, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(V >= 1 && V <= 32, "gruscan_export_dg: bad sizes");
  const int BC = V <= 25 ? 32 : 16, NC = (B + BC - 1) / BC;
  gs::export_dg_kernel<<<T * NC, 256, 0, stream>>>(reinterpret_cast<const gs::bf16*>(dxu), reinterpret_cast<const gs::bf16*>(dxgz),
                                                    reinterpret_cast<const gs::bf16*>(dxgr), reinterpret_cast<gs::bf16*>(dPLu),
                                                    reinterpret_cast<gs::bf16*>(dPLg), T, B, V, NC, BC);
  FMM_CHECK_LAUNCH("gruscan_export_dg");
  return FMM_OK;
}

// dX (B,T,V,Din) from the input gradients of both stages, dXg / dXu (2,T,B,V,Cp): columns c0..c0+Din-1, Din % 8 == 0
int fmm_gruscan_mix_dx(const void* dXg, const void* dXu, const float* S, void* dX, int T, int B, int V, int Cp, int c0, int Din,
                       cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(V >= 1 && V <= 64 && Din % 8 == 0 && c0 % 8 == 0 && c0 + Din <= Cp && Cp % 8 == 0, "gruscan_mix_dx: bad sizes");
  const size_t smem = (size_t)(V * V + V * Din) * sizeof(float);
  gs::mix_dx_kernel<<<T * B, 256, smem, stream>>>(reinterpret_cast<const gs::bf16*>(dXg), reinterpret_cast<const gs::bf16*>(dXu), S,
                                                  reinterpret_cast<gs::bf16*>(dX), T, B, V, Cp, c0, Din);
  FMM_CHECK_LAUNCH("gruscan_mix_dx");
  return FMM_OK;
}

int fmm_gruscan_export_fs(const void* fs, void* ZR, void* LG, void* HC, void* LU, int T, int B, int V, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(V >= 1 && V <= 32, "gruscan_export_fs: bad sizes");
  const int BC = V <= 25 ? 32 : 16, NC = (B + BC - 1) / BC, NPW = V <= 16 ? 2 : 4, MT = BC / 16;
  gs::export_fs_kernel<<<T * NC * gs::CL, 256, 0, stream>>>(reinterpret_cast<const gs::bf16*>(fs), reinterpret_cast<gs::bf16*>(ZR),
                                                            reinterpret_cast<gs::bf16*>(LG), reinterpret_cast<gs::bf16*>(HC),
                                                            reinterpret_cast<gs::bf16*>(LU), T, B, V, NC, NPW, MT);
  FMM_CHECK_LAUNCH("gruscan_export_fs");
  return FMM_OK;
}

}  // extern "C"
