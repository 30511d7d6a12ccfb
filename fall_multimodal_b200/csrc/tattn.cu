// Time-axis attention of the TRAGCN transformer layers (reference TA.py:55-62: A = softmax(q k^T / sqrt(c), -1), A v per
// (clip, joint)) as ONE flash-style kernel per direction: the (B,V,T,T) score tensor is never written.
//
// One CTA per (clip, joint) head, T <= 320 so q, k, v (and dO) of the head live in shared memory for the whole CTA; warps own
// 16-row tiles. q and k arrive FEATURE-major ((B,F,V,Tp) as the time-as-channel convolutions of TA.py:42-44 leave them), which is
// exactly the k-major storage ldmatrix.trans turns into A / B fragments - no transposes anywhere. bf16 mma.sync.m16n8k16, fp32
// accumulators, online softmax in the exp2 domain, the probability tile goes from the accumulator layout straight into the A
// fragments of the next product.
//   forward : out (B,T,V,64), lse2 (B*V,Tp) = log2-domain log-sum-exp of every query row
//   backward: pass 1, warp = 16 keys  : S^T, dP^T recomputed per 16-query tile -> dV (B,V,T,64), dK (B,F,V,Tp)
//             pass 2, warp = 16 queries: S, dP recomputed per 16-key tile      -> dQ (B,F,V,Tp)
#include "common.cuh"
#include "ptx.cuh"

namespace fmm {

struct TAttnArgs {
  const void* q; const void* k; const void* v;   // (B,F,V,Tp), (B,F,V,Tp), (B,V,T,64) bf16
  void* out;                                     // (B,T,V,64) bf16
  float* lse;                                    // (B*V,Tp) fp32
  const void* dout;                              // (B,T,V,64) bf16
  void* dq; void* dk; void* dv;                  // (B,F,V,Tp), (B,F,V,Tp), (B,V,T,64) bf16
  int B, V, T, Tp, F;
  float scale;
  int v_btvc;                                    // 1: v and dv are (B,T,V,64) like out (the value projection is a plain Linear layer then)
};

namespace ta {

typedef __nv_bfloat16 bf16;
constexpr int NW = 10, NT = NW * 32, C = 64, VSTR = 72;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
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
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void sts16z(uint32_t dst) { asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory"); }

// A fragment (rows m0..m0+15, k0..k0+15) from k-major storage base[k][m]
__device__ __forceinline__ void lda_kmajor(uint32_t base, int str, int m0, int k0, int lane, uint32_t (&a)[4]) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4_trans(base + ((k0 + (mi >> 1) * 8 + r) * str + m0 + (mi & 1) * 8) * 2, a);
}
// A fragment from row-major storage base[m][k]
__device__ __forceinline__ void lda_rowmajor(uint32_t base, int str, int m0, int k0, int lane, uint32_t (&a)[4]) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4(base + ((m0 + (mi & 1) * 8 + r) * str + k0 + (mi >> 1) * 8) * 2, a);
}
// B fragments of two n-tiles (n0..n0+15) x (k0..k0+15): b[0],b[1] -> n0..n0+7, b[2],b[3] -> n0+8..n0+15
__device__ __forceinline__ void ldb_kmajor(uint32_t base, int str, int k0, int n0, int lane, uint32_t (&b)[4]) {   // base[k][n]
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4_trans(base + ((k0 + (mi & 1) * 8 + r) * str + n0 + (mi >> 1) * 8) * 2, b);
}
__device__ __forceinline__ void ldb_nmajor(uint32_t base, int str, int k0, int n0, int lane, uint32_t (&b)[4]) {   // base[n][k]
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4(base + ((n0 + (mi >> 1) * 8 + r) * str + k0 + (mi & 1) * 8) * 2, b);
}

// feature-major operand of one head -> smem [64][qstr] (rows >= F zero)
__device__ __forceinline__ void load_fmajor(uint32_t dst, const bf16* src, int F, int V, int Tp, int qstr) {
  const int c8 = Tp >> 3;
  for (int i = threadIdx.x; i < 64 * c8; i += NT) {
    const int f = i / c8, c = i % c8;
    const uint32_t d = dst + (f * qstr + c * 8) * 2;
    if (f < F) cp16(d, src + (size_t)f * V * Tp + c * 8);
    else sts16z(d);
  }
}
// row-major (t, 64) operand with a row stride -> smem [Tp][72] (rows >= T zero)
__device__ __forceinline__ void load_rows(uint32_t dst, const bf16* src, long long row_stride, int T, int Tp) {
  for (int i = threadIdx.x; i < Tp * 8; i += NT) {
    const int t = i >> 3, c = i & 7;
    const uint32_t d = dst + (t * VSTR + c * 8) * 2;
    if (t < T) cp16(d, src + (size_t)t * row_stride + c * 8);
    else sts16z(d);
  }
}

__global__ void __launch_bounds__(NT, 1) tattn_fwd_kernel(const TAttnArgs p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int T = p.T, Tp = p.Tp, V = p.V, F = p.F, qstr = Tp + 8;
  const uint32_t Qs = smem_u32(smem_raw), Ks = Qs + 64 * qstr * 2, Vs = Ks + 64 * qstr * 2;
  const int h = blockIdx.x, b = h / V, v = h % V;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const bf16* qh = reinterpret_cast<const bf16*>(p.q) + ((size_t)b * F * V + v) * Tp;
  const bf16* kh = reinterpret_cast<const bf16*>(p.k) + ((size_t)b * F * V + v) * Tp;
  load_fmajor(Qs, qh, F, V, Tp, qstr);
  load_fmajor(Ks, kh, F, V, Tp, qstr);
  if (p.v_btvc) load_rows(Vs, reinterpret_cast<const bf16*>(p.v) + (size_t)b * T * V * C + (size_t)v * C, (long long)V * C, T, Tp);
  else load_rows(Vs, reinterpret_cast<const bf16*>(p.v) + (size_t)h * T * C, C, T, Tp);
  cp_wait_all();
  __syncthreads();
  const float sc2 = p.scale * 1.4426950408889634f;
  bf16* out = reinterpret_cast<bf16*>(p.out);
  for (int qt = w; qt < (Tp >> 4); qt += NW) {
    const int m0 = qt * 16;
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lda_kmajor(Qs, qstr, m0, ks * 16, lane, qa[ks]);
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[nt][i] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    for (int kb = 0; kb < (Tp >> 6); ++kb) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bb[4];
          ldb_kmajor(Ks, qstr, ks * 16, kb * 64 + np * 16, lane, bb);
          mma16816(s[2 * np], qa[ks], bb[0], bb[1]);
          mma16816(s[2 * np + 1], qa[ks], bb[2], bb[3]);
        }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int key = kb * 64 + nt * 8 + 2 * tq + (i & 1);
          s[nt][i] = key < T ? s[nt][i] * sc2 : -INFINITY;
          mx[i >> 1] = fmaxf(mx[i >> 1], s[nt][i]);
        }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mx[half] = fmaxf(mx[half], __shfl_xor_sync(0xffffffffu, mx[half], 1));
        mx[half] = fmaxf(mx[half], __shfl_xor_sync(0xffffffffu, mx[half], 2));
        const float mnew = fmaxf(mrow[half], mx[half]);
        const float corr = exp2f(mrow[half] - mnew);
        mrow[half] = mnew;
        float ls = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          s[nt][2 * half] = exp2f(s[nt][2 * half] - mnew);
          s[nt][2 * half + 1] = exp2f(s[nt][2 * half + 1] - mnew);
          ls += s[nt][2 * half] + s[nt][2 * half + 1];
          o[nt][2 * half] *= corr;
          o[nt][2 * half + 1] *= corr;
        }
        lrow[half] = lrow[half] * corr + ls;
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const uint32_t pa[4] = {pk(s[2 * k2][0], s[2 * k2][1]), pk(s[2 * k2][2], s[2 * k2][3]), pk(s[2 * k2 + 1][0], s[2 * k2 + 1][1]),
                                pk(s[2 * k2 + 1][2], s[2 * k2 + 1][3])};
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bb[4];
          ldb_kmajor(Vs, VSTR, kb * 64 + k2 * 16, np * 16, lane, bb);
          mma16816(o[2 * np], pa, bb[0], bb[1]);
          mma16816(o[2 * np + 1], pa, bb[2], bb[3]);
        }
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float l = lrow[half];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv = 1.f / l;
      const int t = m0 + g + 8 * half;
      if (t < T) {
        bf16* orow = out + (((size_t)b * T + t) * V + v) * C + 2 * tq;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(orow + nt * 8) = pk(o[nt][2 * half] * inv, o[nt][2 * half + 1] * inv);
      }
      if (tq == 0 && p.lse) p.lse[(size_t)h * Tp + t] = mrow[half] + log2f(l);
    }
  }
}

__global__ void __launch_bounds__(NT, 1) tattn_bwd_kernel(const TAttnArgs p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int T = p.T, Tp = p.Tp, V = p.V, F = p.F, qstr = Tp + 8;
  const uint32_t Qs = smem_u32(smem_raw), Ks = Qs + 64 * qstr * 2, Vs = Ks + 64 * qstr * 2, dOs = Vs + Tp * VSTR * 2;
  const uint32_t lse_a = dOs + Tp * VSTR * 2, del_a = lse_a + Tp * 4, stg_a = del_a + Tp * 4;
  float* lse_s = reinterpret_cast<float*>(smem_raw + (lse_a - Qs));
  float* del_s = reinterpret_cast<float*>(smem_raw + (del_a - Qs));
  const int h = blockIdx.x, b = h / V, v = h % V;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  uint16_t* stg = reinterpret_cast<uint16_t*>(smem_raw + (stg_a - Qs)) + w * 64 * 24;   // this warp's [64 f][16 + 8] transpose tile
  const bf16* qh = reinterpret_cast<const bf16*>(p.q) + ((size_t)b * F * V + v) * Tp;
  const bf16* kh = reinterpret_cast<const bf16*>(p.k) + ((size_t)b * F * V + v) * Tp;
  const bf16* doh = reinterpret_cast<const bf16*>(p.dout) + (size_t)b * T * V * C + (size_t)v * C;
  const bf16* oh = reinterpret_cast<const bf16*>(p.out) + (size_t)b * T * V * C + (size_t)v * C;
  load_fmajor(Qs, qh, F, V, Tp, qstr);
  load_fmajor(Ks, kh, F, V, Tp, qstr);
  if (p.v_btvc) load_rows(Vs, reinterpret_cast<const bf16*>(p.v) + (size_t)b * T * V * C + (size_t)v * C, (long long)V * C, T, Tp);
  else load_rows(Vs, reinterpret_cast<const bf16*>(p.v) + (size_t)h * T * C, C, T, Tp);
  load_rows(dOs, doh, (long long)V * C, T, Tp);
  for (int t = threadIdx.x; t < Tp; t += NT) lse_s[t] = p.lse[(size_t)h * Tp + t];
  cp_wait_all();
  __syncthreads();
  // delta[t] = sum_c dO[t][c] O[t][c]
  for (int t = threadIdx.x; t < Tp; t += NT) {
    float d = 0.f;
    if (t < T) {
      const uint4* orow = reinterpret_cast<const uint4*>(oh + (size_t)t * V * C);
      const uint4* drow = reinterpret_cast<const uint4*>(smem_raw + (dOs - Qs) + (size_t)t * VSTR * 2);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 a = orow[c], e = drow[c];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ew[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          d += __uint_as_float(aw[i] << 16) * __uint_as_float(ew[i] << 16) + __uint_as_float(aw[i] & 0xffff0000u) * __uint_as_float(ew[i] & 0xffff0000u);
      }
    }
    del_s[t] = d;
  }
  __syncthreads();
  const float sc2 = p.scale * 1.4426950408889634f, sc = p.scale;
  const int ntile = Tp >> 4;

  // writes a (16 x 64) accumulator tile (rows = time positions t0.., columns = features) to a feature-major (B,F,V,Tp) tensor
  auto store_fmajor = [&](const float (&acc)[8][4], bf16* dst_head, int t0) {
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 hv = __float2bfloat16_rn(acc[nt][i]);
        stg[(nt * 8 + 2 * tq + (i & 1)) * 24 + g + 8 * (i >> 1)] = *reinterpret_cast<const uint16_t*>(&hv);
      }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int f = lane + 32 * rr;
      if (f < F) {
        const uint4* s4 = reinterpret_cast<const uint4*>(stg + f * 24);
        uint4* d4 = reinterpret_cast<uint4*>(dst_head + (size_t)f * V * Tp + t0);
        d4[0] = s4[0];
        d4[1] = s4[1];
      }
    }
  };

  // ---------------------------------------------------------------- pass 1: dV, dK (warp = 16 keys)
  for (int kt = w; kt < ntile; kt += NW) {
    const int n0 = kt * 16;
    uint32_t ka[4][4], va[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      lda_kmajor(Ks, qstr, n0, ks * 16, lane, ka[ks]);
      lda_rowmajor(Vs, VSTR, n0, ks * 16, lane, va[ks]);
    }
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[nt][i] = dk[nt][i] = 0.f;
    for (int qt = 0; qt < ntile; ++qt) {
      const int q0 = qt * 16;
      float st[2][4], dp[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) st[nt][i] = dp[nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4];
        ldb_kmajor(Qs, qstr, ks * 16, q0, lane, bb);
        mma16816(st[0], ka[ks], bb[0], bb[1]);
        mma16816(st[1], ka[ks], bb[2], bb[3]);
        ldb_nmajor(dOs, VSTR, ks * 16, q0, lane, bb);
        mma16816(dp[0], va[ks], bb[0], bb[1]);
        mma16816(dp[1], va[ks], bb[2], bb[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int qq = q0 + nt * 8 + 2 * tq + (i & 1), key = n0 + g + 8 * (i >> 1);
          const float pv = (qq < T && key < T) ? exp2f(st[nt][i] * sc2 - lse_s[qq]) : 0.f;
          st[nt][i] = pv;
          dp[nt][i] = pv * (dp[nt][i] - del_s[qq]) * sc;
        }
      const uint32_t pa[4] = {pk(st[0][0], st[0][1]), pk(st[0][2], st[0][3]), pk(st[1][0], st[1][1]), pk(st[1][2], st[1][3])};
      const uint32_t da[4] = {pk(dp[0][0], dp[0][1]), pk(dp[0][2], dp[0][3]), pk(dp[1][0], dp[1][1]), pk(dp[1][2], dp[1][3])};
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldb_kmajor(dOs, VSTR, q0, np * 16, lane, bb);
        mma16816(dv[2 * np], pa, bb[0], bb[1]);
        mma16816(dv[2 * np + 1], pa, bb[2], bb[3]);
        ldb_nmajor(Qs, qstr, q0, np * 16, lane, bb);
        mma16816(dk[2 * np], da, bb[0], bb[1]);
        mma16816(dk[2 * np + 1], da, bb[2], bb[3]);
      }
    }
    bf16* dvh = reinterpret_cast<bf16*>(p.dv) + (p.v_btvc ? (size_t)b * T * V * C + (size_t)v * C : (size_t)h * T * C);
    const size_t dvs = p.v_btvc ? (size_t)V * C : (size_t)C;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int key = n0 + g + 8 * half;
      if (key < T) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          *reinterpret_cast<uint32_t*>(dvh + (size_t)key * dvs + nt * 8 + 2 * tq) = pk(dv[nt][2 * half], dv[nt][2 * half + 1]);
      }
    }
    store_fmajor(dk, reinterpret_cast<bf16*>(p.dk) + ((size_t)b * F * V + v) * Tp, n0);
  }
  // ---------------------------------------------------------------- pass 2: dQ (warp = 16 queries)
  for (int qt = w; qt < ntile; qt += NW) {
    const int q0 = qt * 16;
    uint32_t qa[4][4], doa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      lda_kmajor(Qs, qstr, q0, ks * 16, lane, qa[ks]);
      lda_rowmajor(dOs, VSTR, q0, ks * 16, lane, doa[ks]);
    }
    const float l2[2] = {lse_s[q0 + g], lse_s[q0 + g + 8]}, dl[2] = {del_s[q0 + g], del_s[q0 + g + 8]};
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[nt][i] = 0.f;
    for (int kt = 0; kt < ntile; ++kt) {
      const int n0 = kt * 16;
      float s[2][4], dp[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = dp[nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4];
        ldb_kmajor(Ks, qstr, ks * 16, n0, lane, bb);
        mma16816(s[0], qa[ks], bb[0], bb[1]);
        mma16816(s[1], qa[ks], bb[2], bb[3]);
        ldb_nmajor(Vs, VSTR, ks * 16, n0, lane, bb);
        mma16816(dp[0], doa[ks], bb[0], bb[1]);
        mma16816(dp[1], doa[ks], bb[2], bb[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int key = n0 + nt * 8 + 2 * tq + (i & 1), qq = q0 + g + 8 * (i >> 1);
          const float pv = (qq < T && key < T) ? exp2f(s[nt][i] * sc2 - l2[i >> 1]) : 0.f;
          dp[nt][i] = pv * (dp[nt][i] - dl[i >> 1]) * sc;
        }
      const uint32_t da[4] = {pk(dp[0][0], dp[0][1]), pk(dp[0][2], dp[0][3]), pk(dp[1][0], dp[1][1]), pk(dp[1][2], dp[1][3])};
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldb_nmajor(Ks, qstr, n0, np * 16, lane, bb);
        mma16816(dq[2 * np], da, bb[0], bb[1]);
        mma16816(dq[2 * np + 1], da, bb[2], bb[3]);
      }
    }
    store_fmajor(dq, reinterpret_cast<bf16*>(p.dq) + ((size_t)b * F * V + v) * Tp, q0);
  }
}

}  // namespace ta
}  // namespace fmm

extern "C" {

// mode 0: forward (out, lse); mode 1: backward (dq, dk, dv from dout, out, lse). bf16, 64 value channels, F <= 64, Tp <= 320.
int fmm_tattn(const fmm::TAttnArgs* a, int mode, cudaStream_t stream) {
  using namespace fmm;
  FMM_CHECK_ARG(a && a->B >= 1 && a->V >= 1 && a->T >= 1 && a->Tp >= a->T && a->Tp % 64 == 0 && a->Tp <= 320 && a->F >= 1 && a->F <= 64,
                "tattn: unsupported sizes (T=%d Tp=%d F=%d)", a ? a->T : 0, a ? a->Tp : 0, a ? a->F : 0);
  FMM_CHECK_ARG(mode == 0 || mode == 1, "tattn: mode %d", mode);
  const int qstr = a->Tp + 8;
  size_t smem = (size_t)2 * 64 * qstr * 2 + (size_t)a->Tp * ta::VSTR * 2;
  if (mode == 1) smem += (size_t)a->Tp * ta::VSTR * 2 + (size_t)a->Tp * 8 + (size_t)ta::NW * 64 * 24 * 2;
  auto kern = mode == 0 ? ta::tattn_fwd_kernel : ta::tattn_bwd_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("tattn: smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
    return FMM_ERR_SMEM;
  }
  kern<<<a->B * a->V, ta::NT, smem, stream>>>(*a);
  FMM_CHECK_LAUNCH("tattn");
  return FMM_OK;
}

}  // extern "C"
