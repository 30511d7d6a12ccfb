// Graph-GRU cell glue of the TRAGCN family (reference: GRU.py:17-26 around EmbGCN.py:73-88), fused so
// that one time step is four launches each way: the two per-node GEMMs (bgemm.cu) and, between them,
// ONE kernel per stage that finishes the previous stage's gate maths and prepares the next GEMM's input.
//
// One block per clip; every global access is an 8-channel vector. The concatenated cell input is laid
// out [h (H) | x (Din) | 1 | 0-pad] (state first, so its vectors are aligned whatever Din is; the rows
// of the per-node weights are permuted to match on the host), Cp = padded width, column H+Din carries
// the bias row of the weights.
//
//   forward  mode 0: cat=[h_{t-1}|x_t]                                  -> XCg (mixed S.cat and plain cat)
//            mode 1: zr=sigmoid(pre+silu(lin)) -> ZR,LG ; cat=[r*h_{t-1}|x_t]          -> XCu
//            mode 2: hc=tanh(pre+silu(lin)), h_t=z*h_{t-1}+(1-z)*hc -> HC,LU,H_t ; cat=[h_t|x_{t+1}] -> XCg
//   backward mode 0: update/candidate backward of step t ("bwd1")
//            mode 1: undo candidate-stage mix: dcat=S^T.dXC0+dXC1 ; dx ; dr ; carry+=drh*r ; gate-stage dpre/dlin
//            mode 2: undo gate-stage mix: dx+= ; carry+=dcat_h ; then bwd1 of step t-1
#include "common.cuh"

namespace fmm {

struct CellFwdArgs {
  const void* x;      long long xb, xv;   // x slice (B,V,Din) feeding the cat being built (null: no cat)
  const void* hprev;  long long hb, hv;   // h_{t-1} (null at t = 0)
  const float* S;                         // (V,V) supports
  const void* pre;    const void* lin;    // (B,V,Co) GEMM outputs of the stage being finished, in the activation dtype
  void* zr;  void* lg;                    // (B,V,2H): mode 1 writes, mode 2 reads zr
  void* hc;  void* lu;                    // (B,V,H): mode 2 writes
  void* hout; long long ob, ov;           // H_t slice (mode 2)
  void* xc0; void* xc1;                   // (B,V,Cp) outputs
  int mode, B, V, Din, H, Cp;
};

struct CellBwdArgs {
  const float* S;
  float* carry; float* dz;                // (B,V,H) fp32
  const void* dxc0; const void* dxc1;     // (B,V,Cp) gradients of the stage inputs (modes 1, 2)
  void* dx; long long dxb, dxv;           // dX_t slice (B,V,Din) or null
  const void* hprev; long long hb, hv;    // mode 1: h_{t-1}
  const void* zr; const void* lg;         // mode 1: ZR[t], LG[t]
  void* dpre_g; void* dlin_g;             // mode 1: (B,V,2H)
  // bwd1 operands (mode 0: step t, mode 2: step t-1)
  const void* dH; long long db, dv;       // gradient slice of H_t'
  const void* z1;                         // ZR[t'] (row pitch 2H)
  const void* hprev1; long long hb1, hv1; // h_{t'-1} (null at t' = 0)
  const void* hc1; const void* lu1;       // HC[t'], LU[t']
  void* dpre_u; void* dlin_u;             // (B,V,H)
  int mode, dx_accum, do_bwd1, B, V, Din, H, Cp;
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float dsilu(float x) {
  float s = sigm(x);
  return s * (1.f + x * (1.f - s));
}
template <typename T>
__device__ __forceinline__ float rnd(float v) { return to_f32(from_f32<T>(v)); }

__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) { load8(p, f); }

// Shared-memory rows are stored "split": the 8 channels of group g live as two float4, the low half at
// [g*4] and the high half at [Cp/2 + g*4], so that the float4 reads of consecutive lanes (consecutive
// groups) are 16 bytes apart and bank-conflict free (8 contiguous floats per lane would be a 2-way conflict).
__device__ __forceinline__ int split_idx(int Cp, int c) { return ((c & 4) ? (Cp >> 1) : 0) + ((c >> 3) << 2) + (c & 3); }
__device__ __forceinline__ void sm_st8(float* row, int Cp, int c0, const float (&v)[8]) {
  *reinterpret_cast<float4*>(&row[(c0 >> 3) << 2]) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(&row[(Cp >> 1) + ((c0 >> 3) << 2)]) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void sm_ld8(const float* row, int Cp, int c0, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(&row[(c0 >> 3) << 2]);
  const float4 b = *reinterpret_cast<const float4*>(&row[(Cp >> 1) + ((c0 >> 3) << 2)]);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// acc[0..7] += sum_m coef(m) * rows[m][c0..c0+7]; coef(m) = Ss[n*V+m] (forward, S.cat) or Ss[m*V+n] (backward, S^T.d)
template <bool kTransposed>
__device__ __forceinline__ void mix8(const float* Ss, const float* rows, int V, int Cp, int n, int c0, float (&acc)[8]) {
  const int g4 = (c0 >> 3) << 2, hi = Cp >> 1;
  for (int m = 0; m < V; ++m) {
    const float s = kTransposed ? Ss[m * V + n] : Ss[n * V + m];
    const float4 a = *reinterpret_cast<const float4*>(&rows[m * Cp + g4]);
    const float4 b = *reinterpret_cast<const float4*>(&rows[m * Cp + hi + g4]);
    acc[0] += s * a.x; acc[1] += s * a.y; acc[2] += s * a.z; acc[3] += s * a.w;
    acc[4] += s * b.x; acc[5] += s * b.y; acc[6] += s * b.z; acc[7] += s * b.w;
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) cell_fwd_kernel(const CellFwdArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* cat = sm;                  // [V][Cp]
  float* Ss = sm + p.V * p.Cp;      // [V][V]
  const int b = blockIdx.x, V = p.V, H = p.H, Din = p.Din, Cp = p.Cp, Cin = H + Din;
  const int H8 = H >> 3;
  const T* hprev = reinterpret_cast<const T*>(p.hprev);
  const bool want_cat = p.xc0 != nullptr;
  if (want_cat)
    for (int i = threadIdx.x; i < V * V; i += blockDim.x) Ss[i] = p.S[i];

  if (p.mode == 0) {
    for (int it = threadIdx.x; it < V * H8; it += blockDim.x) {
      const int m = it / H8, j = (it % H8) * 8;
      float h[8];
      if (hprev) load8(hprev + b * p.hb + m * p.hv + j, h);
      else
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = 0.f;
      sm_st8(&cat[m * Cp], Cp, j, h);
    }
  } else if (p.mode == 1) {
    T* zr = reinterpret_cast<T*>(p.zr);
    T* lg = reinterpret_cast<T*>(p.lg);
    const int C8 = 2 * H8;
    for (int it = threadIdx.x; it < V * C8; it += blockDim.x) {
      const int m = it / C8, c = (it % C8) * 8;
      const long long o = ((long long)b * V + m) * 2 * H + c;
      float pr[8], li[8], g[8];
      load8(reinterpret_cast<const T*>(p.pre) + o, pr);
      load8(reinterpret_cast<const T*>(p.lin) + o, li);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = sigm(pr[e] + li[e] * sigm(li[e]));
      store8(zr + o, g);
      store8(lg + o, li);
      if (c >= H && want_cat) {
        const int j = c - H;
        float h[8];
        if (hprev) load8(hprev + b * p.hb + m * p.hv + j, h);
        else
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = rnd<T>(rnd<T>(g[e]) * h[e]);  // r*state is a rounded tensor in the reference
        sm_st8(&cat[m * Cp], Cp, j, h);
      }
    }
  } else {
    const T* zr = reinterpret_cast<const T*>(p.zr);
    T* hc = reinterpret_cast<T*>(p.hc);
    T* lu = reinterpret_cast<T*>(p.lu);
    T* hout = reinterpret_cast<T*>(p.hout);
    for (int it = threadIdx.x; it < V * H8; it += blockDim.x) {
      const int m = it / H8, j = (it % H8) * 8;
      const long long row = (long long)b * V + m;
      float pr[8], li[8], z[8], h[8], c2[8];
      load8(reinterpret_cast<const T*>(p.pre) + row * H + j, pr);
      load8(reinterpret_cast<const T*>(p.lin) + row * H + j, li);
      load8(zr + row * 2 * H + j, z);
      if (hprev) load8(hprev + b * p.hb + m * p.hv + j, h);
      else
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        c2[e] = rnd<T>(tanhf(pr[e] + li[e] * sigm(li[e])));
        h[e] = rnd<T>(z[e] * h[e] + (1.f - z[e]) * c2[e]);
      }
      store8(hc + row * H + j, c2);
      store8(lu + row * H + j, li);
      store8(hout + b * p.ob + m * p.ov + j, h);
      if (want_cat) sm_st8(&cat[m * Cp], Cp, j, h);
    }
  }
  if (!want_cat) return;
  // x part, the constant-1 column and the zero pad
  {
    const T* x = reinterpret_cast<const T*>(p.x);
    const int W = Cp - H;
    const bool vec = (Din & 7) == 0 && ((p.xb | p.xv) & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (vec) {
      const int W8 = W >> 3, D8 = Din >> 3;
      for (int it = threadIdx.x; it < V * W8; it += blockDim.x) {
        const int m = it / W8, c8 = it % W8;
        float v[8];
        if (c8 < D8) load8(x + b * p.xb + m * p.xv + c8 * 8, v);
        else
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (c8 == D8 && e == 0) ? 1.f : 0.f;
        sm_st8(&cat[m * Cp], Cp, H + c8 * 8, v);
      }
    } else {
      for (int i = threadIdx.x; i < V * W; i += blockDim.x) {
        const int m = i / W, c = i % W;
        float v = 0.f;
        if (c < Din) v = to_f32(x[b * p.xb + m * p.xv + c]);
        else if (c == Din) v = 1.f;
        cat[m * Cp + split_idx(Cp, H + c)] = v;
      }
    }
  }
  __syncthreads();
  T* xc0 = reinterpret_cast<T*>(p.xc0);
  T* xc1 = reinterpret_cast<T*>(p.xc1);
  const int Cp8 = Cp >> 3;
  const long long ob = (long long)b * V * Cp;
  for (int it = threadIdx.x; it < V * Cp8; it += blockDim.x) {
    const int n = it / Cp8, c0 = (it % Cp8) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, own[8];
    mix8<false>(Ss, cat, V, Cp, n, c0, acc);
    sm_ld8(&cat[n * Cp], Cp, c0, own);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (c0 + e >= Cin) acc[e] = own[e];  // bias column and pad are not mixed
    store8(xc0 + ob + n * Cp + c0, acc);
    store8(xc1 + ob + n * Cp + c0, own);
  }
}

// bwd1 on one 8-vector: d = carry + dH ; outputs dz, new carry, dpre_u, dlin_u
template <typename T>
__device__ __forceinline__ void bwd1_vec(const CellBwdArgs& p, int b, int m, int j, float (&d)[8]) {
  const int V = p.V, H = p.H;
  const long long row = (long long)b * V + m;
  float g[8], z[8], hp[8], hcv[8], l[8], dzv[8], dp[8], dl[8];
  if (p.dH) {
    load8(reinterpret_cast<const T*>(p.dH) + b * p.db + m * p.dv + j, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] += g[e];
  }
  load8(reinterpret_cast<const T*>(p.z1) + row * 2 * H + j, z);
  if (p.hprev1) load8(reinterpret_cast<const T*>(p.hprev1) + b * p.hb1 + m * p.hv1 + j, hp);
  else
#pragma unroll
    for (int e = 0; e < 8; ++e) hp[e] = 0.f;
  load8(reinterpret_cast<const T*>(p.hc1) + row * H + j, hcv);
  load8(reinterpret_cast<const T*>(p.lu1) + row * H + j, l);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    dzv[e] = d[e] * (hp[e] - hcv[e]);
    dp[e] = d[e] * (1.f - z[e]) * (1.f - hcv[e] * hcv[e]);
    dl[e] = dp[e] * dsilu(l[e]);
    d[e] = d[e] * z[e];
  }
  store8(p.dz + row * H + j, dzv);
  store8(p.carry + row * H + j, d);
  store8(reinterpret_cast<T*>(p.dpre_u) + row * H + j, dp);
  store8(reinterpret_cast<T*>(p.dlin_u) + row * H + j, dl);
}

template <typename T>
__global__ void __launch_bounds__(256, 4) cell_bwd_kernel(const CellBwdArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* d0 = sm;                 // [V][Cp]
  float* Ss = sm + p.V * p.Cp;    // [V][V]
  const int b = blockIdx.x, V = p.V, H = p.H, Din = p.Din, Cp = p.Cp;
  const int H8 = H >> 3, Cp8 = Cp >> 3;
  if (p.mode == 0) {
    for (int it = threadIdx.x; it < V * H8; it += blockDim.x) {
      const int m = it / H8, j = (it % H8) * 8;
      float d[8];
      ld8f(p.carry + ((long long)b * V + m) * H + j, d);
      bwd1_vec<T>(p, b, m, j, d);
    }
    return;
  }
  const long long ob = (long long)b * V * Cp;
  const T* dxc0 = reinterpret_cast<const T*>(p.dxc0);
  const T* dxc1 = reinterpret_cast<const T*>(p.dxc1);
  for (int i = threadIdx.x; i < V * V; i += blockDim.x) Ss[i] = p.S[i];
  for (int it = threadIdx.x; it < V * Cp8; it += blockDim.x) {
    float v[8];
    load8(dxc0 + ob + it * 8, v);
    sm_st8(&d0[(it / Cp8) * Cp], Cp, (it % Cp8) * 8, v);
  }
  __syncthreads();
  for (int it = threadIdx.x; it < V * Cp8; it += blockDim.x) {
    const int m = it / Cp8, c0 = (it % Cp8) * 8;
    if (c0 >= H + Din) continue;  // bias column / pad only
    float acc[8];
    load8(dxc1 + ob + m * Cp + c0, acc);
    mix8<true>(Ss, d0, V, Cp, m, c0, acc);
    if (c0 >= H) {  // x part
      if (p.dx) {
        T* dx = reinterpret_cast<T*>(p.dx) + b * p.dxb + m * p.dxv;
        const int cx = c0 - H;
        if (cx + 8 <= Din && ((p.dxb | p.dxv) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.dx) & 15) == 0) {
          if (p.dx_accum) {
            float old[8];
            load8(dx + cx, old);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += old[e];
          }
          store8(dx + cx, acc);
        } else {
          for (int e = 0; e < 8; ++e) {
            const int c = cx + e;
            if (c < Din) dx[c] = from_f32<T>(p.dx_accum ? to_f32(dx[c]) + acc[e] : acc[e]);
          }
        }
      }
      continue;
    }
    const int j = c0;
    const long long row = (long long)b * V + m;
    float cy[8];
    ld8f(p.carry + row * H + j, cy);
    if (p.mode == 1) {
      float hp[8], zz[8], rr[8], dzv[8], lz[8], lr[8], gz[8], gr[8];
      if (p.hprev) load8(reinterpret_cast<const T*>(p.hprev) + b * p.hb + m * p.hv + j, hp);
      else
#pragma unroll
        for (int e = 0; e < 8; ++e) hp[e] = 0.f;
      const T* zr = reinterpret_cast<const T*>(p.zr) + row * 2 * H;
      const T* lg = reinterpret_cast<const T*>(p.lg) + row * 2 * H;
      load8(zr + j, zz);
      load8(zr + H + j, rr);
      load8(lg + j, lz);
      load8(lg + H + j, lr);
      ld8f(p.dz + row * H + j, dzv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        cy[e] += acc[e] * rr[e];
        const float dr = acc[e] * hp[e];
        gz[e] = dzv[e] * zz[e] * (1.f - zz[e]);
        gr[e] = dr * rr[e] * (1.f - rr[e]);
        lz[e] = gz[e] * dsilu(lz[e]);
        lr[e] = gr[e] * dsilu(lr[e]);
      }
      store8(p.carry + row * H + j, cy);
      T* dpre = reinterpret_cast<T*>(p.dpre_g) + row * 2 * H;
      T* dlin = reinterpret_cast<T*>(p.dlin_g) + row * 2 * H;
      store8(dpre + j, gz);
      store8(dpre + H + j, gr);
      store8(dlin + j, lz);
      store8(dlin + H + j, lr);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) cy[e] += acc[e];
      if (p.do_bwd1) bwd1_vec<T>(p, b, m, j, cy);
      else store8(p.carry + row * H + j, cy);
    }
  }
}

}  // namespace fmm

using namespace fmm;

static int check_cell(int B, int V, int Din, int H, int Cp, int dtype, const char* who) {
  FMM_CHECK_ARG(dtype == FMM_DT_BF16 || dtype == FMM_DT_F32, "%s: bad dtype %d", who, dtype);
  FMM_CHECK_ARG(B > 0 && V > 0 && V <= 32 && Din > 0 && H > 0 && (H % 8) == 0 && (Cp % 8) == 0 && Cp >= Din + H + 1,
                "%s: bad shape (B %d V %d Din %d H %d Cp %d)", who, B, V, Din, H, Cp);
  FMM_CHECK_ARG(sizeof(float) * ((size_t)V * Cp + (size_t)V * V) <= 48 * 1024, "%s: V*Cp too large for shared memory", who);
  return FMM_OK;
}

extern "C" {

int fmm_tg_cell_fwd(const CellFwdArgs* a, int dtype, void* stream) {
  FMM_CHECK_ARG(a != nullptr, "tg_cell_fwd: null descriptor");
  int rc = check_cell(a->B, a->V, a->Din, a->H, a->Cp, dtype, "tg_cell_fwd");
  if (rc != FMM_OK) return rc;
  FMM_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "tg_cell_fwd: bad mode %d", a->mode);
  FMM_CHECK_ARG(a->mode == 0 || (a->pre && a->lin && a->zr), "tg_cell_fwd: missing stage tensors");
  FMM_CHECK_ARG(a->mode != 2 || (a->hc && a->lu && a->hout), "tg_cell_fwd: missing state outputs");
  FMM_CHECK_ARG(!a->xc0 || (a->xc1 && a->x && a->S), "tg_cell_fwd: missing cat operands");
  size_t smem = sizeof(float) * ((size_t)a->V * a->Cp + (size_t)a->V * a->V);
  if (dtype == FMM_DT_BF16) cell_fwd_kernel<__nv_bfloat16><<<a->B, 256, smem, (cudaStream_t)stream>>>(*a);
  else cell_fwd_kernel<float><<<a->B, 256, smem, (cudaStream_t)stream>>>(*a);
  FMM_CHECK_LAUNCH("tg_cell_fwd");
  return FMM_OK;
}

int fmm_tg_cell_bwd(const CellBwdArgs* a, int dtype, void* stream) {
  FMM_CHECK_ARG(a != nullptr, "tg_cell_bwd: null descriptor");
  int rc = check_cell(a->B, a->V, a->Din, a->H, a->Cp, dtype, "tg_cell_bwd");
  if (rc != FMM_OK) return rc;
  FMM_CHECK_ARG(a->mode >= 0 && a->mode <= 2 && a->carry && a->dz, "tg_cell_bwd: bad mode / missing carry");
  FMM_CHECK_ARG(a->mode == 0 || (a->dxc0 && a->dxc1 && a->S), "tg_cell_bwd: missing stage gradients");
  FMM_CHECK_ARG(a->mode != 1 || (a->zr && a->lg && a->dpre_g && a->dlin_g), "tg_cell_bwd: missing gate tensors");
  FMM_CHECK_ARG(!(a->mode == 0 || (a->mode == 2 && a->do_bwd1)) || (a->z1 && a->hc1 && a->lu1 && a->dpre_u && a->dlin_u),
                "tg_cell_bwd: missing update-stage tensors");
  size_t smem = sizeof(float) * ((size_t)a->V * a->Cp + (size_t)a->V * a->V);
  if (dtype == FMM_DT_BF16) cell_bwd_kernel<__nv_bfloat16><<<a->B, 256, smem, (cudaStream_t)stream>>>(*a);
  else cell_bwd_kernel<float><<<a->B, 256, smem, (cudaStream_t)stream>>>(*a);
  FMM_CHECK_LAUNCH("tg_cell_bwd");
  return FMM_OK;
}

}  // extern "C"
