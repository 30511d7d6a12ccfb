"""Python faces of the C-ABI kernels (raw pointers in, status out; see include/fmm_b200.h).

Activations are channels-last ``(N, T, V, C)`` contiguous CUDA tensors, bf16 or fp32.
Every function launches on ``torch.cuda.current_stream()`` and raises on a non-zero status.
"""
from __future__ import annotations

import torch

from . import _lib as L


_err_words: dict[int, torch.Tensor] = {}


def err_word(device: torch.device) -> torch.Tensor:
    """Per-device watchdog word the GEMM kernels write to before trapping on a stuck barrier."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _err_words:
        _err_words[idx] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_words[idx]


# ---------------------------------------------------------------------------------------------
# optional per-launch timing of the GEMM-class kernels (bench.py's roofline leg): CUDA events on
# the launching stream around every tapconv / wgrad call while ``profile`` is a list.
# ---------------------------------------------------------------------------------------------
profile = None


def _timed(kind, flops, nbytes, fn):
    if profile is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    profile.append((kind, flops, nbytes, e0, e1))
    return out


def set_sm_limit(n: int) -> int:
    """Grid budget of the persistent kernels (``fmm_set_sm_limit``): 0 = all SMs; ``sm_count // 2`` while the two trunks of a fusion
    model run on concurrent streams lets their kernels sit side by side. Read at launch time, so a captured graph keeps the value it
    was captured with. Returns the previous value."""
    return int(L.load().fmm_set_sm_limit(int(n)))


class PackedWeight:
    """Weights of one tap-conv, converted to the tensor-core shared-memory images."""

    __slots__ = ("buf", "cout", "cin", "ntaps", "dtype")

    def __init__(self, buf, cout, cin, ntaps, dtype):
        self.buf, self.cout, self.cin, self.ntaps, self.dtype = buf, cout, cin, ntaps, dtype


def tapconv_pack(w: torch.Tensor, cout: int, cin: int, n2: int, k2: int, sn1: int, sn2: int,
                 sk1: int, sk2: int, sm: int, tapmap, act_dtype: torch.dtype,
                 out: torch.Tensor | None = None) -> PackedWeight:
    """Pack fp32 weights ``w`` (any 2-level strided view, see csrc/tapconv.cu) for `tapconv`."""
    L.require_device(w)
    assert w.dtype == torch.float32
    lib = L.load()
    dt = L.dt_of(act_dtype)
    ntaps = len(tapmap)
    nbytes = lib.fmm_tapconv_packed_bytes(cin, cout, ntaps, dt)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    assert out.numel() >= nbytes
    st = lib.fmm_tapconv_pack(L.ptr(w), L.ptr(out), cout, cin, n2, k2, sn1, sn2, sk1, sk2, sm, ntaps,
                              L.int_array(list(tapmap)), dt, L.stream())
    L.check(st, "tapconv_pack")
    return PackedWeight(out, cout, cin, ntaps, act_dtype)


def tapconv(x: torch.Tensor, pw: PackedWeight, out: torch.Tensor, *, shifts, tj: int,
            istride: int = 1, ostride: int = 1, ooff: int = 0, in_scale=None, in_shift=None,
            in_relu: bool = False, bias=None, bias_per_joint: bool = False) -> torch.Tensor:
    """out[n, j*ostride+ooff, v, :] = bias + sum_m f(x[n, j*istride+shifts[m], v, :]) @ W[m]^T."""
    L.require_device(x)
    assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype == pw.dtype
    N, Tin, V, Cin = x.shape
    No, Tout, Vo, Cout = out.shape
    assert (No, Vo) == (N, V) and Cin == pw.cin and Cout == pw.cout and len(shifts) == pw.ntaps
    lib = L.load()

    def run():
        st = lib.fmm_tapconv(L.ptr(x), L.ptr(out), L.ptr(pw.buf), L.ptr(in_scale), L.ptr(in_shift),
                             int(in_relu), L.ptr(bias), int(bias_per_joint), N, V, Tin, Tout, Cin, Cout, tj, istride,
                             ostride, ooff, len(shifts), L.int_array(list(shifts)), L.dt_of(x.dtype),
                             L.ptr(err_word(x.device)), L.stream())
        L.check(st, "tapconv")
        return out

    # multi-tap launches are tensor-pipe bound, the 1x1 channel mixes stream activations (HBM bound)
    return _timed("tapconv_taps" if len(shifts) > 1 else "tapconv_1x1", 2.0 * N * V * tj * Cin * Cout * len(shifts),
                  float(x.numel() + N * tj * V * Cout) * x.element_size(), run)


def wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, *, shifts, istride: int = 1,
          in_scale=None, in_shift=None, in_relu: bool = False, c2: int | None = None,
          s_m: int, s_c1: int = 0, s_c2: int, s_co: int) -> torch.Tensor:
    """dw[m, ci, co] += sum_{n,v,j} f(x[n, j*istride+shifts[m], v, ci]) * dy[n, j, v, co].

    ``dw`` is an fp32 buffer addressed as ``m*s_m + (ci//c2)*s_c1 + (ci%c2)*s_c2 + co*s_co`` and is
    accumulated into with atomics: zero it first.
    """
    L.require_device(x)
    assert x.is_contiguous() and dy.is_contiguous() and x.dtype == dy.dtype and dw.dtype == torch.float32
    N, Tin, V, Cin = x.shape
    Ny, Tj, Vy, Cout = dy.shape
    assert (Ny, Vy) == (N, V)
    lib = L.load()

    def run():
        st = lib.fmm_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), L.ptr(in_scale), L.ptr(in_shift), int(in_relu),
                           N, V, Tin, Tj, Cin, Cout, istride, len(shifts), L.int_array(list(shifts)),
                           c2 if c2 is not None else Cin, s_m, s_c1, s_c2, s_co, L.dt_of(x.dtype),
                           L.ptr(err_word(x.device)), L.stream())
        L.check(st, "wgrad")
        return dw

    return _timed("wgrad_taps" if len(shifts) > 1 else "wgrad_1x1", 2.0 * N * V * Tj * Cin * Cout * len(shifts),
                  float(x.numel() + dy.numel()) * x.element_size(), run)


# ---------------------------------------------------------------------------------------------
# parameter-side algebra of the graph conv (csrc/gcnprep.cu)
# ---------------------------------------------------------------------------------------------
def gcn_prep_fwd(A, imp, bg, dense_idx, bwd_perm, Cout):
    """-> (coef_f [E], coef_b [E], colsum [K,V], bias_eff [V,Cout]) of A*imp; one launch."""
    K, V, _ = A.shape
    E = dense_idx.numel()
    dev = A.device
    coef_f, coef_b = torch.empty(E, device=dev), torch.empty(E, device=dev)
    colsum, bias_eff = torch.empty(K, V, device=dev), torch.empty(V, Cout, device=dev)
    L.check(L.load().fmm_gcn_prep_fwd(L.ptr(A), L.ptr(imp), L.ptr(bg), L.ptr(dense_idx), L.ptr(bwd_perm), L.ptr(coef_f), L.ptr(coef_b),
                                      L.ptr(colsum), L.ptr(bias_eff), K, V, Cout, E, L.stream()), "gcn_prep_fwd")
    return coef_f, coef_b, colsum, bias_eff


def gcn_prep_bwd(A, bg, colsum, TblR, dcoef, dense_idx, dbg, dimp):
    K, V, _ = A.shape
    Cout = bg.numel() // K
    L.check(L.load().fmm_gcn_prep_bwd(L.ptr(A), L.ptr(bg), L.ptr(colsum), L.ptr(TblR), TblR.numel() // (V * Cout), L.ptr(dcoef),
                                      L.ptr(dense_idx), L.ptr(dbg), L.ptr(dimp), K, V, Cout, dense_idx.numel(), L.stream()), "gcn_prep_bwd")


# ---------------------------------------------------------------------------------------------
# data_bn (csrc/databn.cu): clip (N,C,T,V) fp32 -> normalised channels-last activation (N,T,V,C)
# ---------------------------------------------------------------------------------------------
def databn_stats(x, s, q):
    N, C, Tn, V = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and s.dtype == q.dtype == torch.float64
    L.check(L.load().fmm_databn_stats(L.ptr(x), L.ptr(s), L.ptr(q), N, C, Tn, V, L.stream()), "databn_stats")


def databn_apply(x, a, b, y):
    N, C, Tn, V = x.shape
    assert y.shape == (N, Tn, V, C) and y.is_contiguous() and x.is_contiguous()
    L.check(L.load().fmm_databn_apply(L.ptr(x), L.ptr(a), L.ptr(b), L.ptr(y), N, C, Tn, V, L.dt_of(y.dtype), L.stream()), "databn_apply")
    return y


def databn_bwd(dy, x, mean, rstd, dgamma, dbeta):
    N, C, Tn, V = x.shape
    assert dy.shape == (N, Tn, V, C) and dy.is_contiguous() and dgamma.dtype == dbeta.dtype == torch.float64
    L.check(L.load().fmm_databn_bwd(L.ptr(dy), L.ptr(x), L.ptr(mean), L.ptr(rstd), L.ptr(dgamma), L.ptr(dbeta), N, C, Tn, V,
                                    L.dt_of(dy.dtype), L.stream()), "databn_bwd")


# ---------------------------------------------------------------------------------------------
# fused spatial graph convolution (csrc/gcn.cu): aggregation in the GEMM prologue, BN statistics in the epilogue
# ---------------------------------------------------------------------------------------------
def gcn_supported(x_dtype, cin: int, cout: int, V: int) -> bool:
    return x_dtype == torch.bfloat16 and cin % 64 == 0 and cout % 64 == 0 and cout <= 256 and cin <= 512 and V <= 33


def gcn_pack(w: torch.Tensor, K: int, cin: int, cout: int) -> torch.Tensor:
    """W fp32 (K*Cout, Cin[,1,1]) -> tensor-core operand images for `gcn_fwd`."""
    L.require_device(w)
    assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == K * cout * cin
    lib = L.load()
    out = torch.empty(lib.fmm_gcn_packed_bytes(K, cin, cout), dtype=torch.uint8, device=w.device)
    L.check(lib.fmm_gcn_pack(L.ptr(w), L.ptr(out), K, cin, cout, L.stream()), "gcn_pack")
    return out


def partition_degrees(rowptr, K: int, V: int):
    """Maximum in-degree of each adjacency partition from the CSR row pointer (host list; the graph is static)."""
    rp = rowptr.detach().cpu().view(-1).tolist() if torch.is_tensor(rowptr) else list(rowptr)
    return [max(1, max(rp[k * V + w + 1] - rp[k * V + w] for w in range(V))) for k in range(K)]


def gcn_fwd(x, wpk, g, rowptr, src, coef, K, kdeg, bias=None, ch_sum=None, ch_sq=None, xa=None):
    """g = bias[v] + aggregate(x, edges) @ W^T over the flat rows of channels-last bf16 ``x`` (N,T,V,Cin) -> ``g`` (N,T,V,Cout)."""
    L.require_device(x)
    N, Tn, V, Cin = x.shape
    Cout = g.shape[-1]
    assert x.is_contiguous() and g.is_contiguous() and x.dtype == g.dtype == torch.bfloat16 and g.shape[:3] == x.shape[:3]
    assert xa is None or (xa.is_contiguous() and xa.shape == (N, Tn, V, K * Cin) and xa.dtype == x.dtype)
    rows = N * Tn * V

    def run():
        L.check(L.load().fmm_gcn_fwd(L.ptr(x), L.ptr(g), L.ptr(xa), L.ptr(wpk), L.ptr(bias), L.ptr(rowptr), L.ptr(src),
                                     L.ptr(coef), L.int_array(list(kdeg)), L.ptr(ch_sum), L.ptr(ch_sq), _nrep(ch_sum, Cout), rows, V, K, Cin, Cout,
                                     src.numel(), L.ptr(err_word(x.device)), L.stream()), "gcn_fwd")
        return g

    return _timed("gcn_fwd", 2.0 * rows * K * Cin * Cout, float(x.numel() + g.numel()) * 2, run)


def gcn_wgrad(x, dg, dw, rowptr, src, coef, K, kdeg):
    """dw (K*Cout, Cin) fp32 += aggregate(x)^T dg  (zero dw first); the aggregated operand is re-derived, not read."""
    L.require_device(x)
    N, Tn, V, Cin = x.shape
    Cout = dg.shape[-1]
    assert x.is_contiguous() and dg.is_contiguous() and x.dtype == dg.dtype == torch.bfloat16 and dg.shape[:3] == x.shape[:3]
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == K * Cout * Cin
    rows = N * Tn * V

    def run():
        L.check(L.load().fmm_gcn_wgrad(L.ptr(x), L.ptr(dg), L.ptr(dw), L.ptr(rowptr), L.ptr(src), L.ptr(coef),
                                       L.int_array(list(kdeg)), rows, V, K, Cin, Cout, src.numel(), L.ptr(err_word(x.device)),
                                       L.stream()), "gcn_wgrad")
        return dw

    return _timed("gcn_wgrad", 2.0 * rows * K * Cin * Cout, float(x.numel() + dg.numel()) * 2, run)


def gcn_pack_bwd(w: torch.Tensor, K: int, cin: int, cout: int) -> torch.Tensor:
    """W fp32 (K*Cout, Cin) -> operand images of `gcn_bwd` (P = dG . W_k^T for all K partitions of a 64-channel slab at once)."""
    L.require_device(w)
    assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == K * cout * cin
    lib = L.load()
    out = torch.empty(lib.fmm_gcn_packed_bwd_bytes(K, cin, cout), dtype=torch.uint8, device=w.device)
    L.check(lib.fmm_gcn_pack_bwd(L.ptr(w), L.ptr(out), K, cin, cout, L.stream()), "gcn_pack_bwd")
    return out


def gcn_bwd_supported(dtype, cin: int, cout: int, V: int, K: int, max_out_degree: int) -> bool:
    return gcn_supported(dtype, cin, cout, V) and K <= 3 and max_out_degree <= 8


def gcn_bwd(dg, wpk, dx, rowptr, dst, kk, coef, K, max_out_degree, addend=None, x=None, eid=None, dcoef=None, relu_mask=False):
    """dx = addend + A^T-aggregate(dG . W^T) with the GEMM result kept on chip; with x/eid/dcoef also the edge-coefficient gradient.
    relu_mask: store dx * (x > 0) (x = the previous block's ReLU output), see include/fmm_b200.h."""
    L.require_device(dg)
    N, Tn, V, Cout = dg.shape
    Cin = dx.shape[-1]
    assert dg.is_contiguous() and dx.is_contiguous() and dg.dtype == dx.dtype == torch.bfloat16 and dx.shape[:3] == dg.shape[:3]
    assert addend is None or (addend.shape == dx.shape and addend.is_contiguous() and addend.dtype == dx.dtype)
    assert x is None or (x.shape == dx.shape and x.is_contiguous() and x.dtype == dx.dtype)
    rows = N * Tn * V

    def run():
        L.check(L.load().fmm_gcn_bwd(L.ptr(dg), L.ptr(x), L.ptr(addend), L.ptr(dx), L.ptr(wpk), L.ptr(rowptr), L.ptr(dst), L.ptr(kk),
                                     L.ptr(coef), L.ptr(eid), L.ptr(dcoef), int(bool(relu_mask)), int(max_out_degree), rows, V, K, Cin, Cout,
                                     L.ptr(err_word(dg.device)), L.stream()), "gcn_bwd")
        return dx

    nb = dg.numel() + dx.numel() + (x.numel() if x is not None else 0) + (addend.numel() if addend is not None else 0)
    return _timed("gcn_bwd", 2.0 * rows * K * Cin * Cout, float(nb) * 2, run)


# ---------------------------------------------------------------------------------------------
# memory-bound kernels (csrc/elementwise.cu) and per-channel/per-clip kernels (csrc/tiny.cu)
# ---------------------------------------------------------------------------------------------
def _shape4(x):
    N, Tn, V, C = x.shape
    return N, Tn, V, C


def agg_fwd(x, xa, rowptr, src, coef, K):
    N, Tn, V, Cin = _shape4(x)
    assert xa.shape == (N, Tn, V, K * Cin) and x.is_contiguous() and xa.is_contiguous()
    L.check(L.load().fmm_agg_fwd(L.ptr(x), L.ptr(xa), L.ptr(rowptr), L.ptr(src), L.ptr(coef), N, Tn, V, Cin, K,
                                 L.dt_of(x.dtype), L.stream()), "agg_fwd")
    return xa


def agg_bwd(P, addend, dx, rowptr, dst, kk, coef, K, x=None, eid=None, dcoef=None):
    """dx = addend + A^T-aggregate(P); with ``x`` also dcoef[eid[e]] += <x[src e], P[dst e, k e]> (fused)."""
    N, Tn, V, Cin = _shape4(dx)
    assert P.shape == (N, Tn, V, K * Cin) and P.is_contiguous() and dx.is_contiguous()
    assert addend is None or (addend.shape == dx.shape and addend.is_contiguous())
    assert x is None or (x.shape == dx.shape and x.is_contiguous())
    L.check(L.load().fmm_agg_bwd(L.ptr(P), L.ptr(addend), L.ptr(dx), L.ptr(rowptr), L.ptr(dst), L.ptr(kk),
                                 L.ptr(coef), L.ptr(x), L.ptr(eid), L.ptr(dcoef), N, Tn, V, Cin, K,
                                 L.dt_of(dx.dtype), L.stream()), "agg_bwd")
    return dx


def agg_dcoef(x, P, dcoef, src, dst, kk, K):
    N, Tn, V, Cin = _shape4(x)
    assert P.shape == (N, Tn, V, K * Cin)
    L.check(L.load().fmm_agg_dcoef(L.ptr(x), L.ptr(P), L.ptr(dcoef), L.ptr(src), L.ptr(dst), L.ptr(kk),
                                   dcoef.numel(), N, Tn, V, Cin, K, L.dt_of(x.dtype), L.stream()), "agg_dcoef")
    return dcoef


NREP = 16  # replication of the cross-block accumulators (see csrc/elementwise.cu: replica_of_block)


def _nrep(buf, per_replica):
    if buf is None:
        return 1
    assert buf.numel() % per_replica == 0
    return buf.numel() // per_replica


def colstats(x, ch_sum=None, ch_sq=None, nc_sum=None):
    """ch_sum / ch_sq: fp64 [nrep][C] (nrep inferred from the buffer size); nc_sum: fp32 [N][C]."""
    N, Tn, V, C = _shape4(x)
    assert x.is_contiguous()
    L.check(L.load().fmm_colstats(L.ptr(x), L.ptr(ch_sum), L.ptr(ch_sq), L.ptr(nc_sum), _nrep(ch_sum, C), N, Tn, V, C,
                                  L.dt_of(x.dtype), L.stream()), "colstats")


def affine_relu(X, a, b, H):
    N, Tn, V, C = _shape4(X)
    assert X.is_contiguous() and H.shape == X.shape
    L.check(L.load().fmm_affine_relu(L.ptr(X), L.ptr(a), L.ptr(b), L.ptr(H), N, Tn, V, C, L.dt_of(X.dtype), L.stream()),
            "affine_relu")
    return H


def block_out(U, k1, k0, res, ar, br, Y):
    N, Tn, V, C = _shape4(U)
    L.check(L.load().fmm_block_out(L.ptr(U), L.ptr(k1), L.ptr(k0), L.ptr(res), L.ptr(ar), L.ptr(br), L.ptr(Y),
                                   N, Tn, V, C, L.dt_of(U.dtype), L.stream()), "block_out")
    return Y


def blockout_bwd_reduce(dY, Y, U, R, S1, S2, S3):
    N, Tn, V, C = _shape4(U)
    assert dY.is_contiguous() and dY.shape == U.shape
    L.check(L.load().fmm_blockout_bwd_reduce(L.ptr(dY), L.ptr(Y), L.ptr(U), L.ptr(R), L.ptr(S1), L.ptr(S2),
                                             L.ptr(S3), N, Tn, V, C, L.dt_of(U.dtype), L.stream()),
            "blockout_bwd_reduce")


def bn2_bwd_apply(dY, Y, U, R, k1, k2, k3, r1, r2, r3, dU, dR, dPre, sum_dU, sum_dR):
    N, Tn, V, C = _shape4(U)
    L.check(L.load().fmm_bn2_bwd_apply(L.ptr(dY), L.ptr(Y), L.ptr(U), L.ptr(R), L.ptr(k1), L.ptr(k2), L.ptr(k3),
                                       L.ptr(r1), L.ptr(r2), L.ptr(r3), L.ptr(dU), L.ptr(dR), L.ptr(dPre),
                                       L.ptr(sum_dU), L.ptr(sum_dR), _nrep(sum_dU, C), N, Tn, V, C, L.dt_of(U.dtype),
                                       L.stream()),
            "bn2_bwd_apply")


def bn1_bwd_reduce(dH, G, a1, b1, T1, T2):
    N, Tn, V, C = _shape4(G)
    assert dH.is_contiguous() and dH.shape == G.shape
    L.check(L.load().fmm_bn1_bwd_reduce(L.ptr(dH), L.ptr(G), L.ptr(a1), L.ptr(b1), L.ptr(T1), L.ptr(T2), _nrep(T1, C),
                                        N, Tn, V, C, L.dt_of(G.dtype), L.stream()), "bn1_bwd_reduce")


def bn1_bwd_apply(dH, G, a1, b1, c1, c2, c3, dG, Tbl):
    N, Tn, V, C = _shape4(G)
    L.check(L.load().fmm_bn1_bwd_apply(L.ptr(dH), L.ptr(G), L.ptr(a1), L.ptr(b1), L.ptr(c1), L.ptr(c2), L.ptr(c3),
                                       L.ptr(dG), L.ptr(Tbl), _nrep(Tbl, V * C), N, Tn, V, C, L.dt_of(G.dtype), L.stream()),
            "bn1_bwd_apply")


def bn_finalize(ch_sum, ch_sq, count, gamma, beta, rmean, rvar, training, a, b, mean_out, rstd_out,
                momentum=0.1, eps=1e-5):
    C = a.numel()
    L.check(L.load().fmm_bn_finalize(L.ptr(ch_sum), L.ptr(ch_sq), _nrep(ch_sum, C), float(count), L.ptr(gamma), L.ptr(beta),
                                     L.ptr(rmean), L.ptr(rvar), momentum, eps, int(training), L.ptr(a), L.ptr(b),
                                     L.ptr(mean_out), L.ptr(rstd_out), C, L.stream()), "bn_finalize")


def se_fwd(pool, a2, b2, invM, W1, b1, gamma, beta, rmean, rvar, training, W2, b2se, p, h, ah, bh, hmean, hrstd,
           s, k1, k0, momentum=0.1, eps=1e-5):
    N, C = pool.shape
    C4 = h.shape[1]
    L.check(L.load().fmm_se_fwd(L.ptr(pool), L.ptr(a2), L.ptr(b2), invM, L.ptr(W1), L.ptr(b1), L.ptr(gamma),
                                L.ptr(beta), L.ptr(rmean), L.ptr(rvar), momentum, eps, int(training), L.ptr(W2),
                                L.ptr(b2se), L.ptr(p), L.ptr(h), L.ptr(ah), L.ptr(bh), L.ptr(hmean), L.ptr(hrstd),
                                L.ptr(s), L.ptr(k1), L.ptr(k0), N, C, C4, L.stream()), "se_fwd")


def se_bwd(S1, S2, a2, b2, s, p, h, ah, bh, hmean, hrstd, W1, W2, training, dq, dhr, r, dh, dp, dW1, db1, dgamma,
           dbeta, dW2, db2se):
    N, C = S1.shape
    C4 = h.shape[1]
    L.check(L.load().fmm_se_bwd(L.ptr(S1), L.ptr(S2), L.ptr(a2), L.ptr(b2), L.ptr(s), L.ptr(p), L.ptr(h), L.ptr(ah),
                                L.ptr(bh), L.ptr(hmean), L.ptr(hrstd), L.ptr(W1), L.ptr(W2), int(training),
                                L.ptr(dq), L.ptr(dhr), L.ptr(r), L.ptr(dh), L.ptr(dp), L.ptr(dW1), L.ptr(db1),
                                L.ptr(dgamma), L.ptr(dbeta), L.ptr(dW2), L.ptr(db2se), N, C, C4, L.stream()),
            "se_bwd")


def se_bwd_params(dq, r, dh, p, dW1, db1, dW2, db2se):
    """The weight-gradient products of the SE block (after `se_bwd(..., dW1=None, db1=None, dW2=None, db2se=None)`)."""
    N, C = dq.shape
    C4 = dh.shape[1]
    L.check(L.load().fmm_se_bwd_params(L.ptr(dq), L.ptr(r), L.ptr(dh), L.ptr(p), L.ptr(dW1), L.ptr(db1), L.ptr(dW2),
                                       L.ptr(db2se), N, C, C4, L.stream()), "se_bwd_params")


def bn2_bwd_coef(S1, S2, S3, pool, dp, s, a2, mean2, rstd2, ar, meanr, rstdr, M, count, training, k1, k2, k3, r1,
                 r2, r3, dgamma2, dbeta2, dgammar, dbetar):
    N, C = S1.shape
    L.check(L.load().fmm_bn2_bwd_coef(L.ptr(S1), L.ptr(S2), L.ptr(S3), L.ptr(pool), L.ptr(dp), L.ptr(s), L.ptr(a2),
                                      L.ptr(mean2), L.ptr(rstd2), L.ptr(ar), L.ptr(meanr), L.ptr(rstdr), float(M),
                                      float(count), int(training), L.ptr(k1), L.ptr(k2), L.ptr(k3), L.ptr(r1),
                                      L.ptr(r2), L.ptr(r3), L.ptr(dgamma2), L.ptr(dbeta2), L.ptr(dgammar),
                                      L.ptr(dbetar), N, C, L.stream()), "bn2_bwd_coef")


def bn1_bwd_coef(T1, T2, a1, mean1, rstd1, count, training, c1, c2, c3, dgamma, dbeta):
    C = a1.numel()
    L.check(L.load().fmm_bn1_bwd_coef(L.ptr(T1), L.ptr(T2), _nrep(T1, C), L.ptr(a1), L.ptr(mean1), L.ptr(rstd1), float(count),
                                      int(training), L.ptr(c1), L.ptr(c2), L.ptr(c3), L.ptr(dgamma), L.ptr(dbeta),
                                      C, L.stream()), "bn1_bwd_coef")


# ---------------------------------------------------------------------------------------------
# sensor branch (csrc/sensor.cu): channels-last fp32 windows x[N][L][C]
# ---------------------------------------------------------------------------------------------
def conv1d_k5_fwd(x, w, b, y):
    N, Ln, Ci = x.shape
    Co = w.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and y.shape == (N, Ln, Co) and w.is_contiguous()
    L.check(L.load().fmm_conv1d_k5_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), N, Ln, Ci, Co, L.stream()), "conv1d_k5_fwd")
    return y


def bn_relu_pool2_fwd(y, a, b, out):
    N, Ln, C = y.shape
    assert out.shape == (N, Ln // 2, C)
    L.check(L.load().fmm_bn_relu_pool2_fwd(L.ptr(y), L.ptr(a), L.ptr(b), L.ptr(out), N, Ln, C, L.stream()), "bn_relu_pool2_fwd")
    return out


def pool2_bwd(y, a, b, dout, dh):
    N, Ln, C = y.shape
    assert dout.is_contiguous() and dout.shape == (N, Ln // 2, C) and dh.shape == y.shape
    L.check(L.load().fmm_pool2_bwd(L.ptr(y), L.ptr(a), L.ptr(b), L.ptr(dout), L.ptr(dh), N, Ln, C, L.stream()), "pool2_bwd")
    return dh


def conv1d_k5_bwd(x, dy, w, dx, dw, db):
    N, Ln, Ci = x.shape
    Co = w.shape[0]
    assert dy.is_contiguous() and dy.shape == (N, Ln, Co)
    L.check(L.load().fmm_conv1d_k5_bwd(L.ptr(x), L.ptr(dy), L.ptr(w), L.ptr(dx), L.ptr(dw), L.ptr(db), N, Ln, Ci, Co,
                                       L.stream()), "conv1d_k5_bwd")


# ---------------------------------------------------------------------------------------------
# LSTM recurrence (csrc/lstm.cu), fp32, weights stacked per direction
# ---------------------------------------------------------------------------------------------
def lstm_fwd(x, w_ih, w_hh, b_ih, b_hh, out, gates=None, cseq=None):
    N, Tn, I = x.shape
    ndir, G, H = w_hh.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.shape == (N, Tn, ndir * H)
    L.check(L.load().fmm_lstm_fwd(L.ptr(x), L.ptr(w_ih), L.ptr(w_hh), L.ptr(b_ih), L.ptr(b_hh), L.ptr(out), L.ptr(gates),
                                  L.ptr(cseq), N, Tn, I, H, ndir, L.stream()), "lstm_fwd")
    return out


def lstm_infer(x, w_ih, w_hh, b_ih, b_hh, feat, mean_feature: bool):
    """feat (N, ndir*H) = mean over T / last step of the bidirectional LSTM from zero state (tensor-core inference path)."""
    N, Tn, I = x.shape
    ndir, G, H = w_hh.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and feat.shape == (N, ndir * H) and feat.is_contiguous()
    L.check(L.load().fmm_lstm_infer(L.ptr(x), L.ptr(w_ih), L.ptr(w_hh), L.ptr(b_ih), L.ptr(b_hh), L.ptr(feat), N, Tn, I, H, ndir,
                                    int(mean_feature), L.stream()), "lstm_infer")
    return feat


def lstm_bwd(x, w_ih, w_hh, out, gates, cseq, dout, dw_ih, dw_hh, db, dx=None):
    N, Tn, I = x.shape
    ndir, G, H = w_hh.shape
    assert dout.is_contiguous() and dout.shape == out.shape
    L.check(L.load().fmm_lstm_bwd(L.ptr(x), L.ptr(w_ih), L.ptr(w_hh), L.ptr(out), L.ptr(gates), L.ptr(cseq), L.ptr(dout),
                                  L.ptr(dw_ih), L.ptr(dw_hh), L.ptr(db), L.ptr(dx), N, Tn, I, H, ndir, L.stream()), "lstm_bwd")
