"""GPU parity of the memory-bound / tiny kernels (csrc/elementwise.cu, csrc/tiny.cu) against torch
autograd on small composites that mirror the reference block (stgcan.py:112-144)."""
import pytest
import torch
import torch.nn.functional as F

gpu = pytest.mark.gpu
EPS = 1e-5


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def cl(x):  # NCHW -> channels-last (N,T,V,C) contiguous
    return x.permute(0, 2, 3, 1).contiguous()


@gpu
@pytest.mark.parametrize("shape", [(8, 3, 14, 256), (4, 12, 14, 64), (3, 7, 33, 128)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_bn1_relu_backward(shape, dtype):
    from fall_multimodal_b200 import ops

    N, T, V, C = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    G = (torch.randn(N, C, T, V, generator=g) * 1.5 + 0.3).to(dev)
    dH = torch.randn(N, C, T, V, generator=g).to(dev)
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.2).to(dev).requires_grad_(True)
    Gc = cl(G).to(dtype)
    dHc = cl(dH).to(dtype)
    Gr = Gc.float().permute(0, 3, 1, 2).requires_grad_(True)  # the values the kernels see
    H = F.relu(F.batch_norm(Gr, None, None, gamma, beta, True, 0.1, EPS))
    H.backward(dHc.float().permute(0, 3, 1, 2))
    # kernels
    st = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    ops.colstats(Gc, st[:C], st[C:])
    a1, b1, mean1, rstd1 = (torch.empty(C, device=dev) for _ in range(4))
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ops.bn_finalize(st[:C], st[C:], N * T * V, gamma.detach(), beta.detach(), rm, rv, True, a1, b1, mean1, rstd1)
    T1 = torch.zeros(C, dtype=torch.float64, device=dev)
    T2 = torch.zeros(C, dtype=torch.float64, device=dev)
    ops.bn1_bwd_reduce(dHc, Gc, a1, b1, T1, T2)
    c1, c2, c3 = (torch.empty(C, device=dev) for _ in range(3))
    dgam, dbet = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    ops.bn1_bwd_coef(T1, T2, a1, mean1, rstd1, N * T * V, True, c1, c2, c3, dgam, dbet)
    dG = torch.empty_like(Gc)
    Tbl = torch.zeros(V, C, device=dev)
    ops.bn1_bwd_apply(dHc, Gc, a1, b1, c1, c2, c3, dG, Tbl)
    torch.cuda.synchronize()
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert rel(dgam, gamma.grad) < tol
    assert rel(dbet, beta.grad) < tol
    assert rel(dG, cl(Gr.grad)) < tol
    assert rel(Tbl, dG.float().sum((0, 1))) < 1e-4
    # running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    bn(Gr.detach())
    assert rel(rm, bn.running_mean) < 1e-4 and rel(rv, bn.running_var) < 1e-4


@gpu
@pytest.mark.parametrize("shape", [(8, 3, 14, 256), (4, 12, 14, 64), (5, 6, 33, 128)])
@pytest.mark.parametrize("reskind", ["none", "identity", "conv"])
def test_block_tail_forward_backward(shape, reskind):
    """z = bn2(U); s = SE(z); Y = relu(s*z + res): forward values and every gradient."""
    from fall_multimodal_b200 import ops

    N, T, V, C = shape
    C4 = C // 4
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
    U = (rnd(N, C, T, V) + rnd(N, C, 1, 1, sc=0.7)).requires_grad_(True)
    X = rnd(N, C, T, V).requires_grad_(True)          # identity residual input
    Rr = (rnd(N, C, T, V) * 0.8 + 0.1).requires_grad_(True)  # residual conv output (pre BN)
    dY = rnd(N, C, T, V)
    p = {k: v.requires_grad_(True) for k, v in dict(
        g2=torch.rand(C, generator=g).to(dev) + 0.5, b2=rnd(C, sc=0.2), W1=rnd(C4, C, sc=C ** -0.5), b1=rnd(C4, sc=0.1),
        gh=torch.rand(C4, generator=g).to(dev) + 0.5, bh=rnd(C4, sc=0.2), W2=rnd(C, C4, sc=C4 ** -0.5), b2se=rnd(C, sc=0.1),
        gr=torch.rand(C, generator=g).to(dev) + 0.5, br=rnd(C, sc=0.2)).items()}
    # ---- torch reference ----
    z = F.batch_norm(U, None, None, p["g2"], p["b2"], True, 0.1, EPS)
    pool = z.mean((2, 3))
    h = F.batch_norm(F.linear(pool, p["W1"], p["b1"]), None, None, p["gh"], p["bh"], True, 0.1, EPS)
    s = torch.sigmoid(F.linear(F.relu(h), p["W2"], p["b2se"]))
    res = 0 if reskind == "none" else X if reskind == "identity" else F.batch_norm(Rr, None, None, p["gr"], p["br"], True, 0.1, EPS)
    Y = F.relu(z * s[:, :, None, None] + res)
    Y.backward(dY)
    # ---- kernels ----
    f32 = lambda *sh: torch.empty(*sh, device=dev)
    z32 = lambda *sh: torch.zeros(*sh, device=dev)
    Uc, Xc, Rc, dYc = cl(U.detach()), cl(X.detach()), cl(Rr.detach()), cl(dY)
    st = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    poolk = z32(N, C)
    ops.colstats(Uc, st[:C], st[C:], poolk)
    a2, b2, mean2, rstd2 = f32(C), f32(C), f32(C), f32(C)
    ops.bn_finalize(st[:C], st[C:], N * T * V, p["g2"].detach(), p["b2"].detach(), z32(C), z32(C) + 1, True, a2, b2, mean2, rstd2)
    pk, hk, sk, k1, k0 = f32(N, C), f32(N, C4), f32(N, C), f32(N, C), f32(N, C)
    ah, bh, hmean, hrstd = f32(C4), f32(C4), f32(C4), f32(C4)
    M = T * V
    ops.se_fwd(poolk, a2, b2, 1.0 / M, p["W1"].detach(), p["b1"].detach(), p["gh"].detach(), p["bh"].detach(), z32(C4),
               z32(C4) + 1, True, p["W2"].detach(), p["b2se"].detach(), pk, hk, ah, bh, hmean, hrstd, sk, k1, k0)
    ar = br = meanr = rstdr = None
    Rk = None
    if reskind == "conv":
        Rk = Rc
        st3 = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        ops.colstats(Rk, st3[:C], st3[C:])
        ar, br, meanr, rstdr = f32(C), f32(C), f32(C), f32(C)
        ops.bn_finalize(st3[:C], st3[C:], N * T * V, p["gr"].detach(), p["br"].detach(), z32(C), z32(C) + 1, True, ar, br, meanr, rstdr)
    resk = None if reskind == "none" else Xc if reskind == "identity" else Rk
    Yk = torch.empty_like(Uc)
    ops.block_out(Uc, k1, k0, resk, ar, br, Yk)
    torch.cuda.synchronize()
    assert rel(sk, s) < 2e-5
    assert rel(Yk, cl(Y)) < 2e-5
    # backward
    S1, S2 = z32(N, C), z32(N, C)
    S3 = z32(N, C) if Rk is not None else None
    ops.blockout_bwd_reduce(dYc, Yk, Uc, Rk, S1, S2, S3)
    dq, dp, dhr, r_, dh = f32(N, C), f32(N, C), f32(N, C4), f32(N, C4), f32(N, C4)
    dW1, db1, dgh, dbh, dW2, db2se = z32(C4, C), z32(C4), z32(C4), z32(C4), z32(C, C4), z32(C)
    ops.se_bwd(S1, S2, a2, b2, sk, pk, hk, ah, bh, hmean, hrstd, p["W1"].detach(), p["W2"].detach(), True, dq, dhr, r_, dh,
               dp, dW1, db1, dgh, dbh, dW2, db2se)
    kk1, kk3, kk2 = f32(N, C), f32(N, C), f32(C)
    dg2, db2 = z32(C), z32(C)
    r1 = r2 = r3 = dgr = dbr = None
    if Rk is not None:
        r1, r2, r3, dgr, dbr = f32(C), f32(C), f32(C), z32(C), z32(C)
    ops.bn2_bwd_coef(S1, S2, S3, poolk, dp, sk, a2, mean2, rstd2, ar, meanr, rstdr, M, N * M, True, kk1, kk2, kk3, r1, r2,
                     r3, dg2, db2, dgr, dbr)
    dU = torch.empty_like(Uc)
    dR = torch.empty_like(Uc) if Rk is not None else None
    dPre = torch.empty_like(Uc) if reskind == "identity" else None
    sum_dU = torch.zeros(C, dtype=torch.float64, device=dev)
    sum_dR = torch.zeros(C, dtype=torch.float64, device=dev) if Rk is not None else None
    ops.bn2_bwd_apply(dYc, Yk, Uc, Rk, kk1, kk2, kk3, r1, r2, r3, dU, dR, dPre, sum_dU, sum_dR)
    torch.cuda.synchronize()
    tol = 3e-5
    assert rel(dU, cl(U.grad)) < tol
    assert rel(dg2, p["g2"].grad) < tol and rel(db2, p["b2"].grad) < tol
    assert rel(dW1, p["W1"].grad) < tol and rel(dW2, p["W2"].grad) < tol
    assert rel(dgh, p["gh"].grad) < tol and rel(dbh, p["bh"].grad) < tol
    assert rel(db2se, p["b2se"].grad) < tol
    assert sum_dU.abs().max().item() < 1e-3 * dU.abs().sum().item()
    if reskind == "identity":
        assert rel(dPre, cl(X.grad)) < tol
    if reskind == "conv":
        assert rel(dR, cl(Rr.grad)) < tol
        assert rel(dgr, p["gr"].grad) < tol and rel(dbr, p["br"].grad) < tol
