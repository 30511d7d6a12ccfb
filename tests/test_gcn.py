"""GPU: the fused spatial graph convolution (csrc/gcn.cu: adjacency aggregation in the tcgen05 GEMM prologue, BatchNorm
statistics and TMA store in the epilogue) against plain torch fp32 arithmetic of the reference expression
``einsum('nkctv,kvw->nctw', conv1x1(x), A*importance)`` (stgcan.py:50-56,222) on the same bf16 inputs."""
import pytest
import torch

gpu = pytest.mark.gpu

CASES = [  # layout, N, T, Cin, Cout
    ("mediapipe33", 4, 9, 64, 64),
    ("coco_cut", 3, 7, 64, 128),
    ("ntu-rgb+d", 2, 5, 128, 256),
    ("mediapipe33", 2, 6, 256, 256),
    ("coco_mmpose", 5, 3, 128, 128),
    ("mediapipe33", 16, 16, 128, 256),
]


def _setup(layout, N, T, Cin, Cout, dev, seed=0, strategy="spatial"):
    from fall_multimodal_b200.graph import Graph, adjacency_csr

    g = torch.Generator().manual_seed(seed)
    try:
        A = torch.tensor(Graph(layout, strategy).A, dtype=torch.float32)
    except Exception:
        pytest.skip(f"layout {layout} not registered")
    K, V, _ = A.shape
    Ahat = A * (1.0 + 0.2 * torch.randn(K, V, V, generator=g))
    csr = adjacency_csr(A.double().numpy())
    t = lambda a, dt=torch.int32: torch.as_tensor(a).to(device=dev, dtype=dt)
    rowptr, src = t(csr["fwd_rowptr"]), t(csr["fwd_src"])
    coef = Ahat.flatten()[torch.as_tensor(csr["dense_idx"]).long()].contiguous().to(dev)
    x = torch.randn(N, T, V, Cin, generator=g).to(dev, torch.bfloat16)
    W = (torch.randn(K * Cout, Cin, generator=g) / Cin ** 0.5).to(dev)
    bias = (0.1 * torch.randn(V, Cout, generator=g)).to(dev)
    kdeg = [int(v) for v in (A != 0).sum(1).max(1).values.clamp_min(1)]
    return A, Ahat.to(dev), K, V, rowptr, src, coef, x, W, bias, kdeg


@gpu
@pytest.mark.parametrize("layout,N,T,Cin,Cout", CASES)
def test_gcn_fwd_matches_torch(layout, N, T, Cin, Cout):
    from fall_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    A, Ahat, K, V, rowptr, src, coef, x, W, bias, kdeg = _setup(layout, N, T, Cin, Cout, dev)
    wpk = ops.gcn_pack(W, K, Cin, Cout)
    G = torch.full((N, T, V, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    Xa = torch.full((N, T, V, K * Cin), float("nan"), dtype=torch.bfloat16, device=dev)
    s1 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev)
    s2 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev)
    ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias, ch_sum=s1, ch_sq=s2, xa=Xa)
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    # reference: aggregate in fp32, round to bf16 (what the tensor core sees), mix channels with bf16 weights, fp32 accumulate
    xa_ref = torch.einsum("ntvc,kvw->ntwkc", x.float(), Ahat).reshape(N, T, V, K * Cin)
    assert torch.isfinite(Xa.float()).all() and torch.isfinite(G.float()).all()
    e_xa = (Xa.float() - xa_ref).abs().max().item() / xa_ref.abs().max().item()
    assert e_xa < 6e-3, f"aggregated operand err {e_xa:.2e}"
    Wk = W.view(K, Cout, Cin).to(torch.bfloat16).float()
    g_ref = torch.einsum("ntwkc,koc->ntwo", Xa.float().view(N, T, V, K, Cin), Wk) + bias[None, None]
    e_g = (G.float() - g_ref).abs().max().item() / g_ref.abs().max().item()
    assert e_g < 6e-3, f"output err {e_g:.2e}"
    # and against exact arithmetic on the same inputs (bf16 rounding of the operands is the only difference)
    g_exact = torch.einsum("ntwkc,koc->ntwo", xa_ref.double().view(N, T, V, K, Cin), W.view(K, Cout, Cin).double()) + bias.double()
    assert (G.double() - g_exact).abs().max().item() / g_exact.abs().max().item() < 2e-2
    # statistics of what was stored
    Gf = G.double().reshape(-1, Cout)
    sum_ref, sq_ref = Gf.sum(0), (Gf * Gf).sum(0)
    got1, got2 = s1.view(ops.NREP, Cout).sum(0), s2.view(ops.NREP, Cout).sum(0)
    assert (got1 - sum_ref).abs().max().item() <= 1e-4 * max(1.0, Gf.abs().sum(0).max().item())
    assert (got2 - sq_ref).abs().max().item() <= 1e-4 * sq_ref.max().item()


@gpu
@pytest.mark.parametrize("layout,N,T,Cin,Cout", [c for c in CASES if c[4] <= 128] + [("mediapipe33", 37, 20, 128, 128), ("mediapipe33", 64, 16, 64, 64)])
def test_gcn_fwd_tensor_core_aggregation(layout, N, T, Cin, Cout, monkeypatch):
    """The opt-in forward with the adjacency product on the tensor core (FMM_GCN_TC=1: per-frame A_hat . x_f as tcgen05 MMAs into
    TMEM, converted to the operand images of the channel GEMM) against the same references and against the default kernel.
    The edge coefficients are bf16 on this path: one extra rounding of ~2^-9 relative per coefficient."""
    from fall_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    A, Ahat, K, V, rowptr, src, coef, x, W, bias, kdeg = _setup(layout, N, T, Cin, Cout, dev, seed=21)
    if K * V > 128:
        pytest.skip("K*V lanes exceed one MMA tile")
    wpk = ops.gcn_pack(W, K, Cin, Cout)
    outs, stats = [], []
    for tc in ("0", "1"):
        monkeypatch.setenv("FMM_GCN_TC", tc)
        G = torch.full((N, T, V, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
        s1 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev)
        s2 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev)
        ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias, ch_sum=s1, ch_sq=s2)
        torch.cuda.synchronize()
        assert int(ops.err_word(dev).item()) == 0
        assert torch.isfinite(G.float()).all()
        outs.append(G)
        stats.append((s1.view(ops.NREP, Cout).sum(0), s2.view(ops.NREP, Cout).sum(0)))
    g_exact = torch.einsum("ntvc,kvw,koc->ntwo", x.double(), Ahat.double(), W.view(K, Cout, Cin).double()) + bias.double()
    scale = g_exact.abs().max().item()
    e0 = (outs[0].double() - g_exact).abs().max().item() / scale
    e1 = (outs[1].double() - g_exact).abs().max().item() / scale
    assert e1 < 2e-2 and e1 < 2.0 * e0 + 4e-3, f"tensor-core aggregation err {e1:.2e} (default kernel {e0:.2e})"
    assert (outs[1].float() - outs[0].float()).abs().max().item() / scale < 1.2e-2
    # statistics of what THIS path stored
    Gf = outs[1].double().reshape(-1, Cout)
    assert (stats[1][0] - Gf.sum(0)).abs().max().item() <= 1e-4 * max(1.0, Gf.abs().sum(0).max().item())
    assert (stats[1][1] - (Gf * Gf).sum(0)).abs().max().item() <= 1e-4 * (Gf * Gf).sum(0).max().item()


@gpu
def test_gcn_fwd_without_optional_outputs_and_bench_shape():
    """No bias / statistics / xa; a bench-sized launch (N=256, T=64, V=33, 64->64) checked on a row sample."""
    from fall_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    N, T, Cin, Cout = 256, 64, 64, 64
    A, Ahat, K, V, rowptr, src, coef, x, W, bias, kdeg = _setup("mediapipe33", N, T, Cin, Cout, dev, seed=3)
    wpk = ops.gcn_pack(W, K, Cin, Cout)
    G = torch.empty(N, T, V, Cout, dtype=torch.bfloat16, device=dev)
    ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg)
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    for n in (0, 77, 255):
        xa_ref = torch.einsum("tvc,kvw->twkc", x[n].float(), Ahat).to(torch.bfloat16).float()
        g_ref = torch.einsum("twkc,koc->two", xa_ref, W.view(K, Cout, Cin).to(torch.bfloat16).float())
        err = (G[n].float() - g_ref).abs().max().item() / g_ref.abs().max().item()
        assert err < 6e-3, (n, err)


@gpu
@pytest.mark.parametrize("layout,N,T,Cin,Cout", CASES)
def test_gcn_wgrad_matches_torch(layout, N, T, Cin, Cout):
    """dW = aggregate(x)^T dG with the aggregated operand re-derived in the prologue, against fp64 torch on the bf16 inputs."""
    from fall_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    A, Ahat, K, V, rowptr, src, coef, x, W, bias, kdeg = _setup(layout, N, T, Cin, Cout, dev, seed=5)
    g = torch.Generator().manual_seed(9)
    dG = torch.randn(N, T, V, Cout, generator=g).to(dev, torch.bfloat16)
    dW = torch.zeros(K * Cout, Cin, device=dev)
    ops.gcn_wgrad(x, dG, dW, rowptr, src, coef, K, kdeg)
    ops.gcn_wgrad(x, dG, dW, rowptr, src, coef, K, kdeg)      # accumulates
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    xa = torch.einsum("ntvc,kvw->ntwkc", x.float(), Ahat).to(torch.bfloat16).double()     # what the tensor core sees
    ref = 2 * torch.einsum("ntwkc,ntwo->koc", xa, dG.double()).reshape(K * Cout, Cin)
    err = (dW.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, f"wgrad err {err:.2e}"


@gpu
@pytest.mark.parametrize("strategy,layout", [("uniform", "coco_cut"), ("distance", "mediapipe33"), ("uniform", "mediapipe33")])
def test_gcn_generic_partitions(strategy, layout):
    """K = 1 / 2 partitions (uniform / distance strategies, graph.py:76-97 of the reference): the generic producer path."""
    from fall_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    N, T, Cin, Cout = 3, 11, 128, 64
    A, Ahat, K, V, rowptr, src, coef, x, W, bias, kdeg = _setup(layout, N, T, Cin, Cout, dev, seed=7, strategy=strategy)
    if max(kdeg) > 8:
        pytest.skip("in-degree above the kernel's unroll limit")
    G = torch.empty(N, T, V, Cout, dtype=torch.bfloat16, device=dev)
    ops.gcn_fwd(x, ops.gcn_pack(W, K, Cin, Cout), G, rowptr, src, coef, K, kdeg, bias=bias)
    xa = torch.einsum("ntvc,kvw->ntwkc", x.float(), Ahat).to(torch.bfloat16).float()
    g_ref = torch.einsum("ntwkc,koc->ntwo", xa, W.view(K, Cout, Cin).to(torch.bfloat16).float()) + bias[None, None]
    assert (G.float() - g_ref).abs().max().item() / g_ref.abs().max().item() < 6e-3
    dG = torch.randn(N, T, V, Cout, generator=torch.Generator().manual_seed(1)).to(dev, torch.bfloat16)
    dW = torch.zeros(K * Cout, Cin, device=dev)
    ops.gcn_wgrad(x, dG, dW, rowptr, src, coef, K, kdeg)
    ref = torch.einsum("ntwkc,ntwo->koc", xa.double(), dG.double()).reshape(K * Cout, Cin)
    assert (dW.double() - ref).abs().max().item() / ref.abs().max().item() < 2e-3
    assert int(ops.err_word(dev).item()) == 0


@gpu
@pytest.mark.parametrize("layout,N,T,Cin,Cout", CASES)
@pytest.mark.parametrize("with_addend", [False, True])
def test_gcn_bwd_matches_torch(layout, N, T, Cin, Cout, with_addend):
    """dx and d(edge coefficients) of the graph conv with P = dG.W^T kept on chip, against fp64 torch on the bf16 inputs."""
    from fall_multimodal_b200 import ops
    from fall_multimodal_b200.graph import adjacency_csr

    dev = torch.device("cuda:0")
    A, Ahat, K, V, rowptr, src, coef, x, W, bias, kdeg = _setup(layout, N, T, Cin, Cout, dev, seed=11)
    csr = adjacency_csr(A.double().numpy())
    t = lambda a, dt=torch.int32: torch.as_tensor(a).to(device=dev, dtype=dt)
    perm = torch.as_tensor(csr["bwd_perm"]).long().to(dev)
    rowptr_b, dst_b, kk_b, eid_b = t(csr["bwd_rowptr"]), t(csr["dst"])[perm].contiguous(), t(csr["kk"])[perm].contiguous(), perm.int().contiguous()
    coef_b = coef[perm].contiguous()
    maxdeg = int((torch.as_tensor(csr["bwd_rowptr"])[1:] - torch.as_tensor(csr["bwd_rowptr"])[:-1]).max())
    g = torch.Generator().manual_seed(13)
    dG = torch.randn(N, T, V, Cout, generator=g).to(dev, torch.bfloat16)
    addend = torch.randn(N, T, V, Cin, generator=g).to(dev, torch.bfloat16) if with_addend else None
    dx = torch.full((N, T, V, Cin), float("nan"), dtype=torch.bfloat16, device=dev)
    dcoef = torch.zeros(src.numel(), dtype=torch.float32, device=dev)
    ops.gcn_bwd(dG, ops.gcn_pack_bwd(W, K, Cin, Cout), dx, rowptr_b, dst_b, kk_b, coef_b, K, maxdeg, addend=addend, x=x, eid=eid_b,
                dcoef=dcoef)
    # relu_mask: the same result times (x > 0), bit for bit (x plays the previous block's ReLU output)
    dxm = torch.full_like(dx, float("nan"))
    dcoef2 = torch.zeros_like(dcoef)
    ops.gcn_bwd(dG, ops.gcn_pack_bwd(W, K, Cin, Cout), dxm, rowptr_b, dst_b, kk_b, coef_b, K, maxdeg, addend=addend, x=x, eid=eid_b,
                dcoef=dcoef2, relu_mask=True)
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    assert torch.equal(dxm, torch.where(x > 0, dx, torch.zeros_like(dx))), "relu_mask must only zero the entries where x <= 0"
    assert (dcoef2 - dcoef).abs().max().item() <= 1e-5 * dcoef.abs().max().item()
    # reference: P rounded to bf16 (what the staging tile holds), then exact arithmetic
    Wk = W.view(K, Cout, Cin).to(torch.bfloat16).double()
    P = torch.einsum("ntwo,koc->ntwkc", dG.double(), Wk).to(torch.bfloat16).double()
    dx_ref = torch.einsum("ntwkc,kvw->ntvc", P, Ahat.double()) + (addend.double() if with_addend else 0)
    assert torch.isfinite(dx.float()).all()
    e_dx = (dx.double() - dx_ref).abs().max().item() / dx_ref.abs().max().item()
    assert e_dx < 6e-3, f"dx err {e_dx:.2e}"
    dA_ref = torch.einsum("ntvc,ntwkc->kvw", x.double(), P)              # d loss / d (A*importance)[k,v,w]
    dcoef_ref = dA_ref.flatten()[torch.as_tensor(csr["dense_idx"]).long().to(dev)]
    e_dc = (dcoef.double() - dcoef_ref).abs().max().item() / dcoef_ref.abs().max().item()
    assert e_dc < 2e-3, f"dcoef err {e_dc:.2e}"
