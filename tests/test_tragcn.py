"""GPU parity of the TRAGCN family through the C-ABI kernels:
bgemm vs torch.matmul on strided views, the graph-GRU scan / time-axis transformer / head against
the oracle, and the whole TARGCN train step against fixtures generated from the unmodified reference."""
import math

import pytest
import torch

from oracle import tragcn_oracle as TO
from tests.golden_util import check_summary, load

gpu = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------
# bgemm
# ------------------------------------------------------------------------------------------------
def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


@gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 1.5e-2)])
def test_bgemm_plain_and_transposed(dtype, tol):
    from fall_multimodal_b200.tragcn import bgemm
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(0)
    for (G1, G2, M, N, K) in [(1, 1, 64, 64, 32), (2, 3, 70, 45, 100), (1, 5, 25, 128, 72), (3, 1, 300, 62, 300), (1, 1, 1, 11, 777)]:
        A = torch.randn(G1, G2, M, K, generator=g).to(dev, dtype)
        B = torch.randn(G1, G2, K, N, generator=g).to(dev, dtype)
        ref = torch.matmul(A.double(), B.double())
        # row-major A, row-major B (n contiguous)
        Cm = torch.empty(G1, G2, M, N, dtype=torch.float32, device=dev)
        bgemm(A, 0, (G2 * M * K, M * K, K, 1, 0, 0), B, 0, (G2 * K * N, K * N, 1, N, 0, 0), Cm, 0, (G2 * M * N, M * N, N, 1),
              (G1, G2), M, N, (K, 1, 1))
        assert _rel(Cm, ref) < tol * math.sqrt(K / 32 + 1), (G1, G2, M, N, K)
        # transposed storage of both operands (m contiguous / k contiguous)
        At = A.transpose(2, 3).contiguous()   # [K][M]
        Bt = B.transpose(2, 3).contiguous()   # [N][K]
        Cm2 = torch.empty(G1, G2, M, N, dtype=dtype, device=dev)
        bgemm(At, 0, (G2 * M * K, M * K, 1, M, 0, 0), Bt, 0, (G2 * K * N, K * N, K, 1, 0, 0), Cm2, 0, (G2 * M * N, M * N, N, 1),
              (G1, G2), M, N, (K, 1, 1))
        assert _rel(Cm2, ref) < max(tol, 8e-3 if dtype == torch.bfloat16 else 0) * math.sqrt(K / 32 + 1), (G1, G2, M, N, K)


@gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-6), (torch.bfloat16, 2e-2)])
def test_bgemm_epilogues_splitk_multilevel(dtype, tol):
    from fall_multimodal_b200.tragcn import bgemm
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(1)
    M, N, K1, K2, K3 = 50, 37, 3, 5, 14
    A = torch.randn(M, K1, K2, K3, generator=g).to(dev, dtype)
    B = torch.randn(K2, N, K1, K3, generator=g).to(dev, dtype)      # deliberately permuted storage
    bm = torch.randn(M, generator=g).to(dev)
    bn = torch.randn(N, generator=g).to(dev)
    ref = torch.einsum("mabc,bnac->mn", A.double(), B.double())
    out = torch.full((M, N), 0.5, dtype=torch.float32, device=dev)
    bgemm(A, 0, (0, 0, K1 * K2 * K3, K2 * K3, K3, 1), B, 0, (0, 0, K1 * K3, K3, N * K1 * K3, 1), out, 0, (0, 0, N, 1), (1, 1),
          M, N, (K1, K2, K3), alpha=0.25, beta=1, act=0, bias_m=bm, bias_n=bn)
    want = 0.25 * ref + bm.double()[:, None] + bn.double()[None] + 0.5
    assert _rel(out, want) < tol * 3
    out2 = torch.empty(M, N, dtype=dtype, device=dev)
    bgemm(A, 0, (0, 0, K1 * K2 * K3, K2 * K3, K3, 1), B, 0, (0, 0, K1 * K3, K3, N * K1 * K3, 1), out2, 0, (0, 0, N, 1), (1, 1),
          M, N, (K1, K2, K3), act=1)
    assert _rel(out2, ref.clamp_min(0)) < max(tol * 3, 8e-3 if dtype == torch.bfloat16 else 0)
    # split-K with atomics, long contraction, broadcast A (column sums)
    R = 5000
    X = torch.randn(R, N, generator=g).to(dev, dtype)
    one = torch.ones(8, dtype=dtype, device=dev)
    s = torch.zeros(N, dtype=torch.float32, device=dev)
    bgemm(one, 0, (0, 0, 0, 0, 0, 0), X, 0, (0, 0, 1, N, 0, 0), s, 0, (0, 0, 0, 1), (1, 1), 1, N, (R, 1, 1), splitk=13)
    assert _rel(s, X.double().sum(0)) < 1e-5
    # negative stride along one contraction level
    Y = torch.randn(3, 40, generator=g).to(dev, dtype)
    Wm = torch.randn(20, 3, generator=g).to(dev, dtype)
    o3 = torch.empty(20, 30, dtype=torch.float32, device=dev)
    # o3[m][n] = sum_j Wm[m][j] * Y[j][n + 2 - j]
    bgemm(Wm, 0, (0, 0, 3, 1, 0, 0), Y, 2, (0, 0, 1, 40 - 1, 0, 0), o3, 0, (0, 0, 30, 1), (1, 1), 20, 30, (3, 1, 1))
    want3 = sum(Wm.double()[:, j:j + 1] * Y.double()[j, 2 - j:2 - j + 30][None] for j in range(3))
    assert _rel(o3, want3) < tol * 2


# ------------------------------------------------------------------------------------------------
# whole model vs the reference fixtures (fp32) and the oracle (bf16, eval)
# ------------------------------------------------------------------------------------------------
def _build(c, adj, dev):
    from fall_multimodal_b200.tragcn import TARGCN
    m = TARGCN(num_nodes=c["V"], adj=adj, seq_len=c["T"])
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == TO.targcn_param_shapes(V=c["V"], T=c["T"])
    m.load_state_dict(TO.fill_targcn(shapes, c["fill_seed"]))
    return m.to(dev).train()


@gpu
@pytest.mark.parametrize("name", ["targcn_v25_t12", "targcn_v14_t30_adj"])
def test_targcn_matches_reference_fixture(name):
    fx = load(name)
    c = fx["config"]
    dev = _dev()
    m = _build(c, fx["adj"], dev)
    assert sum(p.numel() for p in m.parameters()) == fx["n_params"]
    x, tgt = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=c["batch_seed"])
    logits = m(x.to(dev))
    loss = torch.nn.CrossEntropyLoss()(logits, tgt.to(dev))
    loss.backward()
    ref = fx["logits"]
    assert (logits.cpu() - ref).abs().max().item() / ref.abs().max().item() < 1e-4
    assert abs(loss.item() - fx["loss"]) < 1e-4
    assert (logits.argmax(1).cpu() == ref.argmax(1)).all()
    gs = max(v["amax"] if "amax" in v else float(v["full"].abs().max()) for v in fx["grads"].values())
    worst = 0.0
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        worst = max(worst, check_summary(k, p.grad, fx["grads"][k], 1e-4, atol_scale=1e-3 * gs))
    print(name, "worst grad err", worst)


@gpu
def test_targcn_hidden_states_and_input_grad_vs_oracle():
    """Intermediate tensors (both scan layers, both transformer layers) against the fp64 oracle."""
    dev = _dev()
    c = dict(V=14, T=10, B=3, fill_seed=9)
    m = _build(c, None, dev)
    x, _ = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=4)
    sd = {k: v.detach().cpu().double() for k, v in m.state_dict().items()}
    col = {}
    ref = TO.targcn_forward(sd, x.double(), collect=col)
    E = m.node_embeddings.float()
    with torch.no_grad():
        from fall_multimodal_b200.tragcn import _GraphGRUScan
        V = c["V"]
        S = torch.softmax(torch.relu(E @ E.t()), dim=1) + torch.eye(V, device=dev)
        cur = x.to(dev)
        for i, cell in enumerate(m.encoder.dcrnn_cells):
            Cp = (cell.dim_in + cell.hidden_dim + 1 + 7) // 8 * 8
            cur = _GraphGRUScan.apply(cur.contiguous(), S, cell.gate.stage_weights(E, Cp, cell.hidden_dim), cell.update.stage_weights(E, Cp, cell.hidden_dim))
            err = (cur.cpu().double() - col[f"scan{i}"]).abs().max().item() / col[f"scan{i}"].abs().max().item()
            assert err < 2e-5, (i, err)
        out = m.encoder.trans_layer_T(cur)
        err = (out.cpu().double() - col["trans1"]).abs().max().item() / col["trans1"].abs().max().item()
        assert err < 5e-5, err
        assert (m(x.to(dev)).cpu().double() - ref).abs().max().item() / ref.abs().max().item() < 5e-5


def _summary_err(got, ref_summary, truth, floor):
    """(error of ``got``, error of the fixture values) against the fp64 ``truth`` at the fixture's sample points."""
    t = truth.detach().double().flatten().cpu()
    g = got.detach().double().flatten().cpu()
    if "full" in ref_summary:
        r = ref_summary["full"].double().flatten()
        idx = slice(None)
    else:
        idx = ref_summary["idx"]
        r = ref_summary["vals"].double()
    scale = max(t.abs().max().item(), floor)
    return (g[idx] - t[idx]).abs().max().item() / scale, (r - t[idx]).abs().max().item() / scale


@gpu
def test_targcn_bf16_no_worse_than_reference_autocast():
    """bf16 gate: the reference's own autocast path (fixture from the unmodified modules, MF3/main.py:97) is far
    from fp32 on this recurrent model (logits ~0.7 of the max off, gradients 0.14 median) — ours must be at least
    as close to the fp64 truth, tensor by tensor (median) and in the worst case."""
    import statistics
    fx = load("targcn_v25_t16_autocast")
    c = fx["config"]
    dev = _dev()
    m = _build(c, None, dev)
    x, tgt = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=c["batch_seed"])
    sd = {k: v.detach().cpu().double().requires_grad_(not k.endswith("PE.pe")) for k, v in m.state_dict().items()}
    truth = TO.targcn_forward(sd, x.double())
    torch.nn.CrossEntropyLoss()(truth, tgt.double()).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x.to(dev))
        loss = torch.nn.CrossEntropyLoss()(out.float(), tgt.to(dev))
    assert out.dtype == torch.bfloat16
    loss.backward()
    tmax = truth.abs().max().item()
    e_ours = (out.double().cpu() - truth).abs().max().item() / tmax
    e_ref = (fx["logits"].double() - truth).abs().max().item() / tmax
    print(f"logits: ours {e_ours:.3e}, reference autocast {e_ref:.3e}")
    assert e_ours <= max(1.1 * e_ref, 2e-2)
    gs = max(v.grad.abs().max().item() for v in sd.values() if v.grad is not None)
    ours, ref = [], []
    for k, p in m.named_parameters():
        a, b = _summary_err(p.grad, fx["grads"][k], sd[k].grad, 1e-3 * gs)
        ours.append(a)
        ref.append(b)
    print(f"grads: median ours {statistics.median(ours):.3e} / ref {statistics.median(ref):.3e}; "
          f"worst ours {max(ours):.3e} / ref {max(ref):.3e}")
    assert statistics.median(ours) <= 1.1 * statistics.median(ref)
    assert max(ours) <= 1.25 * max(ref)


@gpu
def test_targcn_inference_no_grad_and_errors():
    dev = _dev()
    c = dict(V=14, T=8, B=2, fill_seed=1)
    m = _build(c, None, dev).eval()
    x, _ = TO.synthetic_clips(2, 8, 14, seed=3)
    with torch.no_grad():
        a = m(x.to(dev))
    b = m(x.to(dev))
    assert torch.equal(a, b.detach())
    with pytest.raises(RuntimeError):
        m(x)  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        m(x[:, :5].to(dev))


@pytest.mark.gpu
def test_targcn_wave_chunks_match_unsplit_batch():
    """TARGCN.wave_chunks: a batch that needs a partial extra wave of the persistent scan runs as [full waves..., remainder] on
    concurrent streams; nothing couples clips, so logits and gradients equal the unsplit batch (bf16 scan path; wave_cap forced
    to 32 clips so that 80 clips split as 32 + 32 + 16)."""
    import fall_multimodal_b200 as fmm
    from oracle import tragcn_oracle as TO

    dev = torch.device("cuda:0")
    V, T, B = 25, 12, 80
    m = fmm.TARGCN(num_nodes=V, adj=None, seq_len=T)
    g = torch.Generator().manual_seed(5)
    m.load_state_dict({k: (v if k.endswith("PE.pe") else torch.randn(v.shape, generator=g) * (0.3 if "node_emb" in k else 0.08))
                       for k, v in m.state_dict().items()})
    m = m.to(dev).train()
    x, tgt = TO.synthetic_clips(B, T, V, seed=2)
    x, tgt = x.to(dev), tgt.to(dev)

    def run(split):
        m.wave_split, m.wave_cap = split, 32
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            assert m.wave_chunks(B) == ([32, 32, 16] if split else None)
            assert m.wave_chunks(64) is None and m.wave_chunks(20) is None
            out = m(x)
        torch.nn.CrossEntropyLoss()(out.float(), tgt).backward()
        torch.cuda.synchronize()
        return out.float().detach(), torch.cat([p.grad.flatten().double() for p in m.parameters() if p.grad is not None])

    o0, g0 = run(False)
    o1, g1 = run(True)
    o2, g2 = run(False)
    noise = (g2 - g0).norm().item() / g0.norm().item()
    e_out = (o1 - o0).abs().max().item() / o0.abs().max().item()
    e_g = (g1 - g0).norm().item() / g0.norm().item()
    print(f"wave chunks: logits {e_out:.2e}, gradient rel-L2 {e_g:.2e} (two unsplit runs: {noise:.2e})")
    assert e_out < 1e-2                       # per-clip arithmetic is the same; bf16 rounding of the per-chunk weight staging only
    assert e_g < max(2e-2, 4 * noise)         # weight gradients: per-chunk bf16/fp32 partial sums added in fp32
