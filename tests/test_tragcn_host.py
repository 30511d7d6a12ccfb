"""CPU: host-side logic of the TRAGCN drop-in (no kernels): state_dict layout vs the reference fixture, the
hoisted EmbGCN invariants vs the oracle, loud failure without CUDA."""
import pytest
import torch

from oracle import tragcn_oracle as TO
from tests.golden_util import load


def test_targcn_state_dict_matches_reference_fixture():
    from fall_multimodal_b200 import TARGCN
    for name in ("targcn_v25_t12", "targcn_v14_t30_adj"):
        fx = load(name)
        c = fx["config"]
        m = TARGCN(num_nodes=c["V"], adj=fx["adj"], seq_len=c["T"])
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == fx["shapes"]
        assert sum(p.numel() for p in m.parameters()) == fx["n_params"]
        m.load_state_dict(TO.fill_targcn(fx["shapes"], c["fill_seed"]))      # reference-shaped checkpoints load


def test_sym_norm_adj_and_colscale_match_oracle():
    from fall_multimodal_b200.tragcn import EmbGCN, sym_norm_adj
    g = torch.Generator().manual_seed(0)
    a = torch.rand(14, 14, generator=g)
    adj = (a + a.t()) * (torch.rand(14, 14, generator=g) < 0.4)
    adj = (adj + adj.t()) / 2
    assert torch.allclose(sym_norm_adj(adj), TO.sym_norm_adj(adj), atol=1e-7)
    e = EmbGCN(67, 128, adj, 2, 64)
    assert torch.allclose(e._colscale, torch.softmax(TO.sym_norm_adj(adj), -1).sum(0), atol=1e-7)


@pytest.mark.parametrize("Din,H,Cout", [(3, 64, 128), (64, 64, 64)])
def test_stage_weights_are_the_oracle_invariants_in_kernel_row_order(Din, H, Cout):
    """[0] = E x weights_pool with the bias row, [1] = column-scaled Linear^T with its bias row, rows ordered
    [state | input | bias | pad] (the kernels' cell-input layout) instead of the reference's (input, state)."""
    from fall_multimodal_b200.tragcn import EmbGCN
    V, Cin = 14, Din + H
    g = torch.Generator().manual_seed(1)
    a = torch.rand(V, V, generator=g)
    adj = (a + a.t()) / 2
    m = EmbGCN(Cin, Cout, adj, 2, 64)
    E = torch.randn(V, 64, generator=g)
    sd = {"p.weights_pool": m.weights_pool.detach(), "p.bias_pool": m.bias_pool.detach(),
          "p.linear.weight": m.linear.weight.detach(), "p.linear.bias": m.linear.bias.detach()}
    supports, colscale, weights, bias = TO.emb_gcn_invariants(sd, "p.", E, TO.sym_norm_adj(adj))
    Cp = (Cin + 1 + 7) // 8 * 8
    W = m.stage_weights(E, Cp, H).detach()
    assert W.shape == (2, V, Cp, Cout)
    assert torch.allclose(W[0, :, :H], weights[:, Din:], atol=1e-6) and torch.allclose(W[0, :, H:Cin], weights[:, :Din], atol=1e-6)
    assert torch.allclose(W[0, :, Cin], bias, atol=1e-6)
    lin = colscale[:, None, None] * m.linear.weight.detach().t()[None]
    assert torch.allclose(W[1, :, :H], lin[:, Din:], atol=1e-6) and torch.allclose(W[1, :, H:Cin], lin[:, :Din], atol=1e-6)
    assert torch.allclose(W[1, :, Cin], m.linear.bias.detach()[None].expand(V, Cout), atol=1e-7)
    assert (W[:, :, Cin + 1:] == 0).all()
    # the cell input x_cat . W reproduces the oracle's EmbGCN on a random input (pure torch, fp64)
    x = torch.randn(5, V, Cin, generator=g)
    ref = TO.emb_gcn(x, (supports, colscale, weights, bias), sd, "p.")
    cat = torch.cat([x[..., Din:], x[..., :Din], torch.ones(5, V, 1), torch.zeros(5, V, Cp - Cin - 1)], -1)
    xg = torch.einsum("nm,bmc->bnc", supports, cat)
    xg[..., Cin:] = cat[..., Cin:]                                   # the bias column is not mixed
    pre = torch.einsum("bnc,nco->bno", xg, W[0])
    lin_o = torch.einsum("bnc,nco->bno", cat, W[1])
    assert torch.allclose(pre + lin_o * torch.sigmoid(lin_o), ref, atol=1e-5)


def test_targcn_refuses_cpu_tensors():
    from fall_multimodal_b200 import TARGCN
    m = TARGCN(num_nodes=14, seq_len=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 8, 14, 3))


def test_wave_chunks_split_rule():
    """TARGCN.wave_chunks (host logic): split only when the batch needs a PARTIAL extra wave of the persistent scans."""
    import torch
    import fall_multimodal_b200 as fmm

    m = fmm.TARGCN(num_nodes=25, adj=None, seq_len=8)
    m.compute_dtype = torch.bfloat16
    m.wave_cap = 480                                  # 15 resident clusters x 32 clips (what the library reports on a B200)
    assert m.wave_chunks(512) == [480, 32]
    assert m.wave_chunks(1000) == [480, 480, 40]
    assert m.wave_chunks(480) is None and m.wave_chunks(960) is None and m.wave_chunks(100) is None
    m.wave_split = False
    assert m.wave_chunks(512) is None
    m.wave_split, m.compute_dtype = True, torch.float32   # the fp32 parity mode runs the per-step kernels: nothing to align
    assert m.wave_chunks(512) is None
