"""GPU parity of the row-streaming per-node GEMMs (csrc/pnode.cu) against fp64 torch einsums on the same bf16 operands:
the per-joint products of EmbGCN.py:80-86 over all (t, clip) rows, and the V = 1 case used for the Linear layers of TA.py:33-37."""
import pytest
import torch

gpu = pytest.mark.gpu


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


@gpu
@pytest.mark.parametrize("R,V,K,Cp,c0,ncols", [(1000, 25, 128, 136, 0, 136), (777, 25, 128, 136, 64, 72), (515, 14, 64, 72, 0, 72),
                                               (300, 25, 64, 72, 64, 8), (4099, 1, 64, 64, 0, 64), (130, 1, 128, 128, 0, 128)])
def test_pn_dgrad(R, V, K, Cp, c0, ncols):
    from fall_multimodal_b200.tragcn import pn_dgrad
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(R)
    x = torch.randn(R, V, K, generator=g).to(dev).bfloat16()
    W = (torch.randn(V, Cp, K, generator=g) * 0.1).to(dev).bfloat16()
    bias = torch.randn(Cp, generator=g).to(dev)
    for b, relu in ((None, False), (bias, True)):
        out = torch.full((R, V, Cp), 7.0, dtype=torch.bfloat16, device=dev)
        pn_dgrad(x, W, out, c0=c0, ncols=ncols, bias=b, relu=relu)
        ref = torch.einsum("rnk,nck->rnc", x.double(), W.double())
        if b is not None:
            ref = torch.relu(ref + b.double())
        assert _rel(out[..., c0:c0 + ncols], ref[..., c0:c0 + ncols]) < 6e-3
        untouched = torch.cat([out[..., :c0], out[..., c0 + ncols:]], -1)
        assert untouched.numel() == 0 or (untouched == 7.0).all()


@gpu
@pytest.mark.parametrize("P,R,V,Cp,Co", [(2, 3000, 25, 136, 128), (2, 1111, 14, 72, 64), (1, 5000, 1, 64, 64), (1, 70, 3, 128, 128)])
def test_pn_wgrad(P, R, V, Cp, Co):
    from fall_multimodal_b200.tragcn import pn_wgrad
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(R)
    xc = torch.randn(P, R, V, Cp, generator=g).to(dev).bfloat16()
    dy = torch.randn(P, R, V, Co, generator=g).to(dev).bfloat16()
    dW = torch.empty(P, V, Cp, Co, dtype=torch.float32, device=dev)
    pn_wgrad(xc, dy, dW)
    ref = torch.einsum("prnc,prno->pnco", xc.double(), dy.double())
    assert _rel(dW, ref) < 1e-5


@gpu
@pytest.mark.parametrize("R,V,Cp", [(2000, 25, 136), (333, 14, 72), (5, 30, 144)])
def test_pn_ds(R, V, Cp):
    """Supports gradient partials: sum_{row,c} G[row][n][c] X[row][m][c] (EmbGCN.py:83 backward)."""
    from fall_multimodal_b200 import _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(R)
    G = torch.randn(R, V, Cp, generator=g).to(dev).bfloat16()
    X = torch.randn(R, V, Cp, generator=g).to(dev).bfloat16()
    part = torch.empty(L.load().fmm_pn_ds_parts(), 32, 32, dtype=torch.float32, device=dev)
    L.check(L.load().fmm_pn_ds(G.data_ptr(), X.data_ptr(), part.data_ptr(), R, V, Cp, L.stream()), "pn_ds")
    got = part.sum(0)
    ref = torch.einsum("rnc,rmc->nm", G.double(), X.double())
    assert _rel(got[:V, :V], ref) < 1e-5
    assert got[V:].abs().max().item() == 0 and got[:, V:].abs().max().item() == 0
