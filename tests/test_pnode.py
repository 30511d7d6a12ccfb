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
