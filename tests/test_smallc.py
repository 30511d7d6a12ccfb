"""GPU: the narrow-side 1x1 channel mixes of the first GSTCAN block (csrc/smallc.cu) against torch matmul."""
import pytest
import torch

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("dt,tol", [(torch.float32, 2e-6), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("N,T,V,K,Cin,Cout", [(8, 12, 14, 3, 3, 64), (5, 9, 18, 1, 2, 64), (16, 63, 33, 3, 2, 64), (3, 7, 25, 3, 3, 128)])
def test_smallc_fwd_dgrad_wgrad(dt, tol, N, T, V, K, Cin, Cout):
    from fall_multimodal_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 1000 + V)
    Wg = torch.randn(K * Cout, Cin, 1, 1, generator=g).to(dev)            # k-major like gcn.conv.weight (stgcan.py:53)
    Xa = torch.randn(N, T, V, K * Cin, generator=g).to(dev, dt)
    dG = torch.randn(N, T, V, Cout, generator=g).to(dev, dt)
    bias = torch.randn(V, Cout, generator=g).to(dev)
    Weff = Wg.view(K, Cout, Cin).permute(1, 0, 2).reshape(Cout, K * Cin).double()
    assert ops.smallc_ok(K * Cin, Cout)
    G = torch.empty(N, T, V, Cout, dtype=dt, device=dev)
    ops.smallc_fwd(Xa, Wg, G, Cin, Cin, Cout * Cin, 1, bias=bias, bias_per_joint=True)
    ref = Xa.double() @ Weff.t() + bias.double()[None, None]
    assert (G.double() - ref).abs().max().item() / ref.abs().max().item() < tol
    Pm = torch.empty(N, T, V, K * Cin, dtype=dt, device=dev)
    ops.smallc_dgrad(dG, Wg, Pm, Cin, Cin, Cout * Cin, 1)
    refp = dG.double() @ Weff
    assert (Pm.double() - refp).abs().max().item() / refp.abs().max().item() < tol
    dW = torch.zeros(K * Cout, Cin, 1, 1, device=dev)
    ops.smallc_wgrad(Xa, dG, dW, Cin, Cin, Cout * Cin, 1)
    refw = dG.double().reshape(-1, Cout).t() @ Xa.double().reshape(-1, K * Cin)
    got = dW.view(K, Cout, Cin).permute(1, 0, 2).reshape(Cout, K * Cin).double()
    assert (got - refw).abs().max().item() / refw.abs().max().item() < 5e-6


def test_smallc_shape_gate():
    from fall_multimodal_b200 import ops
    assert ops.smallc_ok(9, 64) and ops.smallc_ok(6, 64) and ops.smallc_ok(16, 256)
    assert not ops.smallc_ok(17, 64) and not ops.smallc_ok(9, 60) and not ops.smallc_ok(9, 192)
