"""GPU parity of the persistent graph-GRU scan (csrc/gruscan.cu; GRU.py:17-27, EmbGCN.py:69-89, TRAGCN.py:158-166):
hidden states and every gradient against an fp64 restatement of the reference cell evaluated by torch autograd, with the
host-driven per-step path (already pinned to the reference fixtures by test_tragcn.py) as the yardstick."""
import pytest
import torch

gpu = pytest.mark.gpu
H = 64


def _ref_scan(x, S, WWg, WWu):
    """fp64 graph-GRU layer from the zero state; WW* (2,V,Cp,Co) rows [h | x | bias | pad] as EmbGCN.stage_weights builds them."""
    B, T, V, Din = x.shape
    Cin = H + Din
    h = x.new_zeros(B, V, H)
    one = x.new_ones(B, V, 1)
    outs = []

    def emb(cat, WW):
        mixed = torch.cat([torch.einsum("nm,bmc->bnc", S, cat), one], -1)
        pre = torch.einsum("bnc,nco->bno", mixed, WW[0][:, :Cin + 1])
        lin = torch.einsum("bnc,nco->bno", torch.cat([cat, one], -1), WW[1][:, :Cin + 1])
        return pre + lin * torch.sigmoid(lin)

    for t in range(T):
        zr = torch.sigmoid(emb(torch.cat([h, x[:, t]], -1), WWg))
        z, r = zr[..., :H], zr[..., H:]
        hc = torch.tanh(emb(torch.cat([r * h, x[:, t]], -1), WWu))
        h = z * h + (1 - z) * hc
        outs.append(h)
    return torch.stack(outs, 1)


def _problem(B, T, V, Din, seed, dev):
    g = torch.Generator().manual_seed(seed)
    Cp = (Din + H + 1 + 7) // 8 * 8
    E = torch.randn(V, 16, generator=g) * 0.5
    S = torch.softmax(torch.relu(E @ E.t()), 1) + torch.eye(V)
    cs = torch.rand(V, generator=g) * 0.5 + 0.75

    def ww(Co):
        Wn = torch.randn(V, Cp, Co, generator=g) * 0.08
        Lw = torch.randn(1, Cp, Co, generator=g) * 0.08
        Wl = cs[:, None, None] * Lw
        Wl[:, H + Din] = Lw[0, H + Din]             # the bias row is not column-scaled
        W = torch.stack([Wn, Wl])
        W[:, :, H + Din + 1:] = 0
        return W

    x = torch.randn(B, T, V, Din, generator=g)
    return x.to(dev), S.to(dev), ww(2 * H).to(dev), ww(H).to(dev), cs.to(dev)


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


@gpu
@pytest.mark.parametrize("B,T,V,Din", [(40, 6, 25, 3), (33, 5, 14, 3), (20, 4, 25, 64), (17, 4, 30, 3), (64, 12, 25, 64)])
def test_persistent_scan_matches_fp64_cell(B, T, V, Din):
    from fall_multimodal_b200.tragcn import _GraphGRUScan, _GraphGRUScanP
    dev = torch.device("cuda:0")
    x, S, WWg, WWu, cs = _problem(B, T, V, Din, 7, dev)
    xb = x.bfloat16()
    # fp64 truth on the bf16-rounded input and weights (what both CUDA paths see)
    leaves = [t.double().requires_grad_(True) for t in (xb, S, WWg.bfloat16(), WWu.bfloat16())]
    truth = _ref_scan(*leaves)
    gout = torch.randn(B, T, V, H, generator=torch.Generator().manual_seed(3)).to(dev)
    truth.backward(gout.bfloat16().double())
    res = {}
    for name, fn in (("steps", _GraphGRUScan), ("persistent", _GraphGRUScanP)):
        ins = [xb.clone().requires_grad_(True), S.clone().requires_grad_(True), WWg.clone().requires_grad_(True),
               WWu.clone().requires_grad_(True)]
        out = fn.apply(*ins) if fn is _GraphGRUScan else fn.apply(*ins, cs)
        out.backward(gout.bfloat16())
        torch.cuda.synchronize()
        res[name] = (_rel(out, truth), [_rel(i.grad, l.grad) for i, l in zip(ins, leaves)])
    print(f"B{B} T{T} V{V} Din{Din}: hidden steps {res['steps'][0]:.3e} / persistent {res['persistent'][0]:.3e}; "
          f"grads (x, S, Wg, Wu) steps {['%.2e' % e for e in res['steps'][1]]} / persistent {['%.2e' % e for e in res['persistent'][1]]}")
    # bf16 tolerance of the north star (2e-2), and never worse than the per-step path by more than rounding noise
    assert res["persistent"][0] < 2e-2
    assert res["persistent"][0] <= 1.25 * res["steps"][0] + 2e-3
    for ep, es in zip(res["persistent"][1], res["steps"][1]):
        assert ep <= 1.25 * es + 3e-3, (ep, es)


@gpu
def test_persistent_scan_inference_no_grad_two_layers_handoff():
    """Two stacked layers under no_grad: the second reads the first's blocked state (no re-blocking) and matches the
    result of feeding it the row-major hidden states."""
    from fall_multimodal_b200 import tragcn as TG
    dev = torch.device("cuda:0")
    B, T, V = 24, 5, 25
    x, S, WWg1, WWu1, cs = _problem(B, T, V, 3, 11, dev)
    _, _, WWg2, WWu2, _ = _problem(B, T, V, 64, 12, dev)
    with torch.no_grad():
        h1 = TG._GraphGRUScanP.apply(x.bfloat16(), S, WWg1, WWu1, cs)
        assert TG._handoff is not None and TG._handoff[0] == h1.data_ptr()
        h2 = TG._GraphGRUScanP.apply(h1, S, WWg2, WWu2, cs)
        TG._handoff = None
        h2b = TG._GraphGRUScanP.apply(h1.clone(), S, WWg2, WWu2, cs)
        TG._handoff = None
    torch.cuda.synchronize()
    assert torch.isfinite(h2).all()
    assert _rel(h2, h2b) < 1e-2
