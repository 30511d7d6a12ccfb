"""GPU: the musa Model through the CUDA kernels against the fixture generated from the unmodified
reference (evaluated in fp64: the first-layer gradients of this net are cancellation-heavy, the reference's own fp32 run
is ~1e-3 from the truth there, so each tensor's gate is max(1e-4, 3x the reference-fp32 error))."""
import pytest
import torch
import torch.nn.functional as F

from oracle import musa_oracle as MO
from oracle import stgcn_oracle as O
from tests.golden_util import load

gpu = pytest.mark.gpu


def _build(fx, dev, ablation=False):
    from fall_multimodal_b200.musa import Ablation, Model, adjGraph
    m = (Ablation if ablation else Model)(num_class=11, num_point=14, max_frame=300, graph=adjGraph(layout="coco_cut", strategy="uniform"), bias=True,
                                          edge=True, block_size=41, embed_dim=64, n_stage=1, act_type="tanh")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == fx["shapes"]
    assert torch.allclose(m.state_dict()["stream_pos.0.A"], fx["A"].float(), atol=1e-7)
    sd = m.state_dict()
    sd.update(MO.fill_musa(shapes, fx["config"]["fill_seed"]))
    m.load_state_dict(sd)
    return m.to(dev)


@gpu
def test_dwconv_and_bn_act_against_torch():
    from fall_multimodal_b200.musa import _BNAct, _DWConv
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for (k, s, C, T) in [(3, 1, 128, 30), (5, 2, 128, 29), (1, 1, 192, 15)]:
        x = torch.randn(4, T, 14, C, generator=g).to(dev).requires_grad_(True)
        w = torch.randn(C, 1, k, 1, generator=g).to(dev).requires_grad_(True)
        b = torch.randn(C, generator=g).to(dev).requires_grad_(True)
        y = _DWConv.apply(x, w, b, k, s, (k - 1) // 2)
        xr = x.detach().double().permute(0, 3, 1, 2).requires_grad_(True)
        wr, br = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
        yr = F.conv2d(xr, wr, br, stride=(s, 1), padding=((k - 1) // 2, 0), groups=C)
        go = torch.randn(yr.shape, generator=g, dtype=torch.float64).to(dev)
        yr.backward(go)
        y.backward(go.permute(0, 2, 3, 1).float())
        assert (y.double().permute(0, 3, 1, 2) - yr).abs().max().item() < 1e-5 * yr.abs().max().item()
        assert (x.grad.double().permute(0, 3, 1, 2) - xr.grad).abs().max().item() < 1e-5 * xr.grad.abs().max().item()
        assert (w.grad.double() - wr.grad).abs().max().item() < 2e-5 * wr.grad.abs().max().item()
        assert (b.grad.double() - br.grad).abs().max().item() < 2e-5 * br.grad.abs().max().item()
    for act, fn in [(0, lambda t: t), (1, torch.relu), (2, torch.tanh), (3, F.leaky_relu)]:
        for training in (True, False):
            C = 64
            x = torch.randn(3, 10, 14, C, generator=g).to(dev).requires_grad_(True)
            r = torch.randn(3, 10, 14, C, generator=g).to(dev).requires_grad_(True)
            gam = (torch.rand(C, generator=g) + 0.5).to(dev).requires_grad_(True)
            bet = torch.randn(C, generator=g).to(dev).requires_grad_(True)
            rm, rv = torch.randn(C, generator=g).to(dev) * 0.1, (torch.rand(C, generator=g) + 0.5).to(dev)
            y = _BNAct.apply(x, gam, bet, rm.clone(), rv.clone(), training, 1e-5, 0.1, r, act)
            xr, rr = x.detach().double().requires_grad_(True), r.detach().double().requires_grad_(True)
            gr, br2 = gam.detach().double().requires_grad_(True), bet.detach().double().requires_grad_(True)
            bn = F.batch_norm(xr.permute(0, 3, 1, 2), rm.double().clone(), rv.double().clone(), gr, br2, training, 0.1, 1e-5)
            yr = fn(bn.permute(0, 2, 3, 1) + rr)
            go = torch.randn(yr.shape, generator=g, dtype=torch.float64).to(dev)
            yr.backward(go)
            y.backward(go.float())
            for a, bb in ((y, yr), (x.grad, xr.grad), (r.grad, rr.grad), (gam.grad, gr.grad), (bet.grad, br2.grad)):
                assert (a.double() - bb).abs().max().item() < 2e-5 * max(bb.abs().max().item(), 1e-3), (act, training)


@gpu
@pytest.mark.parametrize("name,ablation", [("musa_coco_uniform", False), ("musa_ablation_coco_uniform", True)])
def test_musa_model_matches_reference_fixture(name, ablation):
    fx = load(name)
    c = fx["config"]
    dev = torch.device("cuda:0")
    m = _build(fx, dev, ablation).train()
    for mod in m.modules():
        if hasattr(mod, "keep_prob"):
            mod.keep_prob = 1                      # the fixture switches the random DropBlock / Dropout off
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    skel, _, target, _ = O.synthetic_batch(c["N"], c["T"], c["V"], 11, seed=c["batch_seed"])
    out = m(skel.to(dev))
    loss = torch.nn.CrossEntropyLoss()(out, target.to(dev))
    loss.backward()
    ref = fx["logits"].double()
    assert (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item() < 1e-4
    assert abs(loss.item() - fx["loss"]) < 1e-4
    assert torch.equal(out.argmax(1).cpu(), ref.argmax(1))
    gs = max(v["amax"] if "amax" in v else float(v["full"].abs().max()) for v in fx["grads"].values())
    worst = 0.0
    for k, p in m.named_parameters():
        if k not in fx["grads"]:
            continue
        refg = fx["grads"][k]
        g = p.grad.detach().double().flatten().cpu()
        if "full" in refg:
            r, gi = refg["full"].double().flatten(), g
            scale = max(r.abs().max().item(), 1e-3 * gs)
        else:
            r, gi = refg["vals"].double(), g[refg["idx"]]
            scale = max(refg["amax"], 1e-3 * gs)
        err = (gi - r).abs().max().item()
        gate = max(1e-4 * scale, 3.0 * fx["ref32_grad_abs_err"][k])
        worst = max(worst, err / scale)
        assert err <= gate, f"{k}: err {err / scale:.2e} of scale, reference fp32 itself {fx['ref32_grad_abs_err'][k] / scale:.2e}"
    print("musa worst grad err (rel to scale)", worst)
    m.eval()
    with torch.no_grad():
        ev = m(skel.to(dev))
    # the train step above moved the running statistics: compare eval against the oracle on the CURRENT state
    sd = {k: v.detach().cpu().double() if v.is_floating_point() else v.cpu() for k, v in m.state_dict().items()}
    evo = MO.musa_forward(sd, skel.double(), training=False, ablation=ablation)
    assert (ev.double().cpu() - evo).abs().max().item() / evo.abs().max().item() < 1e-4


@gpu
def test_musa_model_bf16_and_dropblock_run():
    fx = load("musa_coco_uniform")
    dev = torch.device("cuda:0")
    m = _build(fx, dev).train()
    skel, _, target, _ = O.synthetic_batch(16, 30, 14, 11, seed=2)
    skel, target = skel.to(dev), target.to(dev)
    for mod in m.modules():
        if hasattr(mod, "keep_prob"):
            mod.keep_prob = 1
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ref = m(skel)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(skel)
    assert out.dtype == torch.bfloat16
    assert (out.float() - ref).abs().max().item() / ref.abs().max().item() < 3e-2
    # default training configuration (keep_prob 0.9, Dropout 0.2): random masks, finite loss and gradients for every parameter
    for mod in m.modules():
        if hasattr(mod, "keep_prob"):
            mod.keep_prob = 0.9
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.2
    m.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.CrossEntropyLoss()(m(skel).float(), target)
    loss.backward()
    assert torch.isfinite(loss)
    # (the `edge` parameters of the SepTemporal blocks only feed the non-differentiable DropBlock mask: no gradient, as in the reference)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for k, p in m.named_parameters() if k in fx["grads"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(skel.cpu())
