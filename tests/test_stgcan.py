"""GPU parity of the drop-in STGCAN (CUDA kernels through the C-ABI) against
  (a) the golden fixtures produced by the unmodified reference (tests/golden/*.pt), and
  (b) the oracle restatement run on the same seeded inputs.
Tolerances (BASELINE.json north_star): 1e-4 relative in fp32, 2e-2 in bf16 (relative to the
tensor's max magnitude), identical predicted labels.
"""
import pytest
import torch

from oracle import stgcn_oracle as O
from tests.golden_util import check_grads, check_summary, load

gpu = pytest.mark.gpu
FP32_TOL = 1e-4
BF16_TOL = 2e-2


def build_from_fixture(fx, dev, compute_dtype):
    from fall_multimodal_b200 import STGCAN

    c = fx["config"]
    m = STGCAN(c["in_ch"], {"layout": c["layout"], "strategy": c["strategy"]}, num_class=c["num_class"])
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == fx["shapes"]
    filled = O.fill_state_dict(fx["shapes"], fx["fill_seed"])
    for k, v in filled.items():
        sd[k] = v
    m.load_state_dict(sd)
    m = m.to(dev)
    m.compute_dtype = compute_dtype
    V = m.A.shape[1]
    skel, _, target, _ = O.synthetic_batch(c["N"], c["T"], V, 11, seed=fx["batch_seed"])
    skel = skel[:, : c["in_ch"]].contiguous().to(dev)
    return m, skel, target.to(dev)


@gpu
@pytest.mark.parametrize("name", ["stgcan_coco_spatial", "stgcan_mp33_spatial", "stgcan_mmpose_uniform_feat"])
def test_stgcan_fp32_matches_reference_fixture(name):
    dev = torch.device("cuda:0")
    fx = load(name)
    m, skel, target = build_from_fixture(fx, dev, torch.float32)
    m.train()
    before = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k}
    out = m(skel, None)
    loss = torch.nn.CrossEntropyLoss()(out, target) if fx["config"]["num_class"] else out.square().mean()
    loss.backward()
    torch.cuda.synchronize()
    ref = fx["logits"].to(dev)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < FP32_TOL, f"logits err {err:.3e}"
    assert abs(loss.item() - fx["loss"]) < 1e-4 * max(1.0, abs(fx["loss"]))
    if fx["config"]["num_class"]:
        assert torch.equal(out.argmax(-1), ref.argmax(-1))
    worst = check_grads({k: p.grad for k, p in m.named_parameters()}, fx["grads"], FP32_TOL)
    print(f"{name}: logits err {err:.2e}, worst grad err {worst:.2e}")
    sd = m.state_dict()
    for k, r in fx["running"].items():
        check_summary(k, sd[k], r, FP32_TOL)
    assert int(sd["data_bn.num_batches_tracked"]) == 1
    # eval mode with the original running statistics
    for k, v in before.items():
        sd[k].copy_(v)
    m.eval()
    with torch.no_grad():
        ev = m(skel, None)
    refe = fx["eval_logits"].to(dev)
    assert (ev - refe).abs().max().item() / refe.abs().max().item() < FP32_TOL


@gpu
@pytest.mark.parametrize("name", ["stgcan_coco_spatial", "stgcan_mp33_spatial"])
def test_stgcan_bf16_within_tolerance(name):
    dev = torch.device("cuda:0")
    fx = load(name)
    m, skel, target = build_from_fixture(fx, dev, None)
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(skel, None)
        assert out.dtype == torch.bfloat16
        loss = torch.nn.CrossEntropyLoss()(out.float(), target)
    loss.backward()
    ref = fx["logits"].to(dev)
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < BF16_TOL, f"bf16 logits err {err:.3e}"
    worst = check_grads({k: p.grad for k, p in m.named_parameters()}, fx["grads"], BF16_TOL * 2.5)
    print(f"{name}: bf16 logits err {err:.2e}, worst grad err {worst:.2e}")


@gpu
def test_stgcan_refuses_cpu():
    from fall_multimodal_b200 import STGCAN

    m = STGCAN(3, {"layout": "coco_cut", "strategy": "spatial"}, num_class=11)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 8, 14), None)
