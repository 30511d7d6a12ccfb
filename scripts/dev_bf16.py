"""Dev helper: bf16-mode errors on a larger batch: mask-consistent and natural-mask, vs torch autocast."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import ZERO_GRAD_SUFFIXES
from tests.test_stgcan import build_from_fixture, oracle_with_masks, _rel_errors
from oracle import stgcn_oracle as O
from fall_multimodal_b200.graph import Graph

layout, N, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
A = Graph(layout, "spatial").A
fx = {"config": dict(in_ch=3, layout=layout, strategy="spatial", num_class=11, N=N, T=T),
      "shapes": O.stgcan_param_shapes(3, A.shape[1], A.shape[0], 11), "fill_seed": 5, "batch_seed": 11}
m, skel, target = build_from_fixture(fx, dev, None)
m.train(); m._engine.debug = {}
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = m(skel, None)
    loss = torch.nn.CrossEntropyLoss()(out.float(), target)
loss.backward()
grads = {k: p.grad for k, p in m.named_parameters()}
ograds, oout, flips, worst_pre = oracle_with_masks(m, fx, skel, target, dev)
e_same = _rel_errors(grads, ograds)

def oracle_run(dtype, autocast):
    sd = {k: v.to(dev) for k, v in O.fill_state_dict(fx["shapes"], fx["fill_seed"]).items()}
    sd["A"] = m.A.clone()
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k and k != "A":
            v.requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        o = O.stgcan_forward(sd, skel.to(dtype), training=True)
        l = O.soft_ce(o, target.to(dtype))
    l.backward()
    return o.detach(), {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
o64, truth = oracle_run(torch.float64, False)
oa, auto = oracle_run(torch.float32, True)
e_nat, e_auto = _rel_errors(grads, truth), _rel_errors(auto, truth)
print(f"{layout} N={N} T={T}: logits mine {((out.double()-o64).abs().max()/o64.abs().max()).item():.2e} autocast {((oa.double()-o64).abs().max()/o64.abs().max()).item():.2e}  flips {flips}")
for nm, e in (("mine, identical ReLU decisions", e_same), ("mine, natural", e_nat), ("torch autocast, natural", e_auto)):
    v = sorted(e.values())
    print(f"  {nm:32s} median {statistics.median(v):.2e}  p90 {v[int(0.9*len(v))]:.2e}  worst {v[-1]:.2e}  ({max(e, key=e.get)})")
