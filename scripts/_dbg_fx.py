import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_stgcan import build_from_fixture, oracle_with_masks
from tests.golden_util import load, ZERO_GRAD_SUFFIXES, grad_scale
dev = torch.device("cuda:0")
for name in ["stgcan_coco_spatial", "stgcan_mmpose_uniform_feat"]:
    fx = load(name)
    m, skel, target = build_from_fixture(fx, dev, torch.float32)
    m.train(); m._engine.debug = {}
    out = m(skel, None)
    loss = torch.nn.CrossEntropyLoss()(out, target) if fx["config"]["num_class"] else out.square().mean()
    loss.backward(); torch.cuda.synchronize()
    grads = {k: p.grad for k, p in m.named_parameters()}
    ograds, oout, flips, worst_pre = oracle_with_masks(m, fx, skel, target, dev)
    gs = max(g.abs().max().item() for g in ograds.values())
    rows = []
    for k, g in grads.items():
        r = ograds[k]
        scale = max(r.abs().max().item(), (1.0 if k.endswith(ZERO_GRAD_SUFFIXES) else 1e-3) * gs)
        e = (g.double() - r).abs().max().item() / scale
        ref = fx["grads"][k]
        ef = None
        if "full" in ref:
            rf = ref["full"].double().flatten().to(dev)
            ef = (g.double().flatten() - rf).abs().max().item() / max(rf.abs().max().item(), 1e-30)
            eo = (r.flatten() - rf).abs().max().item() / max(rf.abs().max().item(), 1e-30)
        rows.append((e, k, ef, eo if ef is not None else None))
    rows.sort(reverse=True)
    print(name, "flips", flips)
    for e, k, ef, eo in rows[:6]:
        print(f"   {k:45s} vs fp64 oracle {e:.2e}   ours vs fixture {ef}   oracle vs fixture {eo}")
    k = "data_bn.weight"
    print("   data_bn.weight ours", grads[k].flatten()[:6].tolist(), "oracle", ograds[k].flatten()[:6].tolist(), "fixture", fx["grads"][k]["full"].flatten()[:6].tolist())
