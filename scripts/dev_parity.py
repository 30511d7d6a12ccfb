"""Dev helper: per-parameter gradient errors of the CUDA STGCAN against a golden fixture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import load, ZERO_GRAD_SUFFIXES, grad_scale
from tests.test_stgcan import build_from_fixture

name = sys.argv[1] if len(sys.argv) > 1 else "stgcan_coco_spatial"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
fx = load(name)
m, skel, target = build_from_fixture(fx, dev, torch.float32 if mode == "fp32" else torch.bfloat16)
m.train()
out = m(skel, None)
loss = torch.nn.CrossEntropyLoss()(out.float(), target) if fx["config"]["num_class"] else out.float().square().mean()
loss.backward()
ref = fx["logits"].to(dev)
print("logits err", ((out.float() - ref).abs().max() / ref.abs().max()).item(), "loss", loss.item(), fx["loss"])
gs = grad_scale(fx["grads"])
rows = []
for k, p in m.named_parameters():
    r = fx["grads"][k]
    g = p.grad.detach().double().flatten().cpu()
    if "full" in r:
        rr = r["full"].double(); e = (g - rr).abs().max().item(); sc = rr.abs().max().item()
    else:
        e = (g[r["idx"]] - r["vals"].double()).abs().max().item(); sc = r["amax"]
    rows.append((e / max(sc, 1e-30), e / gs, k, sc))
rows.sort(reverse=True)
for rel, relg, k, sc in rows[:40]:
    print(f"{rel:10.3e} {relg:10.3e} scale {sc:9.3e} {k}")

# ---- context: how far is the reference's own fp32 result from exact (fp64 oracle) arithmetic? ----
from oracle import stgcn_oracle as O
c = fx["config"]
A = torch.tensor(O.build_adjacency(c["layout"], c["strategy"]), dtype=torch.float64)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in O.fill_state_dict(fx["shapes"], fx["fill_seed"]).items()}
sd64["A"] = A
for k in sd64:
    if sd64[k].is_floating_point() and "running_" not in k and k != "A":
        sd64[k].requires_grad_(True)
sk64 = skel.detach().cpu().double()
lo = O.stgcan_forward(sd64, sk64, training=True)
l64 = O.soft_ce(lo, target.cpu().double()) if c["num_class"] else lo.square().mean()
l64.backward()
print("\nvs fp64 truth:   mine        reference-fp32")
worst_mine = worst_ref = 0
for rel, relg, k, sc in rows[:200]:
    if k.endswith(ZERO_GRAD_SUFFIXES):
        continue
    t = sd64[k].grad.flatten()
    g = dict(m.named_parameters())[k].grad.detach().double().flatten().cpu()
    r = fx["grads"][k]
    tsc = t.abs().max().item()
    em = (g - t).abs().max().item() / tsc
    if "full" in r:
        er = (r["full"].double() - t).abs().max().item() / tsc
    else:
        er = (r["vals"].double() - t[r["idx"]]).abs().max().item() / tsc
    worst_mine, worst_ref = max(worst_mine, em), max(worst_ref, er)
    if em > 2e-5 or er > 2e-5:
        print(f"  {em:10.3e}  {er:10.3e}  {k}")
print("worst vs truth: mine", worst_mine, "reference fp32", worst_ref)
