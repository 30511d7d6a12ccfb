"""dev aid: TARGCN vs the reference fixtures, printing every error instead of asserting."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tragcn_oracle as TO
from tests.golden_util import load
from fall_multimodal_b200.tragcn import TARGCN

dev = torch.device("cuda:0")
for name in ["targcn_v25_t12", "targcn_v14_t30_adj"]:
    fx = load(name); c = fx["config"]
    m = TARGCN(num_nodes=c["V"], adj=fx["adj"], seq_len=c["T"])
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(TO.fill_targcn(shapes, c["fill_seed"]))
    m = m.to(dev).train()
    x, tgt = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=c["batch_seed"])
    sd = {k: v.detach().cpu().double() for k, v in m.state_dict().items()}
    col = {}
    TO.targcn_forward(sd, x.double(), adj=fx["adj"], collect=col)
    logits = m(x.to(dev))
    loss = torch.nn.CrossEntropyLoss()(logits, tgt.to(dev)); loss.backward()
    print(name, "logit err", ((logits.cpu() - fx["logits"]).abs().max() / fx["logits"].abs().max()).item(), "loss", loss.item(), fx["loss"])
    gs = max(v["amax"] if "amax" in v else float(v["full"].abs().max()) for v in fx["grads"].values())
    for k, p in m.named_parameters():
        ref = fx["grads"][k]
        g = p.grad.detach().double().flatten().cpu()
        if "full" in ref:
            r = ref["full"].double().flatten(); e = (g - r).abs().max().item(); s = r.abs().max().item()
        else:
            e = (g[ref["idx"]] - ref["vals"].double()).abs().max().item(); s = ref["amax"]
        flag = "  <<<<" if e / max(s, 1e-3 * gs) > 1e-4 else ""
        print(f"  {k:60s} err {e / max(s, 1e-3 * gs):.2e} (amax {s:.2e}){flag}")
