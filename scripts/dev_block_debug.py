"""Dev helper: locate the first backward intermediate that deviates from fp64 autograd."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests.golden_util import load
from tests.test_stgcan import build_from_fixture
from oracle import stgcn_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "stgcan_coco_spatial"
dev = torch.device("cuda:0")
fx = load(name)
m, skel, target = build_from_fixture(fx, dev, torch.float32)
m.train()
m._engine.debug = {}
out = m(skel, None)
loss = torch.nn.CrossEntropyLoss()(out, target) if fx["config"]["num_class"] else out.square().mean()
loss.backward()
dbg = m._engine.debug

# fp64 oracle on GPU with retained intermediates
c = fx["config"]
sd = {k: (v.double().to(dev) if v.is_floating_point() else v.to(dev)) for k, v in O.fill_state_dict(fx["shapes"], fx["fill_seed"]).items()}
sd["A"] = torch.tensor(O.build_adjacency(c["layout"], c["strategy"]), dtype=torch.float64, device=dev)
x = skel.double()
N, C, T, V = x.shape
xx = x.permute(0, 3, 1, 2).contiguous().view(N, V * C, T)
xx = O._bn(xx, sd, "data_bn.", True)
xx = xx.view(N, V, C, T).permute(0, 2, 3, 1).contiguous()
inter = []
for i, (cin, cout, stride, res) in enumerate(O.BLOCK_PLAN):
    pre = f"st_gcan_networks.{i}."
    A = sd["A"] * sd[f"edge_importance.{i}"]
    xin = xx
    xin.retain_grad() if xin.requires_grad else None
    if not res:
        r = 0
    elif pre + "residual.0.weight" in sd:
        R = F.conv2d(xin, sd[pre + "residual.0.weight"], sd[pre + "residual.0.bias"], stride=(stride, 1)); R.retain_grad() if R.requires_grad else None
        r = O._bn(R, sd, pre + "residual.1.", True)
    else:
        r = xin
    G = O.graph_conv(xin, A, sd[pre + "gcn.conv.weight"], sd[pre + "gcn.conv.bias"])
    G.requires_grad_(True) if not G.requires_grad else None
    G.retain_grad()
    H = F.relu(O._bn(G, sd, pre + "tcn.0.", True)); H.retain_grad()
    U = F.conv2d(H, sd[pre + "tcn.2.weight"], sd[pre + "tcn.2.bias"], stride=(stride, 1), padding=(4, 0)); U.retain_grad()
    z = O._bn(U, sd, pre + "tcn.3.", True)
    y = O.channel_attention(z, sd, pre + "channel_attention_module.", True)
    Y = F.relu(y + r); Y.retain_grad()
    inter.append(dict(x=xin, G=G, H=H, U=U, Y=Y))
    xx = Y
feat = F.avg_pool2d(xx, xx.shape[2:])
lo = F.conv2d(feat, sd["cls.weight"], sd["cls.bias"]).view(N, -1) if "cls.weight" in sd else feat.view(N, -1)
l64 = O.soft_ce(lo, target.double()) if c["num_class"] else lo.square().mean()
for k in sd:
    if sd[k].is_floating_point() and "running" not in k and k != "A":
        sd[k].requires_grad_(True)
# rerun is needed for param grads; here we only need activation grads, so make the first input a leaf
l64.backward()

def rel(a, b):
    b = b.double(); a = a.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

def cl(t):  # NCHW grad -> channels-last
    return t.permute(0, 2, 3, 1)

for i in reversed(range(7)):
    d, it = dbg[i], inter[i]
    row = [f"blk{i}"]
    row.append(f"dY {rel(d['dY'], cl(it['Y'].grad)):.2e}")
    row.append(f"dU {rel(d['dU'], cl(it['U'].grad)):.2e}")
    row.append(f"dH {rel(d['dH'], cl(it['H'].grad)):.2e}")
    row.append(f"dG {rel(d['dG'], cl(it['G'].grad)):.2e}")
    if it['x'].grad is not None:
        row.append(f"dx {rel(d['dx'], cl(it['x'].grad)):.2e}")
    print("  ".join(row))

print("---- recompute block-6 BN1 backward from the engine's own tensors ----")
d = dbg[6]; b = d["saved"]
G, dH = b["G"].double(), d["dH"].double()
a1, b1, mu, rs = b["a1"].double(), b["b1"].double(), b["mean1"].double(), b["rstd1"].double()
dy1 = dH * ((a1 * G + b1) > 0)
T1 = dy1.sum((0, 1, 2)); T2 = (dy1 * G).sum((0, 1, 2))
print("T1 err", rel(d["T1"], T1), "T2 err", rel(d["T2"], T2))
cnt = G.numel() / G.shape[-1]
t2 = (T2 - mu * T1) * rs
m1, m2 = T1 / cnt, t2 / cnt
dGr = a1 * (dy1 - m1 - (G - mu) * rs * m2)
print("dG(engine) vs recomputed", rel(d["dG"], dGr), " recomputed vs oracle", rel(dGr, cl(inter[6]["G"].grad)))
Gor = cl(inter[6]["G"])
print("G engine vs oracle", rel(b["G"], Gor), "mean1", rel(mu, Gor.mean((0,1,2))), "rstd1", rel(rs, 1/torch.sqrt(Gor.var((0,1,2), unbiased=False)+1e-5)))
print("H oracle vs engine-recomputed", rel(torch.relu(a1*G+b1), cl(inter[6]["H"])))
print("---- fresh autograd BN-ReLU on the oracle's own G / H.grad ----")
Go = inter[6]["G"].detach().clone().requires_grad_(True)
pre = "st_gcan_networks.6."
Hf = F.relu(F.batch_norm(Go, None, None, sd[pre + "tcn.0.weight"].detach(), sd[pre + "tcn.0.bias"].detach(), True, 0.1, 1e-5))
Hf.backward(inter[6]["H"].grad)
print("fresh autograd dG vs oracle-graph G.grad", rel(Go.grad, inter[6]["G"].grad), " vs engine", rel(cl(Go.grad), d["dG"]))
og = cl(inter[6]["G"].grad); mg = d["dG"].double()
print("col sums oracle", og.sum((0,1,2)).abs().max().item(), "mine", mg.sum((0,1,2)).abs().max().item(), "amax", og.abs().max().item(), mg.abs().max().item())
diff = (og - mg)
print("diff per-channel max:", diff.abs().amax((0,1,2))[:8].tolist())
print("diff per-joint max:", diff.abs().amax((0,1,3)).tolist())
print("diff per-n max:", diff.abs().amax((1,2,3)).tolist())
