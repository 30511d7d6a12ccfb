"""One forward + backward of a graph-GRU layer (persistent scan kernels) and of the time-axis attention at the config-4 width
(B=480, V=25) with a short sequence for the scan (T=16) - the workload for the ncu captures under profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import tragcn as TG

dev = torch.device("cuda:0")
torch.manual_seed(0)
V, H = 25, 64
B, T = int(os.environ.get("B", 480)), int(os.environ.get("T", 16))


def ww(Din, Co, cs):
    Cp = (Din + H + 1 + 7) // 8 * 8
    W = torch.stack([torch.randn(V, Cp, Co, device=dev) * 0.05, cs[:, None, None] * torch.randn(1, Cp, Co, device=dev) * 0.05])
    W[:, :, H + Din + 1:] = 0
    return W


cs = torch.rand(V, device=dev) * 0.5 + 0.75
S = torch.softmax(torch.randn(V, V, device=dev), 1) + torch.eye(V, device=dev)
x = torch.randn(B, T, V, 64, device=dev).bfloat16().requires_grad_(True)
Wg, Wu = ww(64, 128, cs).requires_grad_(True), ww(64, 64, cs).requires_grad_(True)
for _ in range(2):
    out = TG._GraphGRUScanP.apply(x, S, Wg, Wu, cs)
    TG._handoff = None
    out.backward(torch.randn_like(out))
# attention: 32 clips x 25 joints heads of 300 x 300
Ba, Ta, F = 32, 300, 62
Tp = 320
q = torch.zeros(Ba, F, V, Tp, device=dev)
k = torch.zeros(Ba, F, V, Tp, device=dev)
q[..., :Ta].normal_()
k[..., :Ta].normal_()
q, k = q.bfloat16().requires_grad_(True), k.bfloat16().requires_grad_(True)
v = torch.randn(Ba, Ta, V, 64, device=dev).bfloat16().requires_grad_(True)
for _ in range(2):
    o = TG._AttentionF.apply(q, k, v, True)
    o.backward(torch.randn_like(o))
torch.cuda.synchronize()
print("ok")
