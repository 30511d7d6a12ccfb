import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import tragcn as TG
from fall_multimodal_b200 import _lib as L
dev = torch.device("cuda:0"); V = 25; T = 100; H = 64
print("max active clusters V=25:", L.load().fmm_gruscan_max_clusters(25), " V=14:", L.load().fmm_gruscan_max_clusters(14), " V=30:", L.load().fmm_gruscan_max_clusters(30))
def ww(Din, Co, cs):
    Cp = (Din + H + 1 + 7) // 8 * 8
    W = torch.stack([torch.randn(V, Cp, Co, device=dev) * 0.05, cs[:, None, None] * torch.randn(1, Cp, Co, device=dev) * 0.05]); W[:, :, H + Din + 1:] = 0
    return W
cs = torch.rand(V, device=dev) * 0.5 + 0.75
S = torch.softmax(torch.randn(V, V, device=dev), 1) + torch.eye(V, device=dev)
orig = TG._scan_call
def hooked(mode, *a, **k):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(mode, *a, **k); e1.record(); torch.cuda.synchronize()
    if mode == 1: print(f"   B={a[0]} NC={a[3]} scan: {e0.elapsed_time(e1):.3f} ms = {e0.elapsed_time(e1) * 1e3 / T:.1f} us/step", flush=True)
TG._scan_call = hooked
for B in (32, 256, 384, 416, 448, 480, 512):
    x = torch.randn(B, T, V, 3, device=dev).bfloat16()
    with torch.no_grad():
        for _ in range(2):
            TG._GraphGRUScanP.apply(x, S, ww(3, 128, cs), ww(3, 64, cs), cs); TG._handoff = None
