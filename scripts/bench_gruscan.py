"""Per-piece timing of the persistent graph-GRU scan at the config-4 shape (B=512, T=300, V=25): input blocking, input-half
pass (gruscan mode 0), forward scan (mode 1), backward scan (mode 2), exports. CUDA events, best of 3."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import tragcn as TG
from fall_multimodal_b200 import _lib as L

B, T, V = int(os.environ.get("B", 512)), int(os.environ.get("T", 300)), 25
dev = torch.device("cuda:0")
torch.manual_seed(0)
H = 64


def ww(Din, Co, cs):
    Cp = (Din + H + 1 + 7) // 8 * 8
    Wn = torch.randn(V, Cp, Co, device=dev) * 0.05
    Lw = torch.randn(1, Cp, Co, device=dev) * 0.05
    W = torch.stack([Wn, cs[:, None, None] * Lw])
    W[:, :, H + Din + 1:] = 0
    return W


def timed(name, fn, n=3):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:40s} {best:9.3f} ms", flush=True)
    return r


cs = torch.rand(V, device=dev) * 0.5 + 0.75
E_ = torch.randn(V, 16, device=dev) * 0.5
S = torch.softmax(torch.relu(E_ @ E_.t()), 1) + torch.eye(V, device=dev)
x = torch.randn(B, T, V, 3, device=dev).bfloat16()
BC, NPW = TG.gruscan_geometry(V)
NC = (B + BC - 1) // BC
timed("block_input (Din=3)", lambda: TG._block_input(x, S, BC, NC))
with torch.no_grad():
    h1 = timed("layer 1 forward, no grad (total)", lambda: TG._GraphGRUScanP.apply(x, S, ww(3, 128, cs), ww(3, 64, cs), cs))
    TG._handoff = None
for Din, xin in ((3, x), (64, h1)):
    Wg, Wu = ww(Din, 128, cs).requires_grad_(True), ww(Din, 64, cs).requires_grad_(True)
    xi = xin.clone().requires_grad_(True)
    out = timed(f"layer Din={Din} forward with grad (total)", lambda: TG._GraphGRUScanP.apply(xi, S, Wg, Wu, cs), n=1)
    g = torch.randn_like(out)
    timed(f"layer Din={Din} backward (total)", lambda: out.backward(g, retain_graph=True), n=1)
    TG._handoff = None
# kernels alone, through the profile hook of _scan_call
orig = TG._scan_call
def hooked(mode, *a, **k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(mode, *a, **k); e1.record(); torch.cuda.synchronize()
    print(f"   gruscan mode {mode} KS={k.get('KS')}: {e0.elapsed_time(e1):.3f} ms", flush=True)
TG._scan_call = hooked
with torch.no_grad():
    for rep in range(2):
        h1 = TG._GraphGRUScanP.apply(x, S, ww(3, 128, cs), ww(3, 64, cs), cs)
        h2 = TG._GraphGRUScanP.apply(h1, S, ww(64, 128, cs), ww(64, 64, cs), cs)
        TG._handoff = None
TG._scan_call = orig
TG.scan_prof = {1: torch.zeros(16, dtype=torch.int64, device=dev), 2: torch.zeros(16, dtype=torch.int64, device=dev)}
with torch.no_grad():
    h1 = TG._GraphGRUScanP.apply(x, S, ww(3, 128, cs), ww(3, 64, cs), cs)
    TG._handoff = None
torch.cuda.synchronize()
c = TG.scan_prof[1].tolist()
names = ["g.wait", "g.items", "g.sync+mix", "g.publish", "g.clsync", "u.wait", "u.items", "u.sync+mix", "u.publish", "u.clsync"]
print("forward scan, cycles per step of CTA 0 / thread 0:", {n: round(v / T) for n, v in zip(names, c)}, "total", round(sum(c) / T))
# the per-step path for comparison (eager launches; the train step replays it from a CUDA graph)
TG.scan_prof = None
with torch.no_grad():
    timed("per-step path, layer 1 forward (eager)", lambda: TG._GraphGRUScan.apply(x, S, ww(3, 128, cs), ww(3, 64, cs)), n=2)
    timed("per-step path, layer 2 forward (eager)", lambda: TG._GraphGRUScan.apply(h1, S, ww(64, 128, cs), ww(64, 64, cs)), n=2)
# backward scan profile (gruscan mode 2)
TG.scan_prof = {2: torch.zeros(16, dtype=torch.int64, device=dev)}
Wg, Wu = ww(64, 128, cs).requires_grad_(True), ww(64, 64, cs).requires_grad_(True)
xi = h1.clone().requires_grad_(True)
out = TG._GraphGRUScanP.apply(xi, S, Wg, Wu, cs)
TG._handoff = None
g = torch.randn_like(out)
TG._scan_call = hooked
out.backward(g)
torch.cuda.synchronize()
c = TG.scan_prof[2].tolist()
names = ["A.elementwise", "A.publish+clsync", "wait", "items", "sync+mix+sync", "post", "publish+clsync"]
print("backward scan, cycles per step of CTA 0 / thread 0:", {n: round(v / T) for n, v in zip(names, c)}, "total", round(sum(c) / T))
