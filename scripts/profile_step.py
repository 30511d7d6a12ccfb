"""dev aid: per-launch timing of the GEMM-class kernels inside one eager train step of the bench workload
(GPU parked first so intervals are kernel time). Prints every launch sorted by time with its roofline figure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fall_multimodal_b200 as fmm
from fall_multimodal_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(42)
B = 256
model = fmm.TwoStreamSTGCAN_CNN1D(3, {"layout": bench.LAYOUT, "strategy": "spatial"}, bench.NUM_CLASS, bench.SENSOR_C, bench.SENSOR_L).to(dev).train()
opt = torch.optim.RMSprop(model.parameters(), lr=1e-3)
skel, sensor, target = (t.to(dev) for t in bench.synthetic(B, 42))
lossf = torch.nn.CrossEntropyLoss()

def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(skel, sensor)
    lossf(out.float(), target).backward()
    opt.step()

orig_tap, orig_wg = ops.tapconv, ops.wgrad
meta = []
def tap(x, pw, out, **kw):
    meta.append(("tapconv", tuple(x.shape), tuple(out.shape), len(kw["shifts"]), kw.get("istride", 1), kw.get("in_scale") is not None))
    return orig_tap(x, pw, out, **kw)
def wg(x, dy, dw, **kw):
    meta.append(("wgrad", tuple(x.shape), tuple(dy.shape), len(kw["shifts"]), kw.get("istride", 1), kw.get("in_scale") is not None))
    return orig_wg(x, dy, dw, **kw)
import fall_multimodal_b200.engine as eng
for _ in range(2):
    step()
eng.ops.tapconv, eng.ops.wgrad = tap, wg
torch.cuda.synchronize()
ops.profile = []
torch.cuda._sleep(int(0.08 * 1.9e9))
step()
torch.cuda.synchronize()
prof, ops.profile = ops.profile, None
rows = []
for (kind, fl, nb, e0, e1), m in zip(prof, meta):
    us = e0.elapsed_time(e1) * 1e3
    rows.append((us, kind, m, fl / us / 1e6, nb / us / 1e3))
tot = sum(r[0] for r in rows)
print(f"{len(rows)} GEMM launches, {tot / 1e3:.2f} ms")
for us, kind, m, tf, gbs in sorted(rows, key=lambda r: -r[0]):
    print(f"{us:7.1f} us {kind:12s} x{m[1]} -> {m[2]} taps={m[3]} s={m[4]} prologue={m[5]}  {tf:6.0f} TF/s {gbs:6.0f} GB/s")
