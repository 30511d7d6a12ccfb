"""Dev helper: timeline of one CUDA-graph replay of the bench step (torch profiler / CUPTI): how well do the two trunks'
kernels overlap, and how much of the step has a tensor-core (GEMM-class) kernel in flight?"""
import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import fall_multimodal_b200 as fmm
from fall_multimodal_b200.graphs import GraphedStep
from fall_multimodal_b200.optim import FusedRMSprop

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("B", 256))
model = fmm.TwoStreamSTGCAN_CNN1D(3, {"layout": bench.LAYOUT, "strategy": "spatial"}, bench.NUM_CLASS, bench.SENSOR_C, bench.SENSOR_L).to(dev).train()
opt = FusedRMSprop(model.parameters(), lr=1e-4)
skel, sensor, target = (t.to(dev) for t in bench.synthetic(B, 42))

def step():
    for p in model.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, loss = model.forward_loss(skel, sensor, target)
    loss.backward()
    opt.step()
    return loss

g = GraphedStep(step, (), warmup=3)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay(); torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
GEMM = ("tapconv_kernel", "wgrad_kernel", "gcn_fwd", "gcn_bwd", "gcn_wgrad")
def union(evs):
    tot = 0.0; cur_s = cur_e = None
    for e in sorted(evs, key=lambda e: e["ts"]):
        s, en = e["ts"], e["ts"] + e["dur"]
        if cur_e is None or s > cur_e:
            if cur_e is not None: tot += cur_e - cur_s
            cur_s, cur_e = s, en
        else:
            cur_e = max(cur_e, en)
    if cur_e is not None: tot += cur_e - cur_s
    return tot
gem = [e for e in ev if any(k in e["name"] for k in GEMM)]
print(f"kernels {len(ev)}  wall {t1 - t0:.0f} us  busy(union) {union(ev):.0f} us  sum of durations {sum(e['dur'] for e in ev):.0f} us")
print(f"GEMM-class: {len(gem)} launches, sum {sum(e['dur'] for e in gem):.0f} us, union {union(gem):.0f} us -> no GEMM-class kernel in flight for {t1 - t0 - union(gem):.0f} us")
print(f"B={B}: no kernel at all in flight for {t1 - t0 - union(ev):.0f} us of the step")
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in sorted(streams.items(), key=lambda x: -sum(e["dur"] for e in x[1])):
    print(f"  stream {s}: {len(es)} kernels, sum {sum(e['dur'] for e in es):.0f} us, span {es[0]['ts'] - t0:.0f}..{max(e['ts'] + e['dur'] for e in es) - t0:.0f}")
    short = [e for e in es if e["dur"] < 12]
    print(f"      kernels under 12 us: {len(short)} launches, {sum(e['dur'] for e in short):.0f} us")
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"].split("(")[0][:70]; agg[n][0] += 1; agg[n][1] += e["dur"]
for n, (c, d) in sorted(agg.items(), key=lambda x: -x[1][1])[:70]:
    print(f"  {d:8.0f} us {c:4d}x  avg {d/c:6.1f}  {n}")
# biggest gaps without any GEMM-class kernel
gaps = []
cur = t0
for e in sorted(gem, key=lambda e: e["ts"]):
    if e["ts"] > cur: gaps.append((e["ts"] - cur, cur - t0))
    cur = max(cur, e["ts"] + e["dur"])
gaps.sort(reverse=True)
print("largest GEMM-free gaps (us, at):", [(round(a), round(b)) for a, b in gaps[:12]])
