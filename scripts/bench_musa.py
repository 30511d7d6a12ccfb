"""BASELINE configs[1] read literally: the musa Model of Multimodal_Fall3/main.py:296-320 (B=256, T=30, V=14, bf16 autocast,
default DropBlock keep_prob 0.9). Prints one JSON line: train clips/s (fwd+bwd+RMSprop), eager and CUDA-graph replay."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stgcn_oracle as O
from fall_multimodal_b200 import _lib
from fall_multimodal_b200.musa import Model, adjGraph
from fall_multimodal_b200.graphs import GraphedStep

dev = torch.device("cuda:0")
torch.manual_seed(42)
B, T, V = 256, 30, 14
m = Model(num_class=11, num_point=V, max_frame=300, graph=adjGraph("coco_cut", "uniform"), bias=True, edge=True, block_size=41,
          embed_dim=64, n_stage=1, act_type="tanh").to(dev).train()
opt = torch.optim.RMSprop(m.parameters(), lr=1e-3, capturable=True)
skel, _, target, _ = O.synthetic_batch(B, T, V, 11, seed=42)
skel, target = skel.to(dev), target.to(dev)
lossf = torch.nn.CrossEntropyLoss()

def step(x, t):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)
    loss = lossf(out.float(), t)
    loss.backward()
    opt.step()
    return loss

def timeit(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for _ in range(3):
    step(skel, target)
l0 = _lib.launch_count
step(skel, target)
launches = _lib.launch_count - l0
eager = timeit(lambda: step(skel, target), 10)
g = GraphedStep(step, (skel, target), warmup=1)
g.replay()
graph = timeit(g.replay, 50)
print(json.dumps({"workload": "musa Model B=256 T=30 V=14 bf16 train step (keep_prob 0.9)", "launches_per_step": launches,
                  "eager_ms": eager, "graph_ms": graph, "clips_per_s_graph": B / graph * 1e3, "loss": g.output.item()}))
if len(sys.argv) > 1 and sys.argv[1] == "prof":
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        step(skel, target); torch.cuda.synchronize()
    rows = sorted(p.key_averages(), key=lambda r: -r.device_time_total)[:22]
    tot = sum(r.device_time_total for r in p.key_averages())
    print(f"kernel time total {tot / 1e3:.2f} ms over {sum(r.count for r in p.key_averages())} launches")
    for r in rows:
        print(f"  {r.key[:80]:80s} n={r.count:4d} {r.device_time_total / 1e3:7.2f} ms  avg {r.device_time_total / r.count:7.1f} us")
