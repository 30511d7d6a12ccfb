"""dev/bench aid: TARGCN train step (BASELINE config 4: T=300, V=25, B=512, bf16) — eager vs CUDA-graph,
with a per-kernel-name time breakdown from torch.profiler on one eager step."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tragcn_oracle as TO
from fall_multimodal_b200 import _lib
from fall_multimodal_b200.tragcn import TARGCN
from fall_multimodal_b200.graphs import GraphedStep

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=512)
ap.add_argument("--T", type=int, default=300)
ap.add_argument("--V", type=int, default=25)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--prof", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = TARGCN(num_nodes=a.V, adj=None, seq_len=a.T)
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict(TO.fill_targcn(shapes, 1))
m = m.to(dev).train()
opt = torch.optim.RMSprop(m.parameters(), lr=1e-3, alpha=0.99, eps=1e-8, capturable=True)
x, tgt = TO.synthetic_clips(a.B, a.T, a.V, seed=42)
x, tgt = x.to(dev), tgt.to(dev)
lossf = torch.nn.CrossEntropyLoss()

def step(x, tgt):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.dtype == "bf16"):
        out = m(x)
    loss = lossf(out.float(), tgt)
    loss.backward()
    opt.step()
    return loss

def timeit(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

print("first losses", [round(step(x, tgt).item(), 4) for _ in range(4)])
l0 = _lib.launch_count
step(x, tgt)
print("launches/step", _lib.launch_count - l0, "peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
gpu_ms, wall_ms = timeit(lambda: step(x, tgt), a.steps)
print(f"eager: {gpu_ms:.1f} ms/step (wall {wall_ms:.1f}) -> {a.B / gpu_ms * 1e3:.0f} clips/s")
if a.prof:
    import fall_multimodal_b200.tragcn as TG
    real = TG.bgemm
    recs = []
    def timed(A, a_off, a_str, B, b_off, b_str, Cm, c_off, c_str, G, M, N, K, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); real(A, a_off, a_str, B, b_off, b_str, Cm, c_off, c_str, G, M, N, K, **kw); e1.record()
        recs.append(((tuple(G), M, N, tuple(K), kw.get("splitk", 1), a_str[2], a_str[3:], b_str[2], b_str[3:]), e0, e1))
    TG.bgemm = timed
    step(x, tgt); torch.cuda.synchronize()
    TG.bgemm = real
    agg = {}
    for k, e0, e1 in recs:
        t = e0.elapsed_time(e1)
        n, tt = agg.get(k, (0, 0.0)); agg[k] = (n + 1, tt + t)
    for k, (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        G, M, N, K, sk = k[:5]
        fl = 2.0 * G[0] * G[1] * M * N * K[0] * K[1] * K[2]
        print(f"  bgemm G{G} M{M} N{N} K{K} sk{sk} a(m{k[5]},k{k[6]}) b(n{k[7]},k{k[8]}): n={n} total {tt:8.2f} ms avg {tt / n * 1e3:8.1f} us  {fl * n / tt / 1e9:7.1f} TF/s")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        step(x, tgt); torch.cuda.synchronize()
    rows = sorted(p.key_averages(), key=lambda r: -r.device_time_total)[:14]
    tot = sum(r.device_time_total for r in p.key_averages())
    print(f"kernel time total {tot / 1e3:.1f} ms")
    for r in rows:
        print(f"  {r.key[:70]:70s} n={r.count:6d} {r.device_time_total / 1e3:8.2f} ms  avg {r.device_time_total / r.count:7.1f} us")
if a.graph:
    t0 = time.perf_counter()
    g = GraphedStep(step, (x, tgt), warmup=1)
    print(f"capture {time.perf_counter() - t0:.1f} s, peak mem GB {torch.cuda.max_memory_allocated() / 2**30:.1f}")
    g.replay()
    gpu_ms, wall_ms = timeit(g.replay, a.steps)
    print(f"graph: {gpu_ms:.1f} ms/step -> {a.B / gpu_ms * 1e3:.0f} clips/s, loss {g.output.item():.4f}")
