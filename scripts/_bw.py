import torch
dev = torch.device("cuda:0")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)
for mb in (138, 276, 1024):
    n = mb * 1024 * 1024 // 2
    x = torch.randn(n, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    xf = x.view(torch.float32)
    us = t(lambda: xf.sum()); print(f"{mb} MB  torch sum(fp32 view): {us:7.1f} us  {mb*1.048576/us*1e3:6.0f} GB/s")
    us = t(lambda: y.copy_(x)); print(f"{mb} MB  copy: {us:7.1f} us  {2*mb*1.048576/us*1e3:6.0f} GB/s (read+write)")
    us = t(lambda: y.zero_()); print(f"{mb} MB  memset: {us:7.1f} us  {mb*1.048576/us*1e3:6.0f} GB/s (write)")
