import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import ops
dev = torch.device("cuda:0")
N, V = 256, 33
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)
for (T, C) in ((64, 64), (16, 256)):
    X = torch.randn(N, T, V, C, device=dev).bfloat16()
    S = X.numel() * 2
    pool = torch.zeros(N, C, device=dev)
    for nrep in (1, 16, 128):
        st = torch.zeros(2 * nrep * C, dtype=torch.float64, device=dev)
        us = timeit(lambda: ops.colstats(X, st[:nrep * C], st[nrep * C:], pool))
        print(f"T={T} C={C} nrep={nrep:3d} sums+pool: {us:6.1f} us {S/us/1e3:6.0f} GB/s")
    st = torch.zeros(2 * 16 * C, dtype=torch.float64, device=dev)
    print(f"T={T} C={C} pool only        : {timeit(lambda: ops.colstats(X, None, None, pool)):6.1f} us")
    print(f"T={T} C={C} sums only (16)   : {timeit(lambda: ops.colstats(X, st[:16 * C], st[16 * C:], None)):6.1f} us")
    print(f"T={T} C={C} torch sum(dim)   : {timeit(lambda: X.float().sum((0,1,2))):6.1f} us   copy: {timeit(lambda: X.clone()):6.1f} us")
