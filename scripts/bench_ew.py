"""Isolated timing of the memory-bound kernels at the bench shapes (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import ops
from fall_multimodal_b200.graph import Graph, adjacency_csr

dev = torch.device("cuda:0")
N, V, K = 256, 33, 3
dt = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
only = sys.argv[1] if len(sys.argv) > 1 else ""

def _graph_us(body, reps=5):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)

_flush_us = None

def timeit(fn, reps=5):
    """GPU time of one launch behind an L2 flush, without the host's launch latency: (flush + kernel) and (flush) are each
    replayed from a CUDA graph and subtracted (events around an eager ctypes launch add ~10 us to a 30 us kernel)."""
    global _flush_us
    fn(); torch.cuda.synchronize()
    if _flush_us is None:
        _flush_us = _graph_us(lambda: flush.zero_())
    return _graph_us(lambda: (flush.zero_(), fn())) - _flush_us

def report(name, us, nbytes):
    print(f"{name:28s} {us:8.1f} us  {nbytes/us/1e3:7.0f} GB/s  ({nbytes/1e6:.0f} MB)")

A = Graph("mediapipe33", "spatial").A
c = adjacency_csr(A)
t = lambda a, d=torch.int32: torch.as_tensor(a).to(device=dev, dtype=d)
rowptr, src, dst, kk = t(c["fwd_rowptr"]), t(c["fwd_src"]), t(c["dst"]), t(c["kk"])
perm = t(c["bwd_perm"], torch.int64)
bwd_rowptr, dst_b, kk_b, eid_b = t(c["bwd_rowptr"]), dst[perm].contiguous(), kk[perm].contiguous(), perm.to(torch.int32)
E = len(c["fwd_src"]); coef = torch.rand(E, device=dev)
NR = ops.NREP
for (T, C) in ((64, 64), (32, 128), (16, 256)):
    S = N * T * V * C * 2
    r = lambda *s: torch.randn(*s, device=dev).to(dt)
    X, Y, U, dY = r(N, T, V, C), r(N, T, V, C), r(N, T, V, C), r(N, T, V, C)
    f32 = lambda *s: torch.rand(*s, device=dev)
    st = torch.zeros(2 * NR * C, dtype=torch.float64, device=dev); pool = torch.zeros(N, C, device=dev)
    tag = f"[T={T},C={C}] "
    if not only or only in "colstats":
        report(tag + "colstats", timeit(lambda: ops.colstats(X, st[:NR * C], st[NR * C:], pool)), S)
    k1, k0 = f32(N, C), f32(N, C)
    out = torch.empty_like(X)
    if not only or only in "block_out":
        report(tag + "block_out(identity)", timeit(lambda: ops.block_out(U, k1, k0, X, None, None, out)), 3 * S)
    S1, S2 = torch.zeros(N, C, device=dev), torch.zeros(N, C, device=dev)
    if not only or only in "blockout_bwd_reduce":
        report(tag + "blockout_bwd_reduce", timeit(lambda: ops.blockout_bwd_reduce(dY, Y, U, None, S1, S2, None)), 3 * S)
    k2 = f32(C); dU = torch.empty_like(X); dPre = torch.empty_like(X)
    sums = torch.zeros(NR * C, dtype=torch.float64, device=dev)
    if not only or only in "bn2_bwd_apply":
        report(tag + "bn2_bwd_apply(identity)", timeit(lambda: ops.bn2_bwd_apply(dY, Y, U, None, k1, k2, k0, None, None, None, dU, None, dPre, sums, None)), 5 * S)
        report(tag + "bn2_bwd_apply(pre-masked)", timeit(lambda: ops.bn2_bwd_apply(dY, None, U, None, k1, k2, k0, None, None, None, dU, None, None, sums, None)), 3 * S)
        report(tag + "blockout_bwd_reduce(pre-masked)", timeit(lambda: ops.blockout_bwd_reduce(dY, None, U, None, S1, S2, None)), 2 * S)
    a1, b1 = f32(C), f32(C) - 0.5
    T1, T2 = torch.zeros(NR * C, dtype=torch.float64, device=dev), torch.zeros(NR * C, dtype=torch.float64, device=dev)
    if not only or only in "bn1_bwd_reduce":
        report(tag + "bn1_bwd_reduce", timeit(lambda: ops.bn1_bwd_reduce(dY, X, a1, b1, T1, T2)), 2 * S)
    Tbl = torch.zeros(NR, V, C, device=dev)
    if not only or only in "bn1_bwd_apply":
        report(tag + "bn1_bwd_apply", timeit(lambda: ops.bn1_bwd_apply(dY, X, a1, b1, a1, b1, a1, dU, Tbl)), 3 * S)
    Xa = torch.empty(N, T, V, K * C, device=dev, dtype=dt)
    if not only or only in "agg_fwd":
        report(tag + "agg_fwd", timeit(lambda: ops.agg_fwd(X, Xa, rowptr, src, coef, K)), 4 * S)
    dcoef = torch.zeros(E, device=dev)
    if not only or only in "agg_bwd":
        report(tag + "agg_bwd(+addend,+dcoef)", timeit(lambda: ops.agg_bwd(Xa, dY, dU, bwd_rowptr, dst_b, kk_b, coef[perm].contiguous(), K, x=X, eid=eid_b, dcoef=dcoef)), 6 * S)
