"""Isolated timing of the fused graph conv (csrc/gcn.cu) at the six bench shapes, next to the round-1 path it replaces
(agg_fwd + 1x1 tapconv + colstats). CUDA events, L2 flushed between runs. Usage: bench_gcn.py [reps] [--wave]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fall_multimodal_b200 import _lib, ops
from fall_multimodal_b200.graph import Graph, adjacency_csr

dev = torch.device("cuda:0")
N, V, K = 256, 33, 3
reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
PEAK = 6460.0


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


A = torch.tensor(Graph("mediapipe33", "spatial").A, dtype=torch.float32)
csr = adjacency_csr(A.double().numpy())
t = lambda a: torch.as_tensor(a).to(device=dev, dtype=torch.int32)
rowptr, src = t(csr["fwd_rowptr"]), t(csr["fwd_src"])
kdeg = ops.partition_degrees(csr['fwd_rowptr'].tolist() if hasattr(csr['fwd_rowptr'], 'tolist') else csr['fwd_rowptr'], K, V)
coef = A.flatten()[torch.as_tensor(csr["dense_idx"]).long()].contiguous().to(dev)
rowptr_b, dst_b, kk_b = t(csr["bwd_rowptr"]), t(csr["dst"])[torch.as_tensor(csr["bwd_perm"]).long().to(dev)].contiguous(), \
    t(csr["kk"])[torch.as_tensor(csr["bwd_perm"]).long().to(dev)].contiguous()
eid_b = torch.as_tensor(csr["bwd_perm"]).to(dev, torch.int32).contiguous()
coef_b = coef[torch.as_tensor(csr["bwd_perm"]).long().to(dev)].contiguous()

for blk, T, Cin, Cout in [(1, 64, 64, 64), (3, 64, 64, 128), (4, 32, 128, 128), (5, 32, 128, 256), (6, 16, 256, 256)]:
    x = torch.randn(N, T, V, Cin, device=dev).to(torch.bfloat16)
    W = torch.randn(K * Cout, Cin, device=dev) * 0.05
    bias = torch.randn(V, Cout, device=dev)
    G = torch.empty(N, T, V, Cout, device=dev, dtype=torch.bfloat16)
    Xa = torch.empty(N, T, V, K * Cin, device=dev, dtype=torch.bfloat16)
    s1 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev)
    s2 = torch.zeros_like(s1)
    wpk = ops.gcn_pack(W, K, Cin, Cout)
    pw = ops.tapconv_pack(W.view(K * Cout, Cin, 1, 1), Cout, K * Cin, Cout, Cin, 0, Cin, Cout * Cin, 1, 0, [0], torch.bfloat16)
    byt = (x.numel() + G.numel()) * 2
    fl = 2.0 * N * T * V * K * Cin * Cout
    t_f = timeit(lambda: ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias, ch_sum=s1, ch_sq=s2))
    t_nostat = timeit(lambda: ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias))
    t_xa = timeit(lambda: ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias, ch_sum=s1, ch_sq=s2, xa=Xa))
    t_agg = timeit(lambda: ops.agg_fwd(x, Xa, rowptr, src, coef, K))
    t_mm = timeit(lambda: ops.tapconv(Xa, pw, G, shifts=[0], tj=T, bias=bias, bias_per_joint=True))
    t_cs = timeit(lambda: ops.colstats(G, s1, s2))
    print(f"block {blk}: T={T:2d} {Cin:3d}->{Cout:3d}  fused {t_f:6.1f} us ({byt / t_f / 1e3:5.0f} GB/s = {byt / t_f / 1e3 / PEAK:.2f} of copy peak, "
          f"{fl / t_f / 1e6:5.0f} TF/s) | no stats {t_nostat:6.1f} | +xa {t_xa:6.1f} | round-1 path agg {t_agg:6.1f} + gemm {t_mm:6.1f} + "
          f"colstats {t_cs:5.1f} = {t_agg + t_mm + t_cs:6.1f} us")
    if hasattr(ops, "gcn_bwd"):
        dGb = torch.randn(N, T, V, Cout, device=dev).to(torch.bfloat16)
        dxb = torch.empty(N, T, V, Cin, device=dev, dtype=torch.bfloat16)
        add = torch.randn(N, T, V, Cin, device=dev).to(torch.bfloat16)
        dco = torch.zeros(src.numel(), device=dev)
        wpb = ops.gcn_pack_bwd(W, K, Cin, Cout)
        maxdeg = int((rowptr_b[1:] - rowptr_b[:-1]).max())
        t_b = timeit(lambda: ops.gcn_bwd(dGb, wpb, dxb, rowptr_b, dst_b, kk_b, coef_b, K, maxdeg, addend=add, x=x, eid=eid_b, dcoef=dco))
        pwT = ops.tapconv_pack(W.view(K * Cout, Cin, 1, 1), K * Cin, Cout, Cin, Cout, Cout * Cin, 1, 0, Cin, 0, [0], torch.bfloat16)
        t_p = timeit(lambda: ops.tapconv(dGb, pwT, Xa, shifts=[0], tj=T))
        t_ab = timeit(lambda: ops.agg_bwd(Xa, add, dxb, rowptr_b, dst_b, kk_b, coef_b, K, x=x, eid=eid_b, dcoef=dco))
        nbb = (dGb.numel() + 3 * dxb.numel()) * 2
        print(f"         bwd-data fused {t_b:6.1f} us ({nbb / t_b / 1e3:5.0f} GB/s of dG+x+addend+dx) | round-1 path P gemm {t_p:6.1f} + agg_bwd {t_ab:6.1f} = {t_p + t_ab:6.1f} us")
    if hasattr(ops, "gcn_wgrad"):
        dG = torch.randn(N, T, V, Cout, device=dev).to(torch.bfloat16)
        dW = torch.zeros(K * Cout, Cin, device=dev)
        t_w = timeit(lambda: ops.gcn_wgrad(x, dG, dW, rowptr, src, coef, K, kdeg))
        t_w0 = timeit(lambda: ops.wgrad(Xa, dG, dW, shifts=[0], c2=Cin, s_m=0, s_c1=Cout * Cin, s_c2=1, s_co=Cin))
        print(f"         wgrad fused {t_w:6.1f} us ({byt / t_w / 1e3:5.0f} GB/s of x+dG) | round-1 wgrad on saved Xa {t_w0:6.1f} us")
