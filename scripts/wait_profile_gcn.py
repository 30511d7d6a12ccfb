"""Dev helper: wait-site profile of the fused graph-conv forward (cycles blocked per mbarrier wait site, summed over the sampling
lanes): wait_profile_gcn.py <T> <Cin> <Cout>"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import ops, _lib
from fall_multimodal_b200.graph import Graph, adjacency_csr
dev = torch.device("cuda:0")
T, Cin, Cout = (int(a) for a in sys.argv[1:4])
N, V, K = 256, 33, 3
A = torch.tensor(Graph("mediapipe33", "spatial").A, dtype=torch.float32)
csr = adjacency_csr(A.double().numpy())
t = lambda a: torch.as_tensor(a).to(device=dev, dtype=torch.int32)
rowptr, src = t(csr["fwd_rowptr"]), t(csr["fwd_src"])
kdeg = ops.partition_degrees([int(v) for v in csr["fwd_rowptr"]], K, V)
coef = A.flatten()[torch.as_tensor(csr["dense_idx"]).long()].contiguous().to(dev)
x = torch.randn(N, T, V, Cin, device=dev).to(torch.bfloat16)
W = torch.randn(K * Cout, Cin, device=dev) * 0.05
bias = torch.randn(V, Cout, device=dev)
G = torch.empty(N, T, V, Cout, device=dev, dtype=torch.bfloat16)
s1 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev); s2 = torch.zeros_like(s1)
wpk = ops.gcn_pack(W, K, Cin, Cout)
fn = lambda: ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias, ch_sum=s1, ch_sq=s2)
lib = _lib.load()
for _ in range(3): fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): fn()
lib.fmm_debug_wait_profile(1, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 32)(); lib.fmm_debug_wait_profile(0, buf)
us = e0.elapsed_time(e1) * 1e3; cyc = us * 1.9e3
print(f"gcn_fwd T={T} {Cin}->{Cout} (FMM_GCN_TC={os.environ.get('FMM_GCN_TC','1')}): {us:.1f} us = {cyc:.0f} cycles")
TAGS = {1: ("converter/producer <- agg_full | raw_full", None), 2: ("converter/producer <- a_empty", None), 3: ("epilogue <- acc_full", 8), 4: ("loader <- empty", 1),
        5: ("weights <- b_empty", 1), 6: ("mma <- b_full(res)", 1), 7: ("main mma <- acc_empty", 1), 8: ("main mma <- a_full", 1), 9: ("main mma <- b_full", 1),
        10: ("agg mma <- fr_full", 1), 11: ("agg mma <- agg_empty", 1), 20: ("epilogue: wait for the previous TMA store", 8),
        21: ("epilogue: TMEM -> bias -> bf16 -> staging", 8), 22: ("epilogue: store issue + statistics", 8)}
nconv = 4 if os.environ.get('FMM_GCN_TC','1') != '0' and Cout <= 128 else 16
for tg, (label, nw) in TAGS.items():
    if buf[tg]:
        nw = nw or nconv
        per = buf[tg] / 148 / nw
        print(f"   {label:44s} {per:9.0f} cycles per warp = {100*per/cyc:5.1f}% of the kernel")
