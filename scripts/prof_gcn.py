"""One launch family of csrc/gcn.cu for ncu: prof_gcn.py <T> <Cin> <Cout> [fwd|wgrad|bwd]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fall_multimodal_b200 import ops
from fall_multimodal_b200.graph import Graph, adjacency_csr

dev = torch.device("cuda:0")
T, Cin, Cout = (int(a) for a in sys.argv[1:4])
what = sys.argv[4] if len(sys.argv) > 4 else "fwd"
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
s1 = torch.zeros(ops.NREP * Cout, dtype=torch.float64, device=dev)
s2 = torch.zeros_like(s1)
wpk = ops.gcn_pack(W, K, Cin, Cout)
for _ in range(3):
    if what == "fwd":
        ops.gcn_fwd(x, wpk, G, rowptr, src, coef, K, kdeg, bias=bias, ch_sum=s1, ch_sq=s2)
    elif what == "wgrad":
        dW = torch.zeros(K * Cout, Cin, device=dev)
        ops.gcn_wgrad(x, G, dW, rowptr, src, coef, K, kdeg)
torch.cuda.synchronize()
print("ok", int(ops.err_word(dev).item()))
