"""One launch family of csrc/tapconv.cu / wgrad.cu for ncu: prof_tap.py <T> <C> [fwd|wgrad] [stride]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fall_multimodal_b200 import ops

dev = torch.device("cuda:0")
T, C = int(sys.argv[1]), int(sys.argv[2])
what = sys.argv[3] if len(sys.argv) > 3 else "fwd"
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 1
N, V, dt = 256, 33, torch.bfloat16
To = (T - 1) // stride + 1
x = torch.randn(N, T, V, C, device=dev).to(dt)
W = torch.randn(C, C, 9, device=dev) * 0.05
sh = list(range(-4, 5))
if what == "fwd":
    pw = ops.tapconv_pack(W, C, C, C, C, 0, C * 9, 0, 9, 1, list(range(9)), dt)
    out = torch.empty(N, To, V, C, device=dev, dtype=dt)
    bias = torch.randn(C, device=dev)
    for _ in range(3):
        ops.tapconv(x, pw, out, shifts=sh, tj=To, istride=stride, bias=bias)
else:
    dy = torch.randn(N, To, V, C, device=dev).to(dt)
    dw = torch.zeros(9, C, C, device=dev)
    for _ in range(3):
        ops.wgrad(x, dy, dw, shifts=sh, istride=stride, s_m=C * C, s_c2=1, s_co=C)
torch.cuda.synchronize()
print("ok", int(ops.err_word(dev).item()))
