"""dev aid: achieved throughput of fmm_bgemm on the TRAGCN call-site shapes (back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200.tragcn import bgemm

dev = torch.device("cuda:0")
dt = torch.bfloat16


def run(name, G, M, N, K, a_str, b_str, c_str, asz, bsz, csz, cdt=dt, reps=20, **kw):
    A = torch.randn(asz, device=dev).to(dt)
    B = torch.randn(bsz, device=dev).to(dt)
    Cm = torch.zeros(csz, device=dev, dtype=cdt)
    f = lambda: bgemm(A, kw.pop("a_off", 0) if False else 0, a_str, B, 0, b_str, Cm, 0, c_str, G, M, N, K, **kw)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * G[0] * G[1] * M * N * K[0] * K[1] * K[2]
    by = (asz + bsz) * 2 + csz * Cm.element_size()
    print(f"{name:28s} {us:9.1f} us  {fl / us / 1e6:7.1f} TF/s  {by / us / 1e3:7.1f} GB/s(alloc)")


n = 4096
run("square4096 kcontig", (1, 1), n, n, (n, 1, 1), (0, 0, n, 1, 0, 0), (0, 0, n, 1, 0, 0), (0, 0, n, 1), n * n, n * n, n * n, reps=5)
run("square4096 b ncontig", (1, 1), n, n, (n, 1, 1), (0, 0, n, 1, 0, 0), (0, 0, 1, n, 0, 0), (0, 0, n, 1), n * n, n * n, n * n, reps=5)
run("square4096 a mcontig", (1, 1), n, n, (n, 1, 1), (0, 0, 1, n, 0, 0), (0, 0, n, 1, 0, 0), (0, 0, n, 1), n * n, n * n, n * n, reps=5)
R = 960000
run("linear R x64x64", (1, 1), R, 64, (64, 1, 1), (0, 0, 64, 1, 0, 0), (0, 0, 64, 1, 0, 0), (0, 0, 64, 1), R * 64, 64 * 64, R * 64)
Bc, V, Cp, Co = 512, 25, 136, 128
run("cell stage fwd B512", (2, V), Bc, Co, (Cp, 1, 1), (Bc * V * Cp, Cp, V * Cp, 1, 0, 0), (V * Cp * Co, Cp * Co, 1, Co, 0, 0),
    (Bc * V * Co, Co, V * Co, 1), 2 * Bc * V * Cp, 2 * V * Cp * Co, 2 * Bc * V * Co, cdt=torch.float32, reps=100)
run("cell stage dgrad B512", (2, V), Bc, Cp, (Co, 1, 1), (Bc * V * Co, Co, V * Co, 1, 0, 0), (V * Cp * Co, Cp * Co, Co, 1, 0, 0),
    (Bc * V * Cp, Cp, V * Cp, 1), 2 * Bc * V * Co, 2 * V * Cp * Co, 2 * Bc * V * Cp, cdt=torch.float32, reps=100)
Bq, T, Cc = 128, 300, 64
run("attn QK^T", (Bq, V), T, T, (Cc, 1, 1), (V * T * Cc, T * Cc, Cc, 1, 0, 0), (V * T * Cc, T * Cc, Cc, 1, 0, 0),
    (V * T * 304, T * 304, 304, 1), Bq * V * T * Cc, Bq * V * T * Cc, Bq * V * T * 304, reps=5)
run("attn PV", (Bq, V), T, Cc, (T, 1, 1), (V * T * 304, T * 304, 304, 1, 0, 0), (V * T * Cc, T * Cc, 1, Cc, 0, 0),
    (T * V * Cc, Cc, V * Cc, 1), Bq * V * T * 304, Bq * V * T * Cc, Bq * V * T * Cc, reps=5)
run("timeconv fwd", (Bq, V), T, 62, (3, T, 1), (0, 0, 3 * T, 1, 3, 0), (T * V * Cc, Cc, 1, 1, V * Cc, 0),
    (V * T * Cc, T * Cc, Cc, 1), T * T * 3, Bq * T * V * Cc + 8, Bq * V * T * Cc, reps=3)
