"""Isolated timing of the GEMM-class kernels at the bench shapes (CUDA events, L2 flushed between runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import ops

dev = torch.device("cuda:0")
N, V = 256, 33
dt = torch.bfloat16
only = sys.argv[1] if len(sys.argv) > 1 else ""
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]

def tap(name, T, Cin, Cout, ntaps, stride, prologue, vbias=False):
    if only and only not in name: return
    x = torch.randn(N, T, V, Cin, device=dev).to(dt)
    W = torch.randn(Cout, Cin, ntaps, device=dev) * 0.05
    To = (T - 1) // stride + 1
    pw = ops.tapconv_pack(W, Cout, Cin, Cout, Cin, 0, Cin * ntaps, 0, ntaps, 1, list(range(ntaps)), dt)
    out = torch.empty(N, To, V, Cout, device=dev, dtype=dt)
    sc = torch.rand(Cin, device=dev) + 0.5 if prologue else None
    sh = torch.randn(Cin, device=dev) if prologue else None
    sh_ = list(range(-(ntaps // 2), ntaps // 2 + 1))
    bias = (torch.randn(V, Cout, device=dev) if vbias else torch.randn(Cout, device=dev)) if vbias is not None else None
    f = lambda: ops.tapconv(x, pw, out, shifts=sh_, tj=To, istride=stride, in_scale=sc, in_shift=sh, in_relu=prologue,
                            bias=bias, bias_per_joint=bool(vbias))
    best, med = timeit(f)
    fl = 2.0 * N * V * To * Cin * Cout * ntaps
    byt = (x.numel() + out.numel()) * 2
    print(f"tapconv {name:18s} T={T:3d} {Cin:4d}->{Cout:4d} taps={ntaps} s={stride}: {best:7.1f} us (med {med:7.1f})  {fl/best/1e6:7.1f} TF/s  {byt/best/1e3:6.0f} GB/s")

def wg(name, T, Cin, Cout, ntaps, stride, prologue):
    if only and only not in name: return
    x = torch.randn(N, T, V, Cin, device=dev).to(dt)
    To = (T - 1) // stride + 1
    dy = torch.randn(N, To, V, Cout, device=dev).to(dt)
    dw = torch.zeros(Cout, Cin, ntaps, device=dev)
    sc = torch.rand(Cin, device=dev) + 0.5 if prologue else None
    sh = torch.randn(Cin, device=dev) if prologue else None
    sh_ = list(range(-(ntaps // 2), ntaps // 2 + 1))
    # dw is accumulated as [tap][co][ci] (ci contiguous), like engine.py does
    f = lambda: ops.wgrad(x, dy, dw, shifts=sh_, istride=stride, in_scale=sc, in_shift=sh, in_relu=prologue, s_m=Cout * Cin, s_c2=1, s_co=Cin)
    best, med = timeit(f)
    fl = 2.0 * N * V * To * Cin * Cout * ntaps
    byt = (x.numel() + dy.numel()) * 2
    print(f"wgrad   {name:18s} T={T:3d} {Cin:4d}->{Cout:4d} taps={ntaps} s={stride}: {best:7.1f} us (med {med:7.1f})  {fl/best/1e6:7.1f} TF/s  {byt/best/1e3:6.0f} GB/s")

tap("tcn_fwd_b1", 64, 64, 64, 9, 1, True)
tap("tcn_fwd_b3", 64, 128, 128, 9, 2, True)   # (approx: block 3 has Cin=Cout=128 for the tcn)
tap("tcn_fwd_b4", 32, 128, 128, 9, 1, True)
tap("tcn_fwd_b6", 16, 256, 256, 9, 1, True)
tap("tcn_plain_b6", 16, 256, 256, 9, 1, False)     # what the engine runs: relu(bn1(G)) is materialised, plain cp.async window
tap("tcn_plain_b4", 32, 128, 128, 9, 1, False)
tap("tcn_dgrad_b1", 64, 64, 64, 9, 1, False)
tap("gcn_fwd_b1", 64, 192, 64, 1, 1, False)
tap("gcn_fwd_b1_jointbias", 64, 192, 64, 1, 1, False, vbias=True)   # the engine's graph conv: per-joint bias
tap("gcn_fwd_b1_nobias", 64, 192, 64, 1, 1, False, vbias=None)
tap("gcn_fwd_b4", 32, 384, 128, 1, 1, False)
tap("gcn_fwd_b4_jointbias", 32, 384, 128, 1, 1, False, vbias=True)
tap("gcn_fwd_b6", 16, 768, 256, 1, 1, False)
tap("P_b1", 64, 64, 192, 1, 1, False)
tap("P_b6", 16, 256, 768, 1, 1, False)
tap("res_b3", 64, 64, 128, 1, 2, False)
wg("tcn_wgrad_b1", 64, 64, 64, 9, 1, True)
wg("tcn_wgrad_b4", 32, 128, 128, 9, 1, True)
wg("tcn_wgrad_b5", 32, 256, 256, 9, 2, True)
wg("tcn_wgrad_b6", 16, 256, 256, 9, 1, True)
wg("tcn_wgrad_plain_b1", 64, 64, 64, 9, 1, False)
wg("tcn_wgrad_plain_b3", 64, 128, 128, 9, 2, False)
wg("tcn_wgrad_plain_b5", 32, 256, 256, 9, 2, False)
wg("tcn_wgrad_plain_b6", 16, 256, 256, 9, 1, False)
wg("tcn_wgrad_plain_b4", 32, 128, 128, 9, 1, False)
wg("gcn_wgrad_b1", 64, 192, 64, 1, 1, False)
wg("gcn_wgrad_b6", 16, 768, 256, 1, 1, False)
wg("res_wgrad_b3", 64, 64, 128, 1, 2, False)
