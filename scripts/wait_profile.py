"""Dev helper: where do the GEMM kernels' warps block? (per mbarrier wait site, cycles summed over threads)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import ops, _lib

dev = torch.device("cuda:0")
N, V, dt = 256, 33, torch.bfloat16
lib = _lib.load()
TAGS = {1: "producer<-win_empty (x256 thr)", 2: "epilogue<-acc_full (x128)", 3: "loader<-b_empty (x1)", 4: "mma<-acc_empty (x1)",
        5: "mma<-win_full (x1)", 6: "mma<-b_full (x1)", 7: "mma<-peer_b_full (x1)", 8: "mma issue region (x1)", 11: "wg producer<-empty (x256)", 12: "wg epilogue<-acc_full (x128)", 13: "wg mma<-full (x1)"}
NTHR = {1: 8, 2: 4, 3: 1, 4: 1, 5: 1, 6: 1, 7: 1, 8: 1, 11: 8, 12: 4, 13: 1}  # sampling warps (lane 0 of each)

def run(name, fn):
    fn(); torch.cuda.synchronize()
    lib.fmm_debug_wait_profile(1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    lib.fmm_debug_wait_profile(0, buf)
    us = e0.elapsed_time(e1) * 1e3
    cyc = us * 1.9e3  # ~1.9 GHz
    print(f"{name}: {us:.1f} us (~{cyc:.0f} cycles per CTA, 148 CTAs)")
    for t, label in TAGS.items():
        if buf[t]:
            per = buf[t] / 148 / NTHR[t]
            print(f"   {label:34s} blocked {per:9.0f} cycles per thread = {100*per/cyc:5.1f}% of the kernel")

def tap(name, T, Cin, Cout, ntaps, stride, prologue):
    x = torch.randn(N, T, V, Cin, device=dev).to(dt)
    W = torch.randn(Cout, Cin, ntaps, device=dev) * 0.05
    To = (T - 1) // stride + 1
    pw = ops.tapconv_pack(W, Cout, Cin, Cout, Cin, 0, Cin * ntaps, 0, ntaps, 1, list(range(ntaps)), dt)
    out = torch.empty(N, To, V, Cout, device=dev, dtype=dt)
    sc = torch.rand(Cin, device=dev) + 0.5 if prologue else None
    sh = torch.randn(Cin, device=dev) if prologue else None
    sh_ = list(range(-(ntaps // 2), ntaps // 2 + 1))
    run("tapconv " + name, lambda: ops.tapconv(x, pw, out, shifts=sh_, tj=To, istride=stride, in_scale=sc, in_shift=sh, in_relu=prologue))

def wg(name, T, Cin, Cout, ntaps, stride, prologue):
    x = torch.randn(N, T, V, Cin, device=dev).to(dt)
    To = (T - 1) // stride + 1
    dy = torch.randn(N, To, V, Cout, device=dev).to(dt)
    dw = torch.zeros(Cout, Cin, ntaps, device=dev)
    sc = torch.rand(Cin, device=dev) + 0.5 if prologue else None
    sh = torch.randn(Cin, device=dev) if prologue else None
    sh_ = list(range(-(ntaps // 2), ntaps // 2 + 1))
    run("wgrad " + name, lambda: ops.wgrad(x, dy, dw, shifts=sh_, istride=stride, in_scale=sc, in_shift=sh, in_relu=prologue, s_m=1, s_c2=ntaps, s_co=Cin * ntaps))

if len(sys.argv) > 1 and sys.argv[1] == "plain":
    tap("tcn_plain_b1", 64, 64, 64, 9, 1, False)
    tap("tcn_plain_b4", 32, 128, 128, 9, 1, False)
    tap("tcn_plain_b6", 16, 256, 256, 9, 1, False)
    sys.exit(0)
tap("tcn_fwd_b1", 64, 64, 64, 9, 1, True)
tap("tcn_dgrad_b1", 64, 64, 64, 9, 1, False)
tap("tcn_fwd_b6", 16, 256, 256, 9, 1, True)
tap("gcn_fwd_b1", 64, 192, 64, 1, 1, False)
tap("P_b1", 64, 64, 192, 1, 1, False)
wg("tcn_wgrad_b1", 64, 64, 64, 9, 1, True)
wg("tcn_wgrad_b6", 16, 256, 256, 9, 1, True)
wg("gcn_wgrad_b1", 64, 192, 64, 1, 1, False)
wg("tcn_wgrad_b1 (plain cp.async)", 64, 64, 64, 9, 1, False)
wg("tcn_wgrad_b4 (plain cp.async)", 32, 128, 128, 9, 1, False)
wg("tcn_wgrad_b6 (plain cp.async)", 16, 256, 256, 9, 1, False)
