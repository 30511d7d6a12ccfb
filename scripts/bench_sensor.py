"""BASELINE config 5: sensor-only branch inference on HAR30-shaped windows (128 steps x 6 channels), batch 8192.
Prints one JSON line per model: windows/s, us per time step, achieved GFLOP/s (9.18 MF / window fwd for the BiLSTM,
SURVEY 8a row 10)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import BiLSTM, CNN_BiLSTM

dev = torch.device("cuda:0")
torch.manual_seed(42)
B, T, I = 8192, 128, 6


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


x = torch.randn(B, T, I, device=dev)
m = BiLSTM(I, 64, 1, 0.3, 11, "mean").to(dev).eval()
with torch.no_grad():
    ms = timeit(lambda: m(None, x))
fl = 2 * T * 2 * 4 * 64 * (I + 64) * B
print(json.dumps({"model": "BiLSTM(6,64) inference", "windows_per_s": B / ms * 1e3, "ms": ms, "us_per_time_step": ms * 1e3 / T,
                  "gflops": fl / ms / 1e6}))
host = torch.randn(B, T, I).pin_memory()
def e2e():
    with torch.no_grad():
        return m(None, host.to(dev, non_blocking=True)).argmax(1).cpu()
ms2 = timeit(e2e, 10)
print(json.dumps({"model": "BiLSTM(6,64) inference e2e (pinned host in, labels out)", "windows_per_s": B / ms2 * 1e3, "ms": ms2}))
x2 = torch.randn(B, 30, 15, device=dev)
m2 = CNN_BiLSTM(64, 1, 0.3, 11, "mean").to(dev).eval()
with torch.no_grad():
    ms3 = timeit(lambda: m2(x2))
print(json.dumps({"model": "CNN_BiLSTM inference (30x15 windows)", "windows_per_s": B / ms3 * 1e3, "ms": ms3}))
