import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fall_multimodal_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
names = {0: "1 thread plain", 32: "lean: elect + lo/hi", 46: "lean +fence+commit+try_wait/4"}
for N in (64, 256):
    for mode, nm in names.items():
        lib.fmm_debug_mma_probe(N, 2000, 0, 0, mode, 148, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        i, t = out.tolist()
        print(f"N={N:3d} {nm:28s}: issue {i/2000:6.1f} cyc/MMA, issue+drain {t/2000:6.1f} cyc/MMA  (floor {128*N/256:.0f})")

print("operand majorness (lean issue path): a_mn / b_mn = 1 means MN-major (the weight-gradient engine's layout)")
for N in (64, 128, 256):
    for a_mn, b_mn in ((0, 0), (1, 0), (0, 1), (1, 1)):
        lib.fmm_debug_mma_probe(N, 2000, a_mn, b_mn, 32, 148, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        i, t = out.tolist()
        print(f"N={N:3d} a_mn={a_mn} b_mn={b_mn}: issue {i/2000:6.1f} cyc/MMA, issue+drain {t/2000:6.1f} cyc/MMA  (floor {128*N/256:.0f})")
