"""Kernel-level time table (torch.profiler / CUPTI) of one TARGCN train step at the config-4 shape, eager."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fall_multimodal_b200 as fmm
from torch.profiler import profile, ProfilerActivity

B, T, V = int(os.environ.get("B", 512)), int(os.environ.get("T", 300)), 25
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = fmm.TARGCN(num_nodes=V, adj=None, seq_len=T).to(dev).train()
for p_ in m.parameters():
    torch.nn.init.normal_(p_, std=0.05)
x = torch.randn(B, T, V, 3, device=dev)
tgt = torch.randint(0, 11, (B,), device=dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)
        loss = torch.nn.CrossEntropyLoss()(out.float(), tgt)
    loss.backward()
    m.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.2f} ms, {sum(e.count for e in rows)} launches")
for e in rows[:32]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {e.count:5d} x  {e.key[:110]}")
