"""Dev helper: which torch ops launch the small kernels in one train step?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import fall_multimodal_b200 as fmm
from fall_multimodal_b200.parallel import GradBuckets
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = fmm.TwoStreamSTGCAN_CNN1D(3, {"layout": "mediapipe33", "strategy": "spatial"}, 11, 15, 30).to(dev).train()
opt = torch.optim.RMSprop(model.parameters(), lr=1e-3, capturable=True)
buckets = GradBuckets([list(model.fc.parameters()) + list(model.cnn.parameters()), list(model.stgcan_2.parameters()), list(model.stgcan_1.parameters())])
skel, sensor, target = [t.to(dev) for t in bench.synthetic(64, 1)]
loss_fn = torch.nn.CrossEntropyLoss()
def step():
    buckets.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(skel, sensor)
    loss = loss_fn(out.float(), target); loss.backward(); buckets.wait(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="count", row_limit=40, max_name_column_width=60))
