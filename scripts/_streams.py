import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import fall_multimodal_b200 as fmm
from fall_multimodal_b200 import engine as E
dev = torch.device("cuda:0")
model = fmm.TwoStreamSTGCAN_CNN1D(3, {"layout": bench.LAYOUT, "strategy": "spatial"}, bench.NUM_CLASS, bench.SENSOR_C, bench.SENSOR_L).to(dev).train()
model.concurrent_streams = True
skel, sensor, target = (t.to(dev) for t in bench.synthetic(32, 42))
of, ob = E.TrunkEngine.forward, E.TrunkEngine.backward
def f(self, *a, **k):
    print("forward on stream", hex(torch.cuda.current_stream().cuda_stream)); return of(self, *a, **k)
def b(self, *a, **k):
    print("backward on stream", hex(torch.cuda.current_stream().cuda_stream)); return ob(self, *a, **k)
E.TrunkEngine.forward, E.TrunkEngine.backward = f, b
print("main stream", hex(torch.cuda.current_stream().cuda_stream))
with torch.autocast("cuda", dtype=torch.bfloat16):
    _, loss = model.forward_loss(skel, sensor, target)
loss.backward()
torch.cuda.synchronize()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    print("--- inside a side stream (as under graph capture):", hex(s.cuda_stream))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, loss = model.forward_loss(skel, sensor, target)
    loss.backward()
torch.cuda.synchronize()
