"""dev aid: run the fp32 fixture step with torch.empty() filled with NaN to expose reads of uninitialised memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.use_deterministic_algorithms(True, warn_only=True)
torch.utils.deterministic.fill_uninitialized_memory = True
import warnings; warnings.simplefilter("ignore")
from tests.golden_util import load
from tests.test_stgcan import build_from_fixture
dev = torch.device("cuda:0")
for name in ["stgcan_coco_spatial", "stgcan_mp33_spatial"]:
    for dt in (torch.float32, torch.bfloat16):
        fx = load(name)
        m, skel, target = build_from_fixture(fx, dev, dt)
        m.train()
        out = m(skel, None)
        loss = torch.nn.CrossEntropyLoss()(out.float(), target)
        loss.backward()
        torch.cuda.synchronize()
        bad = [k for k, p in m.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
        print(name, dt, "loss", loss.item(), "non-finite grads:", bad[:8], len(bad))
