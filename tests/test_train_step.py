"""GPU: TrainStep (graph-captured zero_grad..optimizer.step with on-device loss/accuracy accumulation)
against the plain eager loop of F2/main.py:100-135 on a twin model."""
import copy

import pytest
import torch

from oracle import stgcn_oracle as O

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_matches_eager_loop(use_graph):
    import fall_multimodal_b200 as fmm
    from fall_multimodal_b200.train import TrainStep

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m1 = fmm.TwoStreamSTGCAN_CNN1D(3, {"layout": "coco_cut", "strategy": "spatial"}, 11, 15, 30).to(dev).train()
    m2 = copy.deepcopy(m1)
    batches = [O.synthetic_batch(8, 16, 14, 11, sensor_len=30, sensor_ch=15, seed=s)[:3] for s in (1, 2, 3)]
    loss_fn = torch.nn.CrossEntropyLoss()
    # --- reference loop (eager, host sync every step) ---
    opt1 = torch.optim.RMSprop(m1.parameters(), lr=1e-4)
    ref_losses, hits = [], 0
    for skel, sensor, tgt in batches:
        skel, sensor, tgt = skel.to(dev), sensor.to(dev), tgt.to(dev)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = m1(skel, sensor)
        loss = loss_fn(pred.float(), tgt)
        loss.backward()
        opt1.step()
        m1.zero_grad()
        ref_losses.append(loss.item())
        hits += (pred.argmax(-1) == tgt.argmax(-1)).sum().item()
    # --- TrainStep: warm-up/capture steps would also update the weights, so they run on a scratch copy of the state ---
    opt2 = torch.optim.RMSprop(m2.parameters(), lr=1e-4, capturable=True)
    state = copy.deepcopy(m2.state_dict())
    skel0, sensor0, tgt0 = (t.to(dev) for t in batches[0])
    ts = TrainStep(m2, opt2, loss_fn, (skel0, sensor0), tgt0, use_graph=use_graph, warmup=1)
    m2.load_state_dict(state)                       # undo the warm-up / capture updates
    for st in opt2.state.values():                  # and the optimizer statistics they left behind
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    ts.reset_stats()
    losses = [ts.run((skel.pin_memory(), sensor.pin_memory()), tgt.pin_memory()).item() for skel, sensor, tgt in batches]
    mean_loss, top1 = ts.stats()
    # every step's loss (steps 2 and 3 see the weights the earlier optimizer steps produced) and the device-side statistics
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b)), (losses, ref_losses)
    assert abs(mean_loss - sum(ref_losses) / 3) < 5e-3 * max(1.0, abs(sum(ref_losses) / 3))
    assert abs(top1 - hits / 24) <= 1 / 24 + 1e-9
    # (parameters are not compared one by one: RMSprop's first steps move zero-gradient parameters, e.g. conv biases in
    # front of a train-mode BatchNorm, by +-lr/sqrt(1-alpha) with the sign of rounding noise)
