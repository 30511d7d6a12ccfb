"""CPU tests: the oracle restatement (oracle/stgcn_oracle.py) against the golden fixtures that
oracle/make_golden.py produced by running the UNMODIFIED reference (tests/golden/*.pt)."""
import os

import pytest
import torch

from oracle import stgcn_oracle as O
from tests.golden_util import check_grads, check_summary, load

TOL = 2e-5  # fp32 CPU, same torch ops in a different association


def _leafify(sd):
    out = {}
    for k, v in sd.items():
        out[k] = v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k and k != "A" else v
    return out


def test_graph_matches_reference():
    ref = load("graph_A")
    for key, A in ref.items():
        layout, strategy = key.split("/")
        mine = torch.tensor(O.build_adjacency(layout, strategy))
        assert mine.shape == A.shape, key
        assert torch.equal(mine, A), key


def test_param_counts_known_answers():
    c = load("param_counts")
    shapes = O.stgcan_param_shapes(3, 14, 3, 11)
    n_params = sum(int(torch.tensor(s).prod()) if len(s) else 1 for k, s in shapes.items()
                   if "running_" not in k and "num_batches" not in k and k != "A")
    assert n_params == c["STGCAN"] == 2125507
    assert len(shapes) == c["STGCAN_keys"] == 190
    # fusion: two trunks (256-d features) + BiLSTM + Linear(512+11, 11)   
    t3 = O.stgcan_param_shapes(3, 14, 3, None)
    t2 = O.stgcan_param_shapes(2, 14, 3, None)
    cnt = lambda sh: sum(int(torch.tensor(s).prod()) if len(s) else 1 for k, s in sh.items()
                         if "running_" not in k and "num_batches" not in k and k != "A")
    assert cnt(t3) + cnt(t2) + c["BiLSTM"] + (512 + 11) * 11 + 11 == c["TwoStreamSTGCAN_BiLSTM"] == 4298291


@pytest.mark.parametrize("name", ["stgcan_coco_spatial", "stgcan_mp33_spatial", "stgcan_mmpose_uniform_feat"])
def test_stgcan_oracle_matches_reference(name):
    fx = load(name)
    c = fx["config"]
    A = torch.tensor(O.build_adjacency(c["layout"], c["strategy"]), dtype=torch.float32)
    K, V = A.shape[0], A.shape[1]
    shapes = O.stgcan_param_shapes(c["in_ch"], V, K, c["num_class"])
    assert {k: tuple(v) for k, v in shapes.items()} == fx["shapes"], "state_dict keys/shapes differ from the reference"
    sd = O.fill_state_dict(shapes, fx["fill_seed"])
    sd["A"] = A
    sd = _leafify(sd)
    skel, _, target, _ = O.synthetic_batch(c["N"], c["T"], V, 11, seed=fx["batch_seed"])
    skel = skel[:, : c["in_ch"]].contiguous()
    upd = {}
    logits = O.stgcan_forward(sd, skel, training=True, update=upd)
    loss = O.soft_ce(logits, target) if c["num_class"] else logits.square().mean()
    loss.backward()
    assert abs(float(loss.detach()) - fx["loss"]) < 1e-5
    assert (logits - fx["logits"]).abs().max() <= TOL * fx["logits"].abs().max()
    check_grads({k: sd[k].grad for k in fx["grads"]}, fx["grads"], TOL * 5)
    for k, ref in fx["running"].items():
        check_summary(k, upd[k], ref, TOL)
    with torch.no_grad():
        ev = O.stgcan_forward(sd, skel, training=False)
    assert (ev - fx["eval_logits"]).abs().max() <= TOL * fx["eval_logits"].abs().max()


def test_bilstm_oracle_matches_reference():
    fx = load("bilstm_mean")
    sd = _leafify(O.fill_state_dict(fx["shapes"], fx["fill_seed"]))
    _, sensor, target, _ = O.synthetic_batch(6, 4, 14, 11, sensor_len=30, sensor_ch=15, seed=fx["batch_seed"])
    upd = {}
    logits = O.bilstm_forward(sd, sensor, training=True, update=upd)
    loss = O.soft_ce(logits, target)
    loss.backward()
    assert abs(float(loss.detach()) - fx["loss"]) < 1e-5
    assert (logits - fx["logits"]).abs().max() <= TOL * fx["logits"].abs().max()
    check_grads({k: sd[k].grad for k in fx["grads"]}, fx["grads"], TOL * 5)
    with torch.no_grad():
        ev = O.bilstm_forward(sd, sensor, training=False)
    assert (ev - fx["eval_logits"]).abs().max() <= TOL * fx["eval_logits"].abs().max()


def test_cnn1d_oracle_matches_reference():
    fx = load("cnn1d")
    sd = _leafify(O.fill_state_dict(fx["shapes"], fx["fill_seed"]))
    _, sensor, _, _ = O.synthetic_batch(5, 4, 14, 11, sensor_len=30, sensor_ch=15, seed=fx["batch_seed"])
    out = O.cnn1d_forward(sd, sensor.permute(0, 2, 1).contiguous(), training=True)
    loss = out.square().mean()
    loss.backward()
    assert abs(float(loss.detach()) - fx["loss"]) < 1e-5
    assert (out - fx["logits"]).abs().max() <= TOL * fx["logits"].abs().max()
    check_grads({k: sd[k].grad for k in fx["grads"]}, fx["grads"], TOL * 5)


def test_fusion_oracle_matches_reference():
    fx = load("two_stream_bilstm")
    c = fx["config"]
    A = torch.tensor(O.build_adjacency(c["layout"], c["strategy"]), dtype=torch.float32)
    sd = O.fill_state_dict(fx["shapes"], fx["fill_seed"])
    sd["stgcan_1.A"] = A
    sd["stgcan_2.A"] = A
    sd = _leafify(sd)
    skel, sensor, target, _ = O.synthetic_batch(c["N"], c["T"], 14, 11, sensor_len=c["L"], sensor_ch=c["I"],
                                                seed=fx["batch_seed"])
    logits = O.two_stream_bilstm_forward(sd, skel, sensor, training=True)
    loss = O.soft_ce(logits, target)
    loss.backward()
    assert abs(float(loss.detach()) - fx["loss"]) < 1e-5
    assert (logits - fx["logits"]).abs().max() <= TOL * fx["logits"].abs().max()
    check_grads({k: sd[k].grad for k in fx["grads"]}, fx["grads"], TOL * 5)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_live_reference_cross_check():
    """Build container only: oracle vs the imported reference module on a fresh random case."""
    from oracle import ref_import

    stg, g, bl, comb = ref_import.load_gstcan()
    mod = stg.STGCAN(3, {"layout": "coco_cut", "strategy": "distance"}, num_class=5)
    sd = {k: v.clone() for k, v in mod.state_dict().items()}
    x = torch.randn(3, 3, 11, 14)
    mod.train()
    ref = mod(x, None)
    mine = O.stgcan_forward(sd, x, training=True)
    assert (ref - mine).abs().max() <= 1e-5 * ref.abs().max()
