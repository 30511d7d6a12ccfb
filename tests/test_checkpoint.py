"""CPU: reference checkpoint layouts and notebook key names load into the package modules."""
import io

import torch

from oracle import stgcn_oracle as O


def _notebook_keys(sd):
    """The names the notebook classes (TwoStreamSpatialTemporalGraph + StreamSpatialTemporalGraph) give the same tensors."""
    ren = {"stgcan_1": "pts_stream", "stgcan_2": "mot_stream", "lstm": "sensor", "fc": "fcn", "st_gcan_networks": "st_gcn_networks"}
    return {".".join(ren.get(p, p) for p in k.split(".")): v for k, v in sd.items()}


def test_notebook_state_dict_loads_into_fusion_model():
    import warnings
    from fall_multimodal_b200 import TwoStreamSTGCAN_BiLSTM
    from fall_multimodal_b200.checkpoint import load_reference_checkpoint
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        src = TwoStreamSTGCAN_BiLSTM(3, {"layout": "coco_cut", "strategy": "spatial"}, 11, 15)
        dst = TwoStreamSTGCAN_BiLSTM(3, {"layout": "coco_cut", "strategy": "spatial"}, 11, 15)
    sd = src.state_dict()
    filled = O.fill_state_dict({k: tuple(v.shape) for k, v in sd.items()}, 5)
    sd.update(filled)
    src.load_state_dict(sd)
    nb = {"module." + k: v for k, v in _notebook_keys(src.state_dict()).items()}      # as saved from a DataParallel notebook model
    assert any(k.startswith("module.pts_stream.st_gcn_networks.") for k in nb)
    load_reference_checkpoint(dst, nb)
    for (k1, v1), (k2, v2) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1


def test_reference_checkpoint_dict_round_trip():
    from fall_multimodal_b200 import STGCAN
    from fall_multimodal_b200.checkpoint import load_reference_checkpoint, save_checkpoint
    m = STGCAN(3, {"layout": "coco_cut", "strategy": "spatial"}, 11)
    sd = m.state_dict()
    sd.update(O.fill_state_dict({k: tuple(v.shape) for k, v in sd.items()}, 6))
    m.load_state_dict(sd)
    opt = torch.optim.RMSprop(m.parameters(), lr=1e-3)
    buf = io.BytesIO()
    save_checkpoint(buf, m, optimizer=opt, epoch=7, best_acc=0.5)          # checkpoint.pt layout of F2/main.py:329-338
    buf.seek(0)
    m2 = STGCAN(3, {"layout": "coco_cut", "strategy": "spatial"}, 11)
    rest = load_reference_checkpoint(m2, buf)
    assert rest["epoch"] == 7 and rest["best_acc"] == 0.5 and "optimizer" in rest
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    buf = io.BytesIO()
    save_checkpoint(buf, m)                                                  # best_model.pt layout (:323-326)
    buf.seek(0)
    assert set(torch.load(buf, weights_only=False).keys()) == {"model_weight"}


def test_remap_is_idempotent_and_keeps_unknown_keys():
    from fall_multimodal_b200 import STGCAN
    from fall_multimodal_b200.checkpoint import remap_reference_state_dict
    m = STGCAN(3, {"layout": "coco_cut", "strategy": "spatial"}, 11)
    sd = m.state_dict()
    once = remap_reference_state_dict({"module." + k.replace("st_gcan_networks", "st_gcn_networks"): v for k, v in sd.items()}, m)
    assert set(once) == set(sd)
    assert set(remap_reference_state_dict(once, m)) == set(sd)                       # already-native keys pass through
    odd = remap_reference_state_dict({"something.else": torch.zeros(1)}, m)
    assert list(odd) == ["something.else"]                                            # left for load_state_dict(strict) to report
