"""ORACLE tooling — the bench / cross-check models assembled from the UNMODIFIED reference classes.

``oracle/ref_import.py`` finds the reference tree (``/root/reference`` in the build container, the staged byte-for-byte copy
``oracle/_ref`` on the GPU box). Nothing here restates arithmetic: every layer is the reference's own ``nn.Module``; the only
glue is the late-fusion wiring of ``Fall_2_Spatial_Temporal_SR/Model/combination.py:37-46`` /
``GSTCAN_HAR_conv_10kfold.ipynb#cell1:L362-416`` for the CNN1D sensor variant, whose fusion class exists only as a notebook
cell hard-wired to ``BiLSTM``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ref_import
from . import stgcn_oracle as O

_cache = {}


def gstcan_modules(layout: str):
    """(stgcan module, graph module, bilstm module, combination module) with ``layout`` known to the reference ``Graph``."""
    if "gstcan" not in _cache:
        _cache["gstcan"] = ref_import.load_gstcan()
    stg, g, bl, comb = _cache["gstcan"]
    if layout in O.LAYOUTS and layout not in ("coco_cut", "coco_mmpose"):
        n, e, c = O.LAYOUTS[layout]
        ref_import.install_layout(stg, g, layout, n, e, c)      # 33-node layout by subclassing Graph
    comb.STGCAN = stg.STGCAN
    return stg, g, bl, comb


class RefTwoStreamCNN1D(nn.Module):
    """BASELINE config 2: reference ``STGCAN(3)`` + ``STGCAN(2)`` trunks, the notebook's ``CNN1D`` sensor branch, Linear head.

    Attribute names follow ``combination.py:31-35`` (``stgcan_1``, ``stgcan_2``, ``fc``) + ``cnn``, i.e. the state_dict keys of
    ``fall_multimodal_b200.TwoStreamSTGCAN_CNN1D``, so one ``fill_state_dict`` fills both."""

    def __init__(self, layout, num_class, sensor_channels=15, sensor_len=30):
        super().__init__()
        stg, _, _, _ = gstcan_modules(layout)
        ga = {"layout": layout, "strategy": "spatial"}
        self.stgcan_1 = stg.STGCAN(3, ga, None)
        self.stgcan_2 = stg.STGCAN(2, ga, None)
        ns = ref_import.load_notebook_sensor()
        self.cnn = ns["CNN1D"]()
        if sensor_channels != 15:       # the notebook hard-codes HAR-UP's 15 channels (GSTCAN_UR_conv.ipynb uses 4)
            self.cnn.layer1[0] = nn.Conv1d(sensor_channels, 16, kernel_size=5, padding=2)
        self.fc = nn.Linear(512 + 32 * (sensor_len // 4), num_class)

    def forward(self, skel, sensor):
        mot = skel[:, :2, 1:, :] - skel[:, :2, :-1, :]                      # combination.py:39
        out1 = self.stgcan_1(skel, sensor)
        out2 = self.stgcan_2(mot, sensor)
        out3 = self.cnn(sensor.permute(0, 2, 1)).flatten(1)                 # notebook cell2: CNN1D feature map
        return self.fc(torch.cat([out1, out2, out3], dim=-1))               # combination.py:44-46


def load_filled(mod: nn.Module, seed: int):
    """Fill ``mod`` deterministically (same recipe as the CUDA modules); the adjacency buffers stay the reference's own."""
    sd = mod.state_dict()
    filled = O.fill_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed)
    missing, unexpected = mod.load_state_dict(filled, strict=False)
    assert not unexpected and all(k == "A" or k.endswith(".A") for k in missing), (missing, unexpected)
    return mod


def targcn(V: int, T: int, seed: int = 1):
    """Reference ``TARGCN(num_nodes=V, adj=None)`` with ``seq_len`` re-pointed at ``T`` and a seeded fill
    (the reference leaves its pools uninitialised)."""
    from . import tragcn_oracle as TO

    M = ref_import.load_tragcn(T)
    m = M.TARGCN(num_nodes=V, adj=None)
    sd = m.state_dict()
    m.load_state_dict(TO.fill_targcn({k: tuple(v.shape) for k, v in sd.items()}, seed))
    return m
