"""The notebooks' class names and call conventions over the same kernels (drop-in boundary).

Every experiment notebook of the reference (e.g. ``GSTCAN_HAR_conv_10kfold.ipynb``) carries its own copy of the model:

* ``StreamSpatialTemporalGraph(in_channels, graph_args, num_class=None, edge_importance_weighting=True)`` (#cell1:L297-359) is
  ``STGCAN`` with the block list named ``st_gcn_networks`` and ``forward(x)`` taking the clip only;
* ``TwoStreamSpatialTemporalGraph(graph_args, num_class)`` (#cell1:L362-416) holds ``pts_stream`` / ``mot_stream`` / ``sensor``
  (a ``BiLSTM(15, 64, 1, 0.3, 11, 'mean')``) / ``fcn = Linear(512 + 11, num_class)``, takes ONE tuple ``(pts, mot, ser)`` — the
  motion stream is precomputed by the dataset — and returns ``F.softmax(out, dim=-1)``, which the loop then feeds to
  ``CrossEntropyLoss`` (softmax applied twice; ``forward_loss`` fuses exactly that).

``state_dict`` keys equal the notebooks' (``tsstg-model_best.pth`` checkpoints load with ``load_state_dict`` directly).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .sensor import BiLSTM
from .stgcan import STGCAN, _compute_dtype


class StreamSpatialTemporalGraph(STGCAN):
    def __init__(self, in_channels, graph_args, num_class=None, edge_importance_weighting=True, **kwargs):
        super().__init__(in_channels, graph_args, num_class, edge_importance_weighting, **kwargs)
        blocks = self.st_gcan_networks
        del self.st_gcan_networks
        self.st_gcn_networks = blocks                 # the notebooks' attribute name = their state_dict prefix
        self._engine.block_key = "st_gcn_networks"

    def forward(self, x, sensor=None):
        return super().forward(x, None)


class TwoStreamSpatialTemporalGraph(nn.Module):
    def __init__(self, graph_args, num_class, edge_importance_weighting=True, **kwargs):
        super().__init__()
        self.pts_stream = StreamSpatialTemporalGraph(3, graph_args, None, edge_importance_weighting, **kwargs)
        self.mot_stream = StreamSpatialTemporalGraph(2, graph_args, None, edge_importance_weighting, **kwargs)
        self.fcn = nn.Linear(256 * 2 + 11, num_class)
        self.sensor = BiLSTM(input_size=15, hidden_size=64, num_layers=1, dropout_prob=0.3, num_classes=11, feature="mean")
        self.compute_dtype = None

    def _features(self, inputs):
        pts, mot, ser = inputs
        dt = _compute_dtype(self)
        self.pts_stream.compute_dtype = self.mot_stream.compute_dtype = dt
        return [self.pts_stream.features(pts).float(), self.mot_stream.features(mot).float(), self.sensor(None, ser).float()], dt

    def forward(self, inputs):
        feats, dt = self._features(inputs)
        with torch.autocast("cuda", enabled=False):
            out = torch.softmax(torch.addmm(self.fcn.bias, torch.cat(feats, dim=-1), self.fcn.weight.t()), dim=-1)
        return out.to(dt) if dt == torch.bfloat16 else out

    def forward_loss(self, inputs, target, label_smoothing: float = 0.0):
        """``(pred, loss)`` of ``CrossEntropyLoss()(self(inputs), target)`` — the notebooks' loop (#cell7:L129) — with the Linear,
        both softmaxes and the loss in one fused kernel pair."""
        from .head import linear_cross_entropy

        feats, dt = self._features(inputs)
        pred, loss = linear_cross_entropy(feats, self.fcn.weight, self.fcn.bias, target, pre_softmax=True,
                                          label_smoothing=label_smoothing)
        return (pred.to(dt) if dt == torch.bfloat16 else pred), loss
