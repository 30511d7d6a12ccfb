"""Late-fusion models over the B200 trunks.

* ``TwoStreamSTGCAN`` / ``TwoStreamSTGCAN_CNN1D`` mirror
  ``/root/reference/Fall_2_Spatial_Temporal_SR/Model/combination.py`` (:9-46): joint stream on
  ``skel`` (3 ch), motion stream on the frame difference of xy (2 ch, T-1 frames), optional sensor
  branch, ``cat -> Linear``. State-dict prefixes ``stgcan_1.``, ``stgcan_2.``, ``fc.`` as there.
* ``TwoStreamSTGCAN_CNN1D`` is BASELINE.json's config 2 ("skeleton+accelerometer fusion, GCN +
  1D-CNN sensor branch"): the notebook ``TwoStreamSpatialTemporalGraph`` pattern
  (GSTCAN_HAR_conv_10kfold.ipynb#cell1:L362-416) with the notebook ``CNN1D`` as the sensor branch.
The two trunks and the sensor branch are independent until the concat, so they are issued on
separate CUDA streams.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .sensor import CNN1D, BiLSTM
from .stgcan import STGCAN, _compute_dtype


class _Streams:
    """Side streams for the independent branches of a fusion model (one set per device)."""

    def __init__(self):
        self._s = {}

    def get(self, device, n):
        key = str(device)
        if key not in self._s or len(self._s[key]) < n:
            self._s[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
        return self._s[key][:n]


class TwoStreamSTGCAN(nn.Module):
    def __init__(self, in_channels, graph_args, num_class):
        super().__init__()
        self.stgcan_1 = STGCAN(3, graph_args, num_class=None)
        self.stgcan_2 = STGCAN(2, graph_args, num_class=None)
        self.fc = nn.Linear(256 * 2, num_class)
        self.compute_dtype = None
        self.pre_softmax = False         # True: return softmax(logits) like the notebook models
        self.concurrent_streams = False  # opt-in: run the independent branches on side streams
        self._streams = _Streams()

    def _branches(self, skel, sensor):
        return []

    def _features(self, skel, sensor=None):
        dt = _compute_dtype(self)
        self.stgcan_1.compute_dtype = self.stgcan_2.compute_dtype = dt
        mot = skel[:, :2, 1:] - skel[:, :2, :-1]          # combination.py:13,39
        extra = self._branches(skel, sensor)
        jobs = [lambda: self.stgcan_1.features(skel), lambda: self.stgcan_2.features(mot)] + extra
        if self.concurrent_streams and skel.is_cuda:
            cur = torch.cuda.current_stream()
            outs = []
            for job, st in zip(jobs, self._streams.get(skel.device, len(jobs))):
                st.wait_stream(cur)
                for t in (skel, mot, sensor):
                    if t is not None:
                        t.record_stream(st)
                with torch.cuda.stream(st):
                    outs.append(job())
            for o, st in zip(outs, self._streams.get(skel.device, len(jobs))):
                cur.wait_stream(st)
                o.record_stream(cur)
        else:
            outs = [job() for job in jobs]
        return outs, dt

    def forward(self, skel, sensor=None):
        outs, dt = self._features(skel, sensor)
        x = torch.cat([o.float() for o in outs], dim=-1)
        with torch.autocast("cuda", enabled=False):
            out = torch.addmm(self.fc.bias, x, self.fc.weight.t())
        if self.pre_softmax:                                   # notebook models: F.softmax(out, dim=-1)
            out = torch.softmax(out, dim=-1)
        return out.to(dt) if dt == torch.bfloat16 else out

    def forward_loss(self, skel, sensor, target, label_smoothing: float = 0.0):
        """``(pred, loss)`` of ``CrossEntropyLoss(label_smoothing)(self(skel, sensor), target)`` with the late-fusion Linear and
        the loss in one fused kernel pair (csrc/head.cu; combination.py:44-46 + F2/main.py:113)."""
        from .head import linear_cross_entropy

        outs, dt = self._features(skel, sensor)
        pred, loss = linear_cross_entropy([o.float() for o in outs], self.fc.weight, self.fc.bias, target,
                                          pre_softmax=self.pre_softmax, label_smoothing=label_smoothing)
        return (pred.to(dt) if dt == torch.bfloat16 else pred), loss


class TwoStreamSTGCAN_CNN1D(TwoStreamSTGCAN):
    def __init__(self, in_channels, graph_args, num_class, sensor_channels=15, sensor_len=30):
        super().__init__(in_channels, graph_args, num_class)
        self.cnn = CNN1D(sensor_channels, sensor_len)
        self.fc = nn.Linear(256 * 2 + 32 * (sensor_len // 4), num_class)

    def _branches(self, skel, sensor):
        def run():
            f = self.cnn.forward_channels_last(sensor)        # (N, L/4, 32)
            return f.permute(0, 2, 1).flatten(1)               # torch's (N, 32, L/4).flatten(1) order
        return [run]


class TwoStreamSTGCAN_BiLSTM(TwoStreamSTGCAN):
    """Reference ``TwoStreamSTGCAN_BiLSTM(in_channels, graph_args, num_class, bilstm_input_size=15)``
    (combination.py:27-46): the sensor branch contributes its ``num_class`` logits to the concat;
    state-dict prefixes ``stgcan_1.``, ``stgcan_2.``, ``lstm.``, ``fc.``."""

    def __init__(self, in_channels, graph_args, num_class, bilstm_input_size=15):
        super().__init__(in_channels, graph_args, num_class)
        self.lstm = BiLSTM(input_size=bilstm_input_size, hidden_size=64, num_layers=1, dropout_prob=0.3,
                           num_classes=num_class, feature="mean")
        self.fc = nn.Linear(256 * 2 + num_class, num_class)

    def _branches(self, skel, sensor):
        return [lambda: self.lstm(None, sensor)]


def bone_stream(skel: torch.Tensor, parents: torch.Tensor) -> torch.Tensor:
    """bone[v] = joint[v] - joint[parent(v)] over the layout's neighbour links (root: zero vector)."""
    return skel - skel.index_select(3, parents)


def layout_parents(layout: str) -> list:
    """Parent joint of every joint: breadth-first tree from the layout's centre over its neighbour links."""
    from collections import deque

    from .graph import _LAYOUTS

    num_node, links, center = _LAYOUTS[layout]
    nbr = [[] for _ in range(num_node)]
    for i, j in links:
        nbr[i].append(j)
        nbr[j].append(i)
    parent = list(range(num_node))
    seen, q = {center}, deque([center])
    while q:
        u = q.popleft()
        for w in sorted(nbr[u]):
            if w not in seen:
                seen.add(w)
                parent[w] = u
                q.append(w)
    return parent


class ThreeStreamSTGCAN(nn.Module):
    """Joint / bone / motion 3-stream GCN of BASELINE config 3. The reference file ``3_stream_fall.py`` is
    empty: this follows ``TwoStreamSTGCAN`` (combination.py:9-25) with a third trunk on the
    bone vectors; joint and motion streams are pinned by the reference, the bone stream is not."""

    def __init__(self, in_channels, graph_args, num_class):
        super().__init__()
        self.stgcan_1 = STGCAN(3, graph_args, num_class=None)   # joints (x, y, score)
        self.stgcan_2 = STGCAN(2, graph_args, num_class=None)   # motion (dx, dy), T-1 frames
        self.stgcan_3 = STGCAN(3, graph_args, num_class=None)   # bones
        self.fc = nn.Linear(256 * 3, num_class)
        self.register_buffer("parents", torch.tensor(layout_parents(graph_args["layout"]), dtype=torch.long))
        self.compute_dtype = None

    def forward(self, skel, sensor=None):
        dt = _compute_dtype(self)
        for m in (self.stgcan_1, self.stgcan_2, self.stgcan_3):
            m.compute_dtype = dt
        mot = skel[:, :2, 1:] - skel[:, :2, :-1]
        bone = bone_stream(skel, self.parents)
        x = torch.cat([self.stgcan_1.features(skel), self.stgcan_2.features(mot), self.stgcan_3.features(bone)], dim=-1)
        with torch.autocast("cuda", enabled=False):
            out = torch.addmm(self.fc.bias, x.float(), self.fc.weight.t())
        return out.to(dt) if dt == torch.bfloat16 else out
