"""On-device input pipeline: a recording resident in HBM -> training batches in the model layouts.

Reference: ``3_stream/har_create4_sensor.py:36-47,113-132`` (pose scaling, centre joint, score-weighted targets, sliding
windows), ``Multimodal_Fall3/dataset.py:28-41`` (the same scaling with ``nan_to_num``), ``F2/dataset.py:27`` (the
(T,V,C)->(C,T,V) permute), ``F2/Model/combination.py:39`` (motion stream). The reference does this per sample in numpy on
the host; here the frames are normalised once (``csrc/prep.cu``) and every batch is one gather kernel."""
from __future__ import annotations

import torch

from . import _lib as L

MAIN_IDX_PARTS = (1, 2, 7, 8, -1)     # har_create4_sensor.py:13 (-1 = the appended centre joint)


class WindowPipeline:
    def __init__(self, xys: torch.Tensor, labels: torch.Tensor | None = None, sensors: torch.Tensor | None = None,
                 n_frames: int = 30, main_idx_parts=MAIN_IDX_PARTS, nan_to_num: bool = False):
        """xys (L, J, 3) raw (x, y, score) per frame and joint; labels (L, C) (soft) targets; sensors (L, S)."""
        if not xys.is_cuda:
            raise RuntimeError("WindowPipeline needs CUDA tensors (the product path has no CPU fallback)")
        self.n_frames = n_frames
        xys = xys.to(torch.float64).contiguous()
        Lf, J, _ = xys.shape
        dev = xys.device
        self.frames = torch.empty(Lf, J + 1, 3, dtype=torch.float32, device=dev)
        self.scr = torch.empty(Lf, dtype=torch.float32, device=dev)
        lab64 = labels.to(dev, torch.float64).contiguous() if labels is not None else None
        self.lbw = torch.empty(Lf, lab64.shape[1], dtype=torch.float32, device=dev) if lab64 is not None else None
        self.sensors = sensors.to(dev, torch.float32).contiguous() if sensors is not None else None
        mask = 0
        for j in main_idx_parts:
            if j >= 0:
                mask |= 1 << j
        L.check(L.load().fmm_prep_frames(xys.data_ptr(), L.ptr(lab64), self.frames.data_ptr(), self.scr.data_ptr(), L.ptr(self.lbw),
                                         Lf, J, lab64.shape[1] if lab64 is not None else 0, mask, int(-1 in main_idx_parts),
                                         int(nan_to_num), L.stream()), "prep_frames")

    def __len__(self):
        return self.frames.shape[0] - self.n_frames       # range(xys.shape[0] - n_frames), har_create4_sensor.py:126

    def batch(self, starts: torch.Tensor, motion: bool = True):
        """Windows starting at frames ``starts`` (int tensor) -> (skel (N,3,T,V), mot (N,2,T-1,V) | None, sensor | None, label | None)."""
        dev = self.frames.device
        starts = starts.to(dev, torch.int32).contiguous()
        N, T, Vc = starts.numel(), self.n_frames, self.frames.shape[1]
        skel = torch.empty(N, 3, T, Vc, dtype=torch.float32, device=dev)
        mot = torch.empty(N, 2, T - 1, Vc, dtype=torch.float32, device=dev) if motion else None
        sen = torch.empty(N, T, self.sensors.shape[1], dtype=torch.float32, device=dev) if self.sensors is not None else None
        lab = torch.empty(N, self.lbw.shape[1], dtype=torch.float32, device=dev) if self.lbw is not None else None
        L.check(L.load().fmm_prep_windows(self.frames.data_ptr(), L.ptr(self.lbw), L.ptr(self.sensors), starts.data_ptr(),
                                          skel.data_ptr(), L.ptr(mot), L.ptr(sen), L.ptr(lab), N, T, Vc,
                                          self.lbw.shape[1] if self.lbw is not None else 0,
                                          self.sensors.shape[1] if self.sensors is not None else 0, L.stream()), "prep_windows")
        return skel, mot, sen, lab
