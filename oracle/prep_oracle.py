"""ORACLE (test infrastructure): numpy restatement of the reference's input preparation, the step in front of the
model.

  scale_pose            3_stream/har_create4_sensor.py:36-47 (and Multimodal_Fall3/dataset.py:28-41, which adds nan_to_num)
  centre point          har_create4_sensor.py:113   (mean of joints 1 and 2 appended as joint J)
  score weighting       har_create4_sensor.py:115-119 (main parts x1.5 clipped at 1, mean over joints)
  targets x score       har_create4_sensor.py:121-124
  sliding windows       har_create4_sensor.py:126-132 (n_frames consecutive frames, label = window mean)
  (T,V,C)->(C,T,V)      Fall_2_Spatial_Temporal_SR/dataset.py:27
  motion stream         Fall_2_Spatial_Temporal_SR/Model/combination.py:39

``scale_pose`` / ``seq_label_smoothing`` are pinned against the function bodies of the unmodified script (extracted with
``ast``, the script itself is top-level file IO) in tests/test_prep_oracle.py; the windowing loop is restated from the lines
cited above. Only tests/ may import this file.
"""
from __future__ import annotations

import numpy as np

MAIN_IDX_PARTS = [1, 2, 7, 8, -1]      # har_create4_sensor.py:13


def scale_pose(xy, nan_to_num=False):
    xy = np.array(xy, dtype=np.float64, copy=True)
    if xy.ndim == 2:
        xy = np.expand_dims(xy, 0)
    xy_min = np.nanmin(xy, axis=1)
    xy_max = np.nanmax(xy, axis=1)
    for i in range(xy.shape[0]):
        xy[i] = ((xy[i] - xy_min[i]) / (xy_max[i] - xy_min[i])) * 2 - 1
        if nan_to_num:
            xy[i] = np.nan_to_num(xy[i], copy=True, nan=0.0, posinf=0.0, neginf=0.0)
    return xy


def prepare_frames(xys, labels, nan_to_num=False, main_idx=MAIN_IDX_PARTS):
    """xys (L,J,3) raw (x, y, score); labels (L,C) -> frames (L,J+1,3), per-frame score (L,), weighted labels (L,C)."""
    xys = np.array(xys, dtype=np.float64, copy=True)
    xys[:, :, :2] = scale_pose(xys[:, :, :2], nan_to_num)
    xys = np.concatenate((xys, np.expand_dims((xys[:, 1, :] + xys[:, 2, :]) / 2, 1)), axis=1)
    scr = xys[:, :, -1].copy()
    scr[:, main_idx] = np.minimum(scr[:, main_idx] * 1.5, 1.0)
    scr = scr.mean(1)
    return xys, scr, np.asarray(labels, dtype=np.float64) * scr[:, None]


def make_windows(frames, lbw, sensors, n_frames, starts=None):
    """Sliding windows (stride 1 unless ``starts`` is given) -> model inputs:
    skel (N,3,T,V) float32, motion (N,2,T-1,V), sensor (N,T,S), label (N,C)."""
    L = frames.shape[0]
    starts = range(L - n_frames) if starts is None else starts
    feat = np.stack([frames[i:i + n_frames] for i in starts])               # (N,T,V,3)
    lab = np.stack([lbw[i:i + n_frames].mean(0) for i in starts])
    sen = np.stack([sensors[i:i + n_frames] for i in starts])
    skel = np.transpose(feat.astype(np.float32), (0, 3, 1, 2))              # dataset.py:27 after the float32 cast (:19)
    mot = skel[:, :2, 1:] - skel[:, :2, :-1]                                 # combination.py:39
    return skel, mot, sen.astype(np.float32), lab.astype(np.float32)
