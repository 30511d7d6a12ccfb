"""The reference's dataset files and fold construction, feeding resident device tensors.

``3_stream/har_create4_sensor.py:146-147`` pickles ``(video_name_set, feature_set, sensor_set, labels_set)``: windows
``(N, T, V, 3)`` float64, accelerometer windows ``(N, T, S)``, soft labels ``(N, C)``. ``F2/cv_dataloader.py:137-167`` loads
a list of such files, concatenates them and builds 10 folds with ``KFold(n_splits=10, shuffle=True, random_state=seed)``
over the UNIQUE VIDEO NAMES (all windows of a video stay on one side). The reference then wraps per-sample tuples in a
``DataLoader`` (8 workers, pinned memory); here a fold's tensors are moved to the GPU once, permuted to the model layout
(``F2/dataset.py:27``), and batches are index gathers on the device (shuffle / drop_last like the reference's loaders).
"""
from __future__ import annotations

import pickle
from typing import Iterator, Sequence

import numpy as np
import torch


def load_window_pickles(paths: Sequence[str]):
    """-> (videos: list[str], features (N,T,V,3) float32, sensors (N,T,S) float32, labels (N,C) float32)."""
    videos, feats, sens, labs = [], [], [], []
    for p in paths:
        with open(p, "rb") as fh:
            vid, fts, sr, lbs = pickle.load(fh)
        videos += list(vid)
        feats.append(np.asarray(fts))
        sens.append(np.asarray(sr))
        labs.append(np.asarray(lbs))
    features = np.concatenate(feats, axis=0).astype(np.float32)
    sensors = np.concatenate(sens, axis=0).astype(np.float32)
    labels = np.concatenate(labs, axis=0).astype(np.float32)          # dtype object -> float32 (cv_dataloader.py:149)
    assert len(videos) == features.shape[0] == sensors.shape[0] == labels.shape[0]
    return videos, features, sensors, labels


def video_kfold(videos: Sequence[str], n_splits: int = 10, seed: int = 42):
    """Window index pairs (train_idx, test_idx) per fold; the split is over unique video names (cv_dataloader.py:152-163)."""
    from sklearn.model_selection import KFold
    names = np.unique(videos)
    vid = np.asarray(videos)
    folds = []
    for tr, te in KFold(n_splits=n_splits, shuffle=True, random_state=seed).split(names):
        in_train = np.isin(vid, names[tr])
        folds.append((np.nonzero(in_train)[0], np.nonzero(~in_train)[0]))
    return folds


class ResidentSplit:
    """One side of a fold, resident on ``device`` in the model layouts: skel (N,3,T,V), sensor (N,T,S), label (N,C).

    The device copies are made on first use and dropped by ``release()``: the list ``build_cv_splits`` returns holds every fold,
    but only the folds a worker is actually iterating occupy device memory (a cv worker uses one fold at a time)."""

    def __init__(self, features, sensors, labels, index, device, batch_size: int, shuffle: bool, drop_last: bool, seed: int = 42):
        self._src = (features, sensors, labels)
        self._idx = torch.as_tensor(np.asarray(index), dtype=torch.long)
        self.device = device
        self._dev = None
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self._gen = torch.Generator(device="cpu").manual_seed(seed)

    def _resident(self):
        if self._dev is None:
            features, sensors, labels = self._src
            idx = self._idx
            self._dev = (torch.as_tensor(features)[idx].permute(0, 3, 1, 2).contiguous().to(self.device),     # F2/dataset.py:27
                         torch.as_tensor(sensors)[idx].contiguous().to(self.device),
                         torch.as_tensor(labels)[idx].contiguous().to(self.device))
        return self._dev

    def release(self):
        """Drop the device copies (they are rebuilt on the next use)."""
        self._dev = None

    @property
    def skel(self):
        return self._resident()[0]

    @property
    def sensor(self):
        return self._resident()[1]

    @property
    def label(self):
        return self._resident()[2]

    def __len__(self):
        n = self.num_samples
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    @property
    def num_samples(self):
        return int(self._idx.numel())

    def __iter__(self) -> Iterator:
        skel, sensor, label = self._resident()
        n = skel.shape[0]
        order = torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)
        order = order.to(skel.device)
        for b in range(len(self)):
            sel = order[b * self.batch_size:(b + 1) * self.batch_size]
            yield skel[sel], sensor[sel], label[sel]


def build_cv_splits(paths: Sequence[str], device, batch_size: int = 16, n_splits: int = 10, seed: int = 42):
    """The reference's ``cv_dataloaders`` list (``{"train", "valid", "test"}`` per fold; valid == test, cv_dataloader.py:165-169)
    with resident splits instead of DataLoaders. Returns (folds, num_classes)."""
    videos, features, sensors, labels = load_window_pickles(paths)
    out = []
    for tr, te in video_kfold(videos, n_splits, seed):
        train = ResidentSplit(features, sensors, labels, tr, device, batch_size, shuffle=True, drop_last=True, seed=seed)
        valid = ResidentSplit(features, sensors, labels, te, device, batch_size, shuffle=False, drop_last=False, seed=seed)
        out.append({"train": train, "valid": valid, "test": valid})
    return out, labels.shape[-1]
