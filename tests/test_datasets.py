"""CPU: the reference's window pickles and video-grouped k-fold split, resident batch iteration."""
import pickle

import numpy as np
import torch

from fall_multimodal_b200.datasets import ResidentSplit, build_cv_splits, load_window_pickles, video_kfold


def _write_pickles(tmp_path, n_files=3, per_video=7, videos_per_file=8, T=30, V=14, S=15, C=11):
    rng = np.random.default_rng(0)
    paths = []
    for f in range(n_files):
        vids, feats, sens, labs = [], [], [], []
        for v in range(videos_per_file):
            for _ in range(per_video):
                vids.append(f"Subject{f}Activity{v}")
                feats.append(rng.normal(size=(T, V, 3)))
                sens.append(rng.normal(size=(T, S)))
                labs.append(np.asarray(rng.dirichlet(np.ones(C)), dtype=object))
        p = tmp_path / f"har30_{f}_sensor_new-set(labelXscrw).pkl"
        with open(p, "wb") as fh:       # same tuple layout as har_create4_sensor.py:146-147
            pickle.dump((vids, np.stack(feats), np.stack(sens), np.stack(labs)), fh)
        paths.append(str(p))
    return paths


def test_pickles_and_video_kfold_match_reference_logic(tmp_path):
    from sklearn.model_selection import KFold
    paths = _write_pickles(tmp_path)
    videos, features, sensors, labels = load_window_pickles(paths)
    assert features.shape == (3 * 8 * 7, 30, 14, 3) and features.dtype == np.float32 and labels.dtype == np.float32
    folds = video_kfold(videos, n_splits=10, seed=42)
    assert len(folds) == 10
    names = np.unique(videos)
    ref = list(KFold(n_splits=10, shuffle=True, random_state=42).split(names))        # cv_dataloader.py:155
    for (tr, te), (rtr, rte) in zip(folds, ref):
        # the reference's per-window membership test: `if video in unique_video_names[train_idx]`
        want_tr = [i for i, v in enumerate(videos) if v in names[rtr]]
        assert list(tr) == want_tr and len(tr) + len(te) == len(videos)
        assert not set(np.asarray(videos)[tr]) & set(np.asarray(videos)[te])          # a video never straddles the split
    assert sorted(i for _, te in folds for i in te) == list(range(len(videos)))       # every window is tested exactly once


def test_resident_split_iteration(tmp_path):
    paths = _write_pickles(tmp_path, n_files=1)
    videos, features, sensors, labels = load_window_pickles(paths)
    idx = np.arange(50)
    sp = ResidentSplit(features, sensors, labels, idx, "cpu", batch_size=16, shuffle=True, drop_last=True, seed=1)
    assert len(sp) == 3 and sp.skel.shape == (50, 3, 30, 14)
    assert torch.equal(sp.skel[5], torch.as_tensor(features[5]).permute(2, 0, 1))
    seen = torch.cat([lab for _, _, lab in sp])
    assert seen.shape == (48, 11)
    first = [b[2] for b in sp][0]
    assert not torch.equal(first, seen[:16])                     # reshuffled on the next epoch
    ev = ResidentSplit(features, sensors, labels, idx, "cpu", batch_size=16, shuffle=False, drop_last=False)
    assert len(ev) == 4 and torch.equal(torch.cat([s for s, _, _ in ev]), ev.skel)
    folds, C = build_cv_splits(paths, "cpu", batch_size=8, n_splits=4)
    assert C == 11 and len(folds) == 4 and folds[0]["test"] is folds[0]["valid"]
    assert folds[0]["train"].num_samples + folds[0]["valid"].num_samples == len(videos)


def test_resident_split_is_lazy_and_releasable():
    """Folds share the host arrays; a split copies to its device on first use only and release() drops the copy."""
    import numpy as np
    rng = np.random.default_rng(0)
    feats, sens = rng.standard_normal((12, 5, 14, 3)).astype("float32"), rng.standard_normal((12, 5, 6)).astype("float32")
    labels = np.eye(3, dtype="float32")[rng.integers(0, 3, 12)]
    sp = ResidentSplit(feats, sens, labels, [1, 3, 5, 7, 9], "cpu", batch_size=2, shuffle=False, drop_last=False)
    assert sp._dev is None and sp.num_samples == 5 and len(sp) == 3          # nothing copied yet
    batches = list(sp)
    assert sp._dev is not None and batches[0][0].shape == (2, 3, 5, 14) and batches[-1][0].shape[0] == 1
    assert np.allclose(batches[0][1].numpy(), sens[[1, 3]])
    sp.release()
    assert sp._dev is None
    assert sp.skel.shape == (5, 3, 5, 14) and sp._dev is not None            # rebuilt on demand
