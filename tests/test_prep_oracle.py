"""CPU: the input-preparation oracle against the function bodies of the unmodified reference script (when the reference
tree is mounted; skipped on the GPU box) and against hand-checkable properties."""
import ast
import os

import numpy as np
import pytest

from oracle import prep_oracle as PO

REF = "/root/reference/3_stream/har_create4_sensor.py"


def _reference_functions():
    tree = ast.parse(open(REF).read())
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
    return ns


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_scale_pose_matches_reference_function():
    ref = _reference_functions()["scale_pose"]
    rng = np.random.default_rng(0)
    xy = rng.normal(size=(40, 13, 2)) * 50 + 300
    want = ref(xy.copy())
    got = PO.scale_pose(xy)
    assert np.array_equal(want, got)
    assert np.allclose(got.min(1), -1) and np.allclose(got.max(1), 1)


def test_prepare_frames_and_windows_properties():
    rng = np.random.default_rng(1)
    L, J, C, S, T = 50, 13, 6, 15, 30
    xys = np.concatenate([rng.normal(size=(L, J, 2)) * 40 + 200, rng.uniform(0.2, 1.0, size=(L, J, 1))], -1)
    labels = np.eye(C)[rng.integers(0, C, L)] * 0.9 + 0.1 / (C - 1) * (1 - np.eye(C)[rng.integers(0, C, L)])
    sensors = rng.normal(size=(L, S))
    frames, scr, lbw = PO.prepare_frames(xys, labels)
    assert frames.shape == (L, J + 1, 3)
    assert np.allclose(frames[:, J], (frames[:, 1] + frames[:, 2]) / 2)
    assert (scr <= 1).all() and (scr > 0).all()
    skel, mot, sen, lab = PO.make_windows(frames, lbw, sensors, T)
    assert skel.shape == (L - T, 3, T, J + 1) and mot.shape == (L - T, 2, T - 1, J + 1)
    assert sen.shape == (L - T, T, S) and lab.shape == (L - T, C)
    assert np.array_equal(skel[3, :, 5, :], frames[8].astype(np.float32).T)
    assert np.allclose(lab[2], lbw[2:2 + T].mean(0), atol=1e-7)
    # nan handling of the Multimodal_Fall3 variant: a frame whose joints all coincide scales to 0/0 -> 0
    flat = xys.copy()
    flat[0, :, :2] = 5.0
    f2, _, _ = PO.prepare_frames(flat, labels, nan_to_num=True)
    assert np.array_equal(f2[0, :, :2], np.zeros((J + 1, 2)))
