"""GPU: the on-device input pipeline (csrc/prep.cu) against the numpy oracle of the reference preprocessing."""
import numpy as np
import pytest
import torch

from oracle import prep_oracle as PO

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("nan_to_num", [False, True])
def test_window_pipeline_matches_oracle(nan_to_num):
    from fall_multimodal_b200.inputs import WindowPipeline
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    L, J, C, S, T = 200, 13, 6, 15, 30
    xys = np.concatenate([rng.normal(size=(L, J, 2)) * 40 + 200, rng.uniform(0.2, 1.0, size=(L, J, 1))], -1)
    if nan_to_num:
        xys[7, :, :2] = 3.0          # degenerate frame: 0/0 -> nan -> 0 in the Multimodal_Fall3 variant
    labels = rng.uniform(size=(L, C))
    sensors = rng.normal(size=(L, S))
    frames, scr, lbw = PO.prepare_frames(xys, labels, nan_to_num=nan_to_num)
    pipe = WindowPipeline(torch.from_numpy(xys).to(dev), torch.from_numpy(labels), torch.from_numpy(sensors), n_frames=T,
                          nan_to_num=nan_to_num)
    assert len(pipe) == L - T
    assert np.allclose(pipe.frames.cpu().numpy(), frames.astype(np.float32), atol=1e-6, equal_nan=True)
    assert np.allclose(pipe.scr.cpu().numpy(), scr.astype(np.float32), atol=1e-6)
    starts = [0, 5, 17, L - T - 1, 100, 3]
    skel, mot, sen, lab = PO.make_windows(frames, lbw, sensors, T, starts=starts)
    g_skel, g_mot, g_sen, g_lab = pipe.batch(torch.tensor(starts))
    assert np.allclose(g_skel.cpu().numpy(), skel, atol=1e-6, equal_nan=True)
    assert np.allclose(g_mot.cpu().numpy(), mot, atol=2e-6, equal_nan=True)
    assert np.array_equal(g_sen.cpu().numpy(), sen)
    assert np.allclose(g_lab.cpu().numpy(), lab, atol=1e-6)


def test_window_pipeline_refuses_cpu():
    from fall_multimodal_b200.inputs import WindowPipeline
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        WindowPipeline(torch.zeros(40, 13, 3))
