"""CPU: the k-fold scheduler (folds as independent per-device jobs) and the device-side macro metrics."""
import csv

import numpy as np
import torch

from fall_multimodal_b200.cv import assign_folds, macro_precision_recall_f1, run_cv


def test_assign_folds_round_robin():
    a = assign_folds(10, 8)
    assert a[0] == [0, 8] and a[1] == [1, 9] and a[7] == [7]
    assert sorted(f for w in a for f in w) == list(range(10))
    assert assign_folds(3, 8)[:3] == [[0], [1], [2]]


def test_macro_metrics_match_sklearn():
    from sklearn.metrics import precision_recall_fscore_support
    rng = np.random.default_rng(0)
    for C in (2, 6, 11):
        y = rng.integers(0, C, 500)
        p = np.where(rng.random(500) < 0.7, y, rng.integers(0, C, 500))
        p[p == C - 1] = 0                                    # a class that is never predicted (0/0 precision -> 0)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = precision_recall_fscore_support(y, p, average="macro")
        got = macro_precision_recall_f1(torch.from_numpy(p), torch.from_numpy(y))
        assert abs(got["precision"] - want[0]) < 1e-12 and abs(got["recall"] - want[1]) < 1e-12 and abs(got["f1"] - want[2]) < 1e-12
        assert abs(got["accuracy"] - (y == p).mean()) < 1e-12


def _fake_fold(fold, device):
    # deterministic per-fold "metrics"; proves every fold ran exactly once on some worker
    torch.manual_seed(fold)
    y = torch.randint(0, 4, (200,), device=device)
    p = torch.where(torch.rand(200, device=device) < 0.5 + 0.04 * fold, y, torch.randint(0, 4, (200,), device=device))
    return macro_precision_recall_f1(p, y, 4)


def test_run_cv_two_cpu_workers(tmp_path):
    out = tmp_path / "precision_recall_f1.csv"
    rows = run_cv(_fake_fold, 5, devices=["cpu", "cpu"], out_csv=str(out))
    assert len(rows) == 5
    for f, r in enumerate(rows):
        assert r == {k: float(v) for k, v in _fake_fold(f, torch.device("cpu")).items()}
    table = list(csv.reader(open(out)))
    assert table[0] == ["", "precision", "recall", "f1", "accuracy"] and len(table) == 6 and table[3][0] == "2"


def test_assign_folds_properties():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 16))
    def prop(n_folds, n_workers):
        a = assign_folds(n_folds, n_workers)
        assert len(a) == n_workers
        assert sorted(f for w in a for f in w) == list(range(n_folds))               # every fold exactly once
        sizes = [len(w) for w in a]
        assert max(sizes) - min(sizes) <= 1                                           # balanced to within one fold

    prop()


def test_train_step_and_run_cv_refuse_to_run_without_cuda():
    import pytest
    from fall_multimodal_b200.train import TrainStep
    lin = torch.nn.Linear(4, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TrainStep(lin, torch.optim.SGD(lin.parameters(), lr=0.1), torch.nn.CrossEntropyLoss(), (torch.zeros(3, 4),), torch.zeros(3, 2))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA devices"):
            run_cv(_fake_fold, 2)


def _failing_fold(fold, device):
    if fold == 1:
        raise ValueError("fold 1 blew up")
    return {"precision": 1.0, "recall": 1.0, "f1": 1.0, "accuracy": 1.0}


def _dying_fold(fold, device):
    import os
    if fold == 0:
        os._exit(7)                 # dies without a traceback or a sentinel (what an OOM kill looks like)
    return {"precision": 1.0, "recall": 1.0, "f1": 1.0, "accuracy": 1.0}


def test_run_cv_reports_worker_exception_instead_of_hanging():
    import pytest
    with pytest.raises(RuntimeError, match="fold 1 blew up"):
        run_cv(_failing_fold, 4, devices=["cpu", "cpu"], poll_s=0.2)


def test_run_cv_detects_dead_worker():
    import pytest
    with pytest.raises(RuntimeError, match="exited with code 7"):
        run_cv(_dying_fold, 2, devices=["cpu", "cpu"], poll_s=0.2)
