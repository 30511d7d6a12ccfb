"""k-fold cross-validation as independent per-GPU jobs (BASELINE north star: "k-fold cross-validation folds run as
independent per-GPU jobs"; SURVEY.md 8(e): replicas only, no collective; 8(f) N4).

The reference (``Multimodal_Fall3/model/main_cross_validation.py:282-360``) trains the folds one after another on one device,
collects per fold ``precision_recall_fscore_support(labels, prediction, average='macro')`` plus the top-1 accuracy and writes
``precision_recall_f1.csv`` (columns precision, recall, f1, accuracy; one row per fold). Here the folds are dealt round-robin
to one worker process per device (10 folds on 8 GPUs = two waves); a worker owns its device, runs its folds back to back and
reports the same four numbers, computed on the device from a confusion matrix (one host sync per fold).
"""
from __future__ import annotations

import csv
import os
from typing import Callable, Sequence

import torch
import torch.multiprocessing as mp


def assign_folds(n_folds: int, n_workers: int) -> list[list[int]]:
    """Round-robin: worker w gets folds w, w + n_workers, ... (balanced to within one fold)."""
    return [list(range(w, n_folds, n_workers)) for w in range(max(1, n_workers))]


def macro_precision_recall_f1(prediction: torch.Tensor, labels: torch.Tensor, num_classes: int | None = None):
    """``sklearn.metrics.precision_recall_fscore_support(labels, prediction, average='macro')[:3]`` + accuracy, on the tensors'
    device. Classes are those present in labels or predictions (sklearn's default label set); 0/0 counts as 0."""
    prediction, labels = prediction.reshape(-1).long(), labels.reshape(-1).long()
    C = int(num_classes or (max(int(prediction.max()), int(labels.max())) + 1))
    cm = torch.bincount(labels * C + prediction, minlength=C * C).view(C, C).double()      # rows: truth, cols: prediction
    tp, pred_n, true_n = cm.diag(), cm.sum(0), cm.sum(1)
    present = (pred_n + true_n) > 0
    prec = torch.where(pred_n > 0, tp / pred_n.clamp_min(1), torch.zeros_like(tp))
    rec = torch.where(true_n > 0, tp / true_n.clamp_min(1), torch.zeros_like(tp))
    f1 = torch.where(prec + rec > 0, 2 * prec * rec / (prec + rec).clamp_min(1e-300), torch.zeros_like(tp))
    n = present.sum().clamp_min(1)
    out = torch.stack([prec[present].sum() / n, rec[present].sum() / n, f1[present].sum() / n, tp.sum() / cm.sum().clamp_min(1)])
    p, r, f, acc = out.tolist()
    return {"precision": p, "recall": r, "f1": f, "accuracy": acc}


def _worker(rank: int, devices: Sequence[str], folds: list[list[int]], fold_fn: Callable, queue):
    try:
        dev = torch.device(devices[rank])
        if dev.type == "cuda":
            torch.cuda.set_device(dev)
        for fold in folds[rank]:
            res = fold_fn(fold, dev)
            queue.put(("fold", fold, {k: float(v) for k, v in res.items()}))
        queue.put(("done", rank, None))
    except BaseException:       # report instead of dying silently: the parent would otherwise wait for the sentinel forever
        import traceback

        queue.put(("error", rank, traceback.format_exc()))
        raise


def run_cv(fold_fn: Callable, n_folds: int, devices: Sequence[str] | None = None, out_csv: str | None = None,
           poll_s: float = 5.0):
    """Run ``fold_fn(fold_index, device) -> {"precision","recall","f1","accuracy"}`` for every fold, folds dealt round-robin
    to one process per device. Returns the per-fold dicts in fold order and (optionally) writes the reference's CSV."""
    if devices is None:
        if not torch.cuda.is_available():
            raise RuntimeError("run_cv needs CUDA devices (pass devices=[...] explicitly to override)")
        devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    devices = list(devices)[:max(1, min(len(devices), n_folds))]
    folds = assign_folds(n_folds, len(devices))
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, devices, folds, fold_fn, queue)) for r in range(len(devices))]
    for p in procs:
        p.start()
    import queue as _queue

    results, done, failure = {}, 0, None
    while done < len(procs) and failure is None:
        try:
            kind, a, b = queue.get(timeout=poll_s)
        except _queue.Empty:
            # no message: a worker that died without reporting (OOM kill, CUDA abort, segfault) never sends its sentinel
            dead = [r for r, p in enumerate(procs) if not p.is_alive() and p.exitcode not in (0, None)]
            if dead:
                failure = f"cross-validation worker {dead[0]} exited with code {procs[dead[0]].exitcode} without a result"
            continue
        if kind == "done":
            done += 1
        elif kind == "error":
            failure = f"cross-validation worker {a} failed:\n{b}"
        else:
            results[a] = b
    if failure is not None:
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join()
        raise RuntimeError(failure)
    for p in procs:
        p.join()
        if p.exitcode != 0:
            raise RuntimeError(f"a cross-validation worker exited with code {p.exitcode}")
    rows = [results[f] for f in range(n_folds)]
    if out_csv:
        os.makedirs(os.path.dirname(os.path.abspath(out_csv)), exist_ok=True)
        with open(out_csv, "w", newline="") as fh:      # same layout as pd.DataFrame(...).to_csv: index column + 4 metrics
            w = csv.writer(fh)
            w.writerow(["", "precision", "recall", "f1", "accuracy"])
            for i, r in enumerate(rows):
                w.writerow([i, r["precision"], r["recall"], r["f1"], r["accuracy"]])
    return rows
