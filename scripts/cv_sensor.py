"""BASELINE configs[4]: sensor-only BiLSTM on HAR30-shaped windows (128 steps x 6 channels), 10-fold cross-validation with the
folds as independent per-GPU jobs (fall_multimodal_b200.cv.run_cv), synthetic data. Each fold: train on 9/10 of the windows
(graph-captured TrainStep, RMSprop), batch-8192 inference on the held-out tenth, macro precision / recall / F1 / accuracy ->
precision_recall_f1.csv like Multimodal_Fall3/model/main_cross_validation.py:355-360."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

N_WIN, T, I, C, FOLDS = 81920, 128, 6, 6, 10


def make_data(device):
    """Windows whose class is decided by which accelerometer axis carries a slow oscillation (learnable, balanced)."""
    g = torch.Generator().manual_seed(1234)
    y = torch.randint(0, C, (N_WIN,), generator=g)
    x = torch.randn(N_WIN, T, I, generator=g) * 0.5
    t = torch.arange(T)[None, :, None].float()
    x += (torch.nn.functional.one_hot(y, I)[:, None, :].float() * torch.sin(t * 0.2 + torch.rand(N_WIN, 1, 1, generator=g) * 6.28))
    return x.to(device), y.to(device)


def fold_job(fold, device, epochs=2, batch=8192):
    import warnings
    from fall_multimodal_b200 import BiLSTM
    from fall_multimodal_b200.cv import macro_precision_recall_f1
    from fall_multimodal_b200.train import TrainStep
    torch.manual_seed(42 + fold)
    x, y = make_data(device)
    idx = torch.arange(N_WIN, device=device)
    test = (idx % FOLDS) == fold
    xtr, ytr, xte, yte = x[~test], y[~test], x[test], y[test]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = BiLSTM(I, 64, 1, 0.3, C, "mean").to(device).train()
    opt = torch.optim.RMSprop(m.parameters(), lr=1e-3, capturable=True)
    onehot = torch.nn.functional.one_hot(ytr, C).float()
    ts = TrainStep(m, opt, torch.nn.CrossEntropyLoss(), (None, xtr[:batch]), onehot[:batch], autocast_dtype=None, warmup=1)
    nb = xtr.shape[0] // batch
    for _ in range(epochs):
        perm = torch.randperm(xtr.shape[0], device=device)
        for b in range(nb):
            sel = perm[b * batch:(b + 1) * batch]
            ts.run((None, xtr[sel]), onehot[sel])
    m.eval()
    with torch.no_grad():
        pred = m(None, xte).argmax(1)
    return macro_precision_recall_f1(pred, yte, C)


if __name__ == "__main__":
    from fall_multimodal_b200.cv import run_cv
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--out", default="gpurun_out/precision_recall_f1.csv")
    a = ap.parse_args()
    t0 = time.perf_counter()
    rows = run_cv(fold_job, FOLDS, devices=[f"cuda:{i}" for i in range(a.gpus)], out_csv=a.out)
    dt = time.perf_counter() - t0
    print(json.dumps({"folds": FOLDS, "gpus": a.gpus, "wall_s": dt, "folds_per_min": FOLDS / dt * 60,
                      "accuracy_mean": sum(r["accuracy"] for r in rows) / FOLDS, "f1_mean": sum(r["f1"] for r in rows) / FOLDS}))
