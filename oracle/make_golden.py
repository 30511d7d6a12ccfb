"""ORACLE tooling — generates tests/golden/*.pt by RUNNING THE REFERENCE in the build container.

    python oracle/make_golden.py

Every fixture holds: the config, the reference state_dict key->shape list, the recipe seeds, and
the reference outputs (logits, loss, gradient tensors or their summaries, BN running-stat updates).
Weights and inputs are NOT stored: both sides regenerate them with
``stgcn_oracle.fill_state_dict`` / ``synthetic_batch`` (order-independent, seeded per key).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, stgcn_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FULL_LIMIT = 2048


def summarize(t: torch.Tensor, key: str):
    t = t.detach().double().flatten()
    if t.numel() <= FULL_LIMIT:
        return {"full": t.float().clone()}
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) % (2 ** 31))
    idx = torch.randint(0, t.numel(), (64,), generator=g)
    return {"sum": t.sum().item(), "l1": t.abs().sum().item(), "l2": t.norm().item(),
            "amax": t.abs().max().item(), "idx": idx, "vals": t[idx].float().clone()}


def fill_module(mod, seed):
    sd = mod.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    filled = O.fill_state_dict(shapes, seed)
    for k, v in filled.items():
        sd[k] = v
    mod.load_state_dict(sd)
    return shapes


def run_train_step(mod, call, target):
    mod.train()
    mod.zero_grad()
    before = {k: v.clone() for k, v in mod.state_dict().items() if "running_" in k}
    logits = call()
    loss = torch.nn.CrossEntropyLoss()(logits, target) if target is not None else logits.square().mean()
    loss.backward()
    grads = {k: summarize(p.grad, k) for k, p in mod.named_parameters() if p.grad is not None}
    running = {k: summarize(v, k) for k, v in mod.state_dict().items() if "running_" in k}
    mod.eval()
    with torch.no_grad():
        # eval with the ORIGINAL running stats (the train step above updated them in place)
        sd = mod.state_dict()
        after = {k: sd[k].clone() for k in before}
        for k, v in before.items():
            sd[k].copy_(v)
        ev = call()
        for k, v in after.items():
            sd[k].copy_(v)
    return {"logits": logits.detach().clone(), "loss": float(loss), "grads": grads, "running": running,
            "eval_logits": ev.detach().clone()}


def main():
    assert ref_import.available(), "reference tree not mounted"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    stg, g, bl, comb = ref_import.load_gstcan()
    n33, e33, c33 = O.LAYOUTS["mediapipe33"]
    ref_import.install_layout(stg, g, "mediapipe33", n33, e33, c33)
    comb.STGCAN = stg.STGCAN

    # ---- adjacency matrices ---------------------------------------------------------------
    adj = {}
    for layout in ("coco_cut", "coco_mmpose", "mediapipe33"):
        for strategy in ("uniform", "distance", "spatial"):
            adj[f"{layout}/{strategy}"] = torch.tensor(stg.Graph(layout=layout, strategy=strategy).A)
    torch.save(adj, os.path.join(OUT, "graph_A.pt"))

    # ---- parameter-count known answers ----------------------------------
    counts = {}
    m = comb.TwoStreamSTGCAN_BiLSTM(3, {"layout": "coco_cut", "strategy": "spatial"}, 11, 15)
    counts["TwoStreamSTGCAN_BiLSTM"] = sum(p.numel() for p in m.parameters())
    m = bl.BiLSTM(15, 64, 1, 0.3, 11, "mean")
    counts["BiLSTM"] = sum(p.numel() for p in m.parameters())
    m = stg.STGCAN(3, {"layout": "coco_cut", "strategy": "spatial"}, 11)
    counts["STGCAN"] = sum(p.numel() for p in m.parameters())
    counts["STGCAN_keys"] = len(m.state_dict())
    nbs = ref_import.load_notebook_sensor()
    counts["CNN1D"] = sum(p.numel() for p in nbs["CNN1D"]().parameters())
    assert counts["TwoStreamSTGCAN_BiLSTM"] == 4298291 and counts["BiLSTM"] == 47387, counts
    assert counts["CNN1D"] == 11104, counts  # 3904 conv+bn params + the unused fc(224->32)
    torch.save(counts, os.path.join(OUT, "param_counts.pt"))

    # ---- GSTCAN cases -----------------------------------------------------------------------
    cases = [
        ("stgcan_coco_spatial", dict(in_ch=3, layout="coco_cut", strategy="spatial", num_class=11, N=8, T=12)),
        ("stgcan_mp33_spatial", dict(in_ch=3, layout="mediapipe33", strategy="spatial", num_class=11, N=6, T=16)),
        ("stgcan_mmpose_uniform_feat", dict(in_ch=2, layout="coco_mmpose", strategy="uniform", num_class=None, N=5, T=9)),
    ]
    for name, c in cases:
        mod = stg.STGCAN(c["in_ch"], {"layout": c["layout"], "strategy": c["strategy"]}, num_class=c["num_class"])
        shapes = fill_module(mod, seed=1)
        V = mod.A.shape[1]
        skel, _, target, _ = O.synthetic_batch(c["N"], c["T"], V, 11, seed=7)
        skel = skel[:, : c["in_ch"]].contiguous()
        res = run_train_step(mod, lambda: mod(skel, None), target if c["num_class"] else None)
        torch.save({"config": c, "shapes": shapes, "fill_seed": 1, "batch_seed": 7, **res},
                   os.path.join(OUT, name + ".pt"))
        print(name, "loss", res["loss"])

    # ---- BiLSTM ---------------------------------------------------------------------------------
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = bl.BiLSTM(15, 64, 1, 0.3, 11, "mean")
    shapes = fill_module(mod, seed=2)
    _, sensor, target, _ = O.synthetic_batch(6, 4, 14, 11, sensor_len=30, sensor_ch=15, seed=8)
    res = run_train_step(mod, lambda: mod(None, sensor), target)
    torch.save({"config": dict(I=15, H=64, num_class=11, N=6, L=30), "shapes": shapes, "fill_seed": 2,
                "batch_seed": 8, **res}, os.path.join(OUT, "bilstm_mean.pt"))
    print("bilstm loss", res["loss"])

    # ---- CNN1D (notebook) -----------------------------------------------------------------------
    mod = nbs["CNN1D"]()
    shapes = fill_module(mod, seed=3)
    _, sensor, _, _ = O.synthetic_batch(5, 4, 14, 11, sensor_len=30, sensor_ch=15, seed=9)
    xin = sensor.permute(0, 2, 1).contiguous()
    res = run_train_step(mod, lambda: mod(xin), None)
    torch.save({"config": dict(Cin=15, L=30, N=5), "shapes": shapes, "fill_seed": 3, "batch_seed": 9, **res},
               os.path.join(OUT, "cnn1d.pt"))
    print("cnn1d loss", res["loss"])

    # ---- fusion: TwoStreamSTGCAN_BiLSTM ----------------------------------------------------------
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = comb.TwoStreamSTGCAN_BiLSTM(3, {"layout": "coco_cut", "strategy": "spatial"}, 11, 15)
    shapes = fill_module(mod, seed=4)
    skel, sensor, target, _ = O.synthetic_batch(6, 10, 14, 11, sensor_len=30, sensor_ch=15, seed=10)
    res = run_train_step(mod, lambda: mod(skel, sensor), target)
    torch.save({"config": dict(layout="coco_cut", strategy="spatial", num_class=11, N=6, T=10, L=30, I=15),
                "shapes": shapes, "fill_seed": 4, "batch_seed": 10, **res},
               os.path.join(OUT, "two_stream_bilstm.pt"))
    print("fusion loss", res["loss"])


def main_tragcn():
    """TRAGCN family fixtures from the unmodified root-level reference files."""
    import warnings
    from oracle import tragcn_oracle as TO
    torch.set_num_threads(4)
    cases = [
        ("targcn_v25_t12", dict(V=25, T=12, B=4, adj=None, fill_seed=5, batch_seed=11)),
        ("targcn_v14_t30_adj", dict(V=14, T=30, B=3, adj="rand", fill_seed=6, batch_seed=12)),
        # BASELINE config 4's own clip shape (T=300, V=25; SURVEY D5: seq_len re-pointed), a bounded batch
        ("targcn_v25_t300", dict(V=25, T=300, B=4, adj=None, fill_seed=7, batch_seed=13)),
    ]
    only = os.environ.get("FMM_GOLDEN_ONLY")
    if only:
        cases = [c for c in cases if c[0] in only.split(",")]
    for name, c in cases:
        adj = None
        if c["adj"] == "rand":
            g = torch.Generator().manual_seed(77)
            a = (torch.rand(c["V"], c["V"], generator=g) < 0.3).float() * torch.rand(c["V"], c["V"], generator=g)
            adj = ((a + a.t()) * 0.5).contiguous()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            M = ref_import.load_tragcn(c["T"])
            mod = M.TARGCN(num_nodes=c["V"], adj=adj)
        sd = mod.state_dict()
        shapes = {k: tuple(v.shape) for k, v in sd.items()}
        assert shapes == TO.targcn_param_shapes(V=c["V"], T=c["T"]), "oracle shape table drifted from the reference"
        filled = TO.fill_targcn(shapes, c["fill_seed"])
        assert torch.allclose(filled["encoder.trans_layer_T.PE.pe"], sd["encoder.trans_layer_T.PE.pe"].cpu())
        mod.load_state_dict(filled)
        x, tgt = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=c["batch_seed"])
        mod.train()
        mod.zero_grad()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            logits = mod(x)
        loss = torch.nn.CrossEntropyLoss()(logits, tgt)
        loss.backward()
        grads = {k: summarize(p.grad, k) for k, p in mod.named_parameters() if p.grad is not None}
        torch.save({"config": {k: v for k, v in c.items() if k != "adj"}, "adj": adj, "shapes": shapes,
                    "n_params": sum(p.numel() for p in mod.parameters()),
                    "logits": logits.detach().clone(), "loss": float(loss), "grads": grads},
                   os.path.join(OUT, name + ".pt"))
        print(name, "loss", float(loss), "params", sum(p.numel() for p in mod.parameters()))

    if only:
        return
    # the reference's own bf16 path (MF3/main.py:97 torch.amp.autocast) on CPU: the yardstick for the bf16 gate
    c = dict(V=25, T=16, B=8, fill_seed=2, batch_seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = ref_import.load_tragcn(c["T"])
        mod = M.TARGCN(num_nodes=c["V"], adj=None)
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    mod.load_state_dict(TO.fill_targcn(shapes, c["fill_seed"]))
    x, tgt = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=c["batch_seed"])
    mod.train()
    mod.zero_grad()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits = mod(x)
            loss = torch.nn.CrossEntropyLoss()(logits.float(), tgt)
    loss.backward()
    grads = {k: summarize(p.grad, k) for k, p in mod.named_parameters() if p.grad is not None}
    torch.save({"config": c, "shapes": shapes, "logits": logits.detach().float().clone(), "loss": float(loss.detach()),
                "grads": grads}, os.path.join(OUT, "targcn_v25_t16_autocast.pt"))
    print("targcn_v25_t16_autocast loss", float(loss.detach()))


def main_musa(cls_name="Model", out_name="musa_coco_uniform.pt"):
    """musa Model (and Ablation, musa_model.py:593-686) of Multimodal_Fall3/main.py:307-320: train step with
    DropBlock / Dropout switched off (keep_prob = 1, p = 0: the random masks are the only thing that cannot be pinned) + eval logits."""
    import warnings
    from oracle import musa_oracle as MO
    torch.set_num_threads(4)
    mm = ref_import.load_musa()
    c = dict(N=8, T=30, V=14, fill_seed=3, batch_seed=13)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = getattr(mm, cls_name)(num_class=11, num_point=14, max_frame=300, graph=mm.adjGraph(layout="coco_cut", strategy="uniform"),
                                    bias=True, edge=True, block_size=41, embed_dim=64, n_stage=1, act_type="tanh")
    sd = mod.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    sd.update(MO.fill_musa(shapes, c["fill_seed"]))
    mod.load_state_dict(sd)
    for m in mod.modules():
        if hasattr(m, "keep_prob"):
            m.keep_prob = 1
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    skel, _, target, _ = O.synthetic_batch(c["N"], c["T"], c["V"], 11, seed=c["batch_seed"])
    # fp32 run of the reference (what a user sees) and an fp64 run of the same module (the truth both fp32 implementations
    # are measured against: the first-layer gradients of this net are cancellation-heavy, the reference's own fp32 run is
    # ~1e-3 away from it)
    res32 = run_train_step(mod, lambda: mod(skel), target)
    mod.load_state_dict(sd)
    mod64 = mod.double()
    skel64, target64 = skel.double(), target.double()
    res = run_train_step(mod64, lambda: mod64(skel64), target64)
    g32 = {k: v for k, v in res32["grads"].items()}
    ref32_err = {}
    for k, summ in res["grads"].items():
        a, b = summ, g32[k]
        if "full" in a:
            ref32_err[k] = float((a["full"].double() - b["full"].double()).abs().max())
        else:
            ref32_err[k] = float((a["vals"].double() - b["vals"].double()).abs().max())
    torch.save({"config": c, "shapes": shapes, "A": sd["stream_pos.0.A"].clone(), "n_params": sum(p.numel() for p in mod.parameters()),
                "ref32_logits": res32["logits"], "ref32_grad_abs_err": ref32_err, **res}, os.path.join(OUT, out_name))
    print("musa", cls_name, "loss", res["loss"])


def main_notebook():
    """The notebook fusion model (GSTCAN_HAR_conv_10kfold.ipynb#cell1:L362-416 + #cell2): tuple input, BiLSTM sensor logits in
    the concat, softmax output fed to CrossEntropyLoss. Classes exec()ed from the notebook cells, unmodified."""
    import json
    import warnings
    torch.set_num_threads(4)
    nb = json.load(open(os.path.join(ref_import.REF, "GSTCAN_HAR_conv_10kfold.ipynb")))
    ns = {"device": torch.device("cpu")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec("".join(nb["cells"][2]["source"]), ns)
        exec("".join(nb["cells"][1]["source"]), ns)
        mod = ns["TwoStreamSpatialTemporalGraph"]({"layout": "coco_cut", "strategy": "spatial"}, 11)
    shapes = fill_module(mod, seed=8)
    c = dict(layout="coco_cut", strategy="spatial", num_class=11, N=16, T=20, L=30, I=15)
    skel, sensor, target, _ = O.synthetic_batch(c["N"], c["T"], 14, 11, sensor_len=30, sensor_ch=15, seed=14)
    mot = skel[:, :2, 1:] - skel[:, :2, :-1]
    res = run_train_step(mod, lambda: mod((skel, mot, sensor)), target)
    torch.save({"config": c, "shapes": shapes, "fill_seed": 8, "batch_seed": 14, "n_params": sum(p.numel() for p in mod.parameters()),
                **res}, os.path.join(OUT, "nb_two_stream.pt"))
    print("notebook two-stream loss", res["loss"], "params", sum(p.numel() for p in mod.parameters()))


if __name__ == "__main__":
    if sys.argv[1:] == ["notebook"]:
        main_notebook()
    elif sys.argv[1:] == ["tragcn"]:
        main_tragcn()
    elif sys.argv[1:] == ["musa"]:
        main_musa()
    elif sys.argv[1:] == ["musa_ablation"]:
        main_musa("Ablation", "musa_ablation_coco_uniform.pt")
    else:
        main()
        main_tragcn()
        main_musa()
        main_notebook()
