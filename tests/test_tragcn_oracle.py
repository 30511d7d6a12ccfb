"""CPU: the TRAGCN oracle restatement against fixtures generated from the unmodified reference
(oracle/make_golden.py tragcn) — SURVEY.md 8a rows 15-19, 8(c)."""
import pytest
import torch

from oracle import tragcn_oracle as TO
from tests.golden_util import check_summary, load


@pytest.mark.parametrize("name", ["targcn_v25_t12", "targcn_v14_t30_adj"])
def test_targcn_oracle_matches_reference_fixture(name):
    fx = load(name)
    c = fx["config"]
    shapes = TO.targcn_param_shapes(V=c["V"], T=c["T"])
    assert shapes == fx["shapes"]
    assert sum(int(torch.tensor(s).prod()) for k, s in shapes.items() if not k.endswith("PE.pe")) == fx["n_params"]
    sd = {k: v.requires_grad_(not k.endswith("PE.pe")) for k, v in TO.fill_targcn(shapes, c["fill_seed"]).items()}
    x, tgt = TO.synthetic_clips(c["B"], c["T"], c["V"], seed=c["batch_seed"])
    logits = TO.targcn_forward(sd, x, adj=fx["adj"])
    loss = torch.nn.CrossEntropyLoss()(logits, tgt)
    loss.backward()
    assert (logits - fx["logits"]).abs().max().item() / fx["logits"].abs().max().item() < 2e-6
    assert abs(loss.item() - fx["loss"]) < 1e-6
    gs = max(v["amax"] if "amax" in v else float(v["full"].abs().max()) for v in fx["grads"].values())
    for k, ref in fx["grads"].items():
        check_summary(k, sd[k].grad, ref, 2e-5, atol_scale=1e-3 * gs)


def test_targcn_oracle_fp64_self_consistency():
    """fp32 vs fp64 evaluation of the oracle agree to fp32 rounding (sanity of the parity tolerance)."""
    shapes = TO.targcn_param_shapes(V=14, T=8)
    sd = TO.fill_targcn(shapes, 3)
    x, _ = TO.synthetic_clips(2, 8, 14, seed=1)
    a = TO.targcn_forward(sd, x)
    b = TO.targcn_forward({k: v.double() for k, v in sd.items()}, x.double())
    assert (a.double() - b).abs().max().item() / b.abs().max().item() < 1e-5
