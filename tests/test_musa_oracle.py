"""CPU: the musa Model oracle against the fixture generated from the unmodified reference."""
import torch

from oracle import musa_oracle as MO
from oracle import stgcn_oracle as O
from tests.golden_util import check_grads, load


def test_musa_oracle_matches_reference_fixture():
    fx = load("musa_coco_uniform")
    c = fx["config"]
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in MO.fill_musa(fx["shapes"], c["fill_seed"]).items()}
    for k in fx["shapes"]:
        if k.endswith(".A"):
            sd[k] = fx["A"].double().clone()
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k and not k.endswith(".A"):
            v.requires_grad_(True)
    skel, _, target, _ = O.synthetic_batch(c["N"], c["T"], c["V"], 11, seed=c["batch_seed"])
    skel, target = skel.double(), target.double()          # the fixture is the reference evaluated in fp64
    out = MO.musa_forward(sd, skel, training=True)
    loss = torch.nn.CrossEntropyLoss()(out, target)
    loss.backward()
    assert (out - fx["logits"]).abs().max().item() / fx["logits"].abs().max().item() < 1e-6
    assert abs(loss.item() - fx["loss"]) < 1e-6
    grads = {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
    check_grads(grads, fx["grads"], 1e-6)
    with torch.no_grad():
        ev = MO.musa_forward({k: v.detach() for k, v in sd.items()}, skel, training=False)
    assert (ev - fx["eval_logits"]).abs().max().item() / fx["eval_logits"].abs().max().item() < 1e-6
