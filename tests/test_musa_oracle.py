"""CPU: the musa Model / Ablation oracle against the fixtures generated from the unmodified reference."""
import pytest
import torch

from oracle import musa_oracle as MO
from oracle import stgcn_oracle as O
from tests.golden_util import check_grads, load


@pytest.mark.parametrize("name,ablation", [("musa_coco_uniform", False), ("musa_ablation_coco_uniform", True)])
def test_musa_oracle_matches_reference_fixture(name, ablation):
    fx = load(name)
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
    out = MO.musa_forward(sd, skel, training=True, ablation=ablation)
    loss = torch.nn.CrossEntropyLoss()(out, target)
    loss.backward()
    assert (out - fx["logits"]).abs().max().item() / fx["logits"].abs().max().item() < 1e-6
    assert abs(loss.item() - fx["loss"]) < 1e-6
    grads = {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
    check_grads(grads, fx["grads"], 1e-6)
    with torch.no_grad():
        ev = MO.musa_forward({k: v.detach() for k, v in sd.items()}, skel, training=False, ablation=ablation)
    assert (ev - fx["eval_logits"]).abs().max().item() / fx["eval_logits"].abs().max().item() < 1e-6


@pytest.mark.parametrize("name,cls", [("musa_coco_uniform", "Model"), ("musa_ablation_coco_uniform", "Ablation")])
def test_musa_state_dict_keys_match_reference(name, cls):
    """CPU: constructing the drop-in classes needs no GPU; keys, shapes and parameter count equal the reference module's."""
    import warnings

    from fall_multimodal_b200 import musa
    fx = load(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = getattr(musa, cls)(num_class=11, num_point=14, max_frame=300, graph=musa.adjGraph(layout="coco_cut", strategy="uniform"),
                               bias=True, edge=True, block_size=41, embed_dim=64, n_stage=1, act_type="tanh")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == fx["shapes"]
    assert sum(p.numel() for p in m.parameters()) == fx["n_params"]
