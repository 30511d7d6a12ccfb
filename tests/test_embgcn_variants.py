"""EmbGCN_noGate / EmbGCN_linear (EmbGCN.py:91-123, selected in the reference by editing the import at GRU.py:3-6):
the oracle restatement against golden vectors from the unmodified reference classes (CPU), and the whole TARGCN built on each
variant against the fp64 oracle on the GPU (fp32 parity mode = per-step kernels, bf16 = persistent scan kernels)."""
import os

import pytest
import torch

from oracle import tragcn_oracle as TO

gpu = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "embgcn_variants.pt")


@pytest.mark.parametrize("name", ["noGate", "linear"])
def test_oracle_variant_matches_reference_golden(name):
    fx = torch.load(GOLD)[name]
    fn = TO.emb_gcn_nogate if name == "noGate" else TO.emb_gcn_linear
    y = fn(fx["x"].double(), fx["E"].double(), {k: v.double() for k, v in fx["state_dict"].items()}, "")
    assert (y - fx["y"].double()).abs().max().item() < 1e-5 * fx["y"].abs().max().item()


def test_variant_state_dict_keys_match_reference():
    import fall_multimodal_b200 as fmm
    fx = torch.load(GOLD)
    for name, cls in (("noGate", "EmbGCN_noGate"), ("linear", "EmbGCN_linear")):
        m = fmm.TARGCN(num_nodes=14, adj=None, seq_len=8, gcn=cls)
        keys = {k.split("gate.", 1)[1] for k in m.state_dict() if "dcrnn_cells.0.gate." in k}
        assert keys == set(fx[name]["keys"]), (name, keys)


def _fill(m, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {k: (v if k.endswith("PE.pe") else torch.randn(v.shape, generator=g) * (0.3 if "node_emb" in k else 0.08)) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)


@gpu
@pytest.mark.parametrize("cls,variant", [("EmbGCN_noGate", "noGate"), ("EmbGCN_linear", "linear")])
@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-4, 2e-4), (torch.bfloat16, 2e-2, 6e-2)])
def test_targcn_variants_match_oracle(cls, variant, dtype, tol, gtol):
    import fall_multimodal_b200 as fmm
    dev = torch.device("cuda:0")
    V, T, B = 14, 10, 5
    m = fmm.TARGCN(num_nodes=V, adj=None, seq_len=T, gcn=cls)
    _fill(m, 3)
    m = m.to(dev).train()
    x, tgt = TO.synthetic_clips(B, T, V, seed=8)
    sd = {k: v.detach().cpu().double().requires_grad_(not k.endswith("PE.pe")) for k, v in m.state_dict().items()}
    ref = TO.targcn_forward(sd, x.double(), variant=variant)
    torch.nn.CrossEntropyLoss()(ref, tgt.double()).backward()
    if dtype == torch.bfloat16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(x.to(dev))
    else:
        out = m(x.to(dev))
    torch.nn.CrossEntropyLoss()(out.float(), tgt.to(dev)).backward()
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, err
    gs = max(v.grad.abs().max().item() for v in sd.values() if v.grad is not None)
    worst = 0.0
    for k, p in m.named_parameters():
        e = (p.grad.double().cpu() - sd[k].grad).abs().max().item() / max(sd[k].grad.abs().max().item(), 1e-3 * gs)
        worst = max(worst, e)
    print(cls, dtype, "logits", err, "worst grad", worst)
    assert worst < gtol, worst
