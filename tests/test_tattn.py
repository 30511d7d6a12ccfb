"""GPU parity of the flash-style time-axis attention (csrc/tattn.cu; TA.py:55-62) against an fp64 torch restatement
(softmax(q k^T / sqrt(c), -1) v per (clip, joint)) and the materialising bgemm path it replaces."""
import math

import pytest
import torch

gpu = pytest.mark.gpu


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def _truth(q, k, v, T):
    """q, k (B,F,V,Tp) feature-major, v (B,V,T,C) -> (B,T,V,C) in fp64 (TA.py:47-62)."""
    Q = q[..., :T].permute(0, 2, 3, 1)            # (B,V,T,F)
    K = k[..., :T].permute(0, 2, 3, 1)
    A = torch.softmax(Q @ K.transpose(-1, -2) / math.sqrt(v.shape[-1]), -1)
    return (A @ v).permute(0, 2, 1, 3)


@gpu
@pytest.mark.parametrize("B,V,T,F", [(2, 5, 30, 62), (1, 3, 300, 62), (3, 4, 16, 62), (2, 2, 129, 40)])
def test_flash_attention_matches_fp64(B, V, T, F):
    from fall_multimodal_b200.tragcn import _Attention, _AttentionF
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(T)
    Tp = (T + 63) // 64 * 64
    q = torch.zeros(B, F, V, Tp)
    k = torch.zeros(B, F, V, Tp)
    q[..., :T] = torch.randn(B, F, V, T, generator=g) * 0.8
    k[..., :T] = torch.randn(B, F, V, T, generator=g) * 0.8
    v = torch.randn(B, V, T, 64, generator=g)
    go = torch.randn(B, T, V, 64, generator=g)
    q, k, v, go = (t.to(dev).bfloat16() for t in (q, k, v, go))
    leaves = [t.double().requires_grad_(True) for t in (q, k, v)]
    truth = _truth(*leaves, T)
    truth.backward(go.double())
    res = {}
    for name, fn in (("bgemm", _Attention), ("flash", _AttentionF), ("flash_btvc", _AttentionF)):
        ins = [t.clone().requires_grad_(True) for t in (q, k, v)]
        if name == "flash_btvc":      # v handed over in the layer's own (B,T,V,C) order
            ins[2] = v.permute(0, 2, 1, 3).contiguous().requires_grad_(True)
            out = fn.apply(*ins, True)
        else:
            out = fn.apply(*ins)
        out.backward(go)
        torch.cuda.synchronize()
        gv = ins[2].grad.permute(0, 2, 1, 3) if name == "flash_btvc" else ins[2].grad
        res[name] = [_rel(out, truth), _rel(ins[0].grad[..., :T], leaves[0].grad[..., :T]), _rel(ins[1].grad[..., :T], leaves[1].grad[..., :T]),
                     _rel(gv, leaves[2].grad)]
        if name != "bgemm":     # the padded time columns of dq / dk are written, as zeros
            assert ins[0].grad[..., T:].abs().max().item() == 0 and ins[1].grad[..., T:].abs().max().item() == 0
    print(f"B{B} V{V} T{T} F{F}: (out, dq, dk, dv) bgemm {['%.2e' % e for e in res['bgemm']]} / flash {['%.2e' % e for e in res['flash']]}")
    for name in ("flash", "flash_btvc"):
        for ef, eb in zip(res[name], res["bgemm"]):
            assert ef < 2e-2 and ef <= 1.25 * eb + 3e-3, (name, ef, eb)
