"""GPU parity of the tcgen05 tap-conv engine (csrc/tapconv.cu) against plain torch fp32 math.

The checker restates the op with torch matmuls; in bf16 mode both sides see the same
bf16-rounded operands so only accumulation order and the final store rounding differ.
"""
import pytest
import torch

gpu = pytest.mark.gpu


def ref_tapconv(x, W, shifts, tj, istride, ostride, ooff, Tout, scale, shift, relu, bias, bf16):
    N, Tin, V, Cin = x.shape
    Cout = W.shape[1]
    f = x.float()
    if scale is not None:
        f = f * scale + shift
    if relu:
        f = f.clamp_min(0)
    Wf = W.float()
    if bf16:
        f = f.bfloat16().float()
        Wf = Wf.bfloat16().float()
    out = torch.zeros(N, Tout, V, Cout, device=x.device, dtype=torch.float64)
    for j in range(tj):
        acc = torch.zeros(N, V, Cout, device=x.device, dtype=torch.float64)
        if bias is not None:
            acc += bias.double()
        for m, s in enumerate(shifts):
            ti = j * istride + s
            if 0 <= ti < Tin:
                acc += f[:, ti].double() @ Wf[m].double().t()
        out[:, j * ostride + ooff] = acc
    return out


CASES = [
    # name, N, T, V, Cin, Cout, ntaps/shifts, istride, prologue
    ("1x1_c64", 2, 16, 4, 64, 64, [0], 1, False),
    ("1x1_ragged", 3, 19, 5, 64, 64, [0], 1, False),
    ("t9_s1_bn", 3, 20, 5, 64, 64, list(range(-4, 5)), 1, True),
    ("t9_s2_c128_256", 2, 31, 7, 128, 256, list(range(-4, 5)), 2, True),
    ("1x1_cin9", 2, 16, 33, 9, 64, [0], 1, False),
    ("1x1_cout192", 2, 16, 9, 64, 192, [0], 1, False),
    ("1x1_cout768", 2, 8, 9, 256, 768, [0], 1, False),
    ("1x1_s2_res", 2, 32, 6, 64, 128, [0], 2, False),
    ("t5_dgrad_even", 2, 16, 6, 128, 64, [2, 1, 0, -1, -2], 1, False),
    # more row tiles than CTA pairs (148 SMs = 74 pairs): several waves, odd tile counts, the weight ring wraps many times
    ("t9_c256_waves", 37, 16, 33, 256, 256, list(range(-4, 5)), 1, False),
    ("t9_c128_waves", 21, 32, 33, 128, 128, list(range(-4, 5)), 1, False),
    ("t9_c64_s2_waves", 23, 64, 33, 64, 64, list(range(-4, 5)), 2, True),
]


@gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tapconv_matches_torch(case, dtype):
    from fall_multimodal_b200 import ops

    name, N, T, V, Cin, Cout, shifts, istride, prologue = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.randn(N, T, V, Cin, generator=g).to(dev).to(dtype).contiguous()
    ntaps = len(shifts)
    W = (torch.randn(ntaps, Cout, Cin, generator=g) / (Cin * ntaps) ** 0.5).to(dev).contiguous()
    bias = torch.randn(Cout, generator=g).to(dev)
    scale = shift = None
    if prologue:
        scale = (torch.rand(Cin, generator=g) + 0.5).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.3).to(dev)
    if name == "t5_dgrad_even":
        # transposed-conv phase: out positions 2j of a length-2T output
        tj, ostride, ooff, Tout = T, 2, 0, 2 * T
    else:
        Tout = (T - 1) // istride + 1
        tj, ostride, ooff = Tout, 1, 0
    pw = ops.tapconv_pack(W, Cout, Cin, Cout, Cin, 0, Cin, 0, 1, Cout * Cin, list(range(ntaps)), dtype)
    out = torch.full((N, Tout, V, Cout), 7.0, device=dev, dtype=dtype)
    ops.tapconv(x, pw, out, shifts=shifts, tj=tj, istride=istride, ostride=ostride, ooff=ooff,
                in_scale=scale, in_shift=shift, in_relu=prologue, bias=bias)
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    ref = ref_tapconv(x, W, shifts, tj, istride, ostride, ooff, Tout, scale, shift, prologue, bias,
                      dtype == torch.bfloat16)
    got = out.double()
    if name == "t5_dgrad_even":
        assert torch.all(got[:, 1::2] == 7.0), "rows outside the output phase must be untouched"
        got, ref = got[:, 0::2], ref[:, 0::2]
    denom = ref.abs().max().item()
    err = (got - ref).abs().max().item() / denom
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-5
    assert err < tol, f"{name}: rel-to-max err {err:.3e} (tol {tol})"
