"""GPU parity of the tcgen05 weight-gradient engine (csrc/wgrad.cu) against torch fp64 math."""
import pytest
import torch

gpu = pytest.mark.gpu

CASES = [
    # name, N, T, V, Cin, Cout, shifts, istride, prologue
    ("1x1_c64", 2, 16, 4, 64, 64, [0], 1, False),
    ("1x1_ragged", 3, 19, 5, 64, 64, [0], 1, False),
    ("1x1_c128_256", 3, 19, 5, 128, 256, [0], 1, False),
    ("1x1_cin9", 2, 16, 33, 9, 64, [0], 1, False),
    ("1x1_cin192_c64", 2, 20, 9, 192, 64, [0], 1, False),
    ("1x1_s2_res", 2, 32, 6, 64, 128, [0], 2, False),
    ("t9_s1_bn_c64", 3, 20, 5, 64, 64, list(range(-4, 5)), 1, True),
    ("t9_s2_c128", 2, 31, 7, 128, 128, list(range(-4, 5)), 2, True),
    ("t9_s1_c256", 2, 16, 9, 256, 256, list(range(-4, 5)), 1, True),
    ("big_rows", 16, 64, 33, 64, 64, list(range(-4, 5)), 1, True),
    # plain operands with >= 128 input channels: tensor-map TMA stages on clip-aligned column groups (V = 33: 5 groups per
    # clip, the last one 1 joint + 7 zero columns; V = 14: 6 zero columns; frames outside the clip zero-filled by the TMA)
    ("t9_s1_c128_tma", 5, 40, 33, 128, 128, list(range(-4, 5)), 1, False),
    ("t9_s2_c256_tma", 3, 31, 14, 256, 256, list(range(-4, 5)), 2, False),
    ("t9_c256_t16_tma", 11, 16, 33, 256, 256, list(range(-4, 5)), 1, False),
    ("t5_c128_256_tma", 4, 23, 25, 128, 256, [2, 1, 0, -1, -2], 1, False),
]


@gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_wgrad_matches_torch(case, dtype):
    from fall_multimodal_b200 import ops

    name, N, T, V, Cin, Cout, shifts, istride, prologue = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(4321)
    x = torch.randn(N, T, V, Cin, generator=g).to(dev).to(dtype).contiguous()
    Tj = (T - 1) // istride + 1
    dy = torch.randn(N, Tj, V, Cout, generator=g).to(dev).to(dtype).contiguous()
    scale = shift = None
    if prologue:
        scale = (torch.rand(Cin, generator=g) + 0.5).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.3).to(dev)
    ntaps = len(shifts)
    # torch conv weight layout (Cout, Cin, ntaps)
    dw = torch.zeros(Cout, Cin, ntaps, device=dev, dtype=torch.float32)
    ops.wgrad(x, dy, dw, shifts=shifts, istride=istride, in_scale=scale, in_shift=shift,
              in_relu=prologue, s_m=1, s_c2=ntaps, s_co=Cin * ntaps)
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    f = x.float()
    if prologue:
        f = (f * scale + shift).clamp_min(0)
    if dtype == torch.bfloat16:
        f = f.bfloat16().float()
    f = f.double()
    dyd = dy.double()
    ref = torch.zeros(Cout, Cin, ntaps, device=dev, dtype=torch.float64)
    for m, s in enumerate(shifts):
        for j in range(Tj):
            ti = j * istride + s
            if 0 <= ti < T:
                ref[:, :, m] += torch.einsum("nvo,nvi->oi", dyd[:, j], f[:, ti])
    err = (dw.double() - ref).abs().max().item() / ref.abs().max().item()
    tol = 2e-5 if dtype == torch.bfloat16 else 1e-5
    assert err < tol, f"{name}: rel-to-max err {err:.3e} (tol {tol})"
