"""Helpers shared by the parity tests: load golden fixtures, compare against grad summaries."""
import os
import zlib

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def check_summary(name, got: torch.Tensor, ref: dict, rtol: float, atol_scale: float = 0.0, truth: torch.Tensor = None):
    """Compare tensor ``got`` with a fixture summary (full tensor, or sums/norms + 64 samples).

    Error metric: max |got-ref| relative to max |ref| (plus ``atol_scale`` for tensors whose true
    value is ~0, e.g. conv biases feeding a train-mode BatchNorm).

    ``truth``: the fp64 oracle's value of the same tensor. The fixtures were produced by the reference in fp32 on a
    CPU, so they carry the reference's own rounding noise (measured up to 1.2e-4 on ``data_bn.weight``, the gradient
    that accumulates every layer's noise, while the CUDA path sits at 8e-6 from the fp64 value): the tolerance is never
    tighter than twice the distance between the fixture and the fp64 value.
    """
    g = got.detach().double().flatten().cpu()
    t = truth.detach().double().flatten().cpu() if truth is not None else None
    if "full" in ref:
        r = ref["full"].double().flatten()
        assert g.numel() == r.numel(), f"{name}: numel {g.numel()} vs {r.numel()}"
        scale = max(r.abs().max().item(), atol_scale, 1e-30)
        err = (g - r).abs().max().item() / scale
        if t is not None:
            rtol = max(rtol, 2.0 * (t - r).abs().max().item() / scale)
        assert err <= rtol, f"{name}: rel-to-max err {err:.3e} > {rtol}"
        return err
    scale = max(ref["amax"], atol_scale, 1e-30)
    idx = ref["idx"]
    err = ((g[idx] - ref["vals"].double()).abs().max().item()) / scale
    if t is not None:
        rtol = max(rtol, 2.0 * (t[idx] - ref["vals"].double()).abs().max().item() / scale)
    assert err <= rtol, f"{name}: sampled rel-to-max err {err:.3e} > {rtol}"
    l2 = g.norm().item()
    assert abs(l2 - ref["l2"]) <= rtol * 4 * max(ref["l2"], atol_scale), f"{name}: l2 {l2} vs {ref['l2']}"
    return err


# conv/linear biases that feed a train-mode BatchNorm have an analytically ZERO gradient (the BN
# subtracts the batch mean); the reference produces rounding noise there, so they are compared
# against the global gradient scale instead of their own.
ZERO_GRAD_SUFFIXES = ("tcn.2.bias", "residual.0.bias", "atten.1.bias", "layer1.0.bias", "layer2.0.bias")


def grad_scale(grads: dict) -> float:
    return max(v["amax"] if "amax" in v else float(v["full"].abs().max()) for v in grads.values())


def check_grads(named_grads: dict, ref_grads: dict, rtol: float, truth: dict = None):
    gs = grad_scale(ref_grads)
    worst = 0.0
    for k, ref in ref_grads.items():
        assert k in named_grads and named_grads[k] is not None, f"missing gradient for {k}"
        atol = gs if k.endswith(ZERO_GRAD_SUFFIXES) else 1e-3 * gs
        worst = max(worst, check_summary(k, named_grads[k], ref, rtol, atol_scale=atol,
                                         truth=truth.get(k) if truth is not None else None))
    return worst
