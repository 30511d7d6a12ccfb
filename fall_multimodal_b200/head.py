"""Fused late-fusion head + cross-entropy (csrc/head.cu).

Reference call sites: ``F2/Model/combination.py:37-46`` (``torch.cat`` of the stream features -> ``nn.Linear``),
``F2/main.py:111-113`` / ``:280`` (``CrossEntropyLoss(label_smoothing=...)`` on probability targets, mean over the batch), and
the notebooks' variant that returns ``softmax(logits)`` and feeds THAT to the loss (``GSTCAN_HAR_conv_10kfold.ipynb#cell1:L416``,
``#cell7:L129``) — ``pre_softmax=True``.

``linear_cross_entropy(feats, weight, bias, target)`` returns ``(pred, loss)`` from ONE forward and ONE backward entry point
(3 kernel launches) instead of cat + addmm + log_softmax + mul + sum + mean and their autograd twins. ``pred`` is what the model
would have returned (logits, or probabilities with ``pre_softmax``) and carries no gradient; ``loss.backward()`` fills the
gradients of every feature segment, the weight and the bias. There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib as L


def _fill(args, feats, weight, bias, target, pre_softmax, smoothing):
    for i, f in enumerate(feats):
        args.feat[i] = f.data_ptr()
        args.width[i] = f.shape[1]
    args.nseg = len(feats)
    args.W, args.bias, args.target = weight.data_ptr(), (bias.data_ptr() if bias is not None else None), target.data_ptr()
    args.N, args.C, args.F = target.shape[0], weight.shape[0], weight.shape[1]
    args.pre_softmax, args.smoothing = int(pre_softmax), float(smoothing)


class _LinearCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, target, pre_softmax, smoothing, *feats):
        with torch.autocast("cuda", enabled=False):
            feats = [f.float().contiguous() for f in feats]
            weight, target = weight.float().contiguous(), target.float().contiguous()
            bias_c = bias.float().contiguous() if bias is not None else None
            N, Cc = target.shape
            dev = target.device
            out, prob = torch.empty(N, Cc, device=dev), torch.empty(N, Cc, device=dev)
            prob2 = torch.empty(N, Cc, device=dev) if pre_softmax else None
            loss = torch.zeros((), device=dev)
            a = L.HeadArgs()
            _fill(a, feats, weight, bias_c, target, pre_softmax, smoothing)
            a.out, a.prob, a.prob2, a.loss = out.data_ptr(), prob.data_ptr(), (prob2.data_ptr() if pre_softmax else None), loss.data_ptr()
            L.check(L.load().fmm_head_ce_fwd(C.byref(a), L.stream()), "head_ce_fwd")
        ctx.saved = (feats, weight, bias_c, target, prob, prob2)
        ctx.flags = (pre_softmax, smoothing, bias is not None)
        ctx.mark_non_differentiable(out)
        return out, loss

    @staticmethod
    def backward(ctx, _dout, dloss):
        feats, weight, bias_c, target, prob, prob2 = ctx.saved
        pre_softmax, smoothing, has_bias = ctx.flags
        with torch.autocast("cuda", enabled=False):
            dev = target.device
            N, Cc = target.shape
            a = L.HeadArgs()
            _fill(a, feats, weight, bias_c, target, pre_softmax, smoothing)
            a.prob, a.prob2 = prob.data_ptr(), (prob2.data_ptr() if pre_softmax else None)
            g = dloss.float().contiguous()
            dz = torch.empty(N, Cc, device=dev)
            dW = torch.empty_like(weight)
            db = torch.empty(Cc, device=dev) if has_bias else None
            dfe = []
            for i, f in enumerate(feats):
                need = ctx.needs_input_grad[5 + i]
                d = torch.empty_like(f) if need else None
                dfe.append(d)
                a.dfeat[i] = d.data_ptr() if need else None
            a.gloss, a.dz, a.dW, a.dbias = g.data_ptr(), dz.data_ptr(), dW.data_ptr(), (db.data_ptr() if has_bias else None)
            L.check(L.load().fmm_head_ce_bwd(C.byref(a), L.stream()), "head_ce_bwd")
        ctx.saved = None
        return (dW, db, None, None, None) + tuple(dfe)


def linear_cross_entropy(feats: Sequence[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor | None, target: torch.Tensor,
                         pre_softmax: bool = False, label_smoothing: float = 0.0):
    """``pred, loss`` of ``CrossEntropyLoss(label_smoothing)(Linear(cat(feats)), target)`` with probability ``target`` (N, C);
    class-index targets are one-hot encoded first. At most 4 feature segments and 32 classes."""
    if not target.is_cuda:
        raise RuntimeError("fall_multimodal_b200.linear_cross_entropy runs on CUDA (sm_100a) only; there is no CPU fallback")
    if target.dim() == 1:
        target = torch.nn.functional.one_hot(target, weight.shape[0]).float()
    if len(feats) > 4 or weight.shape[0] > 32:
        raise ValueError("the fused head takes at most 4 feature segments and 32 classes")
    assert sum(f.shape[1] for f in feats) == weight.shape[1], "feature widths do not add up to the Linear's in_features"
    return _LinearCE.apply(weight, bias, target, bool(pre_softmax), float(label_smoothing), *feats)
