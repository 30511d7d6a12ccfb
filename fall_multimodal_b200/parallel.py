"""Batch-sharded data parallelism: bucketed gradient all-reduce launched behind backward.

The reference is single-process / single-device; the only place its hot path
shards naturally is the batch, with ONE exchange step: the gradient sum. One process per GPU
(torchrun), replicated weights, contiguous batch shards.

Gradients are grouped into a few buckets (here: fusion head + sensor branch, motion trunk, joint
trunk). ``zero_grad()`` sets ``p.grad = None`` so autograd hands every produced gradient tensor over
without an accumulate kernel; a post-accumulate hook counts a bucket's parameters down and, when the
last one has its gradient, packs the bucket into its flat buffer (one multi-tensor copy), re-points
``p.grad`` at views of that buffer and queues ``all_reduce(AVG)`` on a communication stream, so the
exchange of one trunk overlaps the backward kernels of the next. ``wait()`` flushes buckets that never
filled up (parameters without a gradient) and joins the communication stream before the optimizer
step. With a single rank nothing is copied or exchanged at all.

BatchNorm statistics are per shard by default (each rank normalises over its own clips), i.e. DP parity
is "every shard matches the single-device result on that shard, gradients are the mean of the shard
gradients" (option b). ``convert_sync_batchnorm(model)`` switches the GSTCAN trunks to
global-batch statistics (option a): DP parity is then "N ranks x B/N clips == one device x B clips",
logits, running statistics and (averaged) gradients alike (engine.py: TrunkEngine.sync_bn).
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, groups: Sequence[Iterable[torch.nn.Parameter]], process_group=None, average: bool = True):
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets = []
        self._comm_stream = None
        self._comm_used = False  # anything queued on the comm stream since the last wait()
        self._handles = []
        self._hook_handles = []
        for params in groups:
            params = [p for p in params if p.requires_grad]
            if not params:
                continue
            b = {"params": params, "pending": len(params), "launched": False, "flat": None, "views": None}
            if self.world > 1:
                total = sum(p.numel() for p in params)
                b["flat"] = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
                off, views = 0, []
                for p in params:
                    views.append(b["flat"][off:off + p.numel()].view_as(p))
                    off += p.numel()
                b["views"] = views
            idx = len(self.buckets)
            self.buckets.append(b)
            for p in params:
                self._hook_handles.append(p.register_post_accumulate_grad_hook(self._make_hook(idx)))
        dev = self.buckets[0]["params"][0].device if self.buckets else torch.device("cpu")
        if dev.type == "cuda" and self.world > 1:
            self._comm_stream = torch.cuda.Stream(device=dev)

    def _make_hook(self, idx):
        def hook(param):
            b = self.buckets[idx]
            b["pending"] -= 1
            if b["pending"] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if self.world == 1 or b["launched"]:
            return
        b["launched"] = True
        have = [(v, p.grad) for v, p in zip(b["views"], b["params"]) if p.grad is not None]
        missing = [v for v, p in zip(b["views"], b["params"]) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v in missing:
            v.zero_()
        for v, p in zip(b["views"], b["params"]):
            if p.grad is not None:
                p.grad = v
        flat = b["flat"]
        op = dist.ReduceOp.AVG if (self.average and flat.is_cuda) else dist.ReduceOp.SUM
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream(flat.device))
            self._comm_used = True
            with torch.cuda.stream(self._comm_stream):
                self._handles.append(dist.all_reduce(flat, op=op, group=self.pg, async_op=True))
        else:
            dist.all_reduce(flat, op=op, group=self.pg)
            if self.average:
                flat.div_(self.world)

    def close(self):
        """Remove the post-accumulate hooks (a second GradBuckets on the same parameters would otherwise double count)."""
        for h in self._hook_handles:
            h.remove()
        self._hook_handles.clear()

    def zero_grad(self):
        """Call instead of ``model.zero_grad()``: gradients become None (no zero-fill, no accumulate kernels)."""
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["launched"] = False
            for p in b["params"]:
                p.grad = None

    def wait(self):
        """Flush incomplete buckets and join the gradient exchange before the optimizer step."""
        for b in self.buckets:
            if not b["launched"]:
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        if self._comm_stream is not None and self._comm_used:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._comm_used = False


def convert_sync_batchnorm(model: torch.nn.Module, process_group=None, strict: bool = False) -> torch.nn.Module:
    """SyncBN for the GSTCAN trunks of ``model`` (same call shape as ``torch.nn.SyncBatchNorm.convert_sync_batchnorm``).

    Every BatchNorm inside a trunk - ``data_bn``, the two of each block's ``tcn``, the residual one and the squeeze-excite
    block's BatchNorm over the batch axis (stgcan.py:63-70, 110-133, 213-218) - then uses the statistics of the global batch:
    per block and direction at most three small collectives (all-reduce of the fused (sum, sum^2) accumulators, one all-gather
    of the per-clip (N, C) rows of the SE / block-tail reductions), enqueued on the compute stream so they are captured with the
    step's CUDA graph. All ranks must hold the same number of clips. The sensor branch's BatchNorm1d layers (CNN1D / BiLSTM
    tail: a few hundred channels-rows per step) keep per-shard statistics; ``strict=True`` raises if the model has any.
    ``process_group=False`` switches the trunks back to per-shard statistics."""
    from .stgcan import STGCAN

    n = 0
    trunk_bn = set()
    for m in model.modules():
        if isinstance(m, STGCAN):
            m._engine.sync_bn = None if process_group is False else (True if process_group is None else process_group)
            n += 1
            trunk_bn.update(id(x) for x in m.modules() if isinstance(x, torch.nn.modules.batchnorm._BatchNorm))
    if n == 0:
        raise ValueError("convert_sync_batchnorm: the model has no GSTCAN trunk")
    if strict:
        other = [k for k, x in model.named_modules()
                 if isinstance(x, torch.nn.modules.batchnorm._BatchNorm) and id(x) not in trunk_bn]
        if other:
            raise NotImplementedError(f"convert_sync_batchnorm(strict=True): BatchNorm layers outside the trunks stay per shard: {other}")
    return model
