"""Batch-sharded data parallelism: bucketed gradient all-reduce launched behind backward.

The reference is single-process / single-device; the only place its hot path
shards naturally is the batch, with ONE exchange step: the gradient sum. One process per GPU
(torchrun), replicated weights, contiguous batch shards. Gradients live in a few flat buckets
(``p.grad`` are views into them); a post-accumulate hook counts a bucket's parameters down and, when
the last one has its gradient, queues ``all_reduce(AVG)`` for the whole bucket on a communication
stream, so the exchange of one trunk overlaps the backward kernels of the next.
``wait()`` joins the communication stream before the optimizer step.

BatchNorm statistics are per shard (each rank normalises over its own clips), i.e. DP parity is
"every shard matches the single-device result on that shard, gradients are the mean of the shard
gradients" (option b).
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
        self._pending = []
        self._handles = []
        self._comm_stream = None
        self._comm_used = False  # anything queued on the comm stream since the last wait()
        for params in groups:
            params = [p for p in params if p.requires_grad]
            if not params:
                continue
            total = sum(p.numel() for p in params)
            flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
            off = 0
            for p in params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            idx = len(self.buckets)
            self.buckets.append({"flat": flat, "params": params, "n": len(params)})
            self._pending.append(len(params))
            for p in params:
                p.register_post_accumulate_grad_hook(self._make_hook(idx))
        dev = self.buckets[0]["flat"].device if self.buckets else torch.device("cpu")
        if dev.type == "cuda":
            self._comm_stream = torch.cuda.Stream(device=dev)

    def _make_hook(self, idx):
        def hook(param):
            self._pending[idx] -= 1
            if self._pending[idx] == 0:
                self._launch(idx)
        return hook

    def _launch(self, idx):
        flat = self.buckets[idx]["flat"]
        if self.world == 1:
            return
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

    def zero_grad(self):
        """Call instead of ``model.zero_grad()`` (keeps ``p.grad`` pointing into the buckets)."""
        for i, b in enumerate(self.buckets):
            b["flat"].zero_()
            self._pending[i] = b["n"]
            for p in b["params"]:
                if p.grad is None or p.grad.data_ptr() < b["flat"].data_ptr():
                    raise RuntimeError("p.grad was detached from its bucket (use GradBuckets.zero_grad)")

    def wait(self):
        """Join the gradient exchange before the optimizer step."""
        for h in self._handles:
            h.wait()
        self._handles.clear()
        if self._comm_stream is not None and self._comm_used:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._comm_used = False
