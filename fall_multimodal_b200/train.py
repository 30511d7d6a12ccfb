"""The reference train step as ONE device-side unit.

``F2/main.py:100-135`` runs, per mini-batch: host->device copies, autocast forward, CrossEntropy on soft
targets, backward, ``optimizer.step()``, ``model.zero_grad()``, then ``loss.item()`` and a top-k accuracy on the
host — two synchronisations per iteration. ``TrainStep`` keeps the same arithmetic (stock ``torch.optim``
optimizer in ``capturable`` mode, stock loss) but

  * captures zero_grad .. optimizer.step as a CUDA graph over static input buffers (``graphs.GraphedStep``),
  * all-reduces the gradients in buckets behind backward when a process group is up (``parallel.GradBuckets``),
  * accumulates the running loss and the top-1 hit count on the device; ``stats()`` is the only host sync.

Usage (mirrors the reference loop)::

    ts = TrainStep(model, optimizer, loss_fn, example_inputs=(skel, sensor), example_target=label_onehot)
    for skel, sensor, label in loader:          # pinned host tensors or device tensors
        ts.run((skel, sensor), label)
    mean_loss, top1 = ts.stats()                # one sync per epoch / logging interval
"""
from __future__ import annotations

from typing import Sequence

import torch

from .graphs import GraphedStep
from .parallel import GradBuckets


class TrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn, example_inputs: Sequence,
                 example_target: torch.Tensor, autocast_dtype: torch.dtype | None = torch.bfloat16,
                 bucket_groups: Sequence | None = None, use_graph: bool = True, warmup: int = 2):
        dev = example_target.device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs CUDA tensors (the product path has no CPU fallback)")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.autocast_dtype = autocast_dtype
        self.buckets = GradBuckets(bucket_groups if bucket_groups is not None else [list(model.parameters())])
        # static buffers the graph reads; run() copies each batch into them
        self.inputs = tuple(None if t is None else t.detach().clone() for t in example_inputs)
        self.target = example_target.detach().clone()
        self.loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        self.hits = torch.zeros((), dtype=torch.int64, device=dev)
        self.seen = 0
        self._graphed = GraphedStep(self._step, (), warmup=warmup) if use_graph else None
        self.reset_stats()   # the warm-up / capture runs above are real optimizer steps but not part of the statistics

    def _step(self):
        self.buckets.zero_grad()
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            pred = self.model(*self.inputs)
        loss = self.loss_fn(pred.float(), self.target)
        loss.backward()
        self.buckets.wait()
        self.optimizer.step()
        with torch.no_grad():
            self.loss_sum += loss.detach()
            labels = self.target.argmax(-1) if self.target.dim() == 2 else self.target
            self.hits += (pred.argmax(-1) == labels).sum()
        return loss

    def run(self, inputs: Sequence, target: torch.Tensor) -> torch.Tensor:
        """One train step on a new batch (same shapes as the example batch). Returns the (device) loss tensor."""
        for dst, src in zip(self.inputs, inputs):
            if dst is not None and src is not dst:
                dst.copy_(src, non_blocking=True)
        if target is not self.target:
            self.target.copy_(target, non_blocking=True)
        self.seen += self.target.shape[0]
        return self._graphed.replay() if self._graphed is not None else self._step()

    def stats(self):
        """(mean loss per step, top-1 accuracy) since the last reset — the only host synchronisation."""
        steps = max(1, self.seen // self.target.shape[0])
        return self.loss_sum.item() / steps, self.hits.item() / max(1, self.seen)

    def reset_stats(self):
        self.loss_sum.zero_()
        self.hits.zero_()
        self.seen = 0
