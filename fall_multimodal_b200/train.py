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
    """``lr`` handling under graph capture: a Python-float learning rate would be baked into the captured optimizer kernels,
    so every ``param_group["lr"]`` is replaced by a 0-d device tensor the captured step reads. Stock ``torch.optim.lr_scheduler``
    schedulers ``fill_`` such a tensor in place; schedulers that *assign* a float (timm's, the reference's ``optimizer.py``
    cosine/step factories) are picked up by ``run()``, which copies a re-assigned value back into the tensor. ``set_lr`` does the
    same explicitly. Constructing a TrainStep leaves model, buffers and optimizer state exactly as they were: the warm-up and
    capture steps run on the real tensors (the graph must own their addresses) and are rolled back afterwards."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn, example_inputs: Sequence,
                 example_target: torch.Tensor, autocast_dtype: torch.dtype | None = torch.bfloat16,
                 bucket_groups: Sequence | None = None, use_graph: bool = True, warmup: int = 2,
                 max_norm: float | None = None, fused_loss: bool = False, label_smoothing: float = 0.0):
        dev = example_target.device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs CUDA tensors (the product path has no CPU fallback)")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.autocast_dtype = autocast_dtype
        self.max_norm = max_norm           # MF3/main.py:108 clip_grad_norm_(model.parameters(), max_norm) before the step
        # fused_loss: late-fusion Linear + CrossEntropy(label_smoothing) in one kernel pair (model.forward_loss, csrc/head.cu)
        # instead of model(...) followed by loss_fn
        self.fused_loss, self.label_smoothing = bool(fused_loss), float(label_smoothing)
        if self.fused_loss and not hasattr(model, "forward_loss"):
            raise ValueError("fused_loss=True needs a model with forward_loss(skel, sensor, target) (the fusion models)")
        if use_graph:
            for g in optimizer.param_groups:
                if not g.get("capturable", False) and not getattr(optimizer, "graph_safe", False):
                    raise ValueError("TrainStep(use_graph=True) needs an optimizer built with capturable=True (or a graph-safe "
                                     "optimizer such as fall_multimodal_b200.optim.FusedRMSprop): a non-capturable torch.optim "
                                     "optimizer would replay with its step count / lr frozen at capture time")
        self._lr = []
        for g in optimizer.param_groups:
            lr = g["lr"]
            t = lr if torch.is_tensor(lr) and lr.is_cuda else torch.tensor(float(lr), dtype=torch.float32, device=dev)
            g["lr"] = t
            self._lr.append(t)
        self.buckets = GradBuckets(bucket_groups if bucket_groups is not None else [list(model.parameters())])
        # static buffers the graph reads; run() copies each batch into them
        self.inputs = tuple(None if t is None else t.detach().clone() for t in example_inputs)
        self.target = example_target.detach().clone()
        self.loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        self.hits = torch.zeros((), dtype=torch.int64, device=dev)
        self.seen = 0
        self._graphed = None
        if use_graph:
            snap = self._snapshot()
            self._graphed = GraphedStep(self._step, (), warmup=warmup)
            self._restore(snap)
        self.reset_stats()

    # -- state roll-back around warm-up / capture (in place: the graph holds the tensors' addresses) --
    def _snapshot(self):
        model_sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        opt = {}
        for p, st in self.optimizer.state.items():
            opt[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return model_sd, opt

    def _restore(self, snap):
        model_sd, opt = snap
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model_sd[k])
            for p, st in self.optimizer.state.items():
                old = opt.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()      # state created lazily by the warm-up steps: back to its initial value
            for p in self.model.parameters():
                p.grad = None

    def set_lr(self, lr: float, group: int | None = None):
        """Set the learning rate the (captured) optimizer step reads; all groups when ``group`` is None."""
        for i, t in enumerate(self._lr):
            if group is None or group == i:
                t.fill_(float(lr))

    def _sync_lr(self):
        for g, t in zip(self.optimizer.param_groups, self._lr):
            if g["lr"] is not t:            # a scheduler assigned a new value instead of filling the tensor
                t.fill_(float(g["lr"]))
                g["lr"] = t

    def _step(self):
        self.buckets.zero_grad()
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            if self.fused_loss:
                pred, loss = self.model.forward_loss(*self.inputs, self.target, label_smoothing=self.label_smoothing)
            else:
                pred = self.model(*self.inputs)
        if not self.fused_loss:
            loss = self.loss_fn(pred.float(), self.target)
        loss.backward()
        self.buckets.wait()
        if self.max_norm is not None and getattr(self.optimizer, "max_norm", None) is None:   # FusedRMSprop clips inside its step
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_norm, foreach=True)
        self.optimizer.step()
        with torch.no_grad():
            self.loss_sum += loss.detach()
            labels = self.target.argmax(-1) if self.target.dim() == 2 else self.target
            self.hits += (pred.argmax(-1) == labels).sum()
        return loss

    def run(self, inputs: Sequence, target: torch.Tensor) -> torch.Tensor:
        """One train step on a new batch (same shapes as the example batch). Returns the (device) loss tensor."""
        self._sync_lr()
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

    def close(self):
        """Detach from the model (removes the gradient hooks) so another TrainStep can be built on it."""
        self.buckets.close()
        self._graphed = None
