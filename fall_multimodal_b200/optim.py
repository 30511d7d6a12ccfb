"""Multi-tensor RMSprop on the device (csrc/optim.cu): the reference's optimizer (``F2/optimizer.py:20-21``:
``torch.optim.RMSprop(parameters, lr)``; every shipped config selects it) with the unscale + ``clip_grad_norm_`` step of
``Multimodal_Fall3/main.py:103-113`` folded in.

One kernel launch updates every parameter tensor (two with clipping / loss scaling: the global gradient norm first). ``lr``, the
norm and the inverse loss scale live in device memory, so the whole step is CUDA-graph capturable and a scheduler can change the
learning rate between replays (``param_group["lr"]`` may be assigned a float or filled in place; the step re-reads it).
``state_dict()`` uses torch's RMSprop layout (``square_avg``, ``step``), so reference checkpoints' optimizer states load.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class FusedRMSprop(torch.optim.Optimizer):
    graph_safe = True     # TrainStep: no capturable flag needed

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0.0, max_norm: float | None = None):
        if lr < 0 or eps < 0 or alpha < 0 or weight_decay < 0:
            raise ValueError("invalid RMSprop hyper-parameter")
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay))
        self.max_norm = max_norm
        self.inv_scale = None          # device scalar 1 / loss_scale (GradScaler-style); None = gradients are unscaled
        self._plans = {}

    def _lr_tensor(self, group, dev):
        lr = group["lr"]
        t = group.get("_lr_dev")
        if t is None or t.device != dev:
            t = torch.zeros((), dtype=torch.float32, device=dev)
            group["_lr_dev"] = t
            group["_lr_seen"] = None
        if torch.is_tensor(lr):
            if lr is not t:
                t.copy_(lr.detach().to(dev, torch.float32), non_blocking=True)
        elif group.get("_lr_seen") != float(lr):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedRMSprop: lr changed on the host during graph capture; set it before capturing "
                                   "or hold it in a device tensor")
            t.fill_(float(lr))
            group["_lr_seen"] = float(lr)
        return t

    def _plan(self, gi, group):
        """Device tables for the group's tensors that currently have gradients (rebuilt when any pointer changes)."""
        ps = [p for p in group["params"] if p.grad is not None]
        if not ps:
            return None
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.dtype == torch.float32):
                raise RuntimeError("FusedRMSprop needs contiguous fp32 CUDA parameters and gradients (no CPU fallback)")
            st = self.state[p]
            if "square_avg" not in st:
                st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        key = tuple((p.data_ptr(), p.grad.data_ptr() if p.grad.is_contiguous() else -1, self.state[p]["square_avg"].data_ptr()) for p in ps)
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key:
            return plan
        dev = ps[0].device
        chunk = L.load().fmm_opt_chunk()
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        rows, ct, co = [], [], []
        for i, (p, g) in enumerate(zip(ps, grads)):
            rows.append([p.data_ptr(), g.data_ptr(), self.state[p]["square_avg"].data_ptr(), p.numel()])
            for off in range(0, p.numel(), chunk):
                ct.append(i)
                co.append(off)
        # pinned staging copies stay referenced by the plan: inside a CUDA graph the H2D copies become memcpy nodes that
        # re-read them at every replay
        host = [torch.tensor(rows, dtype=torch.int64).pin_memory(), torch.tensor(ct, dtype=torch.int32).pin_memory(),
                torch.tensor(co, dtype=torch.int64).pin_memory()]
        plan = {"key": key, "params": ps, "grads": grads, "host": host,
                "tensors": host[0].to(dev, non_blocking=True), "ct": host[1].to(dev, non_blocking=True),
                "co": host[2].to(dev, non_blocking=True), "n": len(ct),
                "norm": torch.zeros((), dtype=torch.float32, device=dev)}
        self._plans[gi] = plan
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        plans = [(g, self._plan(i, g)) for i, g in enumerate(self.param_groups)]
        plans = [(g, pl) for g, pl in plans if pl is not None]
        need_norm = self.max_norm is not None or self.inv_scale is not None
        norm = None
        if need_norm and plans:
            norm = plans[0][1]["norm"]          # one global norm over all groups (clip_grad_norm_(model.parameters()))
            norm.zero_()
            for _, pl in plans:
                L.check(lib.fmm_grad_norm_sq(L.ptr(pl["tensors"]), L.ptr(pl["ct"]), L.ptr(pl["co"]), pl["n"], L.ptr(norm),
                                             L.stream()), "grad_norm_sq")
        for g, pl in plans:
            lr = self._lr_tensor(g, pl["tensors"].device)
            L.check(lib.fmm_rmsprop_step(L.ptr(pl["tensors"]), L.ptr(pl["ct"]), L.ptr(pl["co"]), pl["n"], L.ptr(lr),
                                         float(g["alpha"]), float(g["eps"]), float(g["weight_decay"]), L.ptr(norm),
                                         float(self.max_norm or 0.0), L.ptr(self.inv_scale), L.stream()), "rmsprop_step")
        self._steps = getattr(self, "_steps", 0) + 1
        return loss

    def state_dict(self):
        for st in self.state.values():       # torch.optim.RMSprop's layout: a step count next to square_avg
            st["step"] = torch.tensor(float(getattr(self, "_steps", 0)))
        sd = super().state_dict()
        for g in sd["param_groups"]:          # device-side helpers are not part of the checkpoint
            g.pop("_lr_dev", None)
            g.pop("_lr_seen", None)
            if torch.is_tensor(g.get("lr")):
                g["lr"] = float(g["lr"])
        return sd
