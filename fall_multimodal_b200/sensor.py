"""Accelerometer branch: the notebook ``CNN1D`` over the CUDA kernels of csrc/sensor.cu.

Reference: ``/root/reference/GSTCAN_HAR_conv_10kfold.ipynb#cell2:L6-27`` (CNN1D: two
Conv1d(k5,p2)-BatchNorm1d-ReLU-MaxPool1d(2) stages; the ``fc`` layer exists but is unused in
forward) — same attribute names / state_dict keys; ``forward(x)`` takes ``(N, Cin, L)`` and
returns the ``(N, 32, L//4)`` feature map like the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

EPS, MOMENTUM = 1e-5, 0.1


def _stage_fwd(x, w, b, bn_w, bn_b, rm, rv, training):
    """x (N,L,Ci) -> conv out y (N,L,Co), BN scale/shift + stats, pooled (N,L//2,Co)."""
    dev = x.device
    N, Ln, _ = x.shape
    Co = w.shape[0]
    y = torch.empty(N, Ln, Co, device=dev)
    ops.conv1d_k5_fwd(x, w, b, y)
    st = torch.zeros(2 * Co, dtype=torch.float64, device=dev)
    if training:
        ops.colstats(y.view(N, Ln, 1, Co), st[:Co], st[Co:])
    a, sh, mean, rstd = (torch.empty(Co, device=dev) for _ in range(4))
    ops.bn_finalize(st[:Co], st[Co:], N * Ln, bn_w, bn_b, rm, rv, training, a, sh, mean, rstd, MOMENTUM, EPS)
    out = torch.empty(N, Ln // 2, Co, device=dev)
    ops.bn_relu_pool2_fwd(y, a, sh, out)
    return y, a, sh, mean, rstd, out


def _stage_bwd(x, y, a, sh, mean, rstd, w, dout, training, need_dx):
    dev = x.device
    N, Ln, Ci = x.shape
    Co = w.shape[0]
    dh = torch.empty_like(y)
    ops.pool2_bwd(y, a, sh, dout, dh)
    T1 = torch.zeros(2 * Co, dtype=torch.float64, device=dev)
    y4, dh4 = y.view(N, Ln, 1, Co), dh.view(N, Ln, 1, Co)
    ops.bn1_bwd_reduce(dh4, y4, a, sh, T1[:Co], T1[Co:])
    c1, c2, c3 = (torch.empty(Co, device=dev) for _ in range(3))
    dgam, dbet = torch.zeros(Co, device=dev), torch.zeros(Co, device=dev)
    ops.bn1_bwd_coef(T1[:Co], T1[Co:], a, mean, rstd, N * Ln, training, c1, c2, c3, dgam, dbet)
    dy = torch.empty_like(y)
    ops.bn1_bwd_apply(dh4, y4, a, sh, c1, c2, c3, dy.view(N, Ln, 1, Co), None)
    dx = torch.empty_like(x) if need_dx else None
    dw, db = torch.empty_like(w), torch.empty(Co, device=dev)
    ops.conv1d_k5_bwd(x, dy, w, dx, dw, db)
    return dx, dw, db, dgam, dbet


class _CNN1DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, training, x_cl, w1, b1, g1, be1, rm1, rv1, w2, b2, g2, be2, rm2, rv2):
        with torch.autocast("cuda", enabled=False):
            x_cl = x_cl.float().contiguous()
            s1 = _stage_fwd(x_cl, w1, b1, g1, be1, rm1, rv1, training)
            s2 = _stage_fwd(s1[5], w2, b2, g2, be2, rm2, rv2, training)
        ctx.saved = (x_cl, s1, s2, w1, w2, training)
        return s2[5]

    @staticmethod
    def backward(ctx, dout):
        x_cl, s1, s2, w1, w2, training = ctx.saved
        with torch.autocast("cuda", enabled=False):
            dp1, dw2, db2, dg2, dbe2 = _stage_bwd(s1[5], s2[0], s2[1], s2[2], s2[3], s2[4], w2, dout.float().contiguous(),
                                                  training, True)
            _, dw1, db1, dg1, dbe1 = _stage_bwd(x_cl, s1[0], s1[1], s1[2], s1[3], s1[4], w1, dp1, training, False)
        ctx.saved = None
        return (None, None, dw1, db1, dg1, dbe1, None, None, dw2, db2, dg2, dbe2, None, None)


class CNN1D(nn.Module):
    """Parameter container with the notebook's layout; the math runs in csrc/sensor.cu."""

    def __init__(self, in_channels: int = 15, seq_len: int = 30):
        super().__init__()
        self.layer1 = nn.Sequential(nn.Conv1d(in_channels, 16, kernel_size=5, padding=2), nn.BatchNorm1d(16), nn.ReLU(),
                                    nn.MaxPool1d(2))
        self.layer2 = nn.Sequential(nn.Conv1d(16, 32, kernel_size=5, padding=2), nn.BatchNorm1d(32), nn.ReLU(),
                                    nn.MaxPool1d(2))
        self.fc = nn.Linear(32 * (seq_len // 4), 32)  # present (and unused) in the reference too

    def forward_channels_last(self, x_cl):
        """x_cl: (N, L, Cin) as the dataloader delivers it -> (N, L//4, 32) channels-last feature map."""
        if not x_cl.is_cuda:
            raise RuntimeError("fall_multimodal_b200.CNN1D runs on CUDA (sm_100a) only; there is no CPU fallback")
        l1, l2 = self.layer1, self.layer2
        out = _CNN1DFn.apply(self.training, x_cl, l1[0].weight, l1[0].bias, l1[1].weight, l1[1].bias, l1[1].running_mean,
                             l1[1].running_var, l2[0].weight, l2[0].bias, l2[1].weight, l2[1].bias, l2[1].running_mean,
                             l2[1].running_var)
        if self.training:
            with torch.no_grad():
                l1[1].num_batches_tracked += 1
                l2[1].num_batches_tracked += 1
        return out

    def forward(self, x):
        """x: (N, Cin, L) like the reference; returns (N, 32, L//4)."""
        return self.forward_channels_last(x.permute(0, 2, 1)).permute(0, 2, 1)
