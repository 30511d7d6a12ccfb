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


# ------------------------------------------------------------------------------------------------
# BiLSTM sensor branch (reference: F2/Model/bilstm.py:5-58)
# ------------------------------------------------------------------------------------------------
class _LSTMFn(torch.autograd.Function):
    """Bidirectional single-layer LSTM from zero state: x (N,T,I) -> (N,T,2H), persistent-CTA kernels."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        with torch.autocast("cuda", enabled=False):
            x = x.float().contiguous()
            wi = torch.stack([w_ih, w_ih_r]).float().contiguous()
            wh = torch.stack([w_hh, w_hh_r]).float().contiguous()
            bi = torch.stack([b_ih, b_ih_r]).float().contiguous()
            bh = torch.stack([b_hh, b_hh_r]).float().contiguous()
            N, Tn, _ = x.shape
            H = wh.shape[2]
            out = torch.empty(N, Tn, 2 * H, device=x.device)
            need = any(ctx.needs_input_grad)
            gates = torch.empty(2, N, Tn, 4 * H, device=x.device) if need else None
            cseq = torch.empty(2, N, Tn, H, device=x.device) if need else None
            ops.lstm_fwd(x, wi, wh, bi, bh, out, gates, cseq)
        ctx.saved = (x, wi, wh, out, gates, cseq)
        ctx.need_dx = ctx.needs_input_grad[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wi, wh, out, gates, cseq = ctx.saved
        with torch.autocast("cuda", enabled=False):
            dwi, dwh = torch.zeros_like(wi), torch.zeros_like(wh)
            db = torch.zeros(wi.shape[0], wi.shape[1], device=x.device)
            dx = torch.zeros_like(x) if ctx.need_dx else None
            ops.lstm_bwd(x, wi, wh, out, gates, cseq, dout.float().contiguous(), dwi, dwh, db, dx)
        ctx.saved = None
        # b_ih and b_hh receive the same values but must not share storage: autograd adopts the returned tensors as
        # .grad, and an in-place op on aliased gradients (clip_grad_norm_, a second backward) would hit the memory twice
        return dx, dwi[0], dwh[0], db[0], db[0].clone(), dwi[1], dwh[1], db[1], db[1].clone()


class ChannelAttention(nn.Module):
    """Parameter container of the sensor-branch gate (bilstm.py:5-19)."""

    def __init__(self, input_size, reduce_rate=1 / 8):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(input_size, int(input_size * reduce_rate)), nn.ReLU(),
                                       nn.Linear(int(input_size * reduce_rate), input_size), nn.Sigmoid())


class BiLSTM(nn.Module):
    """``BiLSTM(input_size, hidden_size, num_layers, dropout_prob, num_classes=1, feature='last'|'mean')``
    with the reference's state_dict keys (bilstm.py:21-39); ``forward(skel, sensor)`` ignores ``skel``
    (bilstm.py:41). The recurrence runs in csrc/lstm.cu; the (N,128) tail (mean/last, BatchNorm1d,
    channel gate, Linear) is a handful of tiny torch ops in fp32."""

    def __init__(self, input_size, hidden_size, num_layers, dropout_prob, num_classes=1, feature="last"):
        super().__init__()
        if num_layers != 1 or hidden_size != 64:
            raise NotImplementedError("the persistent LSTM kernel implements the reference configuration: 1 layer, H=64")
        self.input_size, self.hidden_size, self.num_layers, self.num_classes = input_size, hidden_size, num_layers, num_classes
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # dropout with one layer is a no-op, as in the reference
            self.lstm1 = nn.LSTM(input_size, hidden_size, num_layers, batch_first=True, bidirectional=True,
                                 dropout=dropout_prob)
        self.batchnorm = nn.BatchNorm1d(hidden_size * 2)
        self.channelattention = ChannelAttention(hidden_size * 2)
        self.feature = feature
        self.fast_inference = True     # no-grad forward: tensor-core recurrence (csrc/lstm_tc.cu) instead of the training kernel
        self.fc = nn.Sequential(nn.Flatten(), nn.Linear(hidden_size * 2, num_classes))

    def features(self, sensor):
        if not sensor.is_cuda:
            raise RuntimeError("fall_multimodal_b200.BiLSTM runs on CUDA (sm_100a) only; there is no CPU fallback")
        l = self.lstm1
        needs_grad = torch.is_grad_enabled() and (sensor.requires_grad or any(p.requires_grad for p in l.parameters()))
        if not needs_grad and self.fast_inference and sensor.shape[2] <= 39:
            # inference: the recurrence on tensor cores (csrc/lstm_tc.cu), only the (N, 128) feature leaves the kernel
            with torch.autocast("cuda", enabled=False):
                x = sensor.float().contiguous()
                wi = torch.stack([l.weight_ih_l0, l.weight_ih_l0_reverse]).float().contiguous()
                wh = torch.stack([l.weight_hh_l0, l.weight_hh_l0_reverse]).float().contiguous()
                bi = torch.stack([l.bias_ih_l0, l.bias_ih_l0_reverse]).float().contiguous()
                bh = torch.stack([l.bias_hh_l0, l.bias_hh_l0_reverse]).float().contiguous()
                feat = torch.empty(x.shape[0], 2 * self.hidden_size, device=x.device)
                return ops.lstm_infer(x, wi, wh, bi, bh, feat, self.feature != "last")
        out = _LSTMFn.apply(sensor, l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0, l.weight_ih_l0_reverse,
                            l.weight_hh_l0_reverse, l.bias_ih_l0_reverse, l.bias_hh_l0_reverse)
        return out[:, -1, :] if self.feature == "last" else out.mean(dim=1)      # bilstm.py:52-55

    def forward(self, skel, sensor):
        with torch.autocast("cuda", enabled=False):
            out = self.features(sensor)
            out = self.batchnorm(out)                                             # bilstm.py:56
            att = self.channelattention.attention
            out = out * att(out)                                                  # bilstm.py:16-19
            return self.fc(out)                                                   # bilstm.py:58


class CNN_BiLSTM(nn.Module):
    """Notebook sensor branch ``CNN_BiLSTM(hidden_size, num_layers, dropout_prob, num_classes, feature)``
    (GSTCAN_HAR_conv_10kfold.ipynb#cell2:L85-100): CNN1D feature map (N,32,L//4) fed as a length-L//4
    sequence of 32-d vectors to ``BiLSTM(32, 64, 1, 0.3, 11, 'mean')``. Like the reference, the constructor
    arguments other than ``num_classes`` / ``feature`` are ignored in favour of those constants; submodule
    names ``cnn`` / ``bilstm`` give the reference's state_dict keys. ``forward(x)`` takes ``(N, L, Cin)``."""

    def __init__(self, hidden_size=64, num_layers=1, dropout_prob=0.3, num_classes=11, feature="mean",
                 in_channels: int = 15, seq_len: int = 30):
        super().__init__()
        self.cnn = CNN1D(in_channels, seq_len)
        self.bilstm = BiLSTM(input_size=32, hidden_size=64, num_layers=1, dropout_prob=0.3, num_classes=num_classes,
                             feature=feature)

    def forward(self, x):
        feat = self.cnn.forward_channels_last(x)       # (N, L//4, 32): already the (batch, seq, feature) the LSTM wants
        return self.bilstm(None, feat)
