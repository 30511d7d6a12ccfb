"""Drop-in GSTCAN modules over the B200 kernels.

Same class names, constructor arguments, ``forward(skel, sensor)`` signatures and ``state_dict``
keys as ``/root/reference/Fall_2_Spatial_Temporal_SR/Model/stgcan.py`` (STGCAN :147-228, st_gcan
:79-144, GraphConvolution :8-56, Channel_Attention :59-74), so reference checkpoints load and the
reference ``main.py`` train loop (autocast -> model(skel, sensor) -> CrossEntropyLoss ->
backward -> optimizer.step, main.py:111-132) runs unchanged. The torch.nn layers below are only
PARAMETER CONTAINERS (they give the reference's key names and default initialisation); their own
``forward`` is never used: the whole trunk runs as one ``torch.autograd.Function`` whose forward
and backward launch the hand-written CUDA kernels (engine.py). There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import BLOCK_PLAN, TrunkEngine
from .graph import Graph


class GraphConvolution(nn.Module):
    """Parameter holder of the 1x1 (K*Cout <- Cin) channel-mixing conv (stgcan.py:33-48)."""

    def __init__(self, in_channels, out_channels, kernel_size, t_kernel_size=1, t_stride=1, t_padding=0,
                 t_dilation=1, bias=True):
        super().__init__()
        if (t_kernel_size, t_stride, t_padding, t_dilation, bias) != (1, 1, 0, 1, True):
            raise NotImplementedError("the B200 graph-conv kernel implements the 1x1, stride-1, biased configuration "
                                      "that st_gcan uses")
        self.kernel_size = kernel_size
        self.conv = nn.Conv2d(in_channels, out_channels * kernel_size, kernel_size=(1, 1))


class Channel_Attention(nn.Module):
    """Parameter holder of the squeeze-excite block (stgcan.py:60-70)."""

    def __init__(self, out_channels):
        super().__init__()
        c4 = int(out_channels / 4)
        self.atten = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(out_channels, c4, 1), nn.BatchNorm2d(c4),
                                   nn.ReLU(), nn.Conv2d(c4, out_channels, 1), nn.Sigmoid())


class st_gcan(nn.Module):
    """Parameter holder of one spatial-temporal block (stgcan.py:100-136)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2 and kernel_size[0] == 9, "temporal kernel is 9 (stgcan.py:177)"
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is not used by the reference configs (stgcan.py:120, p=0)")
        self.gcn = GraphConvolution(in_channels, out_channels, kernel_size[1])
        self.tcn = nn.Sequential(nn.BatchNorm2d(out_channels), nn.ReLU(inplace=False),
                                 nn.Conv2d(out_channels, out_channels, (kernel_size[0], 1), (stride, 1), (4, 0)),
                                 nn.BatchNorm2d(out_channels), nn.Dropout(dropout, inplace=True))
        if residual and not (in_channels == out_channels and stride == 1):
            self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                                          nn.BatchNorm2d(out_channels))
        self.channel_attention_module = Channel_Attention(out_channels)


def _compute_dtype(module) -> torch.dtype:
    forced = getattr(module, "compute_dtype", None)
    if forced is not None:
        return forced
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    return torch.float32


class _TrunkFn(torch.autograd.Function):
    """One GSTCAN trunk: skel (N,C,T,V) -> pooled feature (N,256), CUDA kernels both ways."""

    @staticmethod
    def forward(ctx, engine, names, training, dt, skel, *tensors):
        P = dict(zip(names, tensors))
        need_grad = any(ctx.needs_input_grad)
        with torch.autocast("cuda", enabled=False):
            feat, sv = engine.forward(P, skel, training, dt, need_grad)
        ctx.engine, ctx.names, ctx.P, ctx.sv = engine, names, P, sv
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        with torch.autocast("cuda", enabled=False):
            grads = ctx.engine.backward(ctx.P, ctx.sv, dfeat.contiguous().float())
        out = tuple(grads.get(n) for n in ctx.names)
        ctx.sv = None
        return (None, None, None, None, None) + out


class STGCAN(nn.Module):
    """Spatial-temporal graph conv network with channel attention (reference STGCAN, stgcan.py:147-228).

    Args and shapes as the reference: ``STGCAN(in_channels, graph_args, num_class=None,
    edge_importance_weighting=True)``; input ``(N, in_channels, T, V)``, output ``(N, num_class)`` or the
    ``(N, 256)`` pooled feature when ``num_class is None``.
    """

    def __init__(self, in_channels, graph_args, num_class=None, edge_importance_weighting=True, **kwargs):
        super().__init__()
        if not edge_importance_weighting:
            raise NotImplementedError("edge_importance_weighting=False is broken in the reference (stgcan.py:203)")
        graph = Graph(**graph_args)
        A = torch.tensor(graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer("A", A)
        K = A.size(0)
        kernel_size = (9, K)
        kwargs0 = {k: v for k, v in kwargs.items() if k != "dropout"}
        self.data_bn = nn.BatchNorm1d(in_channels * A.size(1))
        blocks = []
        for i, (cin, cout, stride, res) in enumerate(BLOCK_PLAN):
            if i == 0:
                blocks.append(st_gcan(in_channels, cout, kernel_size, stride, residual=False, **kwargs0))
            else:
                blocks.append(st_gcan(cin, cout, kernel_size, stride, **kwargs))
        self.st_gcan_networks = nn.ModuleList(blocks)
        self.edge_importance = nn.ParameterList([nn.Parameter(torch.ones(A.size())) for _ in self.st_gcan_networks])
        self.num_class = num_class
        if num_class is not None:
            self.cls = nn.Conv2d(256, num_class, kernel_size=1)
        self.in_channels = in_channels
        self.compute_dtype = None  # None: bf16 under torch.autocast(bfloat16), else fp32
        self._engine = TrunkEngine(A, in_channels, "st_gcan_networks")

    def _trunk_tensors(self):
        names, tensors = [], []
        for k, v in self.named_parameters():
            if not k.startswith("cls."):
                names.append(k), tensors.append(v)
        for k, v in self.named_buffers():
            names.append(k), tensors.append(v)
        return names, tensors

    def features(self, skel):
        if not skel.is_cuda:
            raise RuntimeError("fall_multimodal_b200.STGCAN runs on CUDA (sm_100a) only; there is no CPU fallback")
        names, tensors = self._trunk_tensors()
        return _TrunkFn.apply(self._engine, names, self.training, _compute_dtype(self), skel.contiguous(), *tensors)

    def forward(self, skel, sensor=None):
        dt = _compute_dtype(self)
        feat = self.features(skel)
        if self.num_class is None:
            return feat.to(dt) if dt == torch.bfloat16 else feat
        with torch.autocast("cuda", enabled=False):
            out = torch.addmm(self.cls.bias, feat, self.cls.weight.view(self.num_class, 256).t())
        return out.to(dt) if dt == torch.bfloat16 else out
