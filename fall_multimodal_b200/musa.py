"""The musa ``Model`` that ``Multimodal_Fall3/main.py:307-320`` trains, over CUDA kernels.

Reference: ``Multimodal_Fall3/model/musa_model.py`` — ``embed`` :384-406, ``SpatialGraphConv`` :101-146,
``SepTemporal_Block`` :148-199, ``DepthWiseSeparableConv_*`` :422-458, ``Sep_TCN`` :461-474, ``Classification_Module``
:476-490, ``Model`` :492-591. Same class names, constructor arguments, ``forward(x[N,3,T,V])`` and state_dict keys.

Activations are kept channels-last ``(N,T,V,C)``. The 1x1 convolutions and the joint mixing ``einsum('nctv,cvw->nctw')``
are strided batched GEMMs (csrc/bgemm.cu), the depthwise temporal convolutions, the folded BatchNorm + residual +
activation and their hand-written backward passes run in csrc/musa.cu, BatchNorm statistics in the shared colstats /
bn_finalize kernels. The (N,515)->(N,11) classifier tail and, in training with keep_prob < 1, the two DropBlock
mask generators (``torch.bernoulli`` / ``randperm`` exactly as the reference draws them) are a handful of torch ops.
There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib as L
from . import ops
from .graph import Graph
from .stgcan import _compute_dtype
from .tragcn import _Linear, _splitk, bgemm

ACT = {"linear": 0, None: 0, "relu": 1, "tanh": 2, "leakyrelu01": 3}


def adjGraph(layout="coco_cut", strategy="uniform", max_hop=1, dilation=1):
    """``adjGraph`` of musa_model.py:201-323 (same construction as F2/Model/graph.py)."""
    return Graph(layout=layout, strategy=strategy, max_hop=max_hop, dilation=dilation)


def _dt(t):
    return L.dt_of(t.dtype)


class _BNAct(Function):
    """y = act(BatchNorm2d(x) (+ res)) on channels-last x; batch statistics in training (running stats updated)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, rmean, rvar, training, eps, momentum, res, act):
        with torch.autocast("cuda", enabled=False):
            x = x.contiguous()
            Cc = x.shape[-1]
            rows = x.numel() // Cc
            dev = x.device
            st = torch.zeros(2 * ops.NREP * Cc, dtype=torch.float64, device=dev)
            if training:
                ops.colstats(x.view(x.shape[0], -1, 1, Cc) if x.dim() != 4 else x, st[:ops.NREP * Cc], st[ops.NREP * Cc:])
            a, b, mean, rstd = (torch.empty(Cc, dtype=torch.float32, device=dev) for _ in range(4))
            ops.bn_finalize(st[:ops.NREP * Cc], st[ops.NREP * Cc:], rows, gamma.float(), beta.float(), rmean, rvar, training, a, b,
                            mean, rstd, momentum, eps)
            res = res.to(x.dtype).contiguous() if res is not None else None
            y = torch.empty_like(x)
            L.check(L.load().fmm_affine_act(x.data_ptr(), a.data_ptr(), b.data_ptr(), L.ptr(res), y.data_ptr(), rows, Cc, act,
                                            _dt(x), L.stream()), "affine_act")
        ctx.saved = (x, y if act else None, a, mean, rstd)
        ctx.cfg = (training, act, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, a, mean, rstd = ctx.saved
        training, act, has_res = ctx.cfg
        ctx.saved = None
        with torch.autocast("cuda", enabled=False):
            Cc = x.shape[-1]
            rows = x.numel() // Cc
            dy = dy.to(x.dtype).contiguous()
            S = torch.zeros(2, Cc, dtype=torch.float64, device=x.device)
            L.check(L.load().fmm_bn_act_bwd_reduce(dy.data_ptr(), L.ptr(y), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                   S[0].data_ptr(), S[1].data_ptr(), rows, Cc, act, _dt(x), L.stream()),
                    "bn_act_bwd_reduce")
            dx = torch.empty_like(x)
            dres = torch.empty_like(x) if has_res else None
            L.check(L.load().fmm_bn_act_bwd_apply(dy.data_ptr(), L.ptr(y), x.data_ptr(), a.data_ptr(), mean.data_ptr(),
                                                  rstd.data_ptr(), S[0].data_ptr(), S[1].data_ptr(), 1.0 / rows, int(training),
                                                  dx.data_ptr(), L.ptr(dres), rows, Cc, act, _dt(x), L.stream()), "bn_act_bwd_apply")
        return dx, S[1].float(), S[0].float(), None, None, None, None, None, dres, None


def bn_act(x, bn: nn.BatchNorm2d, res=None, act=0):
    y = _BNAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, bn.eps, bn.momentum, res, act)
    if bn.training:
        with torch.no_grad():
            bn.num_batches_tracked += 1
    return y


class _DWConv(Function):
    """Depthwise (k x 1) temporal convolution, groups = C, on channels-last x (N,T,V,C); w (C,1,k,1)."""

    @staticmethod
    def forward(ctx, x, w, b, k, stride, pad):
        with torch.autocast("cuda", enabled=False):
            x = x.contiguous()
            N, T, V, Cc = x.shape
            To = (T + 2 * pad - k) // stride + 1
            wf = w.float().reshape(Cc, k).contiguous()
            out = torch.empty(N, To, V, Cc, dtype=x.dtype, device=x.device)
            L.check(L.load().fmm_dwconv_fwd(x.data_ptr(), wf.data_ptr(), L.ptr(b.float().contiguous() if b is not None else None),
                                            out.data_ptr(), N, T, To, V, Cc, k, stride, pad, _dt(x), L.stream()), "dwconv_fwd")
        ctx.saved = (x, wf)
        ctx.cfg = (k, stride, pad, To, b is not None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, wf = ctx.saved
        k, stride, pad, To, has_b = ctx.cfg
        ctx.saved = None
        with torch.autocast("cuda", enabled=False):
            N, T, V, Cc = x.shape
            dy = dy.to(x.dtype).contiguous()
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                L.check(L.load().fmm_dwconv_bwd_data(dy.data_ptr(), wf.data_ptr(), dx.data_ptr(), N, T, To, V, Cc, k, stride, pad,
                                                     _dt(x), L.stream()), "dwconv_bwd_data")
            dw = torch.zeros(Cc, k, dtype=torch.float32, device=x.device)
            db = torch.zeros(Cc, dtype=torch.float32, device=x.device)
            L.check(L.load().fmm_dwconv_bwd_weight(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, T, To, V, Cc, k,
                                                   stride, pad, _dt(x), L.stream()), "dwconv_bwd_weight")
        return dx, dw.view(Cc, 1, k, 1), (db if has_b else None), None, None, None


class _JointMix(Function):
    """y[n,t,w,c] = sum_v Ae[v,w] x[n,t,v,c]  (``einsum('nctv,cvw->nctw')`` with the uniform (1,V,V) adjacency, :141)."""

    @staticmethod
    def forward(ctx, x, Ae):
        with torch.autocast("cuda", enabled=False):
            x = x.contiguous()
            N, T, V, Cc = x.shape
            Aq = Ae.to(x.dtype).contiguous()
            y = torch.empty_like(x)
            bgemm(Aq, 0, (0, 0, 1, V, 0, 0), x, 0, (V * Cc, 0, 1, Cc, 0, 0), y, 0, (V * Cc, 0, Cc, 1), (N * T, 1), V, Cc, (V, 1, 1))
        ctx.saved = (x, Aq)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Aq = ctx.saved
        ctx.saved = None
        with torch.autocast("cuda", enabled=False):
            N, T, V, Cc = x.shape
            dy = dy.to(x.dtype).contiguous()
            dx = torch.empty_like(x)
            bgemm(Aq, 0, (0, 0, V, 1, 0, 0), dy, 0, (V * Cc, 0, 1, Cc, 0, 0), dx, 0, (V * Cc, 0, Cc, 1), (N * T, 1), V, Cc, (V, 1, 1))
            dA = torch.zeros(V, V, dtype=torch.float32, device=x.device)
            bgemm(x, 0, (0, 0, Cc, V * Cc, 1, 0), dy, 0, (0, 0, Cc, V * Cc, 1, 0), dA, 0, (0, 0, V, 1), (1, 1), V, V, (N * T, Cc, 1),
                  splitk=_splitk(V, V, 1, N * T * Cc))
        return dx, dA


class _Pool(Function):
    """mean over (T, V) per clip and channel: (N,T,V,C) -> (N,C) fp32."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, T, V, Cc = x.shape
        s = torch.zeros(N, Cc, dtype=torch.float32, device=x.device)
        ops.colstats(x, None, None, s)
        ctx.shape, ctx.dt = x.shape, x.dtype
        return s / (T * V)

    @staticmethod
    def backward(ctx, dp):
        N, T, V, Cc = ctx.shape
        return (dp / (T * V)).to(ctx.dt)[:, None, None, :].expand(N, T, V, Cc).contiguous()


def conv1x1(x, conv: nn.Conv2d, stride=1):
    if stride != 1:
        x = x[:, ::stride].contiguous()
    return _Linear.apply(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias, False, None)


def dwconv(x, conv: nn.Conv2d):
    return _DWConv.apply(x, conv.weight, conv.bias, conv.kernel_size[0], conv.stride[0], conv.padding[0])


# ------------------------------------------------------------------------------------------------
# DropBlock mask generators (training, keep_prob < 1): the reference's own torch draws on the channels-last tensor
# ------------------------------------------------------------------------------------------------
def _drop_ske(x, keep_prob, Ae, training):
    if not training or keep_prob == 1:
        return x
    n, t, v, c = x.shape
    a = x.detach().abs().float().mean(dim=(1, 3))
    a = a / a.sum() * a.numel()
    gamma = (1.0 - keep_prob) / (1 + 1.92)
    M = torch.matmul(torch.bernoulli(torch.clamp(a * gamma, max=1.0)), Ae.detach().float()).reshape(n, v)
    M = torch.where(M > 0.001, torch.ones_like(M), M)
    M = torch.where(M < 0.5, torch.zeros_like(M), M)
    mask = 1 - M
    return x * (mask * (mask.numel() / mask.sum())).to(x.dtype)[:, None, :, None]


def _drop_t(x, keep_prob, block_size, training):
    if not training or keep_prob == 1:
        return x
    n, t, v, c = x.shape
    a = x.detach().abs().float().mean(dim=(2, 3))
    a = a / a.sum() * a.numel()
    gamma = (1.0 - keep_prob) / block_size
    M = torch.bernoulli(torch.clamp(a * gamma, max=1.0))[:, None, :]
    Msum = F.max_pool1d(M, kernel_size=block_size, stride=1, padding=block_size // 2)[:, 0]
    mask = 1 - Msum[:, torch.randperm(t, device=x.device)]
    return x * (mask * (mask.numel() / mask.sum())).to(x.dtype)[:, :, None, None]


# ------------------------------------------------------------------------------------------------
# modules (parameter containers with the reference's attribute names)
# ------------------------------------------------------------------------------------------------
def _act_code(name):
    if name not in ("relu", "tanh", "linear", None):
        raise NotImplementedError(f"act_type {name!r}: the CUDA path implements relu / tanh / linear")
    return ACT[name]


class SpatialGraphConv(nn.Module):
    def __init__(self, in_channel, out_channel, max_graph_distance, bias, edge, A, act_type, keep_prob, block_size, num_point,
                 residual=True, **kwargs):
        super().__init__()
        self.keep_prob, self.num_point, self.block_size = keep_prob, num_point, block_size
        self.gcn = nn.Conv2d(in_channel, out_channel, 1, bias=bias)
        self.A = nn.Parameter(A.clone(), requires_grad=False)
        self.edge = nn.Parameter(torch.ones_like(self.A)) if edge else 1
        self.act = _act_code(act_type)
        self.bn = nn.BatchNorm2d(out_channel)
        self.residual = nn.Sequential(nn.Conv2d(in_channel, out_channel, 1, bias=bias), nn.BatchNorm2d(out_channel))

    def forward(self, x):
        Ae = (self.A * self.edge)[0]
        r = conv1x1(x, self.residual[0])
        g = _JointMix.apply(conv1x1(x, self.gcn), Ae)
        if self.training and self.keep_prob != 1:
            kp, bs = self.keep_prob, self.block_size
            y = _drop_t(_drop_ske(bn_act(g, self.bn), kp, Ae, True), kp, bs, True) + \
                _drop_t(_drop_ske(bn_act(r, self.residual[1]), kp, Ae, True), kp, bs, True)
            return {0: lambda t: t, 1: torch.relu, 2: torch.tanh}[self.act](y)
        return bn_act(g, self.bn, res=bn_act(r, self.residual[1]), act=self.act)


class SepTemporal_Block(nn.Module):
    def __init__(self, channel, temporal_window_size, bias, act_type, edge, A, num_point, keep_prob, block_size, expand_ratio,
                 stride=1, residual=True, **kwargs):
        super().__init__()
        if expand_ratio > 0:
            raise NotImplementedError("expand_ratio > 0 is not used by Model / main.py")
        self.keep_prob, self.num_point, self.block_size, self.stride = keep_prob, num_point, block_size, stride
        padding = (temporal_window_size - 1) // 2
        self.act = _act_code(act_type)
        self.depth_conv = nn.Sequential(nn.Conv2d(channel, channel, (temporal_window_size, 1), (stride, 1), (padding, 0),
                                                  groups=channel, bias=bias), nn.BatchNorm2d(channel))
        self.point_conv = nn.Sequential(nn.Conv2d(channel, channel, 1, bias=bias), nn.BatchNorm2d(channel))
        if stride == 1:
            self.residual = nn.Identity()
        else:
            self.residual = nn.Sequential(nn.Conv2d(channel, channel, 1, (stride, 1), bias=bias), nn.BatchNorm2d(channel))
        self.A = nn.Parameter(A.clone(), requires_grad=False)
        self.edge = nn.Parameter(torch.ones_like(self.A)) if edge else 1

    def forward(self, x):
        res = x if self.stride == 1 else bn_act(conv1x1(x, self.residual[0], self.stride), self.residual[1])
        d = bn_act(dwconv(x, self.depth_conv[0]), self.depth_conv[1], act=self.act)
        p = conv1x1(d, self.point_conv[0])
        if self.training and self.keep_prob != 1:
            kp, bs = self.keep_prob, self.block_size
            Ae = (self.A * self.edge)[0]
            y = _drop_t(_drop_ske(bn_act(p, self.point_conv[1]), kp, Ae, True), kp, bs, True) + \
                _drop_t(_drop_ske(res, kp, Ae, True), kp, bs, True)
            return {0: lambda t: t, 1: torch.relu, 2: torch.tanh}[self.act](y)
        return bn_act(p, self.point_conv[1], res=res, act=self.act)


class _DWS(nn.Module):
    def __init__(self, in_features, out_features, k):
        super().__init__()
        self.seq = nn.Sequential(nn.Conv2d(in_features, in_features, kernel_size=(k, 1), padding=((k - 1) // 2, 0), groups=in_features),
                                 nn.BatchNorm2d(in_features), nn.LeakyReLU(), nn.Conv2d(in_features, out_features, kernel_size=1),
                                 nn.BatchNorm2d(out_features))
        self.relu = nn.ReLU(inplace=False)

    def forward(self, x):
        y = bn_act(dwconv(x, self.seq[0]), self.seq[1], act=ACT["leakyrelu01"])
        return bn_act(conv1x1(y, self.seq[3]), self.seq[4], act=ACT["relu"])


class DepthWiseSeparableConv_3x1_1x1(_DWS):
    def __init__(self, in_features, out_features):
        super().__init__(in_features, out_features, 3)


class DepthWiseSeparableConv_1x1_1x1(_DWS):
    def __init__(self, in_features, out_features):
        super().__init__(in_features, out_features, 1)


class Sep_TCN(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        middle = int((out_features - in_features) / 2) + in_features
        self.sep31 = DepthWiseSeparableConv_3x1_1x1(in_features, middle)
        self.sep11 = DepthWiseSeparableConv_1x1_1x1(middle, out_features)
        self.shortcut = nn.Conv2d(in_features, out_features, kernel_size=1)

    def forward(self, x):
        return self.sep11(self.sep31(x)) + conv1x1(x, self.shortcut)


class cnn1x1(nn.Module):
    def __init__(self, dim1=3, dim2=3, bias=True):
        super().__init__()
        self.cnn = nn.Conv2d(dim1, dim2, kernel_size=1, bias=bias)


class embed(nn.Module):
    def __init__(self, dim, dim1, att_type=None, norm=False, bias=False):
        super().__init__()
        if norm:
            raise NotImplementedError("embed(norm=True) is not used by Model")
        self.cnn = nn.Sequential(cnn1x1(dim, dim1, bias=bias), nn.ReLU())

    def forward(self, x):
        c = self.cnn[0].cnn
        W = c.weight.view(c.out_channels, c.in_channels)
        pad = (-c.in_channels) % 8
        if pad:   # 3 / 2 input channels: zero-pad to 8 so the GEMM fetches 16-byte vectors (the pad columns meet zero weights)
            x, W = F.pad(x, (0, pad)), F.pad(W, (0, pad))
        return _Linear.apply(x, W, c.bias, True, None)


class Classification_Module(nn.Module):
    def __init__(self, in_features, numclass):
        super().__init__()
        self.seq = nn.Sequential(nn.Linear(in_features, 128), nn.LeakyReLU(), nn.LayerNorm(128), nn.LeakyReLU(), nn.Dropout(0.2),
                                 nn.Linear(128, numclass))

    def forward(self, x):
        return self.seq(x)


class Model(nn.Module):
    """``Model(num_class, num_point, max_frame, graph, bias, edge, block_size, embed_dim=32, n_stage=2, act_type='relu')``
    (musa_model.py:492-545); ``forward(x[N,3,T,V]) -> (N,num_class)``."""

    _sep_tcn_tail = True   # Ablation (musa_model.py:593-686): the same two streams without the closing Sep_TCN

    def __init__(self, num_class, num_point, max_frame, graph, bias, edge, block_size, embed_dim=32, n_stage=2, act_type="relu"):
        super().__init__()
        self.num_classes = num_class
        tw, mgd, keep_prob = 3, 2, 0.9
        A = torch.as_tensor(graph.A, dtype=torch.float32)
        if A.shape[0] != 1:
            raise NotImplementedError("the joint mixing einsum('nctv,cvw->nctw') needs a (1,V,V) adjacency (strategy='uniform')")
        self.joint_embed_pos = embed(3, embed_dim, att_type="stja", norm=False, bias=bias)
        self.joint_embed_mos = embed(2, embed_dim, att_type="stja", norm=False, bias=bias)
        pos, mot = [], []
        for _ in range(n_stage):
            for lst in (pos, mot):
                lst += [SpatialGraphConv(embed_dim, embed_dim * 2, mgd, bias, edge, A, act_type, keep_prob, block_size, num_point),
                        SepTemporal_Block(embed_dim * 2, tw, bias, act_type, edge, A, num_point, keep_prob, block_size, 0, stride=1),
                        SepTemporal_Block(embed_dim * 2, tw + 2, bias, act_type, edge, A, num_point, keep_prob, block_size, 0, stride=2)]
            embed_dim *= 2
        if self._sep_tcn_tail:
            pos += [Sep_TCN(embed_dim, embed_dim * 2)]
            mot += [Sep_TCN(embed_dim, embed_dim * 2)]
            embed_dim *= 2
        self.stream_pos = nn.Sequential(*pos)
        self.stream_mot = nn.Sequential(*mot)
        self.fc = Classification_Module(embed_dim * 2 + 3, num_class)       # :545 (embed_dim*4+3) / :643 (embed_dim*2+3)
        self.compute_dtype = None

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("fall_multimodal_b200.musa.Model runs on CUDA (sm_100a) only; there is no CPU fallback")
        dt = _compute_dtype(self)
        with torch.autocast("cuda", enabled=False):
            mot = x[:, :2, :-1] - x[:, :2, 1:]                                    # :549
            pts_cl = x.permute(0, 2, 3, 1).to(dt).contiguous()                    # (N,T,V,3)
            mot_cl = mot.permute(0, 2, 3, 1).to(dt).contiguous()
        out = self.stream_pos(self.joint_embed_pos(pts_cl))
        out2 = self.stream_mot(self.joint_embed_mos(mot_cl))
        with torch.autocast("cuda", enabled=False):
            feat = torch.cat([_Pool.apply(out), _Pool.apply(out2), x.float().flatten(2).mean(2)], dim=-1)   # :574-585
            y = self.fc(feat)
        return y.to(dt) if dt == torch.bfloat16 else y


class Ablation(Model):
    """``Ablation(...)`` (musa_model.py:593-686): `Model` without the two closing ``Sep_TCN`` stages - the pooled features are
    the outputs of the last ``SepTemporal_Block`` of each stream, the classifier sees ``2 * C + 3`` features. Same constructor
    arguments, state_dict keys and forward signature as the reference class."""

    _sep_tcn_tail = False
