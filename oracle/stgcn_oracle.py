"""ORACLE — test infrastructure only (never imported by the product path).

Plain-PyTorch CPU restatement of the hot path of musaru/Fall_Multimodal: the GSTCAN trunk, the
sensor branches (bi-LSTM, CNN1D) and the late-fusion heads, written as pure functions over a
``state_dict`` so that gradients come from torch autograd. Each function cites the reference
file:line it follows (paths relative to /root/reference; F2 = Fall_2_Spatial_Temporal_SR).

Pinning: the reference publishes no numerical golden vectors, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
``oracle/make_golden.py`` (imports the unmodified reference modules) and committed under
``tests/golden/``. ``tests/test_oracle.py`` replays them. Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this file.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# Graph (F2/Model/graph.py:6-126)
# ------------------------------------------------------------------------------------------------
LAYOUTS = {
    # name: (num_node, neighbor_link, center)                                   graph.py:33-55
    "coco_cut": (14, [(6, 4), (4, 2), (2, 13), (13, 1), (5, 3), (3, 1), (12, 10), (10, 8), (8, 2),
                      (11, 9), (9, 7), (7, 1), (13, 0)], 13),
    "coco_mmpose": (18, [(0, 1), (1, 3), (0, 2), (2, 4), (17, 0), (17, 6), (6, 8), (8, 10), (17, 5),
                         (5, 7), (7, 9), (17, 12), (12, 14), (14, 16), (17, 11), (11, 13), (13, 15)], 17),
    # 33-node layout of BASELINE.json (not in the reference): MediaPipe-Pose
    # landmarks with its 35 POSE_CONNECTIONS, centred on the nose. Fed as the same A to both sides.
    "mediapipe33": (33, [(0, 1), (1, 2), (2, 3), (3, 7), (0, 4), (4, 5), (5, 6), (6, 8), (9, 10),
                         (11, 12), (11, 13), (13, 15), (15, 17), (15, 19), (15, 21), (17, 19),
                         (12, 14), (14, 16), (16, 18), (16, 20), (16, 22), (18, 20), (11, 23),
                         (12, 24), (23, 24), (23, 25), (24, 26), (25, 27), (26, 28), (27, 29),
                         (28, 30), (29, 31), (30, 32), (27, 31), (28, 32)], 0),
    # 25-node NTU layout used by config 4 (MF3/model/musa_model.py:252-263, 1-based there)
    "ntu-rgb+d": (25, [(i - 1, j - 1) for (i, j) in
                       [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21),
                        (10, 9), (11, 10), (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1),
                        (18, 17), (19, 18), (20, 19), (22, 23), (23, 8), (24, 25), (25, 12)]], 20),
}


def hop_distance(num_node, edge, max_hop=1):
    """graph.py:103-115 — hop count through powers of the adjacency matrix."""
    A = np.zeros((num_node, num_node))
    for i, j in edge:
        A[j, i] = 1
        A[i, j] = 1
    hop = np.zeros((num_node, num_node)) + np.inf
    powers = [np.linalg.matrix_power(A, d) for d in range(max_hop + 1)]
    arrive = np.stack(powers) > 0
    for d in range(max_hop, -1, -1):
        hop[arrive[d]] = d
    return hop


def normalize_digraph(A):
    """graph.py:118-126 — column normalisation A D^-1."""
    Dl = A.sum(0)
    Dn = np.zeros_like(A)
    for i in range(A.shape[0]):
        if Dl[i] > 0:
            Dn[i, i] = Dl[i] ** (-1)
    return A @ Dn


def build_adjacency(layout="coco_cut", strategy="uniform", max_hop=1, dilation=1):
    """graph.py:20-100 — returns A (K, V, V) float64."""
    num_node, neighbor, center = LAYOUTS[layout]
    edge = [(i, i) for i in range(num_node)] + list(neighbor)
    hop = hop_distance(num_node, edge, max_hop)
    valid_hop = range(0, max_hop + 1, dilation)
    adj = np.zeros((num_node, num_node))
    for h in valid_hop:
        adj[hop == h] = 1
    nadj = normalize_digraph(adj)
    if strategy == "uniform":
        return nadj[None].copy()
    if strategy == "distance":
        A = np.zeros((len(valid_hop), num_node, num_node))
        for i, h in enumerate(valid_hop):
            A[i][hop == h] = nadj[hop == h]
        return A
    if strategy == "spatial":
        out = []
        for h in valid_hop:
            root = np.zeros((num_node, num_node))
            close = np.zeros((num_node, num_node))
            further = np.zeros((num_node, num_node))
            for i in range(num_node):
                for j in range(num_node):
                    if hop[j, i] == h:
                        if hop[j, center] == hop[i, center]:
                            root[j, i] = nadj[j, i]
                        elif hop[j, center] > hop[i, center]:
                            close[j, i] = nadj[j, i]
                        else:
                            further[j, i] = nadj[j, i]
            if h == 0:
                out.append(root)
            else:
                out.append(root + close)
                out.append(further)
        return np.stack(out)
    raise ValueError("This strategy is not supported!")


# ------------------------------------------------------------------------------------------------
# GSTCAN (F2/Model/stgcan.py)
# ------------------------------------------------------------------------------------------------
BLOCK_PLAN = [(None, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 128, 2, True),
              (128, 128, 1, True), (128, 256, 2, True), (256, 256, 1, True)]  # stgcan.py:182-194


def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5, update=None):
    """nn.BatchNorm{1,2}d in train (batch stats) or eval (running stats) mode."""
    w, b = sd[prefix + "weight"], sd[prefix + "bias"]
    rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    if training:
        rm_new, rv_new = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm_new, rv_new, w, b, True, momentum, eps)
        if update is not None:
            update[prefix + "running_mean"] = rm_new
            update[prefix + "running_var"] = rv_new
        return y
    return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)


def stgcan_param_shapes(in_channels, V, K, num_class):
    """state_dict keys/shapes of STGCAN (stgcan.py:166-208), in registration order."""
    shapes = {"A": (K, V, V)}
    def bn(p, c):
        shapes[p + "weight"] = (c,); shapes[p + "bias"] = (c,)
        shapes[p + "running_mean"] = (c,); shapes[p + "running_var"] = (c,)
        shapes[p + "num_batches_tracked"] = ()
    bn("data_bn.", in_channels * V)
    for i, (cin, cout, stride, res) in enumerate(BLOCK_PLAN):
        cin = in_channels if cin is None else cin
        p = f"st_gcan_networks.{i}."
        shapes[p + "gcn.conv.weight"] = (K * cout, cin, 1, 1)
        shapes[p + "gcn.conv.bias"] = (K * cout,)
        bn(p + "tcn.0.", cout)
        shapes[p + "tcn.2.weight"] = (cout, cout, 9, 1)
        shapes[p + "tcn.2.bias"] = (cout,)
        bn(p + "tcn.3.", cout)
        if res and (cin != cout or stride != 1):
            shapes[p + "residual.0.weight"] = (cout, cin, 1, 1)
            shapes[p + "residual.0.bias"] = (cout,)
            bn(p + "residual.1.", cout)
        c4 = int(cout / 4)
        a = p + "channel_attention_module.atten."
        shapes[a + "1.weight"] = (c4, cout, 1, 1); shapes[a + "1.bias"] = (c4,)
        bn(a + "2.", c4)
        shapes[a + "4.weight"] = (cout, c4, 1, 1); shapes[a + "4.bias"] = (cout,)
    for i in range(len(BLOCK_PLAN)):
        shapes[f"edge_importance.{i}"] = (K, V, V)
    if num_class is not None:
        shapes["cls.weight"] = (num_class, 256, 1, 1)
        shapes["cls.bias"] = (num_class,)
    return shapes


def graph_conv(x, A, w, b):
    """GraphConvolution.forward, stgcan.py:50-56 (1x1 conv to K*C channels, k-major split, einsum)."""
    K = A.shape[0]
    y = F.conv2d(x, w, b)
    n, kc, t, v = y.shape
    y = y.view(n, K, kc // K, t, v)
    return torch.einsum("nkctv,kvw->nctw", y, A).contiguous()


def channel_attention(x, sd, prefix, training, update=None):
    """Channel_Attention, stgcan.py:59-74: x * sigmoid(W2 relu(BN(W1 avgpool(x))))."""
    a = prefix + "atten."
    s = F.adaptive_avg_pool2d(x, (1, 1))
    s = F.conv2d(s, sd[a + "1.weight"], sd[a + "1.bias"])
    s = _bn(s, sd, a + "2.", training, update=update)
    s = F.relu(s)
    s = F.conv2d(s, sd[a + "4.weight"], sd[a + "4.bias"])
    return x * torch.sigmoid(s)


def _relu(x, masks, site):
    """ReLU; with ``masks`` the 0/1 decision is taken from the given tensor instead of sign(x).

    Used by the parity tests to compare gradients GIVEN IDENTICAL ReLU DECISIONS: two correct fp32
    implementations disagree on sign(x) for the handful of elements with |x| ~ 1e-7*max, and each
    such flip moves a gradient by one whole element, which no tolerance on the sums can absorb.
    """
    if masks is None:
        return F.relu(x)
    return x * masks[site].to(x.dtype)


def st_gcan_block(x, A, sd, prefix, stride, residual, training, attention=True, update=None, masks=None, site=0):
    """st_gcan.forward, stgcan.py:138-144: relu(CA(tcn(gcn(x, A))) + res)."""
    if not residual:
        res = 0
    elif (prefix + "residual.0.weight") in sd:
        res = F.conv2d(x, sd[prefix + "residual.0.weight"], sd[prefix + "residual.0.bias"], stride=(stride, 1))
        res = _bn(res, sd, prefix + "residual.1.", training, update=update)
    else:
        res = x
    y = graph_conv(x, A, sd[prefix + "gcn.conv.weight"], sd[prefix + "gcn.conv.bias"])
    y = _bn(y, sd, prefix + "tcn.0.", training, update=update)            # stgcan.py:112
    y = _relu(y, masks, site)                                             # :113
    y = F.conv2d(y, sd[prefix + "tcn.2.weight"], sd[prefix + "tcn.2.bias"], stride=(stride, 1), padding=(4, 0))  # :114-118
    y = _bn(y, sd, prefix + "tcn.3.", training, update=update)            # :119 (Dropout p=0 is identity)
    if attention:
        y = channel_attention(y, sd, prefix + "channel_attention_module.", training, update=update)
    return _relu(y + res, masks, site + 1)


def stgcan_forward(sd, skel, training=True, update=None, block_key="st_gcan_networks", masks=None):
    """STGCAN.forward, stgcan.py:210-228. ``sd``: state_dict-like mapping of tensors.

    Returns (N, num_class) logits, or the (N, 256) pooled feature when the dict has no ``cls.*``.
    """
    N, C, T, V = skel.shape
    x = skel.permute(0, 3, 1, 2).contiguous().view(N, V * C, T)           # :213-214
    x = _bn(x, sd, "data_bn.", training, update=update)                   # :215
    x = x.view(N, V, C, T).permute(0, 2, 3, 1).contiguous().view(N, C, T, V)  # :216-218
    A = sd["A"]
    for i, (cin, cout, stride, res) in enumerate(BLOCK_PLAN):
        x = st_gcan_block(x, A * sd[f"edge_importance.{i}"], sd, f"{block_key}.{i}.", stride, res,
                          training, update=update, masks=masks, site=2 * i)  # :221-222
    x = F.avg_pool2d(x, x.shape[2:])                                      # :224
    if "cls.weight" in sd:
        x = F.conv2d(x, sd["cls.weight"], sd["cls.bias"])                 # :225
    return x.view(x.size(0), -1)                                          # :226


# ------------------------------------------------------------------------------------------------
# Sensor branches
# ------------------------------------------------------------------------------------------------
def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM (cuDNN semantics): gates i,f,g,o; zero initial state."""
    N, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = x[:, t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def bilstm_forward(sd, sensor, training=True, feature="mean", update=None):
    """BiLSTM.forward, F2/Model/bilstm.py:41-58."""
    fwd = lstm_direction(sensor, sd["lstm1.weight_ih_l0"], sd["lstm1.weight_hh_l0"],
                         sd["lstm1.bias_ih_l0"], sd["lstm1.bias_hh_l0"], False)
    bwd = lstm_direction(sensor, sd["lstm1.weight_ih_l0_reverse"], sd["lstm1.weight_hh_l0_reverse"],
                         sd["lstm1.bias_ih_l0_reverse"], sd["lstm1.bias_hh_l0_reverse"], True)
    out = torch.cat([fwd, bwd], dim=2)                                    # :48
    out = out[:, -1, :] if feature == "last" else out.mean(dim=1)         # :52-55
    out = _bn(out, sd, "batchnorm.", training, update=update)             # :56
    w = F.linear(out, sd["channelattention.attention.0.weight"], sd["channelattention.attention.0.bias"])
    w = F.relu(w)
    w = torch.sigmoid(F.linear(w, sd["channelattention.attention.2.weight"], sd["channelattention.attention.2.bias"]))
    out = out * w                                                         # bilstm.py:16-19
    return F.linear(out, sd["fc.1.weight"], sd["fc.1.bias"])              # :58


def cnn1d_forward(sd, x, training=True, update=None):
    """Notebook CNN1D.forward (GSTCAN_HAR_conv_10kfold.ipynb#cell2:L6-27); x is (N, Cin, L)."""
    for layer in ("layer1", "layer2"):
        x = F.conv1d(x, sd[f"{layer}.0.weight"], sd[f"{layer}.0.bias"], padding=2)
        x = _bn(x, sd, f"{layer}.1.", training, update=update)
        x = F.max_pool1d(F.relu(x), 2)
    return x


def cnn_bilstm_forward(sd, x, training=True, feature="mean", update=None):
    """Notebook CNN_BiLSTM.forward (GSTCAN_HAR_conv_10kfold.ipynb#cell2:L85-100); x is (N, L, Cin)."""
    feat = cnn1d_forward(_sub(sd, "cnn."), x.permute(0, 2, 1), training, _PrefixDict(update, "cnn."))
    return bilstm_forward(_sub(sd, "bilstm."), feat.permute(0, 2, 1), training, feature, _PrefixDict(update, "bilstm."))


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def two_stream_bilstm_forward(sd, skel, sensor, training=True, update=None):
    """TwoStreamSTGCAN_BiLSTM.forward, F2/Model/combination.py:37-46 (raw logits)."""
    mot = skel[:, :2, 1:] - skel[:, :2, :-1]                              # :39
    pts = stgcan_forward(_sub(sd, "stgcan_1."), skel, training, _PrefixDict(update, "stgcan_1."))
    mo = stgcan_forward(_sub(sd, "stgcan_2."), mot, training, _PrefixDict(update, "stgcan_2."))
    sen = bilstm_forward(_sub(sd, "lstm."), sensor, training, "mean", _PrefixDict(update, "lstm."))
    x = torch.cat((pts, mo, sen), dim=-1)                                 # :45
    return F.linear(x, sd["fc.weight"], sd["fc.bias"])                    # :46


def two_stream_cnn_forward(sd, skel, sensor, training=True, update=None):
    """BASELINE config 2: pts + motion GSTCAN trunks and the notebook CNN1D sensor branch, late
    fusion by concat -> Linear (the notebook TwoStreamSpatialTemporalGraph pattern,
    GSTCAN_HAR_conv_10kfold.ipynb#cell1:L362-416, with CNN1D's flattened feature map as the third
    input; raw logits). State-dict prefixes follow combination.py:31-35 with ``cnn.`` for the branch.
    """
    mot = skel[:, :2, 1:] - skel[:, :2, :-1]
    pts = stgcan_forward(_sub(sd, "stgcan_1."), skel, training, _PrefixDict(update, "stgcan_1."))
    mo = stgcan_forward(_sub(sd, "stgcan_2."), mot, training, _PrefixDict(update, "stgcan_2."))
    sen = cnn1d_forward(_sub(sd, "cnn."), sensor.permute(0, 2, 1), training, _PrefixDict(update, "cnn."))
    x = torch.cat((pts, mo, sen.flatten(1)), dim=-1)
    return F.linear(x, sd["fc.weight"], sd["fc.bias"])


def three_stream_forward(sd, skel, parents, training=True, update=None):
    """BASELINE config 3 (bone stream NOT in the reference, SURVEY D1: parity unpinned for it):
    joints + motion as combination.py:9-25, plus bones = joint - joint[parent]; cat -> Linear(768, C)."""
    mot = skel[:, :2, 1:] - skel[:, :2, :-1]
    bone = skel - skel.index_select(3, parents)
    f1 = stgcan_forward(_sub(sd, "stgcan_1."), skel, training, _PrefixDict(update, "stgcan_1."))
    f2 = stgcan_forward(_sub(sd, "stgcan_2."), mot, training, _PrefixDict(update, "stgcan_2."))
    f3 = stgcan_forward(_sub(sd, "stgcan_3."), bone, training, _PrefixDict(update, "stgcan_3."))
    return F.linear(torch.cat((f1, f2, f3), dim=-1), sd["fc.weight"], sd["fc.bias"])


class _PrefixDict:
    """Writes ``update[prefix + k] = v`` into a parent dict (running-stat updates of sub-modules)."""

    def __init__(self, parent, prefix):
        self.parent, self.prefix = parent, prefix

    def __setitem__(self, k, v):
        if self.parent is not None:
            self.parent[self.prefix + k] = v

    def __bool__(self):
        return self.parent is not None


# ------------------------------------------------------------------------------------------------
# Deterministic parameter fill + synthetic inputs + the train step (F2/main.py:104-132)
# ------------------------------------------------------------------------------------------------
from synth import fill_state_dict, synthetic_batch  # noqa: E402,F401  (generators live in /synth.py)


def soft_ce(logits, target):
    """torch.nn.CrossEntropyLoss with probability targets, mean over the batch (main.py:113,280)."""
    return -(target * F.log_softmax(logits.float(), dim=-1)).sum(-1).mean()
