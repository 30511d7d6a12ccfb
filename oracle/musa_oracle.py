"""ORACLE (test infrastructure): functional torch restatement of the musa ``Model`` that Multimodal_Fall3/main.py trains
(Multimodal_Fall3/model/musa_model.py: embed :384-406, SpatialGraphConv :101-146, SepTemporal_Block :148-199, Sep_TCN
:461-474, Classification_Module :476-490, Model.forward :547-591). DropBlock (:39-99) and Dropout are identities here
(eval mode, or keep_prob = 1 / p = 0), which is the configuration the fixtures are generated in. Pinned against the
unmodified reference module by tests/golden/musa_*.pt (oracle/make_golden.py musa). Only tests/ may import this file."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn(x, sd, p, training, eps=1e-5):
    if training:
        return F.batch_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, eps)


def _act(x, name):
    return {"tanh": torch.tanh, "relu": torch.relu, "linear": lambda t: t}[name](x)


def spatial_graph_conv(x, sd, p, training, act):
    res = _bn(F.conv2d(x, sd[p + "residual.0.weight"], sd[p + "residual.0.bias"]), sd, p + "residual.1.", training)    # :127
    g = F.conv2d(x, sd[p + "gcn.weight"], sd[p + "gcn.bias"])                                                          # :129
    g = torch.einsum("nctv,cvw->nctw", g, (sd[p + "A"] * sd[p + "edge"]).expand(g.shape[1], -1, -1))                   # :141
    return _act(_bn(g, sd, p + "bn.", training) + res, act)                                                            # :144-146


def sep_temporal_block(x, sd, p, training, act, k, stride):
    C = x.shape[1]
    if stride == 1:
        res = x                                                                                                         # :176
    else:
        res = _bn(F.conv2d(x, sd[p + "residual.0.weight"], sd[p + "residual.0.bias"], stride=(stride, 1)), sd, p + "residual.1.", training)
    d = F.conv2d(x, sd[p + "depth_conv.0.weight"], sd[p + "depth_conv.0.bias"], stride=(stride, 1), padding=((k - 1) // 2, 0), groups=C)
    d = _act(_bn(d, sd, p + "depth_conv.1.", training), act)                                                           # :195
    pt = _bn(F.conv2d(d, sd[p + "point_conv.0.weight"], sd[p + "point_conv.0.bias"]), sd, p + "point_conv.1.", training)
    return _act(pt + res, act)                                                                                          # :198-199


def _dws(x, sd, p, training, k):
    C = x.shape[1]
    y = F.conv2d(x, sd[p + "seq.0.weight"], sd[p + "seq.0.bias"], padding=((k - 1) // 2, 0), groups=C)
    y = F.leaky_relu(_bn(y, sd, p + "seq.1.", training))
    y = _bn(F.conv2d(y, sd[p + "seq.3.weight"], sd[p + "seq.3.bias"]), sd, p + "seq.4.", training)
    return torch.relu(y)                                                                                                # :436 / :456


def sep_tcn(x, sd, p, training):
    res = F.conv2d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"])
    return _dws(_dws(x, sd, p + "sep31.", training, 3), sd, p + "sep11.", training, 1) + res                            # :469-474


def stream(x, sd, p, training, act, n_stage, tail=True):
    i = 0
    for _ in range(n_stage):
        x = spatial_graph_conv(x, sd, f"{p}{i}.", training, act)
        x = sep_temporal_block(x, sd, f"{p}{i + 1}.", training, act, 3, 1)
        x = sep_temporal_block(x, sd, f"{p}{i + 2}.", training, act, 5, 2)
        i += 3
    return sep_tcn(x, sd, f"{p}{i}.", training) if tail else x


def musa_forward(sd, x, training=False, act="tanh", n_stage=1, ablation=False):
    """Model.forward (:547-591) / Ablation.forward (:646-686: no closing Sep_TCN); x (N,3,T,V) -> (N,num_class)."""
    N = x.shape[0]
    mot = x[:, :2, :-1] - x[:, :2, 1:]                                                                                  # :549
    pos = torch.relu(F.conv2d(x, sd["joint_embed_pos.cnn.0.cnn.weight"], sd["joint_embed_pos.cnn.0.cnn.bias"]))
    mo = torch.relu(F.conv2d(mot, sd["joint_embed_mos.cnn.0.cnn.weight"], sd["joint_embed_mos.cnn.0.cnn.bias"]))
    out = stream(pos, sd, "stream_pos.", training, act, n_stage, tail=not ablation)
    out2 = stream(mo, sd, "stream_mot.", training, act, n_stage, tail=not ablation)
    feat = torch.cat([out.flatten(2).mean(2), out2.flatten(2).mean(2), x.flatten(2).mean(2)], dim=-1)                  # :574-585
    h = F.leaky_relu(F.linear(feat, sd["fc.seq.0.weight"], sd["fc.seq.0.bias"]))
    h = F.leaky_relu(F.layer_norm(h, (h.shape[-1],), sd["fc.seq.2.weight"], sd["fc.seq.2.bias"]))
    return F.linear(h, sd["fc.seq.5.weight"], sd["fc.seq.5.bias"])                                                      # Dropout: identity here


def fill_musa(shapes, seed=0):
    import math
    import zlib
    import numpy as np
    sd = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) % (2 ** 31))
        if k.endswith(".A"):
            continue
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith(".edge"):
            sd[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif len(shp) == 1:
            sd[k] = torch.rand(shp, generator=g) + 0.5
        else:
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(int(np.prod(shp[1:])))
    return sd
