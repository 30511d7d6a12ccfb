"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's
TRAGCN family in plain torch ops.

  EmbGCN            /root/reference/EmbGCN.py:59-89      (emb_gcn_invariants + emb_gcn)
  GRU graph cell    /root/reference/GRU.py:8-29          (gru_cell)
  AVWDCRNN scan     /root/reference/TRAGCN.py:150-169    (encoder)
  PositionalEncoding / Transform / transformer_layer   /root/reference/TA.py:22-108
  TARGCN + head     /root/reference/TRAGCN.py:177-224    (targcn_forward)

The only liberty taken is hoisting the per-call recomputation of the loop-invariant tensors (adaptive
adjacency, per-node weights) out of the time loop, which does not change a single arithmetic
operation. Pinned against the reference itself: tests/golden/targcn_*.pt are produced by
oracle/make_golden.py from the unmodified reference modules and checked in tests/test_oracle.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def sym_norm_adj(adj: torch.Tensor) -> torch.Tensor:
    """EmbGCN.py:13-26 + the constructor's softmax (:63-64; implicit dim of a 2-D tensor = 1)."""
    W = adj.detach().cpu().double().numpy()
    n = W.shape[0]
    W = W + 0.5 * np.identity(n)
    D = np.diag(1.0 / np.sum(W, axis=1))
    out = np.dot(np.dot(np.sqrt(D), W), np.sqrt(D))
    return F.softmax(torch.from_numpy(out).to(torch.float32), dim=1)


from synth import fill_targcn, positional_encoding, synthetic_clips  # noqa: E402,F401  (generators live in /synth.py)


def emb_gcn_invariants(sd, prefix, E, sym):
    """Everything in EmbGCN.forward that does not depend on x (EmbGCN.py:73-83)."""
    V = E.shape[0]
    supports = F.softmax(F.relu(E @ E.t()), dim=1) + torch.eye(V, dtype=E.dtype, device=E.device)      # :73-74
    colscale = torch.softmax(sym.to(E.device, E.dtype), dim=-1).sum(0)                            # :77 "nm,bmc->bmc" sums n
    weights = torch.einsum("nd,dio->nio", E, sd[prefix + "weights_pool"])              # :80
    bias = E @ sd[prefix + "bias_pool"]                                                 # :81
    return supports, colscale, weights, bias


def emb_gcn_nogate(x, E, sd, prefix):
    """EmbGCN_noGate.forward (EmbGCN.py:99-109): the per-node product alone."""
    V = E.shape[0]
    supports = F.softmax(F.relu(E @ E.t()), dim=1) + torch.eye(V, dtype=E.dtype, device=E.device)
    weights = torch.einsum("nd,dio->nio", E, sd[prefix + "weights_pool"])
    bias = E @ sd[prefix + "bias_pool"]
    return torch.einsum("bni,nio->bno", torch.einsum("nm,bmc->bnc", supports, x), weights) + bias


def emb_gcn_linear(x, E, sd, prefix):
    """EmbGCN_linear.forward (EmbGCN.py:116-123): one shared Linear on the adjacency-mixed input."""
    V = E.shape[0]
    supports = F.softmax(F.relu(E @ E.t()), dim=1) + torch.eye(V, dtype=E.dtype, device=E.device)
    return F.linear(torch.einsum("nm,bmc->bnc", supports, x), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])


def emb_gcn(x, inv, sd, prefix):
    if isinstance(inv, tuple) and len(inv) == 2:          # (variant name, E): the import-time variants of GRU.py:3-6
        return (emb_gcn_nogate if inv[0] == "noGate" else emb_gcn_linear)(x, inv[1], sd, prefix)
    supports, colscale, weights, bias = inv
    x_static = F.linear(colscale[None, :, None] * x, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])  # :77-78
    x_g = torch.einsum("nm,bmc->bnc", supports, x)                                     # :83
    x_gconv = torch.einsum("bni,nio->bno", x_g, weights) + bias                        # :86
    return x_gconv + torch.sigmoid(x_static) * x_static                                 # :88


def gru_cell(x, state, inv_g, inv_u, sd, prefix, H):
    """GRU.py:17-26."""
    zr = torch.sigmoid(emb_gcn(torch.cat((x, state), -1), inv_g, sd, prefix + "gate."))
    z, r = torch.split(zr, H, dim=-1)
    hc = torch.tanh(emb_gcn(torch.cat((x, r * state), -1), inv_u, sd, prefix + "update."))
    return z * state + (1 - z) * hc


def transform(x, sd, prefix):
    """TA.Transform.forward (TA.py:40-69); x (B,T,V,C). The convolutions treat TIME as the channel axis."""
    c = x.shape[-1]
    q = F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]).permute(0, 2, 1, 3)   # (B,V,T,C-2)
    k = F.conv2d(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]).permute(0, 2, 3, 1)   # (B,V,C-2,T)
    v = F.linear(x, sd[prefix + "vff.weight"], sd[prefix + "vff.bias"]).permute(0, 2, 1, 3)       # (B,V,T,C)
    A = torch.softmax(torch.matmul(q, k) / (c ** 0.5), -1)
    val = torch.matmul(A, v).permute(0, 2, 1, 3) + x
    val = F.layer_norm(val, (c,), sd[prefix + "ln.weight"], sd[prefix + "ln.bias"])
    y = F.linear(F.relu(F.linear(val, sd[prefix + "ff.0.weight"], sd[prefix + "ff.0.bias"])),
                 sd[prefix + "ff.2.weight"], sd[prefix + "ff.2.bias"]) + val
    return F.layer_norm(y, (c,), sd[prefix + "lnff.weight"], sd[prefix + "lnff.bias"])


def encoder(sd, x, sym, num_layers=2, H=64, trans_layers=2, collect=None, variant=None):
    """AVWDCRNN.forward (TRAGCN.py:150-169) from the zero state; x (B,T,V,Din) -> (B,T,V,H)."""
    E = sd["node_embeddings"]
    B, T, V, _ = x.shape
    cur = x
    for i in range(num_layers):
        p = f"encoder.dcrnn_cells.{i}."
        inv_g = emb_gcn_invariants(sd, p + "gate.", E, sym) if variant is None else (variant, E)
        inv_u = emb_gcn_invariants(sd, p + "update.", E, sym) if variant is None else (variant, E)
        state = torch.zeros(B, V, H, dtype=x.dtype, device=x.device)
        states = []
        for t in range(T):
            state = gru_cell(cur[:, t], state, inv_g, inv_u, sd, p, H)
            states.append(state)
        cur = torch.stack(states, dim=1)
        if collect is not None:
            collect[f"scan{i}"] = cur
    cur = cur + positional_encoding(T, H).to(cur.device, cur.dtype)                                   # TA.py:98
    for l in range(trans_layers):
        cur = transform(cur, sd, f"encoder.trans_layer_T.trans_layers.{l}.")
        if collect is not None:
            collect[f"trans{l}"] = cur
    return cur


def targcn_forward(sd, source, adj=None, horizon=30, output_dim=64, num_layers=2, collect=None, variant=None):
    """TARGCN.forward (TRAGCN.py:209-224); source (B,T,V,D) -> (B,num_classes)."""
    V = source.shape[2]
    adj = torch.ones(V, V) if adj is None else adj                                        # :191
    sym = sym_norm_adj(adj)
    out = encoder(sd, source, sym, num_layers=num_layers, collect=collect, variant=variant)[:, -6:]       # :215
    out = F.conv2d(out, sd["end_conv.weight"], sd["end_conv.bias"])                       # (B, horizon*C, V, 1)
    out = out.squeeze(-1).reshape(-1, horizon, output_dim, V).permute(0, 1, 3, 2)         # :220-221
    feat = out.permute(0, 3, 1, 2).mean(dim=(2, 3))                                       # AdaptiveAvgPool2d(1)+Flatten
    return F.linear(feat, sd["fc.2.weight"], sd["fc.2.bias"])


def targcn_param_shapes(V=25, T=30, D=3, H=64, E=64, horizon=30, output_dim=64, num_classes=11, num_layers=2):
    sh = {"node_embeddings": (V, E)}
    for i in range(num_layers):
        cin = (D if i == 0 else H) + H
        for name, co in (("gate", 2 * H), ("update", H)):
            p = f"encoder.dcrnn_cells.{i}.{name}."
            sh[p + "weights_pool"] = (E, cin, co)
            sh[p + "bias_pool"] = (E, co)
            sh[p + "linear.weight"] = (co, cin)
            sh[p + "linear.bias"] = (co,)
    for l in range(2):
        p = f"encoder.trans_layer_T.trans_layers.{l}."
        sh[p + "vff.weight"], sh[p + "vff.bias"] = (H, H), (H,)
        for c in ("conv1", "conv2"):
            sh[p + c + ".weight"], sh[p + c + ".bias"] = (T, T, 1, 3), (T,)
        for n in ("ln", "lnff"):
            sh[p + n + ".weight"], sh[p + n + ".bias"] = (H,), (H,)
        for n in ("ff.0", "ff.2"):
            sh[p + n + ".weight"], sh[p + n + ".bias"] = (H, H), (H,)
    sh["encoder.trans_layer_T.PE.pe"] = (1, T, 1, H)
    sh["end_conv.weight"], sh["end_conv.bias"] = (horizon * output_dim, 6, 1, H), (horizon * output_dim,)
    sh["fc.2.weight"], sh["fc.2.bias"] = (num_classes, output_dim), (num_classes,)
    return sh


