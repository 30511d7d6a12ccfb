"""Deterministic synthetic inputs and parameter fills shared by bench.py, the tests and the oracle.

Neutral test-data utilities (no model arithmetic): seeded clips / sensor windows / soft targets of the shapes SURVEY.md 8(d)
prescribes, and order-independent deterministic state_dict fills so that the CUDA modules, the oracle restatement and the
staged reference modules start from identical weights. Lives outside ``oracle/`` so the product arm of ``bench.py`` imports
nothing from the checker.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def fill_state_dict(shapes, seed=0):
    """Order-independent deterministic fill: every key gets its own generator seeded by its name.

    Conv/linear weights ~ N(0, 1/fan_in); biases small; BN weight in [0.5,1.5]; running stats
    plausible; edge_importance around 1; ``A`` is NOT filled here (build_adjacency provides it).
    """
    import zlib
    sd = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) % (2 ** 31))
        if k == "A" or k.endswith(".A"):
            continue
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shp, generator=g) + 0.5
        elif "edge_importance" in k:
            sd[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif len(shp) == 1:  # BN / norm weight
            sd[k] = torch.rand(shp, generator=g) + 0.5
        else:
            fan_in = int(np.prod(shp[1:]))
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
    return sd


def synthetic_batch(N, T, V, num_class=11, sensor_len=30, sensor_ch=15, seed=42, per_clip=True):
    """xy in [-1,1], score ~ U(0,1), sensor ~ N(0,1), label-smoothed soft targets.

    ``per_clip`` gives every clip its own pose scale/offset and sensor gain (different subjects at
    different positions). Without it all clips have near-identical pooled statistics and the
    squeeze-excite BatchNorm over N (stgcan.py:66) becomes ill-conditioned: its output is then
    dominated by fp32 rounding and no two correct fp32 implementations agree to 1e-4.
    """
    g = torch.Generator().manual_seed(seed)
    skel = torch.empty(N, 3, T, V)
    skel[:, :2] = torch.rand(N, 2, T, V, generator=g) * 2 - 1
    skel[:, 2] = torch.rand(N, T, V, generator=g)
    sensor = torch.randn(N, sensor_len, sensor_ch, generator=g)
    if per_clip:
        scale = 0.3 + 0.7 * torch.rand(N, 1, 1, 1, generator=g)
        off = torch.rand(N, 2, 1, 1, generator=g) - 0.5
        skel[:, :2] = (skel[:, :2] * scale * 0.5 + off).clamp(-1, 1)
        skel[:, 2:] = skel[:, 2:] * (0.5 + 0.5 * torch.rand(N, 1, 1, 1, generator=g))
        sensor = sensor * (0.5 + torch.rand(N, 1, 1, generator=g)) + 0.3 * torch.randn(N, 1, sensor_ch, generator=g)
    labels = torch.randint(0, num_class, (N,), generator=g)
    eps = 0.1
    target = torch.full((N, num_class), eps / (num_class - 1))
    target[torch.arange(N), labels] = 1 - eps
    return skel, sensor, target, labels


def positional_encoding(T: int, C: int) -> torch.Tensor:
    """TA.py:73-83 -> (1, T, 1, C)."""
    pe = torch.zeros(T, C)
    position = torch.arange(0, T).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, C, 2) * -(math.log(10000.0) / C))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).unsqueeze(2)


def fill_targcn(shapes, seed=0):
    """Deterministic, order-independent fill (per-key generators) scaled so the recurrence is lively:
    pools ~ N(0, 0.02^2... scaled by fan-in), embeddings ~ N(0,1) like the reference constructor."""
    import zlib
    sd = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) % (2 ** 31))
        if k.endswith("PE.pe"):
            sd[k] = positional_encoding(shp[1], shp[3])
        elif k == "node_embeddings":
            sd[k] = torch.randn(shp, generator=g) * 0.5
        elif k.endswith("weights_pool"):
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(shp[0] * shp[1]) * 2.0
        elif k.endswith("bias_pool"):
            sd[k] = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif len(shp) == 1:
            sd[k] = torch.rand(shp, generator=g) + 0.5
        else:
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(int(np.prod(shp[1:])))
    return sd


def synthetic_clips(B, T, V, D=3, num_class=11, seed=42):
    """(B,T,V,D) poses in [-1,1] with a per-clip modulation + soft targets."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, V, D, generator=g) * 2 - 1
    x[..., 2] = torch.rand(B, T, V, generator=g)
    x = x * (0.5 + torch.rand(B, 1, 1, 1, generator=g)) + 0.3 * torch.randn(B, 1, 1, D, generator=g)
    lab = torch.randint(0, num_class, (B,), generator=g)
    tgt = torch.full((B, num_class), 0.1 / (num_class - 1))
    tgt[torch.arange(B), lab] = 0.9
    tgt = tgt * (0.5 + 0.5 * torch.rand(B, 1, generator=g))
    return x, tgt
