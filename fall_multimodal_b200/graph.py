"""Skeleton graphs and the sparse (CSR) form of the partitioned adjacency the kernels consume.

Same public face as the reference ``Graph`` (``/root/reference/Fall_2_Spatial_Temporal_SR/Model/graph.py:6-100``):
``Graph(layout, strategy, max_hop, dilation).A`` is a float64 ``(K, V, V)`` array. Hop distances are
computed with a breadth-first search instead of dense matrix powers; the partition rules are the
ST-GCN ones (uniform / distance / spatial-configuration).
"""
from __future__ import annotations

from collections import deque

import numpy as np

_LAYOUTS = {
    "coco_cut": (14, [(6, 4), (4, 2), (2, 13), (13, 1), (5, 3), (3, 1), (12, 10), (10, 8), (8, 2), (11, 9),
                      (9, 7), (7, 1), (13, 0)], 13),
    "coco_mmpose": (18, [(0, 1), (1, 3), (0, 2), (2, 4), (17, 0), (17, 6), (6, 8), (8, 10), (17, 5), (5, 7),
                         (7, 9), (17, 12), (12, 14), (14, 16), (17, 11), (11, 13), (13, 15)], 17),
    "openpose": (18, [(4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2),
                      (5, 1), (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14)], 1),
    "ntu-rgb+d": (25, [(0, 1), (1, 20), (2, 20), (3, 2), (4, 20), (5, 4), (6, 5), (7, 6), (8, 20), (9, 8),
                       (10, 9), (11, 10), (12, 0), (13, 12), (14, 13), (15, 14), (16, 0), (17, 16), (18, 17),
                       (19, 18), (21, 22), (22, 7), (23, 24), (24, 11)], 20),
    # MediaPipe-Pose, 33 landmarks / 35 connections (the V=33 shape of BASELINE.json), nose-centred
    "mediapipe33": (33, [(0, 1), (1, 2), (2, 3), (3, 7), (0, 4), (4, 5), (5, 6), (6, 8), (9, 10), (11, 12),
                         (11, 13), (13, 15), (15, 17), (15, 19), (15, 21), (17, 19), (12, 14), (14, 16),
                         (16, 18), (16, 20), (16, 22), (18, 20), (11, 23), (12, 24), (23, 24), (23, 25),
                         (24, 26), (25, 27), (26, 28), (27, 29), (28, 30), (29, 31), (30, 32), (27, 31),
                         (28, 32)], 0),
}


def register_layout(name: str, num_node: int, neighbor_link, center: int) -> None:
    """Add a custom skeleton layout (list of undirected joint pairs)."""
    _LAYOUTS[name] = (int(num_node), [tuple(e) for e in neighbor_link], int(center))


def _hops(num_node, links, max_hop):
    nbr = [[] for _ in range(num_node)]
    for i, j in links:
        nbr[i].append(j)
        nbr[j].append(i)
    hop = np.full((num_node, num_node), np.inf)
    for s in range(num_node):
        hop[s, s] = 0
        q = deque([(s, 0)])
        seen = {s}
        while q:
            u, d = q.popleft()
            if d == max_hop:
                continue
            for w in nbr[u]:
                if w not in seen:
                    seen.add(w)
                    hop[s, w] = d + 1
                    q.append((w, d + 1))
    return hop


class Graph:
    def __init__(self, layout="coco_cut", strategy="uniform", max_hop=1, dilation=1):
        if layout not in _LAYOUTS:
            raise ValueError("This layout is not supported!")
        self.max_hop, self.dilation = max_hop, dilation
        self.num_node, links, self.center = _LAYOUTS[layout]
        self.edge = [(i, i) for i in range(self.num_node)] + list(links)
        self.hop_dis = _hops(self.num_node, links, max_hop)
        self.A = self._partition(strategy)

    def _partition(self, strategy):
        V = self.num_node
        hops = list(range(0, self.max_hop + 1, self.dilation))
        reach = np.zeros((V, V))
        for h in hops:
            reach[self.hop_dis == h] = 1
        deg = reach.sum(0)
        norm = reach / np.where(deg > 0, deg, 1)[None, :]  # column-normalised A D^-1
        if strategy == "uniform":
            return norm[None].copy()
        if strategy == "distance":
            return np.stack([np.where(self.hop_dis == h, norm, 0.0) for h in hops])
        if strategy == "spatial":
            dc = self.hop_dis[:, self.center]
            parts = []
            for h in hops:
                on = self.hop_dis == h                       # [j, i]
                same = on & (dc[:, None] == dc[None, :])
                closer = on & (dc[:, None] > dc[None, :])
                further = on & ~same & ~closer
                if h == 0:
                    parts.append(np.where(same, norm, 0.0))
                else:
                    parts.append(np.where(same | closer, norm, 0.0))
                    parts.append(np.where(further, norm, 0.0))
            return np.stack(parts)
        raise ValueError("This strategy is not supported!")


def adjacency_csr(A: np.ndarray):
    """Sparse views of the (K,V,V) adjacency pattern (the learned edge importance only rescales it).

    Returns a dict of int32 numpy arrays:
      fwd_rowptr[K*V+1], fwd_src[E]            in-edges of (k, w), edge order "fwd"
      dense_idx[E], kk[E], dst[E]              per edge (fwd order): flat index into (K,V,V), k, w
      bwd_rowptr[V+1], bwd_perm[E]             out-edges of v; bwd_perm maps bwd position -> fwd edge id
    """
    K, V, _ = A.shape
    src, dst, kk, dense = [], [], [], []
    rowptr = [0]
    for k in range(K):
        for w in range(V):
            for v in range(V):
                if A[k, v, w] != 0:
                    src.append(v), dst.append(w), kk.append(k), dense.append((k * V + v) * V + w)
            rowptr.append(len(src))
    src_a = np.asarray(src, np.int32)
    order = np.argsort(src_a, kind="stable").astype(np.int32)
    bwd_rowptr = np.zeros(V + 1, np.int32)
    np.add.at(bwd_rowptr, src_a + 1, 1)
    bwd_rowptr = np.cumsum(bwd_rowptr).astype(np.int32)
    return {
        "fwd_rowptr": np.asarray(rowptr, np.int32), "fwd_src": src_a,
        "dst": np.asarray(dst, np.int32), "kk": np.asarray(kk, np.int32),
        "dense_idx": np.asarray(dense, np.int64), "bwd_rowptr": bwd_rowptr, "bwd_perm": order,
    }
