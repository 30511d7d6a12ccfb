"""CPU tests of the product-side graph builder against the oracle / reference-generated fixtures."""
import numpy as np
import pytest
import torch

from fall_multimodal_b200.graph import Graph, adjacency_csr
from oracle import stgcn_oracle as O
from tests.golden_util import load


@pytest.mark.parametrize("layout", ["coco_cut", "coco_mmpose", "mediapipe33", "ntu-rgb+d"])
@pytest.mark.parametrize("strategy", ["uniform", "distance", "spatial"])
@pytest.mark.parametrize("max_hop", [1, 2])
def test_graph_equals_oracle(layout, strategy, max_hop):
    mine = Graph(layout, strategy, max_hop=max_hop).A
    ref = O.build_adjacency(layout, strategy, max_hop=max_hop)
    assert mine.shape == ref.shape
    assert np.array_equal(mine, ref)


def test_graph_equals_reference_fixture():
    for key, A in load("graph_A").items():
        layout, strategy = key.split("/")
        assert torch.equal(torch.tensor(Graph(layout, strategy).A), A), key


def test_unknown_layout_raises():
    with pytest.raises(ValueError):
        Graph("nope")
    with pytest.raises(ValueError):
        Graph("coco_cut", "nope")


def test_csr_round_trip():
    A = Graph("mediapipe33", "spatial").A
    csr = adjacency_csr(A)
    K, V, _ = A.shape
    E = len(csr["fwd_src"])
    assert E == int((A != 0).sum()) == 33 + 2 * 35
    dense = np.zeros(K * V * V)
    dense[csr["dense_idx"]] = 1
    assert np.array_equal(dense.reshape(K, V, V) != 0, A != 0)
    # fwd CSR rows are (k, w); bwd CSR rows are v and point back at fwd edge ids
    for k in range(K):
        for w in range(V):
            lo, hi = csr["fwd_rowptr"][k * V + w], csr["fwd_rowptr"][k * V + w + 1]
            assert sorted(csr["fwd_src"][lo:hi]) == sorted(np.nonzero(A[k, :, w])[0])
    for v in range(V):
        ids = csr["bwd_perm"][csr["bwd_rowptr"][v]:csr["bwd_rowptr"][v + 1]]
        assert all(csr["fwd_src"][i] == v for i in ids)
        assert len(ids) == int((A[:, v, :] != 0).sum())
