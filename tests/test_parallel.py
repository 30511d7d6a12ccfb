"""CPU (gloo, world_size 2) test of the bucketed gradient all-reduce used for batch-sharded DP."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fall_multimodal_b200.parallel import GradBuckets

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    buckets = GradBuckets([list(net[0].parameters()), list(net[2].parameters())])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)
    for step in range(2):  # second step checks that zero_grad re-arms the hooks
        buckets.zero_grad()
        ((net(x[shard]) - y[shard]) ** 2).mean().backward()
        buckets.wait()
    got = [p.grad.clone() for p in net.parameters()]
    # single-process result on the full batch (mean of shard means == full mean for equal shards)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    ((ref(x) - y) ** 2).mean().backward()
    ok = all(torch.allclose(a, b.grad, atol=1e-6) for a, b in zip(got, ref.parameters()))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_convert_sync_batchnorm_marks_every_trunk():
    """Host logic of SyncBN selection (parallel.convert_sync_batchnorm); the exchange itself is GPU-tested on two NCCL ranks
    (tests/test_parity_baseline.py::test_syncbn_two_ranks_match_global_batch)."""
    import pytest
    import fall_multimodal_b200 as fmm
    from fall_multimodal_b200.parallel import convert_sync_batchnorm

    ga = {"layout": "mediapipe33", "strategy": "spatial"}
    m3 = fmm.ThreeStreamSTGCAN(3, ga, 11)
    assert all(t._engine.sync_bn is None for t in (m3.stgcan_1, m3.stgcan_2, m3.stgcan_3))
    assert convert_sync_batchnorm(m3, strict=True) is m3
    assert all(t._engine.sync_bn is True for t in (m3.stgcan_1, m3.stgcan_2, m3.stgcan_3))
    assert m3.stgcan_1._engine._sync(True) is None        # no process group initialised: per-device statistics
    assert m3.stgcan_1._engine._sync(False) is None
    convert_sync_batchnorm(m3, process_group=False)
    assert all(t._engine.sync_bn is None for t in (m3.stgcan_1, m3.stgcan_2, m3.stgcan_3))
    m2 = fmm.TwoStreamSTGCAN_CNN1D(3, ga, 11, 15, 30)
    convert_sync_batchnorm(m2)                              # trunks converted, the sensor branch stays per shard
    assert m2.stgcan_1._engine.sync_bn is True
    with pytest.raises(NotImplementedError):
        convert_sync_batchnorm(m2, strict=True)
    with pytest.raises(ValueError):
        convert_sync_batchnorm(fmm.CNN1D(15, 30))
