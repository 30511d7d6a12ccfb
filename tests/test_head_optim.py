"""GPU: fused late-fusion head + cross-entropy (csrc/head.cu), multi-tensor RMSprop (csrc/optim.cu) and the data_bn kernels
(csrc/databn.cu) against the torch expressions of the reference call sites (combination.py:44-46, F2/main.py:113,280,
F2/optimizer.py:20-21, MF3/main.py:103-113, stgcan.py:212-218), evaluated in fp64."""
import pytest
import torch

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("pre_softmax", [False, True])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_linear_cross_entropy_matches_torch(pre_softmax, smoothing):
    from fall_multimodal_b200.head import linear_cross_entropy

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    N, C, widths = 37, 11, (256, 256, 224)
    feats = [torch.randn(N, w, generator=g).to(dev).requires_grad_(i != 2) for i, w in enumerate(widths)]
    W = (torch.randn(C, sum(widths), generator=g) * 0.05).to(dev).requires_grad_(True)
    b = (torch.randn(C, generator=g) * 0.1).to(dev).requires_grad_(True)
    lab = torch.randint(0, C, (N,), generator=g)
    tgt = torch.full((N, C), 0.1 / (C - 1))
    tgt[torch.arange(N), lab] = 0.9
    tgt = (tgt * (0.5 + 0.5 * torch.rand(N, 1, generator=g))).to(dev)          # score-weighted soft labels: rows do not sum to 1
    pred, loss = linear_cross_entropy(feats, W, b, tgt, pre_softmax=pre_softmax, label_smoothing=smoothing)
    (loss * 1.7).backward()
    # reference in fp64
    f64 = [f.detach().double().requires_grad_(True) for f in feats]
    W64, b64 = W.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    z = torch.cat(f64, 1) @ W64.t() + b64
    o = torch.softmax(z, -1) if pre_softmax else z
    rloss = torch.nn.CrossEntropyLoss(label_smoothing=smoothing)(o, tgt.double())
    (rloss * 1.7).backward()
    assert (pred.double() - o).abs().max().item() < 1e-5 * max(1.0, o.abs().max().item())
    assert abs(loss.item() - rloss.item()) < 1e-5 * max(1.0, abs(rloss.item()))
    rel = lambda a, r: (a.double() - r).abs().max().item() / max(r.abs().max().item(), 1e-12)
    assert rel(W.grad, W64.grad) < 1e-4 and rel(b.grad, b64.grad) < 1e-4
    assert rel(feats[0].grad, f64[0].grad) < 1e-4 and rel(feats[1].grad, f64[1].grad) < 1e-4
    assert feats[2].grad is None
    # class-index targets
    pred2, loss2 = linear_cross_entropy([f.detach() for f in feats], W.detach(), b.detach(), lab.to(dev), pre_softmax=pre_softmax)
    r2 = torch.nn.CrossEntropyLoss()(o.detach(), lab.to(dev))
    assert abs(loss2.item() - r2.item()) < 1e-5 * max(1.0, abs(r2.item()))


@gpu
def test_fusion_forward_loss_matches_forward_plus_ce():
    import fall_multimodal_b200 as fmm
    from oracle import stgcn_oracle as O

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    m = fmm.TwoStreamSTGCAN_CNN1D(3, {"layout": "coco_cut", "strategy": "spatial"}, 11, 15, 30).to(dev).train()
    m.compute_dtype = torch.float32
    skel, sensor, tgt = (t.to(dev) for t in O.synthetic_batch(8, 16, 14, 11, sensor_len=30, sensor_ch=15, seed=2)[:3])
    pred, loss = m.forward_loss(skel, sensor, tgt, label_smoothing=0.05)
    loss.backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    out = m(skel, sensor)
    l2 = torch.nn.CrossEntropyLoss(label_smoothing=0.05)(out, tgt)
    l2.backward()
    assert (pred - out).abs().max().item() < 1e-5 * out.abs().max().item() and abs(loss.item() - l2.item()) < 1e-5
    gs = max(p.grad.abs().max().item() for p in m.parameters() if p.grad is not None)
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        scale = max(p.grad.abs().max().item(), 1e-3 * gs)      # conv biases in front of a train-mode BN: analytically zero
        assert (g1[k] - p.grad).abs().max().item() / scale < 2e-4, k       # (atomics reorder fp32 sums between the two runs)


@gpu
@pytest.mark.parametrize("mode", ["plain", "clip", "scaled", "overflow"])
def test_fused_rmsprop_matches_torch(mode):
    from fall_multimodal_b200.optim import FusedRMSprop

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    shapes = [(64, 3, 1, 1), (257,), (33, 33, 3), (5000,), (1,), (128, 64, 9, 1)]
    P1 = [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]
    P2 = [p.detach().clone().requires_grad_(True) for p in P1]
    max_norm = 0.5 if mode in ("clip", "scaled") else None
    o1 = FusedRMSprop(P1, lr=1e-3, weight_decay=0.01 if mode == "plain" else 0.0, max_norm=max_norm)
    o2 = torch.optim.RMSprop(P2, lr=1e-3, weight_decay=0.01 if mode == "plain" else 0.0)
    scale = 1024.0 if mode in ("scaled", "overflow") else 1.0
    if scale != 1.0:
        o1.inv_scale = torch.tensor(1.0 / scale, device=dev)
    for step in range(4):
        grads = [torch.randn(s, generator=g).to(dev) * (step + 1) for s in shapes]
        if step == 2:
            o1.param_groups[0]["lr"] = 5e-4      # a scheduler assigning a new value
            o2.param_groups[0]["lr"] = 5e-4
        for p1, p2, gr in zip(P1, P2, grads):
            p1.grad = (gr * scale).clone()
            p2.grad = gr.clone()
        if mode == "overflow" and step == 1:
            P1[3].grad[17] = float("inf")
            before = [p.detach().clone() for p in P1]
            o1.step()
            assert all(torch.equal(a, b) for a, b in zip(before, P1)), "a non-finite gradient must skip the step"
            for p1, gr in zip(P1, grads):          # the reference loop retries with a smaller scale: replay the same step unscaled
                p1.grad = (gr * scale).clone()
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_(P2, max_norm)
        o1.step()
        o2.step()
    for p1, p2 in zip(P1, P2):
        assert (p1 - p2).abs().max().item() < 2e-6 * max(1.0, p2.abs().max().item())
    sd = o1.state_dict()
    assert set(sd["state"][0].keys()) >= {"square_avg", "step"} and isinstance(sd["param_groups"][0]["lr"], float)


@gpu
def test_fused_rmsprop_under_cuda_graph_with_lr_change():
    from fall_multimodal_b200.optim import FusedRMSprop

    dev = torch.device("cuda:0")
    p = torch.ones(1000, device=dev, requires_grad=True)
    lr = torch.tensor(1e-2, device=dev)
    opt = FusedRMSprop([p], lr=lr, max_norm=10.0)
    gbuf = torch.full((1000,), 0.5, device=dev)

    def step():
        p.grad = gbuf * 1.0
        opt.step()

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    v0 = p.detach().clone()
    lr.fill_(0.0)
    graph.replay()
    assert torch.equal(p.detach(), v0)
    lr.fill_(1e-2)
    graph.replay()
    assert (p.detach() < v0).all()


@gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_databn_kernels_match_batchnorm1d(dtype):
    from fall_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    N, C, T, V = 5, 3, 9, 14
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, C, T, V, generator=g).to(dev)
    bn = torch.nn.BatchNorm1d(V * C).to(dev).double().train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(V * C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(V * C, generator=g))
    rm0, rv0 = bn.running_mean.clone().float(), bn.running_var.clone().float()
    # reference: stgcan.py:213-218
    xr = x.double().permute(0, 3, 1, 2).contiguous().view(N, V * C, T)
    yr = bn(xr).view(N, V, C, T).permute(0, 2, 3, 1)                       # (N,C,T,V)
    st = torch.zeros(2 * V * C, dtype=torch.float64, device=dev)
    a, b, mean, rstd = (torch.empty(V * C, device=dev) for _ in range(4))
    ops.databn_stats(x, st[:V * C], st[V * C:])
    ops.bn_finalize(st[:V * C], st[V * C:], N * T, bn.weight.float(), bn.bias.float(), rm0, rv0, True, a, b, mean, rstd)
    y = ops.databn_apply(x, a, b, torch.empty(N, T, V, C, dtype=dtype, device=dev))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (y.double().permute(0, 3, 1, 2) - yr).abs().max().item() < tol * yr.abs().max().item()
    assert (rm0.double() - bn.running_mean).abs().max().item() < 1e-6 and (rv0.double() - bn.running_var).abs().max().item() < 1e-6
    dy = torch.randn(N, T, V, C, generator=g).to(dev, dtype)
    (yr * dy.double().permute(0, 3, 1, 2)).sum().backward()
    dgb = torch.zeros(2 * V * C, dtype=torch.float64, device=dev)
    ops.databn_bwd(dy, x, mean, rstd, dgb[:V * C], dgb[V * C:])
    assert (dgb[:V * C] - bn.weight.grad).abs().max().item() < 1e-4 * bn.weight.grad.abs().max().item()
    assert (dgb[V * C:] - bn.bias.grad).abs().max().item() < 1e-4 * bn.bias.grad.abs().max().item()


@gpu
@pytest.mark.parametrize("I,Tn,N", [(6, 128, 200), (15, 30, 70), (4, 30, 64), (32, 7, 130)])
@pytest.mark.parametrize("feature", ["mean", "last"])
def test_lstm_tensor_core_inference_matches_training_kernel_and_torch(I, Tn, N, feature):
    """csrc/lstm_tc.cu (3xTF32 mma.sync recurrence) against nn.LSTM in fp64 and the scalar training kernel, ragged batch sizes."""
    from fall_multimodal_b200 import BiLSTM

    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    m = BiLSTM(I, 64, 1, 0.3, 11, feature).to(dev).eval()
    x = torch.randn(N, Tn, I, device=dev) * 1.5
    with torch.no_grad():
        fast = m.features(x)
        m.fast_inference = False
        slow = m.features(x)
        ref_lstm = torch.nn.LSTM(I, 64, 1, batch_first=True, bidirectional=True).to(dev).double()
        ref_lstm.load_state_dict({k: v.double() for k, v in m.lstm1.state_dict().items()})
        o = ref_lstm(x.double())[0]
        ref = o[:, -1, :] if feature == "last" else o.mean(1)
    assert fast.shape == (N, 128)
    e_fast = (fast.double() - ref).abs().max().item() / ref.abs().max().item()
    e_slow = (slow.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"I={I} T={Tn} {feature}: tensor-core path {e_fast:.2e}, scalar kernel {e_slow:.2e} vs fp64 nn.LSTM")
    assert e_fast < 2e-5 and e_slow < 2e-5
