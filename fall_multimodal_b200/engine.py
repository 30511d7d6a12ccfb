"""Kernel sequencing for one GSTCAN trunk: forward and hand-written backward.

This is the host-side mirror of ``STGCAN.forward`` (``/root/reference/Fall_2_Spatial_Temporal_SR/
Model/stgcan.py:210-228``) and of what autograd would do for it, expressed as launches of the
C-ABI kernels (csrc/*.cu). Per block (stgcan.py:138-144):

    Xa = aggregate(x, A*importance)           agg_fwd           (einsum :54, reassociated onto the input)
    G  = Xa @ Wg + bias_eff[v]                tapconv 1x1       (conv :51)
    U  = conv9x1(relu(bn1(G)), stride)        colstats + bn_finalize + tapconv (BN/ReLU in the prologue)
    s  = SE(mean_tv(bn2(U)))                  colstats(+pool) + bn_finalize + se_fwd
    y  = relu(s*bn2(U) + res)                 block_out         (res: none | x | bn_r(conv1x1_s(x)))

All tensors are allocated by PyTorch; the kernels only see raw pointers. Tiny once-per-step glue
(data_bn on the 3-channel input, the adjacency coefficient prep, the classifier head) is torch.
"""
from __future__ import annotations

import torch

from . import ops
from .graph import adjacency_csr

BLOCK_PLAN = [(None, 64, 1, "none"), (64, 64, 1, "identity"), (64, 64, 1, "identity"), (64, 128, 2, "conv"),
              (128, 128, 1, "identity"), (128, 256, 2, "conv"), (256, 256, 1, "identity")]
EPS = 1e-5
MOMENTUM = 0.1


class _Arena:
    """One zero-filled fp64 and one fp32 buffer per pass, carved into accumulators (one memset each)."""

    def __init__(self, device, n64, n32):
        self.b64 = torch.zeros(n64, dtype=torch.float64, device=device)
        self.b32 = torch.zeros(n32, dtype=torch.float32, device=device)
        self.o64 = 0
        self.o32 = 0

    def f64(self, n):
        out = self.b64[self.o64:self.o64 + n]
        self.o64 += n
        assert self.o64 <= self.b64.numel()
        return out

    def f32(self, *shape):
        n = 1
        for s in shape:
            n *= s
        out = self.b32[self.o32:self.o32 + n]
        self.o32 += (n + 3) // 4 * 4
        assert self.o32 <= self.b32.numel()
        return out.view(*shape)


class TrunkEngine:
    """Static plan of one trunk: block shapes and the sparse adjacency, per device."""

    def __init__(self, A: torch.Tensor, in_channels: int, block_key: str = "st_gcan_networks"):
        self.K, self.V, _ = A.shape
        self.in_channels = in_channels
        self.block_key = block_key
        self.blocks = [(in_channels if cin is None else cin, cout, s, res) for cin, cout, s, res in BLOCK_PLAN]
        self._csr_np = adjacency_csr(A.detach().cpu().double().numpy())
        self.E = len(self._csr_np["fwd_src"])
        self.kdeg = ops.partition_degrees([int(v) for v in self._csr_np["fwd_rowptr"]], self.K, self.V)
        self._dev = {}
        self.debug = None  # dev aid: set to a dict to capture backward intermediates per block
        # Weight gradients are only needed when backward returns: they are queued on a side stream so
        # they overlap the dgrad -> BatchNorm-backward -> aggregation chain of the same and later blocks.
        self.materialize_h = True  # bf16: write relu(bn1(G)) once instead of transforming it in two GEMM prologues
        self.overlap_wgrad = False  # measured: no gain, two persistent GEMM CTAs cannot share an SM
        self.side_params = True     # tiny parameter-gradient-only launches (SE weight products, conv-bias / edge-importance sums) on a side stream
        self.fused_gcn = True        # bf16, C % 64 == 0: csrc/gcn.cu instead of agg_fwd + 1x1 tapconv + colstats
        self.fused_gcn_wgrad = True   # weight gradient re-derives the aggregated operand (no saved Xa)
        self.fused_gcn_bwd = True     # dx / d(edge importance): GEMM + transposed aggregation in one kernel (no P tensor)
        rp = self._csr_np["bwd_rowptr"]
        self.max_out_deg = int(max(int(rp[i + 1]) - int(rp[i]) for i in range(self.V)))
        self._wstreams = {}
        self._pstreams = {}
        self.premask = True   # fused graph-conv backward stores dx already masked by the previous block's ReLU
        # SyncBN (SURVEY 8(e) option a): None = per-shard statistics; True / a process group = every BatchNorm of the trunk
        # (data_bn, BN1, BN2, residual BN, the SE block's BatchNorm over the batch axis) normalises with the statistics of the
        # GLOBAL batch. Set through parallel.convert_sync_batchnorm. Per block and direction at most three small collectives:
        # forward  all-reduce(BN1 sums), all-reduce(BN2 sums) [+ residual-BN sums in the two strided blocks], all-gather(SE pool);
        # backward all-gather(per-clip sums S1..S3 of the block tail), all-reduce(BN1 gradient sums).
        # The per-clip (N,C) tensors of the SE / BN2-coefficient kernels are tiny, so they are gathered and every rank runs those
        # kernels on the global N (unchanged kernels, global BatchNorm over N for free) and keeps its own rows.
        self.sync_bn = None

    def csr(self, device):
        key = str(device)
        if key not in self._dev:
            c = self._csr_np
            t = lambda a, dt=torch.int32: torch.as_tensor(a).to(device=device, dtype=dt)
            d = {"fwd_rowptr": t(c["fwd_rowptr"]), "fwd_src": t(c["fwd_src"]), "dst": t(c["dst"]), "kk": t(c["kk"]),
                 "dense_idx": t(c["dense_idx"], torch.int64), "bwd_rowptr": t(c["bwd_rowptr"]),
                 "bwd_perm": t(c["bwd_perm"], torch.int64)}
            d["eid_b"] = d["bwd_perm"].to(torch.int32).contiguous()
            d["dst_b"] = d["dst"][d["bwd_perm"]].contiguous()
            d["kk_b"] = d["kk"][d["bwd_perm"]].contiguous()
            self._dev[key] = d
        return self._dev[key]

    # ------------------------------------------------------------------------------------------
    def _prepare(self, P: dict, dt: torch.dtype, need_grad: bool, dev, V: int):
        """Everything that depends on the parameters only - the (K,V,V) adjacency algebra of every block and all weight
        images of the forward AND backward GEMMs (~110 tiny launches, 0.6 ms serialised) - queued on a side stream at the start
        of the step, off the critical path of the activation chain; block i waits for its own event."""
        K = self.K
        csr = self.csr(dev)
        cur = torch.cuda.current_stream(dev)
        key = (str(dev), cur.cuda_stream)
        if key not in self._pstreams:
            self._pstreams[key] = torch.cuda.Stream(device=dev)
        side = self._pstreams[key]
        side.wait_stream(cur)
        prep = []
        A = P["A"]
        with torch.cuda.stream(side):
            for i, (Cin, Cout, s, reskind) in enumerate(self.blocks):
                pre = f"{self.block_key}.{i}."
                d = {}
                d["coef_f"], d["coef_b"], d["colsum"], d["bias_eff"] = ops.gcn_prep_fwd(
                    A, P[f"edge_importance.{i}"], P[pre + "gcn.conv.bias"], csr["dense_idx"], csr["bwd_perm"], Cout)
                Wg, Wt = P[pre + "gcn.conv.weight"], P[pre + "tcn.2.weight"]
                d["fused"] = self.fused_gcn and ops.gcn_supported(dt, Cin, Cout, V)
                if d["fused"]:
                    d["wg"] = ops.gcn_pack(Wg.view(K * Cout, Cin), K, Cin, Cout)
                else:
                    d["wg"] = ops.tapconv_pack(Wg, Cout, K * Cin, Cout, Cin, 0, Cin, Cout * Cin, 1, 0, [0], dt)
                d["wt"] = ops.tapconv_pack(Wt, Cout, Cout, Cout, Cout, 0, Cout * 9, 0, 9, 1, list(range(9)), dt)
                if reskind == "conv":
                    Wr = P[pre + "residual.0.weight"]
                    d["wr"] = ops.tapconv_pack(Wr, Cout, Cin, Cout, Cin, 0, Cin, 0, 1, 0, [0], dt)
                if need_grad:
                    if s == 1:
                        d["wt_d"] = [ops.tapconv_pack(Wt, Cout, Cout, Cout, Cout, 0, 9, 0, Cout * 9, 1, list(range(9)), dt)]
                    else:
                        d["wt_d"] = [ops.tapconv_pack(Wt, Cout, Cout, Cout, Cout, 0, 9, 0, Cout * 9, 1, [0, 2, 4, 6, 8], dt),
                                     ops.tapconv_pack(Wt, Cout, Cout, Cout, Cout, 0, 9, 0, Cout * 9, 1, [1, 3, 5, 7], dt)]
                    d["fused_bwd"] = self.fused_gcn_bwd and ops.gcn_bwd_supported(dt, Cin, Cout, V, K, self.max_out_deg)
                    if d["fused_bwd"]:
                        d["wg_b"] = ops.gcn_pack_bwd(Wg.view(K * Cout, Cin), K, Cin, Cout)
                    else:
                        d["wg_b"] = ops.tapconv_pack(Wg, K * Cin, Cout, Cin, Cout, Cout * Cin, 1, 0, Cin, 0, [0], dt)
                    if reskind == "conv":
                        d["wr_d"] = ops.tapconv_pack(Wr, Cin, Cout, Cin, Cout, 0, 1, 0, Cin, 0, [0], dt)
                for v in d.values():
                    for t in (v if isinstance(v, list) else [v]):
                        buf = t.buf if isinstance(t, ops.PackedWeight) else t
                        if torch.is_tensor(buf):
                            buf.record_stream(cur)
                d["ev"] = torch.cuda.Event()
                d["ev"].record(side)
                prep.append(d)
        return prep

    def _sync(self, training: bool):
        """(torch.distributed, group, world, rank) when batch statistics are exchanged in this pass, else None."""
        if not training or self.sync_bn is None or self.sync_bn is False:
            return None
        import torch.distributed as dist

        if not dist.is_initialized():
            return None
        group = None if self.sync_bn is True else self.sync_bn
        world = dist.get_world_size(group)
        if world == 1:
            return None
        return dist, group, world, dist.get_rank(group)

    # ------------------------------------------------------------------------------------------
    def forward(self, P: dict, skel: torch.Tensor, training: bool, dt: torch.dtype, need_grad: bool):
        """Returns (pooled feature (N,256) fp32, saved-state dict for backward)."""
        dev = skel.device
        N, C, T, V = skel.shape
        assert C == self.in_channels and V == self.V, "input does not match the trunk's channels / joints"
        K = self.K
        csr = self.csr(dev)
        A = P["A"]
        NR = ops.NREP
        arena = _Arena(dev, 7 * 6 * NR * 256 + 4096 + 2 * V * C, 8 * N * 256 + 4096)
        sv = {"blocks": [], "N": N, "dt": dt, "training": training}
        sy = self._sync(training)           # SyncBN: every rank must hold the same number of clips
        W = sy[2] if sy else 1
        sv["sync"] = sy
        prep = self._prepare(P, dt, need_grad, dev, V)
        sv["prep"] = prep
        cur_stream = torch.cuda.current_stream(dev)

        # ---- data_bn (stgcan.py:213-218): per-(v,c) BatchNorm1d over (N,T), csrc/databn.cu; the normalised clip is written
        # once, already channels-last in the compute dtype ----
        xin = skel.float().contiguous()
        VC = V * C
        st0 = arena.f64(2 * VC)
        a0, b0, mean0, rstd0 = (torch.empty(VC, dtype=torch.float32, device=dev) for _ in range(4))
        if training:
            ops.databn_stats(xin, st0[:VC], st0[VC:])
            if sy:
                sy[0].all_reduce(st0, group=sy[1])
        ops.bn_finalize(st0[:VC], st0[VC:], N * T * W, P["data_bn.weight"], P["data_bn.bias"], P["data_bn.running_mean"],
                        P["data_bn.running_var"], training, a0, b0, mean0, rstd0)
        x = ops.databn_apply(xin, a0, b0, torch.empty(N, T, V, C, dtype=dt, device=dev))
        sv["data_bn"] = (xin, mean0, rstd0)

        for i, (Cin, Cout, s, reskind) in enumerate(self.blocks):
            pre = f"{self.block_key}.{i}."
            b = {"T": T}
            pp = prep[i]
            cur_stream.wait_event(pp["ev"])       # this block's coefficient tables / weight images (side stream, _prepare)
            coef_f, coef_b, colsum, bias_eff = pp["coef_f"], pp["coef_b"], pp["colsum"], pp["bias_eff"]
            Wg = P[pre + "gcn.conv.weight"]
            G = torch.empty(N, T, V, Cout, dtype=dt, device=dev)
            a1, b1, mean1, rstd1 = (torch.empty(Cout, dtype=torch.float32, device=dev) for _ in range(4))
            st = arena.f64(2 * NR * Cout)
            fused = pp["fused"]
            if fused:
                # north-star graph conv: adjacency aggregation in the tcgen05 GEMM prologue, BN1 statistics in its epilogue
                # (csrc/gcn.cu); the K-times wider aggregated tensor only reaches HBM while the old wgrad still wants it
                Xa = torch.empty(N, T, V, K * Cin, dtype=dt, device=dev) if (need_grad and not self.fused_gcn_wgrad) else None
                ops.gcn_fwd(x, pp["wg"], G, csr["fwd_rowptr"], csr["fwd_src"], coef_f, K,
                            self.kdeg, bias=bias_eff, ch_sum=st[:NR * Cout] if training else None,
                            ch_sq=st[NR * Cout:] if training else None, xa=Xa)
            else:
                Xa = torch.empty(N, T, V, K * Cin, dtype=dt, device=dev)
                ops.agg_fwd(x, Xa, csr["fwd_rowptr"], csr["fwd_src"], coef_f, K)
                ops.tapconv(Xa, pp["wg"], G, shifts=[0], tj=T, bias=bias_eff, bias_per_joint=True)
                # BN1 statistics (stgcan.py:112), applied inside the temporal conv's prologue
                if training:
                    ops.colstats(G, st[:NR * Cout], st[NR * Cout:])
            if sy:
                sy[0].all_reduce(st, group=sy[1])
            ops.bn_finalize(st[:NR * Cout], st[NR * Cout:], N * T * V * W, P[pre + "tcn.0.weight"], P[pre + "tcn.0.bias"],
                            P[pre + "tcn.0.running_mean"], P[pre + "tcn.0.running_var"], training, a1, b1, mean1, rstd1)

            # temporal conv 9x1 (stgcan.py:114-118)
            To = (T - 1) // s + 1
            Wt = P[pre + "tcn.2.weight"]
            pw_t = pp["wt"]
            U = torch.empty(N, To, V, Cout, dtype=dt, device=dev)
            Hm = None
            if self.materialize_h and dt == torch.bfloat16:
                # H = relu(bn1(G)) written once: the conv (and later its wgrad) stream it with plain cp.async
                Hm = ops.affine_relu(G, a1, b1, torch.empty_like(G))
                ops.tapconv(Hm, pw_t, U, shifts=list(range(-4, 5)), tj=To, istride=s, bias=P[pre + "tcn.2.bias"])
            else:
                ops.tapconv(G, pw_t, U, shifts=list(range(-4, 5)), tj=To, istride=s, in_scale=a1, in_shift=b1,
                            in_relu=True, bias=P[pre + "tcn.2.bias"])

            # BN2 statistics (:119) + SE pooling (:64) in one pass over U
            a2, b2, mean2, rstd2 = (torch.empty(Cout, dtype=torch.float32, device=dev) for _ in range(4))
            st2 = arena.f64(2 * NR * Cout)
            pool = arena.f32(N, Cout)
            if training:
                ops.colstats(U, st2[:NR * Cout], st2[NR * Cout:], pool)
            else:
                ops.colstats(U, None, None, pool)
            if sy:
                sy[0].all_reduce(st2, group=sy[1])
            ops.bn_finalize(st2[:NR * Cout], st2[NR * Cout:], N * To * V * W, P[pre + "tcn.3.weight"], P[pre + "tcn.3.bias"],
                            P[pre + "tcn.3.running_mean"], P[pre + "tcn.3.running_var"], training, a2, b2, mean2, rstd2)

            # squeeze-excite (stgcan.py:63-70)
            C4 = int(Cout / 4)
            ca = pre + "channel_attention_module.atten."
            f32 = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=dev)
            Ns = N * W
            if sy:   # the SE block's BatchNorm runs over the batch axis: gather the pooled rows, evaluate the block on the global N
                pool_l, pool = pool, f32(Ns, Cout)
                sy[0].all_gather_into_tensor(pool, pool_l.contiguous(), group=sy[1])
            p_, h_, s_, k1, k0 = f32(Ns, Cout), f32(Ns, C4), f32(Ns, Cout), f32(Ns, Cout), f32(Ns, Cout)
            ah, bh, hmean, hrstd = f32(C4), f32(C4), f32(C4), f32(C4)
            M = To * V
            ops.se_fwd(pool, a2, b2, 1.0 / M, P[ca + "1.weight"], P[ca + "1.bias"], P[ca + "2.weight"],
                       P[ca + "2.bias"], P[ca + "2.running_mean"], P[ca + "2.running_var"], training,
                       P[ca + "4.weight"], P[ca + "4.bias"], p_, h_, ah, bh, hmean, hrstd, s_, k1, k0)
            if sy:
                k1, k0 = k1[sy[3] * N:(sy[3] + 1) * N], k0[sy[3] * N:(sy[3] + 1) * N]

            # residual branch (stgcan.py:123-133)
            R = ar = br = meanr = rstdr = pw_r = None
            if reskind == "conv":
                Wr = P[pre + "residual.0.weight"]
                pw_r = pp["wr"]
                R = torch.empty(N, To, V, Cout, dtype=dt, device=dev)
                ops.tapconv(x, pw_r, R, shifts=[0], tj=To, istride=s, bias=P[pre + "residual.0.bias"])
                ar, br, meanr, rstdr = f32(Cout), f32(Cout), f32(Cout), f32(Cout)
                st3 = arena.f64(2 * NR * Cout)
                if training:
                    ops.colstats(R, st3[:NR * Cout], st3[NR * Cout:])
                    if sy:
                        sy[0].all_reduce(st3, group=sy[1])
                ops.bn_finalize(st3[:NR * Cout], st3[NR * Cout:], N * To * V * W, P[pre + "residual.1.weight"],
                                P[pre + "residual.1.bias"], P[pre + "residual.1.running_mean"],
                                P[pre + "residual.1.running_var"], training, ar, br, meanr, rstdr)
                res = R
            elif reskind == "identity":
                res = x
            else:
                res = None
            Y = torch.empty(N, To, V, Cout, dtype=dt, device=dev)
            ops.block_out(U, k1, k0, res, ar, br, Y)

            if need_grad:
                b.update(x=x, Xa=Xa, G=G, H=Hm, U=U, R=R, Y=Y, a1=a1, b1=b1, mean1=mean1, rstd1=rstd1, a2=a2, b2=b2,
                         mean2=mean2, rstd2=rstd2, pool=pool, p=p_, h=h_, s=s_, ah=ah, bh=bh, hmean=hmean,
                         hrstd=hrstd, ar=ar, meanr=meanr, rstdr=rstdr, coef_f=coef_f, coef_b=coef_b, colsum=colsum, To=To)
                sv["blocks"].append(b)
            x, T = Y, To

        M = T * V
        poolY = arena.f32(N, 256)
        ops.colstats(x, None, None, poolY)
        feat = poolY / M
        sv["M_last"], sv["T_last"] = M, T
        if training:
            with torch.no_grad():
                torch._foreach_add_([v for k, v in P.items() if k.endswith("num_batches_tracked")], 1)
        return feat, sv

    # ------------------------------------------------------------------------------------------
    def backward(self, P: dict, sv: dict, dfeat: torch.Tensor) -> dict:
        """dfeat: (N,256) fp32 gradient of the pooled feature. Returns {param name: fp32 gradient}."""
        dev = dfeat.device
        N, dt, training = sv["N"], sv["dt"], sv["training"]
        K, V = self.K, self.V
        csr = self.csr(dev)
        A = P["A"]
        grads = {}
        NR = ops.NREP
        wsize = sum(9 * co * co + K * co * ci + (co * ci if r == "conv" else 0) + 4 * K * V * V
                    for ci, co, _, r in self.blocks)
        arena = _Arena(dev, 7 * 4 * NR * 256 + 4096 + 2 * V * self.in_channels,
                       40 * N * 256 + 7 * NR * V * 256 + 7 * 2 * 256 * 64 + 65536 + wsize + 64 * len(self.blocks))
        f32 = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=dev)
        z32 = lambda *sh: torch.zeros(*sh, dtype=torch.float32, device=dev)
        dY = (dfeat / sv["M_last"]).to(dt)[:, None, None, :].expand(N, sv["T_last"], V, 256).contiguous()
        cur = torch.cuda.current_stream(dev)
        sy = sv.get("sync")
        W = sy[2] if sy else 1
        Ns = N * W
        rows = slice(sy[3] * N, (sy[3] + 1) * N) if sy else slice(None)
        # SyncBN: gradients of the BatchNorm / SE parameters come out of kernels that ran on the GLOBAL batch; divided by the
        # world size at the end they are this rank's share, so the data-parallel gradient average reproduces the global sum
        glob = []
        ws = None
        if self.overlap_wgrad or self.side_params:
            key = (str(dev), cur.cuda_stream)
            if key not in self._wstreams:
                self._wstreams[key] = torch.cuda.Stream(device=dev)
            ws = self._wstreams[key]
        keep = []  # operands of side-stream launches stay referenced until the join below

        def wgrad_async(*args, **kw):
            if ws is None or not self.overlap_wgrad:
                return ops.wgrad(*args, **kw)
            ws.wait_stream(cur)
            keep.append(args)
            with torch.cuda.stream(ws):
                return ops.wgrad(*args, **kw)

        def side(fn, *args):
            """Parameter-gradient-only launches (nothing downstream in this step reads them): off the activation chain."""
            if ws is None:
                return fn(*args)
            ws.wait_stream(cur)
            keep.append(args)
            with torch.cuda.stream(ws):
                return fn(*args)

        def rep_sum(t, c):
            """fp32 total of the NR fp64 replica rows of a per-channel accumulator (conv-bias gradients)."""
            out = t.view(NR, c).sum(0).float()
            out.record_stream(cur)      # allocated on the side stream, consumed on the main one after the join
            return out

        # premasked: dY already carries the ReLU mask of this block's output - the fused graph-conv backward of the block
        # above read that output anyway (edge gradient) and stored dx * (x > 0); Y is then never read again down here
        premasked = False
        for i in reversed(range(len(self.blocks))):
            Cin, Cout, s, reskind = self.blocks[i]
            pre = f"{self.block_key}.{i}."
            b = sv["blocks"][i]
            pp = sv["prep"][i]
            T, To = b["T"], b["To"]
            x, Xa, G, U, R, Y = b["x"], b["Xa"], b["G"], b["U"], b["R"], b["Y"]
            C4 = int(Cout / 4)
            M = To * V
            ca = pre + "channel_attention_module.atten."

            # ---- relu / residual / SE scale: per-(n,c) reductions of dpre = dY*(Y>0) ----
            S1, S2 = arena.f32(N, Cout), arena.f32(N, Cout)
            S3 = arena.f32(N, Cout) if R is not None else None
            Ym = None if premasked else Y
            ops.blockout_bwd_reduce(dY, Ym, U, R, S1, S2, S3)
            if sy:   # one gather per block: the per-clip sums of every rank, (world, nS, N, C) -> nS x (world * N, C)
                loc = torch.stack([S1, S2] + ([S3] if S3 is not None else []))
                gat = torch.empty((W,) + tuple(loc.shape), dtype=torch.float32, device=dev)
                sy[0].all_gather_into_tensor(gat, loc, group=sy[1])
                gat = gat.transpose(0, 1).reshape(loc.shape[0], Ns, Cout)
                S1, S2, S3 = gat[0], gat[1], (gat[2] if S3 is not None else None)

            # ---- SE backward (tiny) ----
            dq, dp = f32(Ns, Cout), f32(Ns, Cout)
            dhr, r_, dh = f32(Ns, C4), f32(Ns, C4), f32(Ns, C4)
            dW1, db1, dgh, dbh = arena.f32(C4, Cout), arena.f32(C4), arena.f32(C4), arena.f32(C4)
            dW2, db2se = arena.f32(Cout, C4), arena.f32(Cout)
            ops.se_bwd(S1, S2, b["a2"], b["b2"], b["s"], b["p"], b["h"], b["ah"], b["bh"], b["hmean"], b["hrstd"],
                       P[ca + "1.weight"], P[ca + "4.weight"], training, dq, dhr, r_, dh, dp, None, None, dgh, dbh,
                       None, None)
            side(ops.se_bwd_params, dq, r_, dh, b["p"], dW1, db1, dW2, db2se)
            grads[ca + "1.weight"] = dW1.view(C4, Cout, 1, 1)
            grads[ca + "1.bias"] = db1
            grads[ca + "2.weight"] = dgh
            grads[ca + "2.bias"] = dbh
            grads[ca + "4.weight"] = dW2.view(Cout, C4, 1, 1)
            grads[ca + "4.bias"] = db2se
            glob += [dW1, db1, dgh, dbh, dW2, db2se]

            # ---- BN2 (+ residual BN) backward folded into affine coefficients ----
            k1, k3, k2 = f32(Ns, Cout), f32(Ns, Cout), f32(Cout)
            dg2, db2 = arena.f32(Cout), arena.f32(Cout)
            r1 = r2 = r3 = dgr = dbr = None
            if R is not None:
                r1, r2, r3 = f32(Cout), f32(Cout), f32(Cout)
                dgr, dbr = arena.f32(Cout), arena.f32(Cout)
            ops.bn2_bwd_coef(S1, S2, S3, b["pool"], dp, b["s"], b["a2"], b["mean2"], b["rstd2"], b["ar"],
                             b["meanr"], b["rstdr"], M, Ns * M, training, k1, k2, k3, r1, r2, r3, dg2, db2, dgr, dbr)
            grads[pre + "tcn.3.weight"], grads[pre + "tcn.3.bias"] = dg2, db2
            glob += [dg2, db2] + ([dgr, dbr] if R is not None else [])
            if sy:
                k1, k3 = k1[rows], k3[rows]
            dU = torch.empty_like(U)
            dR = torch.empty_like(R) if R is not None else None
            dPre = torch.empty_like(Y) if (reskind == "identity" and not premasked) else None
            sum_dU = arena.f64(NR * Cout)
            sum_dR = arena.f64(NR * Cout) if R is not None else None
            ops.bn2_bwd_apply(dY, Ym, U, R, k1, k2, k3, r1, r2, r3, dU, dR, dPre, sum_dU, sum_dR)
            grads[pre + "tcn.2.bias"] = side(rep_sum, sum_dU, Cout)

            # ---- temporal conv: wgrad + dgrad ----
            Wt = P[pre + "tcn.2.weight"]
            # accumulated as [tap][co][ci] (input channel contiguous: the 32 lanes of an atomic land in one line),
            # handed to autograd as the (co, ci, tap, 1) view of that buffer
            dWs = arena.f32(9, Cout, Cout)
            if b["H"] is not None:
                wgrad_async(b["H"], dU, dWs, shifts=list(range(-4, 5)), istride=s, s_m=Cout * Cout, s_c2=1, s_co=Cout)
            else:
                wgrad_async(G, dU, dWs, shifts=list(range(-4, 5)), istride=s, in_scale=b["a1"], in_shift=b["b1"],
                            in_relu=True, s_m=Cout * Cout, s_c2=1, s_co=Cout)
            grads[pre + "tcn.2.weight"] = dWs.permute(1, 2, 0).unsqueeze(-1)
            dH = torch.empty_like(G)
            if s == 1:
                ops.tapconv(dU, pp["wt_d"][0], dH, shifts=[4 - m for m in range(9)], tj=T)
            else:
                ops.tapconv(dU, pp["wt_d"][0], dH, shifts=[2, 1, 0, -1, -2], tj=(T + 1) // 2, ostride=2, ooff=0)
                if T // 2 > 0:
                    ops.tapconv(dU, pp["wt_d"][1], dH, shifts=[2, 1, 0, -1], tj=T // 2, ostride=2, ooff=1)

            # ---- BN1 + ReLU backward ----
            T12 = arena.f64(2 * NR * Cout)
            T1, T2 = T12[:NR * Cout], T12[NR * Cout:]
            ops.bn1_bwd_reduce(dH, G, b["a1"], b["b1"], T1, T2)
            if sy:
                sy[0].all_reduce(T12, group=sy[1])
            c1, c2, c3 = f32(Cout), f32(Cout), f32(Cout)
            dg1, db1n = arena.f32(Cout), arena.f32(Cout)
            ops.bn1_bwd_coef(T1, T2, b["a1"], b["mean1"], b["rstd1"], N * T * V * W, training, c1, c2, c3, dg1, db1n)
            grads[pre + "tcn.0.weight"], grads[pre + "tcn.0.bias"] = dg1, db1n
            glob += [dg1, db1n]
            dG = torch.empty_like(G)
            TblR = arena.f32(NR, V, Cout)
            ops.bn1_bwd_apply(dH, G, b["a1"], b["b1"], c1, c2, c3, dG, TblR)

            # ---- graph conv: wgrad, bias, dgrad through the weights, edge importance ----
            Wg = P[pre + "gcn.conv.weight"]
            dWg = arena.f32(K * Cout, Cin, 1, 1)
            if Xa is None:   # fused graph conv: the aggregated operand is re-derived in the wgrad prologue (csrc/gcn.cu)
                ops.gcn_wgrad(x, dG, dWg, csr["fwd_rowptr"], csr["fwd_src"], b["coef_f"], K, self.kdeg)
            else:
                wgrad_async(Xa, dG, dWg, shifts=[0], c2=Cin, s_m=0, s_c1=Cout * Cin, s_c2=1, s_co=Cin)
            grads[pre + "gcn.conv.weight"] = dWg
            fused_bwd = pp["fused_bwd"]
            dcoef = arena.f32(self.E)
            Pm = None
            if not fused_bwd:
                Pm = torch.empty(N, T, V, K * Cin, dtype=dt, device=dev)
                ops.tapconv(dG, pp["wg_b"], Pm, shifts=[0], tj=T)
            fused_dcoef = Cin % 8 == 0
            if not fused_dcoef:
                ops.agg_dcoef(x, Pm, dcoef, csr["fwd_src"], csr["dst"], csr["kk"], K)

            # ---- residual branch ----
            addend = None
            if reskind == "identity":
                addend = dY if premasked else dPre   # gradient of the identity shortcut = the masked dY
            elif reskind == "conv":
                Wr = P[pre + "residual.0.weight"]
                dWr = arena.f32(Cout, Cin, 1, 1)
                wgrad_async(x, dR, dWr, shifts=[0], istride=s, s_m=0, s_c2=1, s_co=Cin)
                grads[pre + "residual.0.weight"] = dWr
                grads[pre + "residual.0.bias"] = side(rep_sum, sum_dR, Cout)
                grads[pre + "residual.1.weight"], grads[pre + "residual.1.bias"] = dgr, dbr
                addend = torch.zeros_like(x)
                ops.tapconv(dR, pp["wr_d"], addend, shifts=[0], tj=To, ostride=s, ooff=0)

            dx = torch.empty_like(x)
            coef_b = b["coef_b"]
            mask_dx = fused_bwd and i > 0 and self.premask   # x of block i > 0 is the ReLU output of block i - 1
            if fused_bwd:
                # P = dG.W^T stays on chip: GEMM + transposed aggregation + edge-coefficient gradient in one kernel (csrc/gcn.cu)
                ops.gcn_bwd(dG, pp["wg_b"], dx, csr["bwd_rowptr"], csr["dst_b"],
                            csr["kk_b"], coef_b, K, self.max_out_deg, addend=addend, x=x, eid=csr["eid_b"], dcoef=dcoef,
                            relu_mask=mask_dx)
            elif fused_dcoef:
                ops.agg_bwd(Pm, addend, dx, csr["bwd_rowptr"], csr["dst_b"], csr["kk_b"], coef_b, K, x=x,
                            eid=csr["eid_b"], dcoef=dcoef)
            else:
                ops.agg_bwd(Pm, addend, dx, csr["bwd_rowptr"], csr["dst_b"], csr["kk_b"], coef_b, K)
            # d(conv bias) and d(edge importance) from the per-joint sums of dG and the per-edge sums: one launch
            dbg, dimp = arena.f32(K * Cout), arena.f32(K, V, V)
            side(ops.gcn_prep_bwd, A, P[pre + "gcn.conv.bias"], b["colsum"], TblR, dcoef, csr["dense_idx"], dbg, dimp)
            grads[pre + "gcn.conv.bias"] = dbg
            grads[f"edge_importance.{i}"] = dimp
            if self.debug is not None:
                self.debug[i] = dict(dY=dY, dU=dU, dH=dH, dG=dG, P=Pm, dx=dx, S1=S1, S2=S2, dp=dp, dR=dR, c1=c1, c2=c2,
                                     c3=c3, T1=T1, T2=T2, saved=b)
            dY = dx
            premasked = mask_dx

        if ws is not None:
            cur.wait_stream(ws)
            keep.clear()
        if sy:
            torch._foreach_mul_(glob, 1.0 / W)
        # ---- data_bn backward (input itself needs no gradient) ----
        xin, mean0, rstd0 = sv["data_bn"]
        VC = V * self.in_channels
        dgb = arena.f64(2 * VC)
        ops.databn_bwd(dY, xin, mean0, rstd0, dgb[:VC], dgb[VC:])
        grads["data_bn.weight"] = dgb[:VC].float()
        grads["data_bn.bias"] = dgb[VC:].float()
        return grads
