"""Differentiable operators of the hot path: ``torch.autograd.Function`` shells whose forward AND backward
are sequences of lib3dinfomax_b200 kernel launches (3dinfomax_b200/kernels.py).  autograd is used as the tape only.

The central one is ``fc``: the reference's FCLayer (models/base_layers.py:100-111) applied to a *virtual*
concatenation of K-segments, each optionally row-gathered and row-scaled, so that
``cat([h[src], h[dst], e])`` (models/pna.py:249) and ``cat([h, agg, agg*amp, agg*att])``
(models/pna.py:207,232) are never materialised.
"""
import os
import threading

import torch

from . import kernels as K

# ---------------------------------------------------------------------------------------------------------------
# Weight-gradient side stream.  dW = dY^T x of the big FC layers is a leaf of the backward graph: nothing downstream
# of it runs before the optimizer.  Its split-K tensor-core kernels occupy one CTA per SM with most of the shared
# memory but few warps, while the BatchNorm / aggregation / segment-sum backward kernels of the next layers are
# bandwidth-bound and light on shared memory — so the dW GEMMs are issued on a second stream (a parallel branch of
# the captured step graph) and joined once, when backward() ends.  Only taken when the gradient accumulates straight
# into FusedAdam's flat buffer (nothing is handed back to autograd).  I3D_DW_STREAM=0 keeps everything on one stream.
# ---------------------------------------------------------------------------------------------------------------
_dw_lock = threading.Lock()
_dw_streams = {}          # device index -> side stream
_dw_join_pending = set()  # (device index, autograd graph-task id): a join callback is queued for that backward()
DW_MIN_ROWS = int(os.environ.get("I3D_DW_MIN_ROWS", "2048"))     # smaller GEMMs stay on the main stream


def _dw_min_rows():
    return int(os.environ.get("I3D_DW_MIN_ROWS", DW_MIN_ROWS))


def _dw_fork(device, main, used):
    """Side stream ordered after everything enqueued on ``main`` so far, or None.  ``used``: tensors the side work
    reads that may be freed while it still runs (caching-allocator bookkeeping).  Backward nodes run on autograd's
    worker thread, the end-of-backward callback on the thread that called backward(): state is global, not
    thread-local, and the "join queued" flag is per backward() call (graph task), so concurrent backward passes of
    several trainers on one GPU (train.py --multithreaded_seeds) each get their own join."""
    if os.environ.get("I3D_DW_STREAM", "1") == "0":
        return None
    task = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else -1
    if task < 0:
        return None                       # not inside backward(): nothing would join the side stream
    key = (device.index, task)
    with _dw_lock:
        side = _dw_streams.get(device.index)
        if side is None:
            side = _dw_streams[device.index] = torch.cuda.Stream(device=device)
        queue = key not in _dw_join_pending
        _dw_join_pending.add(key)
    side.wait_stream(main)
    for t in used:
        if t is not None:
            t.record_stream(side)
    if queue:
        def join():
            with _dw_lock:
                _dw_join_pending.discard(key)
            torch.cuda.current_stream(device).wait_stream(side)

        torch.autograd.Variable._execution_engine.queue_callback(join)      # runs when this backward() finishes
    return side


def _grad_ready(param, stream):
    """tell the optimizer that ``param``'s weight-gradient kernels have been issued on ``stream`` (FusedAdam overlaps the
    data-parallel all-reduce of a layer's gradients with the backward of the layers below it)"""
    opt = getattr(param, "_i3d_optim", None)
    if opt is not None and opt.process_group is not None:
        dev = param.device
        streams = [torch.cuda.current_stream(dev)]
        side = _dw_streams.get(dev.index)          # earlier weights of the group may have gone to the side stream
        if side is not None:
            streams.append(side)
        if stream is not None and stream not in streams:
            streams.append(stream)
        opt.notify_grad_ready(param, streams)


class Seg:
    """One K-segment of an FC input.

    x            [R, K] fp32
    idx          optional int32 [M]: output row m reads x[idx[m]]
    scale        optional fp32 [M]: per-output-row multiplier (degree scaler)
    inv_rowptr   for gathered segments: CSR over R whose row r lists the output rows m with idx[m]==r
    inv_idx      optional position map of that CSR (None: the output rows of r are contiguous)
    """
    __slots__ = ("x", "idx", "scale", "inv_rowptr", "inv_idx")

    def __init__(self, x, idx=None, scale=None, inv_rowptr=None, inv_idx=None):
        if idx is not None and inv_rowptr is None:
            raise ValueError("a gathered segment needs its inverse CSR for the backward pass")
        if idx is not None and scale is not None:
            raise ValueError("gather and scale cannot be combined on one segment")
        self.x, self.idx, self.scale, self.inv_rowptr, self.inv_idx = x, idx, scale, inv_rowptr, inv_idx


class FCConfig:
    """valid: optional device int32 scalar — the number of valid leading rows when the batch is padded to a shape
    bucket (trainer.BucketedStep).  Rows beyond it stay out of the BatchNorm statistics, leave the normalisation as
    zeros and get an exactly-zero gradient (kernels.*: the *_v entry points of include/i3d.h)."""
    __slots__ = ("segs", "act", "has_bn", "training", "running_mean", "running_var", "nbt", "momentum", "eps", "valid")

    def __init__(self, segs, act, has_bn, training, running_mean=None, running_var=None, nbt=None, momentum=0.1,
                 eps=1e-5, valid=None):
        self.segs, self.act, self.has_bn, self.training = segs, act, has_bn, training
        self.running_mean, self.running_var, self.nbt = running_mean, running_var, nbt
        self.momentum, self.eps = momentum, eps
        self.valid = valid


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class _FC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, W, b, gamma, beta, residual, *xs):
        segs = cfg.segs
        M = segs[0].idx.numel() if segs[0].idx is not None else xs[0].shape[0]
        Fout = W.shape[0]
        Y = torch.empty(M, Fout, dtype=torch.float32, device=W.device)
        gsegs, off = [], 0
        for s, x in zip(segs, xs):
            k = x.shape[1]
            gsegs.append({"A": x, "B": W[:, off:off + k], "K": k, "a_idx": s.idx, "scale": s.scale})
            off += k
        if off != W.shape[1]:
            raise ValueError("segment widths %d do not add up to the weight's in_features %d" % (off, W.shape[1]))
        # weights prepared once per step by the trainer's WeightPrep (one launch for all layers) when available
        prep = getattr(W, "_i3d_prep", None)
        ready = None
        if prep is not None and K.gemm_nt_prepared_ok(M, Fout, gsegs):
            ent = prep.entry(W, ("f",) + tuple(g["K"] for g in gsegs), Fout, [(g["B"], g["K"]) for g in gsegs], False)
            ready = ent if prep.ready(ent) else None
        save = None
        sums = None
        if cfg.has_bn and cfg.training:
            # statistics from the GEMM epilogue
            _, sums = K.gemm(K.NT, M, Fout, gsegs, Y, bias=b, stats_act=cfg.act, prepared=ready,
                             arena=getattr(W, "_i3d_arena", None), valid=cfg.valid)
        else:
            K.gemm(K.NT, M, Fout, gsegs, Y, bias=b, prepared=ready)
        if cfg.has_bn:
            O, save = K.bn_apply(Y, cfg.act, sums, cfg.running_mean, cfg.running_var, cfg.nbt, gamma, beta,
                                 cfg.momentum, cfg.eps, cfg.training, residual, valid=cfg.valid)
        else:
            O = K.act_fwd(Y, cfg.act) if cfg.act != 0 else Y
            if residual is not None:
                O = K.add(O, residual)
        ctx.cfg = cfg
        ctx.M = M
        ctx.has_res = residual is not None
        ctx.w_param = W           # the Parameter object itself: backward looks for FusedAdam's gradient view on it
        ctx.save_for_backward(W, Y, save, gamma, *xs)
        return O

    @staticmethod
    def backward(ctx, dO):
        cfg = ctx.cfg
        W, Y, save, gamma = ctx.saved_tensors[:4]
        xs = ctx.saved_tensors[4:]
        segs = cfg.segs
        M, Fout = ctx.M, W.shape[0]
        if dO.dim() != 2 or dO.stride(1) != 1:
            dO = dO.contiguous()
        need_b = ctx.needs_input_grad[2]
        dgamma = dbeta = None
        if cfg.has_bn:
            arena = getattr(ctx.w_param, "_i3d_arena", None)
            dbz = K.dbias_buffer(Fout, W.device) if need_b else None
            # Linear -> (no activation) -> train-mode BatchNorm: the bias is cancelled by the mean subtraction, its
            # gradient is EXACTLY zero (sum over rows of gr (dO - mean(dO) - xhat mean(dO xhat)) = 0); the reference's
            # value is the fp32 rounding noise of that sum.  dbz was zeroed by bn_bwd_reduce: no second reduction needed.
            zero_db = need_b and cfg.training and cfg.act == 0
            dY, db, dgamma, dbeta = K.bn_bwd(dO, Y, cfg.act, cfg.training, save, gamma, need_b and not zero_db,
                                             dbias_zeroed=dbz, valid=cfg.valid, arena=arena)
            if zero_db:
                db = dbz[:Fout]          # (contiguous zeros: autograd keeps it as .grad without a layout copy)
        elif cfg.act != 0 or cfg.valid is not None:
            # (padded batches: the pass also writes the exact zeros of the padding rows)
            dY, db, _, _ = K.bn_bwd_apply(dO, Y, cfg.act, False, False, None, None, None, need_b, valid=cfg.valid)
        else:
            dY = dO
            db = K.colsum(dO) if need_b else None
        sums_cache = {}

        def row_sums(i, s):
            """[R, Fout]: for every row r of a gathered segment's source, the sum of dY over the output rows reading it"""
            if i not in sums_cache:
                sums_cache[i] = K.segment_sum_fwd(dY, s.inv_rowptr, s.inv_idx)
            return sums_cache[i]

        # weight gradient, one column block per segment: dW[:, off:off+k] = dY^T (scale * gather(x))
        dW = None
        if ctx.needs_input_grad[1]:
            # FusedAdam gives FC weights a view of its flat (pre-zeroed) gradient buffer: accumulate straight into it
            # and hand autograd nothing; otherwise one memset and the split-K column-block GEMMs accumulate into dW
            direct = getattr(ctx.w_param, "_i3d_grad_view", None)
            acc = direct if direct is not None else torch.zeros_like(W)
            # dY^T x[idx] == (sum of the dY rows that read each x row)^T x: reduce over the gather first (the same row
            # sums feed dx below), then a GEMM over the R source rows instead of the M gathered ones
            via_sums = [s.idx is not None and x.shape[0] < M for s, x in zip(segs, xs)]
            a_ops = [row_sums(i, s) if v else dY for i, (v, s) in enumerate(zip(via_sums, segs))]
            side = None
            if direct is not None and M >= _dw_min_rows() and Fout >= 64:
                side = _dw_fork(W.device, torch.cuda.current_stream(W.device), [dY] + a_ops + list(xs))
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                off = 0
                for i, (s, x) in enumerate(zip(segs, xs)):
                    k = x.shape[1]
                    if via_sums[i]:
                        K.gemm(K.TN, Fout, k, [{"A": a_ops[i], "B": x, "K": x.shape[0]}], acc[:, off:off + k],
                               accumulate=True)
                    else:
                        K.gemm(K.TN, Fout, k, [{"A": dY, "B": x, "K": M, "b_idx": s.idx, "scale": s.scale}],
                               acc[:, off:off + k], accumulate=True)
                    off += k
            dW = None if direct is not None else acc
            if direct is not None:
                _grad_ready(ctx.w_param, side)
        # input gradients, one NN GEMM per distinct input tensor (segments sharing a tensor are K-segments of it)
        dxs = [None] * len(xs)
        groups = {}
        off = 0
        for i, (s, x) in enumerate(zip(segs, xs)):
            k = x.shape[1]
            if ctx.needs_input_grad[6 + i]:
                groups.setdefault(x.data_ptr(), []).append((i, s, x, off))
            off += k
        Wt = None
        prep = getattr(ctx.w_param, "_i3d_prep", None)
        for members in groups.values():
            x0 = members[0][2]
            R, k = x0.shape
            dx = torch.empty(R, k, dtype=torch.float32, device=W.device)
            # dx = dy W is fed to the tensor-core NT kernel as dy (W^T)^T: the transposed hi/lo copies come from the
            # per-step WeightPrep when there is one, else from one small transpose of the weight per backward
            via_nt = R >= 256 and Fout % 4 == 0 and k % 4 == 0
            nn = []
            for (_i, s, _x, o) in members:
                a = dY if s.idx is None else row_sums(_i, s)
                nn.append({"A": a, "K": Fout, "scale": s.scale, "_o": o})
            ready = None
            if via_nt and prep is not None and K.gemm_nt_prepared_ok(R, k, nn):
                ent = prep.entry(ctx.w_param, ("b", k) + tuple(m[3] for m in members), k,
                                 [(W[:, m[3]:m[3] + k], Fout) for m in members], True)
                ready = ent if prep.ready(ent) else None
            if ready is None:
                if via_nt and Wt is None:
                    Wt = K.transpose(W)                                  # [in_features, out_features]
                for g in nn:
                    g["B"] = Wt[g["_o"]:g["_o"] + k, :] if via_nt else W[:, g["_o"]:g["_o"] + k]
            K.gemm(K.NT if via_nt else K.NN, R, k, nn, dx, prepared=ready)
            dxs[members[0][0]] = dx
        dres = dO if (ctx.has_res and ctx.needs_input_grad[5]) else None
        return (None, dW, db, dgamma, dbeta, dres) + tuple(dxs)


def fc(segs, W, b, act, bn=None, training=True, residual=None, valid=None):
    """FCLayer over a virtual concat.  ``bn``: None or (gamma, beta, running_mean, running_var, nbt, momentum, eps).
    ``valid``: device int32 scalar, valid leading rows of a padded batch (FCConfig)."""
    if bn is None:
        cfg = FCConfig(segs, act, False, training, valid=valid)
        gamma = beta = None
    else:
        gamma, beta, rm, rv, nbt, mom, eps = bn
        cfg = FCConfig(segs, act, True, training, rm, rv, nbt, mom, eps, valid=valid)
    return _FC.apply(cfg, W, b, gamma, beta, residual, *[s.x for s in segs])


class _FCPostMerged(torch.autograd.Function):
    """FCLayer over cat[h, agg, agg*amp_D, agg*att_D] (models/pna.py:207-211,232) evaluated as [h | agg] @ Wm_D^T with
    one merged weight per in-degree bucket (kernels.DegreePlan / MergedPosttransWeights): K = 5F instead of 13F in the
    forward, in dx and in dW.  Same tail as ``_FC`` (activation -> BatchNorm, residual)."""

    @staticmethod
    def forward(ctx, cfg, plan, merged, W, b, gamma, beta, residual, h, agg):
        N, F = h.shape
        Fout = W.shape[0]
        if agg.shape != (N, 4 * F) or W.shape[1] != 13 * F:
            raise ValueError("merged posttrans expects agg [N,4F] and a [Fout,13F] weight")
        merged.refresh(W)
        Y = torch.empty(N, Fout, dtype=torch.float32, device=W.device)
        segs = [{"A": h, "K": F, "a_idx": plan.perm}, {"A": agg, "K": 4 * F, "a_idx": plan.perm}]
        sums = save = None
        if cfg.has_bn and cfg.training:
            _, sums = K.gemm_nt_bucketed(plan, Fout, segs, Y, b, merged.fwd_hi, merged.fwd_lo, stats_act=cfg.act,
                                         arena=getattr(W, "_i3d_arena", None), valid=cfg.valid)
        else:
            K.gemm_nt_bucketed(plan, Fout, segs, Y, b, merged.fwd_hi, merged.fwd_lo)
        if cfg.has_bn:
            O, save = K.bn_apply(Y, cfg.act, sums, cfg.running_mean, cfg.running_var, cfg.nbt, gamma, beta,
                                 cfg.momentum, cfg.eps, cfg.training, residual, valid=cfg.valid)
        else:
            O = K.act_fwd(Y, cfg.act) if cfg.act != 0 else Y
            if residual is not None:
                O = K.add(O, residual)
        ctx.cfg, ctx.plan, ctx.merged, ctx.w_param = cfg, plan, merged, W
        ctx.has_res = residual is not None
        ctx.res_is_h = residual is h
        ctx.save_for_backward(W, Y, save, gamma, h, agg)
        return O

    @staticmethod
    def backward(ctx, dO):
        cfg, plan, merged = ctx.cfg, ctx.plan, ctx.merged
        W, Y, save, gamma, h, agg = ctx.saved_tensors
        N, F = h.shape
        Fout = W.shape[0]
        if dO.dim() != 2 or dO.stride(1) != 1:
            dO = dO.contiguous()
        need_b = ctx.needs_input_grad[4]
        dgamma = dbeta = None
        if cfg.has_bn:
            arena = getattr(ctx.w_param, "_i3d_arena", None)
            dbz = K.dbias_buffer(Fout, W.device) if need_b else None
            # Linear -> (no activation) -> train-mode BatchNorm: the bias is cancelled by the mean subtraction, its
            # gradient is EXACTLY zero (sum over rows of gr (dO - mean(dO) - xhat mean(dO xhat)) = 0); the reference's
            # value is the fp32 rounding noise of that sum.  dbz was zeroed by bn_bwd_reduce: no second reduction needed.
            zero_db = need_b and cfg.training and cfg.act == 0
            dY, db, dgamma, dbeta = K.bn_bwd(dO, Y, cfg.act, cfg.training, save, gamma, need_b and not zero_db,
                                             dbias_zeroed=dbz, valid=cfg.valid, arena=arena)
            if zero_db:
                db = dbz[:Fout]          # (contiguous zeros: autograd keeps it as .grad without a layout copy)
        elif cfg.act != 0 or cfg.valid is not None:
            dY, db, _, _ = K.bn_bwd_apply(dO, Y, cfg.act, False, False, None, None, None, need_b, valid=cfg.valid)
        else:
            dY = dO
            db = K.colsum(dO) if need_b else None
        dW = None
        if ctx.needs_input_grad[3]:
            direct = getattr(ctx.w_param, "_i3d_grad_view", None)
            acc = direct if direct is not None else torch.zeros_like(W)
            side = None
            if direct is not None and N >= _dw_min_rows() and Fout >= 64:
                side = _dw_fork(W.device, torch.cuda.current_stream(W.device), [dY, h, agg])
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                K.gemm(K.TN, Fout, F, [{"A": dY, "B": h, "K": N}], acc[:, :F], accumulate=True)
                dWb = torch.zeros(plan.n_buckets, Fout, 4 * F, dtype=torch.float32, device=W.device)
                K.gemm_tn_chunked(plan, dY, agg, dWb)
                K.posttrans_unmerge(dWb, acc, F)
            dW = None if direct is not None else acc
            if direct is not None:
                _grad_ready(ctx.w_param, side)
        dh = dagg = None
        if ctx.needs_input_grad[8] or ctx.needs_input_grad[9]:
            # d[h | agg] = dY Wm_D, rows scattered back to node order by the epilogue
            buf = torch.empty(N, 5 * F, dtype=torch.float32, device=W.device)
            K.gemm_nt_bucketed(plan, 5 * F, [{"A": dY, "K": Fout, "a_idx": plan.perm}], buf, None, merged.bwd_hi,
                               merged.bwd_lo)
            dh, dagg = buf[:, :F], buf[:, F:]
        dres = dO if (ctx.has_res and ctx.needs_input_grad[7]) else None
        if dres is not None and dh is not None and ctx.res_is_h:
            # the residual IS the layer input h (models/pna.py:211): both gradients go to the same tensor, summed here by
            # one kernel instead of autograd's strided add (dh is a column slice of the [N, 5F] buffer)
            dh, dres = K.add_rows(dh, dres), None
        return None, None, None, dW, db, dgamma, dbeta, dres, dh, dagg


def fc_post_merged(plan, merged, h, agg, W, b, act, bn=None, training=True, residual=None, valid=None):
    """Degree-merged posttrans FCLayer; arguments as ``fc``."""
    if bn is None:
        cfg = FCConfig(None, act, False, training, valid=valid)
        gamma = beta = None
    else:
        gamma, beta, rm, rv, nbt, mom, eps = bn
        cfg = FCConfig(None, act, True, training, rm, rv, nbt, mom, eps, valid=valid)
    return _FCPostMerged.apply(cfg, plan, merged, W, b, gamma, beta, residual, h, agg)


class EdgeCodes:
    """Per-batch index tensors of the factored edge layer: CSR gather lists + the bond-feature code of every edge."""
    __slots__ = ("src_csr", "dst_csr", "rowptr", "out_rowptr", "out_pos", "code32", "code64", "col_off", "n_codes")

    def __init__(self, st, code_csr, n_codes):
        self.src_csr, self.dst_csr, self.rowptr = st.src_csr, st.dst_csr, st.rowptr
        self.out_rowptr, self.out_pos = st.out_rowptr, st.out_pos
        self.code64 = code_csr.reshape(-1, 1).contiguous()
        self.code32 = code_csr.to(torch.int32)
        self.col_off = torch.zeros(1, dtype=torch.int32, device=code_csr.device)
        self.n_codes = int(n_codes)


class _BondTables(torch.autograd.Function):
    """T_l = combo W_e,l^T for every message-passing layer l (W_e,l = columns [col0, col0+F) of the layer's first edge-MLP
    weight): the `e` K-segment of cat[h[src], h[dst], e] W^T as a table over the distinct bond-feature combinations
    (``combo`` [n_codes, F] = their embeddings).  One launch for all layers, forward and backward (i3d_bond.cu)."""

    @staticmethod
    def forward(ctx, col0, weight_grads, combo, *Ws):
        combo = combo.contiguous()
        Ts = K.bond_tables_fwd(combo, list(Ws), col0)
        ctx.col0, ctx.weight_grads = col0, weight_grads
        ctx.w_params = Ws
        ctx.save_for_backward(combo, *Ws)
        return tuple(Ts)

    @staticmethod
    def backward(ctx, *dTs):
        combo = ctx.saved_tensors[0]
        Ws = list(ctx.saved_tensors[1:])
        dTs = [None if d is None else d.contiguous() for d in dTs]
        # FC weights owned by FusedAdam accumulate straight into their slice of the flat gradient buffer
        direct = [getattr(w, "_i3d_grad_view", None) for w in ctx.w_params]
        need = [n and ctx.weight_grads for n in ctx.needs_input_grad[3:]]
        accs = [d if (d is not None and n) else (torch.zeros_like(w) if n else None) for d, w, n in zip(direct, Ws, need)]
        dcombo = K.bond_tables_bwd(combo, Ws, ctx.col0, dTs, accs, want_dcombo=ctx.needs_input_grad[2])
        grads = [None if (d is not None or not n) else a for d, a, n in zip(direct, accs, need)]
        return (None, None, dcombo) + tuple(grads)


def bond_tables(combo, weights, col0, weight_grads=True):
    """list of per-layer tables T_l [n_codes, Fout] (see _BondTables).  weight_grads=False: the W_e columns of the
    weight gradients are accumulated by the layers themselves (fc_edge_factored(..., combo=combo))."""
    return list(_BondTables.apply(int(col0), bool(weight_grads), combo, *weights))


class _FCEdgeFactored(torch.autograd.Function):
    """First FCLayer of the edge MLP over cat[h[src], h[dst], e] (models/pna.py:249-252) in factored form:
    (h W_s^T)[src] + (h W_d^T)[dst] + T[code] with P = h [W_s; W_d]^T ONE node-level GEMM (N rows, K = F) and
    T = combo W_e^T the layer's bond-feature table (ops.bond_tables).  The edge-level work left is one gather-add pass
    with the BatchNorm statistics fused (i3d_edge_gather_add).
    Backward: dP = per-node row sums of dY over the out- / in-edges (what the h[src], h[dst] gathers need anyway),
    dT = dY summed per code; every GEMM of the layer is node-level.  The weight gradient returned / accumulated here
    covers the W_s, W_d columns; the W_e columns get theirs from _BondTables.backward."""

    @staticmethod
    def forward(ctx, cfg, g, W, b, gamma, beta, h, T, combo):
        N, F = h.shape
        Fout = W.shape[0]
        M = g.src_csr.numel()
        if W.shape[1] != 3 * F or T.shape != (g.n_codes, Fout):
            raise ValueError("factored edge layer expects a [Fout, 3F] weight and a [n_codes, Fout] table")
        dev = W.device
        P = torch.empty(N, 2 * Fout, dtype=torch.float32, device=dev)
        prep = getattr(W, "_i3d_prep", None)
        ready = None
        if prep is not None and K.gemm_nt_prepared_ok(N, 2 * Fout, [{"A": h, "K": F}]):
            ent = prep.entry_stacked(W, (F,), [W[:, :F], W[:, F:2 * F]])
            ready = ent if prep.ready(ent) else None
        if ready is not None:
            K.gemm(K.NT, N, 2 * Fout, [{"A": h, "K": F}], P, prepared=ready)
        else:
            K.gemm(K.NT, N, Fout, [{"A": h, "B": W[:, :F], "K": F}], P[:, :Fout])
            K.gemm(K.NT, N, Fout, [{"A": h, "B": W[:, F:2 * F], "K": F}], P[:, Fout:])
        Y = torch.empty(M, Fout, dtype=torch.float32, device=dev)
        want_stats = cfg.has_bn and cfg.training
        sums = K.edge_gather_add(P, g.src_csr, g.dst_csr, T.contiguous(), g.code32, b, Y,
                                 stats_act=cfg.act if want_stats else None, arena=getattr(W, "_i3d_arena", None),
                                 valid=cfg.valid)
        save = None
        if cfg.has_bn:
            O, save = K.bn_apply(Y, cfg.act, sums, cfg.running_mean, cfg.running_var, cfg.nbt, gamma, beta,
                                 cfg.momentum, cfg.eps, cfg.training, None, valid=cfg.valid)
        else:
            O = K.act_fwd(Y, cfg.act) if cfg.act != 0 else Y
        ctx.cfg, ctx.g, ctx.w_param = cfg, g, W
        ctx.save_for_backward(W, Y, save, gamma, h, combo)
        return O

    @staticmethod
    def backward(ctx, dO):
        cfg, g = ctx.cfg, ctx.g
        W, Y, save, gamma, h, combo = ctx.saved_tensors
        N, F = h.shape
        Fout = W.shape[0]
        dev = W.device
        if dO.dim() != 2 or dO.stride(1) != 1:
            dO = dO.contiguous()
        need_b = ctx.needs_input_grad[3]
        dgamma = dbeta = None
        arena = getattr(ctx.w_param, "_i3d_arena", None)
        if cfg.has_bn:
            dbz = K.dbias_buffer(Fout, dev) if need_b else None
            # Linear -> (no activation) -> train-mode BatchNorm: the bias is cancelled by the mean subtraction, its
            # gradient is EXACTLY zero (sum over rows of gr (dO - mean(dO) - xhat mean(dO xhat)) = 0); the reference's
            # value is the fp32 rounding noise of that sum.  dbz was zeroed by bn_bwd_reduce: no second reduction needed.
            zero_db = need_b and cfg.training and cfg.act == 0
            dY, db, dgamma, dbeta = K.bn_bwd(dO, Y, cfg.act, cfg.training, save, gamma, need_b and not zero_db,
                                             dbias_zeroed=dbz, valid=cfg.valid, arena=arena)
            if zero_db:
                db = dbz[:Fout]          # (contiguous zeros: autograd keeps it as .grad without a layout copy)
        elif cfg.act != 0 or cfg.valid is not None:
            dY, db, _, _ = K.bn_bwd_apply(dO, Y, cfg.act, False, False, None, None, None, need_b, valid=cfg.valid)
        else:
            dY = dO
            db = K.colsum(dO) if need_b else None
        # dP: per-node sums of dY over the node's out-edges (h[src] term) and in-edges (h[dst] term)
        Rs = K.segment_sum_fwd(dY, g.out_rowptr, g.out_pos)
        Rd = K.segment_sum_fwd(dY, g.rowptr, None)
        # dT: dY summed per bond-feature code (shared-memory pre-reduction of the embedding-gradient kernel)
        dT = K.embed_sum_bwd(g.code64, g.col_off, None, dY, g.n_codes, g.n_codes) if ctx.needs_input_grad[7] else None
        dW = None
        if ctx.needs_input_grad[2]:
            direct = getattr(ctx.w_param, "_i3d_grad_view", None)
            acc = direct if direct is not None else torch.zeros_like(W)
            side = None
            if direct is not None and N >= _dw_min_rows() and Fout >= 64:
                side = _dw_fork(dev, torch.cuda.current_stream(dev), [Rs, Rd, h, dT, combo])
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                K.gemm(K.TN, Fout, F, [{"A": Rs, "B": h, "K": N}], acc[:, :F], accumulate=True)
                K.gemm(K.TN, Fout, F, [{"A": Rd, "B": h, "K": N}], acc[:, F:2 * F], accumulate=True)
                if dT is not None and combo is not None:
                    # the W_e columns: dT^T combo.  Done here, per layer, so that the layer's whole weight gradient is
                    # final when this backward returns (the data-parallel all-reduce of the layer starts right away)
                    K.bond_tables_bwd(combo, [W], 2 * F, [dT], [acc], want_dcombo=False)
            dW = None if direct is not None else acc
            if direct is not None:
                _grad_ready(ctx.w_param, side)
        dh = None
        if ctx.needs_input_grad[6]:
            dh = torch.empty(N, F, dtype=torch.float32, device=dev)
            nn = [{"A": Rs, "K": Fout}, {"A": Rd, "K": Fout}]
            prep = getattr(ctx.w_param, "_i3d_prep", None)
            via_nt = N >= 256 and Fout % 4 == 0 and F % 4 == 0
            ready = None
            if via_nt and prep is not None and K.gemm_nt_prepared_ok(N, F, nn):
                ent = prep.entry(ctx.w_param, ("b", F, 0, F), F, [(W[:, :F], Fout), (W[:, F:2 * F], Fout)], True)
                ready = ent if prep.ready(ent) else None
            if ready is None:
                if via_nt:
                    Wt = K.transpose(W[:, :2 * F])                         # [2F, Fout]
                    nn[0]["B"], nn[1]["B"] = Wt[:F, :], Wt[F:, :]
                else:
                    nn[0]["B"], nn[1]["B"] = W[:, :F], W[:, F:2 * F]
            K.gemm(K.NT if via_nt else K.NN, N, F, nn, dh, prepared=ready)
        return None, None, dW, db, dgamma, dbeta, dh, dT, None


def fc_edge_factored(g, h, T, W, b, act, bn=None, training=True, valid=None, combo=None):
    """FCLayer over cat[h[src], h[dst], e] in factored form (``g``: EdgeCodes, ``T``: the layer's table from
    ``bond_tables``); other arguments as ``fc``.  ``combo``: the combination embeddings the tables were built from —
    when given, this layer's backward also accumulates the W_e columns of the weight gradient (dT^T combo) and
    ``bond_tables`` must be called with ``weight_grads=False``."""
    if bn is None:
        cfg = FCConfig(None, act, False, training, valid=valid)
        gamma = beta = None
    else:
        gamma, beta, rm, rv, nbt, mom, eps = bn
        cfg = FCConfig(None, act, True, training, rm, rv, nbt, mom, eps, valid=valid)
    return _FCEdgeFactored.apply(cfg, g, W, b, gamma, beta, h, T, None if combo is None else combo.detach())


class _EmbedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, col_off, perm, table, max_dim):
        ctx.save_for_backward(idx, col_off, perm)
        ctx.rows = table.shape[0]
        ctx.max_dim = max_dim
        return K.embed_sum_fwd(idx, col_off, perm, table)

    @staticmethod
    def backward(ctx, g):
        idx, col_off, perm = ctx.saved_tensors
        return None, None, None, K.embed_sum_bwd(idx, col_off, perm, g, ctx.rows, ctx.max_dim), None


def embed_sum(idx, col_off, perm, table, max_dim=0):
    """max_dim: largest per-column vocabulary (lets the backward pre-reduce in shared memory)."""
    return _EmbedSum.apply(idx, col_off, perm, table, int(max_dim))


class _PNAAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msg, rowptr):
        out = K.pna_aggregate_fwd(msg, rowptr)
        ctx.save_for_backward(msg, out, rowptr)
        return out

    @staticmethod
    def backward(ctx, g):
        msg, out, rowptr = ctx.saved_tensors
        if g.stride(1) != 1:
            g = g.contiguous()
        return K.pna_aggregate_bwd(g, msg, out, rowptr), None


def pna_aggregate(msg, rowptr):
    """[E,F] CSR-ordered messages -> [N,4F] = [mean|max|min|std] (models/pna.py:17-37,221-235)."""
    return _PNAAggregate.apply(msg, rowptr)


class _Readout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ptr, ops):
        out = K.segment_readout_fwd(x, ptr, ops)
        ctx.ops = ops
        ctx.save_for_backward(x, out, ptr)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, ptr = ctx.saved_tensors
        return K.segment_readout_bwd(g, x, out, ptr, ctx.ops), None, None


def readout(x, graph_ptr, ops):
    """cat([dgl.readout_nodes(g,'feat',op) for op in ops]) (models/pna.py:133-134)."""
    return _Readout.apply(x, graph_ptr, tuple(K.RO[o] for o in ops))


class _SegmentReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rowptr, rowid, mean, addend):
        ctx.mean = mean
        ctx.has_add = addend is not None
        ctx.save_for_backward(rowptr, rowid)
        return K.segment_sum_fwd(x, rowptr, None, mean, addend)

    @staticmethod
    def backward(ctx, g):
        rowptr, rowid = ctx.saved_tensors
        g = g.contiguous()
        return K.segment_sum_bwd(g, rowptr, rowid, ctx.mean), None, None, None, (g if ctx.has_add else None)


def segment_reduce(x_csr, rowptr, rowid, mean, addend=None):
    """fn.sum / fn.mean over in-edges (models/net3d.py:94-96) fused with the ``m_sum + feat`` add (net3d.py:122)."""
    return _SegmentReduce.apply(x_csr, rowptr, rowid, bool(mean), addend)


class _SoftGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msg, ws, bs):
        m, w = K.soft_gate_fwd(msg, ws, bs)
        ctx.save_for_backward(msg, w, ws)
        return m

    @staticmethod
    def backward(ctx, gm):
        msg, w, ws = ctx.saved_tensors
        gmsg, gws, gbs = K.soft_gate_bwd(gm, msg, w, ws)
        return gmsg, gws.view_as(ws), gbs


def soft_gate(msg, weight, bias):
    """message * sigmoid(soft_edge_network(message)) (models/net3d.py:117-118)."""
    return _SoftGate.apply(msg, weight, bias)


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        ctx.act = act
        ctx.save_for_backward(x)
        return K.act_fwd(x, act)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return K.act_bwd(g, x, ctx.act), None


def activation(x, name):
    return _Act.apply(x, K.ACT[name])


class _Broadcast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, M):
        return K.broadcast_rows(vec, M)

    @staticmethod
    def backward(ctx, g):
        return K.colsum(g.contiguous()), None


def broadcast_rows(vec, M):
    """node_embedding[None, :].expand(N, -1) (models/net3d.py:61)."""
    return _Broadcast.apply(vec, M)


class _ScaleRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(s)
        return K.scale_rows(x, s)

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        if g.stride(1) != 1:
            g = g.contiguous()
        return K.scale_rows(g, s), None


def scale_rows(x, s):
    """x * s[:, None] with a per-row factor that needs no gradient (graph_norm, models/pna_original.py:258-259)."""
    return _ScaleRows.apply(x, s.reshape(-1).contiguous())


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return K.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a, b)


class _NTXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, C, tau, norm, eps, row_offset, inv_B):
        z1 = z1.contiguous()
        z2 = z2.contiguous()
        B, D = z1.shape
        Bc = z2.shape[0] // C
        n1 = K.row_norms(z1) if norm else None
        n2 = K.row_norms(z2) if norm else None
        P = torch.empty(B, Bc * C, dtype=torch.float32, device=z1.device)
        K.gemm(K.NT, B, Bc * C, [{"A": z1, "B": z2, "K": D}], P)
        rowstats, loss_rows = K.ntxent_rows_fwd(P, B, Bc, C, n1, n2, norm, eps, tau, row_offset)
        loss = K.sum_scaled(loss_rows, inv_B)
        ctx.args = (B, Bc, C, D, tau, norm, eps, row_offset, inv_B)
        ctx.save_for_backward(z1, z2, n1, n2, P, rowstats)
        return loss

    @staticmethod
    def backward(ctx, gout):
        z1, z2, n1, n2, P, rowstats = ctx.saved_tensors
        B, Bc, C, D, tau, norm, eps, row_offset, inv_B = ctx.args
        gout = gout.contiguous().float()
        if getattr(ctx, "consumed", False):
            raise RuntimeError("NTXent backward overwrites its saved probabilities in place: call backward once")
        ctx.consumed = True
        G = P
        dn1, dn2 = K.ntxent_rows_bwd(G, B, Bc, C, n1, n2, norm, eps, tau, row_offset, rowstats, gout, inv_B)
        dz1 = torch.empty_like(z1)
        if B >= 256 and (Bc * C) % 4 == 0:
            K.gemm(K.NT, B, D, [{"A": G, "B": K.transpose(z2), "K": Bc * C}], dz1)     # tensor-core NT path
        else:
            K.gemm(K.NN, B, D, [{"A": G, "B": z2, "K": Bc * C}], dz1)
        dz2 = torch.empty_like(z2)
        K.gemm(K.TN, Bc * C, D, [{"A": G, "B": z1, "K": B}], dz2)
        if norm:
            K.norm_bwd_accum(z1, n1, dn1, dz1)
            K.norm_bwd_accum(z2, n2, dn2, dz2)
        return dz1, dz2, None, None, None, None, None, None


def ntxent(z1, z2, conformers, tau, norm, eps, row_offset=0, total_rows=None):
    """-mean_i log(pos_i / neg_i) over exp(sim/tau) (commons/losses.py:143-155, 225-246).

    z1 [B,D] are the local rows, z2 [Bc*C, D] the (possibly all-gathered) columns, molecule-major."""
    B = z1.shape[0]
    inv_B = 1.0 / float(total_rows if total_rows is not None else B)
    return _NTXent.apply(z1, z2, int(conformers), float(tau), bool(norm), float(eps), int(row_offset), inv_B)
