"""Fused Adam over flat parameter storage (trainer/trainer.py:119-123 ``optim.step()``; torch.optim.Adam maths).

The reference builds ``torch.optim.Adam`` with two parameter groups — BatchNorm parameters with
``weight_decay=0`` and everything else (trainer/self_supervised_trainer.py:78-86) — and steps ~190 small tensors
with a foreach loop.  Here each group's parameters are re-homed as views into ONE flat fp32 buffer (the
``nn.Parameter`` objects and their state-dict keys are untouched), so a step is:

    one multi-tensor pack of the gradients -> [one NCCL all-reduce when data parallel] -> one Adam kernel per group.

FC weights (flagged ``_i3d_direct_grad`` by base_layers._Linear) skip even the pack: their ``.grad`` IS a view of the
flat gradient buffer and ``ops._FC.backward`` accumulates the weight-gradient GEMM into it, so ``zero_grad()`` is one
memset per group and must be called between steps (the reference loop does, trainer/trainer.py:121-123).

``param_groups`` keeps torch's layout (list of dicts with 'params', 'lr', 'betas', 'eps', 'weight_decay'), so the
reference's ``WarmUpWrapper`` (trainer/lr_schedulers.py), which rewrites ``group['lr']`` every step, drives it as is.
"""
import os

import torch

from . import kernels as K


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 graph_safe=False):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = []
        self.process_group = process_group
        self.graph_safe = graph_safe
        self._flat = []
        # data parallel on NVSwitch: the flat buffers live in symmetric (multicast) memory and the step is ONE kernel that
        # reduces the gradients through the switch, applies Adam to this rank's slice and multicast-stores the result
        # (i3d_adam_step_nvls); I3D_NVLS_ADAM=0 or no multicast support -> NCCL all-reduce + the local Adam kernel
        self.nvls = process_group is not None and os.environ.get("I3D_NVLS_ADAM", "1") != "0"
        for g in groups:
            pg = dict(self.defaults)
            pg.update(g)
            pg["params"] = [p for p in pg["params"]]
            self.param_groups.append(pg)
            self._flat.append(self._flatten(pg["params"], process_group if self.nvls else None))
        if self.nvls and not all(fl is None or fl.get("mc") for fl in self._flat):
            self.nvls = False
        self._step = 0
        self._needs_zero = False
        self.post_step_hooks = []          # e.g. WeightPrep.invalidate: the Adam kernel bypasses tensor versions
        dev = self._device()
        self._step_dev = torch.ones(1, dtype=torch.int64, device=dev) if graph_safe else None
        self._hyper_dev = [torch.zeros(6, dtype=torch.float64, device=dev) for _ in self.param_groups] \
            if graph_safe else None
        self._hyper_host = [None] * len(self.param_groups)
        # gradient all-reduce overlapped with backward (data parallel): see set_overlap_groups
        self._overlap = []            # [(flat index, start, end, [params])]
        self._overlap_of = {}         # id(param) -> overlap group index
        self._pending = {}            # overlap group index -> ids of the params whose gradient is not final yet
        self._reduced = []            # [(flat index, start, end)] ranges already all-reduced in this step
        self._comm_stream = None

    def _device(self):
        for g in self.param_groups:
            for p in g["params"]:
                return p.device
        return torch.device("cuda")

    @staticmethod
    def _symmetric(n, dev, group):
        """[4n] fp32 in symmetric memory bound to a multicast address, or (None, None, None) when unavailable"""
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(4 * n, dtype=torch.float32, device=dev)
            h = symm.rendezvous(t, group)
            if not h.has_multicast_support or not h.multicast_ptr:
                return None, None, None
            t.zero_()
            return t, h, int(h.multicast_ptr)
        except Exception:                     # no symmetric-memory support in this build / topology: NCCL path
            return None, None, None

    @staticmethod
    def _flatten(params, symm_group=None):
        if not params:
            return None
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs CUDA parameters: the 3dinfomax_b200 path has no CPU fallback")
        sizes = [p.numel() for p in params]
        # flat layout: the directly-accumulated FC weights first (in parameter order, so the weights of one message-
        # passing layer are ONE contiguous range that can be all-reduced as soon as that layer's backward is done), the
        # packed small tensors (biases, BatchNorm, embeddings) behind them.  Offsets are kept per parameter index.
        order = [i for i, p in enumerate(params) if getattr(p, "_i3d_direct_grad", False)] + \
                [i for i, p in enumerate(params) if not getattr(p, "_i3d_direct_grad", False)]
        offs, acc = [0] * len(params), 0
        for i in order:
            offs[i] = acc
            acc += (sizes[i] + 3) // 4 * 4   # keep every view 16-byte aligned for the vectorised kernels
        sym, handle, mc = (None, None, None)
        if symm_group is not None:
            sym, handle, mc = FusedAdam._symmetric(acc, dev, symm_group)
        flat = torch.zeros(acc, dtype=torch.float32, device=dev) if sym is None else sym[:acc]
        with torch.no_grad():
            for p, o, n in zip(params, offs, sizes):
                if p.dtype != torch.float32:
                    raise TypeError("fp32 parameters only")
                flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = flat[o:o + n].view(p.shape)
        gflat = torch.zeros_like(flat) if sym is None else sym[acc:2 * acc]
        direct = []
        for p, o, n in zip(params, offs, sizes):
            d = bool(getattr(p, "_i3d_direct_grad", False))
            direct.append(d)
            if d:
                p._i3d_grad_view = gflat[o:o + n].view(p.shape)
                p.grad = p._i3d_grad_view
        mflat = torch.zeros_like(flat) if sym is None else sym[2 * acc:3 * acc]
        vflat = torch.zeros_like(flat) if sym is None else sym[3 * acc:4 * acc]
        return {"p": flat, "g": gflat, "m": mflat, "v": vflat, "direct": direct, "sym": sym, "handle": handle, "mc": mc,
                "off": torch.tensor(offs, dtype=torch.int64, device=dev),
                "len": torch.tensor(sizes, dtype=torch.int64, device=dev), "offs": offs, "sizes": sizes,
                "ptr_key": None, "ptrs": None}

    def _check_views(self, params, fl):
        base = fl["p"].data_ptr()
        for p, o in zip(params, fl["offs"]):
            if p.data_ptr() != base + 4 * o:
                raise RuntimeError("a parameter was re-allocated after FusedAdam was built (e.g. module.to()); "
                                   "rebuild the optimizer")

    # --- overlapping the gradient all-reduce with backward ------------------------------------------------------
    def set_overlap_groups(self, groups):
        """``groups``: lists of directly-accumulated FC weights (one list per message-passing layer).  When every weight
        of a group has reported its gradient final (``notify_grad_ready``, called by the backward operators right
        after they issue the weight-gradient GEMMs), the group's contiguous range of the flat gradient buffer is
        all-reduced on a communication stream, while backward goes on with the layers below.  ``step()`` joins that
        stream and all-reduces whatever was not covered.  No-op without a process group."""
        self._overlap, self._overlap_of = [], {}
        if self.process_group is None or self.nvls or os.environ.get("I3D_AR_OVERLAP", "1") == "0":
            return                           # (the fused NVLS step reduces inside the optimizer kernel)
        for params in groups:
            where = {}
            for fi, (g, fl) in enumerate(zip(self.param_groups, self._flat)):
                if fl is None:
                    continue
                for k, p in enumerate(g["params"]):
                    where[id(p)] = (fi, fl["offs"][k], fl["offs"][k] + (fl["sizes"][k] + 3) // 4 * 4, fl["direct"][k])
            locs = [where.get(id(p)) for p in params]
            if not locs or any(l is None or not l[3] or l[0] != locs[0][0] for l in locs):
                continue                                           # not all direct / not in one flat buffer
            locs.sort(key=lambda l: l[1])
            if any(a[2] != b[1] for a, b in zip(locs, locs[1:])):
                continue                                           # not contiguous
            gi = len(self._overlap)
            self._overlap.append((locs[0][0], locs[0][1], locs[-1][2], list(params)))
            for p in params:
                self._overlap_of[id(p)] = gi
        self._reset_pending()

    def _reset_pending(self):
        self._pending = {gi: {id(p) for p in grp[3]} for gi, grp in enumerate(self._overlap)}
        self._reduced = []

    def notify_grad_ready(self, param, streams=None):
        """The weight-gradient kernels of ``param`` have been issued on ``streams`` (default: the current stream)."""
        gi = self._overlap_of.get(id(param))
        if gi is None:
            return
        pend = self._pending.get(gi)
        if not pend:
            return
        pend.discard(id(param))
        if pend:
            return
        fi, s, e, _ = self._overlap[gi]
        fl = self._flat[fi]
        dev = fl["g"].device
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        comm = self._comm_stream
        for st in (streams or [torch.cuda.current_stream(dev)]):
            comm.wait_stream(st)
        with torch.cuda.stream(comm):
            torch.distributed.all_reduce(fl["g"][s:e], group=self.process_group)
        self._reduced.append((fi, s, e))

    def _allreduce_rest(self, fi, fl):
        """all-reduce the parts of flat gradient buffer ``fi`` that no overlap group has covered"""
        done = sorted((s, e) for f, s, e in self._reduced if f == fi)
        pos, n = 0, fl["g"].numel()
        for s, e in done + [(n, n)]:
            if s > pos:
                torch.distributed.all_reduce(fl["g"][pos:s], group=self.process_group)
            pos = max(pos, e)

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        if self._needs_zero:
            raise RuntimeError("FusedAdam: call optimizer.zero_grad() between steps — FC weight gradients accumulate "
                               "in place in the flat buffer (module.zero_grad() does not clear it)")
        self._step += 1
        if self.nvls:
            # pack every group's small gradients first, then ONE cross-rank barrier, the fused kernels, ONE barrier
            for gi, (g, fl) in enumerate(zip(self.param_groups, self._flat)):
                if fl is not None:
                    self._pack(g, fl)
            first = next(fl for fl in self._flat if fl is not None)
            first["handle"].barrier()
        for gi, (g, fl) in enumerate(zip(self.param_groups, self._flat)):
            if fl is None:
                continue
            if not self.nvls:
                self._pack(g, fl)
            self._finish_group(gi, g, fl, grad_scale)
        if self.nvls:
            first["handle"].barrier()
        if self.graph_safe:
            K.add_i64(self._step_dev, 1)
        for hook in self.post_step_hooks:
            hook()

    def _pack(self, g, fl):
        if True:
            params = g["params"]
            self._check_views(params, fl)
            # directly-accumulated weights already live in fl["g"]; parameters that took no part in the step keep
            # the zeros zero_grad() left there (Adam still decays their moments)
            gbase = fl["g"].data_ptr()
            grads = [None if (d or p.grad is None or p.grad.data_ptr() == gbase + 4 * o) else p.grad
                     for p, d, o in zip(params, fl["direct"], fl["offs"])]
            self._needs_zero = self._needs_zero or any(fl["direct"])
            def elem_stride(x):
                """element stride of a gradient the pack kernel can read: contiguous (1) or a 1-D strided view (the
                bias gradients of the FC layers live one per 32-byte sector, kernels.DBIAS_STRIDE)"""
                if x.is_contiguous():
                    return 1
                if x.dim() == 1 and x.stride(0) > 0:
                    return x.stride(0)
                raise RuntimeError("gradients must be contiguous (or 1-D strided) fp32")

            key = tuple(0 if x is None else x.data_ptr() for x in grads)
            if key != fl["ptr_key"]:
                for x in grads:
                    if x is not None and x.dtype != torch.float32:
                        raise RuntimeError("gradients must be fp32")
                live = [i for i, x in enumerate(grads) if x is not None]
                # one pinned table (pointers, offsets, lengths) + async copy: legal inside CUDA-graph capture (the
                # graph re-reads the pinned host memory on replay, so it is kept alive)
                host = torch.tensor([[key[i] for i in live], [fl["offs"][i] for i in live],
                                     [fl["sizes"][i] for i in live], [elem_stride(grads[i]) for i in live]],
                                    dtype=torch.int64).reshape(4, len(live))
                if live:
                    host = host.pin_memory()
                # a captured graph re-reads its pinned table on every replay, so tables created during a capture are
                # kept for good; in eager mode the copy below is stream-ordered and the previous table is only kept
                # until this one replaces it (no unbounded growth when gradient addresses change every step)
                if torch.cuda.is_current_stream_capturing():
                    fl.setdefault("ptrs_host_captured", []).append(host)
                else:
                    fl["ptrs_host_prev"] = fl.get("ptrs_host_eager")
                    fl["ptrs_host_eager"] = host
                table = torch.empty(4, len(live), dtype=torch.int64, device=fl["p"].device)
                table.copy_(host, non_blocking=True)
                fl["ptrs"], fl["poff"], fl["plen"], fl["pstride"] = table[0], table[1], table[2], table[3]
                fl["ptr_key"] = key
            if fl["ptrs"].numel():
                K.multi_copy(fl["ptrs"], fl["poff"], fl["plen"], fl["g"], True, fl["pstride"])

    def _finish_group(self, gi, g, fl, grad_scale):
        if True:
            b1, b2 = g["betas"]
            if self.nvls:
                # one kernel: switch-side gradient reduction + Adam on this rank's slice + multicast store of p, m, v.
                # The signal-pad barriers order the ranks: every gradient written before (taken once, after the LAST
                # group's pack, see below), every store landed after.
                rank = torch.distributed.get_rank(self.process_group)
                world = torch.distributed.get_world_size(self.process_group)
                if self.graph_safe:
                    hyper = (float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                             float(grad_scale))
                    if hyper != self._hyper_host[gi]:
                        self._upload_hyper(gi, hyper)
                    K.adam_step_nvls(fl["mc"], fl["sym"], fl["p"].numel(), rank, world, 0, b1, b2, g["eps"],
                                     g["weight_decay"], grad_scale, 1, self._hyper_dev[gi], self._step_dev)
                else:
                    K.adam_step_nvls(fl["mc"], fl["sym"], fl["p"].numel(), rank, world, g["lr"], b1, b2, g["eps"],
                                     g["weight_decay"], grad_scale, self._step)
                return
            if self.process_group is not None:
                if self._reduced and self._comm_stream is not None:
                    torch.cuda.current_stream(fl["g"].device).wait_stream(self._comm_stream)    # early layer all-reduces
                self._allreduce_rest(gi, fl)
            if self.graph_safe:
                hyper = (float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                         float(grad_scale))
                if hyper != self._hyper_host[gi]:
                    self._upload_hyper(gi, hyper)
                K.adam_step(fl["p"], fl["g"], fl["m"], fl["v"], 0, b1, b2, g["eps"], g["weight_decay"], grad_scale, 1,
                            self._hyper_dev[gi], self._step_dev)
            else:
                K.adam_step(fl["p"], fl["g"], fl["m"], fl["v"], g["lr"], b1, b2, g["eps"], g["weight_decay"],
                            grad_scale, self._step)

    def _upload_hyper(self, gi, hyper):
        host = torch.tensor(hyper, dtype=torch.float64).pin_memory()
        self._hyper_dev[gi].copy_(host, non_blocking=True)
        self._hyper_host[gi] = hyper

    def sync_hyper(self, grad_scale=1.0):
        """graph_safe mode: push changed lr / betas / ... to the device copy BEFORE replaying a captured step."""
        for gi, g in enumerate(self.param_groups):
            b1, b2 = g["betas"]
            hyper = (float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                     float(grad_scale))
            if hyper != self._hyper_host[gi]:
                self._upload_hyper(gi, hyper)

    def packed_grads(self):
        """{parameter: view of its gradient inside the flat buffer} as the last ``step()`` saw them (after the pack and,
        when data parallel, the all-reduce).  Valid until the next ``zero_grad()``; this is how gradients are read
        back after a captured step, where ``p.grad`` of the packed parameters lives in the graph's private pool."""
        out = {}
        for g, fl in zip(self.param_groups, self._flat):
            if fl is None:
                continue
            for p, o, n in zip(g["params"], fl["offs"], fl["sizes"]):
                out[p] = fl["g"][o:o + n].view(p.shape)
        return out

    def zero_grad(self, set_to_none=True):
        self._needs_zero = False
        if self._overlap:
            self._reset_pending()
        for g, fl in zip(self.param_groups, self._flat):
            if fl is None:
                continue
            fl["g"].zero_()                                   # one memset: packed copies and direct views alike
            for p, d in zip(g["params"], fl["direct"]):
                if d:
                    p.grad = p._i3d_grad_view                 # undo a module.zero_grad() that dropped the view
                elif p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    # --- torch.optim.Adam-compatible checkpoint format (trainer/self_supervised_trainer.py:88-97) -------------
    def state_dict(self):
        state, groups, idx = {}, [], 0
        for g, fl in zip(self.param_groups, self._flat):
            ids = []
            for k, p in enumerate(g["params"]):
                o, n = fl["offs"][k], fl["sizes"][k]
                state[idx] = {"step": torch.tensor(float(self._step)),
                              "exp_avg": fl["m"][o:o + n].view(p.shape).clone(),
                              "exp_avg_sq": fl["v"][o:o + n].view(p.shape).clone()}
                ids.append(idx)
                idx += 1
            pg = {k: v for k, v in g.items() if k != "params"}
            pg["params"] = ids
            groups.append(pg)
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd):
        idx = 0
        for g, fl, sg in zip(self.param_groups, self._flat, sd["param_groups"]):
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in sg:
                    g[k] = sg[k]
            for k, p in enumerate(g["params"]):
                st = sd["state"].get(idx)
                if st is not None:
                    o, n = fl["offs"][k], fl["sizes"][k]
                    fl["m"][o:o + n].copy_(st["exp_avg"].reshape(-1))
                    fl["v"][o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    self._step = int(st["step"])
                idx += 1
        if self.graph_safe:
            self._step_dev.fill_(self._step + 1)
