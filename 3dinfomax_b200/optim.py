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
        for g in groups:
            pg = dict(self.defaults)
            pg.update(g)
            pg["params"] = [p for p in pg["params"]]
            self.param_groups.append(pg)
            self._flat.append(self._flatten(pg["params"]))
        self._step = 0
        self._needs_zero = False
        self.post_step_hooks = []          # e.g. WeightPrep.invalidate: the Adam kernel bypasses tensor versions
        dev = self._device()
        self._step_dev = torch.ones(1, dtype=torch.int64, device=dev) if graph_safe else None
        self._hyper_dev = [torch.zeros(6, dtype=torch.float64, device=dev) for _ in self.param_groups] \
            if graph_safe else None
        self._hyper_host = [None] * len(self.param_groups)

    def _device(self):
        for g in self.param_groups:
            for p in g["params"]:
                return p.device
        return torch.device("cuda")

    @staticmethod
    def _flatten(params):
        if not params:
            return None
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs CUDA parameters: the 3dinfomax_b200 path has no CPU fallback")
        sizes = [p.numel() for p in params]
        offs, acc = [], 0
        for n in sizes:
            offs.append(acc)
            acc += (n + 3) // 4 * 4          # keep every view 16-byte aligned for the vectorised kernels
        flat = torch.zeros(acc, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o, n in zip(params, offs, sizes):
                if p.dtype != torch.float32:
                    raise TypeError("fp32 parameters only")
                flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = flat[o:o + n].view(p.shape)
        gflat = torch.zeros_like(flat)
        direct = []
        for p, o, n in zip(params, offs, sizes):
            d = bool(getattr(p, "_i3d_direct_grad", False))
            direct.append(d)
            if d:
                p._i3d_grad_view = gflat[o:o + n].view(p.shape)
                p.grad = p._i3d_grad_view
        return {"p": flat, "g": gflat, "m": torch.zeros_like(flat), "v": torch.zeros_like(flat), "direct": direct,
                "off": torch.tensor(offs, dtype=torch.int64, device=dev),
                "len": torch.tensor(sizes, dtype=torch.int64, device=dev), "offs": offs, "sizes": sizes,
                "ptr_key": None, "ptrs": None}

    def _check_views(self, params, fl):
        base = fl["p"].data_ptr()
        for p, o in zip(params, fl["offs"]):
            if p.data_ptr() != base + 4 * o:
                raise RuntimeError("a parameter was re-allocated after FusedAdam was built (e.g. module.to()); "
                                   "rebuild the optimizer")

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        if self._needs_zero:
            raise RuntimeError("FusedAdam: call optimizer.zero_grad() between steps — FC weight gradients accumulate "
                               "in place in the flat buffer (module.zero_grad() does not clear it)")
        self._step += 1
        for gi, (g, fl) in enumerate(zip(self.param_groups, self._flat)):
            if fl is None:
                continue
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
            if self.process_group is not None:
                torch.distributed.all_reduce(fl["g"], group=self.process_group)
            b1, b2 = g["betas"]
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
        if self.graph_safe:
            K.add_i64(self._step_dev, 1)
        for hook in self.post_step_hooks:
            hook()

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
