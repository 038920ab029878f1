"""Host-side mirror of the reference's contrastive pre-training step.

``SelfSupervisedTrainer.forward_pass`` / ``process_batch`` follow trainer/self_supervised_trainer.py:24-29 and
trainer/trainer.py:116-124: 2-D forward, 3-D forward, contrastive loss, ``backward()``, ``optim.step()``,
scheduler hook, ``zero_grad()``.  Optimizer grouping follows self_supervised_trainer.py:78-86 (parameters whose
name contains 'batch_norm' form a weight_decay=0 group).  Epoch loops, TensorBoard, checkpoints and metrics are the
reference's own (out of scope): the modules here are plain ``nn.Module``s, so the reference's
``Trainer`` can drive them unchanged.

``CapturedStep`` replays the whole step (CSR build -> both encoders -> loss -> backward -> gradient pack ->
[all-reduce] -> Adam) as ONE CUDA graph fed from static device buffers: at these sizes the step is launch-bound
(~350 kernels of a few microseconds each), and a graph is the B200-idiomatic way to remove the host from the loop.
"""
import math
from itertools import chain

import numpy as np
import torch

from . import dist as D
from .graph import GraphBatch
from .kernels import StatsArena, WeightPrep
from .optim import FusedAdam


class Trainer:
    """The reference's supervised trainer (trainer/trainer.py:26-124) for the fine-tuning configs
    (configs_clean/tune_QM9_homo.yml: PNA with target_dim 1, torch's own L1Loss, Adam): ``forward_pass`` is
    trainer/trainer.py:111-114 — the last entry of the batch tuple is the targets, the first the model's arguments —
    and ``process_batch`` :116-124.  ``SelfSupervisedTrainer`` below adds the second encoder."""

    def __init__(self, model, loss_func, device="cuda", optimizer_params=None, lr_scheduler=None, process_group=None,
                 graph_safe=False, transfer_layers=(), exclude_from_transfer=(), frozen_layers=(), transferred_lr=None):
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.model3d = None
        self._init_common(loss_func, lr_scheduler, process_group)
        self.stream3d = None
        self.transfer_layers, self.exclude_from_transfer = tuple(transfer_layers), tuple(exclude_from_transfer)
        self.frozen_layers, self.transferred_lr = tuple(frozen_layers), transferred_lr
        self.initialize_optimizer(optimizer_params or {"lr": 1e-3}, graph_safe)

    def _init_common(self, loss_func, lr_scheduler, process_group):
        self.loss_func = loss_func
        self.process_group = process_group
        self.world = torch.distributed.get_world_size(process_group) if D.is_distributed() else 1
        self.rank = torch.distributed.get_rank(process_group) if D.is_distributed() else 0
        self.optim_steps = 0
        self.lr_scheduler = lr_scheduler

    def encoders(self):
        return [self.model] if self.model3d is None else [self.model, self.model3d]

    def param_groups(self, named, optimizer_params):
        """trainer/trainer.py:216-238: BatchNorm parameters without weight decay, new parameters, parameters transferred
        from a pre-trained checkpoint (own learning rate), frozen parameters (lr 0) — in this order, which is the order
        the reference's ordered warm-up walks through"""
        keys = list(self.model.state_dict().keys())
        transferred = {k for k in keys if any(t in k for t in self.transfer_layers)
                       and not any(x in k for x in self.exclude_from_transfer)}
        frozen = {k for k in keys if any(f in k for f in self.frozen_layers)}
        frozen_params = [v for k, v in named if k in frozen]
        transferred_params = [v for k, v in named if k in transferred]
        new_params = [v for k, v in named if k not in transferred and "batch_norm" not in k and k not in frozen]
        batch_norm_params = [v for k, v in named if "batch_norm" in k and k not in transferred and k not in frozen]
        transfer_lr = optimizer_params.get("lr", 1e-3) if self.transferred_lr is None else self.transferred_lr
        groups = []
        if batch_norm_params:
            groups.append({"params": batch_norm_params, "weight_decay": 0})
        groups.append({"params": new_params})
        if transferred_params:
            groups.append({"params": transferred_params, "lr": transfer_lr})
        if frozen_params:
            groups.append({"params": frozen_params, "lr": 0})
        return groups

    def initialize_optimizer(self, optimizer_params, graph_safe=False):
        named = list(chain(*[m.named_parameters() for m in self.encoders()]))
        self.optim = FusedAdam(self.param_groups(named, optimizer_params),
                               process_group=self.process_group if self.world > 1 else None, graph_safe=graph_safe,
                               **optimizer_params)
        # FC weights get their tf32 hi/lo operand copies (plain and transposed) from one launch per step
        self.prep = WeightPrep(self.device)
        # BatchNorm statistics scratch of the whole step, zeroed by one memset at the start of forward_pass
        self.arena = StatsArena(self.device)
        for _, v in named:
            if getattr(v, "_i3d_direct_grad", False):
                v._i3d_prep = self.prep
                v._i3d_arena = self.arena
                v._i3d_optim = self.optim
        self.optim.post_step_hooks.append(self.prep.invalidate)
        # data parallel: the FC weights of one message-passing layer form one contiguous range of the flat gradient
        # buffer, all-reduced as soon as that layer's backward has issued its weight-gradient kernels
        layers = {}
        for k, v in self.model.named_parameters():
            if getattr(v, "_i3d_direct_grad", False) and ".mp_layers." in k:
                layers.setdefault(k.split(".mp_layers.")[1].split(".")[0], []).append(v)
        self.optim.set_overlap_groups(list(layers.values()))

    # --- state that a step mutates (used to make CUDA-graph warm-up runs side-effect free) ---------------------
    def snapshot_state(self):
        """Copies of everything ``process_batch`` mutates: parameters, Adam moments and step counters, BatchNorm
        buffers.  ``restore_state`` puts them back (same storage: views handed to modules / captured graphs stay valid)."""
        o = self.optim
        bufs = [b for m in self.encoders() for b in m.buffers()]
        return {"flat": [None if fl is None else (fl["p"].clone(), fl["m"].clone(), fl["v"].clone()) for fl in o._flat],
                "step": o._step, "step_dev": None if o._step_dev is None else o._step_dev.clone(),
                "bufs": [b.clone() for b in bufs], "optim_steps": self.optim_steps}

    @torch.no_grad()
    def restore_state(self, snap):
        o = self.optim
        for fl, c in zip(o._flat, snap["flat"]):
            if fl is not None:
                fl["p"].copy_(c[0]), fl["m"].copy_(c[1]), fl["v"].copy_(c[2])
        o._step = snap["step"]
        if o._step_dev is not None:
            o._step_dev.copy_(snap["step_dev"])
        bufs = [b for m in self.encoders() for b in m.buffers()]
        for b, c in zip(bufs, snap["bufs"]):
            b.copy_(c)
        self.optim_steps = snap["optim_steps"]
        self.prep.invalidate()

    def forward_pass(self, batch):
        self.arena.reset()
        self.prep.refresh()
        targets = batch[-1]                       # the last entry of the batch tuple is always the targets
        predictions = self.model(*batch[0])
        return self.loss_func(predictions, targets), predictions, targets

    def process_batch(self, batch, optim=True):
        loss, predictions, targets = self.forward_pass(batch)
        if optim:
            loss.backward()
            self.optim.step()
            self.after_optim_step()
            self.optim.zero_grad()
            self.optim_steps += 1
        return loss, predictions.detach(), targets.detach()

    def after_optim_step(self):
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()


class SelfSupervisedTrainer(Trainer):
    def __init__(self, model, model3d, loss_func, device="cuda", optimizer_params=None, lr_scheduler=None,
                 process_group=None, graph_safe=False, overlap_encoders=True):
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.model3d = model3d.to(self.device)      # moved before the optimizer is built (self_supervised_trainer.py:16)
        self._init_common(loss_func, lr_scheduler, process_group)
        # the 2-D and 3-D encoders share nothing until the loss: the 3-D one (small, latency-bound kernels) runs on
        # its own stream, forward and backward (autograd replays a node on the stream its forward ran on), and fills
        # the SMs the 2-D encoder's one-CTA-per-SM GEMMs leave idle.  Inside a captured step this is a forked branch.
        self.stream3d = torch.cuda.Stream(device=self.device) if overlap_encoders else None
        if self.stream3d is not None:
            from . import kernels as _K
            _K.NO_FUSED_BN_BWD_STREAMS.add(self.stream3d.cuda_stream)     # one spinning grid barrier per device at a time
        self.initialize_optimizer(optimizer_params or {"lr": 8e-5}, graph_safe)

    def param_groups(self, named, optimizer_params):
        """trainer/self_supervised_trainer.py:78-86 over both encoders"""
        normal_params = [v for k, v in named if "batch_norm" not in k]
        batch_norm_params = [v for k, v in named if "batch_norm" in k]
        return [{"params": batch_norm_params, "weight_decay": 0}, {"params": normal_params}]

    def forward_pass(self, batch):
        self.arena.reset()
        self.prep.refresh()
        info2d, info3d, *rest = tuple(batch)
        view3d_all = None
        if self.stream3d is not None:
            main = torch.cuda.current_stream(self.device)
            self.stream3d.wait_stream(main)
            with torch.cuda.stream(self.stream3d):
                view3d = self.model3d(*info3d)
                if self.world > 1:
                    # the 3-D encoder is done long before the 2-D one: its all-gather (and, in backward, the
                    # reduce-scatter autograd replays on this stream) hides behind the 2-D encoder
                    view3d_all = D.all_gather_rows(view3d, self.process_group)
            view2d = self.model(*info2d, *rest)      # *snorm_n of PNAOriginal (self_supervised_trainer.py:25-26)
            main.wait_stream(self.stream3d)
            view3d.record_stream(main)
            if view3d_all is not None:
                view3d_all.record_stream(main)
        else:
            view2d = self.model(*info2d, *rest)
            view3d = self.model3d(*info3d)
        if self.world > 1:
            # global negative set: every rank's 3-D embeddings; the local rows sit at row_offset in column space
            b_local = view2d.shape[0]
            if view3d_all is None:
                view3d_all = D.all_gather_rows(view3d, self.process_group)
            loss = self.loss_func(view2d, view3d_all, row_offset=self.rank * b_local,
                                  total_rows=self.world * b_local)
        else:
            loss = self.loss_func(view2d, view3d, nodes_per_graph=None)
        return loss, view2d, view3d


class CapturedStep:
    """One training step as a CUDA graph over static input buffers.

    ``load(batch_np)`` copies a collated batch (numpy dict, 3dinfomax_b200/synthetic.py layout; pinned host staging) into the
    static device buffers — shapes must match the capture; ``run()`` replays; ``loss`` is a device scalar.
    """

    def __init__(self, trainer, example_g2, example_g3, warmup=3, keep_warmup_updates=False):
        if not trainer.optim.graph_safe:
            raise ValueError("CapturedStep needs SelfSupervisedTrainer(..., graph_safe=True): with host-side "
                             "hyper-parameters the learning rate and Adam's step count would be frozen into the graph")
        if getattr(trainer.model, "needs_snorm", False):
            raise NotImplementedError("CapturedStep feeds forward(graph) models; the tower models' forward(graph, snorm_n) "
                                      "is driven by BucketedStep, which builds snorm_n on the device")
        self.tr = trainer
        dev = trainer.device
        self.static = {
            "src": example_g2.edges()[0].clone(), "dst": example_g2.edges()[1].clone(),
            "bnn": example_g2.batch_num_nodes().clone(), "x": example_g2.ndata["feat"].clone(),
            "e": example_g2.edata["feat"].clone(),
            "src3": example_g3.edges()[0].clone(), "dst3": example_g3.edges()[1].clone(),
            "bnn3": example_g3.batch_num_nodes().clone(), "d3": example_g3.edata["d"].clone(),
        }
        self.n2, self.n3 = example_g2.number_of_nodes(), example_g3.number_of_nodes()
        self.max_in_degree = getattr(example_g2, "max_in_degree", None)   # sizes the degree plan inside the graph
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.graph = None
        # the warm-up runs are real steps on the example batch (lazy initialisation, allocator warm-up): unless asked
        # otherwise everything they changed (weights, Adam state, BatchNorm buffers, step counters) is put back
        snap = None if keep_warmup_updates else trainer.snapshot_state()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_body()
                self.tr.optim_steps += 1
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if snap is not None:
            trainer.restore_state(snap)
        trainer.prep.refresh()          # operand table of the entries the warm-up registered: built eagerly, not captured
        torch.cuda.synchronize(dev)
        from . import lib as _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_body()
        self.tr.optim._step -= 1                    # capture ran the Python side of step() without executing it
        self.launches_per_step = _lib.launch_count() - n0

    def _graphs(self):
        s = self.static
        g2 = GraphBatch(s["src"], s["dst"], s["bnn"], None, {"feat": s["x"]}, {"feat": s["e"]}, self.n2,
                        self.max_in_degree)
        g3 = GraphBatch(s["src3"], s["dst3"], s["bnn3"], None, {}, {"d": s["d3"]}, self.n3)
        return g2, g3

    def _step_body(self):
        self.tr.optim.zero_grad(set_to_none=True)   # inside the graph: one memset per parameter group
        g2, g3 = self._graphs()
        loss, _, _ = self.tr.forward_pass(([g2], [g3]))
        loss.backward()
        self.tr.optim.step()
        self.loss.copy_(loss.detach())

    def load(self, g2, g3):
        """Copy a new batch of IDENTICAL shape into the static buffers (async on the current stream)."""
        s = self.static
        if self.max_in_degree is not None:
            md = getattr(g2, "max_in_degree", None)
            if md is None or md > self.max_in_degree:
                raise ValueError("captured step was built for max_in_degree <= %d, the new batch has %r"
                                 % (self.max_in_degree, md))
        pairs = [("src", g2.edges()[0]), ("dst", g2.edges()[1]), ("bnn", g2.batch_num_nodes()),
                 ("x", g2.ndata["feat"]), ("e", g2.edata["feat"]), ("src3", g3.edges()[0]), ("dst3", g3.edges()[1]),
                 ("bnn3", g3.batch_num_nodes()), ("d3", g3.edata["d"])]
        for k, t in pairs:
            if t.shape != s[k].shape:
                raise ValueError("captured step needs identical shapes (%s: %s vs %s)" % (k, tuple(t.shape),
                                                                                        tuple(s[k].shape)))
            s[k].copy_(t, non_blocking=True)

    def run(self):
        self.tr.optim.sync_hyper()
        self.graph.replay()
        self.tr.optim_steps += 1
        self.tr.optim._step += 1
        if self.tr.lr_scheduler is not None:
            self.tr.lr_scheduler.step()
        return self.loss


class BucketLadder:
    """Shape buckets for batches of ``batch_size`` molecules drawn from ``store``.

    N, E and E3 of a batch are sums over its molecules, so they concentrate around ``B * mean`` with a standard
    deviation of ``sqrt(B) * std`` (QM9, B = 512: 0.7 % of N, 1.5 % of E3).  Level k holds every batch whose three sizes
    stay below ``B * mean + (k + 1) * step * sqrt(B) * std`` (rounded up to ``align`` rows): ~84 % of the batches of an
    epoch land in level 0, ~14 % in level 1, and the padding costs 1-3 % of the rows.  Every rank of a data-parallel job
    derives the same ladder from the same store."""

    def __init__(self, store, batch_size, conformers=1, step=1.0, align=128):
        n = store.n_atoms.astype(np.float64)
        e = store.n_edges.astype(np.float64)
        e3 = float(conformers) * n * (n - 1.0)
        B = int(batch_size)
        self.B, self.C, self.step, self.align = B, int(conformers), float(step), int(align)
        self.mean = [B * float(x.mean()) for x in (n, e, e3)]
        self.sd = [max(math.sqrt(B) * float(x.std()), 1.0) for x in (n, e, e3)]
        self.lo = [int(np.sort(x)[:B].sum()) for x in (n, e, e3)] if B <= len(n) else None

    def caps(self, level):
        """(n_cap, e_cap, e3_cap) of ``level``; n_cap counts 2-D nodes (the 3-D graph has conformers * n_cap)"""
        up = lambda x: int(-(-int(math.ceil(x)) // self.align) * self.align)
        return tuple(max(up(m + (level + 1) * self.step * sd), self.align) for m, sd in zip(self.mean, self.sd))

    def level_of(self, sizes):
        lv = 0
        for x, m, sd in zip(sizes, self.mean, self.sd):
            lv = max(lv, int(math.ceil((x - m) / (self.step * sd))) - 1)
        while any(x > c for x, c in zip(sizes, self.caps(lv))):
            lv += 1
        return lv


class _Bucket:
    __slots__ = ("graph", "meta", "caps", "launches", "replays")


class BucketedStep:
    """The batch loop of ``train.py`` (train.py:595-598 -> trainer/trainer.py:116-124) on captured CUDA graphs although
    every batch of an epoch has a different (N, E, E3): batches are built ON THE DEVICE from a ``PackedMoleculeStore``
    padded to a small ladder of shape buckets, one captured graph per (batch size, bucket, train/eval).

        run = BucketedStep(trainer, store, conformers=1)
        for idx in sampler:                  # host int array of molecule ids, any batch size (the last one is short)
            loss = run.step(idx)             # device scalar; run.predictions / run.targets = z2d / z3d of the step

    Per step the host computes the batch's three sizes (O(B) integer adds), picks the bucket, uploads 7B+3 int64 of
    metadata from an event-guarded pinned ring and replays the graph; the graph holds collate (both graphs + CSR
    structures, no sort), both encoders, the loss, backward, [NCCL], Adam.  Padding rows are masked where they would
    matter (BatchNorm statistics and counts; see include/i3d.h "Padding convention") and are exact zeros in every
    gradient, so a padded step equals the unpadded one up to fp32 summation order (tests: case_bucketed_step).

    Data parallel: every rank replays its own bucket's graph — all graphs issue the same collectives in the same order
    with the same sizes.  Capturing needs an eager warm-up run (with collectives), so under DP the levels
    ``0..dp_levels-1`` of a batch size are captured together, on all ranks, the first time that batch size is seen; a
    batch beyond the ladder runs eagerly (same collectives)."""

    def __init__(self, trainer, store, conformers=1, sigma_step=1.0, dp_levels=6, keep_graph_outputs=True):
        if not trainer.optim.graph_safe:
            raise ValueError("BucketedStep needs SelfSupervisedTrainer(..., graph_safe=True)")
        self.supervised = trainer.model3d is None       # Trainer: batch = ([2-D graph], targets[idx])
        if self.supervised and store.targets is None:
            raise ValueError("supervised steps need `targets` [M, T] in the packed store")
        self.tr, self.store, self.C = trainer, store, int(conformers)
        self.sigma_step, self.dp_levels = float(sigma_step), int(dp_levels)
        self.pool = torch.cuda.graph_pool_handle()
        self.buckets, self.ladders = {}, {}
        dev = trainer.device
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.keep_outputs = keep_graph_outputs
        self.predictions = self.targets = None
        self._out = {}
        self.stats = {"steps": 0, "captures": 0, "eager": 0, "levels": {}}
        for m in trainer.encoders():
            gnn = getattr(m, "node_gnn", None)
            if gnn is not None and hasattr(gnn, "keep_edge_side_effects"):
                gnn.keep_edge_side_effects = False        # nobody sees the graph object of a captured step

    def ladder(self, B):
        lad = self.ladders.get(B)
        if lad is None:
            lad = self.ladders[B] = BucketLadder(self.store, B, self.C, self.sigma_step)
        return lad

    # ---- one step's kernel sequence (captured) ---------------------------------------------------------------
    def _body(self, bk, B, train):
        tr = self.tr
        if train:
            tr.optim.zero_grad(set_to_none=True)
        g2, g3 = self.store.collate_padded(bk.meta, B, *bk.caps, conformers=self.C, need_3d=not self.supervised)
        info2d = [g2] + ([self._snorm(g2)] if getattr(tr.model, "needs_snorm", False) else [])
        if self.supervised:
            from .collate import metadata_views
            batch = (info2d, self.store.targets.index_select(0, metadata_views(bk.meta, B)["idx"]))
        else:
            batch = (info2d, [g3])
        if train:
            loss, z2, z3 = tr.forward_pass(batch)
            loss.backward()
            tr.optim.step()
        else:
            with torch.no_grad():
                loss, z2, z3 = tr.forward_pass(batch)
        self.loss.copy_(loss.detach())
        if self.keep_outputs:
            o2, o3 = self._outputs(B, z2, z3)
            o2.copy_(z2.detach()), o3.copy_(z3.detach())

    @staticmethod
    def _snorm(g2):
        """snorm_n of the tower models (graph_collate, datasets/custom_collate.py:96-98): sqrt(1 / atoms of the node's
        molecule) per node, [N, 1] — fixed-shape device ops (capturable); padding nodes get the last molecule's value"""
        nn_ = g2.batch_num_nodes()
        B, n = nn_.numel(), g2.number_of_nodes()
        ends = torch.cumsum(nn_, 0)
        node = torch.arange(n, device=nn_.device, dtype=ends.dtype)
        gid = torch.searchsorted(ends, node, right=True).clamp_(max=B - 1)
        return nn_.to(torch.float32).rsqrt()[gid].unsqueeze(1)

    def _outputs(self, B, z2, z3):
        o = self._out.get(B)
        if o is None:
            o = self._out[B] = (torch.zeros_like(z2), torch.zeros_like(z3))
        return o

    def _capture(self, B, level, train, example_idx):
        from . import lib as _lib
        from .collate import metadata_len
        tr, dev = self.tr, self.tr.device
        bk = _Bucket()
        bk.caps = self.ladder(B).caps(level)
        bk.meta = torch.zeros(metadata_len(B), dtype=torch.int64, device=dev)
        bk.replays = 0
        self.store.stage_metadata(example_idx, dev_out=bk.meta)
        snap = tr.snapshot_state()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._body(bk, B, train)                    # eager warm-up: lazy initialisation, persistent scratch
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        tr.restore_state(snap)                          # ... without training on the example batch
        tr.prep.refresh()                               # operand table built eagerly (never inside a capture)
        torch.cuda.synchronize(dev)
        n0 = _lib.launch_count()
        bk.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(bk.graph, pool=self.pool):
            self._body(bk, B, train)
        if train:
            tr.optim._step -= 1                         # capture ran the Python side of step() without executing it
        bk.launches = _lib.launch_count() - n0
        self.buckets[(B, level, train)] = bk
        self.stats["captures"] += 1
        return bk

    def _bucket(self, B, level, train, idx):
        bk = self.buckets.get((B, level, train))
        if bk is not None:
            return bk
        if self.tr.world > 1:
            if level >= self.dp_levels:
                return None                             # beyond the pre-captured ladder: eager step
            # all ranks see a new batch size at the same step: capture the whole ladder in lockstep
            for lv in range(self.dp_levels):
                if (B, lv, train) not in self.buckets:
                    ex = idx if self.ladder(B).level_of(self.store.batch_sizes(idx, self.C)) <= lv else \
                        self._small_example(B)
                    self._capture(B, lv, train, ex)
            return self.buckets[(B, level, train)]
        return self._capture(B, level, train, idx)

    def _small_example(self, B):
        """B molecule ids whose batch fits level 0 (the smallest molecules of the store)"""
        return np.argsort(self.store.n_atoms, kind="stable")[:B]

    # ---- public ------------------------------------------------------------------------------------------------
    def step(self, idx, train=True):
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int64).reshape(-1)
        B = int(len(idx))
        tr = self.tr
        sizes = self.store.batch_sizes(idx, self.C)
        level = self.ladder(B).level_of(sizes)
        bk = self._bucket(B, level, bool(train), idx)
        self.stats["steps"] += 1
        if bk is None:
            return self._eager(idx, train)
        self.stats["levels"][level] = self.stats["levels"].get(level, 0) + 1
        self.store.stage_metadata(idx, dev_out=bk.meta)
        if train:
            tr.optim.sync_hyper()
        bk.graph.replay()
        bk.replays += 1
        if train:
            tr.optim_steps += 1
            tr.optim._step += 1
            tr.optim._needs_zero = False
            if tr.lr_scheduler is not None:
                tr.lr_scheduler.step()
        if self.keep_outputs:
            self.predictions, self.targets = self._out[B]
        return self.loss

    def evaluate(self, idx):
        """forward + loss only, modules in whatever mode the caller put them (trainer/trainer.py:72-75: the reference
        switches ``model`` to eval and leaves ``model3d`` in train mode)"""
        return self.step(idx, train=False)

    def _eager(self, idx, train):
        self.stats["eager"] += 1
        if self.C != 1:
            raise RuntimeError("batch beyond the captured ladder: the eager fallback handles one conformer per molecule")
        g2, g3 = self.store.collate(idx)
        info2d = [g2] + ([self._snorm(g2)] if getattr(self.tr.model, "needs_snorm", False) else [])
        if self.supervised:
            ix = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(self.tr.device)
            batch = (info2d, self.store.targets.index_select(0, ix))
        else:
            batch = (info2d, [g3])
        if train:
            loss, z2, z3 = self.tr.process_batch(batch)
        else:
            with torch.no_grad():
                loss, z2, z3 = self.tr.forward_pass(batch)
        self.loss.copy_(loss.detach())
        self.predictions, self.targets = z2.detach(), z3.detach()
        return self.loss

    @property
    def launches_per_step(self):
        tot = sum(b.launches * max(b.replays, 1) for b in self.buckets.values())
        cnt = sum(max(b.replays, 1) for b in self.buckets.values())
        return tot / max(cnt, 1)
