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
from itertools import chain

import torch

from . import dist as D
from .graph import GraphBatch
from .kernels import StatsArena, WeightPrep
from .optim import FusedAdam


class SelfSupervisedTrainer:
    def __init__(self, model, model3d, loss_func, device="cuda", optimizer_params=None, lr_scheduler=None,
                 process_group=None, graph_safe=False, overlap_encoders=True):
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.model3d = model3d.to(self.device)      # moved before the optimizer is built (self_supervised_trainer.py:16)
        self.loss_func = loss_func
        self.process_group = process_group
        self.world = torch.distributed.get_world_size(process_group) if D.is_distributed() else 1
        self.rank = torch.distributed.get_rank(process_group) if D.is_distributed() else 0
        self.optim_steps = 0
        self.lr_scheduler = lr_scheduler
        # the 2-D and 3-D encoders share nothing until the loss: the 3-D one (small, latency-bound kernels) runs on
        # its own stream, forward and backward (autograd replays a node on the stream its forward ran on), and fills
        # the SMs the 2-D encoder's one-CTA-per-SM GEMMs leave idle.  Inside a captured step this is a forked branch.
        self.stream3d = torch.cuda.Stream(device=self.device) if overlap_encoders else None
        self.initialize_optimizer(optimizer_params or {"lr": 8e-5}, graph_safe)

    def initialize_optimizer(self, optimizer_params, graph_safe=False):
        named = list(chain(self.model.named_parameters(), self.model3d.named_parameters()))
        normal_params = [v for k, v in named if "batch_norm" not in k]
        batch_norm_params = [v for k, v in named if "batch_norm" in k]
        self.optim = FusedAdam([{"params": batch_norm_params, "weight_decay": 0}, {"params": normal_params}],
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
        self.optim.post_step_hooks.append(self.prep.invalidate)

    def forward_pass(self, batch):
        self.arena.reset()
        self.prep.refresh()
        info2d, info3d, *rest = tuple(batch)
        if self.stream3d is not None:
            main = torch.cuda.current_stream(self.device)
            self.stream3d.wait_stream(main)
            with torch.cuda.stream(self.stream3d):
                view3d = self.model3d(*info3d)
            view2d = self.model(*info2d)
            main.wait_stream(self.stream3d)
            view3d.record_stream(main)
        else:
            view2d = self.model(*info2d)
            view3d = self.model3d(*info3d)
        if self.world > 1:
            # global negative set: every rank's 3-D embeddings; the local rows sit at row_offset in column space
            b_local = view2d.shape[0]
            view3d_all = D.all_gather_rows(view3d, self.process_group)
            loss = self.loss_func(view2d, view3d_all, row_offset=self.rank * b_local,
                                  total_rows=self.world * b_local)
        else:
            loss = self.loss_func(view2d, view3d, nodes_per_graph=None)
        return loss, view2d, view3d

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


class CapturedStep:
    """One training step as a CUDA graph over static input buffers.

    ``load(batch_np)`` copies a collated batch (numpy dict, 3dinfomax_b200/synthetic.py layout; pinned host staging) into the
    static device buffers — shapes must match the capture; ``run()`` replays; ``loss`` is a device scalar.
    """

    def __init__(self, trainer, example_g2, example_g3, warmup=3):
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
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_body()
                self.tr.optim_steps += 1
        torch.cuda.current_stream(dev).wait_stream(side)
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
