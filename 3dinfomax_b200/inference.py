"""Fingerprint / inference path: ``inference.py:208-214`` of the reference —

    test_loader = DataLoader(test_data, batch_size=2, collate_fn=graph_only_collate)
    for batch in test_loader: fingerprints_list.append(model(batch))

— over a device-resident ``PackedMoleculeStore``, as forward-only CUDA graphs per shape bucket.

Two modes:

* ``mode="reference"``: exactly what the reference script computes.  It never calls ``model.eval()`` nor ``no_grad``, so
  every batch of 2 molecules is normalised with ITS OWN BatchNorm statistics and the running statistics drift; kept for
  parity with fingerprints produced by the reference (the modules stay in train mode).
* ``mode="eval"`` (default): eval-mode forward with every BatchNorm folded into a neighbouring Linear
  (``fold_batch_norm``), so the forward has no normalisation pass at all.  FCLayer is Linear -> activation -> BatchNorm
  (models/base_layers.py:100-111): in eval mode BN is the per-column affine ``a x + c`` with
  ``a = gamma / sqrt(running_var + eps)``, ``c = beta - running_mean a``.  With activation 'none' it folds into the
  layer's own Linear (``W' = diag(a) W``, ``b' = a b + c``); behind an activation it folds into the NEXT Linear of the
  same MLP (``W_next' = W_next diag(a)``, ``b_next' = b_next + W_next c``).  For the shipped PNA that removes all 22
  BatchNorms (pretrans 0 -> pretrans 1, pretrans 1 and posttrans into themselves, readout 0 -> readout 1).

There is no CPU fallback: the store and the modules must live on a CUDA device (the folding algebra itself is plain
tensor arithmetic on the parameters and runs anywhere — tests/test_host_logic.py checks it on the CPU).
"""
import copy

import numpy as np
import torch

from .base_layers import MLP


def bn_affine(bn):
    """(a, c) with eval-mode BatchNorm(x) = a * x + c (float64 for the fold, cast back by the caller)"""
    a = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    c = bn.bias.detach().double() - bn.running_mean.detach().double() * a
    return a, c


@torch.no_grad()
def fold_mlp(mlp):
    """Fold the eval-mode BatchNorms of one ``MLP`` in place; returns the number of BatchNorms removed.  A BatchNorm that
    sits behind an activation in the LAST layer of the MLP has no Linear to fold into and is kept."""
    fcs = list(mlp.fully_connected)
    removed = 0
    for i, fc in enumerate(fcs):
        bn = fc.batch_norm
        if bn is None:
            continue
        a, c = bn_affine(bn)
        W, b = fc.linear.weight, fc.linear.bias
        if fc.act == 0:                                   # Linear -> BN: into the layer's own weights
            W.copy_((a[:, None] * W.double()).to(W.dtype))
            b.copy_((a * b.double() + c).to(b.dtype))
        elif i + 1 < len(fcs):                            # Linear -> act -> BN -> Linear_next: into the next weights
            Wn, bn_next = fcs[i + 1].linear.weight, fcs[i + 1].linear.bias
            if Wn.shape[1] != a.numel():
                continue                                  # the next layer reads something else as well: keep the BN
            bn_next.copy_((bn_next.double() + Wn.double() @ c).to(bn_next.dtype))
            Wn.copy_((Wn.double() * a[None, :]).to(Wn.dtype))
        else:
            continue
        fc.batch_norm = None
        removed += 1
    return removed


def fold_batch_norm(model):
    """Deep copy of ``model`` (PNA / Net3D / any module built from base_layers.MLP) in eval mode with every foldable
    BatchNorm folded away.  The copy shares nothing with the original; its state dict no longer has the folded
    ``batch_norm.*`` entries, so it is an inference artefact, not a checkpoint."""
    # (parameters driven by a trainer carry references to its optimizer / scratch arenas: not part of the model)
    stripped = []
    for p in model.parameters():
        for k in [k for k in p.__dict__ if k.startswith("_i3d_") and k != "_i3d_direct_grad"]:
            stripped.append((p, k, p.__dict__.pop(k)))
    try:
        m = copy.deepcopy(model).eval()
    finally:
        for p, k, v in stripped:
            p.__dict__[k] = v
    for p in m.parameters():
        p.grad = None
    m.folded_batch_norms = sum(fold_mlp(x) for x in m.modules() if isinstance(x, MLP))
    for p in m.parameters():
        p.requires_grad_(False)
    return m


class Fingerprinter:
    """``fingerprints = Fingerprinter(model, store)(indices)`` — the loop of inference.py:208-214 on shape-bucketed
    forward-only CUDA graphs fed by the device collate.  Returns a [len(indices), target_dim] tensor on the device."""

    def __init__(self, model, store, batch_size=2, mode="eval", sigma_step=1.0):
        from .trainer import BucketLadder
        if mode not in ("eval", "reference"):
            raise ValueError("mode must be 'eval' or 'reference'")
        self.store, self.B, self.mode = store, int(batch_size), mode
        self.model = fold_batch_norm(model).to(store.device) if mode == "eval" else model.to(store.device).train()
        gnn = getattr(self.model, "node_gnn", None)
        if gnn is not None and hasattr(gnn, "keep_edge_side_effects"):
            gnn.keep_edge_side_effects = False
        self._ladder_cls, self.sigma_step = BucketLadder, float(sigma_step)
        self.ladders, self.buckets = {}, {}
        self.pool = torch.cuda.graph_pool_handle()

    def _ladder(self, B):
        lad = self.ladders.get(B)
        if lad is None:
            lad = self.ladders[B] = self._ladder_cls(self.store, B, 1, self.sigma_step)
        return lad

    def _forward(self, meta, B, caps):
        g2, _ = self.store.collate_padded(meta, B, *caps, conformers=1, need_3d=False)
        with torch.no_grad():
            return self.model(g2)

    def _bucket(self, B, level, idx):
        caps = self._ladder(B).caps(level)
        key = (B,) + tuple(caps[:2])            # small batches: many levels round to the same capacities
        bk = self.buckets.get(key)
        if bk is not None:
            return bk
        from .collate import metadata_len
        dev = self.store.device
        meta = torch.zeros(metadata_len(B), dtype=torch.int64, device=dev)
        self.store.stage_metadata(idx, dev_out=meta)
        # eager warm-up (lazy initialisation) with the BatchNorm buffers put back: 'reference' mode updates them
        bufs = [b for b in self.model.buffers()]
        snap = [b.clone() for b in bufs]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            out = self._forward(meta, B, caps)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for b, s in zip(bufs, snap):
                b.copy_(s)
        static_out = torch.zeros_like(out)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self.pool):
            static_out.copy_(self._forward(meta, B, caps))
        bk = self.buckets[key] = (graph, meta, static_out)
        return bk

    def __call__(self, indices=None):
        store = self.store
        idx_all = np.arange(len(store)) if indices is None else np.asarray(indices, dtype=np.int64).reshape(-1)
        outs = []
        for i in range(0, len(idx_all), self.B):
            idx = idx_all[i:i + self.B]
            B = len(idx)
            level = self._ladder(B).level_of(store.batch_sizes(idx, 1))
            graph, meta, out = self._bucket(B, level, idx)
            store.stage_metadata(idx, dev_out=meta)
            graph.replay()
            outs.append(out.clone())
        return torch.cat(outs, dim=0) if outs else torch.zeros(0, device=store.device)
