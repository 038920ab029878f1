"""``PNA`` — drop-in for the reference's ``model_type: 'PNA'`` (models/pna.py:90-252) on B200.

Same constructor kwargs (unknown ones swallowed by ``**kwargs``, models/pna.py:115), same state-dict keys
(SURVEY.md §8b), same input contract (a batched graph with int64 ``ndata['feat']`` [N,9] and
``edata['feat']`` [E,3]) and the same visible side effect (``ndata['feat']`` ends as the final node states).

Per layer (models/pna.py:199-213), in kernels:
  pretrans   : ONE gather-fused GEMM over the virtual cat[h[src], h[dst], e] -> ReLU/BN -> GEMM -> BN
               (edge rows are produced directly in CSR order, so each node's mailbox is contiguous);
  aggregate  : ONE segmented reduction writing [mean|max|min|std]  (replaces DGL degree bucketing, ~25 launches
               per distinct degree plus a host sync);
  posttrans  : ONE GEMM over the virtual cat[h, agg, agg*ln(D+1), agg/ln(D+1)] (degree scalers applied to the
               operand tile as it is staged, never materialised) -> BN -> + residual.
"""
import os

import torch
from torch import nn

from . import ops
from .base_layers import MLP
from .graph import graph_structure
from .synthetic import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS

_AGG_ORDER = ["mean", "max", "min", "std"]
_SCALER_ORDER = ["identity", "amplification", "attenuation"]


class _SummedEmbedding(nn.Module):
    """Sum of per-column embeddings (commons/mol_encoder.py:10-73) evaluated by one gather-sum kernel."""

    def __init__(self, dims, emb_dim, list_name):
        super().__init__()
        tables = nn.ModuleList()
        for d in dims:
            e = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(e.weight.data)          # mol_encoder.py:27,62
            tables.append(e)
        setattr(self, list_name, tables)
        self._list_name = list_name
        offs, acc = [], 0
        for d in dims:
            offs.append(acc)
            acc += d
        self.register_buffer("_col_off", torch.tensor(offs, dtype=torch.int32), persistent=False)
        self._max_dim = max(dims)

    def forward(self, idx, perm=None):
        tables = getattr(self, self._list_name)
        # pointer plumbing only: one contiguous table so the kernel takes a single base pointer; autograd's
        # cat backward splits the table gradient back onto the per-column nn.Embedding weights
        table = torch.cat([e.weight for e in tables], dim=0)
        return ops.embed_sum(idx.contiguous(), self._col_off, perm, table, self._max_dim)


class AtomEncoder(_SummedEmbedding):
    def __init__(self, emb_dim):
        super().__init__(ATOM_FEATURE_DIMS, emb_dim, "atom_embedding_list")


class BondEncoder(_SummedEmbedding):
    def __init__(self, emb_dim):
        super().__init__(BOND_FEATURE_DIMS, emb_dim, "bond_embedding_list")


class PNALayer(nn.Module):
    def __init__(self, in_dim, out_dim, in_dim_edges, aggregators, scalers, activation="relu",
                 last_activation="none", dropout=0.0, residual=True, pairwise_distances=False, mid_batch_norm=False,
                 last_batch_norm=False, batch_norm_momentum=0.1, avg_d=None, posttrans_layers=2, pretrans_layers=1):
        super().__init__()
        if list(aggregators) != _AGG_ORDER:
            raise NotImplementedError("fused aggregation kernel implements aggregators %s (every shipped config); "
                                      "got %s" % (_AGG_ORDER, list(aggregators)))
        if list(scalers) != _SCALER_ORDER:
            raise NotImplementedError("posttrans GEMM folds scalers %s (every shipped config); got %s"
                                      % (_SCALER_ORDER, list(scalers)))
        if pairwise_distances:
            raise NotImplementedError("pairwise_distances=True is unused by the target configs")
        if in_dim_edges != in_dim:
            raise NotImplementedError("edge feature width must equal the node width (models/pna.py:150)")
        self.residual = bool(residual) and in_dim == out_dim
        self.pretrans = MLP(in_dim=2 * in_dim + in_dim_edges, hidden_size=in_dim, out_dim=in_dim,
                            mid_batch_norm=mid_batch_norm, last_batch_norm=last_batch_norm, layers=pretrans_layers,
                            mid_activation=activation, dropout=dropout, last_activation=last_activation,
                            batch_norm_momentum=batch_norm_momentum)
        self.posttrans = MLP(in_dim=(len(aggregators) * len(scalers) + 1) * in_dim, hidden_size=out_dim,
                             out_dim=out_dim, layers=posttrans_layers, mid_activation=activation,
                             last_activation=last_activation, dropout=dropout, mid_batch_norm=mid_batch_norm,
                             last_batch_norm=last_batch_norm, batch_norm_momentum=batch_norm_momentum)

    def forward(self, st, h, ef_csr, edge_codes=None, table=None, combo=None):
        # models/pna.py:203,237-252 — edge MLP over cat[h[src], h[dst], e], rows emitted in CSR order
        if edge_codes is not None:
            # factored first layer: node-level GEMM + 60-row bond-feature table + one gather-add pass (ops._FCEdgeFactored)
            fcs = self.pretrans.fully_connected
            msg = fcs[0].forward_edge_factored(edge_codes, h, table, st.e_valid, combo)
            for i in range(1, len(fcs)):
                msg = fcs[i](msg, None, st.e_valid)
        else:
            msg = self.pretrans([ops.Seg(h, idx=st.src_csr, inv_rowptr=st.out_rowptr, inv_idx=st.out_pos),
                                 ops.Seg(h, idx=st.dst_csr, inv_rowptr=st.rowptr),
                                 ops.Seg(ef_csr)], valid=st.e_valid)
        # models/pna.py:206,221-235
        agg = ops.pna_aggregate(msg, st.rowptr)
        # models/pna.py:207-211 — cat[h, agg, agg*amp, agg*att] -> posttrans -> + h
        res = h if self.residual else None
        if st.plan is not None:
            # the degree scalers depend on the in-degree only: nodes grouped by degree share one merged weight
            # W_id + amp_D W_amp + att_D W_att, and the first posttrans FC runs with K = 5F instead of 13F
            fcs = self.posttrans.fully_connected
            x = fcs[0].forward_merged(st.plan, h, agg, res if len(fcs) == 1 else None, st.n_valid)
            for i in range(1, len(fcs)):
                x = fcs[i](x, res if i == len(fcs) - 1 else None, st.n_valid)
            return x, msg
        return self.posttrans([ops.Seg(h), ops.Seg(agg), ops.Seg(agg, scale=st.amp), ops.Seg(agg, scale=st.att)],
                              residual=res, valid=st.n_valid), msg


class PNAGNN(nn.Module):
    def __init__(self, hidden_dim, aggregators, scalers, residual=True, pairwise_distances=False, activation="relu",
                 last_activation="none", mid_batch_norm=False, last_batch_norm=False, batch_norm_momentum=0.1,
                 propagation_depth=5, dropout=0.0, posttrans_layers=1, pretrans_layers=1, **kwargs):
        super().__init__()
        self.mp_layers = nn.ModuleList([
            PNALayer(in_dim=hidden_dim, out_dim=int(hidden_dim), in_dim_edges=hidden_dim, aggregators=aggregators,
                     scalers=scalers, pairwise_distances=pairwise_distances, residual=residual, dropout=dropout,
                     activation=activation, last_activation=last_activation, mid_batch_norm=mid_batch_norm,
                     last_batch_norm=last_batch_norm, avg_d={"log": 1.0}, posttrans_layers=posttrans_layers,
                     pretrans_layers=pretrans_layers, batch_norm_momentum=batch_norm_momentum)
            for _ in range(propagation_depth)])
        self.atom_encoder = AtomEncoder(emb_dim=hidden_dim)
        self.bond_encoder = BondEncoder(emb_dim=hidden_dim)
        self.keep_edge_side_effects = True
        # every combination of the categorical bond features (5*6*2 = 60 rows) and the mixed-radix weights that turn an
        # edge's feature row into its combination index: lets the `e` K-segment of the edge MLP become a table lookup
        dims = list(BOND_FEATURE_DIMS)
        n_codes = 1
        for d in dims:
            n_codes *= d
        self.n_bond_codes = n_codes
        mult, acc = [], 1
        for d in reversed(dims):
            mult.append(acc)
            acc *= d
        mult = list(reversed(mult))
        grid = torch.stack(torch.meshgrid(*[torch.arange(d) for d in dims], indexing="ij"), dim=-1).reshape(-1, len(dims))
        self.register_buffer("_bond_combos", grid.long().contiguous(), persistent=False)
        self.register_buffer("_bond_mult", torch.tensor(mult, dtype=torch.int64), persistent=False)

    def forward(self, graph):
        st = graph_structure(graph)
        x_atom, e_attr = graph.ndata["feat"], graph.edata["feat"]
        if x_atom.dtype != torch.int64 or e_attr.dtype != torch.int64:
            raise TypeError("PNA expects int64 categorical features in ndata['feat'] / edata['feat'] "
                            "(the graph is consumed by forward, as in the reference)")
        h = self.atom_encoder(x_atom)                              # models/pna.py:162
        factored = (os.environ.get("I3D_PRETRANS", "factored") != "gemm" and self.n_bond_codes <= 256
                    and e_attr.shape[1] == self._bond_mult.numel())
        edge_codes = tables = ef_csr = combo = None
        if factored:
            # models/pna.py:163 — the bond embedding of an edge is one of n_bond_codes rows: embed the combinations once,
            # and turn the `e` segment of every layer's edge MLP into a table over them (one launch for all layers)
            combo = self.bond_encoder(self._bond_combos)
            F = h.shape[1]
            tables = ops.bond_tables(combo, [l.pretrans.fully_connected[0].linear.weight for l in self.mp_layers], 2 * F,
                                     weight_grads=False)
            code_csr = getattr(st, "code_csr", None)                # emitted by the device collate when it built `st`
            if code_csr is None:
                code_csr = (e_attr * self._bond_mult).sum(dim=1)[st.eid.long()]
            edge_codes = ops.EdgeCodes(st, code_csr, self.n_bond_codes)
        else:
            ef_csr = self.bond_encoder(e_attr, perm=st.eid)        # models/pna.py:163, emitted in CSR order
        graph.ndata["feat"] = h
        if self.keep_edge_side_effects:
            with torch.no_grad():
                graph.edata["feat"] = self.bond_encoder(e_attr)    # edge-id order, as the reference leaves it
        msg = None
        for li, layer in enumerate(self.mp_layers):
            h, msg = layer(st, h, ef_csr, edge_codes, None if tables is None else tables[li], combo if factored else None)
            graph.ndata["feat"] = h                                # models/pna.py:213
        if self.keep_edge_side_effects and msg is not None:
            # models/pna.py:203,252: apply_edges leaves the last layer's messages in edata['e'] (edge-id order; ours
            # are in CSR order: position k holds edge eid[k]).  A pure permutation copy, outside the autograd graph.
            with torch.no_grad():
                e = torch.zeros_like(msg)
                e[st.eid.long()] = msg
                graph.edata["e"] = e
        return st, h


class PNA(nn.Module):
    """Message passing network over the 2-D molecular graph (no 3-D information)."""

    def __init__(self, hidden_dim, target_dim, aggregators, scalers, readout_aggregators, readout_batchnorm=True,
                 readout_hidden_dim=None, readout_layers=2, residual=True, pairwise_distances=False,
                 activation="relu", last_activation="none", mid_batch_norm=False, last_batch_norm=False,
                 propagation_depth=5, dropout=0.0, posttrans_layers=1, pretrans_layers=1, batch_norm_momentum=0.1,
                 **kwargs):
        super().__init__()
        self.node_gnn = PNAGNN(hidden_dim=hidden_dim, aggregators=aggregators, scalers=scalers, residual=residual,
                               pairwise_distances=pairwise_distances, activation=activation,
                               last_activation=last_activation, mid_batch_norm=mid_batch_norm,
                               last_batch_norm=last_batch_norm, propagation_depth=propagation_depth, dropout=dropout,
                               posttrans_layers=posttrans_layers, pretrans_layers=pretrans_layers,
                               batch_norm_momentum=batch_norm_momentum)
        if readout_hidden_dim is None:
            readout_hidden_dim = hidden_dim
        self.readout_aggregators = list(readout_aggregators)
        self.output = MLP(in_dim=hidden_dim * len(self.readout_aggregators), hidden_size=readout_hidden_dim,
                          mid_batch_norm=readout_batchnorm, out_dim=target_dim, layers=readout_layers,
                          batch_norm_momentum=batch_norm_momentum)

    def forward(self, graph):
        st, h = self.node_gnn(graph)
        ro = ops.readout(h, st.graph_ptr, self.readout_aggregators)      # models/pna.py:133-134
        return self.output(ro)                                           # models/pna.py:135
