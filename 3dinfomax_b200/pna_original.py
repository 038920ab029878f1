"""``PNAOriginal`` — drop-in for the reference's tower PNA (``model_type: 'PNAOriginal'``,
models/pna_original.py:119-319; configs/contrastive_training_pna_original.yml) on the same kernels as ``PNA``.

Same constructor kwargs, same ``forward(graph, snorm_n)`` signature, same state-dict keys:
``node_gnn.embedding_h / embedding_e`` (AtomEncoder / BondEncoder, the latter ``edge_hidden_dim`` wide),
``node_gnn.layers.{l}.towers.{t}.{pretrans,posttrans}.fully_connected.{j}.{linear,batch_norm}``,
``node_gnn.layers.{l}.mixing_network.{weight,bias}``, ``node_gnn.MLP_layer.FC_layers.{i}`` (present but unused in the
reference too, :185) and ``output.FC_layers.{i}`` (MLPReadout, models/base_layers.py:149-164).

Per tower (PNATower.forward, :240-261): the pretrans MLP runs as ONE gather-fused GEMM over the virtual
cat[h_t[src], h_t[dst], e] (rows emitted in CSR order), the fused multi-aggregator kernel reduces the mailbox, the
posttrans MLP runs over the virtual cat[h_t, A, A*amp, A*att] with amp = ln(D+1)/avg_d, att = avg_d/ln(D+1) (the
SCALAR avg_d of this model, :28-35), then ``* snorm_n`` when ``graph_norm``.  Towers read column slices of ``h``
in place (``divide_input``) — the slices are views, the GEMM loader takes their leading dimension.  Tower widths of
the shipped configs (14 = 70/5, 50 = 200/4) are not multiples of 4, so tower-by-tower these GEMMs take the fp32 SIMT
kernel.

Fused towers (default when the layer divides its input between >1 towers of single-FC MLPs and the total widths are
multiples of 4, i.e. the "hidden 200, 4 towers" shape): T towers over disjoint column slices are ONE layer with
block-diagonal weights.  The per-tower parameters stay what they are (state-dict parity); every step assembles
``W_pre [F, 2F + Fe]`` and ``W_post [F_out, 13F]`` from them (two einsums against the TxT identity, differentiable, so
the block gradients flow back to the tower parameters), and the layer runs as one tensor-core GEMM each over the full
width — the aggregation and BatchNorm are column-wise, so concatenated towers are exactly the per-tower results.  The
towers' BatchNorm buffers are re-homed as views of one [F_out] buffer (same keys, same values).  This path also takes
the shape-bucketed (padded) batches of trainer.BucketedStep.
"""
import math
import os

import torch
from torch import nn

from . import kernels as K
from . import ops
from .base_layers import MLP, _Linear, activation_code
from .graph import graph_structure
from .pna import AtomEncoder, BondEncoder, _AGG_ORDER, _SCALER_ORDER


def _linear_default_init(lin, in_dim):
    """nn.Linear's default initialisation (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for both)."""
    bound = 1.0 / math.sqrt(in_dim)
    with torch.no_grad():
        lin.weight.uniform_(-bound, bound)
        lin.bias.uniform_(-bound, bound)


class MLPReadout(nn.Module):
    """models/base_layers.py:149-164: L halving ReLU layers and a linear head, plain nn.Linear parameters."""

    def __init__(self, input_dim, output_dim, L=2):
        super().__init__()
        dims = [input_dim // 2 ** l for l in range(L + 1)]
        layers = []
        for l in range(L):
            layers.append(_Linear(dims[l], dims[l + 1]))
            _linear_default_init(layers[-1], dims[l])
        layers.append(_Linear(dims[L], output_dim))
        _linear_default_init(layers[-1], dims[L])
        self.FC_layers = nn.ModuleList(layers)
        self.L = L

    def forward(self, x):
        relu, none = activation_code("relu"), activation_code("none")
        for l, lin in enumerate(self.FC_layers):
            x = ops.fc([ops.Seg(x)], lin.weight, lin.bias, relu if l < self.L else none, None, self.training)
        return x


class PNATower(nn.Module):
    def __init__(self, in_dim, out_dim, dropout, graph_norm, mid_batch_norm, last_batch_norm, aggregators, scalers,
                 avg_d, use_3d, pretrans_layers, posttrans_layers, edge_features, edge_hidden_dim):
        super().__init__()
        if dropout:
            raise NotImplementedError("dropout > 0 has no kernel (the shipped configs use 0.0)")
        if use_3d:
            raise NotImplementedError("use_3d=True (pairwise distances in the message) is not used by the target configs")
        if list(aggregators) != _AGG_ORDER or list(scalers) != _SCALER_ORDER:
            raise NotImplementedError("the fused aggregation implements aggregators %s x scalers %s"
                                      % (_AGG_ORDER, _SCALER_ORDER))
        self.graph_norm = graph_norm
        self.edge_features = edge_features
        self.pretrans = MLP(in_dim=2 * in_dim + (edge_hidden_dim if edge_features else 0), hidden_size=in_dim,
                            out_dim=in_dim, layers=pretrans_layers, mid_activation="relu", last_activation="none")
        self.posttrans = MLP(in_dim=(len(aggregators) * len(scalers) + 1) * in_dim, hidden_size=out_dim,
                             mid_batch_norm=mid_batch_norm, last_batch_norm=last_batch_norm, out_dim=out_dim,
                             layers=posttrans_layers, mid_activation="relu", last_activation="none")

    def forward(self, st, scalers, h, ef_csr, snorm_n):
        segs = [ops.Seg(h, idx=st.src_csr, inv_rowptr=st.out_rowptr, inv_idx=st.out_pos),
                ops.Seg(h, idx=st.dst_csr, inv_rowptr=st.rowptr)]
        if self.edge_features:
            segs.append(ops.Seg(ef_csr))
        msg = self.pretrans(segs)                                            # :207-221
        agg = ops.pna_aggregate(msg, st.rowptr)                              # :232-237, [N, 4 * in_dim]
        amp, att = scalers
        x = self.posttrans([ops.Seg(h), ops.Seg(agg), ops.Seg(agg, scale=amp), ops.Seg(agg, scale=att)])   # :252-255
        if self.graph_norm:
            x = ops.scale_rows(x, snorm_n)                                   # :258-259
        return x


class PNALayer(nn.Module):
    def __init__(self, in_dim, out_dim, aggregators, scalers, avg_d, dropout, graph_norm, mid_batch_norm, use_3d,
                 last_batch_norm, towers=1, pretrans_layers=1, posttrans_layers=1, divide_input=True, residual=False,
                 edge_features=False, edge_hidden_dim=0):
        super().__init__()
        assert (not divide_input) or in_dim % towers == 0, \
            "if divide_input is set the number of towers has to divide in_dim"
        assert out_dim % towers == 0, "the number of towers has to divide the last_layer_dim"
        assert avg_d is not None
        self.divide_input = divide_input
        self.input_tower = in_dim // towers if divide_input else in_dim
        self.output_tower = out_dim // towers
        self.in_dim, self.out_dim = in_dim, out_dim
        self.residual = bool(residual) and in_dim == out_dim
        self.towers = nn.ModuleList([
            PNATower(in_dim=self.input_tower, out_dim=self.output_tower, aggregators=aggregators, scalers=scalers,
                     avg_d=avg_d, pretrans_layers=pretrans_layers, posttrans_layers=posttrans_layers,
                     mid_batch_norm=mid_batch_norm, last_batch_norm=last_batch_norm, dropout=dropout, use_3d=use_3d,
                     graph_norm=graph_norm, edge_features=edge_features, edge_hidden_dim=edge_hidden_dim)
            for _ in range(towers)])
        self.mixing_network = _Linear(out_dim, out_dim)
        _linear_default_init(self.mixing_network, out_dim)
        self._leaky = activation_code("leakyrelu")

    # ---- fused towers -------------------------------------------------------------------------------------
    fuse_towers = True          # class default; set False on an instance (or I3D_TOWERS=loop) for the tower-by-tower path

    def _fusable(self):
        t0 = self.towers[0]
        return (self.fuse_towers and os.environ.get("I3D_TOWERS", "fused") != "loop" and self.divide_input
                and len(self.towers) > 1 and self.in_dim % 4 == 0 and self.out_dim % 4 == 0
                and len(t0.pretrans.fully_connected) == 1 and len(t0.posttrans.fully_connected) == 1)

    def _fused_bn(self, fcs, tag):
        """(gamma, beta, running_mean, running_var, nbt, momentum, eps) over the concatenated towers, or None.  The
        per-tower buffers become views of shared storage the first time (and again after a ``.to()`` replaced them)."""
        bns = [fc.batch_norm for fc in fcs]
        if bns[0] is None:
            return None
        T, w = len(bns), bns[0].running_mean.numel()
        big = self.__dict__.get("_bn_big_" + tag)
        ok = big is not None and big["rm"].device == bns[0].running_mean.device and all(
            bn.running_mean.data_ptr() == big["rm"].data_ptr() + 4 * t * w
            and bn.running_var.data_ptr() == big["rv"].data_ptr() + 4 * t * w
            and bn.num_batches_tracked.data_ptr() == big["nbt"].data_ptr() + 8 * t for t, bn in enumerate(bns))
        if not ok:
            with torch.no_grad():
                big = {"rm": torch.cat([bn.running_mean.reshape(-1) for bn in bns]).contiguous(),
                       "rv": torch.cat([bn.running_var.reshape(-1) for bn in bns]).contiguous(),
                       "nbt": torch.stack([bn.num_batches_tracked.reshape(()) for bn in bns]).contiguous()}
                for t, bn in enumerate(bns):
                    bn._buffers["running_mean"] = big["rm"][t * w:(t + 1) * w]
                    bn._buffers["running_var"] = big["rv"][t * w:(t + 1) * w]
                    bn._buffers["num_batches_tracked"] = big["nbt"][t]
            self.__dict__["_bn_big_" + tag] = big
        gamma = torch.cat([bn.weight for bn in bns])
        beta = torch.cat([bn.bias for bn in bns])
        return (gamma, beta, big["rm"], big["rv"], big["nbt"][0], bns[0].momentum, bns[0].eps)

    def _block_diagonal(self, Ws, blocks, width):
        """[T, rows, blocks * width] tower weights -> [T * rows, blocks * T * width]: block b of tower t lands on the
        columns of tower t inside block b of the full-width operand (zeros elsewhere)"""
        T, rows = Ws.shape[0], Ws.shape[1]
        eye = torch.eye(T, dtype=Ws.dtype, device=Ws.device)
        return torch.einsum("tobf,ts->tobsf", Ws.reshape(T, rows, blocks, width), eye).reshape(T * rows,
                                                                                             blocks * T * width)

    def _forward_fused(self, st, scalers, h, ef_csr, snorm_n):
        T, ft, ot = len(self.towers), self.input_tower, self.output_tower
        pre = [tw.pretrans.fully_connected[0] for tw in self.towers]
        post = [tw.posttrans.fully_connected[0] for tw in self.towers]
        t0 = self.towers[0]
        # pretrans (:207-221): cat[h_t[src], h_t[dst], e] W_t^T for all towers = cat[h[src], h[dst], e] W_pre^T
        Wp = torch.stack([fc.linear.weight for fc in pre])                   # [T, ft, 2 ft (+ Fe)]
        W_pre = self._block_diagonal(Wp[:, :, :2 * ft], 2, ft)
        segs = [ops.Seg(h, idx=st.src_csr, inv_rowptr=st.out_rowptr, inv_idx=st.out_pos),
                ops.Seg(h, idx=st.dst_csr, inv_rowptr=st.rowptr)]
        if t0.edge_features:
            W_pre = torch.cat([W_pre, Wp[:, :, 2 * ft:].reshape(T * ft, -1)], dim=1)
            segs.append(ops.Seg(ef_csr))
        b_pre = torch.cat([fc.linear.bias for fc in pre])
        msg = ops.fc(segs, W_pre.contiguous(), b_pre, pre[0].act, self._fused_bn(pre, "pre"), self.training, None,
                     st.e_valid)
        agg = ops.pna_aggregate(msg, st.rowptr)                              # :232-237, [N, 4 F], tower-major columns
        amp, att = scalers
        # posttrans (:252-255): 13 blocks [h | 4 aggregators x 3 scalers], each tower-major
        W_post = self._block_diagonal(torch.stack([fc.linear.weight for fc in post]), 13, ft)
        b_post = torch.cat([fc.linear.bias for fc in post])
        bn = self._fused_bn(post, "post")
        x = ops.fc([ops.Seg(h), ops.Seg(agg), ops.Seg(agg, scale=amp), ops.Seg(agg, scale=att)], W_post.contiguous(),
                   b_post, post[0].act, bn, self.training, None, st.n_valid)
        if bn is not None and self.training and T > 1:
            with torch.no_grad():
                self.__dict__["_bn_big_post"]["nbt"][1:] += 1      # the kernel counted the batch on tower 0's counter
        if t0.graph_norm:
            x = ops.scale_rows(x, snorm_n)                                   # :258-259
        return ops.fc([ops.Seg(x)], self.mixing_network.weight, self.mixing_network.bias, self._leaky, None,
                      self.training, h if self.residual else None, st.n_valid)    # :315-318

    def forward(self, st, scalers, h, ef_csr, snorm_n):
        if self._fusable():
            return self._forward_fused(st, scalers, h, ef_csr, snorm_n)
        if st.n_valid is not None:
            raise NotImplementedError("shape-bucketed (padded) batches need the fused-tower path of PNAOriginal")
        outs = []
        for t, tower in enumerate(self.towers):
            ht = h[:, t * self.input_tower:(t + 1) * self.input_tower] if self.divide_input else h      # :307-313
            outs.append(tower(st, scalers, ht, ef_csr, snorm_n))
        # virtual concatenation of the tower outputs as K-segments of the mixing GEMM (at most 4 segments per GEMM)
        segs = [ops.Seg(o) for o in outs] if len(outs) <= 4 else [ops.Seg(torch.cat(outs, dim=1))]
        return ops.fc(segs, self.mixing_network.weight, self.mixing_network.bias, self._leaky, None, self.training,
                      residual=h if self.residual else None)                 # :315-318


class PNAGNNOriginal(nn.Module):
    def __init__(self, hidden_dim, last_layer_dim, in_feat_dropout, dropout, propagation_depth, graph_norm,
                 mid_batch_norm, last_batch_norm, residual, aggregators, scalers, avg_d, use_3d, towers,
                 divide_input_first, divide_input_last, edge_feat, edge_hidden_dim, pretrans_layers, posttrans_layers,
                 gru_enable, device):
        super().__init__()
        if gru_enable:
            raise NotImplementedError("gru_enable=True is not used by the target configs")
        if in_feat_dropout:
            raise NotImplementedError("in_feat_dropout > 0 has no kernel (the shipped configs use 0.0)")
        self.edge_feat = edge_feat
        self.avg_d = float(avg_d)
        self.embedding_h = AtomEncoder(hidden_dim)
        if edge_feat:
            self.embedding_e = BondEncoder(edge_hidden_dim)
        mk = lambda out_dim, divide: PNALayer(
            in_dim=hidden_dim, out_dim=out_dim, dropout=dropout, graph_norm=graph_norm, mid_batch_norm=mid_batch_norm,
            last_batch_norm=last_batch_norm, use_3d=use_3d, residual=residual, aggregators=aggregators,
            scalers=scalers, avg_d=avg_d, towers=towers, edge_features=edge_feat, edge_hidden_dim=edge_hidden_dim,
            divide_input=divide, pretrans_layers=pretrans_layers, posttrans_layers=posttrans_layers)
        self.layers = nn.ModuleList([mk(hidden_dim, divide_input_first) for _ in range(propagation_depth - 1)]
                                    + [mk(last_layer_dim, divide_input_last)])
        self.MLP_layer = MLPReadout(hidden_dim, 1)                           # :185 (never called, kept for the keys)

    def forward(self, graph, snorm_n):
        st = graph_structure(graph)
        x_atom, e_attr = graph.ndata["feat"], graph.edata["feat"]
        if x_atom.dtype != torch.int64 or e_attr.dtype != torch.int64:
            raise TypeError("PNAOriginal expects int64 categorical features in ndata['feat'] / edata['feat']")
        h = self.embedding_h(x_atom)                                         # :188
        ef_csr = self.embedding_e(e_attr, perm=st.eid) if self.edge_feat else None   # :191, rows in CSR order
        scalers = K.degree_scalers(st.rowptr, self.avg_d)                    # :28-35 with the scalar avg_d
        snorm = snorm_n.to(device=h.device, dtype=torch.float32).reshape(-1).contiguous()
        for conv in self.layers:
            h = conv(st, scalers, h, ef_csr, snorm)                          # :193-197
        graph.ndata["feat"] = h                                              # :199
        return st, h


class PNAOriginal(nn.Module):
    needs_snorm = True          # forward(graph, snorm_n): trainer.BucketedStep builds snorm_n on the device

    def __init__(self, hidden_dim, last_layer_dim, target_dim, in_feat_dropout, dropout, last_batch_norm,
                 mid_batch_norm, propagation_depth, readout_aggregators, readout_hidden_dim, readout_layers,
                 aggregators, scalers, avg_d, residual, posttrans_layers, pretrans_layers, device, edge_hidden_dim,
                 graph_norm, use_3d=False, gru_enable=False, divide_input_last=True, divide_input_first=True,
                 edge_feat=True, towers=1, **kwargs):
        super().__init__()
        self.node_gnn = PNAGNNOriginal(
            hidden_dim=hidden_dim, last_layer_dim=last_layer_dim, last_batch_norm=last_batch_norm,
            mid_batch_norm=mid_batch_norm, in_feat_dropout=in_feat_dropout, dropout=dropout, aggregators=aggregators,
            scalers=scalers, residual=residual, avg_d=avg_d, propagation_depth=propagation_depth,
            posttrans_layers=posttrans_layers, device=device, pretrans_layers=pretrans_layers, gru_enable=gru_enable,
            use_3d=use_3d, edge_hidden_dim=edge_hidden_dim, divide_input_first=divide_input_first,
            divide_input_last=divide_input_last, edge_feat=edge_feat, graph_norm=graph_norm, towers=towers)
        self.readout_aggregators = list(readout_aggregators)
        self.output = MLPReadout(last_layer_dim * len(self.readout_aggregators), target_dim)

    def forward(self, g, snorm_n):
        st, h = self.node_gnn(g, snorm_n)                                    # :142-145
        ro = ops.readout(h, st.graph_ptr, self.readout_aggregators)          # :147-148
        return self.output(ro)
