"""``Net3D`` — drop-in for the reference's ``model3d_type: 'Net3D'`` (models/net3d.py:14-125) on B200.

Input: a batched complete graph per conformer whose only information is ``edata['d']`` fp32 [E3,1], the
pairwise distances.  Same kwargs / state-dict keys as the reference.  All edge tensors are kept in CSR
(dst-sorted) order internally, so the builtin mean/sum reduce is a contiguous segment reduction and
the h[src], h[dst] gathers are fused into the message GEMM.
"""
import torch
from torch import nn

from . import ops
from . import kernels as K
from .base_layers import MLP
from .graph import graph_structure


class Net3DLayer(nn.Module):
    def __init__(self, edge_dim, reduce_func, hidden_dim, batch_norm, batch_norm_momentum, dropout, mid_activation,
                 message_net_layers, update_net_layers):
        super().__init__()
        if edge_dim != hidden_dim:
            raise NotImplementedError("edge width must equal the hidden width (models/net3d.py:36)")
        self.message_network = MLP(in_dim=hidden_dim * 2 + edge_dim, hidden_size=hidden_dim, out_dim=hidden_dim,
                                   mid_batch_norm=batch_norm, last_batch_norm=batch_norm,
                                   batch_norm_momentum=batch_norm_momentum, layers=message_net_layers,
                                   mid_activation=mid_activation, dropout=dropout, last_activation=mid_activation)
        if reduce_func not in ("sum", "mean"):
            raise ValueError("reduce function not supported: ", reduce_func)     # models/net3d.py:98
        self.reduce_mean = reduce_func == "mean"
        self.update_network = MLP(in_dim=hidden_dim, hidden_size=hidden_dim, out_dim=hidden_dim,
                                  mid_batch_norm=batch_norm, last_batch_norm=batch_norm,
                                  batch_norm_momentum=batch_norm_momentum, layers=update_net_layers,
                                  mid_activation=mid_activation, dropout=dropout, last_activation="None")
        self.soft_edge_network = nn.Linear(hidden_dim, 1)

    def forward(self, st, h, d, update_edges):
        # models/net3d.py:112-118
        msg = self.message_network([ops.Seg(h, idx=st.src_csr, inv_rowptr=st.out_rowptr, inv_idx=st.out_pos),
                                    ops.Seg(h, idx=st.dst_csr, inv_rowptr=st.rowptr),
                                    ops.Seg(d)], valid=st.e_valid)
        d_next = ops.add(d, msg) if update_edges else None            # edges.data['d'] += message
        m = ops.soft_gate(msg, self.soft_edge_network.weight, self.soft_edge_network.bias)
        # fn.sum / fn.mean over in-edges, fused with "+ feat"  (models/net3d.py:94-96,122)
        agg = ops.segment_reduce(m, st.rowptr, st.dst_csr, self.reduce_mean, addend=h)
        # models/net3d.py:120-125
        return self.update_network(agg, residual=h, valid=st.n_valid), d_next


class Net3D(nn.Module):
    def __init__(self, node_dim, edge_dim, hidden_dim, target_dim, readout_aggregators, batch_norm=False,
                 node_wise_output_layers=2, readout_batchnorm=True, batch_norm_momentum=0.1, reduce_func="sum",
                 dropout=0.0, propagation_depth=4, readout_layers=2, readout_hidden_dim=None, fourier_encodings=0,
                 activation="SiLU", update_net_layers=2, message_net_layers=2, use_node_features=False, **kwargs):
        super().__init__()
        if use_node_features:
            raise NotImplementedError("use_node_features=True is unused by the target configs")
        self.fourier_encodings = fourier_encodings
        edge_in_dim = 1 if fourier_encodings == 0 else 2 * fourier_encodings + 1
        self.edge_input = MLP(in_dim=edge_in_dim, hidden_size=hidden_dim, out_dim=hidden_dim,
                              mid_batch_norm=batch_norm, last_batch_norm=batch_norm,
                              batch_norm_momentum=batch_norm_momentum, layers=1, mid_activation=activation,
                              dropout=dropout, last_activation=activation)
        self.node_embedding = nn.Parameter(torch.empty((hidden_dim,)))
        nn.init.normal_(self.node_embedding)
        self.mp_layers = nn.ModuleList([
            Net3DLayer(edge_dim=hidden_dim, hidden_dim=hidden_dim, batch_norm=batch_norm,
                       batch_norm_momentum=batch_norm_momentum, dropout=dropout, mid_activation=activation,
                       reduce_func=reduce_func, message_net_layers=message_net_layers,
                       update_net_layers=update_net_layers) for _ in range(propagation_depth)])
        self.node_wise_output_layers = node_wise_output_layers
        if node_wise_output_layers > 0:
            self.node_wise_output_network = MLP(in_dim=hidden_dim, hidden_size=hidden_dim, out_dim=hidden_dim,
                                                mid_batch_norm=batch_norm, last_batch_norm=batch_norm,
                                                batch_norm_momentum=batch_norm_momentum,
                                                layers=node_wise_output_layers, mid_activation=activation,
                                                dropout=dropout, last_activation="None")
        if readout_hidden_dim is None:
            readout_hidden_dim = hidden_dim
        self.readout_aggregators = list(readout_aggregators)
        self.output = MLP(in_dim=hidden_dim * len(self.readout_aggregators), hidden_size=readout_hidden_dim,
                          mid_batch_norm=readout_batchnorm, batch_norm_momentum=batch_norm_momentum,
                          out_dim=target_dim, layers=readout_layers)

    def forward(self, graph):
        st = graph_structure(graph, need_scalers=False)
        dist = graph.edata["d"]
        if dist.dtype != torch.float32:
            raise TypeError("Net3D expects fp32 distances in edata['d']")
        h = ops.broadcast_rows(self.node_embedding, st.N)                      # models/net3d.py:61
        graph.ndata["feat"] = h
        dist = dist.reshape(-1).contiguous()
        # commons/utils.py:103-110, emitted in CSR order (k = 0: the raw distance column only)
        e_in = K.fourier_encode(dist, st.eid, self.fourier_encodings)
        d = ops.activation(self.edge_input(e_in, valid=st.e_valid), "silu")     # models/net3d.py:80-81
        n_layers = len(self.mp_layers)
        for i, layer in enumerate(self.mp_layers):
            h, d = layer(st, h, d, update_edges=i + 1 < n_layers)
        if self.node_wise_output_layers > 0:
            h = self.node_wise_output_network(h, valid=st.n_valid)              # models/net3d.py:70-71
        graph.ndata["feat"] = h
        ro = ops.readout(h, st.graph_ptr, self.readout_aggregators)            # models/net3d.py:73-74
        return self.output(ro)                                                  # models/net3d.py:75
