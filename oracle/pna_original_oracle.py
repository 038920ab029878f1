"""TEST INFRASTRUCTURE — CPU restatement of the reference's tower PNA (``model_type: 'PNAOriginal'``,
models/pna_original.py:119-319) in functional fp32 torch.  Only tests/ may import this; the product path
(3dinfomax_b200/pna_original.py) never does.  Parity pinned: oracle/pin_pna_original.py runs the reference's own
``PNAOriginal`` (unmodified, under the dgl shim of oracle/ref_under_shim.py) on seeded graphs and weights and asserts
equality with this file (tests/golden/pna_original_*.npz).

Structure followed:
  PNAOriginal.forward (:142-149)       h = AtomEncoder(x), e = BondEncoder(a) (edge_hidden_dim wide), L x PNALayer,
                                        readout cat over readout_aggregators, MLPReadout (base_layers.py:149-164)
  PNALayer.forward (:304-319)          towers on h[:, t*Ft:(t+1)*Ft] (divide_input) or on the full h, cat,
                                        LeakyReLU(mixing_network(.)), + h_in when in_dim == out_dim and residual
  PNATower.forward (:240-261)          pretrans MLP (no BatchNorm) over cat[h_src, h_dst, e]; reduce = aggregators x
                                        scalers with the SCALAR avg_d (:28-35: amp = ln(D+1)/avg_d, att = avg_d/ln(D+1));
                                        posttrans MLP (BatchNorm per mid/last_batch_norm) over cat[h, agg]; * snorm_n
                                        when graph_norm; dropout (0 in the configs)
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle as O

DEFAULTS = dict(in_feat_dropout=0.0, dropout=0.0, pretrans_layers=1, posttrans_layers=1, divide_input_first=True,
                divide_input_last=True, edge_feat=True, towers=1, use_3d=False, gru_enable=False, residual=True)

# configs/contrastive_training_pna_original.yml:50-85
CONTRASTIVE_PNA_ORIGINAL = dict(
    target_dim=256, hidden_dim=70, last_layer_dim=70, mid_batch_norm=True, last_batch_norm=True, graph_norm=True,
    readout_batchnorm=True, edge_hidden_dim=30, readout_hidden_dim=100, readout_layers=2, dropout=0.0,
    in_feat_dropout=0.0, propagation_depth=4, towers=5, divide_input_first=False, divide_input_last=True,
    aggregators=["mean", "max", "min", "std"], scalers=["identity", "amplification", "attenuation"],
    readout_aggregators=["mean", "max", "min", "sum"], pretrans_layers=1, posttrans_layers=1, residual=True, gru=False)


def cfg(**kw):
    c = dict(DEFAULTS)
    c.update(kw)
    return c


def layer_dims(c):
    """[(in_dim, out_dim, divide_input)] per PNALayer (pna_original.py:164-181)."""
    H, Lst, depth = c["hidden_dim"], c["last_layer_dim"], c["propagation_depth"]
    return [(H, H, c["divide_input_first"])] * (depth - 1) + [(H, Lst, c["divide_input_last"])]


def mlp_spec(in_dim, hidden, out_dim, layers, mid_bn, last_bn):
    """[(din, dout, act, bn)] of base_layers.MLP with mid_activation relu / last_activation none (:119-142)."""
    if layers <= 1:
        return [(in_dim, out_dim, "none", last_bn)]
    spec = [(in_dim, hidden, "relu", mid_bn)]
    spec += [(hidden, hidden, "relu", mid_bn)] * (layers - 2)
    return spec + [(hidden, out_dim, "none", last_bn)]


def tower_specs(c, in_dim, out_dim, divide):
    T = c["towers"]
    it = in_dim // T if divide else in_dim
    ot = out_dim // T
    n_agg = len(c["aggregators"]) * len(c["scalers"])
    pre = mlp_spec(2 * it + (c["edge_hidden_dim"] if c["edge_feat"] else 0), it, it, c["pretrans_layers"], False, False)
    post = mlp_spec((n_agg + 1) * it, ot, ot, c["posttrans_layers"], c["mid_batch_norm"], c["last_batch_norm"])
    return it, ot, pre, post


def init_state(c, seed, trained_scale=True):
    """Seeded state dict with the reference's keys and shapes (checked by the pin script with strict=True)."""
    gen = torch.Generator().manual_seed(seed)
    st = OrderedDict()
    rnd = lambda *s, std=1.0: torch.randn(*s, generator=gen) * std
    H = c["hidden_dim"]
    for i, d in enumerate(O.ATOM_DIMS):
        st["node_gnn.embedding_h.atom_embedding_list.%d.weight" % i] = rnd(d, H, std=0.09)
    for i, d in enumerate(O.BOND_DIMS):
        st["node_gnn.embedding_e.bond_embedding_list.%d.weight" % i] = rnd(d, c["edge_hidden_dim"], std=0.09)

    def mlp(prefix, spec):
        for j, (din, dout, _act, bn) in enumerate(spec):
            p = "%s.fully_connected.%d" % (prefix, j)
            st[p + ".linear.weight"] = rnd(dout, din, std=0.6 / math.sqrt(din))
            st[p + ".linear.bias"] = rnd(dout, std=0.05)
            if bn:
                st[p + ".batch_norm.weight"] = 1.0 + 0.03 * rnd(dout)
                st[p + ".batch_norm.bias"] = 0.05 * rnd(dout)
                st[p + ".batch_norm.running_mean"] = 0.1 * rnd(dout)
                st[p + ".batch_norm.running_var"] = 0.05 + torch.rand(dout, generator=gen)
                st[p + ".batch_norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for l, (din, dout, divide) in enumerate(layer_dims(c)):
        _it, _ot, pre, post = tower_specs(c, din, dout, divide)
        for t in range(c["towers"]):
            mlp("node_gnn.layers.%d.towers.%d.pretrans" % (l, t), pre)
            mlp("node_gnn.layers.%d.towers.%d.posttrans" % (l, t), post)
        st["node_gnn.layers.%d.mixing_network.weight" % l] = rnd(dout, dout, std=0.6 / math.sqrt(dout))
        st["node_gnn.layers.%d.mixing_network.bias" % l] = rnd(dout, std=0.05)

    def readout(prefix, din, dout, L=2):
        dims = [din // 2 ** i for i in range(L + 1)]
        for i in range(L):
            st["%s.FC_layers.%d.weight" % (prefix, i)] = rnd(dims[i + 1], dims[i], std=0.6 / math.sqrt(dims[i]))
            st["%s.FC_layers.%d.bias" % (prefix, i)] = rnd(dims[i + 1], std=0.05)
        st["%s.FC_layers.%d.weight" % (prefix, L)] = rnd(dout, dims[L], std=0.6 / math.sqrt(dims[L]))
        st["%s.FC_layers.%d.bias" % (prefix, L)] = rnd(dout, std=0.05)

    readout("node_gnn.MLP_layer", H, 1)                                  # :185, present in the state dict, unused
    readout("output", c["last_layer_dim"] * len(c["readout_aggregators"]), c["target_dim"])
    return st


def reduce_scalar_avg(g, msg, aggregators, scalers, avg_d):
    """PNATower.reduce_func (:232-237) through DGL degree bucketing, scalers with the scalar avg_d (:28-35)."""
    width = len(aggregators) * len(scalers) * msg.shape[1]
    out = torch.zeros(g.n, width, dtype=msg.dtype)
    for D, nodes, eids in g.buckets():
        mail = msg[eids.reshape(-1)].reshape(len(nodes), D, -1)
        h = torch.cat([O._aggregate(mail, a) for a in aggregators], dim=1)
        parts = []
        for s in scalers:
            if s == "identity":
                parts.append(h)
            elif s == "amplification":
                parts.append(h * (np.log(D + 1) / avg_d))
            elif s == "attenuation":
                parts.append(h * (avg_d / np.log(D + 1)))
            else:
                raise NotImplementedError(s)
        out = out.index_put((nodes,), torch.cat(parts, dim=1))
    return out


def mlp_readout(x, st, prefix, L=2):
    for i in range(L):                                                   # base_layers.py:158-164
        x = torch.relu(F.linear(x, st["%s.FC_layers.%d.weight" % (prefix, i)], st["%s.FC_layers.%d.bias" % (prefix, i)]))
    return F.linear(x, st["%s.FC_layers.%d.weight" % (prefix, L)], st["%s.FC_layers.%d.bias" % (prefix, L)])


def forward(st, c, g, x_atom, e_attr, snorm_n, avg_d, training=True):
    """PNAOriginal.forward(g, snorm_n) -> [B, target_dim]."""
    mom = 0.1                                                            # MLP default batch_norm_momentum
    h = O.embed_sum(x_atom, st, "node_gnn.embedding_h.atom_embedding_list.%d.weight", x_atom.shape[1])
    e = O.embed_sum(e_attr, st, "node_gnn.embedding_e.bond_embedding_list.%d.weight", e_attr.shape[1])
    T = c["towers"]
    for l, (din, dout, divide) in enumerate(layer_dims(c)):
        it, ot, pre, post = tower_specs(c, din, dout, divide)
        outs = []
        for t in range(T):
            ht = h[:, t * it:(t + 1) * it] if divide else h              # :307-313
            p = "node_gnn.layers.%d.towers.%d" % (l, t)
            z = torch.cat([ht[g.src], ht[g.dst], e], dim=1)              # :217-218
            msg = O.mlp_forward(z, st, p + ".pretrans", pre, training, mom)
            agg = reduce_scalar_avg(g, msg, c["aggregators"], c["scalers"], avg_d)
            x = O.mlp_forward(torch.cat([ht, agg], dim=1), st, p + ".posttrans", post, training, mom)   # :252-255
            if c["graph_norm"]:
                x = x * snorm_n                                          # :258-259
            outs.append(x)
        hc = torch.cat(outs, dim=1)
        ho = F.leaky_relu(F.linear(hc, st["node_gnn.layers.%d.mixing_network.weight" % l],
                                   st["node_gnn.layers.%d.mixing_network.bias" % l]))            # :315
        h = h + ho if (c["residual"] and din == dout) else ho           # :317-318
    ro = torch.cat([O.segment_readout(h, g, op) for op in c["readout_aggregators"]], dim=-1)
    return mlp_readout(ro, st, "output")
