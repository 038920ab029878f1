"""TEST INFRASTRUCTURE — CPU oracle for the 3DInfomax hot path.  NOT part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product (3dinfomax_b200/) never does and has no CPU fallback.

What it is: a from-scratch, functional, eager-PyTorch fp32 restatement on the CPU of the
reference's PNA / Net3D / NTXent forward pass (autograd supplies the backward), following the
reference op by op, including DGL's degree-bucketed reduce, so that it doubles as the timed
"reference CPU path" on hosts where /root/reference and DGL do not exist (the GPU box).

Parity status: PINNED.  oracle/make_golden.py runs the reference's own unmodified modules
(oracle/ref_under_shim.py) in the build container and (a) asserts this file agrees with them
(eval: exact; train: rounding level) and (b) writes tests/golden/*.npz, which
tests/test_oracle_golden.py re-checks everywhere.  The reference itself ships no tests or
golden vectors for this path; DGL semantics are DGL's documented behaviour.

State is a flat dict {state_dict key: tensor} with exactly the reference's key names
(SURVEY.md §8b), so one dict loads into the reference modules, this oracle and the product.

Reference lines followed (relative to /root/reference):
  models/base_layers.py:93-111,119-147   FC layer order Linear -> act -> BN; MLP stacking; init
  commons/mol_encoder.py:34-42,65-73     summed categorical embeddings
  models/pna.py:17-37,57-68              aggregators, scalers
  models/pna.py:131-135,161-166,199-252  PNA / PNAGNN / PNALayer forward
  commons/utils.py:103-110               fourier_encode_dist
  models/net3d.py:57-81,108-125          Net3D / Net3DLayer forward
  commons/losses.py:143-155,225-246      NTXent, NTXentMultiplePositives
  trainer/self_supervised_trainer.py:24-29,78-86; trainer/trainer.py:116-124   step + optimizer groups
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

ATOM_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]   # ogb.utils.features.get_atom_feature_dims()
BOND_DIMS = [5, 6, 2]                          # ogb.utils.features.get_bond_feature_dims()
BN_EPS = 1e-5                                  # nn.BatchNorm1d default (models/base_layers.py:87)
AGG_EPS = 1e-5                                 # models/pna.py:14

PNA_DEFAULTS = dict(readout_batchnorm=True, readout_hidden_dim=None, readout_layers=2, residual=True,
                    pairwise_distances=False, activation="relu", last_activation="none", mid_batch_norm=False,
                    last_batch_norm=False, propagation_depth=5, dropout=0.0, posttrans_layers=1,
                    pretrans_layers=1, batch_norm_momentum=0.1)          # models/pna.py:95-115
NET3D_DEFAULTS = dict(batch_norm=False, node_wise_output_layers=2, readout_batchnorm=True, batch_norm_momentum=0.1,
                      reduce_func="sum", dropout=0.0, propagation_depth=4, readout_layers=2,
                      readout_hidden_dim=None, fourier_encodings=0, activation="SiLU", update_net_layers=2,
                      message_net_layers=2, use_node_features=False)    # models/net3d.py:15-18

PRETRAIN_QM9_PNA = dict(target_dim=256, hidden_dim=200, mid_batch_norm=True, last_batch_norm=True,
                        readout_batchnorm=True, batch_norm_momentum=0.93, readout_hidden_dim=200, readout_layers=2,
                        dropout=0.0, propagation_depth=7, aggregators=["mean", "max", "min", "std"],
                        scalers=["identity", "amplification", "attenuation"],
                        readout_aggregators=["min", "max", "mean"], pretrans_layers=2, posttrans_layers=1,
                        residual=True)                                   # configs_clean/pre-train_QM9.yml:50-78
PRETRAIN_QM9_NET3D = dict(target_dim=256, hidden_dim=20, hidden_edge_dim=20, node_wise_output_layers=0,
                          message_net_layers=1, update_net_layers=1, reduce_func="mean", fourier_encodings=4,
                          propagation_depth=1, dropout=0.0, batch_norm=True, readout_batchnorm=True,
                          batch_norm_momentum=0.93, readout_hidden_dim=20, readout_layers=1,
                          readout_aggregators=["min", "max", "mean"])    # configs_clean/pre-train_QM9.yml:81-102


# ----------------------------------------------------------------------------------------------
# graph container (what the hot path needs from a batched DGLGraph)
# ----------------------------------------------------------------------------------------------
class OGraph:
    def __init__(self, src, dst, batch_num_nodes, batch_num_edges=None):
        self.src = torch.as_tensor(src).long()
        self.dst = torch.as_tensor(dst).long()
        self.bnn = torch.as_tensor(batch_num_nodes).long()
        self.bne = None if batch_num_edges is None else torch.as_tensor(batch_num_edges).long()
        self.n = int(self.bnn.sum())
        self.deg = torch.bincount(self.dst, minlength=self.n)
        # in-edges of every node, ascending edge id (DGL mailbox order) — bit-exact CSR reference
        self.order = torch.argsort(self.dst, stable=True)
        self.rowptr = torch.zeros(self.n + 1, dtype=torch.long)
        self.rowptr[1:] = torch.cumsum(self.deg, 0)
        self.seg = torch.repeat_interleave(torch.arange(len(self.bnn)), self.bnn)
        self._buckets = None

    def buckets(self):
        if self._buckets is None:
            out = []
            for d in torch.unique(self.deg).tolist():
                if d == 0:
                    continue
                nodes = torch.nonzero(self.deg == d).flatten()
                eids = self.order[self.rowptr[nodes][:, None] + torch.arange(d)[None, :]]
                out.append((d, nodes, eids))
            self._buckets = out
        return self._buckets


# ----------------------------------------------------------------------------------------------
# architecture -> list of FC layers, state initialisation
# ----------------------------------------------------------------------------------------------
def _norm_act(a):
    a = "none" if a is None else str(a).lower()
    if a not in ("relu", "silu", "none", "sigmoid", "tanh", "leakyrelu", "elu", "softplus"):
        raise AssertionError("Unhandled activation function")
    return a


def mlp_spec(in_dim, out_dim, layers, hidden_size=None, mid_activation="relu", last_activation="none",
             mid_batch_norm=False, last_batch_norm=False):
    """[(in, out, act, bn)] exactly as MLP.__init__ stacks FCLayers (models/base_layers.py:119-142)."""
    if layers <= 1:
        return [(in_dim, out_dim, _norm_act(last_activation), bool(last_batch_norm))]
    spec = [(in_dim, hidden_size, _norm_act(mid_activation), bool(mid_batch_norm))]
    for _ in range(layers - 2):
        spec.append((hidden_size, hidden_size, _norm_act(mid_activation), bool(mid_batch_norm)))
    spec.append((hidden_size, out_dim, _norm_act(last_activation), bool(last_batch_norm)))
    return spec


def pna_cfg(**kw):
    c = dict(PNA_DEFAULTS)
    c.update(kw)
    if c["readout_hidden_dim"] is None:
        c["readout_hidden_dim"] = c["hidden_dim"]
    if c["dropout"]:
        raise NotImplementedError("dropout>0 is unused by the target configs")
    if c["pairwise_distances"]:
        raise NotImplementedError("pairwise_distances=True is unused by the target configs")
    return c


def net3d_cfg(**kw):
    c = dict(NET3D_DEFAULTS)
    c.update(kw)
    if c["readout_hidden_dim"] is None:
        c["readout_hidden_dim"] = c["hidden_dim"]
    if c["dropout"]:
        raise NotImplementedError("dropout>0 is unused by the target configs")
    if c["use_node_features"]:
        raise NotImplementedError("use_node_features=True is unused by the target configs")
    if c["reduce_func"] not in ("sum", "mean"):
        raise ValueError("reduce function not supported: ", c["reduce_func"])
    return c


def pna_specs(c):
    """OrderedDict {prefix: mlp spec} for every MLP in PNA (models/pna.py:116-129,186-197)."""
    Fh = c["hidden_dim"]
    n_in = (len(c["aggregators"]) * len(c["scalers"]) + 1) * Fh
    s = OrderedDict()
    for l in range(c["propagation_depth"]):
        s["node_gnn.mp_layers.%d.pretrans" % l] = mlp_spec(3 * Fh, Fh, c["pretrans_layers"], Fh, c["activation"],
                                                           c["last_activation"], c["mid_batch_norm"],
                                                           c["last_batch_norm"])
        s["node_gnn.mp_layers.%d.posttrans" % l] = mlp_spec(n_in, Fh, c["posttrans_layers"], Fh, c["activation"],
                                                            c["last_activation"], c["mid_batch_norm"],
                                                            c["last_batch_norm"])
    s["output"] = mlp_spec(Fh * len(c["readout_aggregators"]), c["target_dim"], c["readout_layers"],
                           c["readout_hidden_dim"], "relu", "none", c["readout_batchnorm"], False)
    return s


def net3d_specs(c):
    H = c["hidden_dim"]
    act, bn = c["activation"], c["batch_norm"]
    e_in = 1 if c["fourier_encodings"] == 0 else 2 * c["fourier_encodings"] + 1
    s = OrderedDict()
    s["edge_input"] = mlp_spec(e_in, H, 1, H, act, act, bn, bn)                                    # net3d.py:22-25
    for l in range(c["propagation_depth"]):
        s["mp_layers.%d.message_network" % l] = mlp_spec(3 * H, H, c["message_net_layers"], H, act, act, bn, bn)
        s["mp_layers.%d.update_network" % l] = mlp_spec(H, H, c["update_net_layers"], H, act, "none", bn, bn)
    if c["node_wise_output_layers"] > 0:
        s["node_wise_output_network"] = mlp_spec(H, H, c["node_wise_output_layers"], H, act, "none", bn, bn)
    s["output"] = mlp_spec(H * len(c["readout_aggregators"]), c["target_dim"], c["readout_layers"],
                           c["readout_hidden_dim"], "relu", "none", c["readout_batchnorm"], False)
    return s


def _xavier_uniform(shape, gain, gen):
    fan_out, fan_in = shape[0], shape[1]
    a = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2 - 1) * a


def _init_mlp(state, prefix, spec, gen):
    for i, (din, dout, _act, bn) in enumerate(spec):
        p = "%s.fully_connected.%d" % (prefix, i)
        state[p + ".linear.weight"] = _xavier_uniform((dout, din), 1.0 / din, gen)      # base_layers.py:96
        state[p + ".linear.bias"] = torch.zeros(dout)                                    # base_layers.py:98
        if bn:
            state[p + ".batch_norm.weight"] = torch.ones(dout)
            state[p + ".batch_norm.bias"] = torch.zeros(dout)
            state[p + ".batch_norm.running_mean"] = torch.zeros(dout)
            state[p + ".batch_norm.running_var"] = torch.ones(dout)
            state[p + ".batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def init_pna_state(c, seed, trained_scale=False):
    """Fresh state with the reference's initialisers; ``trained_scale`` rescales to the shipped
    checkpoint's per-class statistics so activations are not degenerate."""
    gen = torch.Generator().manual_seed(seed)
    st = OrderedDict()
    Fh = c["hidden_dim"]
    specs = pna_specs(c)
    for l in range(c["propagation_depth"]):
        for k in ("pretrans", "posttrans"):
            _init_mlp(st, "node_gnn.mp_layers.%d.%s" % (l, k), specs["node_gnn.mp_layers.%d.%s" % (l, k)], gen)
    for i, d in enumerate(ATOM_DIMS):
        st["node_gnn.atom_encoder.atom_embedding_list.%d.weight" % i] = _xavier_uniform((d, Fh), 1.0, gen)
    for i, d in enumerate(BOND_DIMS):
        st["node_gnn.bond_encoder.bond_embedding_list.%d.weight" % i] = _xavier_uniform((d, Fh), 1.0, gen)
    _init_mlp(st, "output", specs["output"], gen)
    if trained_scale:
        _to_trained_scale(st, gen, net3d=False)
    return st


def init_net3d_state(c, seed, trained_scale=False):
    gen = torch.Generator().manual_seed(seed)
    st = OrderedDict()
    specs = net3d_specs(c)
    st["node_embedding"] = torch.randn(c["hidden_dim"], generator=gen)                   # net3d.py:31-32
    _init_mlp(st, "edge_input", specs["edge_input"], gen)
    for l in range(c["propagation_depth"]):
        _init_mlp(st, "mp_layers.%d.message_network" % l, specs["mp_layers.%d.message_network" % l], gen)
        _init_mlp(st, "mp_layers.%d.update_network" % l, specs["mp_layers.%d.update_network" % l], gen)
        H = c["hidden_dim"]
        k = 1.0 / math.sqrt(H)                                                           # nn.Linear default init
        st["mp_layers.%d.soft_edge_network.weight" % l] = (torch.rand((1, H), generator=gen) * 2 - 1) * k
        st["mp_layers.%d.soft_edge_network.bias" % l] = (torch.rand((1,), generator=gen) * 2 - 1) * k
    if "node_wise_output_network" in specs:
        _init_mlp(st, "node_wise_output_network", specs["node_wise_output_network"], gen)
    _init_mlp(st, "output", specs["output"], gen)
    if trained_scale:
        _to_trained_scale(st, gen, net3d=True)
    return st


def _to_trained_scale(st, gen, net3d):
    """Seeded weights whose per-class std matches the shipped checkpoint."""
    for k in list(st.keys()):
        t = st[k]
        if k.endswith("linear.weight") or k.endswith("soft_edge_network.weight"):
            std = 0.03 if net3d else 0.012
            if net3d and "soft_edge" in k:
                std = 0.7
            if net3d and k.startswith("output"):
                std = 0.008
            st[k] = torch.randn(t.shape, generator=gen) * std
        elif k.endswith("linear.bias") or k.endswith("soft_edge_network.bias"):
            st[k] = torch.randn(t.shape, generator=gen) * (0.05 if net3d else 0.01)
        elif k.endswith("batch_norm.weight"):
            st[k] = 1.0 + 0.03 * torch.randn(t.shape, generator=gen)
        elif k.endswith("batch_norm.bias"):
            st[k] = 0.05 * torch.randn(t.shape, generator=gen)
        elif k.endswith("running_mean"):
            st[k] = 0.1 * torch.randn(t.shape, generator=gen)
        elif k.endswith("running_var"):
            st[k] = 0.05 + torch.rand(t.shape, generator=gen)
        elif "embedding_list" in k:
            st[k] = torch.randn(t.shape, generator=gen) * 0.09
    return st


def param_keys(st):
    return [k for k in st if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def as_leaf_params(st):
    """Clone; float tensors that are parameters get requires_grad."""
    out = OrderedDict()
    pk = set(param_keys(st))
    for k, v in st.items():
        t = v.detach().clone()
        if k in pk:
            t.requires_grad_(True)
        out[k] = t
    return out


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------
def _act(x, a):
    if a == "none":
        return x
    if a == "relu":
        return torch.relu(x)
    if a == "silu":
        return F.silu(x)
    if a == "sigmoid":
        return torch.sigmoid(x)
    if a == "tanh":
        return torch.tanh(x)
    raise NotImplementedError(a)


def fc_forward(x, st, p, act, bn, training, momentum):
    """FCLayer.forward (models/base_layers.py:100-111): Linear -> activation -> (dropout) -> BatchNorm."""
    h = F.linear(x, st[p + ".linear.weight"], st[p + ".linear.bias"])
    h = _act(h, act)
    if bn:
        h = F.batch_norm(h, st[p + ".batch_norm.running_mean"], st[p + ".batch_norm.running_var"],
                         st[p + ".batch_norm.weight"], st[p + ".batch_norm.bias"], training, momentum, BN_EPS)
        if training:
            st[p + ".batch_norm.num_batches_tracked"] += 1
    return h


def mlp_forward(x, st, prefix, spec, training, momentum):
    for i, (_din, _dout, act, bn) in enumerate(spec):                    # base_layers.py:144-147
        x = fc_forward(x, st, "%s.fully_connected.%d" % (prefix, i), act, bn, training, momentum)
    return x


def embed_sum(idx, st, fmt, n_cols):
    out = 0
    for c in range(n_cols):                                              # mol_encoder.py:34-42, 65-73
        out = out + F.embedding(idx[:, c], st[fmt % c])
    return out


def _aggregate(mail, name):
    """Aggregators over the mailbox axis dim=-2 (models/pna.py:17-37)."""
    if name == "mean":
        return mail.mean(dim=-2)
    if name == "max":
        return mail.max(dim=-2)[0]
    if name == "min":
        return mail.min(dim=-2)[0]
    if name == "sum":
        return mail.sum(dim=-2)
    if name in ("std", "var"):
        mean = mail.mean(dim=-2)
        var = torch.relu((mail * mail).mean(dim=-2) - mean * mean)
        return var if name == "var" else torch.sqrt(var + AGG_EPS)
    raise NotImplementedError("aggregator %s is unused by the target configs" % name)


def _scale(h, name, D):
    """Degree scalers with avg_d['log'] hard-coded to 1.0 (models/pna.py:57-68,153)."""
    if name == "identity":
        return h
    if name == "amplification":
        return h * (np.log(D + 1) / 1.0)
    if name == "attenuation":
        return h * (1.0 / np.log(D + 1))
    raise NotImplementedError(name)


def pna_reduce(g, msg, aggregators, scalers):
    """DGL update_all with a UDF reduce = degree bucketing (models/pna.py:206,221-235)."""
    width = len(aggregators) * (len(scalers) if len(scalers) > 1 else 1) * msg.shape[1]
    out = torch.zeros(g.n, width, dtype=msg.dtype)
    for D, nodes, eids in g.buckets():
        mail = msg[eids.reshape(-1)].reshape(len(nodes), D, -1)
        h = torch.cat([_aggregate(mail, a) for a in aggregators], dim=-1)
        if len(scalers) > 1:
            h = torch.cat([_scale(h, s, D) for s in scalers], dim=-1)
        out = out.index_put((nodes,), h)
    return out


def segment_readout(x, g, op):
    B = len(g.bnn)
    if op in ("sum", "mean"):
        out = torch.zeros(B, x.shape[1], dtype=x.dtype).index_add(0, g.seg, x)
        return out / g.bnn.to(x.dtype)[:, None] if op == "mean" else out
    # DGL's segment_reduce max/min keeps the arg index (first row attaining the extremum, strict compare) and its
    # backward sends the whole gradient there -- not an even split between ties
    return x.gather(0, _first_extremum_rows(x.detach(), g.seg, B, op))


def _first_extremum_rows(x, seg, B, op):
    red = {"max": "amax", "min": "amin"}[op]
    idx = seg[:, None].expand_as(x)
    val = torch.zeros(B, x.shape[1], dtype=x.dtype).scatter_reduce(0, idx, x, red, include_self=False)
    n = x.shape[0]
    pos = torch.arange(n)[:, None].expand_as(x)
    cand = torch.where(x == val[seg], pos, torch.full_like(pos, n))
    return torch.zeros(B, x.shape[1], dtype=torch.long).scatter_reduce(0, idx, cand, "amin", include_self=False)


def pna_forward(st, c, g, x_atom, e_attr, training=True, taps=None):
    """PNA.forward (models/pna.py:131-135) -> [B, target_dim].  ``taps`` (dict) collects intermediates."""
    mom = c["batch_norm_momentum"]
    specs = pna_specs(c)
    h = embed_sum(x_atom, st, "node_gnn.atom_encoder.atom_embedding_list.%d.weight", x_atom.shape[1])
    ef = embed_sum(e_attr, st, "node_gnn.bond_encoder.bond_embedding_list.%d.weight", e_attr.shape[1])
    if taps is not None:
        taps["h0"], taps["ef"] = h, ef
    for l in range(c["propagation_depth"]):
        pre = "node_gnn.mp_layers.%d" % l
        z = torch.cat([h[g.src], h[g.dst], ef], dim=-1)                                  # pna.py:248-249
        e = mlp_forward(z, st, pre + ".pretrans", specs[pre + ".pretrans"], training, mom)  # pna.py:252
        agg = pna_reduce(g, e, c["aggregators"], c["scalers"])                           # pna.py:206
        hn = mlp_forward(torch.cat([h, agg], dim=-1), st, pre + ".posttrans", specs[pre + ".posttrans"],
                         training, mom)                                                  # pna.py:207-209
        if taps is not None:
            taps["msg%d" % l], taps["agg%d" % l] = e, agg
        h = hn + h if c["residual"] else hn                                              # pna.py:210-211
        if taps is not None:
            taps["h%d" % (l + 1)] = h
    ro = torch.cat([segment_readout(h, g, op) for op in c["readout_aggregators"]], dim=-1)  # pna.py:133-134
    if taps is not None:
        taps["readout"] = ro
    return mlp_forward(ro, st, "output", specs["output"], training, mom)                # pna.py:135


def fourier_encode(d, num_encodings):
    """commons/utils.py:103-110 for d [E,1] -> [E, 2k+1] = [sin(d/2^i)..., cos(d/2^i)..., d]."""
    scales = 2 ** torch.arange(num_encodings, dtype=d.dtype)
    x = d.reshape(-1, 1) / scales[None, :]
    return torch.cat([x.sin(), x.cos(), d.reshape(-1, 1)], dim=-1)


def net3d_forward(st, c, g, d, training=True, taps=None):
    """Net3D.forward (models/net3d.py:57-75) -> [B*C, target_dim]."""
    mom = c["batch_norm_momentum"]
    specs = net3d_specs(c)
    H = c["hidden_dim"]
    h = st["node_embedding"][None, :].expand(g.n, H)                                     # net3d.py:61
    e = fourier_encode(d, c["fourier_encodings"]) if c["fourier_encodings"] > 0 else d   # net3d.py:63-64
    e = F.silu(mlp_forward(e, st, "edge_input", specs["edge_input"], training, mom))     # net3d.py:80-81
    if taps is not None:
        taps["d0"] = e
    deg = g.deg.clamp(min=1).to(e.dtype)[:, None]
    for l in range(c["propagation_depth"]):
        pre = "mp_layers.%d" % l
        msg = mlp_forward(torch.cat([h[g.src], h[g.dst], e], dim=-1), st, pre + ".message_network",
                          specs[pre + ".message_network"], training, mom)               # net3d.py:112-115
        e = e + msg                                                                      # net3d.py:116
        w = torch.sigmoid(F.linear(msg, st[pre + ".soft_edge_network.weight"],
                                   st[pre + ".soft_edge_network.bias"]))                 # net3d.py:117
        m = torch.zeros(g.n, H, dtype=msg.dtype).index_add(0, g.dst, msg * w)           # fn.sum / fn.mean
        if c["reduce_func"] == "mean":
            m = m / deg
        hn = mlp_forward(m + h, st, pre + ".update_network", specs[pre + ".update_network"], training, mom)
        if taps is not None:
            taps["msg%d" % l], taps["m%d" % l] = msg, m
        h = hn + h                                                                       # net3d.py:120-125
    if c["node_wise_output_layers"] > 0:
        h = mlp_forward(h, st, "node_wise_output_network", specs["node_wise_output_network"], training, mom)
    ro = torch.cat([segment_readout(h, g, op) for op in c["readout_aggregators"]], dim=-1)
    if taps is not None:
        taps["readout"] = ro
    return mlp_forward(ro, st, "output", specs["output"], training, mom)


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def ntxent(z1, z2, tau=0.5, norm=True):
    """NTXent.forward (commons/losses.py:143-155); regularisers (weights 0 in target configs) omitted."""
    sim = z1 @ z2.t()
    if norm:
        sim = sim / (z1.norm(dim=1)[:, None] * z2.norm(dim=1)[None, :] + 1e-8)
    sim = torch.exp(sim / tau)
    pos = torch.diagonal(sim)
    return -torch.log(pos / (sim.sum(dim=1) - pos)).mean()


def ntxent_multiple_positives(z1, z2, tau=0.5, norm=True):
    """NTXentMultiplePositives.forward (commons/losses.py:225-246): z2 is [B*C, dim], molecule-major."""
    B, dim = z1.shape
    z2 = z2.view(B, -1, dim)
    sim = torch.einsum("ik,juk->iju", z1, z2)
    if norm:
        sim = sim / (z1.norm(dim=1)[:, None, None] * z2.norm(dim=2)[None, :, :])
    sim = torch.exp(sim / tau).sum(dim=2)
    pos = torch.diagonal(sim)
    return -torch.log(pos / (sim.sum(dim=1) - pos)).mean()


LOSSES = {"NTXent": ntxent, "NTXentMultiplePositives": ntxent_multiple_positives}


# ----------------------------------------------------------------------------------------------
# one optimisation step (trainer/self_supervised_trainer.py:24-29,78-86; trainer/trainer.py:116-124)
# ----------------------------------------------------------------------------------------------
class OracleTrainer:
    """PNA + Net3D + loss + backward + Adam over both nets (BN params in a weight_decay=0 group)."""

    def __init__(self, c2d, c3d, st2d, st3d, loss="NTXent", tau=0.1, lr=8e-5, betas=(0.9, 0.999), eps=1e-8,
                 net3d_training=True):
        self.c2d, self.c3d = c2d, c3d
        self.st2d, self.st3d = as_leaf_params(st2d), as_leaf_params(st3d)
        self.loss_fn, self.tau = LOSSES[loss], tau
        named = [(k, self.st2d[k]) for k in param_keys(self.st2d)] + [(k, self.st3d[k]) for k in param_keys(self.st3d)]
        bn = [v for k, v in named if "batch_norm" in k]
        rest = [v for k, v in named if "batch_norm" not in k]
        self.optim = torch.optim.Adam([{"params": bn, "weight_decay": 0}, {"params": rest}], lr=lr, betas=betas,
                                      eps=eps)
        self.net3d_training = net3d_training    # the reference never calls model3d.eval() (trainer.py:72,75)

    def forward(self, g2, x_atom, e_attr, g3, d3, training=True):
        z2 = pna_forward(self.st2d, self.c2d, g2, x_atom, e_attr, training)
        z3 = net3d_forward(self.st3d, self.c3d, g3, d3, self.net3d_training)
        return self.loss_fn(z2, z3, tau=self.tau), z2, z3

    def step(self, g2, x_atom, e_attr, g3, d3):
        loss, z2, z3 = self.forward(g2, x_atom, e_attr, g3, d3, True)
        loss.backward()
        self.optim.step()
        self.optim.zero_grad()
        return loss.detach(), z2.detach(), z3.detach()


def graphs_from_batch(b):
    """numpy batch dict (3dinfomax_b200/synthetic.py layout) -> (g2, x_atom, e_attr, g3, d3)."""
    g2 = OGraph(b["src"], b["dst"], b["num_nodes"], b["num_edges"])
    g3 = OGraph(b["src3"], b["dst3"], b["num_nodes3"], b["num_edges3"])
    return (g2, torch.as_tensor(b["x_atom"]).long(), torch.as_tensor(b["e_attr"]).long(), g3,
            torch.as_tensor(b["d3"]).float())


# ----------------------------------------------------------------------------------------------
# bit-exact integer reference for the CSR builder (numpy; argsort(dst, stable))
# ----------------------------------------------------------------------------------------------
def csr_reference(src, dst, n):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    eid = np.argsort(dst, kind="stable").astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=n), out=rowptr[1:])
    return rowptr, src[eid].astype(np.int32), eid


# ---------------------------------------------------------------------------------------------------------------
# Contrastive metrics (trainer/metrics.py:240-334, 444-463; pos_mask == None).  Pinned on the reference's own classes
# by oracle/pin_metrics.py (bit-equal on the committed vectors).
# ---------------------------------------------------------------------------------------------------------------
def contrastive_metrics(x1, x2, threshold=0.5):
    """(positive_similarity, negative_similarity, true_positive_rate, true_negative_rate, contrastive_accuracy)"""
    B = x1.shape[0]
    if x1.shape != x2.shape:
        x2 = x2[:B]                                                              # metrics.py:243-244
    sim = torch.einsum("ik,jk->ij", x1, x2)
    sim = sim / torch.einsum("i,j->ij", x1.norm(dim=1), x2.norm(dim=1))          # metrics.py:246-250
    pos = torch.nn.functional.cosine_similarity(x1, x2)                         # metrics.py:331 (global vs global)
    positive_similarity = ((pos + 1) / 2).mean(dim=0)
    diag = sim[range(B), range(B)]
    negative_similarity = (((sim.sum(dim=1) - diag) / (B - 1) + 1) / 2).mean(dim=0)   # metrics.py:459-462
    preds = (sim + 1) / 2 > threshold
    pos_mask = torch.eye(B)
    neg_mask = 1 - pos_mask
    tp = B - ((preds.long() - pos_mask) * pos_mask).count_nonzero()             # metrics.py:256-257
    tn = B * (B - 1) - (((~preds).long() - neg_mask) * neg_mask).count_nonzero()  # metrics.py:279-280
    tpr, tnr = tp / B, tn / (B * (B - 1))
    return torch.stack([positive_similarity, negative_similarity, tpr, tnr, (tpr + tnr) / 2])


# ---------------------------------------------------------------------------------------------------------------
# The other four metrics the pre-training configs log (configs_clean/pre-train_QM9.yml: uniformity, alignment,
# batch_variance, dimension_covariance): trainer/metrics.py:161-176, 212-230 with commons/losses.py:946-964.
# Pinned on the reference's own classes by oracle/pin_metrics.py.
# ---------------------------------------------------------------------------------------------------------------
def _cov_loss(x):
    """commons/losses.py:954-959"""
    B, D = x.shape
    x = x - x.mean(dim=0)
    cov = (x.T @ x) / (B - 1)
    off = cov.flatten()[:-1].view(D - 1, D + 1)[:, 1:].flatten()
    return off.pow(2).sum() / D


def _uniformity(x, t=2):
    """one term of commons/losses.py:946-951"""
    return torch.pdist(x, p=2).pow(2).mul(-t).exp().mean().log()


def embedding_metrics(x1, x2, alpha=2):
    """(dimension_covariance, batch_variance, alignment, uniformity) of embeddings x1 [B,D], x2 [>=B,D]:
    DimensionCovariance = cov_loss(x1) + cov_loss(x2)                       (metrics.py:161-166)
    BatchVariance       = x1.std(0).mean() + x2.std(0).mean()               (metrics.py:169-174)
    Alignment(alpha)    = ||x1 - x2[:B]||_2^alpha averaged over the batch   (metrics.py:212-220)
    Uniformity          = uniformity_loss(x1, x2) with its default t = 2    (metrics.py:223-229: self.t is not passed on)"""
    dc = _cov_loss(x1) + _cov_loss(x2)
    bv = x1.std(dim=0).mean() + x2.std(dim=0).mean()
    al = (x1 - x2[:len(x1)]).norm(dim=1).pow(alpha).mean()
    un = (_uniformity(x1) + _uniformity(x2)) / 2
    return torch.stack([dc, bv, al, un])
