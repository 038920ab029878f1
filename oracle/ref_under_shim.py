"""TEST INFRASTRUCTURE — container-only ground truth.  Never imported by the product.

Runs the reference's OWN hot-path modules (models/pna.py, models/net3d.py,
models/base_layers.py, commons/mol_encoder.py, commons/losses.py), unmodified, from
/root/reference, under a minimal pure-torch stand-in for the slice of DGL they touch
(SURVEY.md §8c, Appendix B).  DGL / ogb / rdkit are not installable here (no network).

It exists to (i) pin oracle/oracle.py against the real reference code and (ii) emit the
committed golden vectors under tests/golden/ (oracle/make_golden.py).  /root/reference
does not exist on the GPU box: nothing that runs there may import this file.

DGL semantics implemented (documented DGL behaviour):
  * apply_edges(udf): ONE udf call over all edges in edge-id order.
  * update_all(udf_msg, udf_reduce): nodes bucketed by in-degree; the mailbox of a bucket
    is [N_D, D, F] with the D axis in ascending edge id; zero in-degree rows are zero.
  * update_all(udf_msg, fn.sum|fn.mean, apply_udf): segment sum / mean over in-edges.
  * readout_nodes(g, feat, op): per-graph segment reduce.
"""
import importlib.util
import os
import sys
import types

import torch

REF = os.environ.get("I3D_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "pna.py"))


class _View:
    def __init__(self, src=None, dst=None, data=None, mailbox=None):
        self.src, self.dst, self.data, self.mailbox = src, dst, data, mailbox


class ShimGraph:
    """Stands in for a *batched* DGLGraph (only the API the hot path touches)."""

    def __init__(self, src, dst, num_nodes, batch_num_nodes, batch_num_edges):
        self.src = src.long()
        self.dst = dst.long()
        self.n = int(num_nodes)
        self._bnn = batch_num_nodes.long()
        self._bne = batch_num_edges.long()
        self.ndata = {}
        self.edata = {}

    def number_of_nodes(self):
        return self.n

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    def edges(self):
        return self.src, self.dst

    def apply_edges(self, f):
        v = _View(src={k: t[self.src] for k, t in self.ndata.items()},
                  dst={k: t[self.dst] for k, t in self.ndata.items()}, data=self.edata)
        self.edata.update(f(v))

    def apply_nodes(self, f):
        self.ndata.update(f(_View(data=self.ndata)))

    def update_all(self, message_func, reduce_func, apply_node_func=None):
        v = _View(src={k: t[self.src] for k, t in self.ndata.items()},
                  dst={k: t[self.dst] for k, t in self.ndata.items()}, data=self.edata)
        msgs = message_func(v)
        deg = torch.bincount(self.dst, minlength=self.n)
        if isinstance(reduce_func, tuple):
            op, mk, ok = reduce_func
            m = msgs[mk]
            acc = torch.zeros((self.n,) + m.shape[1:], dtype=m.dtype).index_add_(0, self.dst, m)
            if op == "mean":
                acc = acc / deg.clamp(min=1).to(m.dtype).view(-1, *([1] * (m.dim() - 1)))
            self.ndata[ok] = acc
        else:
            order = torch.argsort(self.dst, stable=True)
            starts = torch.cumsum(deg, 0) - deg
            res = None
            for d in torch.unique(deg).tolist():
                if d == 0:
                    continue
                nodes = torch.nonzero(deg == d).flatten()
                eids = order[(starts[nodes][:, None] + torch.arange(d)[None, :])].reshape(-1)
                mail = {k: t[eids].reshape(len(nodes), d, *t.shape[1:]) for k, t in msgs.items()}
                out = reduce_func(_View(data={k: t[nodes] for k, t in self.ndata.items()}, mailbox=mail))
                if res is None:
                    res = {k: torch.zeros((self.n,) + t.shape[1:], dtype=t.dtype) for k, t in out.items()}
                for k, t in out.items():
                    res[k] = res[k].index_put((nodes,), t)
            if res is not None:
                self.ndata.update(res)
        if apply_node_func is not None:
            self.apply_nodes(apply_node_func)


def _readout_nodes(g, feat, op="sum"):
    x = g.ndata[feat]
    B = len(g._bnn)
    seg = torch.repeat_interleave(torch.arange(B), g._bnn)
    if op in ("sum", "mean"):
        out = torch.zeros((B,) + x.shape[1:], dtype=x.dtype).index_add_(0, seg, x)
        if op == "mean":
            out = out / g._bnn.to(x.dtype).view(-1, *([1] * (x.dim() - 1)))
        return out
    # DGL segment_reduce max/min: value of the FIRST row attaining the extremum; backward routes the gradient there
    red = {"max": "amax", "min": "amin"}[op]
    xd = x.detach()
    idx = seg.view(-1, *([1] * (x.dim() - 1))).expand_as(x)
    val = torch.zeros((B,) + x.shape[1:], dtype=x.dtype).scatter_reduce(0, idx, xd, red, include_self=False)
    n = x.shape[0]
    pos = torch.arange(n).view(-1, *([1] * (x.dim() - 1))).expand_as(x)
    cand = torch.where(xd == val[seg], pos, torch.full_like(pos, n))
    first = torch.zeros((B,) + x.shape[1:], dtype=torch.long).scatter_reduce(0, idx, cand, "amin", include_self=False)
    return x.gather(0, first)


_LOADED = {}


def load_reference():
    """Returns a namespace with the reference's PNA, Net3D, NTXent, NTXentMultiplePositives."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF)
    saved = {k: sys.modules.get(k) for k in
             ("dgl", "dgl.function", "ogb", "ogb.utils", "ogb.utils.features", "models", "commons",
              "commons.utils", "commons.mol_encoder", "models.base_layers", "models.pna", "models.net3d",
              "commons.losses")}
    dgl = types.ModuleType("dgl")
    dgl.DGLGraph = ShimGraph
    dgl.readout_nodes = _readout_nodes
    fn = types.ModuleType("dgl.function")
    fn.sum = lambda msg, out: ("sum", msg, out)
    fn.mean = lambda msg, out: ("mean", msg, out)
    dgl.function = fn
    ogb = types.ModuleType("ogb")
    ogbu = types.ModuleType("ogb.utils")
    ogbf = types.ModuleType("ogb.utils.features")
    ogbf.get_atom_feature_dims = lambda: [119, 4, 12, 12, 10, 6, 6, 2, 2]
    ogbf.get_bond_feature_dims = lambda: [5, 6, 2]
    sys.modules.update({"dgl": dgl, "dgl.function": fn, "ogb": ogb, "ogb.utils": ogbu, "ogb.utils.features": ogbf})
    pk_models = types.ModuleType("models")
    pk_models.__path__ = []
    pk_commons = types.ModuleType("commons")
    pk_commons.__path__ = []
    sys.modules["models"] = pk_models
    sys.modules["commons"] = pk_commons
    # commons/utils.py imports half the world; only fourier_encode_dist (lines 103-110) is needed.
    cu = types.ModuleType("commons.utils")
    with open(os.path.join(REF, "commons", "utils.py")) as fh:
        lines = fh.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def fourier_encode_dist"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("def "))
    exec(compile("import torch\n" + "\n".join(lines[start:end]), "commons/utils.py[fourier]", "exec"), cu.__dict__)
    sys.modules["commons.utils"] = cu
    mods = {}
    for name, rel in [("commons.mol_encoder", "commons/mol_encoder.py"),
                      ("models.base_layers", "models/base_layers.py"),
                      ("models.pna", "models/pna.py"),
                      ("models.net3d", "models/net3d.py"),
                      ("commons.losses", "commons/losses.py")]:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    _LOADED.update(PNA=mods["models.pna"].PNA, Net3D=mods["models.net3d"].Net3D,
                   NTXent=mods["commons.losses"].NTXent,
                   NTXentMultiplePositives=mods["commons.losses"].NTXentMultiplePositives,
                   fourier_encode_dist=cu.fourier_encode_dist, ShimGraph=ShimGraph)
    # do not leave stub packages behind for unrelated imports
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return types.SimpleNamespace(**_LOADED)


def checkpoint_paths():
    run = os.path.join(REF, "runs", "PNA_qmugs_NTXentMultiplePositives_620000_123_25-08_09-19-52")
    return os.path.join(run, "best_checkpoint_35epochs.pt"), os.path.join(run, "train_arguments.yaml")
