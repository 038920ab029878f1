"""TEST INFRASTRUCTURE — container-only: pins oracle/collate_oracle.py on the reference's own batch construction.

Loads datasets/qm9_dataset.py and datasets/custom_collate.py UNMODIFIED from /root/reference (by file path, with stub
modules for the imports that are not installable here: dgl, rdkit, ogb, torch_geometric), fills a ``QM9Dataset``
object with a seeded synthetic packed store (bypassing ``__init__``, which needs rdkit to build the store), runs the
reference's ``__getitem__`` -> ``get_graph`` / ``get_complete_graph`` and ``contrastive_collate`` and asserts that
``collate_reference`` returns the same batch: integer arrays bit-exact, distances bit-exact (both are fp32
sqrt(sum of squares)).  Writes tests/golden/collate_qm9.npz (inputs' seed + outputs) for the GPU-box tests.

DGL calls stubbed (documented DGL semantics): ``dgl.graph((src, dst), num_nodes=n)`` — edge ids in the given order;
``g.ndata / g.edata / g.edges()``; ``dgl.batch(graphs)`` — node ids offset by the cumulative node counts, order kept.

    python -m oracle.pin_collate
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("I3D_REFERENCE_ROOT", "/root/reference")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _G:
    def __init__(self, src, dst, num_nodes=None, bnn=None, bne=None):
        self.src, self.dst = src.long(), dst.long()
        self.n = int(num_nodes) if num_nodes is not None else (int(max(src.max(), dst.max())) + 1 if len(src) else 0)
        self.ndata, self.edata = {}, {}
        self._bnn = bnn if bnn is not None else torch.tensor([self.n])
        self._bne = bne if bne is not None else torch.tensor([len(self.src)])

    def edges(self):
        return self.src, self.dst

    def number_of_nodes(self):
        return self.n

    def to(self, device):
        return self

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne


def _dgl_graph(edges, num_nodes=None, device=None):
    return _G(edges[0], edges[1], num_nodes)


def _dgl_batch(graphs):
    off, src, dst = 0, [], []
    for g in graphs:
        src.append(g.src + off)
        dst.append(g.dst + off)
        off += g.n
    out = _G(torch.cat(src), torch.cat(dst), off, torch.cat([g._bnn for g in graphs]),
             torch.cat([g._bne for g in graphs]))
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs])
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs])
    return out


def load_reference_collate():
    stubs = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        stubs[name] = m
        return m

    mod("dgl", graph=_dgl_graph, batch=_dgl_batch, DGLGraph=_G, heterograph=None)
    mod("torch_geometric")
    mod("ogb"), mod("ogb.utils")
    mod("ogb.utils.features", atom_to_feature_vector=None, bond_to_feature_vector=None,
        get_atom_feature_dims=lambda: [119, 4, 12, 12, 10, 6, 6, 2, 2], get_bond_feature_dims=lambda: [5, 6, 2])
    mod("rdkit", Chem=None), mod("rdkit.Chem"), mod("rdkit.Chem.rdmolops", GetAdjacencyMatrix=None)
    pk = mod("commons")
    pk.__path__ = []
    mod("commons.spherical_encoding", dist_emb=None)
    mod("commons.utils", get_adj_matrix=None)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        out = {}
        for name, rel in [("_ref_qm9_dataset", "datasets/qm9_dataset.py"), ("_ref_custom_collate", "datasets/custom_collate.py"),
                          ("_ref_qmugs_dataset", "datasets/qmugs_dataset.py")]:
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            out[name] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    load_reference_collate.qmugs = out["_ref_qmugs_dataset"].QMugsDataset
    return out["_ref_qm9_dataset"].QM9Dataset, out["_ref_custom_collate"].contrastive_collate


def reference_batch_conformers(store, idx, conformers):
    """The reference's QMugsDataset.__getitem__ with return_types ['dgl_graph', 'conformations']
    (configs_clean/pre-train_QMugs.yml; datasets/qmugs_dataset.py:76-96,149-166) + contrastive_collate."""
    _, contrastive_collate = load_reference_collate()
    QMugsDataset = load_reference_collate.qmugs
    ds = object.__new__(QMugsDataset)
    t = torch.from_numpy
    ds.device = "cpu"
    ds.num_conformers = int(conformers)
    ds.return_types = ["dgl_graph", "conformations"]
    ds.features_tensor = t(store["atom_features"])
    ds.e_features_tensor = t(store["edge_features"])
    ds.conformations = t(store["conformations"]).float()
    ds.coordinates = ds.conformations[:, :3].float()                      # qmugs_dataset.py:44
    ds.edge_indices = t(store["edge_indices"])
    ds.meta_dict = {"chembl_ids": np.arange(len(store["n_atoms"])), "edge_slices": t(store["edge_slices"]),
                    "atom_slices": t(store["atom_slices"]), "n_atoms": t(store["n_atoms"])}
    ds.dgl_graphs, ds.pairwise, ds.complete_graphs, ds.conformer_graphs = {}, {}, {}, {}
    (g2,), (g3,) = contrastive_collate([ds[int(i)] for i in idx])
    return {"src": g2.src.numpy(), "dst": g2.dst.numpy(), "x_atom": g2.ndata["feat"].numpy(),
            "e_attr": g2.edata["feat"].numpy(), "num_nodes": g2.batch_num_nodes().numpy(),
            "num_edges": g2.batch_num_edges().numpy(), "src3": g3.src.numpy(), "dst3": g3.dst.numpy(),
            "d3": g3.edata["d"].numpy(), "num_nodes3": g3.batch_num_nodes().numpy(),
            "num_edges3": g3.batch_num_edges().numpy()}


def reference_batch(store, idx):
    QM9Dataset, contrastive_collate = load_reference_collate()
    ds = object.__new__(QM9Dataset)               # __init__ builds the store with rdkit; the store is given here
    t = torch.from_numpy
    ds.device = "cpu"
    ds.return_types = ["dgl_graph", "complete_graph3d"]      # configs_clean/pre-train_QM9.yml:12-14
    ds.features_tensor = t(store["atom_features"])
    ds.e_features_tensor = t(store["edge_features"])
    ds.coordinates = t(store["coordinates"])
    ds.edge_indices = t(store["edge_indices"])
    ds.meta_dict = {"mol_id": np.arange(len(store["n_atoms"])), "edge_slices": t(store["edge_slices"]),
                    "atom_slices": t(store["atom_slices"]), "n_atoms": t(store["n_atoms"])}
    ds.dgl_graphs, ds.pairwise, ds.complete_graphs = {}, {}, {}
    (g2,), (g3,) = contrastive_collate([ds[int(i)] for i in idx])
    return {"src": g2.src.numpy(), "dst": g2.dst.numpy(), "x_atom": g2.ndata["feat"].numpy(),
            "e_attr": g2.edata["feat"].numpy(), "num_nodes": g2.batch_num_nodes().numpy(),
            "num_edges": g2.batch_num_edges().numpy(), "src3": g3.src.numpy(), "dst3": g3.dst.numpy(),
            "d3": g3.edata["d"].numpy(), "num_nodes3": g3.batch_num_nodes().numpy(),
            "num_edges3": g3.batch_num_edges().numpy()}


CASES = {"collate_qm9": (77, 40, "qm9", [3, 17, 0, 39, 17, 8, 21]),          # repeated index on purpose
         "collate_qmugs": (78, 12, "qmugs", [11, 2, 5])}


CONFORMER_CASES = {"collate_qmugs_c3": (79, 10, "qmugs", [7, 1, 4, 1], 3)}


def main():
    from oracle.collate_oracle import collate_reference, make_store
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (seed, M, shape, idx) in CASES.items():
        store = make_store(seed, M, shape)
        ref = reference_batch(store, idx)
        mine = collate_reference(store, idx)
        for k, v in ref.items():
            assert mine[k].dtype == v.dtype and mine[k].shape == v.shape, (name, k, mine[k].dtype, v.dtype, mine[k].shape, v.shape)
            assert np.array_equal(mine[k], v), "oracle != reference for %s/%s" % (name, k)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), seed=seed, n_molecules=M,
                            idx=np.array(idx), **ref)
        print("pinned %s: %d molecules, N=%d E=%d E3=%d — oracle == reference (bit exact)"
              % (name, len(idx), len(ref["x_atom"]), len(ref["src"]), len(ref["src3"])))
    # multi-conformer batches (BASELINE configs 3-4: NTXentMultiplePositives over 3 conformers per molecule)
    import importlib
    from oracle.collate_oracle import collate_reference_conformers
    syn = importlib.import_module("3dinfomax_b200.synthetic")
    for name, (seed, M, shape, idx, C) in CONFORMER_CASES.items():
        store = syn.make_store(seed, M, shape, conformers=C)
        ref = reference_batch_conformers(store, idx, C)
        mine = collate_reference_conformers(store, idx, C)
        for k, v in ref.items():
            assert mine[k].dtype == v.dtype and mine[k].shape == v.shape, (name, k, mine[k].dtype, v.dtype, mine[k].shape, v.shape)
            assert np.array_equal(mine[k], v), "oracle != reference for %s/%s" % (name, k)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), seed=seed, n_molecules=M,
                            idx=np.array(idx), conformers=C, **ref)
        print("pinned %s: %d molecules x %d conformers, N3=%d E3=%d — oracle == reference (bit exact)"
              % (name, len(idx), C, int(ref["num_nodes3"].sum()), len(ref["src3"])))


if __name__ == "__main__":
    main()
