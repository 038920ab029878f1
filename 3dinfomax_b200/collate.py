"""Device-resident packed molecule store + on-device batch construction.

The reference builds every batch in Python: B calls of ``QM9Dataset.__getitem__`` (datasets/qm9_dataset.py:189-244 —
``get_graph`` slices the packed store, ``get_complete_graph`` builds the complete digraph and its distances), then
``contrastive_collate`` = ``dgl.batch`` (datasets/custom_collate.py:105-114), then a host->device copy of both graphs
(4.8 MB per batch of 512 QM9 molecules).  Here the store of the reference's processed file
(qm9_dataset.py:454-467) lives in HBM, and a step uploads ONLY the molecule indices plus three offset arrays
(one pinned buffer, 29 KB at B = 512); two kernels (i3d_collate_2d / i3d_collate_3d) emit the batched graphs in the
reference's exact node / edge order, with the 3-D complete graphs and their distances generated implicitly.

    store = PackedMoleculeStore(store_dict, device)          # store_dict: numpy arrays, see synthetic.make_store
    g2, g3 = store.collate(idx)                               # idx: host int array of molecule ids (sampler output)
    loss, *_ = trainer.process_batch(([g2], [g3]))

There is no CPU fallback: the store must live on a CUDA device.
"""
import numpy as np
import torch

from . import lib as _lib
from .graph import GraphBatch

_FIELDS = ("n_atoms", "atom_slices", "edge_slices", "edge_indices", "atom_features", "edge_features", "coordinates")


class PackedMoleculeStore:
    def __init__(self, store, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("PackedMoleculeStore lives in HBM: the 3dinfomax_b200 path has no CPU fallback")
        for k in _FIELDS:
            if k not in store:
                raise KeyError("packed store is missing %r (fields of the reference's processed file)" % k)
        self.device = device
        self.M = int(len(store["n_atoms"]))
        # host copies of the O(M) metadata: sizes of a batch are computed on the host without touching the device
        self.n_atoms = np.ascontiguousarray(store["n_atoms"], dtype=np.int64)
        self.atom_slices_h = np.ascontiguousarray(store["atom_slices"], dtype=np.int64)
        self.edge_slices_h = np.ascontiguousarray(store["edge_slices"], dtype=np.int64)
        if len(self.atom_slices_h) != self.M + 1 or len(self.edge_slices_h) != self.M + 1:
            raise ValueError("atom_slices / edge_slices must have M+1 entries with a leading 0")
        self.n_edges = np.diff(self.edge_slices_h)
        ei = np.ascontiguousarray(store["edge_indices"], dtype=np.int64)
        self.Etot = int(ei.shape[1])
        # per-molecule maximum in-degree (host, once): lets the encoder size its degree plan without a device sync
        deg = np.zeros(int(self.atom_slices_h[-1]) + 1, dtype=np.int64)
        mol_of_edge = np.repeat(np.arange(self.M), self.n_edges)
        np.add.at(deg, ei[1] + self.atom_slices_h[:-1][mol_of_edge], 1)
        self.max_in_degree = np.zeros(self.M, dtype=np.int64)
        nz = self.n_atoms > 0
        if nz.any():
            self.max_in_degree[nz] = np.maximum.reduceat(deg[:-1], self.atom_slices_h[:-1][nz])
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
        self.atom_slices = t(self.atom_slices_h, np.int64)
        self.edge_slices = t(self.edge_slices_h, np.int64)
        self.edge_indices = t(ei, np.int64)
        self.atom_features = t(store["atom_features"], np.int64)
        self.edge_features = t(store["edge_features"], np.int64)
        self.coordinates = t(store["coordinates"], np.float32)
        self.CA = int(self.atom_features.shape[1])
        self.CE = int(self.edge_features.shape[1])
        self._staging = {}

    def __len__(self):
        return self.M

    def _stage(self, B):
        """pinned host + device buffer holding the per-batch metadata (layout of ``batch_metadata``)"""
        st = self._staging.get(B)
        if st is None:
            n = metadata_len(B)
            st = (torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.int64, device=self.device))
            self._staging[B] = st
        return st

    def collate(self, idx):
        """Batched (2-D bond graph, 3-D complete graph) of molecules ``idx`` — contrastive_collate of the reference."""
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int64).reshape(-1)
        B = int(len(idx))
        if B == 0:
            raise ValueError("empty batch")
        if idx.min() < 0 or idx.max() >= self.M:
            raise IndexError("molecule index out of range [0, %d)" % self.M)
        host, dev = self._stage(B)
        N, E, E3 = batch_metadata(host.numpy(), idx, self.n_atoms, self.n_edges)
        dev.copy_(host, non_blocking=True)                       # the ONLY host->device traffic of the batch
        v = metadata_views(dev, B)
        i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=self.device)
        src, dst, x_atom, e_attr = i64(E), i64(E), i64(N, self.CA), i64(E, self.CE)
        src3, dst3 = i64(E3), i64(E3)
        d3 = torch.empty(E3, 1, dtype=torch.float32, device=self.device)
        L = _lib.load()
        s = torch.cuda.current_stream().cuda_stream
        p = lambda t: t.data_ptr()
        _lib.check(L.i3d_collate_2d(p(v["idx"]), B, p(self.atom_slices), p(self.edge_slices), p(self.edge_indices),
                                    self.Etot, p(self.atom_features), self.CA, p(self.edge_features), self.CE,
                                    p(v["node_ptr"]), p(v["edge_ptr"]), N, E, p(src), p(dst), p(x_atom), p(e_attr), s),
                   "i3d_collate_2d")
        _lib.check(L.i3d_collate_3d(p(v["idx"]), B, p(self.atom_slices), p(self.coordinates), p(v["node_ptr"]),
                                    p(v["edge3_ptr"]), E3, p(src3), p(dst3), p(d3), s), "i3d_collate_3d")
        # the count tensors are fresh copies: the staging buffer is overwritten by the next batch of this size
        g2 = GraphBatch(src, dst, v["num_nodes"].clone(), v["num_edges"].clone(), {"feat": x_atom}, {"feat": e_attr}, N,
                        max_in_degree=int(self.max_in_degree[idx].max()))
        g3 = GraphBatch(src3, dst3, v["num_nodes"].clone(), v["num_edges3"].clone(), {}, {"d": d3}, N)
        return g2, g3


def metadata_len(B):
    return B + 3 * (B + 1) + 3 * B


def metadata_views(buf, B):
    """slices of the metadata buffer: idx[B], node_ptr / edge_ptr / edge3_ptr [B+1], num_nodes / num_edges / num_edges3 [B]"""
    o, out = 0, {}
    for name, n in (("idx", B), ("node_ptr", B + 1), ("edge_ptr", B + 1), ("edge3_ptr", B + 1), ("num_nodes", B),
                    ("num_edges", B), ("num_edges3", B)):
        out[name] = buf[o:o + n]
        o += n
    return out


def batch_metadata(h, idx, n_atoms, n_edges):
    """Fill the int64 host array ``h`` (metadata_len(B) entries) for molecules ``idx``; returns (N, E, E3).
    Pure host arithmetic over O(B) integers (numpy): the per-step replacement of dgl.batch's bookkeeping."""
    B = len(idx)
    v = metadata_views(h, B)
    n, ne = n_atoms[idx], n_edges[idx]
    v["idx"][:] = idx
    for ptr, cnt, counts in (("node_ptr", "num_nodes", n), ("edge_ptr", "num_edges", ne),
                             ("edge3_ptr", "num_edges3", n * (n - 1))):
        v[cnt][:] = counts
        v[ptr][0] = 0
        np.cumsum(counts, out=v[ptr][1:])
    return int(v["node_ptr"][B]), int(v["edge_ptr"][B]), int(v["edge3_ptr"][B])
