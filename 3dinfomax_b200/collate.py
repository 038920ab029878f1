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
from .graph import GraphBatch, GraphStructure

_FIELDS = ("n_atoms", "atom_slices", "edge_slices", "edge_indices", "atom_features", "edge_features", "coordinates")


class _StagingRing:
    """Pinned host buffers for the per-batch metadata, each guarded by a CUDA event: a slot is rewritten only after
    the host->device copy that last read it has completed.  (One buffer per batch size, rewritten every call, lets
    the host overwrite the pinned memory of batch k while its asynchronous copy is still queued behind earlier GPU
    work — the kernels of batch k would then see the indices of batch k+1.)"""

    def __init__(self, n, slots=4):
        self.slots = [(torch.empty(n, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(slots)]
        self.used = [False] * slots
        self.next = 0

    def acquire(self):
        i = self.next
        self.next = (i + 1) % len(self.slots)
        host, ev = self.slots[i]
        if self.used[i]:
            ev.synchronize()
        self.used[i] = True
        return host, ev


def local_csr(n_atoms, atom_slices, edge_slices, edge_indices):
    """Molecule-local CSR arrays of a packed store (host, numpy, once per store; int32):
    in_rowptr_l[Ntot]   in-edges of the earlier atoms of the same molecule
    in_eid_l[Etot]      local edge id held by local CSR position p (destination-sorted, edge-id stable)
    out_rowptr_l[Ntot]  same for the CSR-ordered edges sorted by source
    out_pos_l[Etot]     local CSR position held by local out-slot s
    dgl.batch only offsets ids, so the CSR of any batch is these arrays plus the batch offsets (i3d_collate_2d_struct)."""
    M = len(n_atoms)
    n_edges = np.diff(edge_slices)
    mol_of_edge = np.repeat(np.arange(M), n_edges)
    mol_of_atom = np.repeat(np.arange(M), n_atoms)
    a0, e0 = atom_slices[:-1], edge_slices[:-1]
    gsrc = edge_indices[0] + a0[mol_of_edge]
    gdst = edge_indices[1] + a0[mol_of_edge]
    order = np.argsort(gdst, kind="stable")                      # store edge id at global CSR position
    in_eid_l = order - e0[mol_of_edge]                           # position p and edge order[p] share the molecule
    atoms = np.arange(int(atom_slices[-1]))
    in_rowptr_l = np.searchsorted(gdst[order], atoms, side="left") - e0[mol_of_atom]
    src_csr = gsrc[order]
    order2 = np.argsort(src_csr, kind="stable")                  # CSR position at global out-slot
    out_pos_l = order2 - e0[mol_of_edge]
    out_rowptr_l = np.searchsorted(src_csr[order2], atoms, side="left") - e0[mol_of_atom]
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return i32(in_rowptr_l), i32(in_eid_l), i32(out_rowptr_l), i32(out_pos_l)


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
        # conformers: [Ntot, 3*C] fp32, conformer c in columns [3c, 3c+3) (datasets/qmugs_dataset.py `conformations`);
        # absent = the single conformer of `coordinates` (QM9)
        conf = store.get("conformations") if hasattr(store, "get") else None
        self.conformations = self.coordinates if conf is None else t(conf, np.float32)
        self.n_conformers = int(self.conformations.shape[1]) // 3
        # regression targets of the fine-tuning configs ([M, T] fp32; `targets` of the processed file, already
        # normalised by the dataset — datasets/qm9_dataset.py:176-187); absent for pre-training stores
        tg = store.get("targets") if hasattr(store, "get") else None
        self.targets = None if tg is None else t(np.asarray(tg).reshape(self.M, -1), np.float32)
        self.CA = int(self.atom_features.shape[1])
        self.CE = int(self.edge_features.shape[1])
        # molecule-local CSR arrays: the structure of a batch is emitted by the collate kernel without any sort
        lc = local_csr(self.n_atoms, self.atom_slices_h, self.edge_slices_h, ei)
        self.in_rowptr_l, self.in_eid_l, self.out_rowptr_l, self.out_pos_l = (torch.from_numpy(a).to(device) for a in lc)
        self.max_in_degree_all = int(self.max_in_degree.max()) if self.M else 0
        # mixed-radix weights of the categorical edge-feature row (OGB bond feature vocabulary sizes 5, 6, 2 ->
        # (12, 2, 1); commons/mol_encoder.py:4-7): the collate kernel emits each edge's combination index with them
        from .synthetic import BOND_FEATURE_DIMS
        mult, acc = [], 1
        for d in reversed(BOND_FEATURE_DIMS[:self.CE] if self.CE <= len(BOND_FEATURE_DIMS) else [1] * self.CE):
            mult.append(acc)
            acc *= d
        self.code_mult = torch.tensor(list(reversed(mult)), dtype=torch.int64, device=device)
        self._staging = {}

    def __len__(self):
        return self.M

    def _stage(self, B):
        """event-guarded ring of pinned host buffers for the per-batch metadata (layout of ``batch_metadata``)"""
        st = self._staging.get(B)
        if st is None:
            st = self._staging[B] = _StagingRing(metadata_len(B))
        return st

    def stage_metadata(self, idx, dev_out=None):
        """Host part of a batch: validate ``idx``, fill a pinned metadata buffer and enqueue its copy to the device
        (the ONLY host->device traffic of the batch).  Returns (device metadata buffer, (N, E, E3))."""
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int64).reshape(-1)
        B = int(len(idx))
        if B == 0:
            raise ValueError("empty batch")
        if idx.min() < 0 or idx.max() >= self.M:
            raise IndexError("molecule index out of range [0, %d)" % self.M)
        host, ev = self._stage(B).acquire()
        sizes = batch_metadata(host.numpy(), idx, self.n_atoms, self.n_edges)
        dev = torch.empty(metadata_len(B), dtype=torch.int64, device=self.device) if dev_out is None else dev_out
        dev.copy_(host, non_blocking=True)
        ev.record()                                  # the pinned slot is free again once this copy has run
        return dev, sizes

    def collate(self, idx):
        """Batched (2-D bond graph, 3-D complete graph) of molecules ``idx`` — contrastive_collate of the reference."""
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int64).reshape(-1)
        B = int(len(idx))
        dev, (N, E, E3) = self.stage_metadata(idx)               # fresh device buffer per batch
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
        g2 = GraphBatch(src, dst, v["num_nodes"], v["num_edges"], {"feat": x_atom}, {"feat": e_attr}, N,
                        max_in_degree=int(self.max_in_degree[idx].max()))
        g3 = GraphBatch(src3, dst3, v["num_nodes"], v["num_edges3"], {}, {"d": d3}, N)
        return g2, g3

    def collate_padded(self, meta, B, n_cap, e_cap, e3_cap, conformers=1, need_3d=True):
        """Kernel-only part of a shape-bucketed batch (CUDA-graph capturable): from the DEVICE metadata buffer ``meta``
        (stage_metadata) emit both graphs padded to the bucket capacities, together with their CSR structures
        (i3d_collate_2d_struct / i3d_collate_3d_struct: no sort, no atomics).  Valid sizes stay on the device."""
        C = int(conformers)
        if C < 1 or C > self.n_conformers:
            raise ValueError("store holds %d conformer(s) per molecule, %d requested" % (self.n_conformers, C))
        v = metadata_views(meta, B)
        dev = self.device
        i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=dev)
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        n3_cap = C * n_cap
        src, dst, x_atom, e_attr = i64(e_cap), i64(e_cap), i64(n_cap, self.CA), i64(e_cap, self.CE)
        rowptr, out_rowptr, graph_ptr = i32(n_cap + 1), i32(n_cap + 1), i32(B + 1)
        src_csr, dst_csr, eid, out_pos = i32(e_cap), i32(e_cap), i32(e_cap), i32(e_cap)
        code_csr = i64(e_cap)
        if not need_3d:                                     # forward-only 2-D use (inference.Fingerprinter)
            e3_cap = 0
        src3, dst3 = i64(e3_cap), i64(e3_cap)
        d3 = torch.empty(e3_cap, 1, dtype=torch.float32, device=dev)
        rowptr3, graph_ptr3, nn3 = i32(n3_cap + 1), i32(B * C + 1), i64(B * C)
        src_csr3, dst_csr3, eid3, out_pos3 = i32(e3_cap), i32(e3_cap), i32(e3_cap), i32(e3_cap)
        L = _lib.load()
        s = torch.cuda.current_stream().cuda_stream
        p = lambda t: t.data_ptr()
        _lib.check(L.i3d_collate_2d_struct(p(v["idx"]), B, p(self.atom_slices), p(self.edge_slices), p(self.edge_indices),
                                           self.Etot, p(self.atom_features), self.CA, p(self.edge_features), self.CE,
                                           p(self.in_rowptr_l), p(self.in_eid_l), p(self.out_rowptr_l),
                                           p(self.out_pos_l), p(v["node_ptr"]), p(v["edge_ptr"]), n_cap, e_cap, p(src),
                                           p(dst), p(x_atom), p(e_attr), p(rowptr), p(src_csr), p(dst_csr), p(eid),
                                           p(out_rowptr), p(out_pos), p(graph_ptr), p(self.code_mult), p(code_csr), s),
                   "i3d_collate_2d_struct")
        g2 = GraphBatch(src, dst, v["num_nodes"], v["num_edges"], {"feat": x_atom}, {"feat": e_attr}, n_cap,
                        max_in_degree=self.max_in_degree_all)
        g2._i3d_struct = GraphStructure.from_parts(n_cap, e_cap, B, rowptr, src_csr, dst_csr, eid, out_rowptr, out_pos,
                                                   graph_ptr, True, self.max_in_degree_all, padded=True)
        g2._i3d_struct.code_csr = code_csr                  # bond-feature combination index of every CSR-ordered edge
        if not need_3d:
            return g2, None
        _lib.check(L.i3d_collate_3d_struct(p(v["idx"]), B, C, p(self.atom_slices), p(self.conformations),
                                           int(self.conformations.shape[1]), p(v["node_ptr"]), p(v["edge3_ptr"]), n3_cap,
                                           e3_cap, p(src3), p(dst3), p(d3), p(rowptr3), p(src_csr3), p(dst_csr3), p(eid3),
                                           p(out_pos3), p(graph_ptr3), p(nn3), s), "i3d_collate_3d_struct")
        g3 = GraphBatch(src3, dst3, nn3, None, {}, {"d": d3}, n3_cap)
        g3._i3d_struct = GraphStructure.from_parts(n3_cap, e3_cap, B * C, rowptr3, src_csr3, dst_csr3, eid3, rowptr3,
                                                   out_pos3, graph_ptr3, False, None, padded=True)
        return g2, g3

    def batch_sizes(self, idx, conformers=1):
        """(N, E, E3) of the batch ``idx`` from the host copies of the store metadata (no device work)."""
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        n = self.n_atoms[idx]
        return int(n.sum()), int(self.n_edges[idx].sum()), int(conformers) * int((n * (n - 1)).sum())


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
