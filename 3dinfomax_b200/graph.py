"""Graph batch contract of the hot path.

The reference feeds ``PNA.forward`` / ``Net3D.forward`` a batched ``dgl.DGLGraph``
(datasets/custom_collate.py:105-114).  DGL is not installable in this image, so:

  * ``GraphBatch`` is a light stand-in with the DGL attribute names the hot path and its callers touch
    (``ndata``, ``edata``, ``edges()``, ``batch_num_nodes()``, ``batch_num_edges()``, ``number_of_nodes()``,
    ``to()``), built by ``batch_from_numpy`` / ``GraphBatch(...)``;
  * a real ``DGLGraph`` is accepted as-is when ``dgl`` is importable — ``graph_structure`` only calls
    ``g.edges()`` (edge-id order), ``g.batch_num_nodes()`` and ``g.number_of_nodes()``.

``GraphStructure`` is what replaces DGL's per-call degree bucketing (models/pna.py:206): a destination-sorted,
edge-id-stable int32 CSR built ON THE DEVICE once per batch and cached on the graph object.
"""
import os

import torch

from . import kernels as K


class GraphBatch:
    """Batched graph: node ids of graph k are offset by the cumulative node counts (dgl.batch semantics)."""

    def __init__(self, src, dst, batch_num_nodes, batch_num_edges=None, ndata=None, edata=None, num_nodes=None,
                 max_in_degree=None):
        # max_in_degree: optional HOST integer known at collate time; lets the 2-D encoder group nodes by in-degree
        # (degree-merged posttrans weights) without a device->host sync.  None: the generic 13F-wide path is used.
        self.max_in_degree = None if max_in_degree is None else int(max_in_degree)
        self._src = src
        self._dst = dst
        self._bnn = batch_num_nodes
        self._bne = batch_num_edges
        self._n = int(num_nodes) if num_nodes is not None else int(batch_num_nodes.sum().item())
        self.ndata = dict(ndata or {})
        self.edata = dict(edata or {})
        self._i3d_struct = None

    # --- DGL-compatible surface -------------------------------------------------------------
    def edges(self):
        return self._src, self._dst

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self):
        return int(self._src.numel())

    num_edges = number_of_edges

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    @property
    def batch_size(self):
        return int(self._bnn.numel())

    @property
    def device(self):
        return self._src.device

    def to(self, device, non_blocking=False):
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        g = GraphBatch(mv(self._src), mv(self._dst), mv(self._bnn), mv(self._bne),
                       {k: mv(v) for k, v in self.ndata.items()}, {k: mv(v) for k, v in self.edata.items()}, self._n,
                       self.max_in_degree)
        if self._i3d_struct is not None and torch.device(device) == self.device:
            g._i3d_struct = self._i3d_struct
        return g

    def pin_memory(self):
        pm = lambda t: None if t is None else t.pin_memory()
        return GraphBatch(pm(self._src), pm(self._dst), pm(self._bnn), pm(self._bne),
                          {k: pm(v) for k, v in self.ndata.items()}, {k: pm(v) for k, v in self.edata.items()}, self._n,
                          self.max_in_degree)


def batch_from_numpy(b, device="cpu", pin=False):
    """numpy batch dict (synthetic.make_batch layout) -> (graph2d, graph3d) GraphBatch pair on ``device``."""
    t = lambda a: torch.from_numpy(a)
    import numpy as np
    max_deg = int(np.bincount(b["dst"]).max()) if len(b["dst"]) else 0
    g2 = GraphBatch(t(b["src"]), t(b["dst"]), t(b["num_nodes"]), t(b["num_edges"]),
                    {"feat": t(b["x_atom"])}, {"feat": t(b["e_attr"])}, max_in_degree=max_deg)
    g3 = GraphBatch(t(b["src3"]), t(b["dst3"]), t(b["num_nodes3"]), t(b["num_edges3"]), {}, {"d": t(b["d3"])})
    if pin:
        g2, g3 = g2.pin_memory(), g3.pin_memory()
    if torch.device(device).type != "cpu":
        g2, g3 = g2.to(device, non_blocking=pin), g3.to(device, non_blocking=pin)
    return g2, g3


class GraphStructure:
    """Device-resident CSR views of one batched graph (all int32).

    rowptr[N+1], src_csr[E], dst_csr[E], eid[E]      in-edges of node v = rows [rowptr[v], rowptr[v+1]),
                                                     ascending edge id == argsort(dst, stable)  (bit exact)
    out_rowptr[N+1], out_pos[E]                      out-edges of node v as POSITIONS into the CSR edge order
    graph_ptr[B+1]                                   node range of each molecule
    amp[N], att[N]                                   ln(D+1), 1/ln(D+1) degree scalers (fp32)
    plan                                             kernels.DegreePlan (nodes grouped by in-degree) or None
    """

    MAX_PLAN_BUCKETS = 16
    # shape-bucketed (padded) batches only: device int32 scalars with the number of valid nodes / edges; rows behind
    # them are padding (include/i3d.h "Padding convention").  None = every row is valid.
    n_valid = None
    e_valid = None
    code_csr = None      # optional int64 [E]: bond-feature combination index of every CSR-ordered edge (device collate)

    @classmethod
    def from_parts(cls, N, E, B, rowptr, src_csr, dst_csr, eid, out_rowptr, out_pos, graph_ptr, need_scalers=True,
                   max_in_degree=None, padded=False):
        """Structure whose CSR arrays were emitted by the collate kernels (i3d_collate_2d_struct / _3d_struct)."""
        st = cls.__new__(cls)
        st.N, st.E, st.B = int(N), int(E), int(B)
        st.rowptr, st.src_csr, st.dst_csr, st.eid = rowptr, src_csr, dst_csr, eid
        st.out_rowptr, st.out_pos, st.graph_ptr = out_rowptr, out_pos, graph_ptr
        if padded:
            st.n_valid = graph_ptr[st.B:st.B + 1]            # graph_ptr[B] = number of valid nodes
            st.e_valid = rowptr[st.N:st.N + 1]               # rowptr[n_cap] = number of valid edges
        st.amp = st.att = None
        st.plan = None
        if need_scalers:
            st.amp, st.att = K.degree_scalers(rowptr)
            st._maybe_plan(max_in_degree)
        return st

    def _maybe_plan(self, max_in_degree):
        use_plan = os.environ.get("I3D_POSTTRANS", "merged") != "generic"
        min_nodes = int(os.environ.get("I3D_PLAN_MIN_NODES", "256"))
        if (use_plan and max_in_degree is not None and self.N >= min_nodes
                and int(max_in_degree) + 1 <= self.MAX_PLAN_BUCKETS):
            self.plan = K.DegreePlan(self.rowptr, int(max_in_degree) + 1)

    def __init__(self, src, dst, batch_num_nodes, num_nodes, need_out=True, need_scalers=True, max_in_degree=None):
        if not src.is_cuda:
            raise RuntimeError("graph tensors must live on a CUDA device: the 3dinfomax_b200 path has no CPU fallback")
        src = src.contiguous()
        dst = dst.contiguous()
        if src.dtype != torch.int64:
            src, dst = src.long(), dst.long()
        self.N = int(num_nodes)
        self.E = int(src.numel())
        self.B = int(batch_num_nodes.numel())
        self.rowptr, self.src_csr, self.dst_csr, self.eid = K.csr_build(dst, src, self.N)
        if need_out:
            self.out_rowptr, _, _, self.out_pos = K.csr_build(self.src_csr, self.dst_csr, self.N)
        else:
            self.out_rowptr = self.out_pos = None
        bnn = batch_num_nodes.to(device=src.device, dtype=torch.int64).contiguous()
        self.graph_ptr = K.segment_ptr(bnn)
        if need_scalers:
            self.amp, self.att = K.degree_scalers(self.rowptr)
        else:
            self.amp = self.att = None
        self.plan = None
        if need_scalers:          # tiny batches (< I3D_PLAN_MIN_NODES): padding to whole tiles per bucket outweighs K
            self._maybe_plan(max_in_degree)

    def check_plan(self):
        """Host check (device->host sync) that no node exceeded the ``max_in_degree`` the plan was built for."""
        if self.plan is not None and int(self.plan.overflow.item()) != 0:
            raise RuntimeError("a node's in-degree exceeds the graph's max_in_degree hint (%d): the degree-merged "
                               "posttrans results are invalid" % (self.plan.n_buckets - 1))


def annotate_max_in_degree(graph):
    """Set ``graph.max_in_degree`` from the edge list (one device->host sync if the graph lives on the GPU).  Call it
    at collate time for graphs that do not come from ``batch_from_numpy`` (e.g. a DGLGraph) to enable the degree plan."""
    _, dst = graph.edges()
    n = graph.number_of_nodes()
    md = int(torch.bincount(dst.long(), minlength=max(n, 1)).max().item()) if dst.numel() else 0
    try:
        graph.max_in_degree = md
    except AttributeError:
        pass
    return md


def graph_structure(graph, need_out=True, need_scalers=True):
    """CSR structure of ``graph`` (GraphBatch or DGLGraph), built once and cached on the object."""
    st = getattr(graph, "_i3d_struct", None)
    if st is not None and (st.out_pos is not None or not need_out) and (st.amp is not None or not need_scalers):
        return st
    src, dst = graph.edges()
    st = GraphStructure(src, dst, graph.batch_num_nodes(), graph.number_of_nodes(), need_out, need_scalers,
                        getattr(graph, "max_in_degree", None))
    if st.plan is not None and os.environ.get("I3D_CHECK_PLAN", "0") == "1":
        st.check_plan()
    try:
        graph._i3d_struct = st
    except AttributeError:
        pass
    return st
