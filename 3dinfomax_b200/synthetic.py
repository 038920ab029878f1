"""Seeded synthetic molecule batches with the shapes of the reference's QM9 / QMugs inputs.

There is no network (no QM9 download, no rdkit), so tests and bench.py draw batches from this
generator.  What it reproduces from the reference's data contract:

  * 2-D bond graph: directed edges stored in both directions with the reverse pair adjacent
    (datasets/qm9_dataset.py:431-435), int64 OGB atom features [N,9] / bond features [E,3]
    within the OGB vocabulary sizes (commons/mol_encoder.py:4-7).
  * 3-D graph: complete digraph without self loops, ``src = repeat_interleave(arange(n), n-1)``,
    dst ascending (datasets/qm9_dataset.py:210-219), ``edata['d']`` = Euclidean distance
    fp32 [E3,1] (datasets/qm9_dataset.py:241-242).
  * batching = block-diagonal offsetting of node ids, order preserved (dgl.batch,
    datasets/custom_collate.py:105-114); conformers molecule-major (custom_collate.py:155-157).

Atom-count histogram = the empirical QM9 one; QMugs-shaped = round(N(40,10)) in [6,100].
Everything is numpy + a seeded ``default_rng`` so the same seed gives the same batch on any box.
"""
import numpy as np

ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_FEATURE_DIMS = [5, 6, 2]

_QM9_HIST = {3: 2, 4: 4, 5: 5, 6: 12, 7: 19, 8: 66, 9: 180, 10: 485, 11: 1066, 12: 2184, 13: 4001, 14: 6738,
             15: 10216, 16: 13875, 17: 17052, 18: 17532, 19: 18129, 20: 12435, 21: 13048, 22: 4447, 23: 6286,
             24: 705, 25: 1903, 26: 58, 27: 350, 29: 33}


def _sample_atom_counts(rng, batch_size, shape):
    if shape == "qm9":
        ks = np.array(sorted(_QM9_HIST))
        p = np.array([_QM9_HIST[k] for k in ks], dtype=np.float64)
        return rng.choice(ks, size=batch_size, p=p / p.sum())
    if shape == "qmugs":
        return np.clip(np.rint(rng.normal(40.0, 10.0, size=batch_size)), 6, 100).astype(np.int64)
    raise ValueError("unknown shape %r" % (shape,))


def _unit(rng):
    v = rng.normal(size=3)
    return v / (np.linalg.norm(v) + 1e-12)


def make_molecule(rng, n):
    """Returns (bonds [(u,v)...] undirected in creation order, coords [n,3] float64, is_heavy [n])."""
    nh = max(1, int(round(0.49 * n)))
    nh = min(nh, n)
    # make sure the hydrogens fit: capacity 2*nh + 2 (tree) must cover n - nh
    while nh < n and 2 * nh + 2 < n - nh:
        nh += 1
    deg = np.zeros(n, dtype=np.int64)
    bonds = []
    adj = set()
    pos = np.zeros((n, 3))

    def place(child, parent, length):
        for _ in range(20):
            cand = pos[parent] + length * _unit(rng)
            if child == 0 or np.min(np.linalg.norm(pos[:child] - cand, axis=1)) > 0.95:
                break
        pos[child] = cand

    for i in range(1, nh):
        cands = np.nonzero(deg[:i] < 4)[0]
        # keep one valence free where needed so that hydrogens still fit
        p = int(rng.choice(cands))
        bonds.append((p, i))
        adj.add((p, i))
        deg[p] += 1
        deg[i] += 1
        place(i, p, 1.5)
    n_h = n - nh
    rings = int(rng.poisson(1.6)) if nh >= 3 else 0
    for _ in range(rings):
        free = 4 * nh - int(deg[:nh].sum())
        if free - 2 < n_h:
            break
        cands = np.nonzero(deg[:nh] < 4)[0]
        if len(cands) < 2:
            break
        ok = False
        for _ in range(8):
            u, v = rng.choice(cands, size=2, replace=False)
            u, v = int(min(u, v)), int(max(u, v))
            if (u, v) not in adj:
                ok = True
                break
        if not ok:
            continue
        bonds.append((u, v))
        adj.add((u, v))
        deg[u] += 1
        deg[v] += 1
    for i in range(nh, n):
        cands = np.nonzero(deg[:nh] < 4)[0]
        if len(cands) == 0:
            cands = np.arange(nh)
        p = int(rng.choice(cands))
        bonds.append((p, i))
        deg[p] += 1
        deg[i] += 1
        place(i, p, 1.09)
    heavy = np.zeros(n, dtype=bool)
    heavy[:nh] = True
    return bonds, pos, heavy


def complete_graph_edges(n):
    """src = repeat_interleave(arange(n), n-1); dst = all other nodes ascending (qm9_dataset.py:210-219)."""
    ar = np.arange(n, dtype=np.int64)
    src = np.repeat(ar, n - 1)
    dst = np.concatenate([np.concatenate([ar[:i], ar[i + 1:]]) for i in range(n)]) if n > 1 else np.zeros(0, np.int64)
    return src, dst


def make_batch(seed, batch_size, shape="qm9", conformers=1, conformer_noise=0.3):
    """Returns a dict of numpy arrays describing one collated batch (2-D graph + 3-D conformer graphs)."""
    rng = np.random.default_rng(seed)
    counts = _sample_atom_counts(rng, batch_size, shape)
    src_l, dst_l, ef_l, xf_l, nn_l, ne_l = [], [], [], [], [], []
    src3_l, dst3_l, d3_l, nn3_l, ne3_l, xyz_l = [], [], [], [], [], []
    off = 0
    off3 = 0
    heavy_z = np.array([5, 6, 7, 8])  # OGB atomic-number index = Z-1: C,N,O,F
    for n in counts.tolist():
        bonds, pos, heavy = make_molecule(rng, n)
        nb = len(bonds)
        b = np.array(bonds, dtype=np.int64).reshape(nb, 2)
        s = np.empty(2 * nb, dtype=np.int64)
        t = np.empty(2 * nb, dtype=np.int64)
        s[0::2], t[0::2] = b[:, 0], b[:, 1]
        s[1::2], t[1::2] = b[:, 1], b[:, 0]
        bf = np.stack([rng.integers(0, 4, size=nb), np.zeros(nb, dtype=np.int64), rng.integers(0, 2, size=nb)], 1)
        ef = np.repeat(bf, 2, axis=0)
        xf = np.stack([rng.integers(0, d, size=n) for d in ATOM_FEATURE_DIMS], 1)
        xf[:, 0] = np.where(heavy, rng.choice(heavy_z, size=n), 0)
        src_l.append(s + off)
        dst_l.append(t + off)
        ef_l.append(ef)
        xf_l.append(xf)
        nn_l.append(n)
        ne_l.append(2 * nb)
        off += n
        cs, cd = complete_graph_edges(n)
        for c in range(conformers):
            p = pos if c == 0 else pos + rng.normal(0.0, conformer_noise, size=pos.shape)
            d = np.linalg.norm(p[cs] - p[cd], axis=1)
            src3_l.append(cs + off3)
            dst3_l.append(cd + off3)
            d3_l.append(d)
            nn3_l.append(n)
            ne3_l.append(len(cs))
            xyz_l.append(p)
            off3 += n
    return {
        "batch_size": int(batch_size), "conformers": int(conformers),
        "x_atom": np.concatenate(xf_l).astype(np.int64),
        "e_attr": np.concatenate(ef_l).astype(np.int64).reshape(-1, 3),
        "src": np.concatenate(src_l), "dst": np.concatenate(dst_l),
        "num_nodes": np.array(nn_l, dtype=np.int64), "num_edges": np.array(ne_l, dtype=np.int64),
        "src3": np.concatenate(src3_l), "dst3": np.concatenate(dst3_l),
        "d3": np.concatenate(d3_l).astype(np.float32).reshape(-1, 1),
        "num_nodes3": np.array(nn3_l, dtype=np.int64), "num_edges3": np.array(ne3_l, dtype=np.int64),
        "xyz3": np.concatenate(xyz_l).astype(np.float32),
    }


def slice_batch(b, lo, hi):
    """Molecules [lo, hi) of a collated batch as a collated batch of their own (node ids re-based): the shard a
    data-parallel rank works on."""
    C = int(b["conformers"])
    nn, ne = b["num_nodes"], b["num_edges"]
    nn3, ne3 = b["num_nodes3"], b["num_edges3"]
    n0, n1 = int(nn[:lo].sum()), int(nn[:hi].sum())
    e0, e1 = int(ne[:lo].sum()), int(ne[:hi].sum())
    m0, m1 = int(nn3[:lo * C].sum()), int(nn3[:hi * C].sum())
    f0, f1 = int(ne3[:lo * C].sum()), int(ne3[:hi * C].sum())
    return {
        "batch_size": hi - lo, "conformers": C,
        "x_atom": b["x_atom"][n0:n1], "e_attr": b["e_attr"][e0:e1],
        "src": b["src"][e0:e1] - n0, "dst": b["dst"][e0:e1] - n0,
        "num_nodes": nn[lo:hi], "num_edges": ne[lo:hi],
        "src3": b["src3"][f0:f1] - m0, "dst3": b["dst3"][f0:f1] - m0, "d3": b["d3"][f0:f1],
        "num_nodes3": nn3[lo * C:hi * C], "num_edges3": ne3[lo * C:hi * C], "xyz3": b["xyz3"][m0:m1],
    }


def make_store(seed, n_molecules, shape="qm9", conformers=1, conformer_noise=0.3):
    """A packed molecule store with the fields of the reference's processed file (datasets/qm9_dataset.py:454-467):
    n_atoms [M], atom_slices / edge_slices [M+1] (leading 0), edge_indices [2, Etot] with molecule-LOCAL node ids,
    atom_features int64 [Ntot, 9], edge_features int64 [Etot, 3], coordinates fp32 [Ntot, 3].
    conformers > 1 adds ``conformations`` fp32 [Ntot, 3*conformers] (datasets/qmugs_dataset.py:96-101: conformer c in
    columns [3c, 3c+3), conformer 0 == coordinates); drawn from a separate generator so the other fields do not change."""
    rng = np.random.default_rng(seed)
    counts = _sample_atom_counts(rng, n_molecules, shape)
    heavy_z = np.array([5, 6, 7, 8])
    ei_l, ef_l, xf_l, xyz_l = [], [], [], []
    atom_slices, edge_slices = [0], [0]
    for n in counts.tolist():
        bonds, pos, heavy = make_molecule(rng, n)
        nb = len(bonds)
        b = np.array(bonds, dtype=np.int64).reshape(nb, 2)
        s = np.empty(2 * nb, dtype=np.int64)
        t = np.empty(2 * nb, dtype=np.int64)
        s[0::2], t[0::2] = b[:, 0], b[:, 1]
        s[1::2], t[1::2] = b[:, 1], b[:, 0]
        bf = np.stack([rng.integers(0, 4, size=nb), np.zeros(nb, dtype=np.int64), rng.integers(0, 2, size=nb)], 1)
        xf = np.stack([rng.integers(0, d, size=n) for d in ATOM_FEATURE_DIMS], 1)
        xf[:, 0] = np.where(heavy, rng.choice(heavy_z, size=n), 0)
        ei_l.append(np.stack([s, t]))
        ef_l.append(np.repeat(bf, 2, axis=0))
        xf_l.append(xf)
        xyz_l.append(pos)
        atom_slices.append(atom_slices[-1] + n)
        edge_slices.append(edge_slices[-1] + 2 * nb)
    out = {"n_atoms": counts.astype(np.int64), "atom_slices": np.array(atom_slices, dtype=np.int64),
           "edge_slices": np.array(edge_slices, dtype=np.int64),
           "edge_indices": np.concatenate(ei_l, axis=1).astype(np.int64),
           "atom_features": np.concatenate(xf_l).astype(np.int64),
           "edge_features": np.concatenate(ef_l).astype(np.int64).reshape(-1, 3),
           "coordinates": np.concatenate(xyz_l).astype(np.float32)}
    if conformers > 1:
        rng2 = np.random.default_rng([seed, 77])
        xyz = out["coordinates"].astype(np.float64)
        cols = [xyz] + [xyz + rng2.normal(0.0, conformer_noise, size=xyz.shape) for _ in range(conformers - 1)]
        out["conformations"] = np.concatenate(cols, axis=1).astype(np.float32)
    return out
