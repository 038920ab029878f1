"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's batch construction for the pre-training path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this; the product path
(3dinfomax_b200/collate.py) never does.  Parity pinned: oracle/pin_collate.py runs the reference's own
``QM9Dataset.__getitem__`` / ``contrastive_collate`` (datasets/qm9_dataset.py:189-244, datasets/custom_collate.py:105-114)
under a stub of the three DGL calls they make and asserts equality with this file on a seeded store; the vectors are
committed as tests/golden/collate_*.npz.

What the reference does per step (one molecule at a time, in Python):
  * ``get_graph`` (qm9_dataset.py:219-229): 2-D bond graph of molecule i from the packed store —
    ``edge_indices[:, e_start:e_end]`` (molecule-local node ids), ``atom_features[start:start+n]`` int64 [n,9],
    ``edge_features[e_start:e_end]`` int64 [e,3], ``coordinates[start:start+n]``.
  * ``get_complete_graph`` (qm9_dataset.py:231-244) with ``get_pairwise`` (:207-217): complete digraph without self
    loops, ``src = repeat_interleave(arange(n), n-1)``, ``dst`` = the other nodes in ascending order;
    ``edata['d'] = ||x[src] - x[dst]||_2`` fp32 [n(n-1), 1].
  * ``contrastive_collate`` (custom_collate.py:105-114) = ``dgl.batch`` of both lists: node ids of molecule k are offset
    by the node counts of molecules 0..k-1, node / edge order preserved, ``batch_num_nodes`` / ``batch_num_edges`` kept.
"""
import numpy as np


def make_store(seed, n_molecules, shape="qm9"):
    """A packed molecule store with the fields of the reference's processed file (qm9_dataset.py:454-467), filled from
    the seeded synthetic molecule generator (there is no rdkit / QM9 download here)."""
    import importlib
    syn = importlib.import_module("3dinfomax_b200.synthetic")
    return syn.make_store(seed, n_molecules, shape)


def pairwise_edges(n):
    """qm9_dataset.py:207-217 — (src, dst) of the complete digraph without self loops, reference order."""
    ar = np.arange(n, dtype=np.int64)
    src = np.repeat(ar, n - 1)
    dst = np.concatenate([np.concatenate([ar[:i], ar[i + 1:]]) for i in range(n)]) if n > 1 else np.zeros(0, np.int64)
    return src, dst


def norm3_fp32(diff):
    """torch.norm(diff, p=2, dim=-1) for fp32 [.., 3] as torch evaluates it (probe in oracle/pin_collate.py: bit-equal
    on 200k random vectors): a fused-multiply-add chain acc = x*x; acc = fma(y, y, acc); acc = fma(z, z, acc) rounded
    to fp32 after every step, then an fp32 square root.  The fma is emulated in fp64 (products of fp32 are exact)."""
    d = diff.astype(np.float64)
    r32 = lambda v: v.astype(np.float32).astype(np.float64)
    acc = r32(d[..., 0] * d[..., 0])
    acc = r32(d[..., 1] * d[..., 1] + acc)
    acc = r32(d[..., 2] * d[..., 2] + acc)
    return np.sqrt(acc.astype(np.float32))


def collate_reference_conformers(store, idx, conformers):
    """Same for the multi-conformer datasets (datasets/qmugs_dataset.py:149-166, return type 'conformations'): the 3-D
    entry of molecule i is dgl.batch of ``conformers`` complete graphs, conformer c with coordinates
    ``conformations[start:start+n, 3c:3c+3]`` (conformer 0 == ``coordinates``) and the same pairwise edge order;
    contrastive_collate (custom_collate.py:105-114) then batches molecule-major.  (The reference adds N(0, 0.05) noise to
    a conformer that is identical to conformer 0 — random, not part of the deterministic restatement; the synthetic
    stores never contain such duplicates.)"""
    b = collate_reference(store, idx)
    idx = np.asarray(idx, dtype=np.int64)
    C = int(conformers)
    conf = store["conformations"] if C > 1 else store["coordinates"]
    src3, dst3, d3, nn3, ne3 = [], [], [], [], []
    off = 0
    for k, i in enumerate(idx):
        n = int(store["n_atoms"][i])
        a0 = int(store["atom_slices"][i])
        s3, t3 = pairwise_edges(n)
        for c in range(C):
            x = conf[a0:a0 + n, 3 * c:3 * c + 3].astype(np.float32)
            d3.append(norm3_fp32(x[s3] - x[t3])[:, None])
            src3.append(s3 + off)
            dst3.append(t3 + off)
            nn3.append(n)
            ne3.append(n * (n - 1))
            off += n
    b.update(src3=np.concatenate(src3).astype(np.int64), dst3=np.concatenate(dst3).astype(np.int64),
             d3=np.concatenate(d3).astype(np.float32), num_nodes3=np.array(nn3, dtype=np.int64),
             num_edges3=np.array(ne3, dtype=np.int64), conformers=C, batch_size=len(idx))
    return b


def collate_reference(store, idx):
    """numpy dict in the layout of 3dinfomax_b200.synthetic.make_batch for molecules ``idx`` of ``store``."""
    idx = np.asarray(idx, dtype=np.int64)
    n_atoms = store["n_atoms"][idx]
    a0 = store["atom_slices"][idx]
    e0, e1 = store["edge_slices"][idx], store["edge_slices"][idx + 1]
    off = np.concatenate([[0], np.cumsum(n_atoms)])
    src, dst, x_atom, e_attr, src3, dst3, d3 = [], [], [], [], [], [], []
    for k in range(len(idx)):
        n = int(n_atoms[k])
        ei = store["edge_indices"][:, e0[k]:e1[k]]
        src.append(ei[0] + off[k])
        dst.append(ei[1] + off[k])
        x_atom.append(store["atom_features"][a0[k]:a0[k] + n])
        e_attr.append(store["edge_features"][e0[k]:e1[k]])
        s3, t3 = pairwise_edges(n)
        x = store["coordinates"][a0[k]:a0[k] + n].astype(np.float32)
        d3.append(norm3_fp32(x[s3] - x[t3])[:, None])
        src3.append(s3 + off[k])
        dst3.append(t3 + off[k])
    cat = lambda xs, dt, tail=(): (np.concatenate(xs).astype(dt) if xs else np.zeros((0,) + tail, dt))
    return {"src": cat(src, np.int64), "dst": cat(dst, np.int64), "x_atom": cat(x_atom, np.int64, (9,)),
            "e_attr": cat(e_attr, np.int64, (3,)), "num_nodes": n_atoms.astype(np.int64),
            "num_edges": (e1 - e0).astype(np.int64), "src3": cat(src3, np.int64), "dst3": cat(dst3, np.int64),
            "d3": cat(d3, np.float32, (1,)), "num_nodes3": n_atoms.astype(np.int64),
            "num_edges3": (n_atoms * (n_atoms - 1)).astype(np.int64)}
