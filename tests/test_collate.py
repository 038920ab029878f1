"""Batch construction: the numpy oracle against the vectors produced by the reference's own
QM9Dataset.__getitem__ + contrastive_collate (oracle/pin_collate.py), and the host-side metadata of the device collate."""
import importlib
import os

import numpy as np
import pytest

from oracle import collate_oracle as CO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C = importlib.import_module("3dinfomax_b200.collate")
syn = importlib.import_module("3dinfomax_b200.synthetic")


@pytest.mark.parametrize("name,shape", [("collate_qm9", "qm9"), ("collate_qmugs", "qmugs")])
def test_oracle_matches_reference_vectors(name, shape):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    store = CO.make_store(int(g["seed"]), int(g["n_molecules"]), shape)
    got = CO.collate_reference(store, g["idx"])
    for k in ("src", "dst", "x_atom", "e_attr", "num_nodes", "num_edges", "src3", "dst3", "d3", "num_nodes3", "num_edges3"):
        assert got[k].dtype == g[k].dtype and got[k].shape == g[k].shape, k
        assert np.array_equal(got[k], g[k]), k           # integers and fp32 distances bit-exact


def test_pairwise_order_is_the_reference_order():
    s, d = CO.pairwise_edges(4)
    assert s.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    assert d.tolist() == [1, 2, 3, 0, 2, 3, 0, 1, 3, 0, 1, 2]
    s, d = CO.pairwise_edges(1)
    assert len(s) == 0 and len(d) == 0


def test_store_layout_and_batch_metadata():
    store = syn.make_store(5, 30)
    assert store["atom_slices"][0] == 0 and store["edge_slices"][0] == 0
    assert np.array_equal(np.diff(store["atom_slices"]), store["n_atoms"])
    assert store["edge_indices"].shape == (2, store["edge_slices"][-1])
    # molecule-local ids, reverse pairs adjacent (datasets/qm9_dataset.py:431-435)
    for m in range(30):
        e0, e1 = store["edge_slices"][m], store["edge_slices"][m + 1]
        ei = store["edge_indices"][:, e0:e1]
        assert ei.min() >= 0 and ei.max() < store["n_atoms"][m]
        assert np.array_equal(ei[0, 0::2], ei[1, 1::2]) and np.array_equal(ei[1, 0::2], ei[0, 1::2])
    idx = np.array([4, 4, 29, 0])
    h = np.zeros(C.metadata_len(len(idx)), dtype=np.int64)
    N, E, E3 = C.batch_metadata(h, idx, store["n_atoms"], np.diff(store["edge_slices"]))
    ref = CO.collate_reference(store, idx)
    assert (N, E, E3) == (len(ref["x_atom"]), len(ref["src"]), len(ref["src3"]))
    v = C.metadata_views(h, len(idx))
    assert np.array_equal(v["num_nodes"], ref["num_nodes"]) and np.array_equal(v["num_edges3"], ref["num_edges3"])
    assert np.array_equal(v["node_ptr"], np.concatenate([[0], np.cumsum(ref["num_nodes"])]))
    assert np.array_equal(v["edge3_ptr"], np.concatenate([[0], np.cumsum(ref["num_edges3"])]))


def test_store_requires_cuda():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        C.PackedMoleculeStore(syn.make_store(1, 3), "cpu")


def test_bucket_ladder_covers_an_epoch_with_few_levels():
    """trainer.BucketLadder is host logic over the store's per-molecule sizes: capacities are aligned and monotone,
    level_of returns the tightest level that holds a batch, and a shuffled epoch of QM9-shaped batches needs only a
    handful of levels (the number of CUDA graphs a training run captures)."""
    import types
    T = importlib.import_module("3dinfomax_b200.trainer")
    store = syn.make_store(3, 4096, "qm9")
    n_atoms = np.asarray(store["n_atoms"], dtype=np.int64)
    n_edges = np.diff(np.asarray(store["edge_slices"], dtype=np.int64))
    stub = types.SimpleNamespace(n_atoms=n_atoms, n_edges=n_edges)
    B = 512
    lad = T.BucketLadder(stub, B, conformers=1)
    caps = [lad.caps(k) for k in range(6)]
    for k, c in enumerate(caps):
        assert all(int(x) % 128 == 0 for x in c)
        if k:
            assert all(a >= b for a, b in zip(c, caps[k - 1]))
    rng = np.random.default_rng(0)
    levels = []
    for _ in range(5):
        perm = rng.permutation(len(n_atoms))
        for i in range(0, len(perm) - B + 1, B):
            ix = perm[i:i + B]
            n = n_atoms[ix]
            sizes = (int(n.sum()), int(n_edges[ix].sum()), int((n * (n - 1)).sum()))
            lv = lad.level_of(sizes)
            assert all(s <= c for s, c in zip(sizes, lad.caps(lv)))                 # the level holds the batch
            assert lv == 0 or any(s > c for s, c in zip(sizes, lad.caps(lv - 1)))   # ... and is the tightest one
            levels.append(lv)
    assert max(levels) <= 4 and np.mean(np.asarray(levels) == 0) > 0.6
    # padding of the common level stays small
    assert lad.caps(0)[0] <= 1.05 * B * n_atoms.mean() + 128
    # three conformers per molecule triple the 3-D edge capacity only
    lad3 = T.BucketLadder(stub, B, conformers=3)
    assert lad3.caps(0)[:2] == lad.caps(0)[:2] and lad3.caps(0)[2] > 2.5 * lad.caps(0)[2]
