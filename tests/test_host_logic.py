"""CPU: host-side logic of the product — plugin surface, state-dict contract, input generator, failure modes."""
import importlib

import numpy as np
import pytest
import torch

from oracle import oracle as O

i3d = importlib.import_module("3dinfomax_b200")
syn = i3d.synthetic


def test_plugin_names_and_kwargs():
    pna = i3d.PNA(avg_d=1, device="cpu", some_unknown_kwarg=3, **O.PRETRAIN_QM9_PNA)       # train.py:208 call shape
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **O.PRETRAIN_QM9_NET3D)                # train.py:167-172
    assert sum(p.numel() for p in pna.parameters()) == 4981856
    assert sum(p.numel() for p in n3.parameters()) == 17617
    assert isinstance(i3d.NTXent(tau=0.1), torch.nn.Module)
    assert isinstance(i3d.NTXentMultiplePositives(tau=0.1), torch.nn.Module)


def test_state_dict_keys_match_the_reference_layout():
    c2, c3 = O.pna_cfg(**O.PRETRAIN_QM9_PNA), O.net3d_cfg(**O.PRETRAIN_QM9_NET3D)
    st2, st3 = O.init_pna_state(c2, 0), O.init_net3d_state(c3, 0)
    pna = i3d.PNA(**O.PRETRAIN_QM9_PNA)
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, **O.PRETRAIN_QM9_NET3D)
    assert set(pna.state_dict().keys()) == set(st2.keys())
    assert set(n3.state_dict().keys()) == set(st3.keys())
    pna.load_state_dict(st2, strict=True)
    n3.load_state_dict(st3, strict=True)
    for k, v in pna.state_dict().items():
        assert v.shape == st2[k].shape, k
    # optimizer grouping keys on the substring 'batch_norm' (trainer/self_supervised_trainer.py:79-82)
    bn = [k for k, _ in pna.named_parameters() if "batch_norm" in k]
    assert len(bn) == 2 * 22
    assert "node_gnn.mp_layers.0.pretrans.fully_connected.0.linear.weight" in dict(pna.named_parameters())


def test_fresh_init_follows_the_reference_initialisers():
    torch.manual_seed(0)
    pna = i3d.PNA(**O.PRETRAIN_QM9_PNA)
    w = pna.node_gnn.mp_layers[0].pretrans.fully_connected[0].linear.weight
    bound = (1 / 600) * np.sqrt(6 / (600 + 200))                      # xavier_uniform_(w, gain=1/in_dim)
    assert w.abs().max().item() <= bound and w.abs().max().item() > 0.9 * bound
    assert pna.node_gnn.mp_layers[0].pretrans.fully_connected[0].linear.bias.abs().max().item() == 0
    e = pna.node_gnn.atom_encoder.atom_embedding_list[0].weight
    assert e.shape == (119, 200) and e.abs().max().item() <= np.sqrt(6 / (119 + 200)) + 1e-6


def test_unsupported_configs_fail_loudly():
    bad = dict(O.PRETRAIN_QM9_PNA, aggregators=["mean", "sum"])
    with pytest.raises(NotImplementedError):
        i3d.PNA(**bad)
    with pytest.raises(NotImplementedError):
        i3d.PNA(**dict(O.PRETRAIN_QM9_PNA, dropout=0.1))
    with pytest.raises(ValueError):
        i3d.Net3D(node_dim=0, edge_dim=1, **dict(O.PRETRAIN_QM9_NET3D, reduce_func="max"))
    with pytest.raises(AssertionError):
        i3d.PNA(**dict(O.PRETRAIN_QM9_PNA, activation="not_an_activation"))
    # regularisers are supported on one process's embeddings, rejected under data parallelism (row_offset / total_rows)
    reg = i3d.NTXent(tau=0.1, variance_reg=1.0)
    with pytest.raises(NotImplementedError):
        reg._with_regularisers(torch.zeros(()), torch.zeros(4, 8), torch.zeros(4, 8), 1, 4, 16)


def test_no_cpu_fallback():
    pna = i3d.PNA(**O.PRETRAIN_QM9_PNA)
    g2, _ = i3d.batch_from_numpy(syn.make_batch(0, 2), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pna(g2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        i3d.NTXent(tau=0.1)(torch.randn(4, 8), torch.randn(4, 8))


def test_synthetic_batch_contract():
    b = syn.make_batch(7, 64)
    N, E = len(b["x_atom"]), len(b["src"])
    assert b["num_nodes"].sum() == N and b["num_edges"].sum() == E
    # both directions stored, reverse pair adjacent (datasets/qm9_dataset.py:431-435)
    assert np.array_equal(b["src"][0::2], b["dst"][1::2]) and np.array_equal(b["dst"][0::2], b["src"][1::2])
    assert np.array_equal(b["e_attr"][0::2], b["e_attr"][1::2])
    deg = np.bincount(b["dst"], minlength=N)
    assert deg.min() >= 1 and deg.max() <= 4
    for c, d in enumerate(syn.ATOM_FEATURE_DIMS):
        assert b["x_atom"][:, c].min() >= 0 and b["x_atom"][:, c].max() < d
    for c, d in enumerate(syn.BOND_FEATURE_DIMS):
        assert b["e_attr"][:, c].max() < d
    # block diagonal: no edge crosses a molecule boundary
    ptr = np.concatenate([[0], np.cumsum(b["num_nodes"])])
    mol_of = np.repeat(np.arange(64), b["num_nodes"])
    assert np.array_equal(mol_of[b["src"]], mol_of[b["dst"]])
    # complete graphs: n(n-1) edges, src = repeat_interleave(arange(n), n-1) (datasets/qm9_dataset.py:215-217)
    assert np.array_equal(b["num_edges3"], b["num_nodes3"] * (b["num_nodes3"] - 1))
    n0 = int(b["num_nodes3"][0])
    assert np.array_equal(b["src3"][:n0 * (n0 - 1)], np.repeat(np.arange(n0), n0 - 1))
    assert b["d3"].dtype == np.float32 and b["d3"].shape == (len(b["src3"]), 1) and b["d3"].min() > 0
    assert ptr[-1] == N
    again = syn.make_batch(7, 64)
    assert all(np.array_equal(b[k], again[k]) for k in b if isinstance(b[k], np.ndarray))
    c3 = syn.make_batch(1, 4, shape="qmugs", conformers=3)
    assert len(c3["num_nodes3"]) == 12 and np.array_equal(c3["num_nodes3"][0::3], c3["num_nodes"])


def test_graph_batch_mirrors_dgl_surface():
    g2, g3 = i3d.batch_from_numpy(syn.make_batch(0, 3), "cpu")
    assert g2.number_of_nodes() == int(g2.batch_num_nodes().sum()) and g2.batch_size == 3
    s, d = g2.edges()
    assert s.dtype == torch.int64 and len(s) == g2.number_of_edges() == int(g2.batch_num_edges().sum())
    assert g2.ndata["feat"].shape[1] == 9 and g2.edata["feat"].shape[1] == 3 and g3.edata["d"].shape[1] == 1
    assert g2.to("cpu").ndata["feat"].dtype == torch.int64


def test_pna_original_state_dict_keys_match_the_oracle_layout():
    """PNAOriginal builds on the CPU (no kernels run) and owns exactly the reference's state-dict keys / shapes: the
    oracle state (checked against the reference model with strict=True by oracle/pin_pna_original.py) loads strictly."""
    import importlib
    from oracle import pna_original_oracle as PO
    i3d = importlib.import_module("3dinfomax_b200")
    for over in ({}, {"hidden_dim": 200, "last_layer_dim": 200, "towers": 4, "edge_hidden_dim": 200,
                      "divide_input_first": True, "graph_norm": False}):
        c = PO.cfg(**dict(PO.CONTRASTIVE_PNA_ORIGINAL, **over))
        st = PO.init_state(c, 1)
        m = i3d.PNAOriginal(avg_d=2.0, device="cpu", **{k: v for k, v in c.items() if k != "gru"})
        res = m.load_state_dict(st, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        assert [k for k, _ in m.named_parameters()] == [k for k in st if not k.endswith(("running_mean", "running_var",
                                                                                        "num_batches_tracked"))]


def test_stats_arena_bump_allocator():
    """StatsArena host logic (runs on CPU tensors): consecutive zeroed slices, exhaustion -> None, reset re-zeroes."""
    import importlib
    import torch
    K = importlib.import_module("3dinfomax_b200.kernels")
    a = K.StatsArena("cpu", doubles=1000)
    s1, s2 = a.take(400), a.take(400)
    assert s1.data_ptr() + 400 * 8 == s2.data_ptr() and float(s1.abs().sum() + s2.abs().sum()) == 0.0
    assert a.take(400) is None                      # exhausted: callers fall back to their own buffer + memset
    s1.fill_(3.0)
    a.reset()
    assert a.used == 0 and float(a.buf.abs().sum()) == 0.0
    # statistics buffers use the library's strided layout: one accumulator per 128-byte line (I3D_STATS_STRIDE doubles)
    buf, flag = K._stats_buffer(a, 40, "cpu")
    assert flag == K.STATS_PREZEROED and buf.data_ptr() == a.buf.data_ptr() and buf.numel() == 40 * K.STATS_STRIDE
    buf2, flag2 = K._stats_buffer(None, 40, "cpu")
    assert flag2 == 0 and buf2.numel() == 40 * K.STATS_STRIDE
    assert K.StatsArena("cpu").take_ws(400, 8) is None      # no reduction scratch off the GPU


def test_graph_batch_carries_the_degree_hint():
    """max_in_degree (host int that sizes the degree plan) survives pin_memory() / to() and is computed by
    batch_from_numpy; graphs without a hint keep None (generic posttrans path)."""
    import importlib
    import numpy as np
    import torch
    i3d = importlib.import_module("3dinfomax_b200")
    b = i3d.synthetic.make_batch(3, 5)
    g2, g3 = i3d.batch_from_numpy(b, "cpu")
    assert g2.max_in_degree == int(np.bincount(b["dst"]).max()) and g3.max_in_degree is None
    assert g2.to("cpu").max_in_degree == g2.max_in_degree
    plain = i3d.GraphBatch(torch.from_numpy(b["src"]), torch.from_numpy(b["dst"]), torch.from_numpy(b["num_nodes"]))
    assert plain.max_in_degree is None
    assert i3d.annotate_max_in_degree(plain) == g2.max_in_degree and plain.max_in_degree == g2.max_in_degree


def test_degree_plan_reference_layout():
    """the numpy restatement of i3d_degree_plan used by the GPU tests: every node appears once, buckets own whole
    tiles, chunks never cross a bucket"""
    import numpy as np
    from gpu_cases_plan_ref import degree_plan_ref
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 5, size=1000)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    perm, tile_bucket, chunk, over = degree_plan_ref(rowptr, 5, 4)
    assert over == 0 and sorted(perm[perm >= 0].tolist()) == list(range(1000))
    for t, bk in enumerate(tile_bucket):
        rows = perm[t * 128:(t + 1) * 128]
        rows = rows[rows >= 0]
        if bk < 0:
            assert len(rows) == 0
        else:
            assert (deg[rows] == bk).all()
    for r0, n, bk in chunk.reshape(-1, 3):
        if n:
            assert r0 % 128 == 0 and n % 128 == 0 and n <= 4 * 128
            assert set(tile_bucket[r0 // 128:(r0 + n) // 128].tolist()) == {bk}


def test_fold_batch_norm_algebra():
    """inference.fold_batch_norm: eval-mode FCLayer stacks (Linear -> activation -> BatchNorm,
    models/base_layers.py:100-111) with the BatchNorm affine folded into the layer's own Linear (activation 'none') or
    into the next Linear (behind an activation) compute the same function.  Plain torch on the CPU: the parameters of the
    product modules, evaluated with torch ops (the kernels need a GPU; tests/gpu_cases_bucketed.py::case_inference runs
    the folded modules themselves)."""
    import importlib
    import torch
    import torch.nn.functional as F
    i3d = importlib.import_module("3dinfomax_b200")
    inf = importlib.import_module("3dinfomax_b200.inference")
    bl = importlib.import_module("3dinfomax_b200.base_layers")
    torch.manual_seed(0)

    def torch_eval(mlp, x):
        for fc in mlp.fully_connected:
            x = F.linear(x, fc.linear.weight, fc.linear.bias)
            x = {0: lambda t: t, 1: torch.relu, 2: F.silu, 3: lambda t: F.leaky_relu(t, 0.01)}[fc.act](x)
            if fc.batch_norm is not None:
                b = fc.batch_norm
                x = F.batch_norm(x, b.running_mean, b.running_var, b.weight, b.bias, False, 0.0, b.eps)
        return x

    for kw, kept in ((dict(in_dim=24, hidden_size=16, out_dim=16, layers=2, mid_activation="relu", last_activation="none",
                           mid_batch_norm=True, last_batch_norm=True), 0),      # PNA pretrans: 0 -> 1, 1 -> itself
                     (dict(in_dim=24, out_dim=16, layers=1, last_activation="none", last_batch_norm=True), 0),  # posttrans
                     (dict(in_dim=24, hidden_size=16, out_dim=8, layers=2, mid_activation="relu", mid_batch_norm=True), 0),
                     (dict(in_dim=24, out_dim=16, layers=1, last_activation="SiLU", last_batch_norm=True), 1)):  # kept
        m = bl.MLP(**kw).double()
        with torch.no_grad():
            for fc in m.fully_connected:
                fc.linear.weight.normal_(0, 0.3), fc.linear.bias.normal_(0, 0.3)
                if fc.batch_norm is not None:
                    b = fc.batch_norm
                    b.weight.uniform_(0.5, 1.5), b.bias.normal_(0, 0.3), b.running_mean.normal_(0, 0.5)
                    b.running_var.uniform_(0.1, 2.0)
        x = torch.randn(50, 24, dtype=torch.float64)
        want = torch_eval(m, x)
        n_bn = sum(fc.batch_norm is not None for fc in m.fully_connected)
        folded = inf.fold_batch_norm(m)
        assert folded.folded_batch_norms == n_bn - kept
        assert sum(fc.batch_norm is not None for fc in folded.fully_connected) == kept
        got = torch_eval(folded, x)
        assert float((got - want).abs().max()) < 1e-12 * float(want.abs().max() + 1)
        assert all(fc.batch_norm is not None for fc in m.fully_connected if True) == (n_bn == len(m.fully_connected))
    # the shipped PNA: every BatchNorm folds away
    cfg = importlib.import_module("3dinfomax_b200.configs")
    pna = i3d.PNA(avg_d=1, device="cpu", **cfg.PRETRAIN_QM9_MODEL_PARAMETERS)
    f = inf.fold_batch_norm(pna)
    assert f.folded_batch_norms == 22 and not any("batch_norm" in k for k in f.state_dict())
    assert any("batch_norm" in k for k in pna.state_dict())          # the original is untouched


def test_supervised_trainer_parameter_groups_follow_the_reference():
    """trainer/trainer.py:216-238: BatchNorm (no weight decay) / new / transferred (own lr) / frozen (lr 0), in this order;
    transfer_layers ['gnn'] with exclude_from_transfer ['batch_norm'] as in configs_clean/tune_QM9_homo.yml"""
    cfg = importlib.import_module("3dinfomax_b200.configs")
    T = importlib.import_module("3dinfomax_b200.trainer")
    model = i3d.PNA(avg_d=1, **cfg.TUNE_QM9_HOMO_MODEL_PARAMETERS)
    tr = object.__new__(T.Trainer)                       # the grouping is host logic; the constructor needs a GPU
    tr.model, tr.model3d = model, None
    tr.transfer_layers, tr.exclude_from_transfer = ("gnn",), ("batch_norm",)
    tr.frozen_layers, tr.transferred_lr = ("output.fully_connected.1",), 1e-5
    named = list(model.named_parameters())
    groups = tr.param_groups(named, {"lr": 7e-5, "weight_decay": 1e-11})
    ids = lambda ps: {id(p) for p in ps}
    by_name = dict(named)
    assert [sorted(k for k in g if k != "params") for g in groups] == [["weight_decay"], [], ["lr"], ["lr"]]
    bn, new, transferred, frozen = (ids(g["params"]) for g in groups)
    assert groups[0]["weight_decay"] == 0 and groups[2]["lr"] == 1e-5 and groups[3]["lr"] == 0
    for k, p in by_name.items():
        in_gnn = "gnn" in k
        is_bn = "batch_norm" in k
        is_frozen = "output.fully_connected.1" in k
        want = {"transferred": in_gnn and not is_bn, "frozen": is_frozen,
                "bn": is_bn and not is_frozen, "new": not (in_gnn and not is_bn) and not is_bn and not is_frozen}
        assert (id(p) in transferred) == want["transferred"], k
        assert (id(p) in frozen) == want["frozen"], k
        assert (id(p) in bn) == want["bn"], k
        assert (id(p) in new) == want["new"], k
    # every parameter is optimised exactly once
    assert sum(len(g["params"]) for g in groups) == len(named)
    # the contrastive trainer keeps the two groups of trainer/self_supervised_trainer.py:78-86
    ss = object.__new__(T.SelfSupervisedTrainer)
    g2 = ss.param_groups(named, {"lr": 8e-5})
    assert len(g2) == 2 and g2[0]["weight_decay"] == 0 and ids(g2[0]["params"]) == {id(p) for k, p in named if "batch_norm" in k}
