"""TEST INFRASTRUCTURE — container-only: pins oracle/pna_original_oracle.py on the reference's own PNAOriginal.

Loads models/pna_original.py (with commons/mol_encoder.py and models/base_layers.py) UNMODIFIED from /root/reference
under the dgl shim of oracle/ref_under_shim.py, loads the oracle's seeded state dict with strict=True (same keys and
shapes), runs the reference model on a seeded batch in eval and train mode and asserts equality with the oracle
(forward exact in eval, <= 2e-6 in train; parameter gradients <= 1e-5 of the gradient scale).  Writes
tests/golden/pna_original_*.npz (outputs, gradient fingerprints) for the GPU-box tests.

    python -m oracle.pin_pna_original
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import pna_original_oracle as PO  # noqa: E402
from oracle import ref_under_shim as R  # noqa: E402
from oracle.make_golden import grad_fingerprint  # noqa: E402

# name: (batch seed, B, shape, weight seed, avg_d, config)
CASES = {
    "pna_original_towers5": (21, 6, "qm9", 301, 2.3, PO.cfg(**PO.CONTRASTIVE_PNA_ORIGINAL)),
    # the shape BASELINE.json configs[1] names: hidden 200, 4 towers, 4 layers (inputs divided between the towers)
    "pna_original_h200_t4": (22, 4, "qm9", 302, 1.0, PO.cfg(**dict(PO.CONTRASTIVE_PNA_ORIGINAL, hidden_dim=200,
                                                                   last_layer_dim=200, towers=4, edge_hidden_dim=200,
                                                                   divide_input_first=True, graph_norm=False))),
}


def load_reference_pna_original():
    dgl = types.ModuleType("dgl")
    dgl.DGLGraph = R.ShimGraph
    dgl.readout_nodes = R._readout_nodes
    fn = types.ModuleType("dgl.function")
    fn.sum = lambda msg, out: ("sum", msg, out)
    fn.mean = lambda msg, out: ("mean", msg, out)
    dgl.function = fn
    ogb, ogbu, ogbf = types.ModuleType("ogb"), types.ModuleType("ogb.utils"), types.ModuleType("ogb.utils.features")
    ogbf.get_atom_feature_dims = lambda: list(O.ATOM_DIMS)
    ogbf.get_bond_feature_dims = lambda: list(O.BOND_DIMS)
    pk_models, pk_commons = types.ModuleType("models"), types.ModuleType("commons")
    pk_models.__path__, pk_commons.__path__ = [], []
    stubs = {"dgl": dgl, "dgl.function": fn, "ogb": ogb, "ogb.utils": ogbu, "ogb.utils.features": ogbf,
             "models": pk_models, "commons": pk_commons}
    names = list(stubs) + ["commons.mol_encoder", "models.base_layers", "models.pna_original"]
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(stubs)
    try:
        mods = {}
        for name, rel in [("commons.mol_encoder", "commons/mol_encoder.py"), ("models.base_layers", "models/base_layers.py"),
                          ("models.pna_original", "models/pna_original.py")]:
            spec = importlib.util.spec_from_file_location(name, os.path.join(R.REF, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m
            spec.loader.exec_module(m)
            mods[name] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods["models.pna_original"].PNAOriginal


def snorm(num_nodes):
    """graph_collate's snorm_n (datasets/custom_collate.py:96-98): sqrt(1/n) repeated per node, [N, 1]."""
    n = torch.as_tensor(num_nodes).long()
    return torch.repeat_interleave((1.0 / n.float()).sqrt(), n)[:, None]


def main():
    import importlib
    syn = importlib.import_module("3dinfomax_b200.synthetic")
    PNAOriginal = load_reference_pna_original()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (bseed, B, shape, wseed, avg_d, c) in CASES.items():
        b = syn.make_batch(bseed, B, shape=shape)
        g2, xa, ea, _, _ = O.graphs_from_batch(b)
        sn = snorm(b["num_nodes"])
        st = PO.init_state(c, wseed)
        out = {}
        for mode in ("eval", "train"):
            training = mode == "train"
            kw = {k: v for k, v in c.items() if k not in ("gru",)}
            m = PNAOriginal(avg_d=avg_d, device="cpu", **kw)
            m.load_state_dict(st, strict=True)
            m.train(training)
            G = R.ShimGraph(g2.src, g2.dst, g2.n, g2.bnn, g2.bne)
            G.ndata["feat"], G.edata["feat"] = xa.clone(), ea.clone()
            z = m(G, sn)
            o = O.as_leaf_params(st)
            oz = PO.forward(o, c, g2, xa, ea, sn, avg_d, training)
            tol = 0.0 if not training else 2e-6 * float(z.abs().max())
            assert (oz - z).abs().max().item() <= tol, (name, mode, (oz - z).abs().max().item())
            out["z_" + mode] = z.detach().numpy()
            if training:
                w = torch.randn(z.shape, generator=torch.Generator().manual_seed(5))
                (z * w).sum().backward()
                (oz * w).sum().backward()
                named = dict(m.named_parameters())
                scale = max(float(p.grad.abs().max()) for p in named.values() if p.grad is not None)
                keys, fps = [], []
                for k, p in named.items():
                    if p.grad is None:                       # node_gnn.MLP_layer.* is never used in forward
                        assert o[k].grad is None, k
                        continue
                    assert (o[k].grad - p.grad).abs().max().item() <= 1e-5 * scale, (name, k)
                    keys.append(k)
                    fps.append(grad_fingerprint(p.grad))
                out["grad_keys"], out["grad_fp"], out["grad_scale"] = np.array(keys), np.stack(fps), np.float64(scale)
                sd = m.state_dict()
                for k in sd:
                    if k.endswith("running_mean") or k.endswith("running_var"):
                        if ".layers.0.towers.0." in k:
                            out["buf/" + k] = sd[k].numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
        print("pinned %s: B=%d N=%d, z %s — oracle == reference (eval exact, train <= 2e-6 rel, grads <= 1e-5)"
              % (name, B, g2.n, tuple(out["z_eval"].shape)))


if __name__ == "__main__":
    main()
