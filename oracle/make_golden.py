"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the reference's OWN modules.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden

For each case: a seeded synthetic batch (3dinfomax_b200/synthetic.py) and seeded trained-scale weights
(oracle.init_*_state) are pushed through the unmodified reference PNA / Net3D / NTXent classes running under
oracle/ref_under_shim.py; outputs, loss, a compact fingerprint of every parameter gradient, BN running statistics
and the CSR arrays are stored.  While doing so it asserts that oracle/oracle.py reproduces the reference
(forward: exact in eval, <=1e-6 in train; gradients: <=1e-5 of the global gradient scale), which is what
"parity pinned" in the oracle header refers to.  tests/test_oracle_golden.py re-checks the oracle against these
files on any box; tests/test_gpu_parity.py checks the CUDA path against them on the B200.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import ref_under_shim as R  # noqa: E402

CASES = {
    # name: (batch seed, B, shape, conformers, loss, weight seeds (2d, 3d))
    "qm9_b8": (11, 8, "qm9", 1, "NTXent", (101, 202)),
    "qmugs_b6_c3": (12, 6, "qmugs", 3, "NTXentMultiplePositives", (103, 204)),
}
TAU = 0.1
GRAD_SAMPLE = 48


def grad_fingerprint(g):
    """(l2 norm, sum, first GRAD_SAMPLE entries of a fixed stride sample) — compact but sensitive."""
    f = g.detach().reshape(-1).double()
    stride = max(1, f.numel() // GRAD_SAMPLE)
    sample = f[::stride][:GRAD_SAMPLE]
    pad = torch.zeros(GRAD_SAMPLE, dtype=torch.float64)
    pad[:sample.numel()] = sample
    return np.concatenate([[f.norm().item(), f.sum().item()], pad.numpy()])


def run_case(name, ref, syn):
    bseed, B, shape, C, loss_name, (s2, s3) = CASES[name]
    b = syn.make_batch(bseed, B, shape=shape, conformers=C)
    g2, xa, ea, g3, d3 = O.graphs_from_batch(b)
    c2, c3 = O.pna_cfg(**O.PRETRAIN_QM9_PNA), O.net3d_cfg(**O.PRETRAIN_QM9_NET3D)
    st2, st3 = O.init_pna_state(c2, s2, True), O.init_net3d_state(c3, s3, True)
    out = {}

    def ref_graphs():
        G2 = ref.ShimGraph(g2.src, g2.dst, g2.n, g2.bnn, g2.bne)
        G2.ndata["feat"], G2.edata["feat"] = xa.clone(), ea.clone()
        G3 = ref.ShimGraph(g3.src, g3.dst, g3.n, g3.bnn, g3.bne)
        G3.edata["d"] = d3.clone()
        return G2, G3

    ref_loss = getattr(ref, loss_name)(tau=TAU)
    for mode in ("eval", "train"):
        training = mode == "train"
        m2 = ref.PNA(avg_d=1, device="cpu", **O.PRETRAIN_QM9_PNA)
        m3 = ref.Net3D(node_dim=0, edge_dim=1, avg_d=1, **O.PRETRAIN_QM9_NET3D)
        m2.load_state_dict(st2), m3.load_state_dict(st3)
        m2.train(training), m3.train(training)
        G2, G3 = ref_graphs()
        z2, z3 = m2(G2), m3(G3)
        loss = ref_loss(z2, z3)
        o2, o3 = O.as_leaf_params(st2), O.as_leaf_params(st3)
        taps = {}
        oz2 = O.pna_forward(o2, c2, g2, xa, ea, training, taps)
        oz3 = O.net3d_forward(o3, c3, g3, d3, training)
        oloss = O.LOSSES[loss_name](oz2, oz3, tau=TAU)
        tol = 0.0 if not training else 1e-6
        assert (oz2 - z2).abs().max().item() <= tol and (oz3 - z3).abs().max().item() <= tol, (name, mode)
        assert abs(oloss.item() - loss.item()) <= 1e-6, (name, mode)
        out["z2d_" + mode] = z2.detach().numpy()
        out["z3d_" + mode] = z3.detach().numpy()
        out["loss_" + mode] = np.float32(loss.item())
        if training:
            loss.backward()
            oloss.backward()
            named = [("2d." + k, p) for k, p in m2.named_parameters()] + [("3d." + k, p) for k, p in m3.named_parameters()]
            scale = max(p.grad.abs().max().item() for _, p in named)
            keys, fps = [], []
            for k, p in named:
                og = (o2 if k.startswith("2d.") else o3)[k[3:]].grad
                assert (og - p.grad).abs().max().item() <= 1e-5 * scale, (name, k)
                keys.append(k)
                fps.append(grad_fingerprint(p.grad))
            out["grad_keys"] = np.array(keys)
            out["grad_fp"] = np.stack(fps)
            out["grad_scale"] = np.float64(scale)
            for k in ("node_gnn.mp_layers.0.pretrans.fully_connected.1.linear.weight",
                      "node_gnn.mp_layers.6.posttrans.fully_connected.0.batch_norm.weight",
                      "output.fully_connected.1.linear.bias"):
                out["grad2d/" + k] = dict(m2.named_parameters())[k].grad.numpy()
            for k, p in m3.named_parameters():
                out["grad3d/" + k] = p.grad.numpy()
            sd2, sd3 = m2.state_dict(), m3.state_dict()
            for k in ("node_gnn.mp_layers.0.pretrans.fully_connected.0.batch_norm.running_mean",
                      "node_gnn.mp_layers.0.pretrans.fully_connected.0.batch_norm.running_var",
                      "node_gnn.mp_layers.6.posttrans.fully_connected.0.batch_norm.running_var"):
                out["buf2d/" + k] = sd2[k].numpy()
            for k in sd3:
                if "running" in k:
                    out["buf3d/" + k] = sd3[k].numpy()
            # aggregation of layer 0 as the reference's reduce_func leaves it: [N, 12F]
            out["agg0_head"] = taps["agg0"][:48].detach().numpy()
            out["msg0_head"] = taps["msg0"][:64].detach().numpy()
    # integer structure: argsort(dst, stable) as DGL's mailbox order
    rowptr, col, eid = O.csr_reference(b["src"], b["dst"], g2.n)
    out["csr_rowptr"], out["csr_col"], out["csr_eid"] = rowptr, col, eid
    out["meta"] = np.array([bseed, B, C, s2, s3], dtype=np.int64)
    return out


def main():
    ref = R.load_reference()
    syn = importlib.import_module("3dinfomax_b200.synthetic")
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in CASES:
        torch.manual_seed(0)
        out = run_case(name, ref, syn)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, "%.1f kB" % (os.path.getsize(path) / 1e3), "loss_train", out["loss_train"])


if __name__ == "__main__":
    main()
