"""TEST INFRASTRUCTURE — container-only: golden vectors for BASELINE config 5 (configs_clean/tune_QM9_homo.yml):
the reference's own PNA (under the dgl shim) with the fine-tuning head — readout [min, max, mean, sum], target_dim 1,
batch_norm_momentum 0.1 — and ``loss_func: L1Loss`` on a seeded batch with seeded targets.  Asserts that
oracle/oracle.py reproduces it (eval exact, train <= 1e-6, gradients <= 1e-5 of scale) and writes
tests/golden/finetune_homo_b12.npz.

    python -m oracle.pin_finetune
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import ref_under_shim as R  # noqa: E402
from oracle.make_golden import grad_fingerprint  # noqa: E402

# configs_clean/tune_QM9_homo.yml:47-77
TUNE_QM9_HOMO = dict(target_dim=1, hidden_dim=200, mid_batch_norm=True, last_batch_norm=True, readout_batchnorm=True,
                     batch_norm_momentum=0.1, readout_hidden_dim=200, readout_layers=2, dropout=0.0,
                     propagation_depth=7, aggregators=["mean", "max", "min", "std"],
                     scalers=["identity", "amplification", "attenuation"],
                     readout_aggregators=["min", "max", "mean", "sum"], pretrans_layers=2, posttrans_layers=1,
                     residual=True)
CASE = ("finetune_homo_b12", 31, 12, 401)          # name, batch seed, B, weight seed


def targets(B):
    return torch.randn(B, 1, generator=torch.Generator().manual_seed(77))


def main():
    syn = importlib.import_module("3dinfomax_b200.synthetic")
    ref = R.load_reference()
    name, bseed, B, wseed = CASE
    b = syn.make_batch(bseed, B)
    g2, xa, ea, _, _ = O.graphs_from_batch(b)
    c = O.pna_cfg(**TUNE_QM9_HOMO)
    st = O.init_pna_state(c, wseed, True)
    y = targets(B)
    out = {}
    for mode in ("eval", "train"):
        training = mode == "train"
        m = ref.PNA(avg_d=1, device="cpu", **TUNE_QM9_HOMO)
        m.load_state_dict(st)
        m.train(training)
        G = ref.ShimGraph(g2.src, g2.dst, g2.n, g2.bnn, g2.bne)
        G.ndata["feat"], G.edata["feat"] = xa.clone(), ea.clone()
        z = m(G)
        loss = torch.nn.L1Loss()(z, y)                                       # train.py: globals()['L1Loss']()
        o = O.as_leaf_params(st)
        oz = O.pna_forward(o, c, g2, xa, ea, training)
        oloss = torch.nn.functional.l1_loss(oz, y)
        tol = 0.0 if not training else 1e-6
        assert (oz - z).abs().max().item() <= tol and abs(oloss.item() - loss.item()) <= 1e-6, (mode,)
        out["z_" + mode] = z.detach().numpy()
        out["loss_" + mode] = np.float32(loss.item())
        if training:
            loss.backward()
            oloss.backward()
            named = dict(m.named_parameters())
            scale = max(float(p.grad.abs().max()) for p in named.values())
            keys, fps = [], []
            for k, p in named.items():
                assert (o[k].grad - p.grad).abs().max().item() <= 1e-5 * scale, k
                keys.append(k)
                fps.append(grad_fingerprint(p.grad))
            out["grad_keys"], out["grad_fp"], out["grad_scale"] = np.array(keys), np.stack(fps), np.float64(scale)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print("pinned %s: z %s, L1 loss eval %.6f train %.6f — oracle == reference" %
          (name, out["z_eval"].shape, out["loss_eval"], out["loss_train"]))


if __name__ == "__main__":
    main()
