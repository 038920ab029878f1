"""Where does the fp32 distance between the CUDA path and the float64 oracle come from?  One B = 512 train-mode step,
gradients of every parameter against the float64 oracle, for the GEMM backends (tensor-core 3xTF32 / fp32 SIMT).
usage (GPU box): python tools/accuracy_probe.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_cases_bucketed as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

i3d = G.i3d
DEV = "cuda"


def main():
    lr = 8e-5
    B, M = 512, 1200
    store = G.syn.make_store(55, M, "qm9")
    idx = np.random.default_rng(12).integers(0, M, size=B)
    b = G._ref_batch(store, idx, 1)
    c2, c3 = O.pna_cfg(**O.PRETRAIN_QM9_PNA), O.net3d_cfg(**O.PRETRAIN_QM9_NET3D)
    st2, st3 = O.init_pna_state(c2, 71, True), O.init_net3d_state(c3, 72, True)
    tl, tz2, tz3, tg = G._fp64_truth(c2, c3, st2, st3, "NTXent", b, lr)
    otr = O.OracleTrainer(c2, c3, st2, st3, loss="NTXent", tau=0.1, lr=lr)
    ol, oz2, oz3, og = G._oracle_step(otr, b)
    scale = max(float(g.abs().max()) for g in tg.values())
    print("cpu fp32 oracle: z2d %.2e grads %.2e" % (G.rel(oz2, tz2), G._grad_errors(og, tg)[0]))
    L = i3d.lib.load()
    for name, backend, env in (("tensor-core 3xTF32 (merged)", 0, {}), ("tensor-core 3xTF32 (generic 13F)", 0, {"I3D_POSTTRANS": "generic"}),
                               ("fp32 SIMT (merged->generic)", 1, {})):
        os.environ.pop("I3D_POSTTRANS", None)
        os.environ.update(env)
        L.i3d_gemm_backend(backend)
        pna = i3d.PNA(avg_d=1, device=DEV, **O.PRETRAIN_QM9_PNA)
        n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **O.PRETRAIN_QM9_NET3D)
        pna.load_state_dict(st2), n3.load_state_dict(st3)
        tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=0.1), DEV, {"lr": lr})
        g2, g3 = i3d.batch_from_numpy(b, DEV)
        try:
            l, z2, z3 = tr.forward_pass(([g2], [g3]))
            l.backward()
            tr.optim.step()
        except Exception as e:
            print(name, "failed:", str(e)[:200])
            continue
        torch.cuda.synchronize()
        named = G._named(pna, n3)
        pg = tr.optim.packed_grads()
        grads = {k: pg[p] for k, p in named.items()}
        g, t, tn = G._grad_errors(grads, tg)
        print("%-36s z2d %.2e z3d %.2e grads(global) %.2e grads(per tensor) %.2e %s" % (name, G.rel(z2, tz2), G.rel(z3, tz3), g, t, tn))
        rows = []
        for k, ref in tg.items():
            rows.append((float((grads[k].cpu().double() - ref).abs().max()) / scale, k))
        rows.sort(reverse=True)
        for e, k in rows[:5]:
            print("      %.2e %s" % (e, k))
    L.i3d_gemm_backend(0)


if __name__ == "__main__":
    main()
