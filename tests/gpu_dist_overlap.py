"""Data-parallel training steps with the per-layer gradient all-reduce overlapped with backward (FusedAdam
.set_overlap_groups) against the same steps with ONE all-reduce after backward (I3D_AR_OVERLAP=0) — run on the GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/gpu_dist_overlap.py

The all-reduced gradients of the first step must agree to fp32 summation-order noise (the split-K weight-gradient GEMMs
accumulate with atomics, so not even two runs of the SAME schedule are bit-identical: relative L2 <= 1e-5), and the loss
trajectories of 4 steps must agree to 2e-4 (Adam turns noise-level gradients into +-lr steps, tests/gpu_cases.py::
case_train_steps).  Both the eager trainer and the shape-bucketed captured step (NCCL inside the CUDA graphs) are run."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (seeded weights only)


def run(overlap, bucketed, rank, world, dev, store_np, batches, nvls=False):
    os.environ["I3D_AR_OVERLAP"] = "1" if overlap else "0"
    os.environ["I3D_NVLS_ADAM"] = "1" if nvls else "0"
    i3d = importlib.import_module("3dinfomax_b200")
    cfg = importlib.import_module("3dinfomax_b200.configs")
    c2, c3 = O.pna_cfg(**cfg.PRETRAIN_QM9_MODEL_PARAMETERS), O.net3d_cfg(**cfg.PRETRAIN_QM9_MODEL3D_PARAMETERS)
    st2, st3 = O.init_pna_state(c2, 5, True), O.init_net3d_state(c3, 6, True)
    pna = i3d.PNA(avg_d=1, device=dev, **cfg.PRETRAIN_QM9_MODEL_PARAMETERS)
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **cfg.PRETRAIN_QM9_MODEL3D_PARAMETERS)
    pna.load_state_dict(st2), n3.load_state_dict(st3)
    tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=0.1), dev, {"lr": 8e-5}, process_group=dist.group.WORLD,
                                   graph_safe=bucketed)
    store = i3d.PackedMoleculeStore(store_np, dev)
    groups = len(tr.optim._overlap)
    losses, grads = [], None
    runner = i3d.BucketedStep(tr, store, dp_levels=3) if bucketed else None
    for ix in batches:
        if bucketed:
            losses.append(float(runner.step(ix).item()))
        else:
            g2, g3 = store.collate(ix)
            loss, _, _ = tr.forward_pass(([g2], [g3]))
            loss.backward()
            tr.optim.step()
            losses.append(float(loss.item()))
        if grads is None:
            # Adam's first moment after the first step = (1 - beta1) x the all-reduced gradient: linear in the reduced
            # gradient and available in every mode (the fused NVLS step never materialises the reduced gradient)
            torch.cuda.synchronize()
            grads = torch.cat([fl["m"].clone() for fl in tr.optim._flat if fl is not None])
        if not bucketed:
            tr.optim.zero_grad()
    torch.cuda.synchronize()
    # replicas must stay identical: every rank's parameters equal rank 0's
    flat = torch.cat([fl["p"] for fl in tr.optim._flat if fl is not None]).clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(flat, ref))
    return grads, losses, (groups, same, tr.optim.nvls)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    i3d = importlib.import_module("3dinfomax_b200")
    store_np = i3d.synthetic.make_store(21, 1500)
    rng = np.random.default_rng(100 + rank)
    batches = [rng.choice(1500, size=256, replace=False) for _ in range(4)]
    ok = True
    for bucketed in (False, True):
        p0, l0, (g0, same0, _) = run(False, bucketed, rank, world, dev, store_np, batches)       # one NCCL all-reduce
        for tag, overlap, nvls in (("NCCL all-reduce per layer, overlapped with backward", True, False),
                                   ("fused NVLS step (multimem reduce + Adam + multicast store)", False, True)):
            p1, l1, (g1, same1, is_nvls) = run(overlap, bucketed, rank, world, dev, store_np, batches, nvls)
            diff = float((p1 - p0).double().norm() / p0.double().norm())
            dl = max(abs(a - b) for a, b in zip(l1, l0))
            good = diff <= 1e-5 and dl <= 2e-4 and same0 and same1 and all(np.isfinite(l1))
            good = good and ((g1 >= 1 and g0 == 0) if overlap else True)
            ok = ok and good
            if rank == 0:
                print("%s | %s%s: reduced gradient of step 1 (via Adam's exp_avg), relative L2 difference to the single "
                      "all-reduce %.3e (tol 1e-5) | losses %s vs %s (max diff %.1e, tol 2e-4) | replicas identical %s | %s"
                      % ("bucketed/captured" if bucketed else "eager", tag,
                         "" if not nvls else (" [active]" if is_nvls else " [NOT AVAILABLE: NCCL fallback ran]"), diff,
                         ["%.5f" % x for x in l1], ["%.5f" % x for x in l0], dl, same1, "ok" if good else "FAIL"),
                      flush=True)
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("DIST OVERLAP CHECK:", "PASS" if int(t.item()) == 0 else "FAIL", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
