#!/usr/bin/env python
"""In-situ per-kernel GPU time of the pre-training step (warm caches, real launch order) via torch.profiler (CUPTI).

    python tools/step_profile.py [--batch 512] [--steps 5] [--mode eager|graph] > gpurun_out/step_profile.txt
    ncu --nvtx --nvtx-include "timed_step/" ... python tools/step_profile.py --nvtx     # one eager step inside an NVTX range

ncu serialises launches and flushes caches before each one, so its per-launch times overstate kernels whose inputs
are L2-resident in the real step; this table is what the step actually spends.  Not a bench: no number printed here
is a throughput claim.
"""
import argparse
import importlib
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--mode", default="eager", choices=["eager", "graph"])
    ap.add_argument("--nvtx", action="store_true", help="no CUPTI profile: one step inside the NVTX range 'timed_step' (for ncu)")
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    i3d = importlib.import_module("3dinfomax_b200")
    cfg = importlib.import_module("3dinfomax_b200.configs")
    dev = torch.device("cuda", 0)
    torch.manual_seed(123)
    pna = i3d.PNA(avg_d=1, device=dev, **cfg.PRETRAIN_QM9_MODEL_PARAMETERS)
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **cfg.PRETRAIN_QM9_MODEL3D_PARAMETERS)
    graph = args.mode == "graph"
    # the step bench.py times: device collate padded to a shape bucket -> both encoders -> loss -> backward -> Adam,
    # here launched EAGERLY (the kernel sequence BucketedStep captures) so that CUPTI attributes every kernel
    import numpy as np
    from importlib import import_module
    tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=0.1), dev, {"lr": 8e-5}, graph_safe=True)
    store = i3d.PackedMoleculeStore(i3d.synthetic.make_store(7, 4 * args.batch), dev)
    run = i3d.BucketedStep(tr, store)
    rng = np.random.default_rng(3)
    idxs = [rng.permutation(4 * args.batch)[:args.batch] for _ in range(8)]
    tmod = import_module("3dinfomax_b200.trainer")
    cmod = import_module("3dinfomax_b200.collate")
    bk = tmod._Bucket()
    lad = run.ladder(args.batch)
    bk.caps = lad.caps(max(lad.level_of(store.batch_sizes(ix)) for ix in idxs))
    bk.meta = torch.zeros(cmod.metadata_len(args.batch), dtype=torch.int64, device=dev)

    def step(i):
        if graph:
            return run.step(idxs[i % 8])
        store.stage_metadata(idxs[i % 8], dev_out=bk.meta)
        run._body(bk, args.batch, True)
        return run.loss

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    if args.nvtx:
        torch.cuda.nvtx.range_push("timed_step")
        step(3)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            step(i)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    t_min, t_max = None, None
    for ev in prof.events():
        if str(getattr(ev, "device_type", "")).endswith("CUDA") and ev.device_time_total > 0:
            name = re.sub(r"\(.*$", "", ev.name)[:100]
            agg[name][0] += 1
            agg[name][1] += ev.device_time_total
            tr_ = ev.time_range
            t_min = tr_.start if t_min is None else min(t_min, tr_.start)
            t_max = tr_.end if t_max is None else max(t_max, tr_.end)
    total = sum(v[1] for v in agg.values())
    span = (t_max - t_min) if t_min is not None else 0.0
    print("torch.profiler (CUPTI), %s mode, batch %d, %d steps: sum of kernel time %.1f us/step, GPU span %.1f us/step"
          % (args.mode, args.batch, args.steps, total / args.steps, span / args.steps))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6.2f%% %6.1f x/step %8.1f us  %9.1f us/step  %s" % (100 * us / total, n / args.steps, us / n,
                                                                     us / args.steps, name))


if __name__ == "__main__":
    main()
