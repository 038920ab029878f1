#!/usr/bin/env python
"""Where PackedMoleculeStore.collate spends its time: host (cProfile) and device (CUDA events around the two kernels)."""
import cProfile
import importlib
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
i3d = importlib.import_module("3dinfomax_b200")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    store = i3d.PackedMoleculeStore(i3d.synthetic.make_store(7, 4 * B), dev)
    rng = np.random.default_rng(11)
    idxs = [rng.integers(0, len(store), size=B) for _ in range(60)]
    for ix in idxs[:5]:
        store.collate(ix)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for ix in idxs[5:55]:
        store.collate(ix)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    g = torch.cuda.CUDAGraph()
    ix = idxs[55]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        store.collate(ix)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g2, g3 = store.collate(ix)
    e1.record()
    torch.cuda.synchronize()
    print("one collate, events around the call (host-bound upper bound of device time): %.1f us" % (e0.elapsed_time(e1) * 1e3))


if __name__ == "__main__":
    main()
