#!/usr/bin/env python
"""Warm per-launch time of the weight-gradient (TN) GEMM kernels over K, for both tensor-core backends
(0 = transposing generic kernel, 2 = MN-major warp-specialised kernel).  Tuning aid, not a bench.

    python tools/tn_sweep.py [M N]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
i3d = importlib.import_module("3dinfomax_b200")
K = importlib.import_module("3dinfomax_b200.kernels")
from gemm_bench import graph_time  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    L = i3d.lib.load()
    dev = torch.device("cuda", 0)
    for Kd in (4736, 9472, 18944, 37888, 75776):
        A, B = torch.randn(Kd, M, device=dev), torch.randn(Kd, N, device=dev)
        C = torch.zeros(M, N, device=dev)
        row = []
        for backend in (0, 2):
            L.i3d_gemm_backend(backend)
            row.append(graph_time(lambda: K.gemm(K.TN, M, N, [{"A": A, "B": B, "K": Kd}], C, accumulate=True)))
        L.i3d_gemm_backend(0)
        print("M %d N %d K %6d: generic %.2f us   mn-major ws %.2f us" % (M, N, Kd, row[0], row[1]), flush=True)


if __name__ == "__main__":
    main()
