"""Timing of the column-statistics kernels (bn_bwd_reduce / bn_bwd_apply / bn_apply / edge_gather_add) the way they run in
the step: 40 launches per CUDA graph, 20 replays, rotating over 8 buffer sets (cold) or one (warm).
usage (GPU box): [I3D_COL_CTAS_PER_SM=n] python tools/col_bench.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
i3d = importlib.import_module("3dinfomax_b200")
K = i3d.kernels
dev = torch.device("cuda", 0)


def graph_time(fn, reps=20, inner=40):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for i in range(3):
            fn(i)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(inner):
            fn(i)
    gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * inner) * 1e3


def main():
    F = 200
    for M in (19072, 9216):
        RP = 8
        Y = [torch.randn(M, F, device=dev) for _ in range(RP)]
        dO = [torch.randn(M, F, device=dev) for _ in range(RP)]
        save = torch.cat([torch.zeros(F), torch.ones(F)]).to(dev)
        gamma = torch.ones(F, device=dev)
        sums = torch.zeros(2 * F, dtype=torch.float64, device=dev)
        rm, rv = torch.zeros(F, device=dev), torch.ones(F, device=dev)
        nbt = torch.zeros((), dtype=torch.long, device=dev)
        s2 = K.bn_bwd_reduce(dO[0], Y[0], 1, save)
        res = {}
        for tag, rot in (("cold", True), ("warm", False)):
            pick = (lambda i: i % RP) if rot else (lambda i: 0)
            res["bn_bwd_reduce/" + tag] = graph_time(lambda i: K.bn_bwd_reduce(dO[pick(i)], Y[pick(i)], 1, save))
            res["bn_bwd_apply/" + tag] = graph_time(lambda i: K.bn_bwd_apply(dO[pick(i)], Y[pick(i)], 1, True, True, save, gamma, s2))
            res["bn_apply/" + tag] = graph_time(lambda i: K.bn_apply(Y[pick(i)], 1, sums, rm, rv, nbt, gamma, gamma, 0.1, 1e-5, False, None))
            res["colstats/" + tag] = graph_time(lambda i: K.act_colstats(Y[pick(i)], 1))
        print("M=%d F=%d  (31 MB read per bwd launch at M=19072: 4.7 us at 6.56 TB/s)" % (M, F))
        for k, v in res.items():
            print("   %-24s %7.2f us" % (k, v))


if __name__ == "__main__":
    main()
