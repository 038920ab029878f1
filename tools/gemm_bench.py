#!/usr/bin/env python
"""Warm per-launch time of the hot path's GEMM shapes (CUDA-graph replay, 20 launches per graph).  Tuning aid:
I3D_WS_FORCE="<BN>,<OCC>" overrides the NT tile choice for N <= 208.

    python tools/gemm_bench.py [batch]
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
i3d = importlib.import_module("3dinfomax_b200")
K = importlib.import_module("3dinfomax_b200.kernels")
G = importlib.import_module("3dinfomax_b200.graph")


def graph_time(fn, reps=10, inner=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * inner)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    os.environ["I3D_PLAN_MIN_NODES"] = "0"
    g2, _ = i3d.batch_from_numpy(i3d.synthetic.make_batch(1000, B), dev)
    st = G.graph_structure(g2)
    N, E, F = st.N, st.E, 200
    rnd = lambda *s: torch.randn(*s, device=dev)
    h, agg, ef, dY_e, dY_n = rnd(N, F), rnd(N, 4 * F), rnd(E, F), rnd(E, F), rnd(N, F)
    W1, W2, W3 = rnd(F, 3 * F) * 0.05, rnd(F, F) * 0.05, rnd(F, 13 * F) * 0.05
    b = rnd(F) * 0.1
    res = {"B": B, "N": N, "E": E, "force": os.environ.get("I3D_WS_FORCE", "")}
    Ye, Yn = torch.empty(E, F, device=dev), torch.empty(N, F, device=dev)

    segs1 = [{"A": h, "B": W1[:, :F], "K": F, "a_idx": st.src_csr}, {"A": h, "B": W1[:, F:2 * F], "K": F, "a_idx": st.dst_csr},
             {"A": ef, "B": W1[:, 2 * F:], "K": F}]
    res["edge_fc1_fwd_K600"] = graph_time(lambda: K.gemm(K.NT, E, F, segs1, Ye, bias=b, stats_act=1))
    segs2 = [{"A": ef, "B": W2, "K": F}]
    res["edge_fc2_fwd_K200"] = graph_time(lambda: K.gemm(K.NT, E, F, segs2, Ye, bias=b, stats_act=0))
    segs3 = [{"A": h, "B": W3[:, :F], "K": F}, {"A": agg, "B": W3[:, F:5 * F], "K": 4 * F},
             {"A": agg, "B": W3[:, 5 * F:9 * F], "K": 4 * F, "scale": st.amp},
             {"A": agg, "B": W3[:, 9 * F:], "K": 4 * F, "scale": st.att}]
    res["node_post_generic_K2600"] = graph_time(lambda: K.gemm(K.NT, N, F, segs3, Yn, bias=b, stats_act=0))
    merged = K.MergedPosttransWeights(F, F, st.plan.n_buckets, dev)
    merged.refresh(W3)
    msegs = [{"A": h, "K": F, "a_idx": st.plan.perm}, {"A": agg, "K": 4 * F, "a_idx": st.plan.perm}]
    res["node_post_merged_K1000"] = graph_time(
        lambda: K.gemm_nt_bucketed(st.plan, F, msegs, Yn, b, merged.fwd_hi, merged.fwd_lo, stats_act=0))
    buf = torch.empty(N, 5 * F, device=dev)
    res["node_post_merged_dx_N1000_K200"] = graph_time(
        lambda: K.gemm_nt_bucketed(st.plan, 5 * F, [{"A": dY_n, "K": F, "a_idx": st.plan.perm}], buf, None,
                                   merged.bwd_hi, merged.bwd_lo))
    dW = torch.zeros(F, F, device=dev)
    res["tn_dW_edge_K_E"] = graph_time(lambda: K.gemm(K.TN, F, F, [{"A": dY_e, "B": ef, "K": E}], dW, accumulate=True))
    res["tn_dW_edge_gather_K_E"] = graph_time(
        lambda: K.gemm(K.TN, F, F, [{"A": dY_e, "B": h, "K": E, "b_idx": st.src_csr}], dW, accumulate=True))
    res["tn_dW_node_K_N"] = graph_time(lambda: K.gemm(K.TN, F, F, [{"A": dY_n, "B": h, "K": N}], dW, accumulate=True))
    dWb = torch.zeros(st.plan.n_buckets, F, 4 * F, device=dev)
    res["tn_dW_chunked_N800"] = graph_time(lambda: K.gemm_tn_chunked(st.plan, dY_n, agg, dWb))
    dW8 = torch.zeros(F, 4 * F, device=dev)
    res["tn_dW_node_N800_K_N"] = graph_time(lambda: K.gemm(K.TN, F, 4 * F, [{"A": dY_n, "B": agg, "K": N}], dW8, accumulate=True))
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    if os.environ.get("I3D_WS_DEBUG_REPORT"):
        import ctypes
        L = importlib.import_module("3dinfomax_b200.lib").load()
        names = ["kernel_total", "mma_wait_a_full", "mma_wait_b_full", "mma_issue", "tma_wait_mma_done",
                 "stager_wait_mma_done", "stager_main_loop", "epilogue", "launches", "epi_tile_in_smem", "epi_stats_done",
                 "epi_stores_issued"]

        def counters(tag, fn):
            buf = (ctypes.c_ulonglong * 16)()
            torch.cuda.synchronize()
            L.i3d_gemm_debug_counters(buf)
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            L.i3d_gemm_debug_counters(buf)
            n = max(int(buf[8]), 1)
            print(tag, {k: int(buf[i]) // n for i, k in enumerate(names) if k != "launches"}, "launches", int(buf[8]),
                  flush=True)

        counters("merged_K1000", lambda: K.gemm_nt_bucketed(st.plan, F, msegs, Yn, b, merged.fwd_hi, merged.fwd_lo, stats_act=0))
        counters("edge_fc1_K600", lambda: K.gemm(K.NT, E, F, segs1, Ye, bias=b, stats_act=1))
        counters("edge_fc2_K200", lambda: K.gemm(K.NT, E, F, segs2, Ye, bias=b, stats_act=0))
        counters("generic_K2600", lambda: K.gemm(K.NT, N, F, segs3, Yn, bias=b, stats_act=0))


if __name__ == "__main__":
    main()
