"""Micro-benchmark of the PNA aggregation kernels (run on the GPU box): warm (L2-resident, back-to-back) and cold
(rotating over buffers larger than L2) time per launch against the algorithmic bytes of SURVEY.md §8d.

    python tests/gpu_agg_bench.py [batch]          # env I3D_AGG_FWD / I3D_AGG_BWD select tuning variants
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    peak = 6538.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    b = i3d.synthetic.make_batch(1, B)
    g2, _ = i3d.batch_from_numpy(b, dev)
    st = G.graph_structure(g2)
    rowptr = st.rowptr
    N, E, F = g2.number_of_nodes(), st.rowptr[-1].item(), 200
    bytes_f = 4 * F * E + 4 * E + 4 * (N + 1) + 16 * F * N
    bytes_b = 32 * F * N + 8 * F * E + 4 * (N + 1)
    POOL = 8                                                  # 8 x (15 + 30) MB > 126 MB of L2
    msgs = [torch.randn(E, F, device=dev) for _ in range(POOL)]
    outs = [K.pna_aggregate_fwd(m, rowptr) for m in msgs]
    gs = [torch.randn(N, 4 * F, device=dev) for _ in range(POOL)]

    def timeit(fn, reps=20, inner=40):
        """GPU time per launch: `inner` launches captured in a CUDA graph (no host in the loop), replayed `reps` times"""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(inner):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * inner)

    res = {}
    res["fwd_warm_us"] = timeit(lambda i: K.pna_aggregate_fwd(msgs[0], rowptr))
    res["fwd_cold_us"] = timeit(lambda i: K.pna_aggregate_fwd(msgs[i % POOL], rowptr))
    res["bwd_warm_us"] = timeit(lambda i: K.pna_aggregate_bwd(gs[0], msgs[0], outs[0], rowptr))
    res["bwd_cold_us"] = timeit(lambda i: K.pna_aggregate_bwd(gs[i % POOL], msgs[i % POOL], outs[i % POOL], rowptr))
    line = {"B": B, "N": N, "E": E, "fwd_variant": os.environ.get("I3D_AGG_FWD", "default"),
            "bwd_variant": os.environ.get("I3D_AGG_BWD", "default")}
    for k, v in res.items():
        nb = bytes_f if k.startswith("fwd") else bytes_b
        line[k] = round(v, 2)
        line[k.replace("_us", "_frac")] = round(nb / (v * 1e-6) / 1e9 / peak, 3)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
