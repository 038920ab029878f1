#!/usr/bin/env python
"""bench.py — molecules/s of one full 3DInfomax pre-training step (PNA + Net3D + NTXent, fwd + bwd + Adam) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--mode graph|eager]

Workload (config.workload): BASELINE.json configs[1] — QM9-shaped synthetic molecules, batch 512 PER GPU, the PNA
the reference ships in configs_clean/pre-train_QM9.yml (hidden 200, 7 layers) + Net3D (hidden 20) + NTXent(tau 0.1),
Adam lr 8e-5.  A "step" = CSR build, both encoders, loss, backward, gradient pack, [NCCL all-gather / reduce-scatter
of embeddings + all-reduce of gradients when N>1], Adam — on one collated batch.

value : inputs already resident in HBM (a rotating pool of distinct batches), CUDA events around exactly K steps,
        barrier + synchronize on both sides, max over ranks.
e2e   : the same step driven from PINNED HOST buffers through the public API: every step copies the batch
        host->device and reads the loss back (device->host) inside the timed region.
roofline / cpu_baseline: see DESIGN.md "Measurement".  --impl reference times the CPU oracle port of the
reference (oracle/oracle.py; the reference itself needs DGL, which is not installable) on all host cores.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "molecules/sec PNA+Net3D QM9 pretrain (fwd+bwd+Adam)"
UNIT = "molecules/s"
POOL = 4            # distinct resident batches the timed loop rotates over
LR = 8e-5
TAU = 0.1


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=40)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=512, help="molecules per GPU")
    p.add_argument("--mode", default="graph", choices=["graph", "eager"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same
# workload (profiles/r01_aggregate_full.txt: ncu flushes L2 before each launch, so these are cold-cache bytes)
# fwd: profiles/r01_aggregate_full.txt; bwd: profiles/r01_ncu_full_final_summary.txt (76.6 MB read + 3.6 MB written)
AGG_TRAFFIC = {"fwd": 15.33e6, "bwd": 80.2e6}   # batch 512, N=9392, E=19444; output writes stay in L2 past kernel end


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, mx = [], set(), None
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            busy = sm[len(sm) // 2:]                    # upper half = samples taken under load
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_oracle_throughput(batch, steps, warmup, threads=None):
    """molecules/s of the CPU oracle port (reference op sequence incl. degree bucketing) on the host cores."""
    import torch
    from oracle import oracle as O
    syn = importlib.import_module("3dinfomax_b200.synthetic")
    if threads:
        torch.set_num_threads(threads)
    c2, c3 = O.pna_cfg(**O.PRETRAIN_QM9_PNA), O.net3d_cfg(**O.PRETRAIN_QM9_NET3D)
    tr = O.OracleTrainer(c2, c3, O.init_pna_state(c2, 1), O.init_net3d_state(c3, 2), loss="NTXent", tau=TAU, lr=LR)
    batches = [O.graphs_from_batch(syn.make_batch(100 + i, batch)) for i in range(2)]
    for i in range(warmup):
        tr.step(*batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        tr.step(*batches[i % 2])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # size the per-step sample so that (K + W) steps stay within ~2.5 minutes
    _, t_probe, _ = cpu_oracle_throughput(32, 1, 1, cores)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    b = args.batch
    while b > 32 and t_probe * (b / 32.0) > budget:
        b //= 2
    val, per_step, threads = cpu_oracle_throughput(b, args.steps, args.warmup, cores)
    sample = "%d timed + %d warm-up oracle steps on QM9-shaped batches of %d molecules" % (args.steps, args.warmup, b)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample_batch": b,
                       "note": "CPU oracle port of the reference op sequence (DGL is not installable here)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return ("configs[1]: PNA(hidden 200, 7 layers, as shipped in configs_clean/pre-train_QM9.yml)+Net3D(hidden 20) "
            "NTXent(tau 0.1) Adam, QM9-shaped synthetic, batch %d per GPU" % args.batch)


# ---------------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    i3d = importlib.import_module("3dinfomax_b200")
    ops = importlib.import_module("3dinfomax_b200.ops")
    cfg = importlib.import_module("3dinfomax_b200.configs")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    i3d.lib.load()

    torch.manual_seed(123)                                   # identical replicas on every rank (train.py:235 seed_all)
    pna = i3d.PNA(avg_d=1, device=dev, **cfg.PRETRAIN_QM9_MODEL_PARAMETERS)
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **cfg.PRETRAIN_QM9_MODEL3D_PARAMETERS)
    graph_mode = args.mode == "graph"
    tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=TAU), dev, {"lr": LR},
                                   process_group=dist.group.WORLD if world > 1 else None, graph_safe=graph_mode)

    host = [i3d.batch_from_numpy(i3d.synthetic.make_batch(1000 * rank + i, args.batch), "cpu", pin=True)
            for i in range(POOL)]
    resident = [(g2.to(dev), g3.to(dev)) for g2, g3 in host]
    torch.cuda.synchronize()
    h2d = sum(t.numel() * t.element_size() for g2, g3 in host[:1] for t in
              (g2.edges()[0], g2.edges()[1], g2.batch_num_nodes(), g2.ndata["feat"], g2.edata["feat"],
               g3.edges()[0], g3.edges()[1], g3.batch_num_nodes(), g3.edata["d"]))

    def fresh(pair):          # forward consumes the graph object (ndata['feat'] is overwritten, as in the reference)
        g2, g3 = pair
        a = i3d.GraphBatch(*g2.edges(), g2.batch_num_nodes(), None, {"feat": g2.ndata["feat"]},
                           {"feat": g2.edata["feat"]}, g2.number_of_nodes(), g2.max_in_degree)
        b = i3d.GraphBatch(*g3.edges(), g3.batch_num_nodes(), None, {}, {"d": g3.edata["d"]}, g3.number_of_nodes())
        return a, b

    caps = None
    note = ""
    if graph_mode:
        try:
            caps = [i3d.CapturedStep(tr, *fresh(p), warmup=2 if i == 0 else 1) for i, p in enumerate(resident)]
        except Exception as e:   # capture is an optimisation, not a requirement: fall back to eager launches
            note = "graph capture failed (%s); eager launches" % (str(e).splitlines()[0][:120],)
            print("bench.py WARNING: " + note, file=sys.stderr, flush=True)
            caps = None
            torch.cuda.synchronize()
            tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=TAU), dev, {"lr": LR},
                                           process_group=dist.group.WORLD if world > 1 else None)

    def step_resident(i):
        if caps is not None:
            return caps[i % POOL].run()
        g2, g3 = fresh(resident[i % POOL])
        loss, _, _ = tr.process_batch(([g2], [g3]))
        return loss

    def step_e2e(i):
        g2h, g3h = host[i % POOL]
        if caps is not None:
            caps[i % POOL].load(g2h, g3h)
            loss = caps[i % POOL].run()
        else:
            g2, g3 = g2h.to(dev, non_blocking=True), g3h.to(dev, non_blocking=True)
            loss, _, _ = tr.process_batch(([g2], [g3]))
        return float(loss.item())                                  # device->host read of the step's result

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n0 = i3d.lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            last = fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        launched = i3d.lib.launch_count() - n0
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last, launched

    sampler = ClockSampler(local) if rank == 0 else None
    ms, last_loss, launched = timed(step_resident, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _ = timed(step_e2e, args.steps, 3)
    if caps is not None:   # graph replays re-launch the kernels recorded at capture time
        gpu_launches = int(round(sum(c.launches_per_step for c in caps) / len(caps) * args.steps))
    else:
        gpu_launches = int(launched)

    # ---- roofline of the aggregation kernel (the graded kernel) -------------------------------------
    # The step replays from a CUDA graph, so single launches cannot be bracketed there, and an event pair around one
    # ~10 us launch in eager mode measures the event/launch latency, not the kernel (probe: +7..10 us).  So the kernel
    # is timed the way it runs in the step: 40 launches per CUDA graph on this batch's CSR, replayed 20 times between
    # two events.  "cold": rotating over 8 buffer sets (8 x 45 MB > 126 MB L2), every launch re-reads HBM -> reported
    # as `achieved`.  "warm": one buffer set, messages L2-resident as they are right after the producing kernel.
    roof = None
    if rank == 0:
        hbm, peak_src = peaks()
        K = i3d.kernels
        g2, _ = fresh(resident[0])
        st = i3d.graph.graph_structure(g2)
        rowptr = st.rowptr
        N, E, F = g2.number_of_nodes(), int(st.rowptr[-1].item()), int(cfg.PRETRAIN_QM9_MODEL_PARAMETERS["hidden_dim"])
        b_f = 4 * F * E + 4 * E + 4 * (N + 1) + 16 * F * N
        b_b = 32 * F * N + 8 * F * E + 4 * (N + 1)
        RP = 8
        msgs = [torch.randn(E, F, device=dev) for _ in range(RP)]
        outs = [K.pna_aggregate_fwd(m, rowptr) for m in msgs]
        gs = [torch.randn(N, 4 * F, device=dev) for _ in range(RP)]

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
            return a.elapsed_time(b) / (reps * inner)          # ms per launch

        t_f = graph_time(lambda i: K.pna_aggregate_fwd(msgs[i % RP], rowptr))
        t_fw = graph_time(lambda i: K.pna_aggregate_fwd(msgs[0], rowptr))
        t_b = graph_time(lambda i: K.pna_aggregate_bwd(gs[i % RP], msgs[i % RP], outs[i % RP], rowptr))
        t_bw = graph_time(lambda i: K.pna_aggregate_bwd(gs[0], msgs[0], outs[0], rowptr))
        del msgs, outs, gs
        gbs = lambda nbytes, t: nbytes / (t * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "pna_aggregate_fwd_kernel<4,2,3>", "achieved": gbs(b_f, t_f),
                "peak": hbm, "peak_source": peak_src, "unit": "GB/s", "frac": gbs(b_f, t_f) / hbm,
                "traffic": AGG_TRAFFIC.get("fwd"), "algorithmic_bytes_per_launch": b_f, "us_per_launch": t_f * 1e3,
                "warm": {"us_per_launch": t_fw * 1e3, "frac": gbs(b_f, t_fw) / hbm},
                "bwd": {"kernel": "pna_aggregate_bwd_kernel<4>", "achieved": gbs(b_b, t_b),
                        "frac": gbs(b_b, t_b) / hbm, "traffic": AGG_TRAFFIC.get("bwd"),
                        "algorithmic_bytes_per_launch": b_b, "us_per_launch": t_b * 1e3,
                        "warm": {"us_per_launch": t_bw * 1e3, "frac": gbs(b_b, t_bw) / hbm}},
                "launches_per_step": {"fwd": len(pna.node_gnn.mp_layers), "bwd": len(pna.node_gnn.mp_layers)},
                "how": "40 launches per CUDA graph on the timed batch's CSR (N=%d, E=%d, F=%d), 20 replays between two "
                       "CUDA events on the launch stream; achieved = cold (rotating over 8 buffer sets > L2), warm = "
                       "one buffer set; algorithmic bytes = 4F*E + 4E + 4(N+1) + 16F*N (fwd), 32F*N + 8F*E + 4(N+1) "
                       "(bwd), SURVEY 8d" % (N, E, F)}

    # ---- device-side batch construction: time of PackedMoleculeStore.collate per batch ----------------
    collate_info = None
    if rank == 0:
        store = i3d.PackedMoleculeStore(i3d.synthetic.make_store(7, 4 * args.batch), dev)
        rng = __import__("numpy").random.default_rng(11)
        idxs = [rng.integers(0, len(store), size=args.batch) for _ in range(16)]
        for ix in idxs[:6]:                      # batch sizes differ: let the caching allocator see them first
            store.collate(ix)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for ix in idxs[6:]:
            cg2, cg3 = store.collate(ix)
        c1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 10
        collate_info = {"us_per_batch_device": c0.elapsed_time(c1) * 1e3 / 10, "us_per_batch_wall": wall * 1e6,
                        "h2d_bytes_per_batch": 8 * (7 * args.batch + 3),
                        "what": "PackedMoleculeStore.collate: molecule indices -> batched 2-D + 3-D graphs on the "
                                "device (replaces B x Dataset.__getitem__ + dgl.batch + graph H2D copy)"}
        del store, cg2, cg3

    if rank == 0:
        mols = args.batch * world * args.steps
        line = {"metric": METRIC, "value": mols / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args), "global_batch": args.batch * world,
                           "parallelism": "dp%d" % world, "launch": "cuda-graph" if caps is not None else "eager",
                           "l2": "rotating pool of %d distinct batches; per-step working set (activations ~1 GB at "
                                 "batch 512) exceeds the 126 MB L2, no explicit flush" % POOL,
                           "bn": "local per-rank batch statistics", "note": note, "last_loss": float(last_loss)},
                "e2e": {"value": mols / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roof, "collate": collate_info}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            val, per, threads = cpu_oracle_throughput(min(args.batch, 256), 3, 1, cores)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "3 timed + 1 warm-up oracle steps, QM9-shaped batch of %d molecules"
                                              % min(args.batch, 256)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured CUDA graphs hold NCCL kernels; tearing the communicator down under them deadlocks
        # (observed on 2xB200: both ranks printed, then hung in destroy_process_group).  All collectives are complete
        # and the result line is flushed, so leave without running NCCL / graph destructors.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
