"""Kernel- and model-level parity cases: the CUDA path (through the C ABI) against the CPU oracle / plain fp32-fp64
torch references on the same seeded inputs.  Each case returns [(label, error, tolerance)].

Shared by tests/test_gpu_parity.py (pytest, -m gpu) and tests/gpu_diag.py (prints every number without stopping).
Tolerances: integer / index work bit exact (tol 0); fp32 element-wise and reductions 1e-6..1e-5 relative to the
tensor's magnitude; GEMM-containing results 2e-5 relative (fp32 accumulation order differs from MKL);
end-to-end embeddings 1e-4 relative, parameter gradients 1e-3 of the global gradient scale (measured: 6e-5 on the
QM9-shaped case, 8e-4 worst on the 6-molecule QMugs-shaped case where train-mode BN over few rows amplifies fp32
accumulation-order noise through 7 layers; CPU fp32 vs fp64 of the oracle itself differ by 5e-5 there).
"""
import importlib
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
i3d = importlib.import_module("3dinfomax_b200")
K = importlib.import_module("3dinfomax_b200.kernels")
ops = importlib.import_module("3dinfomax_b200.ops")
syn = i3d.synthetic
DEV = "cuda"


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    if a.shape != b.shape:
        return float("inf")
    if a.numel() == 0:
        return 0.0
    if not torch.isfinite(a).all():
        return float("inf")
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def exact(a, b):
    a = torch.as_tensor(a).cpu()
    b = torch.as_tensor(b).cpu()
    return 0.0 if a.shape == b.shape and torch.equal(a.to(b.dtype), b) else 1.0


def gen(seed):
    return torch.Generator().manual_seed(seed)


def random_graph(seed, n, e, isolated=True):
    g = gen(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - (2 if isolated and n > 2 else 0), (e,), generator=g)   # last two nodes: in-degree 0
    return src, dst


# ------------------------------------------------------------------------------------------------ graph
def case_csr():
    out = []
    for tag, (n, e) in {"small": (7, 19), "empty_edges": (5, 0), "bond_like": (3000, 6100),
                        "high_degree": (64, 64 * 63)}.items():
        src, dst = random_graph(1, n, e) if e else (torch.zeros(0, dtype=torch.long),) * 2
        if tag == "high_degree":
            s, d = syn.complete_graph_edges(64)
            src, dst = torch.from_numpy(s), torch.from_numpy(d)
        rowptr, col, eid = O.csr_reference(src.numpy(), dst.numpy(), n)
        r, c, rid, ei = K.csr_build(dst.to(DEV), src.to(DEV), n)
        out += [("csr/%s/rowptr" % tag, exact(r, rowptr), 0), ("csr/%s/col" % tag, exact(c, col), 0),
                ("csr/%s/eid" % tag, exact(ei, eid), 0),
                ("csr/%s/rowid" % tag, exact(rid, dst.numpy()[eid] if e else np.zeros(0)), 0)]
        # int32 variant: sort the CSR-ordered edge list by source -> position map
        if e:
            r2, c2, _, pos = K.csr_build(c, rid, n)
            rp2, col2, eid2 = O.csr_reference(rid.cpu().numpy(), c.cpu().numpy(), n)
            out += [("csr/%s/out_rowptr" % tag, exact(r2, rp2), 0), ("csr/%s/out_pos" % tag, exact(pos, eid2), 0)]
    counts = torch.tensor([3, 0, 5, 1, 29], dtype=torch.int64)
    ptr = K.segment_ptr(counts.to(DEV))
    out.append(("segment_ptr", exact(ptr, torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])), 0))
    big = torch.randint(0, 40, (5000,), generator=gen(2))
    out.append(("segment_ptr/5000", exact(K.segment_ptr(big.to(DEV)),
                                           torch.cat([torch.zeros(1, dtype=torch.long), big.cumsum(0)])), 0))
    rp = torch.tensor([0, 0, 1, 3, 6, 10, 10], dtype=torch.int32)
    amp, att = K.degree_scalers(rp.to(DEV))
    D = (rp[1:] - rp[:-1]).numpy()
    with np.errstate(divide="ignore"):
        ra = np.where(D > 0, np.log(D + 1.0), 0.0).astype(np.float32)
        rt = np.where(D > 0, 1.0 / np.log(D + 1.0), 0.0).astype(np.float32)
    out += [("degree_scalers/amp", exact(amp, ra), 0), ("degree_scalers/att", exact(att, rt), 0)]
    return out


# -------------------------------------------------------------------------------------------- embedding
def case_embed():
    g = gen(3)
    dims = syn.ATOM_FEATURE_DIMS
    out = []
    for Fd in (200, 20, 6):
        tables = [torch.randn(d, Fd, generator=g) for d in dims]
        idx = torch.stack([torch.randint(0, d, (57,), generator=g) for d in dims], 1)
        perm = torch.randperm(57, generator=g).int()
        table = torch.cat(tables).requires_grad_(True)
        off = torch.tensor(np.concatenate([[0], np.cumsum(dims)[:-1]]), dtype=torch.int32)
        ref = sum(F.embedding(idx[perm.long(), c] + int(off[c]), table) for c in range(len(dims)))
        gout = torch.randn(57, Fd, generator=g)
        ref.backward(gout)
        t_dev = table.detach().to(DEV).requires_grad_(True)
        got = ops.embed_sum(idx.to(DEV), off.to(DEV), perm.to(DEV), t_dev)
        got.backward(gout.to(DEV))
        out += [("embed/F%d/fwd" % Fd, rel(got, ref), 1e-6), ("embed/F%d/bwd" % Fd, rel(t_dev.grad, table.grad), 1e-5)]
    return out


# ------------------------------------------------------------------------------------------------- gemm
def _gemm_ref(mode, M, N, segs, bias, C0):
    acc = torch.zeros(M, N, dtype=torch.float64) if C0 is None else C0.double().clone()
    for s in segs:
        A, B = s["A"].double(), s["B"].double()
        if mode in (K.NT, K.NN):
            a = A[s["a_idx"].long()] if s.get("a_idx") is not None else A[:M]
            if s.get("scale") is not None:
                a = a * s["scale"].double()[:, None]
            acc += a @ (B.t() if mode == K.NT else B)
        else:
            a = A[s["a_idx"].long()] if s.get("a_idx") is not None else A
            b = B[s["b_idx"].long()] if s.get("b_idx") is not None else B
            if s.get("scale") is not None:
                a = a * s["scale"].double()[:, None]
            acc += a.t() @ b
    if bias is not None:
        acc += bias.double()[None, :]
    return acc


def _to_dev(segs):
    return [{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in s.items()} for s in segs]


def case_gemm():
    g = gen(4)
    out = []
    rn = lambda *s: torch.randn(*s, generator=g)
    # (tag, mode, M, N, Ks, gather, scale, bias, accumulate)
    for tag, mode, M, N, Ks, gather, scale, bias, accum in [
        ("nt_plain", K.NT, 300, 200, [600], False, False, True, False),
        ("nt_tiny", K.NT, 1, 3, [5], False, False, True, False),
        ("nt_k9", K.NT, 777, 20, [9], False, False, True, False),
        ("nt_gather3", K.NT, 1000, 200, [200, 200, 200], True, False, True, False),
        ("nt_scale4", K.NT, 513, 200, [200, 800, 800, 800], False, True, True, False),
        ("nt_sim", K.NT, 96, 288, [256], False, False, False, False),
        ("nn_plain", K.NN, 300, 600, [200], False, False, False, False),
        ("nn_scale3_acc", K.NN, 257, 800, [200, 200, 200], False, True, False, True),
        ("nn_small", K.NN, 50, 20, [20], False, False, False, False),
        ("tn_plain", K.TN, 200, 600, [5000], False, False, False, False),
        ("tn_gather", K.TN, 200, 200, [3001], True, False, False, False),
        ("tn_scale", K.TN, 200, 800, [2000], False, True, False, False),
        ("tn_acc", K.TN, 20, 60, [999], False, False, False, True),
        ("tn_smallk", K.TN, 256, 20, [7], False, False, False, False),
    ]:
        segs = []
        rows = 400
        for Kd in Ks:
            s = {"K": Kd}
            if mode == K.NT:
                s["A"] = rn(rows if gather else M, Kd)
                s["B"] = rn(N, Kd)
                if gather:
                    s["a_idx"] = torch.randint(0, rows, (M,), generator=g).int()
                if scale and len(segs) > 0:
                    s["scale"] = rn(M)
            elif mode == K.NN:
                s["A"] = rn(M, Kd)
                s["B"] = rn(Kd, N)
                if scale and len(segs) > 0:
                    s["scale"] = rn(M)
            else:
                s["A"] = rn(Kd, M)
                s["B"] = rn(rows if gather else Kd, N)
                if gather:
                    s["b_idx"] = torch.randint(0, rows, (Kd,), generator=g).int()
                if scale:
                    s["scale"] = rn(Kd)
            segs.append(s)
        b = rn(N) if bias else None
        C0 = rn(M, N) if accum else None
        ref = _gemm_ref(mode, M, N, segs, b, C0)
        C = C0.clone().to(DEV) if accum else torch.full((M, N), float("nan"), device=DEV)
        K.gemm(mode, M, N, _to_dev(segs), C, None if b is None else b.to(DEV), accum)
        out.append(("gemm/" + tag, rel(C, ref), 3e-5))
    # strided views: weight column blocks and an output column block, as the FC operator uses them
    W = rn(200, 2600)
    X = rn(321, 800)
    Cbig = torch.zeros(321, 1000, device=DEV)
    K.gemm(K.NT, 321, 200, [{"A": X.to(DEV), "B": W.to(DEV)[:, 1000:1800], "K": 800}], Cbig[:, 400:600])
    out.append(("gemm/strided_views", rel(Cbig[:, 400:600], X.double() @ W[:, 1000:1800].double().t()), 2e-5))
    out.append(("gemm/strided_untouched", float(Cbig[:, :400].abs().max() + Cbig[:, 600:].abs().max()), 0))
    return out


def case_gemm_tc():
    """tcgen05 (3xTF32, TMEM accumulator) backend against fp64 and against the fp32 SIMT backend of the same ABI."""
    g = gen(40)
    out = []
    rn = lambda *s: torch.randn(*s, generator=g)
    L = i3d.lib.load()
    for tag, M, N, Ks, gather, scale, bias, accum in [
        ("fc1_gather3", 19000, 200, [200, 200, 200], True, False, True, False),
        ("fc2_plain", 4111, 200, [200], False, False, True, False),
        ("posttrans_scale4", 9226, 200, [200, 800, 800, 800], False, True, True, False),
        ("k_tail_36", 1000, 200, [36], False, False, False, False),
        ("n256_head", 512, 256, [200], False, False, True, True),
        ("n512_two_tiles", 2048, 512, [256], False, False, False, False),
        ("n32_net3d", 30000, 20, [20, 20, 20], True, False, True, False),
        ("n64", 777, 64, [64], False, False, False, False),
    ]:
        segs, rows = [], 5000
        for Kd in Ks:
            s = {"K": Kd, "A": rn(rows if gather else M, Kd), "B": rn(N, Kd)}
            if gather:
                s["a_idx"] = torch.randint(0, rows, (M,), generator=g).int()
            if scale and len(segs) > 0:
                s["scale"] = rn(M)
            segs.append(s)
        b = rn(N) if bias else None
        C0 = rn(M, N) if accum else None
        ref = _gemm_ref(K.NT, M, N, segs, b, C0)
        dsegs = _to_dev(segs)
        res = {}
        for backend in (0, 1):
            L.i3d_gemm_backend(backend)
            C = C0.clone().to(DEV) if accum else torch.full((M, N), float("nan"), device=DEV)
            K.gemm(K.NT, M, N, dsegs, C, None if b is None else b.to(DEV), accum)
            res[backend] = C
        L.i3d_gemm_backend(0)
        out += [("gemm_tc/%s/vs_fp64" % tag, rel(res[0], ref), 3e-5),
                ("gemm_tc/%s/simt_vs_fp64" % tag, rel(res[1], ref), 2e-5)]
    # TN (weight gradients): split-K over CTAs, operands transposed while staged, fp32 atomics into C
    for tag, M, N, Kd, gather, scale, accum in [
        ("dW_fc2", 200, 200, 19092, False, False, False),
        ("dW_fc1_gather", 200, 200, 19000, True, False, False),
        ("dW_post_scaled", 200, 800, 9226, False, True, False),
        ("dW_post_h", 200, 1000, 9226, False, False, False),
        ("dz2_ntxent", 1536, 256, 512, False, False, False),
        ("dW_small_acc", 20, 60, 30000, False, False, True),
        ("k_tail", 64, 32, 1000, False, False, False),
    ]:
        rows = 5000
        s = {"K": Kd, "A": rn(Kd, M), "B": rn(rows if gather else Kd, N)}
        if gather:
            s["b_idx"] = torch.randint(0, rows, (Kd,), generator=g).int()
        if scale:
            s["scale"] = rn(Kd)
        C0 = rn(M, N) if accum else None
        ref = _gemm_ref(K.TN, M, N, [s], None, C0)
        dsegs = _to_dev([s])
        res = {}
        for backend in (0, 1, 2):
            L.i3d_gemm_backend(backend)
            C = C0.clone().to(DEV) if accum else torch.full((M, N), float("nan"), device=DEV)
            K.gemm(K.TN, M, N, dsegs, C, None, accum)
            res[backend] = C
        L.i3d_gemm_backend(0)
        out += [("gemm_tc/tn_%s/vs_fp64" % tag, rel(res[0], ref), 3e-5),
                ("gemm_tc/tn_%s/mn_major_vs_fp64" % tag, rel(res[2], ref), 3e-5),
                ("gemm_tc/tn_%s/simt_vs_fp64" % tag, rel(res[1], ref), 2e-5)]
    x = rn(777, 200)
    out.append(("transpose", exact(K.transpose(x.to(DEV)[:, 8:72]), x[:, 8:72].t().contiguous()), 0))
    return out


# --------------------------------------------------------------------------------------- FC tail (BN)
def case_weight_prep():
    """Operands prepared once per step (kernels.WeightPrep, one launch for all weights) must give bit-identical
    GEMMs to the per-call split, plain and transposed, and must go stale on any update of the weight."""
    g = gen(41)
    out = []
    rn = lambda *s: torch.randn(*s, generator=g)
    prep = K.WeightPrep(DEV)
    W = rn(200, 616).to(DEV)                     # 600 used + an unused tail, like a wider weight
    W2 = rn(100, 2600).to(DEV)
    M = 3000
    xs = [rn(M, k).to(DEV) for k in (200, 200, 200)]
    sc = rn(M).to(DEV)
    segs = [{"A": xs[0], "K": 200, "B": W[:, 0:200]}, {"A": xs[1], "K": 200, "B": W[:, 200:400], "scale": sc},
            {"A": xs[2], "K": 200, "B": W[:, 400:600]}]
    e1 = prep.entry(W, ("f", 200, 200, 200), 200, [(sg["B"], 200) for sg in segs], False)
    # transposed: dx = dy W[:, 200:400] -> N = 200, K = 200 rows of W
    dy = rn(M, 200).to(DEV)
    e2 = prep.entry(W, ("b", 200, 200), 200, [(W[:, 200:400], 200)], True)
    # K not a multiple of 32, N not a multiple of 32
    x3 = rn(M, 36).to(DEV)
    e3 = prep.entry(W2, ("f", 36), 100, [(W2[:, 4:40], 36)], False)
    dy3 = rn(M, 100).to(DEV)
    e4 = prep.entry(W2, ("b", 36, 4), 36, [(W2[:, 4:40], 100)], True)
    out.append(("weight_prep/stale_before_refresh", float(prep.ready(e1)), 0))
    prep.refresh()
    out.append(("weight_prep/ready_after_refresh", 1.0 - float(prep.ready(e1) and prep.ready(e4)), 0))
    ref = K.gemm(K.NT, M, 200, segs, torch.empty(M, 200, device=DEV))
    got = K.gemm(K.NT, M, 200, segs, torch.empty(M, 200, device=DEV), prepared=e1)
    out.append(("weight_prep/plain_bit_exact", exact(got, ref), 0))
    Wt = K.transpose(W)
    ref = K.gemm(K.NT, M, 200, [{"A": dy, "K": 200, "B": Wt[200:400, :]}], torch.empty(M, 200, device=DEV))
    got = K.gemm(K.NT, M, 200, [{"A": dy, "K": 200}], torch.empty(M, 200, device=DEV), prepared=e2)
    out.append(("weight_prep/transposed_bit_exact", exact(got, ref), 0))
    ref = K.gemm(K.NT, M, 100, [{"A": x3, "K": 36, "B": W2[:, 4:40]}], torch.empty(M, 100, device=DEV))
    got = K.gemm(K.NT, M, 100, [{"A": x3, "K": 36}], torch.empty(M, 100, device=DEV), prepared=e3)
    out.append(("weight_prep/ragged_bit_exact", exact(got, ref), 0))
    W2t = K.transpose(W2)
    ref = K.gemm(K.NT, M, 36, [{"A": dy3, "K": 100, "B": W2t[4:40, :]}], torch.empty(M, 36, device=DEV))
    got = K.gemm(K.NT, M, 36, [{"A": dy3, "K": 100}], torch.empty(M, 36, device=DEV), prepared=e4)
    out.append(("weight_prep/ragged_transposed_bit_exact", exact(got, ref), 0))
    W.add_(1.0)                                   # torch-visible update -> version counter
    out.append(("weight_prep/stale_after_inplace_update", float(prep.ready(e1)), 0))
    prep.refresh()
    prep.invalidate()                             # what FusedAdam's post-step hook does
    out.append(("weight_prep/stale_after_invalidate", float(prep.ready(e1)), 0))
    return out


def case_bn():
    g = gen(5)
    out = []
    for Fd, M in ((200, 1777), (20, 30001), (9, 101), (256, 64)):
        for act_name in ("none", "relu", "silu"):
            for training in (True, False):
                act = K.ACT[act_name]
                Y = (torch.randn(M, Fd, generator=g) * 0.7 + 0.3).requires_grad_(True)
                gamma = (1 + 0.1 * torch.randn(Fd, generator=g)).requires_grad_(True)
                beta = (0.1 * torch.randn(Fd, generator=g)).requires_grad_(True)
                res = torch.randn(M, Fd, generator=g)
                rm, rv = 0.1 * torch.randn(Fd, generator=g), 0.5 + torch.rand(Fd, generator=g)
                rm_ref, rv_ref = rm.clone(), rv.clone()
                h = {"none": lambda t: t, "relu": torch.relu, "silu": F.silu}[act_name](Y)
                ref = F.batch_norm(h, rm_ref, rv_ref, gamma, beta, training, 0.93, 1e-5) + res
                gout = torch.randn(M, Fd, generator=g)
                ref.backward(gout)
                Yd = Y.detach().to(DEV)
                rm_d, rv_d = rm.to(DEV), rv.to(DEV)
                nbt = torch.zeros((), dtype=torch.long, device=DEV)
                sums = K.act_colstats(Yd, act) if training else None
                Od, save = K.bn_apply(Yd, act, sums, rm_d, rv_d, nbt, gamma.detach().to(DEV), beta.detach().to(DEV),
                                      0.93, 1e-5, training, res.to(DEV))
                sums2 = K.bn_bwd_reduce(gout.to(DEV), Yd, act, save)
                dY, db, dgam, dbet = K.bn_bwd_apply(gout.to(DEV), Yd, act, True, training, save,
                                                    gamma.detach().to(DEV), sums2)
                tag = "bn/F%d/%s/%s" % (Fd, act_name, "train" if training else "eval")
                out += [(tag + "/fwd", rel(Od, ref), 1e-5), (tag + "/dY", rel(dY, Y.grad), 2e-5),
                        (tag + "/dgamma", rel(dgam, gamma.grad), 2e-5), (tag + "/dbeta", rel(dbet, beta.grad), 2e-5),
                        (tag + "/dbias", rel(db, Y.grad.sum(0)) if act_name != "none" or not training else
                         float(db.abs().max().item() / (gout.abs().sum(0).max().item())), 5e-5)]
                if training:
                    out += [(tag + "/running_mean", rel(rm_d, rm_ref), 2e-6), (tag + "/running_var", rel(rv_d, rv_ref), 1e-5),
                            (tag + "/nbt", exact(nbt, torch.tensor(1)), 0)]
    # no-BN path + plain activations
    Y = torch.randn(333, 20, generator=g)
    gy = torch.randn(333, 20, generator=g)
    for act_name in ("relu", "silu"):
        Yr = Y.clone().requires_grad_(True)
        r = {"relu": torch.relu, "silu": F.silu}[act_name](Yr)
        r.backward(gy)
        dY, db, _, _ = K.bn_bwd_apply(gy.to(DEV), Y.to(DEV), K.ACT[act_name], False, False, None, None, None)
        out += [("act/%s/fwd" % act_name, rel(K.act_fwd(Y.to(DEV), K.ACT[act_name]), r), 2e-6),
                ("act/%s/bwd" % act_name, rel(K.act_bwd(gy.to(DEV), Y.to(DEV), K.ACT[act_name]), Yr.grad), 2e-6),
                ("act/%s/nobn_dY" % act_name, rel(dY, Yr.grad), 2e-6),
                ("act/%s/nobn_db" % act_name, rel(db, Yr.grad.sum(0)), 1e-5)]
    out.append(("colsum", rel(K.colsum(Y.to(DEV)), Y.sum(0)), 1e-5))
    out.append(("add", rel(K.add(Y.to(DEV), gy.to(DEV)), Y + gy), 0))
    v = torch.randn(20, generator=g)
    out.append(("broadcast_rows", exact(K.broadcast_rows(v.to(DEV), 77), v[None].expand(77, 20)), 0))
    return out


# ------------------------------------------------------------------------------------------ aggregation
def _csr_order(src, dst, n):
    rowptr, col, eid = O.csr_reference(src.numpy(), dst.numpy(), n)
    return torch.from_numpy(rowptr), torch.from_numpy(eid).long()


def case_aggregate():
    out = []
    for tag, (n, e, Fd) in {"bond": (500, 1100, 200), "narrow": (300, 700, 20), "odd": (50, 170, 6),
                            "dense": (40, 1500, 200)}.items():
        g = gen(6)
        src, dst = random_graph(7, n, e)
        og = O.OGraph(src, dst, torch.tensor([n]))
        msg = torch.randn(e, Fd, generator=g)
        msg[3] = msg[1]                       # exact ties -> first-index gradient routing must match torch.max/min
        dst = dst.clone()
        rowptr, eid = _csr_order(src, dst, n)
        msg_ref = msg.clone().requires_grad_(True)
        ref = O.pna_reduce(og, msg_ref, ["mean", "max", "min", "std"], ["identity"])          # [n, 4F], edge-id order in
        gout = torch.randn(n, 4 * Fd, generator=g)
        ref.backward(gout)
        m_csr = msg[eid].to(DEV).requires_grad_(True)
        got = ops.pna_aggregate(m_csr, rowptr.to(DEV))
        got.backward(gout.to(DEV))
        out += [("aggregate/%s/fwd" % tag, rel(got, ref), 1e-6),
                ("aggregate/%s/bwd" % tag, rel(m_csr.grad, msg_ref.grad[eid]), 1e-5),
                ("aggregate/%s/zero_degree_rows" % tag, float(got[n - 2:].abs().max()), 0)]
    # duplicate messages inside one mailbox (symmetric hydrogens): tie routing + D=1 rows + var gate
    src = torch.tensor([1, 2, 3, 0, 0, 0, 4])
    dst = torch.tensor([0, 0, 0, 1, 2, 3, 1])
    g = gen(8)
    base = torch.randn(1, 200, generator=g)
    msg = torch.cat([base, base, base, torch.randn(4, 200, generator=g)])
    og = O.OGraph(src, dst, torch.tensor([5]))
    rowptr, eid = _csr_order(src, dst, 5)
    mr = msg.clone().requires_grad_(True)
    ref = O.pna_reduce(og, mr, ["mean", "max", "min", "std"], ["identity"])
    gout = torch.randn(5, 800, generator=g)
    ref.backward(gout)
    mc = msg[eid].to(DEV).requires_grad_(True)
    got = ops.pna_aggregate(mc, rowptr.to(DEV))
    got.backward(gout.to(DEV))
    out += [("aggregate/ties/fwd", rel(got, ref), 1e-6), ("aggregate/ties/bwd", rel(mc.grad, mr.grad[eid]), 1e-5)]
    return out


def case_segment_ops():
    g = gen(9)
    out = []
    nn_ = torch.tensor([3, 1, 7, 29, 2, 11])
    og = O.OGraph(torch.zeros(0), torch.zeros(0), nn_)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), nn_.cumsum(0)]).int()
    for Fd in (200, 20, 6):
        for ops_ in (["min", "max", "mean"], ["min", "max", "mean", "sum"], ["sum"]):
            x = torch.randn(int(nn_.sum()), Fd, generator=g)
            x[4] = x[5]                                      # tie inside graph 2
            xr = x.clone().requires_grad_(True)
            ref = torch.cat([O.segment_readout(xr, og, o) for o in ops_], -1)
            gout = torch.randn(ref.shape, generator=g)
            ref.backward(gout)
            xd = x.to(DEV).requires_grad_(True)
            got = ops.readout(xd, ptr.to(DEV), ops_)
            got.backward(gout.to(DEV))
            tag = "readout/F%d/%s" % (Fd, "+".join(ops_))
            out += [(tag + "/fwd", rel(got, ref), 1e-6), (tag + "/bwd", rel(xd.grad, xr.grad), 1e-6)]
    # segment mean / sum over CSR rows (+ addend), and the indexed gather-sum used by the FC backward
    n, e, H = 200, 900, 20
    src, dst = random_graph(10, n, e)
    rowptr, eid = _csr_order(src, dst, n)
    rowid = dst[eid].int()
    deg = (rowptr[1:] - rowptr[:-1]).clamp(min=1).float()
    for mean in (True, False):
        x = torch.randn(e, H, generator=g).requires_grad_(True)       # already CSR ordered
        add = torch.randn(n, H, generator=g).requires_grad_(True)
        ref = torch.zeros(n, H).index_add(0, rowid.long(), x)
        ref = (ref / deg[:, None] if mean else ref) + add
        gout = torch.randn(n, H, generator=g)
        ref.backward(gout)
        xd, ad = x.detach().to(DEV).requires_grad_(True), add.detach().to(DEV).requires_grad_(True)
        got = ops.segment_reduce(xd, rowptr.to(DEV), rowid.to(DEV), mean, ad)
        got.backward(gout.to(DEV))
        tag = "segment_%s" % ("mean" if mean else "sum")
        out += [(tag + "/fwd", rel(got, ref), 1e-6), (tag + "/bwd_x", rel(xd.grad, x.grad), 1e-6),
                (tag + "/bwd_addend", rel(ad.grad, add.grad), 0)]
    x = torch.randn(e, H, generator=g)
    idx = torch.randperm(e, generator=g).int()
    ref = torch.zeros(n, H).index_add(0, rowid.long(), x[idx.long()])
    got = K.segment_sum_fwd(x.to(DEV), rowptr.to(DEV), idx.to(DEV))
    out.append(("segment_sum/indexed", rel(got, ref), 1e-6))
    return out


# ------------------------------------------------------------------------------------------------ net3d
def case_net3d_elementwise():
    g = gen(11)
    out = []
    d = torch.rand(1234, generator=g) * 8 + 0.9
    perm = torch.randperm(1234, generator=g).int()
    ref = O.fourier_encode(d[perm.long()].reshape(-1, 1), 4)
    out.append(("fourier/k4", rel(K.fourier_encode(d.to(DEV), perm.to(DEV), 4), ref), 1e-6))
    out.append(("fourier/k0", exact(K.fourier_encode(d.to(DEV), None, 0), d.reshape(-1, 1)), 0))
    for H in (20, 32, 7):
        msg = torch.randn(999, H, generator=g).requires_grad_(True)
        w = (torch.randn(1, H, generator=g) * 0.7).requires_grad_(True)
        b = torch.randn(1, generator=g).requires_grad_(True)
        ref = msg * torch.sigmoid(F.linear(msg, w, b))
        gout = torch.randn(999, H, generator=g)
        ref.backward(gout)
        md, wd, bd = (t.detach().to(DEV).requires_grad_(True) for t in (msg, w, b))
        got = ops.soft_gate(md, wd, bd)
        got.backward(gout.to(DEV))
        out += [("soft_gate/H%d/fwd" % H, rel(got, ref), 2e-6), ("soft_gate/H%d/dmsg" % H, rel(md.grad, msg.grad), 1e-5),
                ("soft_gate/H%d/dw" % H, rel(wd.grad, w.grad), 1e-4), ("soft_gate/H%d/db" % H, rel(bd.grad, b.grad), 1e-4)]
    return out


# ------------------------------------------------------------------------------------------------- loss
def case_ntxent():
    g = gen(12)
    out = []
    for B, C, D in ((37, 1, 256), (24, 3, 256), (130, 1, 64)):
        z1 = torch.randn(B, D, generator=g)
        z2 = torch.randn(B * C, D, generator=g) * 0.5 + 0.2 * z1.repeat_interleave(C, 0)
        for name, fn, eps in (("NTXent", O.ntxent, 1e-8), ("NTXentMultiplePositives", O.ntxent_multiple_positives, 0.0)):
            if name == "NTXent" and C != 1:
                continue
            a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
            ref = fn(a, b, tau=0.1)
            (ref * 1.7).backward()
            ad, bd = z1.to(DEV).requires_grad_(True), z2.to(DEV).requires_grad_(True)
            mod = getattr(i3d, name)(tau=0.1)
            got = mod(ad, bd)
            (got * 1.7).backward()
            tag = "%s/B%d_C%d" % (name, B, C)
            out += [(tag + "/loss", abs(got.item() - ref.item()) / abs(ref.item()), 2e-6),
                    (tag + "/dz1", rel(ad.grad, a.grad), 2e-5), (tag + "/dz2", rel(bd.grad, b.grad), 2e-5)]
    # with the optional regularisers (commons/losses.py:157-162, 250-258; pinned on the reference's functions by
    # oracle/pin_regularisers.py): the CUDA loss + the regulariser terms against the oracle loss + the same terms on the CPU
    for name, fn, C, kw in (("NTXent", O.ntxent, 1, dict(variance_reg=0.3, covariance_reg=0.2, uniformity_reg=0.1)),
                            ("NTXentMultiplePositives", O.ntxent_multiple_positives, 3,
                             dict(variance_reg=0.3, conformer_variance_reg=0.4))):
        B, D = 40, 64
        # (small embeddings: the uniformity term is log mean exp(-2 ||x_i - x_j||^2), which underflows to log(0) in fp32
        #  — in the reference as well — once the squared distances pass ~45)
        z1 = torch.randn(B, D, generator=g) * 0.25
        z2 = torch.randn(B * C, D, generator=g) * 0.2 + 0.2 * z1.repeat_interleave(C, 0)
        mod = getattr(i3d, name)(tau=0.1, **kw)
        a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        ref = fn(a, b, tau=0.1) + mod.regularisers(a, b, C)
        ref.backward()
        ad, bd = z1.to(DEV).requires_grad_(True), z2.to(DEV).requires_grad_(True)
        got = mod(ad, bd)
        got.backward()
        tag = "%s+regularisers/B%d_C%d" % (name, B, C)
        out += [(tag + "/loss", abs(got.item() - ref.item()) / abs(ref.item()), 5e-6),
                (tag + "/dz1", rel(ad.grad, a.grad), 2e-5), (tag + "/dz2", rel(bd.grad, b.grad), 2e-5)]
    # NTXentMultiplePositivesV2 / V3 (commons/losses.py:598-689) against vectors of the reference's own classes
    from oracle import pin_loss_variants as PV
    gold = np.load(os.path.join(ROOT, "tests", "golden", "loss_variants.npz"))
    z1, z2 = PV.inputs()
    for tag, cls, kw in (("v2", i3d.NTXentMultiplePositivesV2, {}), ("v3", i3d.NTXentMultiplePositivesV3, {}),
                         ("v3_reg", i3d.NTXentMultiplePositivesV3, PV.REG)):
        ad, bd = z1.to(DEV).requires_grad_(True), z2.to(DEV).requires_grad_(True)
        got = cls(tau=PV.CASE["tau"], **kw)(ad, bd)
        got.backward()
        out += [("NTXentMultiplePositives_%s/loss_vs_reference" % tag,
                 abs(got.item() - float(gold[tag])), 5e-6),      # absolute: V2 is a difference of two O(1) terms
                ("NTXentMultiplePositives_%s/dz1_vs_reference" % tag, rel(ad.grad, gold[tag + "_dz1"]), 2e-5),
                ("NTXentMultiplePositives_%s/dz2_vs_reference" % tag, rel(bd.grad, gold[tag + "_dz2"]), 2e-5)]
    # local rows against a gathered column set (data-parallel layout): rows 8..15 of a 24-molecule batch
    B, C, D = 24, 3, 256
    z1 = torch.randn(B, D, generator=g)
    z2 = torch.randn(B * C, D, generator=g)
    full = O.ntxent_multiple_positives(z1, z2, tau=0.1)
    parts = [i3d.NTXentMultiplePositives(tau=0.1)(z1[r * 8:(r + 1) * 8].to(DEV), z2.to(DEV), row_offset=r * 8, total_rows=B)
             for r in range(3)]
    out.append(("NTXentMultiplePositives/row_offset_sum", abs(sum(p.item() for p in parts) - full.item()) / abs(full.item()),
                2e-6))
    return out


# -------------------------------------------------------------------------------------------- optimizer
def case_adam():
    g = gen(13)
    out = []
    shapes = [(200, 600), (200,), (7, 3), (1,), (256, 200)]
    params = [torch.randn(*s, generator=g) for s in shapes]
    grads = [[torch.randn(*s, generator=g) * 0.01 for s in shapes] for _ in range(4)]
    ref_p = [p.clone().requires_grad_(True) for p in params]
    ref_opt = torch.optim.Adam([{"params": ref_p[:2], "weight_decay": 0}, {"params": ref_p[2:]}], lr=8e-5)
    dev_p = [torch.nn.Parameter(p.clone().to(DEV)) for p in params]
    for graph_safe in (False, True):
        for p, q in zip(dev_p, params):
            p.data = q.clone().to(DEV)
        for p, q in zip(ref_p, params):
            p.data.copy_(q)
        ref_opt = torch.optim.Adam([{"params": ref_p[:2], "weight_decay": 0}, {"params": ref_p[2:]}], lr=8e-5)
        opt = i3d.FusedAdam([{"params": dev_p[:2], "weight_decay": 0}, {"params": dev_p[2:]}], lr=8e-5,
                            graph_safe=graph_safe)
        for step in range(4):
            lr = 8e-5 * (step + 1) / 4                       # WarmUpWrapper-style lr changes every step
            for grp in ref_opt.param_groups + opt.param_groups:
                grp["lr"] = lr
            for p, q, gr in zip(ref_p, dev_p, grads[step]):
                p.grad = gr.clone()
                q.grad = gr.clone().to(DEV)
            ref_opt.step()
            opt.step()
            opt.zero_grad()
        err = max(rel(q, p) for p, q in zip(ref_p, dev_p))
        upd = max(rel(q.detach().cpu() - p0, p.detach() - p0) for p, q, p0 in zip(ref_p, dev_p, params))
        out += [("adam/graph_safe=%s/params" % graph_safe, err, 1e-6), ("adam/graph_safe=%s/update" % graph_safe, upd, 2e-3)]
    sd = opt.state_dict()
    out.append(("adam/state_dict_exp_avg", rel(sd["state"][0]["exp_avg"], ref_opt.state_dict()["state"][0]["exp_avg"]), 1e-5))
    return out


# -------------------------------------------------------------------------------------------- FC operator
def case_fc():
    """ops.fc over gathered / scaled K-segments against cat + Linear + act + BatchNorm in plain torch."""
    g = gen(14)
    out = []
    n, e, Fd = 300, 700, 40
    src, dst = random_graph(15, n, e, isolated=False)
    st = i3d.GraphStructure(src.to(DEV), dst.to(DEV), torch.tensor([n]), n)
    eid = st.eid.long().cpu()
    src_c, dst_c = src[eid], dst[eid]
    h = torch.randn(n, Fd, generator=g).requires_grad_(True)
    ef = torch.randn(e, Fd, generator=g).requires_grad_(True)
    W = (torch.randn(Fd, 3 * Fd, generator=g) * 0.1).requires_grad_(True)
    b = (torch.randn(Fd, generator=g) * 0.1).requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(Fd, generator=g)).requires_grad_(True)
    bet = (0.1 * torch.randn(Fd, generator=g)).requires_grad_(True)
    rm, rv = torch.zeros(Fd), torch.ones(Fd)
    ref = F.batch_norm(torch.relu(F.linear(torch.cat([h[src_c], h[dst_c], ef], -1), W, b)), rm.clone(), rv.clone(), gam, bet,
                       True, 0.9, 1e-5)
    gout = torch.randn(e, Fd, generator=g)
    ref.backward(gout)
    d = lambda t: t.detach().to(DEV).requires_grad_(True)
    hd, efd, Wd, bd, gd, btd = d(h), d(ef), d(W), d(b), d(gam), d(bet)
    bn = (gd, btd, rm.to(DEV), rv.to(DEV), torch.zeros((), dtype=torch.long, device=DEV), 0.9, 1e-5)
    got = ops.fc([ops.Seg(hd, idx=st.src_csr, inv_rowptr=st.out_rowptr, inv_idx=st.out_pos),
                  ops.Seg(hd, idx=st.dst_csr, inv_rowptr=st.rowptr), ops.Seg(efd)], Wd, bd, K.ACT["relu"], bn, True)
    got.backward(gout.to(DEV))
    out += [("fc/gather/fwd", rel(got, ref), 2e-5), ("fc/gather/dh", rel(hd.grad, h.grad), 5e-5),
            ("fc/gather/def", rel(efd.grad, ef.grad), 5e-5), ("fc/gather/dW", rel(Wd.grad, W.grad), 5e-5),
            ("fc/gather/dgamma", rel(gd.grad, gam.grad), 5e-5), ("fc/gather/dbeta", rel(btd.grad, bet.grad), 5e-5),
            ("fc/gather/db", rel(bd.grad, b.grad), 5e-5)]
    # posttrans-style: cat[h, A, A*amp, A*att] with per-row scalers, no activation, BN, residual
    A = torch.randn(n, 4 * Fd, generator=g).requires_grad_(True)
    amp, att = st.amp.cpu(), st.att.cpu()
    W2 = (torch.randn(Fd, 13 * Fd, generator=g) * 0.05).requires_grad_(True)
    b2 = torch.zeros(Fd, requires_grad=True)
    h2 = torch.randn(n, Fd, generator=g).requires_grad_(True)
    x = torch.cat([h2, A, A * amp[:, None], A * att[:, None]], -1)
    ref = F.batch_norm(F.linear(x, W2, b2), rm.clone(), rv.clone(), gam, bet, True, 0.9, 1e-5) + h2
    gout = torch.randn(n, Fd, generator=g)
    gam.grad = bet.grad = None
    ref.backward(gout)
    Ad, W2d, b2d, h2d, gd, btd = d(A), d(W2), d(b2), d(h2), d(gam), d(bet)
    bn = (gd, btd, rm.to(DEV), rv.to(DEV), torch.zeros((), dtype=torch.long, device=DEV), 0.9, 1e-5)
    got = ops.fc([ops.Seg(h2d), ops.Seg(Ad), ops.Seg(Ad, scale=st.amp), ops.Seg(Ad, scale=st.att)], W2d, b2d,
                 K.ACT["none"], bn, True, residual=h2d)
    got.backward(gout.to(DEV))
    out += [("fc/scaled/fwd", rel(got, ref), 2e-5), ("fc/scaled/dA", rel(Ad.grad, A.grad), 5e-5),
            ("fc/scaled/dh", rel(h2d.grad, h2.grad), 5e-5), ("fc/scaled/dW", rel(W2d.grad, W2.grad), 5e-5)]
    return out


# ------------------------------------------------------------------------------- degree-merged posttrans
from gpu_cases_plan_ref import degree_plan_ref as _degree_plan_ref  # noqa: E402


def case_degree_plan():
    out = []
    cases = {"bond_like": (3000, 6100, None), "tiny": (5, 9, None), "one_bucket": (300, 0, 1),
             "overflow": (400, 3000, 4)}
    for tag, (n, e, nb) in cases.items():
        src, dst = random_graph(3, n, e) if e else (torch.zeros(0, dtype=torch.long),) * 2
        rowptr, _, _ = O.csr_reference(src.numpy(), dst.numpy(), n)
        NB = nb if nb is not None else int(np.diff(rowptr).max()) + 1
        plan = K.DegreePlan(torch.from_numpy(np.asarray(rowptr, dtype=np.int32)).to(DEV), NB)
        perm, tb, ch, over = _degree_plan_ref(rowptr, NB, plan.CHUNK_TILES)
        out += [("degree_plan/%s/perm" % tag, exact(plan.perm, perm), 0),
                ("degree_plan/%s/tile_bucket" % tag, exact(plan.tile_bucket, tb), 0),
                ("degree_plan/%s/chunk_tab" % tag, exact(plan.chunk_tab, ch), 0),
                ("degree_plan/%s/overflow" % tag, abs(int(plan.overflow.item()) - over), 0)]
    b = syn.make_batch(5, 512)
    rowptr, _, _ = O.csr_reference(b["src"], b["dst"], len(b["x_atom"]))
    NB = int(np.diff(rowptr).max()) + 1
    plan = K.DegreePlan(torch.from_numpy(np.asarray(rowptr, dtype=np.int32)).to(DEV), NB)
    perm, tb, ch, over = _degree_plan_ref(rowptr, NB, plan.CHUNK_TILES)
    out += [("degree_plan/qm9_b512/perm", exact(plan.perm, perm), 0),
            ("degree_plan/qm9_b512/chunk_tab", exact(plan.chunk_tab, ch), 0),
            ("degree_plan/qm9_b512/every_node_once",
             exact(np.sort(plan.perm.cpu().numpy()[plan.perm.cpu().numpy() >= 0]), np.arange(len(b["x_atom"]))), 0)]
    return out


def _fc_merged_once(tag, n, e, Fd, seed, act, train, tol_f, tol_g):
    """FCLayer over cat[h, A, A*amp, A*att] (models/pna.py:207-211,232) through the degree-merged weights against the
    concatenation evaluated in fp64 torch."""
    g = gen(seed)
    src, dst = random_graph(seed + 1, n, e, isolated=True)
    rowptr, _, _ = O.csr_reference(src.numpy(), dst.numpy(), n)
    md = int(np.diff(rowptr).max())
    st = _with_env({"I3D_PLAN_MIN_NODES": "0"}, i3d.GraphStructure, src.to(DEV), dst.to(DEV), torch.tensor([n]), n,
                   max_in_degree=md)
    assert st.plan is not None, "max in-degree %d does not fit the degree plan" % md
    amp, att = st.amp.cpu().double(), st.att.cpu().double()
    mk = lambda *shape, s=1.0: (torch.randn(*shape, generator=g) * s)
    A, h2 = mk(n, 4 * Fd), mk(n, Fd)
    W2, b2 = mk(Fd, 13 * Fd, s=0.05), mk(Fd, s=0.1)
    gam, bet = 1 + 0.1 * mk(Fd), 0.1 * mk(Fd)
    rm, rv = 0.1 * mk(Fd), 1 + 0.1 * mk(Fd).abs()
    gout = mk(n, Fd)
    leaf = lambda t: t.double().requires_grad_(True)
    A64, h64, W64, b64, g64, bt64 = leaf(A), leaf(h2), leaf(W2), leaf(b2), leaf(gam), leaf(bet)
    x = torch.cat([h64, A64, A64 * amp[:, None], A64 * att[:, None]], -1)
    y = F.linear(x, W64, b64)
    y = torch.relu(y) if act == "relu" else y
    rm64, rv64 = rm.double().clone(), rv.double().clone()
    ref = F.batch_norm(y, rm64, rv64, g64, bt64, train, 0.9, 1e-5) + h64
    ref.backward(gout.double())
    d = lambda t: t.detach().float().to(DEV).requires_grad_(True)
    Ad, hd, Wd, bd, gd, btd = d(A), d(h2), d(W2), d(b2), d(gam), d(bet)
    rmd, rvd = rm.to(DEV).clone(), rv.to(DEV).clone()
    bn = (gd, btd, rmd, rvd, torch.zeros((), dtype=torch.long, device=DEV), 0.9, 1e-5)
    merged = K.MergedPosttransWeights(Fd, Fd, st.plan.n_buckets, DEV)
    got = ops.fc_post_merged(st.plan, merged, hd, Ad, Wd, bd, K.ACT[act], bn, train, residual=hd)
    got.backward(gout.to(DEV))
    st.check_plan()
    out = [("%s/fwd" % tag, rel(got, ref), tol_f), ("%s/dA" % tag, rel(Ad.grad, A64.grad), tol_g),
           ("%s/dh" % tag, rel(hd.grad, h64.grad), tol_g), ("%s/dW" % tag, rel(Wd.grad, W64.grad), tol_g),
           # under train-mode BN without activation db is mathematically 0: measure against the column sums of |dO|
           ("%s/db" % tag, float((bd.grad.cpu().double() - b64.grad).abs().max() / gout.abs().sum(0).max()), tol_g),
           ("%s/dgamma" % tag, rel(gd.grad, g64.grad), tol_g), ("%s/dbeta" % tag, rel(btd.grad, bt64.grad), tol_g)]
    if train:
        # the batch variance inherits ~2x the relative error of the 3xTF32 GEMM output it is computed from
        out += [("%s/running_mean" % tag, rel(rmd, rm64), 1e-5), ("%s/running_var" % tag, rel(rvd, rv64), 5e-5)]
    # the generic 13F-wide path on the same inputs (same kernels family, different association): both within tolerance
    Ae, he, We, be_, ge, bte = d(A), d(h2), d(W2), d(b2), d(gam), d(bet)
    bn2 = (ge, bte, rm.to(DEV).clone(), rv.to(DEV).clone(), torch.zeros((), dtype=torch.long, device=DEV), 0.9, 1e-5)
    gen_out = ops.fc([ops.Seg(he), ops.Seg(Ae), ops.Seg(Ae, scale=st.amp), ops.Seg(Ae, scale=st.att)], We, be_,
                     K.ACT[act], bn2, train, residual=he)
    out.append(("%s/generic_path_fwd" % tag, rel(gen_out, ref), tol_f))
    return out


def case_fc_merged():
    out = []
    # db under train-mode BN is mathematically zero (bias feeds a BatchNorm): compared at a looser, absolute-ish level
    out += _fc_merged_once("fc_merged/small_train", 300, 700, 40, 50, "none", True, 2e-5, 5e-5)
    out += _fc_merged_once("fc_merged/small_eval_relu", 300, 700, 40, 52, "relu", False, 2e-5, 5e-5)
    out += _fc_merged_once("fc_merged/pna_width_train", 9400, 19500, 200, 54, "none", True, 3e-5, 1e-4)
    return out


def _with_env(env, fn, *a, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn(*a, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def case_golden_merged():
    """the golden vectors of the reference's own modules with the degree-merged posttrans path forced on (the
    8-molecule cases are below the node threshold at which it switches on by itself)"""
    res = _with_env({"I3D_PLAN_MIN_NODES": "0", "I3D_CHECK_PLAN": "1"}, case_golden, "qm9_b8")
    res += _with_env({"I3D_PLAN_MIN_NODES": "0", "I3D_CHECK_PLAN": "1"}, case_golden, "qmugs_b6_c3")
    return [(l.replace("golden/", "golden_merged/"), e, t) for l, e, t in res]


def case_train_steps_merged():
    """training steps against the oracle with the degree-merged posttrans AND the weight-gradient side stream forced
    on (both switch on by themselves only above a size threshold that a 16-molecule batch does not reach)"""
    env = {"I3D_PLAN_MIN_NODES": "0", "I3D_DW_MIN_ROWS": "0"}
    res = _with_env(env, case_train_steps, 16, 2, False, 23)
    res += _with_env(env, case_train_steps, 16, 2, True, 23)
    return [(l.replace("train", "train_merged", 1), e, t) for l, e, t in res]


def case_dw_side_stream():
    """BASELINE config-2 size: parameter gradients of one backward pass with the weight-gradient GEMMs on the side
    stream equal the single-stream ones (split-K atomics make both runs differ at rounding level only)."""
    b = syn.make_batch(9, 512)
    grads = {}
    for mode in ("1", "0"):
        def run():
            c2, c3, st2, st3, pna, n3 = _models(71, 72)
            tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=0.1), DEV, {"lr": 8e-5})
            g2, g3 = i3d.batch_from_numpy(b, DEV)
            loss, _, _ = tr.forward_pass(([g2], [g3]))
            loss.backward()
            torch.cuda.synchronize()
            return loss.item(), {k: p.grad.detach().clone() for k, p in list(pna.named_parameters()) + list(n3.named_parameters())}
        grads[mode] = _with_env({"I3D_DW_STREAM": mode}, run)
    (l1, g1), (l0, g0) = grads["1"], grads["0"]
    scale = max(float(v.abs().max()) for v in g0.values())
    worst = max(float((g1[k] - g0[k]).abs().max()) for k in g0) / scale
    nonzero = min(float(g1[k].abs().max() > 0) for k in g1 if "linear.weight" in k)
    return [("dw_side_stream/loss", abs(l1 - l0), 1e-6), ("dw_side_stream/param_grads_vs_single_stream", worst, 2e-5),
            ("dw_side_stream/every_weight_gradient_written", 1.0 - nonzero, 0)]


# ------------------------------------------------------------------------------------ device collate (N1)
def case_collate():
    """i3d_collate_2d / i3d_collate_3d against the numpy restatement of QM9Dataset.__getitem__ + contrastive_collate
    (oracle/collate_oracle.py, pinned on the reference's code) and against the committed reference vectors."""
    from oracle import collate_oracle as CO
    out = []
    keys = ("src", "dst", "x_atom", "e_attr", "num_nodes", "num_edges", "src3", "dst3", "num_nodes3", "num_edges3")

    def compare(tag, g2, g3, ref):
        got = {"src": g2.edges()[0], "dst": g2.edges()[1], "x_atom": g2.ndata["feat"], "e_attr": g2.edata["feat"],
               "num_nodes": g2.batch_num_nodes(), "num_edges": g2.batch_num_edges(), "src3": g3.edges()[0],
               "dst3": g3.edges()[1], "num_nodes3": g3.batch_num_nodes(), "num_edges3": g3.batch_num_edges()}
        res = [("collate/%s/%s" % (tag, k), exact(got[k], ref[k]), 0) for k in keys]
        # fp32 distance: same fma chain + sqrt as torch.norm on the CPU; bit-equal expected, 1 ulp allowed
        res.append(("collate/%s/d3" % tag, rel(g3.edata["d"], ref["d3"]), 1.2e-7))
        res.append(("collate/%s/d3_bit_equal_fraction_missing" % tag,
                    float((g3.edata["d"].cpu().numpy() != ref["d3"]).mean()), 0.0))
        return res

    for name, shape in (("collate_qm9", "qm9"), ("collate_qmugs", "qmugs")):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        store = CO.make_store(int(g["seed"]), int(g["n_molecules"]), shape)
        ps = i3d.PackedMoleculeStore(store, DEV)
        out += compare("golden_" + name, *ps.collate(g["idx"]), g)
    # BASELINE config-2 size: 512 random molecules out of a 2000-molecule store, with repeats and a 1-atom... (n>=3 in QM9)
    store = CO.make_store(91, 2000, "qm9")
    ps = i3d.PackedMoleculeStore(store, DEV)
    idx = np.random.default_rng(3).integers(0, 2000, size=512)
    g2, g3 = ps.collate(idx)
    ref = CO.collate_reference(store, idx)
    out += compare("qm9_b512", g2, g3, ref)
    out.append(("collate/qm9_b512/max_in_degree_hint", abs(g2.max_in_degree - int(np.bincount(ref["dst"]).max())), 0))
    # the collated batch drives the encoders exactly like a batch built on the host
    c2, c3, st2, st3, pna, n3 = _models(61, 62)
    pna.eval(), n3.eval()
    with torch.no_grad():
        z2, z3 = pna(g2), n3(g3)
        h2, h3 = i3d.batch_from_numpy(dict(ref, batch_size=512, conformers=1), DEV)
        y2, y3 = pna(h2), n3(h3)
    out += [("collate/qm9_b512/pna_on_device_batch_equals_host_batch", rel(z2, y2), 1e-6),
            ("collate/qm9_b512/net3d_on_device_batch_equals_host_batch", rel(z3, y3), 1e-6)]
    # a second batch of the same size reuses the staging buffer: the first batch's graphs must stay intact
    src_before = g2.edges()[0].clone()
    bnn_before = g2.batch_num_nodes().clone()
    ps.collate(np.random.default_rng(4).integers(0, 2000, size=512))
    torch.cuda.synchronize()
    out += [("collate/staging_reuse_keeps_previous_batch", exact(g2.edges()[0], src_before) +
             exact(g2.batch_num_nodes(), bnn_before), 0)]
    return out


# ------------------------------------------------------------------------------------ contrastive metrics (N3)
def case_contrastive_metrics():
    """i3d_contrastive_metrics against the oracle (pinned on trainer/metrics.py) and the committed reference vectors.
    Similarities: 1e-5 absolute (3xTF32 GEMM vs fp32 einsum).  Rates are counts of threshold decisions: a similarity
    within rounding distance of the threshold may flip, so they are compared to within 3 decisions."""
    from oracle.pin_metrics import CASES, embeddings
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    thr = float(g["threshold"])
    out = []
    for name, (seed, B, D, noisy) in CASES.items():
        x1, x2 = embeddings(seed, B, D, noisy)
        ref = torch.from_numpy(g[name + "/ref"]).double()
        got = i3d.contrastive_metrics(x1.to(DEV), x2.to(DEV), thr).double().cpu()
        flip = 3.0 / (B * (B - 1))
        out += [("metrics/%s/positive_similarity" % name, abs(float(got[0] - ref[0])), 1e-5),
                ("metrics/%s/negative_similarity" % name, abs(float(got[1] - ref[1])), 1e-5),
                ("metrics/%s/true_positive_rate" % name, abs(float(got[2] - ref[2])), 3.0 / B),
                ("metrics/%s/true_negative_rate" % name, abs(float(got[3] - ref[3])), flip),
                ("metrics/%s/contrastive_accuracy" % name, abs(float(got[4] - ref[4])), 0.5 * (3.0 / B + flip))]
    # the other four logged metrics (dimension_covariance, batch_variance, alignment, uniformity) against the vectors of
    # the reference's own classes; normalised embeddings give a finite uniformity (unnormalised ones underflow to -inf in
    # fp32, in the reference and here alike)
    for name, (seed, B, D, noisy) in CASES.items():
        x1, x2 = embeddings(seed, B, D, noisy)
        ref4 = g[name + "/ref4"]
        got4 = i3d.embedding_metrics(x1.to(DEV), x2.to(DEV), 2.0).double().cpu().numpy()
        for j, nm in enumerate(("dimension_covariance", "batch_variance", "alignment")):
            out.append(("metrics/%s/%s" % (name, nm), abs(got4[j] - ref4[j]) / abs(ref4[j]), 2e-5))
        if np.isfinite(ref4[3]):
            out.append(("metrics/%s/uniformity" % name, abs(got4[3] - ref4[3]), 1e-3))
        else:
            out.append(("metrics/%s/uniformity(-inf like the reference)" % name, float(got4[3] != ref4[3]), 0))
        n1, n2 = torch.nn.functional.normalize(x1, dim=1), torch.nn.functional.normalize(x2, dim=1)
        want = O.embedding_metrics(n1, n2, 2)
        got = i3d.embedding_metrics(n1.to(DEV), n2.to(DEV), 2.0).cpu()
        out.append(("metrics/%s/normalised_embeddings(all four, relative)" % name,
                    float(((got - want).abs() / want.abs()).max()), 5e-5))
    a4, b4 = embeddings(12, 128, 64, 0)
    a4, b4 = torch.nn.functional.normalize(a4, dim=1).to(DEV), torch.nn.functional.normalize(b4, dim=1).to(DEV)
    vals4 = torch.stack([i3d.DimensionCovariance()(a4, b4), i3d.BatchVariance()(a4, b4), i3d.Alignment(alpha=2)(a4, b4),
                         i3d.Uniformity(t=2)(a4, b4)]).cpu()
    want4 = O.embedding_metrics(a4.cpu(), b4.cpu(), 2)
    out.append(("metrics/modules4_vs_oracle", float(((vals4 - want4).abs() / want4.abs().clamp(min=1e-6)).max()), 5e-5))
    # module surface of trainer/metrics.py: five metric objects share one fused evaluation
    x1, x2 = embeddings(9, 256, 64, 0)
    a, b = x1.to(DEV), x2.to(DEV)
    n0 = i3d.lib.launch_count()
    vals = [i3d.PositiveSimilarity()(a, b), i3d.NegativeSimilarity()(a, b), i3d.TruePositiveRate(thr)(a, b),
            i3d.TrueNegativeRate(thr)(a, b), i3d.ContrastiveAccuracy(thr)(a, b)]
    launches = i3d.lib.launch_count() - n0
    ref = O.contrastive_metrics(x1, x2, thr)
    out.append(("metrics/modules_vs_oracle", float((torch.stack(vals).cpu() - ref).abs().max()), 1e-4))
    # the two threshold-free metrics share one fused evaluation, the three thresholded ones another (6 launches each:
    # 2 row norms, weight split, GEMM, row pass, final reduction) — instead of five separate [B,B] einsum chains
    out.append(("metrics/five_metrics_two_evaluations(launches<=12)", float(max(0, launches - 12)), 0))
    # a NEW batch of embeddings at the SAME address (what the caching allocator hands the trainer for the next batch's
    # predictions) must not be served the previous batch's cached metrics
    m = i3d.PositiveSimilarity()
    ya, yb = embeddings(10, 256, 64, 0)
    ta, tb = ya.to(DEV), yb.to(DEV)
    first = float(m(ta, tb))
    pa, pb = ta.data_ptr(), tb.data_ptr()
    del ta, tb
    za, zb = embeddings(11, 256, 64, 0)
    ua, ub = za.to(DEV), zb.to(DEV)
    same_addr = ua.data_ptr() == pa and ub.data_ptr() == pb
    second = float(m(ua, ub))
    want = float(O.contrastive_metrics(za, zb, 0.5)[0])
    out.append(("metrics/next_batch_at_the_same_address_is_recomputed(same address: %s)" % same_addr, abs(second - want), 1e-5))
    out.append(("metrics/and_differs_from_the_previous_batch", float(abs(first - second) < 1e-7), 0))
    return out


# ------------------------------------------------------------------------------------ tower PNA (N2 / P11)
def case_pna_original():
    """PNAOriginal (towers, scalar avg_d, graph_norm, LeakyReLU mixing, MLPReadout) against the vectors emitted by the
    reference's own models/pna_original.py (oracle/pin_pna_original.py) and the CPU oracle.  Tolerances as for PNA:
    embeddings 1e-4 relative, parameter gradients 1e-3 of the global gradient scale."""
    from oracle import pna_original_oracle as PO
    from oracle.pin_pna_original import CASES as PCASES, snorm
    from oracle.make_golden import grad_fingerprint
    out = []
    for name, (bseed, B, shape, wseed, avg_d, c) in PCASES.items():
        gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        b = syn.make_batch(bseed, B, shape=shape)
        st = PO.init_state(c, wseed)
        sn = snorm(b["num_nodes"])
        for mode in ("eval", "train"):
            kw = {k: v for k, v in c.items() if k != "gru"}
            m = i3d.PNAOriginal(avg_d=avg_d, device=DEV, **kw)
            missing = m.load_state_dict(st, strict=True)
            m = m.to(DEV)
            m.train(mode == "train")
            g2, _ = i3d.batch_from_numpy(b, DEV)
            z = m(g2, sn)
            tag = "%s/%s" % (name, mode)
            out.append((tag + "/z", rel(z, gold["z_" + mode]), 1e-4))
            if mode == "train":
                w = torch.randn(z.shape, generator=torch.Generator().manual_seed(5)).to(DEV)
                (z * w).sum().backward()
                named = dict(m.named_parameters())
                scale = float(gold["grad_scale"])
                worst = 0.0
                for k, fp in zip(gold["grad_keys"], gold["grad_fp"]):
                    mine = grad_fingerprint(named[str(k)].grad.cpu())
                    worst = max(worst, float(np.abs(mine[2:] - fp[2:]).max()) / scale)
                out.append((tag + "/param_grads(all, sampled)", worst, 1e-3))
                sd = m.state_dict()
                for k in gold.files:
                    if k.startswith("buf/"):
                        out.append((tag + "/" + k, rel(sd[k[4:]], gold[k]), 1e-4))
        out.append((name + "/state_dict_keys_identical", float(len(missing.missing_keys) + len(missing.unexpected_keys)), 0))
    # the "hidden 200, 4 towers" shape runs FUSED by default (one block-diagonal layer on the tensor-core path, checked
    # against the reference vectors above); here: the tower-by-tower path on the same weights, and the fused path on a
    # shape-bucketed (padded) batch against the unpadded one
    name = "pna_original_h200_t4"
    bseed, B, shape, wseed, avg_d, c = PCASES[name]
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    kw = {k: v for k, v in c.items() if k != "gru"}
    st = PO.init_state(c, wseed)
    b = syn.make_batch(bseed, B, shape=shape)
    m = i3d.PNAOriginal(avg_d=avg_d, device=DEV, **kw)
    m.load_state_dict(st, strict=True)
    m = m.to(DEV).train()
    out.append((name + "/fused_path_is_the_default", float(not all(L._fusable() for L in m.node_gnn.layers)), 0))
    for L in m.node_gnn.layers:
        L.fuse_towers = False
    g2, _ = i3d.batch_from_numpy(b, DEV)
    out.append((name + "/train/z(tower-by-tower path)", rel(m(g2, snorm(b["num_nodes"])), gold["z_train"]), 1e-4))
    # padded vs unpadded, fused, train mode (BatchNorm statistics must ignore the padding rows), incl. gradients
    store = syn.make_store(77, 200, "qm9")
    ps = i3d.PackedMoleculeStore(store, DEV)
    idx = np.random.default_rng(3).integers(0, 200, size=40)
    res = {}
    for how in ("plain", "padded"):
        m = i3d.PNAOriginal(avg_d=avg_d, device=DEV, **kw)
        m.load_state_dict(st, strict=True)
        m = m.to(DEV).train()
        if how == "plain":
            g2, _ = ps.collate(idx)
        else:
            meta, (N, E, E3) = ps.stage_metadata(idx)
            g2, _ = ps.collate_padded(meta, len(idx), N + 200, E + 300, E3 + 128, need_3d=False)
        z = m(g2, i3d.BucketedStep._snorm(g2))
        w = torch.randn(z.shape, generator=torch.Generator().manual_seed(6)).to(DEV)
        (z * w).sum().backward()
        res[how] = (z.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None},
                    {k: v.detach().clone() for k, v in m.state_dict().items() if "running_" in k})
    out.append((name + "/padded_vs_plain/z", rel(res["padded"][0], res["plain"][0]), 2e-5))
    # (two fp32 evaluations with different summation orders — sort-based vs structure-emitting collate, other tile
    #  boundaries: isolated ReLU sign flips move single gradient rows, see gpu_cases_bucketed.py; the relative L2 error
    #  over all parameters is the tight bound, a padding leak would show there at the 1e-1 level)
    gs = max(float(v.abs().max()) for v in res["plain"][1].values())
    worst = max(res["plain"][1], key=lambda k: float((res["padded"][1][k] - res["plain"][1][k]).abs().max()))
    out.append((name + "/padded_vs_plain/param_grads(all, global scale; worst: %s)" % worst,
                float((res["padded"][1][worst] - res["plain"][1][worst]).abs().max()) / gs, 2e-2))
    num = sum(float((res["padded"][1][k] - v).double().pow(2).sum()) for k, v in res["plain"][1].items())
    den = sum(float(v.double().pow(2).sum()) for v in res["plain"][1].values())
    out.append((name + "/padded_vs_plain/param_grads(all parameters, relative L2)", (num / den) ** 0.5, 5e-3))
    out.append((name + "/padded_vs_plain/bn_running_stats",
                max(rel(res["padded"][2][k], v) for k, v in res["plain"][2].items()), 1e-5))
    return out


# ------------------------------------------------------------------------------------ fine-tuning head (config 5)
def case_finetune_config():
    """BASELINE config 5 (configs_clean/tune_QM9_homo.yml): PNA with the fine-tuning head (readout min/max/mean/sum,
    target_dim 1, BN momentum 0.1) and L1Loss, against the vectors of the reference's own PNA (oracle/pin_finetune.py)."""
    from oracle.pin_finetune import CASE, TUNE_QM9_HOMO, targets
    from oracle.make_golden import grad_fingerprint
    name, bseed, B, wseed = CASE
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    b = syn.make_batch(bseed, B)
    c = O.pna_cfg(**TUNE_QM9_HOMO)
    st = O.init_pna_state(c, wseed, True)
    y = targets(B).to(DEV)
    out = []
    for mode in ("eval", "train"):
        m = i3d.PNA(avg_d=1, device=DEV, **TUNE_QM9_HOMO)
        m.load_state_dict(st)
        m = m.to(DEV)
        m.train(mode == "train")
        g2, _ = i3d.batch_from_numpy(b, DEV)
        z = m(g2)
        loss = torch.nn.L1Loss()(z, y)                    # the reference's loss_func for this config is torch's own
        tag = "finetune/%s" % mode
        out += [(tag + "/prediction", rel(z, gold["z_" + mode]), 1e-4),
                (tag + "/l1_loss", abs(loss.item() - float(gold["loss_" + mode])), 1e-5)]
        if mode == "train":
            loss.backward()
            named = dict(m.named_parameters())
            scale = float(gold["grad_scale"])
            worst = 0.0
            for k, fp in zip(gold["grad_keys"], gold["grad_fp"]):
                mine = grad_fingerprint(named[str(k)].grad.cpu())
                worst = max(worst, float(np.abs(mine[2:] - fp[2:]).max()) / scale)
            # Ground truth: the oracle in float64.  12 molecules (~220 nodes) through 7 train-mode BatchNorm layers and
            # an L1 loss (gradient = sign / B): the CPU fp32 reference's own distance from the truth sets the scale, and
            # the golden fp32 vectors are checked at 1e-3 or twice that distance.
            def oracle_grads(dt):
                s = {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in st.items()}
                s = O.as_leaf_params(s)
                og2, xa, ea, _, _ = O.graphs_from_batch(b)
                zz = O.pna_forward(s, c, og2, xa, ea, True)
                torch.nn.L1Loss()(zz, y.cpu().to(dt)).backward()
                return {k: s[k].grad for k in O.param_keys(s)}

            g64, g32 = oracle_grads(torch.float64), oracle_grads(torch.float32)
            sc = max(float(v.abs().max()) for v in g64.values())
            e_cuda = max(float((named[k].grad.cpu().double() - g64[k]).abs().max()) for k in g64) / sc
            e_cpu = max(float((g32[k].double() - g64[k]).abs().max()) for k in g64) / sc
            out.append((tag + "/param_grads(all, sampled; vs the reference's fp32 vectors)", worst, max(1e-3, 2 * e_cpu + e_cuda)))
            out.append((tag + "/param_grads(all tensors in full; cuda vs fp64 truth)", e_cuda, max(1e-3, 2 * e_cpu)))
            out.append((tag + "/param_grads(all tensors in full; cpu fp32 oracle vs fp64 truth, for scale)", e_cpu, float("inf")))
    return out


# ------------------------------------------------------------------------------------------ whole models
def _models(s2, s3, trained=True):
    c2, c3 = O.pna_cfg(**O.PRETRAIN_QM9_PNA), O.net3d_cfg(**O.PRETRAIN_QM9_NET3D)
    st2, st3 = O.init_pna_state(c2, s2, trained), O.init_net3d_state(c3, s3, trained)
    pna = i3d.PNA(avg_d=1, device=DEV, **O.PRETRAIN_QM9_PNA)
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **O.PRETRAIN_QM9_NET3D)
    pna.load_state_dict(st2)
    n3.load_state_dict(st3)
    return c2, c3, st2, st3, pna.to(DEV), n3.to(DEV)


def case_golden(name="qm9_b8"):
    """CUDA path against the vectors emitted by the reference's own modules (tests/golden)."""
    from oracle.make_golden import CASES, TAU, grad_fingerprint
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    bseed, B, shape, C, loss_name, (s2, s3) = CASES[name]
    b = syn.make_batch(bseed, B, shape=shape, conformers=C)
    out = []
    for mode in ("eval", "train"):
        c2, c3, st2, st3, pna, n3 = _models(s2, s3)
        pna.train(mode == "train"), n3.train(mode == "train")
        g2, g3 = i3d.batch_from_numpy(b, DEV)
        z2, z3 = pna(g2), n3(g3)
        loss = getattr(i3d, loss_name)(tau=TAU)(z2, z3)
        tag = "golden/%s/%s" % (name, mode)
        out += [(tag + "/z2d", rel(z2, gold["z2d_" + mode]), 1e-4), (tag + "/z3d", rel(z3, gold["z3d_" + mode]), 1e-4),
                (tag + "/loss", abs(loss.item() - float(gold["loss_" + mode])), 1e-4)]
        if mode == "train":
            st = g2._i3d_struct
            out += [(tag + "/csr_rowptr", exact(st.rowptr, gold["csr_rowptr"]), 0),
                    (tag + "/csr_col", exact(st.src_csr, gold["csr_col"]), 0),
                    (tag + "/csr_eid", exact(st.eid, gold["csr_eid"]), 0)]
            loss.backward()
            scale = float(gold["grad_scale"])
            named = dict([("2d." + k, p) for k, p in pna.named_parameters()] + [("3d." + k, p) for k, p in n3.named_parameters()])
            worst = 0.0
            for k, fp in zip(gold["grad_keys"], gold["grad_fp"]):
                mine = grad_fingerprint(named[str(k)].grad.cpu())
                worst = max(worst, float(np.abs(mine[2:] - fp[2:]).max()) / scale)
            out.append((tag + "/param_grads(all, sampled)", worst, 1e-3))
            for k in gold.files:
                if k.startswith("grad3d/"):
                    out.append((tag + "/" + k, float(np.abs(named["3d." + k[7:]].grad.cpu().numpy() - gold[k]).max()) / scale, 1e-3))
                if k.startswith("grad2d/"):
                    out.append((tag + "/" + k, float(np.abs(named["2d." + k[7:]].grad.cpu().numpy() - gold[k]).max()) / scale, 1e-3))
                if k.startswith("buf2d/"):
                    out.append((tag + "/" + k, rel(pna.state_dict()[k[6:]], gold[k]), 1e-4))
                if k.startswith("buf3d/"):
                    out.append((tag + "/" + k, rel(n3.state_dict()[k[6:]], gold[k]), 1e-4))
    return out


def case_golden_qmugs():
    return case_golden("qmugs_b6_c3")


def case_train_steps(B=16, steps=3, captured=False, seed=21):
    """Three optimisation steps (fwd, bwd, gradient pack, Adam) against the CPU oracle trainer.

    Adam's first steps move every weight by ~lr*sign(g): a weight whose gradient is at rounding level flips sign
    between ANY two fp32 implementations and the trajectories separate chaotically (measured: 0.05 % of weights after
    one step, which moves the next step's embeddings by 6e-3).  So each step is checked on its own: same loss, same
    update for (almost) all weights, then the oracle's parameters are copied over the CUDA model's before the next step
    (Adam moments are NOT copied: they stay consistent only if every step's gradients were right)."""
    c2, c3, st2, st3, pna, n3 = _models(31, 32)
    otr = O.OracleTrainer(c2, c3, st2, st3, loss="NTXent", tau=0.1, lr=8e-5)
    tr = i3d.SelfSupervisedTrainer(pna, n3, i3d.NTXent(tau=0.1), DEV, {"lr": 8e-5}, graph_safe=captured)
    named = dict([("2d." + k, p) for k, p in pna.named_parameters()] + [("3d." + k, p) for k, p in n3.named_parameters()])
    oparam = lambda k: (otr.st2d if k.startswith("2d.") else otr.st3d)[k[3:]]
    # gradient mathematically zero (a bias that feeds a BatchNorm): Adam turns rounding noise into +-lr steps
    zero_grad = ("pretrans.fully_connected.1.linear.bias", "posttrans.fully_connected.0.linear.bias",
                 "pretrans.fully_connected.0.batch_norm.bias", "update_network.fully_connected.0.linear.bias")

    def sync_params():
        with torch.no_grad():
            for k, p in named.items():
                p.copy_(oparam(k).detach())

    out = []
    tag = "train_captured" if captured else "train"
    b = syn.make_batch(seed, B)
    cap = None
    if captured:
        # CapturedStep's eager warm-up step on the example batch leaves no trace (weights, Adam state and BatchNorm
        # buffers are restored before the capture): nothing to mirror in the oracle
        g2, g3 = i3d.batch_from_numpy(b, DEV)
        cap = i3d.CapturedStep(tr, g2, g3, warmup=1)
    for s in range(steps):
        if not captured:
            b = syn.make_batch(seed + s, B)
        before = {k: oparam(k).detach().clone() for k in named}
        ol, _, _ = otr.step(*O.graphs_from_batch(b))
        g2, g3 = i3d.batch_from_numpy(b, DEV)
        if captured:
            cap.load(g2, g3)
            l = cap.run()
        else:
            l, _, _ = tr.process_batch(([g2], [g3]))
        bad = total = 0
        worst_big = 0.0
        for k, p in named.items():
            if any(z in k for z in zero_grad):
                continue
            ref = oparam(k).detach()
            err = (p.detach().cpu() - ref).abs() / 8e-5                     # in units of one Adam step
            bad += int((err > 0.05).sum())
            total += err.numel()
            moved = (ref - before[k]).abs() > 0.5 * 8e-5                     # weights with a decisive gradient
            if moved.any():
                worst_big = max(worst_big, float(err[moved].median()))
        out += [("%s/step%d/loss" % (tag, s), abs(l.item() - ol.item()), 2e-5),
                ("%s/step%d/fraction_of_weights_off_by_>5%%_of_lr" % (tag, s), bad / max(total, 1), 1e-2),
                ("%s/step%d/median_update_error_in_lr_units" % (tag, s), worst_big, 1e-2)]
        sync_params()
    return out


def case_train_steps_captured():
    return case_train_steps(captured=True)


def case_full_size_properties():
    """BASELINE config-2 size (B=512): properties that do not need the oracle to finish in seconds."""
    out = []
    b = syn.make_batch(5, 512)
    c2, c3, st2, st3, pna, n3 = _models(41, 42)
    pna.eval(), n3.eval()
    with torch.no_grad():
        g2, g3 = i3d.batch_from_numpy(b, DEV)
        z2, z3 = pna(g2), n3(g3)
        # molecules are independent in eval mode: the first 8 molecules alone give the same embeddings
        nb = {k: v for k, v in b.items()}
        n8, e8 = int(b["num_nodes"][:8].sum()), int(b["num_edges"][:8].sum())
        n38, e38 = int(b["num_nodes3"][:8].sum()), int(b["num_edges3"][:8].sum())
        nb.update(x_atom=b["x_atom"][:n8], e_attr=b["e_attr"][:e8], src=b["src"][:e8], dst=b["dst"][:e8],
                  num_nodes=b["num_nodes"][:8], num_edges=b["num_edges"][:8], src3=b["src3"][:e38], dst3=b["dst3"][:e38],
                  d3=b["d3"][:e38], num_nodes3=b["num_nodes3"][:8], num_edges3=b["num_edges3"][:8])
        h2, h3 = i3d.batch_from_numpy(nb, DEV)
        y2, y3 = pna(h2), n3(h3)
        out += [("full/block_independence_2d", rel(y2, z2[:8]), 1e-4), ("full/block_independence_3d", rel(y3, z3[:8]), 2e-5)]
        # oracle on the 8-molecule slice
        og2, xa, ea, og3, d3 = O.graphs_from_batch(nb)
        r2 = O.pna_forward(O.as_leaf_params(st2), c2, og2, xa, ea, False)
        r3 = O.net3d_forward(O.as_leaf_params(st3), c3, og3, d3, False)
        out += [("full/slice_vs_oracle_2d", rel(y2, r2), 1e-4), ("full/slice_vs_oracle_3d", rel(y3, r3), 1e-4)]
        st = g2._i3d_struct
        rowptr, col, eid = O.csr_reference(b["src"], b["dst"], len(b["x_atom"]))
        out += [("full/csr_eid", exact(st.eid, eid), 0), ("full/csr_rowptr", exact(st.rowptr, rowptr), 0)]
        loss = i3d.NTXent(tau=0.1)(z2, z3)
        out.append(("full/loss_vs_oracle_loss_on_gpu_embeddings",
                    abs(loss.item() - O.ntxent(z2.cpu(), z3.cpu(), 0.1).item()), 1e-4))
    return out


from gpu_cases_dp import case_sharded_equals_full  # noqa: E402
from gpu_cases_bucketed import (case_bucketed_step, case_bucketed_step_conformers, case_collate_struct,  # noqa: E402
                                case_edge_factored, case_epoch_many_shapes, case_finetune_step, case_inference, case_staging_ring, case_step_b512,
                                case_step_config3)

ALL_CASES = [case_csr, case_embed, case_gemm, case_gemm_tc, case_weight_prep, case_bn, case_aggregate, case_segment_ops, case_net3d_elementwise,
             case_ntxent, case_adam, case_fc, case_degree_plan, case_fc_merged, case_golden, case_golden_qmugs,
             case_golden_merged, case_train_steps, case_train_steps_captured, case_train_steps_merged,
             case_dw_side_stream, case_full_size_properties, case_collate, case_contrastive_metrics, case_pna_original, case_finetune_config, case_sharded_equals_full,
             case_collate_struct, case_edge_factored, case_staging_ring, case_bucketed_step, case_bucketed_step_conformers, case_step_b512,
             case_step_config3, case_epoch_many_shapes, case_inference, case_finetune_step]
