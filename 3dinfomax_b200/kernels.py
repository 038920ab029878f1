"""Typed tensor-level wrappers around the C ABI (one Python function per entry point of include/i3d.h).

torch is used here for what the task calls plumbing only: device allocations (``torch.empty``),
the current CUDA stream and raw ``data_ptr()``s.  No wrapper computes anything with torch ops.
"""
import ctypes
import os

import torch

from . import lib as _lib

ACT = {"none": 0, "relu": 1, "silu": 2, "leakyrelu": 3}
RO = {"sum": 0, "mean": 1, "max": 2, "min": 3}
NT, NN, TN = 0, 1, 2


STATS_PREZEROED = 0x100      # I3D_STATS_PREZEROED (include/i3d.h)
STATS_STRIDE = 16            # I3D_STATS_STRIDE: one fp64 accumulator per 128-byte line


class StatsArena:
    """fp64 scratch for the BatchNorm statistics of ONE training step, zeroed by ONE memset at the start of the step.

    Every FC layer needs a zeroed [2F] fp64 buffer in the forward (column sums from the GEMM epilogue) and another in
    the backward; zeroing each one separately costs ~50 memset nodes per step, each of which also breaks programmatic
    dependent launch between its neighbours.  ``take`` hands out consecutive slices (host-side bump pointer); the
    slices are only valid until the next ``reset`` and never escape the operator that took them."""

    def __init__(self, device, doubles=1 << 20, scratch_bytes=192 << 20, counters=1 << 12):
        self.buf = torch.zeros(doubles, dtype=torch.float64, device=device)
        self.used = 0
        self.lock = __import__("threading").Lock()
        # workspace of the two-stage column reductions (i3d_reduce_ws): per-CTA partial slots (never initialised) and
        # ticket counters (zeroed ONCE: every kernel leaves its counter at 0), handed out by bump pointers per step
        on_gpu = torch.device(device).type == "cuda"
        self.scratch = torch.empty(scratch_bytes if on_gpu else 0, dtype=torch.uint8, device=device)
        self.counters = torch.zeros(counters, dtype=torch.int32, device=device)
        self.s_used = self.c_used = 0
        self.slot_ctas = 4 * (torch.cuda.get_device_properties(device).multi_processor_count if on_gpu else 148)

    def reset(self):
        self.buf.zero_()                       # one memset node (8 MB) per step
        self.used = 0
        self.s_used = self.c_used = 0

    def take_ws(self, ncols, elem_bytes):
        """i3d_reduce_ws for a column reduction over ``ncols`` columns, or None when the scratch is exhausted"""
        n = (self.slot_ctas * ncols * elem_bytes + 255) // 256 * 256
        with self.lock:
            if self.s_used + n > self.scratch.numel() or self.c_used >= self.counters.numel():
                return None
            ws = _lib.reduce_ws(self.scratch.data_ptr() + self.s_used, n, self.counters.data_ptr() + 4 * self.c_used)
            self.s_used += n
            self.c_used += 1
            return ws

    def take(self, n):
        with self.lock:
            if self.used + n > self.buf.numel():
                return None                    # exhausted: the caller falls back to its own buffer + memset
            out = self.buf[self.used:self.used + n]
            self.used += n
            return out


def _stats_buffer(arena, n, device):
    """(buffer, flag) for ``n`` column statistics in the library's strided layout (n * STATS_STRIDE doubles): a pre-zeroed
    arena slice when there is one, else a fresh tensor the library zeroes itself"""
    n = n * STATS_STRIDE
    t = arena.take(n) if arena is not None else None
    if t is not None:
        return t, STATS_PREZEROED
    return torch.empty(n, dtype=torch.float64, device=device), 0


def _ws(arena, ncols, elem_bytes):
    """(ctypes pointer to an i3d_reduce_ws or None, keep-alive object)"""
    # measured on B200 (batch 512): the single-CTA second stage pulls ~1 MB of slots through one SM (~10 us) and costs
    # more than the atomic tail it removes (bn_bwd_reduce 16.3 -> 26.3 us), so the workspace is opt-in (I3D_TWO_STAGE=1)
    if arena is None or os.environ.get("I3D_TWO_STAGE", "0") != "1":
        return None, None
    ws = arena.take_ws(ncols, elem_bytes)
    return (ctypes.byref(ws) if ws is not None else None), ws


def _L():
    return _lib.load()


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the 3dinfomax_b200 path has no CPU fallback" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))


def _mat(t, name):
    """2-D fp32 tensor with unit inner stride -> (ptr, ld)."""
    _req(t, torch.float32, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("%s must be 2-D with unit inner stride" % name)
    return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))


def _valid(t):
    """device int32 scalar holding the number of valid leading rows of a padded (shape-bucketed) batch, or None"""
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.int32 or t.numel() != 1:
        raise TypeError("valid-row count must be a 1-element int32 CUDA tensor")
    return t.data_ptr()


def _vec(t, dtype, name):
    _req(t, dtype, name)
    if t is not None and not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return _p(t)


# ---------------------------------------------------------------------------------------- graph
def csr_build(key, other, n_rows):
    """Stable CSR by ``key``.  Returns (rowptr[N+1], col[E], rowid[E], eid[E]) int32."""
    E = key.numel()
    dev = key.device
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    rowid = torch.empty(E, dtype=torch.int32, device=dev)
    eid = torch.empty(E, dtype=torch.int32, device=dev)
    ws = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
    if key.dtype == torch.int64:
        fn = _L().i3d_csr_build
        _vec(key, torch.int64, "key"), _vec(other, torch.int64, "other")
    else:
        fn = _L().i3d_csr_build_i32
        _vec(key, torch.int32, "key"), _vec(other, torch.int32, "other")
    _lib.check(fn(_p(key), _p(other), E, n_rows, _p(rowptr), _p(col), _p(rowid), _p(eid), _p(ws), _s()),
               "i3d_csr_build")
    return rowptr, col, rowid, eid


def segment_ptr(counts):
    _vec(counts, torch.int64, "counts")
    B = counts.numel()
    ptr = torch.empty(B + 1, dtype=torch.int32, device=counts.device)
    _lib.check(_L().i3d_segment_ptr(_p(counts), B, _p(ptr), _s()), "i3d_segment_ptr")
    return ptr


def degree_scalers(rowptr, avg_d=None):
    """amp = ln(D+1) [/ avg_d], att = [avg_d] / ln(D+1); avg_d None = the hard-coded 1.0 of models/pna.py:153"""
    N = rowptr.numel() - 1
    amp = torch.empty(N, dtype=torch.float32, device=rowptr.device)
    att = torch.empty(N, dtype=torch.float32, device=rowptr.device)
    if avg_d is None:
        _lib.check(_L().i3d_degree_scalers(_p(rowptr), N, _p(amp), _p(att), _s()), "i3d_degree_scalers")
    else:
        _lib.check(_L().i3d_degree_scalers_avg(_p(rowptr), N, float(avg_d), _p(amp), _p(att), _s()),
                   "i3d_degree_scalers_avg")
    return amp, att


def scale_rows(x, s):
    """y[m, :] = x[m, :] * s[m]"""
    px, ldx = _mat(x, "x")
    _vec(s, torch.float32, "s")
    M, F = x.shape
    if s.numel() != M:
        raise ValueError("one scale per row expected")
    y = torch.empty(M, F, dtype=torch.float32, device=x.device)
    _lib.check(_L().i3d_scale_rows(px, ldx, _p(s), M, F, _p(y), F, _s()), "i3d_scale_rows")
    return y


class DegreePlan:
    """Device arrays of i3d_degree_plan (include/i3d.h): nodes grouped by in-degree into whole 128-row tiles."""
    CHUNK_TILES = 4        # smallest split-K chunk of the weight-gradient GEMM (4 tiles = 512 virtual rows)

    def __init__(self, rowptr, n_buckets, ctas_per_chunk=8, sms=None):
        N = rowptr.numel() - 1
        dev = rowptr.device
        if sms is None:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
        tiles = (N + 127) // 128
        self.n_buckets = int(n_buckets)
        # the chunked dW GEMM launches ctas_per_chunk CTAs (output tiles of the [Fout, 4F] block) per chunk: size the
        # chunks so that all of them fit one wave of the machine
        room = max(1, sms // ctas_per_chunk - self.n_buckets)
        self.CHUNK_TILES = max(DegreePlan.CHUNK_TILES, -(-tiles // room))
        self.N = N
        self.T = tiles + self.n_buckets
        self.Mv = 128 * self.T
        self.CH = (tiles + self.CHUNK_TILES - 1) // self.CHUNK_TILES + self.n_buckets
        self.perm = torch.empty(self.Mv, dtype=torch.int32, device=dev)
        self.tile_bucket = torch.empty(self.T, dtype=torch.int32, device=dev)
        self.chunk_tab = torch.empty(3 * self.CH, dtype=torch.int32, device=dev)
        self.overflow = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(_L().i3d_degree_plan(_vec(rowptr, torch.int32, "rowptr"), N, self.n_buckets, self.CHUNK_TILES,
                                        _p(self.perm), _p(self.tile_bucket), _p(self.chunk_tab), _p(self.overflow),
                                        _s()), "i3d_degree_plan")


def kpad32(k):
    return (k + 31) // 32 * 32


class MergedPosttransWeights:
    """Persistent tf32 hi/lo scratch of the degree-merged posttrans weights of ONE layer (i3d_posttrans_merge)."""

    def __init__(self, Fout, F, n_buckets, device):
        self.Fout, self.F, self.n_buckets = int(Fout), int(F), int(n_buckets)
        # row pitch of the operands: the padded K extent, bumped off multiples of 2 KB (e.g. 1024 floats at F = 200)
        # so that the 200-odd rows of a TMA box do not all map to the same L2 slice
        unpow2 = lambda k: k + 32 if (k * 4) % 2048 == 0 else k
        self.ktf = unpow2(kpad32(F) + kpad32(4 * F)) if os.environ.get("I3D_MERGED_PITCH", "pad") == "pad" else kpad32(F) + kpad32(4 * F)
        self.ktb = kpad32(Fout)
        z = lambda r, c: torch.zeros(r, c, dtype=torch.float32, device=device)     # pad columns stay zero for good
        self.fwd_hi, self.fwd_lo = z(n_buckets * Fout, self.ktf), z(n_buckets * Fout, self.ktf)
        self.bwd_hi, self.bwd_lo = z(n_buckets * 5 * F, self.ktb), z(n_buckets * 5 * F, self.ktb)

    def refresh(self, W):
        pw, ldw = _mat(W, "W")
        if W.shape != (self.Fout, 13 * self.F):
            raise ValueError("posttrans weight must be [Fout, 13F]")
        _lib.check(_L().i3d_posttrans_merge(pw, ldw, self.Fout, self.F, self.n_buckets, _p(self.fwd_hi),
                                            _p(self.fwd_lo), self.ktf, _p(self.bwd_hi), _p(self.bwd_lo), self.ktb, _s()),
                   "i3d_posttrans_merge")


def gemm_nt_bucketed(plan, N, segs, C, bias, b_hi, b_lo, stats_act=None, arena=None, valid=None):
    """C[plan.perm[m], :] = bias + sum_s A_s[a_idx_s[m] or m, :] @ B[bucket(m)]^T over the plan's virtual rows.
    valid: optional device int32 scalar — output rows >= valid[0] (padding of a bucketed batch) stay out of the stats."""
    b_pitch = b_hi.shape[1]
    pc, ldc = _mat(C, "C")
    arr = _seg_array(segs, need_b=False)
    stats, flag = (None, 0)
    if stats_act is not None:
        stats, flag = _stats_buffer(arena, 2 * N, C.device)
    _lib.check(_L().i3d_gemm_nt_bucketed_v(plan.Mv, N, len(segs), arr, pc, ldc, _vec(bias, torch.float32, "bias"),
                                           _p(b_hi), _p(b_lo), b_pitch, plan.n_buckets, _p(plan.tile_bucket),
                                           _p(plan.perm), _p(stats), 0 if stats_act is None else (stats_act | flag),
                                           _valid(valid), _s()), "i3d_gemm_nt_bucketed")
    return C if stats_act is None else (C, stats)


def gemm_tn_chunked(plan, A, B, C):
    """C[b] += sum_{virtual rows m of bucket b} A[perm[m], :]^T B[perm[m], :];  C [n_buckets, A.cols, B.cols] zeroed."""
    pa, lda = _mat(A, "A")
    pb, ldb = _mat(B, "B")
    M, N = A.shape[1], B.shape[1]
    if C.shape != (plan.n_buckets, M, N) or not C.is_contiguous():
        raise ValueError("C must be a contiguous [n_buckets, %d, %d] tensor" % (M, N))
    seg = (_lib.gemm_seg * 1)()
    seg[0].A, seg[0].B, seg[0].a_idx, seg[0].b_idx = pa, pb, _p(plan.perm), _p(plan.perm)
    seg[0].scale = None
    seg[0].K, seg[0].lda, seg[0].ldb = plan.Mv, int(lda), int(ldb)
    _lib.check(_L().i3d_gemm_tn_chunked(M, N, seg, _p(C), N, M * N, _p(plan.chunk_tab), plan.CH, _s()),
               "i3d_gemm_tn_chunked")
    return C


def posttrans_unmerge(dWb, dW, F):
    nb, Fout, F4 = dWb.shape
    pd, ldw = _mat(dW, "dW")
    _lib.check(_L().i3d_posttrans_unmerge(_p(dWb), nb, Fout, F, pd, ldw, _s()), "i3d_posttrans_unmerge")


# ------------------------------------------------------------------------------------ embedding
def embed_sum_fwd(idx, col_off, perm, table):
    _vec(idx, torch.int64, "idx"), _vec(col_off, torch.int32, "col_off"), _vec(perm, torch.int32, "perm")
    _vec(table, torch.float32, "table")
    R = idx.shape[0] if perm is None else perm.numel()
    C, F = idx.shape[1], table.shape[1]
    out = torch.empty(R, F, dtype=torch.float32, device=table.device)
    _lib.check(_L().i3d_embed_sum_fwd(_p(idx), R, C, _p(col_off), _p(perm), _p(table), F, _p(out), _s()),
               "i3d_embed_sum_fwd")
    return out


def embed_sum_bwd(idx, col_off, perm, gout, n_table_rows, max_dim=0):
    gout = gout.contiguous()
    R, F = gout.shape
    C = idx.shape[1]
    gtable = torch.zeros(n_table_rows, F, dtype=torch.float32, device=gout.device)
    _lib.check(_L().i3d_embed_sum_bwd(_p(idx), R, C, _p(col_off), _p(perm), _p(gout), F, _p(gtable),
                                      int(n_table_rows) if max_dim else 0, int(max_dim), _s()),
               "i3d_embed_sum_bwd")
    return gtable


# ----------------------------------------------------------------------------------------- gemm
def _seg_array(segs, need_b=True):
    arr = (_lib.gemm_seg * len(segs))()
    for i, s in enumerate(segs):
        pa, lda = _mat(s["A"], "A")
        arr[i].A = pa
        if need_b:
            arr[i].B, ldb = _mat(s["B"], "B")
            arr[i].ldb = int(ldb)
        arr[i].a_idx = _vec(s.get("a_idx"), torch.int32, "a_idx")
        arr[i].b_idx = _vec(s.get("b_idx"), torch.int32, "b_idx")
        arr[i].scale = _vec(s.get("scale"), torch.float32, "scale")
        arr[i].K, arr[i].lda = int(s["K"]), int(lda)
    return arr


def gemm(mode, M, N, segs, C, bias=None, accumulate=False, stats_act=None, prepared=None, arena=None, valid=None):
    """segs: list of dicts {A, B, K, a_idx?, b_idx?, scale?}; A/B are 2-D views (their stride(0) is the ld).
    stats_act: activation code -> also returns fp64 [2N] column sums of act(C), act(C)^2 (fused BatchNorm statistics).
    prepared: a ready ``PreparedB`` (NT only): B comes from its scratch, the segments need no "B"."""
    pc, ldc = _mat(C, "C")
    L = _L()
    stats, flag = (None, 0)
    if stats_act is not None:
        stats, flag = _stats_buffer(arena, 2 * N, C.device)
    if prepared is not None:
        if mode != NT:
            raise ValueError("prepared operands are for NT GEMMs")
        arr = _seg_array(segs, need_b=False)
        _lib.check(L.i3d_gemm_nt_prepared_v(M, N, len(segs), arr, pc, ldc, _vec(bias, torch.float32, "bias"),
                                            1 if accumulate else 0, _p(prepared.ws), _p(stats),
                                            0 if stats_act is None else (stats_act | flag), _valid(valid), _s()),
                   "i3d_gemm_nt_prepared")
        return C if stats_act is None else (C, stats)
    arr = _seg_array(segs)
    nws = int(L.i3d_gemm_ws_bytes(mode, M, N, len(segs), arr))
    ws = torch.empty(nws, dtype=torch.uint8, device=C.device) if nws else None      # tf32 hi/lo copies of B for TMA
    _lib.check(L.i3d_gemm_ex_v(mode, M, N, len(segs), arr, pc, ldc, _vec(bias, torch.float32, "bias"),
                               1 if accumulate else 0, _p(ws), nws, _p(stats),
                               0 if stats_act is None else (stats_act | flag), _valid(valid), _s()), "i3d_gemm")
    return C if stats_act is None else (C, stats)


def gemm_nt_prepared_ok(M, N, segs):
    return bool(_L().i3d_gemm_nt_prepared_ok(M, N, len(segs), _seg_array(segs, need_b=False)))


class PreparedB:
    """Persistent tf32 hi/lo scratch of one NT GEMM's B operand (a weight, or column blocks of it read transposed)."""
    __slots__ = ("W", "ws", "items", "n_items", "epoch", "w_version")


class WeightPrep:
    """Registry of prepared weight operands refreshed by ONE kernel launch per optimizer step.

    ``entry`` is called by ops._FC with the weight views a GEMM would otherwise split on the fly; ``refresh`` (start
    of a step) rewrites every registered scratch from the current weights; ``invalidate`` (after an optimizer step
    that bypasses torch's version counter, i.e. FusedAdam) marks them stale.  An entry is used only while
    ``ready(entry)``: same epoch AND the weight's autograd version unchanged, so a foreign in-place update of the
    weight silently falls back to the per-call split instead of computing with stale copies."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.entries = {}
        self.epoch = 0
        self._table = None
        self._old_tables = []
        self._pinned = []
        self._tiles = 0

    def entry(self, W, key, N, b_views, transposed):
        """b_views: list of (2-D view of W as the kernel would read it, K).  transposed: views are [K, N]."""
        k = (W.data_ptr(), key, bool(transposed))
        e = self.entries.get(k)
        if e is not None:
            return e
        arr = (_lib.gemm_seg * len(b_views))()
        for i, (b, kk) in enumerate(b_views):
            arr[i].B, ldb = _mat(b, "B")
            arr[i].ldb, arr[i].K = int(ldb), int(kk)
        ktot = sum((kk + 31) // 32 * 32 for _, kk in b_views)
        e = PreparedB()
        e.W = W
        e.ws = torch.empty(2 * N * ktot * 4 + 256, dtype=torch.uint8, device=self.device)
        e.items = (_lib.prep_item * len(b_views))()
        e.n_items = len(b_views)
        tiles = ctypes.c_int(0)
        _lib.check(_L().i3d_gemm_prep_describe(N, len(b_views), arr, 1 if transposed else 0, _p(e.ws), self._tiles,
                                               e.items, ctypes.byref(tiles)), "i3d_gemm_prep_describe")
        self._tiles += tiles.value
        e.epoch, e.w_version = -1, -1
        self.entries[k] = e
        if self._table is not None:
            self._old_tables.append(self._table)            # graphs captured earlier still point at their table
        self._table = None
        return e

    def entry_stacked(self, W, key, views):
        """Prepared B operand made of several weight blocks stacked along N (rows of the operand): ``views`` are 2-D
        views [n_j, K] of W with the same K.  Used for P = h [W_s; W_d]^T (ops._FCEdgeFactored): one GEMM, N = sum n_j."""
        k = (W.data_ptr(), ("stacked",) + tuple(key), False)
        e = self.entries.get(k)
        if e is not None:
            return e
        Kd = int(views[0].shape[1])
        ktot = kpad32(Kd)
        ntot = sum(int(v.shape[0]) for v in views)
        e = PreparedB()
        e.W = W
        e.ws = torch.empty(2 * ntot * ktot * 4 + 256, dtype=torch.uint8, device=self.device)
        base = (e.ws.data_ptr() + 127) & ~127               # same layout as i3d_gemm_prep_describe / gemm_ws_nt
        lo_base = base + ntot * ktot * 4
        e.items = (_lib.prep_item * len(views))()
        e.n_items = len(views)
        row0, tiles = 0, 0
        for j, v in enumerate(views):
            pb, ldb = _mat(v, "B")
            n = int(v.shape[0])
            it = e.items[j]
            it.B, it.hi, it.lo = pb, base + row0 * ktot * 4, lo_base + row0 * ktot * 4
            it.ldb, it.N, it.K, it.kpad, it.ldo, it.col0 = int(ldb), n, Kd, ktot, ktot, 0
            it.transposed, it.tile0 = 0, self._tiles + tiles
            tiles += ((n + 31) // 32) * (ktot // 32)
            row0 += n
        self._tiles += tiles
        e.epoch, e.w_version = -1, -1
        self.entries[k] = e
        if self._table is not None:
            self._old_tables.append(self._table)
        self._table = None
        return e

    def ready(self, e):
        return e.epoch == self.epoch and e.w_version == e.W._version

    def invalidate(self):
        self.epoch += 1

    def refresh(self):
        if not self.entries:
            return
        ents = list(self.entries.values())
        if self._table is None:
            if torch.cuda.is_current_stream_capturing():
                # the upload below would be recorded instead of executed, and the table would live in the capturing
                # graph's private pool: any later EAGER refresh would then run on a table that was never filled
                raise RuntimeError("WeightPrep: the operand table must be built outside CUDA-graph capture — run one "
                                   "eager forward and call refresh() before capturing (CapturedStep / BucketedStep do)")
            raw = b"".join(bytes(e.items) for e in ents)
            host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
            self._pinned.append(host)                       # a captured graph re-reads it on replay
            self._table = torch.empty(len(raw), dtype=torch.uint8, device=self.device)
            self._table.copy_(host, non_blocking=True)
            self._n_items = sum(e.n_items for e in ents)
        _lib.check(_L().i3d_gemm_prep_run(_p(self._table), self._n_items, self._tiles, _s()), "i3d_gemm_prep_run")
        for e in ents:
            e.epoch, e.w_version = self.epoch, e.W._version


def edge_gather_add(P, src, dst, T, code, bias, Y, stats_act=None, arena=None, valid=None):
    """Y[m] = P[src[m], :F] + P[dst[m], F:2F] + T[code[m]] + bias (i3d_edge_gather_add); returns the fp64 [2F] column
    sums of act(Y), act(Y)^2 when ``stats_act`` is given."""
    pp, ldp = _mat(P, "P")
    py, ldy = _mat(Y, "Y")
    M, F = Y.shape
    pt, ldt = (None, 0) if T is None else _mat(T, "T")
    stats, flag = (None, 0)
    ws = keep = None
    if stats_act is not None:
        stats, flag = _stats_buffer(arena, 2 * F, Y.device)
        ws, keep = _ws(arena, 2 * F, 8)
    _lib.check(_L().i3d_edge_gather_add(pp, ldp, _vec(src, torch.int32, "src"), _vec(dst, torch.int32, "dst"), pt, ldt,
                                        _vec(code, torch.int32, "code"), _vec(bias, torch.float32, "bias"), M, F, py, ldy,
                                        _p(stats), 0 if stats_act is None else (stats_act | flag), _valid(valid), ws,
                                        _s()), "i3d_edge_gather_add")
    return stats


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _bond_check(combo, Ws, col0):
    _vec(combo, torch.float32, "combo")
    n_codes, F = combo.shape
    Fout, ldw = Ws[0].shape[0], Ws[0].stride(0)
    for W in Ws:
        _req(W, torch.float32, "W")
        if W.dim() != 2 or W.stride(1) != 1 or W.shape[0] != Fout or W.stride(0) != ldw or W.shape[1] < col0 + F:
            raise ValueError("bond tables: the layers' weights must share shape and leading dimension")
    return n_codes, F, Fout, ldw


def bond_tables_fwd(combo, Ws, col0):
    """T_l = combo @ W_l[:, col0:col0+F]^T for every layer in one launch (i3d_bond_tables_fwd); returns the list of T_l"""
    n_codes, F, Fout, ldw = _bond_check(combo, Ws, col0)
    Ts = [torch.empty(n_codes, Fout, dtype=torch.float32, device=combo.device) for _ in Ws]
    _lib.check(_L().i3d_bond_tables_fwd(_p(combo), n_codes, F, len(Ws), _ptr_array(Ws), ldw, col0, Fout, _ptr_array(Ts),
                                        _s()), "i3d_bond_tables_fwd")
    return Ts


def bond_tables_bwd(combo, Ws, col0, dTs, dWs, want_dcombo=True):
    """dW_l[:, col0:col0+F] += dT_l^T combo (layers whose dT_l / dW_l is None are skipped); returns dcombo or None"""
    n_codes, F, Fout, ldw = _bond_check(combo, Ws, col0)
    for t in dTs:
        if t is not None:
            _vec(t, torch.float32, "dT")
    for t in dWs:
        if t is not None and (t.stride(0) != ldw or t.stride(1) != 1):
            raise ValueError("bond tables: a weight gradient must share the weight's layout")
    dcombo = torch.zeros(n_codes, F, dtype=torch.float32, device=combo.device) if want_dcombo else None
    _lib.check(_L().i3d_bond_tables_bwd(_p(combo), n_codes, F, len(Ws), _ptr_array(Ws), ldw, col0, Fout,
                                        _ptr_array(dTs), _ptr_array(dWs), _p(dcombo), _s()), "i3d_bond_tables_bwd")
    return dcombo


def transpose(x):
    """[R, C] (unit inner stride) -> contiguous [C, R]"""
    px, ldx = _mat(x, "x")
    R, C = x.shape
    out = torch.empty(C, R, dtype=torch.float32, device=x.device)
    _lib.check(_L().i3d_transpose(px, R, C, ldx, _p(out), R, _s()), "i3d_transpose")
    return out


# ------------------------------------------------------------------------------ FC tail (act+BN)
def act_colstats(Y, act):
    py, ldy = _mat(Y, "Y")
    M, F = Y.shape
    sums = torch.empty(2 * F * STATS_STRIDE, dtype=torch.float64, device=Y.device)
    _lib.check(_L().i3d_act_colstats(py, M, F, ldy, act, _p(sums), _s()), "i3d_act_colstats")
    return sums


def bn_apply(Y, act, sums, running_mean, running_var, nbt, gamma, beta, momentum, eps, training, residual, out=None,
             valid=None):
    py, ldy = _mat(Y, "Y")
    M, F = Y.shape
    O = torch.empty(M, F, dtype=torch.float32, device=Y.device) if out is None else out
    po, ldo = _mat(O, "O")
    if residual is not None:
        pr, ldr = _mat(residual, "residual")
        if ldr != ldo:
            raise ValueError("residual must share the output's leading dimension")
    else:
        pr = None
    save = torch.empty(2 * F, dtype=torch.float32, device=Y.device)
    _lib.check(_L().i3d_bn_apply_v(py, M, F, ldy, act, _p(sums), _p(running_mean), _p(running_var), _p(nbt),
                                   _p(gamma), _p(beta), float(momentum), float(eps), 1 if training else 0, _p(save),
                                   pr, po, ldo, _valid(valid), _s()), "i3d_bn_apply")
    return O, save


DBIAS_STRIDE = 8       # bias-gradient accumulators: one float per 32-byte sector (see i3d_bn_bwd_apply_v)


def dbias_buffer(F, device):
    """accumulator for the bias gradient of an FC layer in the spread layout; zeroed by bn_bwd_reduce(zero=...)"""
    return torch.empty(F * DBIAS_STRIDE, dtype=torch.float32, device=device)


def bn_bwd_reduce(dO, Y, act, save, arena=None, zero=None, valid=None):
    """zero: optional fp32 tensor the kernel clears on the way (the dbias accumulator of the following bn_bwd_apply)"""
    pd, ldd = _mat(dO, "dO")
    py, ldy = _mat(Y, "Y")
    M, F = Y.shape
    sums2, flag = _stats_buffer(arena, 2 * F, Y.device)
    ws, keep = _ws(arena, 2 * F, 8)
    _lib.check(_L().i3d_bn_bwd_reduce_v(pd, ldd, py, ldy, M, F, act | flag, _p(save), _p(sums2), _p(zero),
                                        0 if zero is None else zero.numel(), _valid(valid), ws, _s()),
               "i3d_bn_bwd_reduce")
    return sums2


def bn_bwd_apply(dO, Y, act, has_bn, training, save, gamma, sums2, want_dbias=True, dbias_zeroed=None, valid=None,
                 arena=None):
    pd, ldd = _mat(dO, "dO")
    py, ldy = _mat(Y, "Y")
    M, F = Y.shape
    dev = Y.device
    dY = torch.empty(M, F, dtype=torch.float32, device=dev)
    db_stride = 1
    if not want_dbias:
        dbias = None
    elif dbias_zeroed is not None:
        dbias = dbias_zeroed                     # cleared by the preceding bn_bwd_reduce
        db_stride = dbias.numel() // F           # DBIAS_STRIDE floats per column when it comes from dbias_buffer()
    else:
        dbias = torch.zeros(F, dtype=torch.float32, device=dev)
    dgamma = torch.empty(F, dtype=torch.float32, device=dev) if has_bn else None
    dbeta = torch.empty(F, dtype=torch.float32, device=dev) if has_bn else None
    ws, keep = _ws(arena, F, 4) if dbias is not None else (None, None)
    _lib.check(_L().i3d_bn_bwd_apply_v(pd, ldd, py, ldy, M, F, act, 1 if has_bn else 0, 1 if training else 0,
                                       _p(save), _p(gamma), _p(sums2), _p(dY), F, _p(dbias), db_stride, _p(dgamma),
                                       _p(dbeta), _valid(valid), ws, _s()), "i3d_bn_bwd_apply")
    if dbias is not None and db_stride > 1:
        dbias = dbias.view(F, db_stride)[:, 0]   # strided view: FusedAdam's gradient pack reads it with its stride
    return dY, dbias, dgamma, dbeta


# streams on which the one-launch BatchNorm backward is NOT used: its grid barrier spins, so at most one stream per device
# may run it (two such grids beside each other could starve each other of SMs).  The trainer registers the 3-D encoder's
# stream here; the 2-D encoder's (main) stream takes the fused path.
NO_FUSED_BN_BWD_STREAMS = set()


def bn_bwd(dO, Y, act, training, save, gamma, want_dbias=True, dbias_zeroed=None, valid=None, arena=None):
    """Backward of activation -> BatchNorm: (dY, dbias, dgamma, dbeta).  ``bn_bwd_reduce`` + ``bn_bwd_apply`` in one
    call; with I3D_BN_BWD=fused one LAUNCH (i3d_bn_bwd_fused_v) under a trainer (``arena``) on a stream that may spin
    on a grid barrier (measured: 4.130 vs 4.146 ms per step, i.e. nothing; off by default)."""
    fused_ok = (arena is not None and os.environ.get("I3D_BN_BWD", "split") == "fused"
                and os.environ.get("I3D_TWO_STAGE", "0") != "1"
                and torch.cuda.current_stream(Y.device).cuda_stream not in NO_FUSED_BN_BWD_STREAMS)
    bar = arena.take(STATS_STRIDE) if fused_ok else None          # 128 zeroed bytes: the barrier's ticket counter
    if bar is None:
        sums2 = bn_bwd_reduce(dO, Y, act, save, arena=arena, zero=dbias_zeroed, valid=valid)
        return bn_bwd_apply(dO, Y, act, True, training, save, gamma, sums2, want_dbias, dbias_zeroed=dbias_zeroed,
                            valid=valid, arena=arena)
    pd, ldd = _mat(dO, "dO")
    py, ldy = _mat(Y, "Y")
    M, F = Y.shape
    dev = Y.device
    sums2, flag = _stats_buffer(arena, 2 * F, dev)
    dY = torch.empty(M, F, dtype=torch.float32, device=dev)
    db_stride = 1
    if not want_dbias:
        dbias = None
    elif dbias_zeroed is not None:
        dbias = dbias_zeroed                     # cleared by the kernel's first phase
        db_stride = dbias.numel() // F
    else:
        dbias = torch.zeros(F, dtype=torch.float32, device=dev)
    dgamma = torch.empty(F, dtype=torch.float32, device=dev)
    dbeta = torch.empty(F, dtype=torch.float32, device=dev)
    zero = dbias_zeroed
    _lib.check(_L().i3d_bn_bwd_fused_v(pd, ldd, py, ldy, M, F, act | flag, 1 if training else 0, _p(save), _p(gamma),
                                       _p(sums2), _p(dY), F, _p(dbias), db_stride, _p(dgamma), _p(dbeta), _p(zero),
                                       0 if zero is None else zero.numel(), _valid(valid), _p(bar), _s()),
               "i3d_bn_bwd_fused")
    if dbias is not None and db_stride > 1:
        dbias = dbias.view(F, db_stride)[:, 0]
    return dY, dbias, dgamma, dbeta


def act_fwd(x, act):
    _req(x, torch.float32, "x")
    x = x.contiguous()
    y = torch.empty_like(x)
    _lib.check(_L().i3d_act_fwd(_p(x), x.numel(), act, _p(y), _s()), "i3d_act_fwd")
    return y


def act_bwd(gy, x, act):
    _req(gy, torch.float32, "gy"), _req(x, torch.float32, "x")
    gy = gy.contiguous()
    gx = torch.empty_like(x)
    _lib.check(_L().i3d_act_bwd(_p(gy), _p(x), x.numel(), act, _p(gx), _s()), "i3d_act_bwd")
    return gx


# ------------------------------------------------------------------------------------ aggregation
def pna_aggregate_fwd(msg, rowptr):
    _vec(msg, torch.float32, "msg")
    N = rowptr.numel() - 1
    F = msg.shape[1]
    out = torch.empty(N, 4 * F, dtype=torch.float32, device=msg.device)
    _lib.check(_L().i3d_pna_aggregate_fwd(_p(msg), _p(rowptr), N, F, _p(out), 4 * F, _s()), "i3d_pna_aggregate_fwd")
    return out


def pna_aggregate_bwd(g, msg, out, rowptr):
    pg, ldg = _mat(g, "g")
    N = rowptr.numel() - 1
    F = msg.shape[1]
    dmsg = torch.empty_like(msg)
    _lib.check(_L().i3d_pna_aggregate_bwd(pg, ldg, _p(msg), _p(out), 4 * F, _p(rowptr), N, F, _p(dmsg), _s()),
               "i3d_pna_aggregate_bwd")
    return dmsg


def _ops_arr(ops):
    return (ctypes.c_int32 * len(ops))(*ops)


def segment_readout_fwd(x, ptr, ops):
    px, ldx = _mat(x, "x")
    B = ptr.numel() - 1
    F = x.shape[1]
    out = torch.empty(B, len(ops) * F, dtype=torch.float32, device=x.device)
    _lib.check(_L().i3d_segment_readout_fwd(px, ldx, _p(ptr), B, F, len(ops), ctypes.cast(_ops_arr(ops), ctypes.c_void_p),
                                            _p(out), _s()), "i3d_segment_readout_fwd")
    return out


def segment_readout_bwd(g, x, out, ptr, ops):
    g = g.contiguous()
    px, ldx = _mat(x, "x")
    B = ptr.numel() - 1
    F = x.shape[1]
    dx = torch.empty(x.shape[0], F, dtype=torch.float32, device=x.device)
    # rows behind the last graph (padding nodes of a bucketed batch) get a zero gradient
    _lib.check(_L().i3d_segment_readout_bwd_v(_p(g), px, ldx, _p(out), _p(ptr), B, F, len(ops),
                                              ctypes.cast(_ops_arr(ops), ctypes.c_void_p), _p(dx), F, x.shape[0], _s()),
               "i3d_segment_readout_bwd")
    return dx


def segment_sum_fwd(x, rowptr, idx=None, mean=False, addend=None):
    px, ldx = _mat(x, "x")
    N = rowptr.numel() - 1
    F = x.shape[1]
    out = torch.empty(N, F, dtype=torch.float32, device=x.device)
    pa, lda = (None, 0) if addend is None else _mat(addend, "addend")
    _lib.check(_L().i3d_segment_sum_fwd(px, ldx, _p(rowptr), _vec(idx, torch.int32, "idx"), N, F, 1 if mean else 0,
                                        pa, lda, _p(out), F, _s()), "i3d_segment_sum_fwd")
    return out


def segment_sum_bwd(g, rowptr, rowid, mean=False):
    g = g.contiguous()
    E = rowid.numel()
    F = g.shape[1]
    gx = torch.empty(E, F, dtype=torch.float32, device=g.device)
    _lib.check(_L().i3d_segment_sum_bwd(_p(g), _p(rowptr), _p(rowid), E, F, 1 if mean else 0, _p(gx), _s()),
               "i3d_segment_sum_bwd")
    return gx


# ----------------------------------------------------------------------------------------- net3d
def fourier_encode(dist, perm, k):
    _vec(dist, torch.float32, "dist")
    E = dist.numel() if perm is None else perm.numel()
    out = torch.empty(E, 2 * k + 1, dtype=torch.float32, device=dist.device)
    _lib.check(_L().i3d_fourier_encode(_p(dist), _vec(perm, torch.int32, "perm"), E, k, _p(out), _s()),
               "i3d_fourier_encode")
    return out


def soft_gate_fwd(msg, ws, bs):
    _vec(msg, torch.float32, "msg"), _vec(ws, torch.float32, "ws"), _vec(bs, torch.float32, "bs")
    E, H = msg.shape
    m = torch.empty_like(msg)
    w = torch.empty(E, dtype=torch.float32, device=msg.device)
    _lib.check(_L().i3d_soft_gate_fwd(_p(msg), E, H, _p(ws), _p(bs), _p(m), _p(w), _s()), "i3d_soft_gate_fwd")
    return m, w


def soft_gate_bwd(gm, msg, w, ws):
    gm = gm.contiguous()
    E, H = msg.shape
    gmsg = torch.empty_like(msg)
    gws = torch.zeros(H, dtype=torch.float32, device=msg.device)
    gbs = torch.zeros(1, dtype=torch.float32, device=msg.device)
    _lib.check(_L().i3d_soft_gate_bwd(_p(gm), _p(msg), _p(w), E, H, _p(ws), _p(gmsg), _p(gws), _p(gbs), _s()),
               "i3d_soft_gate_bwd")
    return gmsg, gws, gbs


def broadcast_rows(vec, M):
    _vec(vec, torch.float32, "vec")
    F = vec.numel()
    out = torch.empty(M, F, dtype=torch.float32, device=vec.device)
    _lib.check(_L().i3d_broadcast_rows(_p(vec), M, F, _p(out), _s()), "i3d_broadcast_rows")
    return out


def colsum(x):
    px, ldx = _mat(x, "x")
    M, F = x.shape
    out = torch.empty(F, dtype=torch.float32, device=x.device)
    _lib.check(_L().i3d_colsum(px, ldx, M, F, _p(out), _s()), "i3d_colsum")
    return out


def add(a, b):
    _req(a, torch.float32, "a"), _req(b, torch.float32, "b")
    a, b = a.contiguous(), b.contiguous()
    y = torch.empty_like(a)
    _lib.check(_L().i3d_add(_p(a), _p(b), a.numel(), _p(y), _s()), "i3d_add")
    return y


def add_rows(a, b):
    """a + b for two [M, F] fp32 matrices that may be column slices of wider buffers (unit column stride)"""
    pa, lda = _mat(a, "a")
    pb, ldb = _mat(b, "b")
    M, F = a.shape
    if tuple(b.shape) != (M, F):
        raise ValueError("add_rows: shapes differ")
    y = torch.empty(M, F, dtype=torch.float32, device=a.device)
    _lib.check(_L().i3d_add_rows(pa, lda, pb, ldb, M, F, _p(y), F, _s()), "i3d_add_rows")
    return y


# ------------------------------------------------------------------------------------------ loss
def row_norms(z):
    _vec(z, torch.float32, "z")
    R, D = z.shape
    out = torch.empty(R, dtype=torch.float32, device=z.device)
    _lib.check(_L().i3d_row_norms(_p(z), R, D, _p(out), _s()), "i3d_row_norms")
    return out


def ntxent_rows_fwd(P, B, Bc, C, n1, n2, norm, eps, tau, row_offset):
    rowstats = torch.empty(B, 2, dtype=torch.float32, device=P.device)
    loss_rows = torch.empty(B, dtype=torch.float32, device=P.device)
    _lib.check(_L().i3d_ntxent_rows_fwd(_p(P), B, Bc, C, _p(n1), _p(n2), 1 if norm else 0, float(eps), float(tau),
                                        int(row_offset), _p(rowstats), _p(loss_rows), _s()), "i3d_ntxent_rows_fwd")
    return rowstats, loss_rows


def sum_scaled(x, scale):
    out = torch.empty((), dtype=torch.float32, device=x.device)
    _lib.check(_L().i3d_sum_scaled(_p(x), x.numel(), float(scale), _p(out), _s()), "i3d_sum_scaled")
    return out


def ntxent_rows_bwd(P, B, Bc, C, n1, n2, norm, eps, tau, row_offset, rowstats, gout, inv_B):
    dn1 = torch.empty(B, dtype=torch.float32, device=P.device) if norm else None
    dn2 = torch.zeros(Bc * C, dtype=torch.float32, device=P.device) if norm else None
    _lib.check(_L().i3d_ntxent_rows_bwd(_p(P), B, Bc, C, _p(n1), _p(n2), 1 if norm else 0, float(eps), float(tau),
                                        int(row_offset), _p(rowstats), _p(gout), float(inv_B), _p(dn1), _p(dn2),
                                        _s()), "i3d_ntxent_rows_bwd")
    return dn1, dn2


def norm_bwd_accum(z, norms, dn, dz):
    R, D = z.shape
    _lib.check(_L().i3d_norm_bwd_accum(_p(z), _p(norms), _p(dn), R, D, _p(dz), _s()), "i3d_norm_bwd_accum")
    return dz


# ------------------------------------------------------------------------------------- optimizer
def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, grad_scale, step, hyper_dev=None, step_dev=None):
    _lib.check(_L().i3d_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2),
                                  float(eps), float(weight_decay), float(grad_scale), int(step), _p(hyper_dev),
                                  _p(step_dev), _s()), "i3d_adam_step")


def adam_step_nvls(mc_ptr, local, n, rank, world, lr, beta1, beta2, eps, weight_decay, grad_scale, step, hyper_dev=None,
                   step_dev=None):
    """fused all-reduce + Adam + broadcast over multicast memory (i3d_adam_step_nvls); ``local``: the [4n] symmetric
    tensor [p | g | m | v], ``mc_ptr``: its multicast address"""
    _lib.check(_L().i3d_adam_step_nvls(int(mc_ptr), _p(local), int(n), int(rank), int(world), float(lr), float(beta1),
                                       float(beta2), float(eps), float(weight_decay), float(grad_scale), int(step),
                                       _p(hyper_dev), _p(step_dev), _s()), "i3d_adam_step_nvls")


def add_i64(x, delta):
    _lib.check(_L().i3d_add_i64(_p(x), int(delta), _s()), "i3d_add_i64")


def multi_copy(ptrs, off, length, flat, to_flat, stride=None):
    _lib.check(_L().i3d_multi_copy_strided(_p(ptrs), _p(off), _p(length), _p(stride), ptrs.numel(), _p(flat),
                                           1 if to_flat else 0, _s()), "i3d_multi_copy")
