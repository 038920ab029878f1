// PNA hot-path kernels: categorical embedding sums, the fused multi-aggregator segmented reduction
// (mean | max | min | std) with its backward, per-graph readouts, generic segment sums.
//
// All of them are HBM-bound streaming kernels: one thread owns one (row-segment, 16-byte column group)
// pair, a warp reads 512 contiguous bytes per message row, neighbour rows of a node are contiguous
// because messages are produced in CSR order (see DESIGN.md "data layout").
#include <initializer_list>

#include <cstdlib>

#include "i3d_vec.cuh"

namespace i3d {

// ------------------------------------------------------------------------------------------------
// AtomEncoder / BondEncoder (commons/mol_encoder.py:34-42,65-73)
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void embed_sum_fwd_kernel(const int64_t* __restrict__ idx, int64_t R, int C,
                                     const int32_t* __restrict__ col_off, const int32_t* __restrict__ perm,
                                     const float* __restrict__ table, int F, float* __restrict__ out) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = R * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / FV;
    const int c0 = (int)(t - r * FV) * V;
    const int64_t rr = perm ? (int64_t)perm[r] : r;
    Vec<V> acc;
    acc.fill(0.f);
    for (int c = 0; c < C; ++c) {
      const int64_t row = (int64_t)col_off[c] + idx[rr * C + c];
      Vec<V> e;
      e.load(table + row * F + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = __fadd_rn(acc.v[i], e.v[i]);
    }
    acc.store(out + r * F + c0);
  }
}

template <int V>
__global__ void embed_sum_bwd_kernel(const int64_t* __restrict__ idx, int64_t R, int C,
                                     const int32_t* __restrict__ col_off, const int32_t* __restrict__ perm,
                                     const float* __restrict__ gout, int F, float* __restrict__ gtable) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = R * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / FV;
    const int c0 = (int)(t - r * FV) * V;
    const int64_t rr = perm ? (int64_t)perm[r] : r;
    Vec<V> g;
    g.load(gout + r * F + c0);
    for (int c = 0; c < C; ++c) {
      const int64_t row = (int64_t)col_off[c] + idx[rr * C + c];
      float* dst = gtable + row * F + c0;
#pragma unroll
      for (int i = 0; i < V; ++i) atomicAdd(dst + i, g.v[i]);
    }
  }
}

// Same gradient with a per-CTA shared-memory copy of ONE column's table slice: a CTA takes a chunk of rows and one
// categorical column c, accumulates gout rows into acc[value][F] with shared-memory atomics (a column has <= 119
// values, half the atoms are hydrogens: global atomics on those rows serialise), then flushes only the touched rows.
template <int V>
__global__ void __launch_bounds__(256)
    embed_sum_bwd_smem_kernel(const int64_t* __restrict__ idx, int64_t R, int C, const int32_t* __restrict__ col_off,
                              const int32_t* __restrict__ perm, const float* __restrict__ gout, int F,
                              float* __restrict__ gtable, int table_rows, int rows_per_cta) {
  pdl_grid_sync();
  extern __shared__ float acc[];          // [dim][F]
  __shared__ unsigned touched[8];
  const int c = blockIdx.y;
  const int base = col_off[c];
  const int dim = min((c + 1 < C ? col_off[c + 1] : table_rows) - base, 256);
  for (int t = threadIdx.x; t < dim * F; t += blockDim.x) acc[t] = 0.f;
  if (threadIdx.x < 8) touched[threadIdx.x] = 0u;
  __syncthreads();
  const int FV = F / V;
  const int RP = blockDim.x / FV;
  const int cgp = threadIdx.x % FV, rg = threadIdx.x / FV;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(R, r0 + rows_per_cta);
  if (rg < RP) {
    // kEmbRows rows in flight per thread: the chain perm -> idx is two dependent loads per row (and the gradient row a
    // third, independent one); issued one row at a time the loop is pure memory latency
    constexpr int kEmbRows = 8;
    for (int64_t rb = r0 + rg; rb < r1; rb += (int64_t)RP * kEmbRows) {
      int64_t rr[kEmbRows];
      int v[kEmbRows];
      Vec<V> g[kEmbRows];
#pragma unroll
      for (int u = 0; u < kEmbRows; ++u) {
        const int64_t r = rb + (int64_t)u * RP;
        rr[u] = -1;
        if (r < r1) {
          rr[u] = perm ? (int64_t)__ldg(perm + r) : r;
          g[u].load(gout + r * F + cgp * V);
        }
      }
#pragma unroll
      for (int u = 0; u < kEmbRows; ++u) v[u] = rr[u] >= 0 ? (int)__ldg(idx + rr[u] * C + c) : -1;
#pragma unroll
      for (int u = 0; u < kEmbRows; ++u) {
        if (v[u] < 0 || v[u] >= dim) continue;
        float* dst = acc + v[u] * F + cgp * V;
#pragma unroll
        for (int i = 0; i < V; ++i) atomicAdd(dst + i, g[u].v[i]);
        if (cgp == 0) atomicOr(&touched[v[u] >> 5], 1u << (v[u] & 31));
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < dim * F; t += blockDim.x) {
    const int v = t / F;
    if (touched[v >> 5] & (1u << (v & 31))) atomicAdd(gtable + (int64_t)(base + v) * F + (t - v * F), acc[t]);
  }
}

// ------------------------------------------------------------------------------------------------
// PNA aggregation forward (models/pna.py:17-37,221-235).  One thread = one node x one column group.
// Arithmetic is written with explicit round-to-nearest intrinsics so that nvcc does not contract
// mean(x^2) - mean(x)^2 into an FMA: the relu gate of the variance must see the same sign as PyTorch.
// ------------------------------------------------------------------------------------------------
constexpr float kAggEps = 1e-5f;  // EPS, models/pna.py:14

// One thread owns one (node, 16-byte column group).  Bond graphs have in-degree <= 4 almost always, so the fast path
// issues every row load of the item before any arithmetic (predicated, no loop) and a thread works on kAggItems
// items at once: the kernel is a latency chain rowptr -> rows -> stores, and memory-level parallelism per thread is
// what moves it towards the HBM roofline.  Sums run in edge order j = 0..D-1 in both paths and in the backward
// kernel (the relu gate on var must see the same bits).

template <int V>
__device__ __forceinline__ void agg_finish(const Vec<V>& s, const Vec<V>& q, const Vec<V>& mx, const Vec<V>& mn,
                                           int D, float* o, int F) {
  const float fd = (float)D;
  Vec<V> mean, sd;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mean.v[i] = __fdiv_rn(s.v[i], fd);
    const float msq = __fdiv_rn(q.v[i], fd);
    const float var = fmaxf(__fsub_rn(msq, __fmul_rn(mean.v[i], mean.v[i])), 0.f);
    sd.v[i] = __fsqrt_rn(__fadd_rn(var, kAggEps));
  }
  mean.store_cs(o), mx.store_cs(o + F), mn.store_cs(o + 2 * F), sd.store_cs(o + 3 * F);
}

template <int V, int kAggItems, int MINB>
__global__ void __launch_bounds__(256, MINB)
    pna_aggregate_fwd_kernel(const float* __restrict__ msg, const int32_t* __restrict__ rowptr, int64_t N, int F,
                             float* __restrict__ out, int ldo) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = N * FV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t0 < total; t0 += stride * kAggItems) {
    int64_t v[kAggItems];
    int c0[kAggItems], D[kAggItems];
    const float* p[kAggItems];
#pragma unroll
    for (int u = 0; u < kAggItems; ++u) {
      const int64_t t = t0 + u * stride;
      D[u] = -1;                                        // -1: no item
      if (t < total) {
        v[u] = t / FV;
        c0[u] = (int)(t - v[u] * FV) * V;
        const int32_t b = __ldg(rowptr + v[u]), e = __ldg(rowptr + v[u] + 1);
        D[u] = e - b;
        p[u] = msg + (int64_t)b * F + c0[u];
      }
    }
    Vec<V> x[kAggItems][4];
#pragma unroll
    for (int u = 0; u < kAggItems; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < D[u]) x[u][j].load(p[u] + (int64_t)j * F);
#pragma unroll
    for (int u = 0; u < kAggItems; ++u) {
      if (D[u] < 0) continue;
      float* o = out + v[u] * (int64_t)ldo + c0[u];
      Vec<V> s, q, mx, mn;
      if (D[u] == 0) {
        s.fill(0.f);
        s.store_cs(o), s.store_cs(o + F), s.store_cs(o + 2 * F), s.store_cs(o + 3 * F);
        continue;
      }
      s.fill(0.f), q.fill(0.f), mx.fill(-INFINITY), mn.fill(INFINITY);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < D[u]) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            s.v[i] = __fadd_rn(s.v[i], x[u][j].v[i]);
            q.v[i] = __fadd_rn(q.v[i], __fmul_rn(x[u][j].v[i], x[u][j].v[i]));
            mx.v[i] = fmaxf(mx.v[i], x[u][j].v[i]);
            mn.v[i] = fminf(mn.v[i], x[u][j].v[i]);
          }
        }
      for (int k = 4; k < D[u]; ++k) {                  // rare: in-degree > 4
        Vec<V> y;
        y.load(p[u] + (int64_t)k * F);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          s.v[i] = __fadd_rn(s.v[i], y.v[i]);
          q.v[i] = __fadd_rn(q.v[i], __fmul_rn(y.v[i], y.v[i]));
          mx.v[i] = fmaxf(mx.v[i], y.v[i]);
          mn.v[i] = fminf(mn.v[i], y.v[i]);
        }
      }
      agg_finish<V>(s, q, mx, mn, D[u], o, F);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Aggregation forward, bulk-copy staged ("TMA" 1-D: cp.async.bulk global -> shared, mbarrier complete_tx).
// Messages are in CSR order, so the mailboxes of a block of consecutive nodes are ONE contiguous byte range of `msg`:
// a single elected thread moves it with one bulk copy per block, two blocks in flight per CTA (double buffer), and
// the 256 threads reduce out of shared memory.  Memory-level parallelism no longer costs registers (the LDG variant
// holds 2 items x 4 rows x 16 B per thread): ~50 KB per CTA are in flight from the first microsecond, which is what
// a ~10 us kernel needs to approach the HBM roofline.  Persistent grid (<= 2 CTAs per SM), block = kAggNodes nodes;
// a block whose mailbox does not fit a stage (very high degrees) is reduced straight from global memory.
// Sums run in edge order j = 0..D-1 exactly like the LDG kernel and the backward (bit-identical results).
// ------------------------------------------------------------------------------------------------
constexpr int kAggNodes = 16;      // nodes per block: 16 x in-degree <= 4 x 800 B = at most 51 KB per stage at F = 200

__device__ __forceinline__ uint32_t agg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mean | max | min | std of the D rows starting at p (row stride F floats) for one 16-byte column group
__device__ __forceinline__ void agg_reduce_rows(const float* p, int D, int F, float* o) {
  Vec<4> s, q, mx, mn;
  s.fill(0.f), q.fill(0.f), mx.fill(-INFINITY), mn.fill(INFINITY);
  for (int k = 0; k < D; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(p + (int64_t)k * F);
    const float x[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s.v[j] = __fadd_rn(s.v[j], x[j]);
      q.v[j] = __fadd_rn(q.v[j], __fmul_rn(x[j], x[j]));
      mx.v[j] = fmaxf(mx.v[j], x[j]);
      mn.v[j] = fminf(mn.v[j], x[j]);
    }
  }
  agg_finish<4>(s, q, mx, mn, D, o, F);
}

__global__ void __launch_bounds__(256, 2)
    pna_aggregate_fwd_tma_kernel(const float* __restrict__ msg, const int32_t* __restrict__ rowptr, int64_t N, int F,
                                 float* __restrict__ out, int ldo, int stage_rows) {
  pdl_grid_sync();
  extern __shared__ __align__(128) uint8_t agg_smem[];
  float* stage0 = reinterpret_cast<float*>(agg_smem);
  const size_t stage_floats = (size_t)stage_rows * F;
  uint64_t* full = reinterpret_cast<uint64_t*>(agg_smem + 2 * stage_floats * sizeof(float));   // [2]
  int32_t* s_rp = reinterpret_cast<int32_t*>(full + 2);                                        // [2][kAggNodes + 1]
  int32_t* s_direct = s_rp + 2 * (kAggNodes + 1);                                              // [2]
  const int tid = threadIdx.x;
  const int FV = F >> 2;
  const int64_t n_blocks = (N + kAggNodes - 1) / kAggNodes;
  const int64_t G = gridDim.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(agg_smem_u32(&full[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(agg_smem_u32(&full[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // thread 0: request mailbox rows [b, e) into stage st (one bulk copy, completion counted in bytes on full[st])
  auto issue = [&](int32_t b, int32_t e, int st) {
    const int rows = e - b;
    const uint32_t bar = agg_smem_u32(&full[st]);
    if (rows > 0 && rows <= stage_rows) {
      const uint32_t bytes = (uint32_t)rows * (uint32_t)F * 4u;
      s_direct[st] = 0;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       agg_smem_u32(stage0 + (size_t)st * stage_floats)),
                   "l"(msg + (int64_t)b * F), "r"(bytes), "r"(bar)
                   : "memory");
    } else {
      s_direct[st] = 1;                 // empty block, or too many rows for a stage: reduce from global memory
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
  };
  auto block_rows = [&](int64_t blk, int32_t* b, int32_t* e) {
    const int64_t v0 = blk * kAggNodes;
    const int64_t v1 = v0 + kAggNodes < N ? v0 + kAggNodes : N;
    *b = __ldg(rowptr + v0), *e = __ldg(rowptr + v1);
  };

  if (tid == 0) {                       // both stages are requested before any reduction starts
    int32_t b0, e0, b1, e1;
    const int64_t blk0 = blockIdx.x, blk1 = blk0 + G;
    if (blk0 < n_blocks) block_rows(blk0, &b0, &e0);
    if (blk1 < n_blocks) block_rows(blk1, &b1, &e1);
    if (blk0 < n_blocks) issue(b0, e0, 0);
    if (blk1 < n_blocks) issue(b1, e1, 1);
  }
  int it = 0;
  for (int64_t blk = blockIdx.x; blk < n_blocks; blk += G, ++it) {
    const int st = it & 1;
    const int64_t v0 = blk * kAggNodes;
    const int nodes = (int)(v0 + kAggNodes < N ? kAggNodes : N - v0);
    // rows of the block that will reuse this stage: loaded now, consumed after the reduction (no stall on the way)
    const int64_t nxt = blk + 2 * G;
    int32_t nb = 0, ne = 0;
    if (tid == 0 && nxt < n_blocks) block_rows(nxt, &nb, &ne);
    if (tid <= nodes) s_rp[st * (kAggNodes + 1) + tid] = __ldg(rowptr + v0 + tid);
    {   // wait for the bulk copy of this stage (the phase parity flips on every reuse of the stage)
      const uint32_t bar = agg_smem_u32(&full[st]);
      const uint32_t parity = (uint32_t)((it >> 1) & 1);
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "AGG_WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra AGG_DONE_%=;\n"
          "bra AGG_WAIT_%=;\n"
          "AGG_DONE_%=:\n"
          "}\n" ::"r"(bar),
          "r"(parity)
          : "memory");
    }
    __syncthreads();                    // s_rp / s_direct of this stage visible to every thread
    const int32_t* rp = s_rp + st * (kAggNodes + 1);
    const int32_t row0 = rp[0];
    const bool direct = s_direct[st] != 0;
    const float* staged = stage0 + (size_t)st * stage_floats;
    for (int idx = tid; idx < nodes * FV; idx += 256) {
      const int i = idx / FV;
      const int c0 = (idx - i * FV) << 2;
      const int32_t b = rp[i];
      const int D = rp[i + 1] - b;
      float* o = out + (v0 + i) * (int64_t)ldo + c0;
      if (D <= 0) {
        Vec<4> z;
        z.fill(0.f);
        z.store_cs(o), z.store_cs(o + F), z.store_cs(o + 2 * F), z.store_cs(o + 3 * F);
      } else if (!direct) {
        agg_reduce_rows(staged + (size_t)(b - row0) * F + c0, D, F, o);
      } else {
        agg_reduce_rows(msg + (int64_t)b * F + c0, D, F, o);
      }
    }
    __syncthreads();                    // every thread is done with this stage before it is refilled
    if (tid == 0 && nxt < n_blocks) issue(nb, ne, st);
  }
}

// backward: g = [g_mean | g_max | g_min | g_std] already folded over the degree scalers by the GEMM backward
template <int V>
__global__ void __launch_bounds__(256)
    pna_aggregate_bwd_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ msg,
                             const float* __restrict__ out, int ldo, const int32_t* __restrict__ rowptr, int64_t N,
                             int F, float* __restrict__ dmsg) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = N * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / FV;
    const int c0 = (int)(t - v * FV) * V;
    const int32_t b = __ldg(rowptr + v), e = __ldg(rowptr + v + 1);
    const int D = e - b;
    if (D <= 0) continue;
    const float* gp = g + v * (int64_t)ldg + c0;
    const float* op = out + v * (int64_t)ldo + c0;
    Vec<V> gmean, gmax, gmin, gstd, mean, mx, mn, sd;
    gmean.load(gp), gmax.load(gp + F), gmin.load(gp + 2 * F), gstd.load(gp + 3 * F);
    mean.load(op), mx.load(op + F), mn.load(op + 2 * F), sd.load(op + 3 * F);
    const float* p = msg + (int64_t)b * F + c0;
    float* dp = dmsg + (int64_t)b * F + c0;
    // pass 1: sum of squares (same order as forward -> same relu gate), first arg-max / arg-min
    Vec<V> q;
    q.fill(0.f);
    int imax[V], imin[V];
#pragma unroll
    for (int i = 0; i < V; ++i) imax[i] = -1, imin[i] = -1;
    for (int k = 0; k < D; ++k) {
      Vec<V> x;
      x.load(p + (int64_t)k * F);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        q.v[i] = __fadd_rn(q.v[i], __fmul_rn(x.v[i], x.v[i]));
        if (imax[i] < 0 && x.v[i] == mx.v[i]) imax[i] = k;
        if (imin[i] < 0 && x.v[i] == mn.v[i]) imin[i] = k;
      }
    }
    const float fd = (float)D;
    Vec<V> base, cs;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float msq = __fdiv_rn(q.v[i], fd);
      const float var = __fsub_rn(msq, __fmul_rn(mean.v[i], mean.v[i]));
      base.v[i] = gmean.v[i] / fd;
      cs.v[i] = var > 0.f ? gstd.v[i] / (fd * sd.v[i]) : 0.f;
    }
    // pass 2 (rows are L1/L2 resident from pass 1)
    for (int k = 0; k < D; ++k) {
      Vec<V> x, d;
      x.load(p + (int64_t)k * F);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float r = base.v[i] + cs.v[i] * (x.v[i] - mean.v[i]);
        if (k == imax[i]) r += gmax.v[i];
        if (k == imin[i]) r += gmin.v[i];
        d.v[i] = r;
      }
      d.store_cs(dp + (int64_t)k * F);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Degree-merged posttrans weights (models/pna.py:57-68,207-211,232).  For a node of in-degree D the posttrans input
// is cat[h, A, A*a_D, A*t_D] with a_D = (float)ln(D+1), t_D = (float)(1/ln(D+1)) (A = [mean|max|min|std], 4F wide), so
//   cat[...] W^T = h Wh^T + A (W_id + a_D W_amp + t_D W_att)^T
// One merged weight per degree bucket turns the K = 13F GEMM into K = 5F.  merge writes the tf32 hi/lo operands the
// bucketed NT kernel streams by TMA, for the forward ([NB*Fout, kpad(F)+kpad(4F)], K-major) and for dx = dy Wm
// ([NB*5F, kpad(Fout)], i.e. Wm^T); unmerge folds the per-bucket weight gradients back onto W's three column blocks.
// Pad columns of the scratch are never written (the caller zero-fills the buffers once).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bucket_scalers(int b, float* a, float* t) {
  if (b <= 0) {
    *a = 0.f, *t = 0.f;                 // D = 0: the aggregation row is all zeros (DGL zero fill), any finite factor works
  } else {
    const double l = log((double)b + 1.0);
    *a = (float)l, *t = (float)(1.0 / l);
  }
}

__device__ __forceinline__ int kpad32(int k) { return (k + 31) / 32 * 32; }
__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// one CTA = one 32 x 32 tile of (output row n, merged column j) of one bucket: both operand layouts are written with
// coalesced rows (the backward operand is the transpose, staged through shared memory)
__global__ void __launch_bounds__(256)
    posttrans_merge_kernel(const float* __restrict__ W, int ldw, int Fout, int F, int NB, float* __restrict__ fwd_hi,
                           float* __restrict__ fwd_lo, int ktf, float* __restrict__ bwd_hi, float* __restrict__ bwd_lo,
                           int ktb) {
  pdl_grid_sync();
  __shared__ float th[32][33], tl[32][33];
  const int F4 = 4 * F, K5 = 5 * F;
  const int tiles_j = (K5 + 31) / 32, tiles_n = (Fout + 31) / 32;
  const int t = blockIdx.x;
  const int b = t / (tiles_j * tiles_n);
  const int r = t - b * (tiles_j * tiles_n);
  const int n0 = (r / tiles_j) * 32, j0 = (r % tiles_j) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float a, tt;
  bucket_scalers(b, &a, &tt);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + 8 * i, j = j0 + tx;
    float h = 0.f, l = 0.f;
    if (n < Fout && j < K5) {
      const float* w = W + (int64_t)n * ldw;
      float m;
      int col;
      if (j < F) {
        m = __ldg(w + j);
        col = j;
      } else {
        const int c = j - F;
        m = __ldg(w + F + c) + a * __ldg(w + F + F4 + c) + tt * __ldg(w + F + 2 * F4 + c);
        col = kpad32(F) + c;
      }
      h = to_tf32_rna(m);
      l = m - h;
      const int64_t of = ((int64_t)b * Fout + n) * ktf + col;
      fwd_hi[of] = h, fwd_lo[of] = l;
    }
    th[ty + 8 * i][tx] = h, tl[ty + 8 * i][tx] = l;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty + 8 * i, n = n0 + tx;
    if (n < Fout && j < K5) {
      const int64_t ob = ((int64_t)b * K5 + j) * ktb + n;
      bwd_hi[ob] = th[tx][ty + 8 * i], bwd_lo[ob] = tl[tx][ty + 8 * i];
    }
  }
}

// dW[n, F + c] += sum_b dWb[b,n,c];  dW[n, 5F + c] += sum_b a_b dWb[b,n,c];  dW[n, 9F + c] += sum_b t_b dWb[b,n,c]
__global__ void __launch_bounds__(256)
    posttrans_unmerge_kernel(const float* __restrict__ dWb, int NB, int Fout, int F, float* __restrict__ dW, int ldw) {
  pdl_grid_sync();
  const int F4 = 4 * F;
  const int64_t total = (int64_t)Fout * F4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % F4);
    const int n = (int)(t / F4);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int b = 0; b < NB; ++b) {
      float a, tt;
      bucket_scalers(b, &a, &tt);
      const float g = __ldg(dWb + ((int64_t)b * Fout + n) * F4 + c);
      s0 += g, s1 += a * g, s2 += tt * g;
    }
    float* d = dW + (int64_t)n * ldw + F + c;
    d[0] += s0, d[F4] += s1, d[2 * F4] += s2;
  }
}

// ------------------------------------------------------------------------------------------------
// dgl.readout_nodes(graph, 'feat', op)  (models/pna.py:133, models/net3d.py:73)
// ------------------------------------------------------------------------------------------------
struct RoOps {
  int n;
  int op[4];
};

template <int V>
__global__ void segment_readout_fwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ ptr,
                                           int64_t B, int F, RoOps ops, float* __restrict__ out) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = B * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = t / FV;
    const int c0 = (int)(t - g * FV) * V;
    const int32_t b = ptr[g], e = ptr[g + 1];
    Vec<V> s, mx, mn;
    s.fill(0.f), mx.fill(-INFINITY), mn.fill(INFINITY);
    for (int32_t r = b; r < e; ++r) {
      Vec<V> v;
      v.load(x + (int64_t)r * ldx + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s.v[i] = __fadd_rn(s.v[i], v.v[i]);
        mx.v[i] = fmaxf(mx.v[i], v.v[i]);
        mn.v[i] = fminf(mn.v[i], v.v[i]);
      }
    }
    const float cnt = (float)(e - b);
    for (int j = 0; j < ops.n; ++j) {
      Vec<V> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float r;
        if (e <= b) r = 0.f;
        else if (ops.op[j] == I3D_RO_SUM) r = s.v[i];
        else if (ops.op[j] == I3D_RO_MEAN) r = __fdiv_rn(s.v[i], cnt);
        else if (ops.op[j] == I3D_RO_MAX) r = mx.v[i];
        else r = mn.v[i];
        o.v[i] = r;
      }
      o.store(out + g * (int64_t)(ops.n * F) + (int64_t)j * F + c0);
    }
  }
}

template <int V>
__global__ void segment_readout_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, int ldx,
                                           const float* __restrict__ out, const int32_t* __restrict__ ptr, int64_t B,
                                           int F, RoOps ops, float* __restrict__ dx, int lddx, int64_t n_rows) {
  pdl_grid_sync();
  const int FV = F / V;
  // rows behind the last graph (padding nodes of a shape-bucketed batch) belong to no readout: zero gradient
  {
    const int64_t r0 = ptr[B];
    const int64_t ztotal = n_rows > r0 ? (n_rows - r0) * FV : 0;
    Vec<V> z;
    z.fill(0.f);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ztotal; t += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = r0 + t / FV;
      z.store(dx + r * lddx + (int)(t % FV) * V);
    }
  }
  const int64_t total = B * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gi = t / FV;
    const int c0 = (int)(t - gi * FV) * V;
    const int32_t b = ptr[gi], e = ptr[gi + 1];
    if (e <= b) continue;
    const float cnt = (float)(e - b);
    Vec<V> base, gmx, gmn, vmx, vmn;
    base.fill(0.f), gmx.fill(0.f), gmn.fill(0.f), vmx.fill(0.f), vmn.fill(0.f);
    bool has_max = false, has_min = false;
    for (int j = 0; j < ops.n; ++j) {
      Vec<V> gj, oj;
      const int64_t off = gi * (int64_t)(ops.n * F) + (int64_t)j * F + c0;
      gj.load(g + off);
      oj.load(out + off);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (ops.op[j] == I3D_RO_SUM) base.v[i] += gj.v[i];
        else if (ops.op[j] == I3D_RO_MEAN) base.v[i] += gj.v[i] / cnt;
        else if (ops.op[j] == I3D_RO_MAX) gmx.v[i] += gj.v[i], vmx.v[i] = oj.v[i];
        else gmn.v[i] += gj.v[i], vmn.v[i] = oj.v[i];
      }
      has_max |= ops.op[j] == I3D_RO_MAX;
      has_min |= ops.op[j] == I3D_RO_MIN;
    }
    bool done_max[V], done_min[V];
#pragma unroll
    for (int i = 0; i < V; ++i) done_max[i] = !has_max, done_min[i] = !has_min;
    for (int32_t r = b; r < e; ++r) {
      Vec<V> v, d;
      v.load(x + (int64_t)r * ldx + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float acc = base.v[i];
        if (!done_max[i] && v.v[i] == vmx.v[i]) acc += gmx.v[i], done_max[i] = true;
        if (!done_min[i] && v.v[i] == vmn.v[i]) acc += gmn.v[i], done_min[i] = true;
        d.v[i] = acc;
      }
      d.store(dx + (int64_t)r * lddx + c0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic CSR segment sum / mean with optional row indirection and addend
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void segment_sum_fwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ rowptr,
                                       const int32_t* __restrict__ idx, int64_t N, int F, int mean,
                                       const float* __restrict__ addend, int lda, float* __restrict__ out, int ldo) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = N * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / FV;
    const int c0 = (int)(t - v * FV) * V;
    const int32_t b = __ldg(rowptr + v), e = __ldg(rowptr + v + 1);
    Vec<V> s;
    s.fill(0.f);
    for (int32_t k = b; k < e; ++k) {
      const int64_t row = idx ? (int64_t)__ldg(idx + k) : (int64_t)k;
      Vec<V> xv;
      xv.load(x + row * ldx + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) s.v[i] = __fadd_rn(s.v[i], xv.v[i]);
    }
    if (mean) {
      const float d = (float)max(e - b, 1);
#pragma unroll
      for (int i = 0; i < V; ++i) s.v[i] = __fdiv_rn(s.v[i], d);
    }
    if (addend) {
      Vec<V> a;
      a.load(addend + v * (int64_t)lda + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) s.v[i] = __fadd_rn(s.v[i], a.v[i]);
    }
    s.store(out + v * (int64_t)ldo + c0);
  }
}

template <int V>
__global__ void segment_sum_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ rowptr,
                                       const int32_t* __restrict__ rowid, int64_t E, int F, int mean,
                                       float* __restrict__ gx) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = E * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / FV;
    const int c0 = (int)(t - k * FV) * V;
    const int32_t v = __ldg(rowid + k);
    Vec<V> gv;
    if (v < 0) {                       // padding edge of a shape-bucketed batch: in no row
      gv.fill(0.f);
      gv.store(gx + k * F + c0);
      continue;
    }
    gv.load(g + (int64_t)v * F + c0);
    if (mean) {
      const float d = (float)max(__ldg(rowptr + v + 1) - __ldg(rowptr + v), 1);
#pragma unroll
      for (int i = 0; i < V; ++i) gv.v[i] = __fdiv_rn(gv.v[i], d);
    }
    gv.store(gx + k * F + c0);
  }
}

}  // namespace i3d

using namespace i3d;

// grid = enough 256-thread CTAs for work/items, capped at what is co-resident (occupancy x SMs): one full wave
template <typename Kern>
static int occ_grid(Kern kern, int64_t work, int items) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  int64_t need = (work + (int64_t)items * 256 - 1) / ((int64_t)items * 256);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

#define I3D_DISPATCH_VEC(vec_ok, KERNEL, grid, block, stream, ...)          \
  do {                                                                       \
    if (vec_ok)                                                              \
      launch(KERNEL<4>, grid, block, 0, stream, __VA_ARGS__);                    \
    else                                                                     \
      launch(KERNEL<1>, grid, block, 0, stream, __VA_ARGS__);                    \
  } while (0)

extern "C" {

int i3d_embed_sum_fwd(const int64_t* idx, int64_t R, int C, const int32_t* col_off, const int32_t* perm,
                      const float* table, int F, float* out, void* stream) {
  I3D_REQUIRE(R >= 0 && C > 0 && F > 0 && col_off && table && (R == 0 || (idx && out)), "invalid argument");
  if (R == 0) return I3D_OK;
  const bool v4 = can_vec4({table, out}, {F});
  const int64_t work = R * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, embed_sum_fwd_kernel, grid_for(work, 256), 256, as_stream(stream), idx, R, C, col_off, perm,
                   table, F, out);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_embed_sum_bwd(const int64_t* idx, int64_t R, int C, const int32_t* col_off, const int32_t* perm,
                      const float* gout, int F, float* gtable, int table_rows, int max_dim, void* stream) {
  I3D_REQUIRE(R >= 0 && C > 0 && F > 0 && col_off && gtable && (R == 0 || (idx && gout)), "invalid argument");
  if (R == 0) return I3D_OK;
  const bool v4 = can_vec4({gout}, {F});
  const size_t smem = (size_t)max_dim * F * sizeof(float);
  if (table_rows > 0 && max_dim > 0 && max_dim <= 256 && smem <= 200 * 1024 && F / (v4 ? 4 : 1) <= 256) {
    static bool configured = false;
    if (!configured) {
      I3D_CUDA(cudaFuncSetAttribute(embed_sum_bwd_smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      I3D_CUDA(cudaFuncSetAttribute(embed_sum_bwd_smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    // (measured for the one-column case — the bond-code table of the factored edge layer — at batch 512: 37 / 74 / 296
    //  CTAs give 4.61 / 4.39 / 4.24 ms per training step: the kernel wants parallelism, the final atomics are cheap)
    int64_t chunks = (2 * (int64_t)sm_count() + C - 1) / C;
    {
      static int one_col = -1;
      if (one_col < 0) {
        const char* e = getenv("I3D_EMB_BWD_CHUNKS1");
        one_col = e ? atoi(e) : 0;
      }
      if (C == 1 && one_col > 0) chunks = one_col;
    }
    int rows_per_cta = (int)((R + chunks - 1) / chunks);
    if (rows_per_cta < 64) rows_per_cta = 64;
    chunks = (R + rows_per_cta - 1) / rows_per_cta;
    const dim3 grid((unsigned)chunks, C);
    if (v4)
      launch(embed_sum_bwd_smem_kernel<4>, grid, 256, smem, as_stream(stream), idx, R, C, col_off, perm, gout, F, gtable,
                                                                            table_rows, rows_per_cta);
    else
      launch(embed_sum_bwd_smem_kernel<1>, grid, 256, smem, as_stream(stream), idx, R, C, col_off, perm, gout, F, gtable,
                                                                            table_rows, rows_per_cta);
    I3D_LAUNCHED();
    return I3D_OK;
  }
  const int64_t work = R * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, embed_sum_bwd_kernel, grid_for(work, 256), 256, as_stream(stream), idx, R, C, col_off, perm,
                   gout, F, gtable);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_pna_aggregate_fwd(const float* msg, const int32_t* rowptr, int64_t N, int F, float* out, int ldo,
                          void* stream) {
  I3D_REQUIRE(N >= 0 && F > 0 && ldo >= 4 * F && rowptr && (N == 0 || out), "invalid argument");
  if (N == 0) return I3D_OK;
  const bool v4 = can_vec4({msg, out}, {F, ldo});
  const int64_t work = N * (F / (v4 ? 4 : 1));
  cudaStream_t st = as_stream(stream);
  // default: register-staged LDG kernel.  I3D_AGG_FWD=tma selects the bulk-copy (TMA 1-D) staged variant, measured
  // SLOWER on B200 at the target sizes (batch 512: 9.8 vs 7.9 us warm, 11.3 vs 9.4 us cold; batch 2048: 37.2 vs 36.2 us,
  // tests/gpu_agg_bench.py): at <= 2 blocks per CTA its copy -> reduce -> store phases barely overlap.
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("I3D_AGG_FWD");
    variant = (e && e[0] == 't') ? 1 : 0;
  }
  if (variant == 1 && v4 && N >= kAggNodes) {
    // two stages per CTA, two CTAs per SM: a stage holds a block of kAggNodes nodes of in-degree <= 4 (64 rows),
    // fewer if F is wide; blocks with more rows take the in-kernel global-memory path
    int stage_rows = 4 * kAggNodes;
    while (stage_rows > 8 && (size_t)2 * stage_rows * F * 4 > 100 * 1024) stage_rows -= 8;
    const size_t smem = (size_t)2 * stage_rows * F * 4 + 2 * sizeof(uint64_t) + (2 * (kAggNodes + 1) + 2) * sizeof(int32_t);
    static size_t configured = 0;
    if (smem > configured) {
      I3D_CUDA(cudaFuncSetAttribute(pna_aggregate_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    const int64_t n_blocks = (N + kAggNodes - 1) / kAggNodes;
    const int64_t cap = (int64_t)sm_count() * 2;
    launch(pna_aggregate_fwd_tma_kernel, (int)(n_blocks < cap ? n_blocks : cap), 256, smem, st, msg, rowptr, N, F, out,
           ldo, stage_rows);
    I3D_LAUNCHED();
    return I3D_OK;
  }
  // 2 items per thread at >= 3 CTAs/SM: the best of the LDG variants measured on B200 (tests/gpu_agg_bench.py)
  if (v4)
    launch(pna_aggregate_fwd_kernel<4, 2, 3>, occ_grid(pna_aggregate_fwd_kernel<4, 2, 3>, work, 2), 256, 0, st, msg,
           rowptr, N, F, out, ldo);
  else
    launch(pna_aggregate_fwd_kernel<1, 2, 3>, occ_grid(pna_aggregate_fwd_kernel<1, 2, 3>, work, 2), 256, 0, st, msg,
           rowptr, N, F, out, ldo);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_pna_aggregate_bwd(const float* g, int ldg, const float* msg, const float* out, int ldo,
                          const int32_t* rowptr, int64_t N, int F, float* dmsg, void* stream) {
  I3D_REQUIRE(N >= 0 && F > 0 && ldo >= 4 * F && ldg >= 4 * F && rowptr && (N == 0 || (g && out)),
              "invalid argument");
  if (N == 0) return I3D_OK;
  const bool v4 = can_vec4({g, msg, out, dmsg}, {F, ldo, ldg});
  const int64_t work = N * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, pna_aggregate_bwd_kernel, grid_for(work, 256, 8), 256, as_stream(stream), g, ldg, msg, out, ldo,
                   rowptr, N, F, dmsg);
  I3D_LAUNCHED();
  return I3D_OK;
}

static int make_ops(int n_ops, const int32_t* ops, RoOps* r) {
  if (n_ops < 1 || n_ops > 4 || !ops) return -1;
  r->n = n_ops;
  for (int i = 0; i < 4; ++i) r->op[i] = i < n_ops ? ops[i] : 0;
  for (int i = 0; i < n_ops; ++i)
    if (ops[i] < 0 || ops[i] > 3) return -1;
  return 0;
}

int i3d_segment_readout_fwd(const float* x, int ldx, const int32_t* ptr, int64_t B, int F, int n_ops,
                            const int32_t* ops, float* out, void* stream) {
  RoOps r;
  I3D_REQUIRE(B >= 0 && F > 0 && ldx >= F && ptr && make_ops(n_ops, ops, &r) == 0, "invalid argument");
  if (B == 0) return I3D_OK;
  const bool v4 = can_vec4({x, out}, {F, ldx});
  const int64_t work = B * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, segment_readout_fwd_kernel, grid_for(work, 128), 128, as_stream(stream), x, ldx, ptr, B, F, r,
                   out);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_segment_readout_bwd(const float* g, const float* x, int ldx, const float* out, const int32_t* ptr,
                            int64_t B, int F, int n_ops, const int32_t* ops, float* dx, int lddx, void* stream) {
  return i3d_segment_readout_bwd_v(g, x, ldx, out, ptr, B, F, n_ops, ops, dx, lddx, 0, stream);
}

int i3d_segment_readout_bwd_v(const float* g, const float* x, int ldx, const float* out, const int32_t* ptr,
                              int64_t B, int F, int n_ops, const int32_t* ops, float* dx, int lddx, int64_t n_rows,
                              void* stream) {
  RoOps r;
  I3D_REQUIRE(B >= 0 && F > 0 && ldx >= F && lddx >= F && ptr && make_ops(n_ops, ops, &r) == 0, "invalid argument");
  if (B == 0) return I3D_OK;
  const bool v4 = can_vec4({g, x, out, dx}, {F, ldx, lddx});
  const int64_t work = B * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, segment_readout_bwd_kernel, grid_for(work, 128), 128, as_stream(stream), g, x, ldx, out, ptr, B,
                   F, r, dx, lddx, n_rows);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_segment_sum_fwd(const float* x, int ldx, const int32_t* rowptr, const int32_t* idx, int64_t N, int F,
                        int mean, const float* addend, int lda, float* out, int ldo, void* stream) {
  I3D_REQUIRE(N >= 0 && F > 0 && ldx >= F && ldo >= F && rowptr && (N == 0 || out), "invalid argument");
  if (N == 0) return I3D_OK;
  const bool v4 = can_vec4({x, addend, out}, {F, ldx, ldo, addend ? lda : 0});
  const int64_t work = N * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, segment_sum_fwd_kernel, grid_for(work, 256), 256, as_stream(stream), x, ldx, rowptr, idx, N, F,
                   mean, addend, lda, out, ldo);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_segment_sum_bwd(const float* g, const int32_t* rowptr, const int32_t* rowid, int64_t E, int F, int mean,
                        float* gx, void* stream) {
  I3D_REQUIRE(E >= 0 && F > 0 && rowptr && (E == 0 || (g && rowid && gx)), "invalid argument");
  if (E == 0) return I3D_OK;
  const bool v4 = can_vec4({g, gx}, {F});
  const int64_t work = E * (F / (v4 ? 4 : 1));
  I3D_DISPATCH_VEC(v4, segment_sum_bwd_kernel, grid_for(work, 256), 256, as_stream(stream), g, rowptr, rowid, E, F,
                   mean, gx);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_posttrans_merge(const float* W, int ldw, int Fout, int F, int n_buckets, float* fwd_hi, float* fwd_lo,
                        int fwd_pitch, float* bwd_hi, float* bwd_lo, int bwd_pitch, void* stream) {
  I3D_REQUIRE(W && Fout > 0 && F > 0 && ldw >= 13 * F && n_buckets >= 1 && n_buckets <= 16 && fwd_hi && fwd_lo &&
                  bwd_hi && bwd_lo, "invalid argument");
  I3D_REQUIRE(fwd_pitch >= (F + 31) / 32 * 32 + (4 * F + 31) / 32 * 32 && bwd_pitch >= (Fout + 31) / 32 * 32,
              "operand pitch smaller than the padded K extent");
  const int tiles = n_buckets * ((Fout + 31) / 32) * ((5 * F + 31) / 32);
  launch(posttrans_merge_kernel, tiles, 256, 0, as_stream(stream), W, ldw, Fout, F, n_buckets, fwd_hi, fwd_lo,
         fwd_pitch, bwd_hi, bwd_lo, bwd_pitch);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_posttrans_unmerge(const float* dWb, int n_buckets, int Fout, int F, float* dW, int ldw, void* stream) {
  I3D_REQUIRE(dWb && dW && Fout > 0 && F > 0 && ldw >= 13 * F && n_buckets >= 1 && n_buckets <= 16,
              "invalid argument");
  launch(posttrans_unmerge_kernel, grid_for((int64_t)Fout * 4 * F, 256), 256, 0, as_stream(stream), dWb, n_buckets,
         Fout, F, dW, ldw);
  I3D_LAUNCHED();
  return I3D_OK;
}
}
