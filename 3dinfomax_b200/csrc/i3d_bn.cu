// FCLayer tail kernels: activation + train-mode BatchNorm1d statistics / apply / backward
// (models/base_layers.py:100-111: Linear -> activation -> BatchNorm, statistics over ALL rows of the batch).
//
// Column statistics are accumulated in fp64: Net3D's BN inputs have mean^2 >> var,
// where an fp32 E[x^2]-E[x]^2 loses the variance.  All kernels are HBM-bound row streams: a thread owns one
// 16-byte column group and walks rows with a fixed stride, so column partials stay in registers.
#include <cstdlib>
#include <initializer_list>

#include "i3d_vec.cuh"

namespace i3d {

constexpr int kColThreads = 256;

// thread -> (column group cg, row phase rg); rows visited: blockIdx*RP + rg, += gridDim*RP
struct ColMap {
  int cg, rg, RP;
  bool active;
};
__device__ __forceinline__ ColMap col_map(int FV) {
  ColMap m;
  m.RP = kColThreads / FV;
  m.cg = threadIdx.x % FV;
  m.rg = threadIdx.x / FV;
  m.active = m.rg < m.RP;
  return m;
}

// block-reduce NV*V doubles per column group over the row phases, then atomically add to global
template <int V, int NV>
__device__ __forceinline__ void block_col_reduce_f64(double (&acc)[NV][V], const ColMap& m, int FV, int F,
                                                     double* __restrict__ gsum) {
  extern __shared__ double sh[];  // [NV][V][kColThreads]
  for (int a = 0; a < NV; ++a)
#pragma unroll
    for (int i = 0; i < V; ++i) sh[(a * V + i) * kColThreads + threadIdx.x] = m.active ? acc[a][i] : 0.0;
  __syncthreads();
  if (m.active && m.rg == 0) {
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        double s = 0.0;
        for (int r = 0; r < m.RP; ++r) s += sh[(a * V + i) * kColThreads + r * FV + m.cg];
        atomicAdd(gsum + ((int64_t)a * F + m.cg * V + i) * I3D_STATS_STRIDE, s);
      }
  }
}

// Same-address atomics serialise in L2 (~33 ns each, measured: +148 CTAs = +4.9 us on the gather-add kernel), so with
// ~300-450 CTAs the atomic tail of a column-statistics kernel was 10-15 us of a 16-30 us launch.  Two-stage variant:
// every CTA stores its column partials in its OWN slot of `slots` (plain coalesced stores), takes a ticket, and the
// last CTA to arrive sums the slots in slot order and stores the result (no pre-zeroed output needed, and the sums no
// longer depend on the order in which CTAs retire).  `counter` must be 0 on entry and is reset to 0 on exit.
template <typename T>
__device__ __forceinline__ void last_block_col_sum(const T* __restrict__ slots, int ncols, T* __restrict__ out,
                                                   unsigned* __restrict__ counter, int out_stride = 1) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int G = gridDim.x;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int b = 0;
    for (; b + 4 <= G; b += 4) {
      s0 += __ldcg(slots + (int64_t)b * ncols + c);
      s1 += __ldcg(slots + (int64_t)(b + 1) * ncols + c);
      s2 += __ldcg(slots + (int64_t)(b + 2) * ncols + c);
      s3 += __ldcg(slots + (int64_t)(b + 3) * ncols + c);
    }
    for (; b < G; ++b) s0 += __ldcg(slots + (int64_t)b * ncols + c);
    out[(int64_t)c * out_stride] = (s0 + s1) + (s2 + s3);
  }
  if (threadIdx.x == 0) *counter = 0u;
}

template <int V, int NV>
__device__ __forceinline__ void block_col_reduce_f64_ws(double (&acc)[NV][V], const ColMap& m, int FV, int F,
                                                        double* __restrict__ gsum, double* __restrict__ slots,
                                                        unsigned* __restrict__ counter) {
  if (!slots) {
    block_col_reduce_f64<V, NV>(acc, m, FV, F, gsum);
    return;
  }
  extern __shared__ double sh[];  // [NV][V][kColThreads]
  for (int a = 0; a < NV; ++a)
#pragma unroll
    for (int i = 0; i < V; ++i) sh[(a * V + i) * kColThreads + threadIdx.x] = m.active ? acc[a][i] : 0.0;
  __syncthreads();
  double* mine = slots + (int64_t)blockIdx.x * NV * F;
  if (m.active && m.rg == 0) {
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        double s = 0.0;
        for (int r = 0; r < m.RP; ++r) s += sh[(a * V + i) * kColThreads + r * FV + m.cg];
        mine[(int64_t)a * F + m.cg * V + i] = s;
      }
  }
  last_block_col_sum<double>(slots, NV * F, gsum, counter, I3D_STATS_STRIDE);
}

template <int V>
__global__ void __launch_bounds__(kColThreads)
    act_colstats_kernel(const float* __restrict__ Y, int64_t M, int F, int ldy, int act, double* __restrict__ sums,
                        const int32_t* __restrict__ m_valid, double* __restrict__ slots, unsigned* __restrict__ counter) {
  pdl_grid_sync();
  if (m_valid) M = min(M, (int64_t)*m_valid);
  const int FV = F / V;
  const ColMap m = col_map(FV);
  double acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[0][i] = 0.0, acc[1][i] = 0.0;
  if (m.active) {
#pragma unroll 4
    for (int64_t r = (int64_t)blockIdx.x * m.RP + m.rg; r < M; r += (int64_t)gridDim.x * m.RP) {
      Vec<V> y;
      y.load(Y + r * ldy + m.cg * V);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const double h = (double)act_apply(y.v[i], act);
        acc[0][i] += h;
        acc[1][i] += h * h;
      }
    }
  }
  block_col_reduce_f64_ws<V, 2>(acc, m, FV, F, sums, slots, counter);
}

// per-column (alpha, beta) with O = h*alpha + beta, exactly the folded form PyTorch's CPU batch_norm uses
__device__ __forceinline__ void bn_column_terms(int c, int64_t M, int F, const double* __restrict__ sums,
                                                const float* __restrict__ running_mean,
                                                const float* __restrict__ running_var,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                float eps, int training, float* mean_out, float* rstd_out,
                                                double* var_biased) {
  double mean, var;
  if (training) {
    mean = sums[(int64_t)c * I3D_STATS_STRIDE] / (double)M;
    var = sums[(int64_t)(F + c) * I3D_STATS_STRIDE] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
  } else {
    mean = (double)running_mean[c];
    var = (double)running_var[c];
  }
  *mean_out = (float)mean;
  *rstd_out = (float)(1.0 / sqrt(var + (double)eps));
  *var_biased = var;
}

template <int V>
__global__ void __launch_bounds__(256)
    bn_apply_kernel(const float* __restrict__ Y, int64_t M, int F, int ldy, int act, const double* __restrict__ sums,
                    float* __restrict__ running_mean, float* __restrict__ running_var,
                    int64_t* __restrict__ num_batches_tracked, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float momentum, float eps, int training,
                    float* __restrict__ save_mean_rstd, const float* __restrict__ residual, float* __restrict__ O,
                    int ldo, const int32_t* __restrict__ m_valid) {
  pdl_grid_sync();
  // shape-bucketed batches: rows [*m_valid, M) are padding — excluded from the statistics, written as zeros
  const int64_t Mrows = M;
  if (m_valid) M = min(M, (int64_t)*m_valid);
  extern __shared__ float shf[];  // alpha[F], beta[F]
  float* s_alpha = shf;
  float* s_beta = shf + F;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float mean, rstd;
    double var;
    bn_column_terms(c, M, F, sums, running_mean, running_var, gamma, beta, eps, training, &mean, &rstd, &var);
    const float a = rstd * gamma[c];
    s_alpha[c] = a;
    s_beta[c] = beta[c] - mean * a;
    if (blockIdx.x == 0) {
      save_mean_rstd[c] = mean;
      save_mean_rstd[F + c] = rstd;
      if (training) {
        const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
  __syncthreads();
  const int FV = F / V;
  const int64_t total = Mrows * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / FV;
    const int c0 = (int)(t - r * FV) * V;
    Vec<V> y, o;
    if (r >= M) {
      o.fill(0.f);
      o.store(O + r * ldo + c0);
      continue;
    }
    y.load(Y + r * ldy + c0);
#pragma unroll
    for (int i = 0; i < V; ++i) o.v[i] = act_apply(y.v[i], act) * s_alpha[c0 + i] + s_beta[c0 + i];
    if (residual) {
      Vec<V> rr;
      rr.load(residual + r * ldo + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = __fadd_rn(o.v[i], rr.v[i]);
    }
    o.store(O + r * ldo + c0);
  }
}

// rows are walked kColBatch at a time: all loads of a batch are issued before any arithmetic, so a thread keeps
// 2*kColBatch 16-byte requests in flight (these kernels are latency-bound at 2 CTAs/SM otherwise)
constexpr int kColBatch = 4;

template <int V>
__global__ void __launch_bounds__(kColThreads)
    bn_bwd_reduce_kernel(const float* __restrict__ dO, int ldd, const float* __restrict__ Y, int ldy, int64_t M,
                         int F, int act, const float* __restrict__ save_mean_rstd, double* __restrict__ sums2,
                         float* __restrict__ zero_buf, int zero_n, const int32_t* __restrict__ m_valid,
                         double* __restrict__ slots, unsigned* __restrict__ counter) {
  pdl_grid_sync();
  if (m_valid) M = min(M, (int64_t)*m_valid);      // padding rows may hold anything (never read)
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < zero_n; i += blockDim.x) zero_buf[i] = 0.f;
  const int FV = F / V;
  const ColMap m = col_map(FV);
  double acc[2][V];
  float mean[V], rstd[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    acc[0][i] = 0.0, acc[1][i] = 0.0;
    mean[i] = m.active ? save_mean_rstd[m.cg * V + i] : 0.f;
    rstd[i] = m.active ? save_mean_rstd[F + m.cg * V + i] : 0.f;
  }
  if (m.active) {
    const int64_t stride = (int64_t)gridDim.x * m.RP;
    const float* yp = Y + m.cg * V;
    const float* dp = dO + m.cg * V;
    int64_t r = (int64_t)blockIdx.x * m.RP + m.rg;
    for (; r + (kColBatch - 1) * stride < M; r += kColBatch * stride) {
      Vec<V> y[kColBatch], d[kColBatch];
#pragma unroll
      for (int j = 0; j < kColBatch; ++j) {
        y[j].load(yp + (r + j * stride) * ldy);
        d[j].load(dp + (r + j * stride) * ldd);
      }
      // fp32 inside a batch of 4 rows, fp64 across batches
      float s0[V], s1[V];
#pragma unroll
      for (int i = 0; i < V; ++i) s0[i] = 0.f, s1[i] = 0.f;
#pragma unroll
      for (int j = 0; j < kColBatch; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xhat = (act_apply(y[j].v[i], act) - mean[i]) * rstd[i];
          s0[i] += d[j].v[i];
          s1[i] = fmaf(d[j].v[i], xhat, s1[i]);
        }
#pragma unroll
      for (int i = 0; i < V; ++i) acc[0][i] += (double)s0[i], acc[1][i] += (double)s1[i];
    }
    for (; r < M; r += stride) {
      Vec<V> y, d;
      y.load(yp + r * ldy);
      d.load(dp + r * ldd);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xhat = (act_apply(y.v[i], act) - mean[i]) * rstd[i];
        acc[0][i] += (double)d.v[i];
        acc[1][i] += (double)d.v[i] * (double)xhat;
      }
    }
  }
  block_col_reduce_f64_ws<V, 2>(acc, m, FV, F, sums2, slots, counter);
}

template <int V>
__global__ void __launch_bounds__(kColThreads)
    bn_bwd_apply_kernel(const float* __restrict__ dO, int ldd, const float* __restrict__ Y, int ldy, int64_t M, int F,
                        int act, int has_bn, int training, const float* __restrict__ save_mean_rstd,
                        const float* __restrict__ gamma, const double* __restrict__ sums2, float* __restrict__ dY,
                        int lddy, float* __restrict__ dbias, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        const int32_t* __restrict__ m_valid, float* __restrict__ slots,
                        unsigned* __restrict__ counter, int db_stride) {
  pdl_grid_sync();
  // shape-bucketed batches: rows [*m_valid, M) are padding — dY is written as exact zeros there (their dO / Y are
  // never read), so nothing downstream (dW = dY^T x, dx = dY W, row sums) sees them
  const int64_t Mrows = M;
  if (m_valid) M = min(M, (int64_t)*m_valid);
  const int FV = F / V;
  const ColMap m = col_map(FV);
  float mean[V], rstd[V], k1[V], k2[V], gr[V], db[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    db[i] = 0.f;
    mean[i] = 0.f, rstd[i] = 1.f, k1[i] = 0.f, k2[i] = 0.f, gr[i] = 1.f;
    if (m.active && has_bn) {
      const int c = m.cg * V + i;
      mean[i] = save_mean_rstd[c];
      rstd[i] = save_mean_rstd[F + c];
      gr[i] = gamma[c] * rstd[i];
      const double s_do = sums2[(int64_t)c * I3D_STATS_STRIDE], s_dx = sums2[(int64_t)(F + c) * I3D_STATS_STRIDE];
      if (training) {
        k1[i] = (float)(s_do / (double)M);
        k2[i] = (float)(s_dx / (double)M);
      }
      if (blockIdx.x == 0 && m.rg == 0) {
        dbeta[c] = (float)s_do;
        dgamma[c] = (float)s_dx;
      }
    }
  }
  if (m.active) {
    const int64_t stride = (int64_t)gridDim.x * m.RP;
    const float* yp = Y + m.cg * V;
    const float* dp = dO + m.cg * V;
    float* op = dY + m.cg * V;
    auto one = [&](const Vec<V>& y, const Vec<V>& d, int64_t row) {
      Vec<V> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float dh = d.v[i];
        if (has_bn) {
          const float xhat = (act_apply(y.v[i], act) - mean[i]) * rstd[i];
          dh = gr[i] * (d.v[i] - k1[i] - xhat * k2[i]);
        }
        const float dy = dh * act_grad(y.v[i], act);
        o.v[i] = dy;
        db[i] += dy;
      }
      o.store(op + row * lddy);
    };
    int64_t r = (int64_t)blockIdx.x * m.RP + m.rg;
    for (; r + (kColBatch - 1) * stride < M; r += kColBatch * stride) {
      Vec<V> y[kColBatch], d[kColBatch];
#pragma unroll
      for (int j = 0; j < kColBatch; ++j) {
        y[j].load(yp + (r + j * stride) * ldy);
        d[j].load(dp + (r + j * stride) * ldd);
      }
#pragma unroll
      for (int j = 0; j < kColBatch; ++j) one(y[j], d[j], r + j * stride);
    }
    for (; r < M; r += stride) {
      Vec<V> y, d;
      y.load(yp + r * ldy);
      d.load(dp + r * ldd);
      one(y, d, r);
    }
    if (Mrows > M) {
      Vec<V> z;
      z.fill(0.f);
      // first padding row of this thread's row phase
      int64_t rz = (int64_t)blockIdx.x * m.RP + m.rg;
      if (rz < M) rz += ((M - rz + stride - 1) / stride) * stride;
      for (; rz < Mrows; rz += stride) z.store(op + rz * lddy);
    }
  }
  if (dbias) {
    extern __shared__ double shd[];
    float* sh = reinterpret_cast<float*>(shd);  // [V][kColThreads]
#pragma unroll
    for (int i = 0; i < V; ++i) sh[i * kColThreads + threadIdx.x] = m.active ? db[i] : 0.f;
    __syncthreads();
    if (m.active && m.rg == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float s = 0.f;
        for (int r = 0; r < m.RP; ++r) s += sh[i * kColThreads + r * FV + m.cg];
        if (slots) slots[(int64_t)blockIdx.x * F + m.cg * V + i] = s;
        else atomicAdd(dbias + (int64_t)(m.cg * V + i) * db_stride, s);      // db_stride 8: one float per 32-byte sector
      }
    }
    if (slots) last_block_col_sum<float>(slots, F, dbias, counter, db_stride);     // (stores: dbias need not be zeroed)
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm backward in ONE launch: the column reduction (bn_bwd_reduce_kernel) and the element-wise pass
// (bn_bwd_apply_kernel) separated by a grid-wide barrier instead of a kernel boundary.  Same row mapping and the same
// arithmetic as the two kernels; every thread keeps its rows of dO in shared memory between the phases (Y is re-read, it
// is L2 resident), the column totals come back through L2.  The barrier is a ticket counter (zeroed by the caller) with
// acquire polling, which needs every CTA of the grid resident at once: the host caps the grid at the occupancy of this
// kernel and only takes this path on ONE stream at a time (two such grids spinning beside each other could starve each
// other of SMs; kernels.bn_bwd), everything else — other streams, wider problems — runs the two-kernel path.
// ------------------------------------------------------------------------------------------------
constexpr int kFusedMaxRows = 16;      // rows of dO per thread kept in shared memory

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int V>
__global__ void __launch_bounds__(kColThreads, 2)
    bn_bwd_fused_kernel(const float* __restrict__ dO, int ldd, const float* __restrict__ Y, int ldy, int64_t M, int F,
                        int act, int training, const float* __restrict__ save_mean_rstd,
                        const float* __restrict__ gamma, double* __restrict__ sums2, float* __restrict__ dY, int lddy,
                        float* __restrict__ dbias, int db_stride, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        float* __restrict__ zero_buf, int zero_n, const int32_t* __restrict__ m_valid,
                        unsigned* __restrict__ barrier) {
  pdl_grid_sync();
  extern __shared__ double sh[];                                           // [2][V][kColThreads] doubles, then the slab
  float* slab = reinterpret_cast<float*>(sh + 2 * V * kColThreads);        // [kFusedMaxRows][kColThreads][V] floats
  const int64_t Mrows = M;
  if (m_valid) M = min(M, (int64_t)*m_valid);
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < zero_n; i += blockDim.x) zero_buf[i] = 0.f;
  const int FV = F / V;
  const ColMap m = col_map(FV);
  const int64_t stride = (int64_t)gridDim.x * m.RP;
  const int64_t r_first = (int64_t)blockIdx.x * m.RP + m.rg;
  const float* yp = Y + m.cg * V;
  const float* dp = dO + m.cg * V;
  float* mine = slab + (size_t)threadIdx.x * V;
  // ---------------- phase 1: column sums of dO and dO * xhat (identical to bn_bwd_reduce_kernel) ----------------
  {
    double acc[2][V];
    float mean[V], rstd[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      acc[0][i] = 0.0, acc[1][i] = 0.0;
      mean[i] = m.active ? save_mean_rstd[m.cg * V + i] : 0.f;
      rstd[i] = m.active ? save_mean_rstd[F + m.cg * V + i] : 0.f;
    }
    if (m.active) {
      int64_t r = r_first;
      int k = 0;
      for (; r + (kColBatch - 1) * stride < M; r += kColBatch * stride, k += kColBatch) {
        Vec<V> y[kColBatch], d[kColBatch];
#pragma unroll
        for (int j = 0; j < kColBatch; ++j) {
          y[j].load(yp + (r + j * stride) * ldy);
          d[j].load(dp + (r + j * stride) * ldd);
        }
        float s0[V], s1[V];
#pragma unroll
        for (int i = 0; i < V; ++i) s0[i] = 0.f, s1[i] = 0.f;
#pragma unroll
        for (int j = 0; j < kColBatch; ++j) {
          d[j].store(mine + (size_t)(k + j) * kColThreads * V);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float xhat = (act_apply(y[j].v[i], act) - mean[i]) * rstd[i];
            s0[i] += d[j].v[i];
            s1[i] = fmaf(d[j].v[i], xhat, s1[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) acc[0][i] += (double)s0[i], acc[1][i] += (double)s1[i];
      }
      for (; r < M; r += stride, ++k) {
        Vec<V> y, d;
        y.load(yp + r * ldy);
        d.load(dp + r * ldd);
        d.store(mine + (size_t)k * kColThreads * V);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xhat = (act_apply(y.v[i], act) - mean[i]) * rstd[i];
          acc[0][i] += (double)d.v[i];
          acc[1][i] += (double)d.v[i] * (double)xhat;
        }
      }
    }
    block_col_reduce_f64<V, 2>(acc, m, FV, F, sums2);
  }
  // ---------------- grid barrier: every CTA's partial sums (and CTA 0's zero fill) are in L2 ----------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(barrier, 1u);
    while (ld_acquire_u32(barrier) < gridDim.x) __nanosleep(32);
  }
  __syncthreads();
  // ---------------- phase 2: dY (identical to bn_bwd_apply_kernel with has_bn = 1) ----------------
  float mean[V], rstd[V], k1[V], k2[V], gr[V], db[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    db[i] = 0.f;
    mean[i] = 0.f, rstd[i] = 1.f, k1[i] = 0.f, k2[i] = 0.f, gr[i] = 1.f;
    if (m.active) {
      const int c = m.cg * V + i;
      mean[i] = save_mean_rstd[c];
      rstd[i] = save_mean_rstd[F + c];
      gr[i] = gamma[c] * rstd[i];
      const double s_do = __ldcg(sums2 + (int64_t)c * I3D_STATS_STRIDE);
      const double s_dx = __ldcg(sums2 + (int64_t)(F + c) * I3D_STATS_STRIDE);
      if (training) {
        k1[i] = (float)(s_do / (double)M);
        k2[i] = (float)(s_dx / (double)M);
      }
      if (blockIdx.x == 0 && m.rg == 0) {
        dbeta[c] = (float)s_do;
        dgamma[c] = (float)s_dx;
      }
    }
  }
  if (m.active) {
    float* op = dY + m.cg * V;
    int k = 0;
    for (int64_t r = r_first; r < M; r += stride, ++k) {
      Vec<V> y, d, o;
      y.load(yp + r * ldy);
      d.load_shared(mine + (size_t)k * kColThreads * V);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xhat = (act_apply(y.v[i], act) - mean[i]) * rstd[i];
        const float dh = gr[i] * (d.v[i] - k1[i] - xhat * k2[i]);
        const float dy = dh * act_grad(y.v[i], act);
        o.v[i] = dy;
        db[i] += dy;
      }
      o.store(op + r * lddy);
    }
    if (Mrows > M) {
      Vec<V> z;
      z.fill(0.f);
      int64_t rz = r_first;
      if (rz < M) rz += ((M - rz + stride - 1) / stride) * stride;
      for (; rz < Mrows; rz += stride) z.store(op + rz * lddy);
    }
  }
  if (dbias) {
    float* shf = reinterpret_cast<float*>(sh);  // [V][kColThreads] (the reduction scratch of phase 1 is free again)
#pragma unroll
    for (int i = 0; i < V; ++i) shf[i * kColThreads + threadIdx.x] = m.active ? db[i] : 0.f;
    __syncthreads();
    if (m.active && m.rg == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float s = 0.f;
        for (int r = 0; r < m.RP; ++r) s += shf[i * kColThreads + r * FV + m.cg];
        atomicAdd(dbias + (int64_t)(m.cg * V + i) * db_stride, s);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Factored first layer of the edge MLP (models/pna.py:237-252).  cat[h[src], h[dst], e] W^T splits by columns of W into
// (h W_s^T)[src] + (h W_d^T)[dst] + e W_e^T: the two h terms are ONE node-level GEMM P = h [W_s;W_d]^T
// (N rows instead of E, K = F instead of 3F) and, because the bond features take only prod(5,6,2) = 60 distinct values
// (commons/mol_encoder.py:4-7), e W_e^T is a 60-row table T looked up by the edge's feature code.  This kernel is what
// is left at edge level:  Y[m,:] = P[src[m], 0:F] + P[dst[m], F:2F] + T[code[m], :] + bias, with the FCLayer's train-mode
// BatchNorm statistics (column sums of act(Y), act(Y)^2 in fp64) fused, as the GEMM epilogue does for the other layers.
// HBM/L2-bound: P (N x 2F) and T stay L2 / L1 resident, Y is written once.  Negative src/dst (padding edges) read zeros.
// ------------------------------------------------------------------------------------------------
constexpr int kGaBatch = 2;      // rows in flight per thread: 3 CTAs/SM x 2 rows beat 2 CTAs/SM x 4 rows (registers)

template <int V>
__global__ void __launch_bounds__(kColThreads, 3)
    edge_gather_add_kernel(const float* __restrict__ P, int ldp, const int32_t* __restrict__ src,
                           const int32_t* __restrict__ dst, const float* __restrict__ T, int ldt,
                           const int32_t* __restrict__ code, const float* __restrict__ bias, int64_t M, int F,
                           float* __restrict__ Y, int ldy, double* __restrict__ stats, int act,
                           const int32_t* __restrict__ m_valid, double* __restrict__ slots,
                           unsigned* __restrict__ counter) {
  pdl_grid_sync();
  const int FV = F / V;
  const ColMap m = col_map(FV);
  const int64_t Mv = m_valid ? min(M, (int64_t)*m_valid) : M;
  double acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[0][i] = 0.0, acc[1][i] = 0.0;
  if (m.active) {
    const int c0 = m.cg * V;
    Vec<V> b;
    if (bias) b.load(bias + c0); else b.fill(0.f);
    const int64_t stride = (int64_t)gridDim.x * m.RP;
    // the gather indices of a batch of rows are loaded one batch ahead, so that no row load waits on an index load
    int32_t ns[kGaBatch], nd[kGaBatch], nc[kGaBatch];
    auto load_idx = [&](int64_t rb) {
#pragma unroll
      for (int j = 0; j < kGaBatch; ++j) {
        const int64_t r = rb + j * stride;
        ns[j] = nd[j] = -1, nc[j] = 0;
        if (r < M) ns[j] = __ldg(src + r), nd[j] = __ldg(dst + r), nc[j] = code ? __ldg(code + r) : 0;
      }
    };
    load_idx((int64_t)blockIdx.x * m.RP + m.rg);
    for (int64_t r0 = (int64_t)blockIdx.x * m.RP + m.rg; r0 < M; r0 += kGaBatch * stride) {
      int32_t is[kGaBatch], id[kGaBatch], ic[kGaBatch];
#pragma unroll
      for (int j = 0; j < kGaBatch; ++j) is[j] = ns[j], id[j] = nd[j], ic[j] = nc[j];
      load_idx(r0 + kGaBatch * stride);
      Vec<V> a[kGaBatch], d[kGaBatch], t[kGaBatch];
#pragma unroll
      for (int j = 0; j < kGaBatch; ++j) {
        a[j].fill(0.f), d[j].fill(0.f), t[j].fill(0.f);
        if (is[j] >= 0) a[j].load(P + (int64_t)is[j] * ldp + c0);
        if (id[j] >= 0) d[j].load(P + (int64_t)id[j] * ldp + F + c0);
        if (T && r0 + j * stride < M) t[j].load(T + (int64_t)ic[j] * ldt + c0);
      }
#pragma unroll
      for (int j = 0; j < kGaBatch; ++j) {
        const int64_t r = r0 + j * stride;
        if (r >= M) continue;
        Vec<V> y;
#pragma unroll
        for (int i = 0; i < V; ++i)
          y.v[i] = __fadd_rn(__fadd_rn(__fadd_rn(a[j].v[i], d[j].v[i]), t[j].v[i]), b.v[i]);
        y.store(Y + r * ldy + c0);
        if (stats && r < Mv) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const double h = (double)act_apply(y.v[i], act);
            acc[0][i] += h;
            acc[1][i] = fma(h, h, acc[1][i]);
          }
        }
      }
    }
  }
  if (stats) block_col_reduce_f64_ws<V, 2>(acc, m, FV, F, stats, slots, counter);
}

__global__ void act_fwd_kernel(const float* __restrict__ x, int64_t n, int act, float* __restrict__ y) {
  pdl_grid_sync();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = act_apply(x[i], act);
}
__global__ void act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, int64_t n, int act,
                               float* __restrict__ gx) {
  pdl_grid_sync();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    gx[i] = gy[i] * act_grad(x[i], act);
}

// column sums in fp32 (bias gradients of layers without BN, node_embedding gradient)
template <int V>
__global__ void __launch_bounds__(kColThreads)
    colsum_kernel(const float* __restrict__ x, int ldx, int64_t M, int F, float* __restrict__ out) {
  pdl_grid_sync();
  const int FV = F / V;
  const ColMap m = col_map(FV);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (m.active) {
    for (int64_t r = (int64_t)blockIdx.x * m.RP + m.rg; r < M; r += (int64_t)gridDim.x * m.RP) {
      Vec<V> v;
      v.load(x + r * ldx + m.cg * V);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += v.v[i];
    }
  }
  extern __shared__ double shd[];
  float* sh = reinterpret_cast<float*>(shd);
#pragma unroll
  for (int i = 0; i < V; ++i) sh[i * kColThreads + threadIdx.x] = m.active ? acc[i] : 0.f;
  __syncthreads();
  if (m.active && m.rg == 0) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float s = 0.f;
      for (int r = 0; r < m.RP; ++r) s += sh[i * kColThreads + r * FV + m.cg];
      atomicAdd(out + m.cg * V + i, s);
    }
  }
}

static inline int col_grid(int64_t M, int FV) {
  const int RP = kColThreads / FV;
  // every CTA ends with 2F same-address atomics (LTS serialises them per address), so the grid is capped at a few
  // CTAs per SM; I3D_COL_CTAS_PER_SM overrides the cap for tuning runs
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = getenv("I3D_COL_CTAS_PER_SM");
    per_sm = e ? atoi(e) : 2;
    if (per_sm < 1 || per_sm > 16) per_sm = 2;
  }
  int64_t need = (M + (int64_t)RP * 8 - 1) / ((int64_t)RP * 8);  // >= 8 rows per thread
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace i3d

using namespace i3d;

extern "C" {

int i3d_act_colstats(const float* Y, int64_t M, int F, int ldy, int act, double* sums, void* stream) {
  return i3d_act_colstats_v(Y, M, F, ldy, act, sums, nullptr, nullptr, stream);
}

// two-stage reduction workspace usable for `grid` CTAs x `ncols` columns of `elem` bytes?
static inline bool rws_ok(const i3d_reduce_ws* r, int grid, int ncols, size_t elem) {
  return r && r->slots && r->counter && (size_t)grid * ncols * elem <= (size_t)r->slot_bytes;
}

int i3d_act_colstats_v(const float* Y, int64_t M, int F, int ldy, int act, double* sums, const int32_t* m_valid,
                       const i3d_reduce_ws* rws, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldy >= F && sums && (M == 0 || Y), "invalid argument");
  cudaStream_t s = as_stream(stream);
  I3D_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * F * I3D_STATS_STRIDE, s));
  if (M == 0) return I3D_OK;
  const bool v4 = can_vec4({Y}, {F, ldy});
  const int V = v4 ? 4 : 1, FV = F / V;
  I3D_REQUIRE(FV <= kColThreads, "feature width too large (F <= 1024 when 16B-aligned, else F <= 256)");
  const size_t smem = sizeof(double) * 2 * V * kColThreads;
  const int grid = col_grid(M, FV);
  const bool two = rws_ok(rws, grid, 2 * F, sizeof(double));
  double* slots = two ? static_cast<double*>(rws->slots) : nullptr;
  unsigned* counter = two ? rws->counter : nullptr;
  if (v4)
    launch(act_colstats_kernel<4>, grid, kColThreads, smem, s, Y, M, F, ldy, act, sums, m_valid, slots, counter);
  else
    launch(act_colstats_kernel<1>, grid, kColThreads, smem, s, Y, M, F, ldy, act, sums, m_valid, slots, counter);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_bn_apply(const float* Y, int64_t M, int F, int ldy, int act, const double* sums, float* running_mean,
                 float* running_var, int64_t* num_batches_tracked, const float* gamma, const float* beta,
                 float momentum, float eps, int training, float* save_mean_rstd, const float* residual, float* O,
                 int ldo, void* stream) {
  return i3d_bn_apply_v(Y, M, F, ldy, act, sums, running_mean, running_var, num_batches_tracked, gamma, beta, momentum,
                        eps, training, save_mean_rstd, residual, O, ldo, nullptr, stream);
}

int i3d_bn_apply_v(const float* Y, int64_t M, int F, int ldy, int act, const double* sums, float* running_mean,
                   float* running_var, int64_t* num_batches_tracked, const float* gamma, const float* beta,
                   float momentum, float eps, int training, float* save_mean_rstd, const float* residual, float* O,
                   int ldo, const int32_t* m_valid, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldy >= F && ldo >= F && gamma && beta && save_mean_rstd && running_mean &&
                  running_var && (!training || sums) && (M == 0 || (Y && O)),
              "invalid argument");
  I3D_REQUIRE(!(training && M < 2), "Expected more than 1 value per channel when training");
  if (M == 0) return I3D_OK;
  const bool v4 = can_vec4({Y, O, residual}, {F, ldy, ldo});
  const int64_t work = M * (F / (v4 ? 4 : 1));
  const size_t smem = sizeof(float) * 2 * F;
  cudaStream_t s = as_stream(stream);
  if (v4)
    launch(bn_apply_kernel<4>, grid_for(work, 256), 256, smem, s, Y, M, F, ldy, act, sums, running_mean, running_var,
                                                              num_batches_tracked, gamma, beta, momentum, eps,
                                                              training, save_mean_rstd, residual, O, ldo, m_valid);
  else
    launch(bn_apply_kernel<1>, grid_for(work, 256), 256, smem, s, Y, M, F, ldy, act, sums, running_mean, running_var,
                                                              num_batches_tracked, gamma, beta, momentum, eps,
                                                              training, save_mean_rstd, residual, O, ldo, m_valid);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_bn_bwd_reduce_ex(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act,
                         const float* save_mean_rstd, double* sums2, float* zero_buf, int zero_n, void* stream) {
  return i3d_bn_bwd_reduce_v(dO, ldd, Y, ldy, M, F, act, save_mean_rstd, sums2, zero_buf, zero_n, nullptr, nullptr,
                             stream);
}

int i3d_bn_bwd_reduce_v(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act,
                        const float* save_mean_rstd, double* sums2, float* zero_buf, int zero_n,
                        const int32_t* m_valid, const i3d_reduce_ws* rws, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldy >= F && ldd >= F && sums2 && save_mean_rstd && (M == 0 || (Y && dO)) &&
                  zero_n >= 0 && (zero_n == 0 || zero_buf),
              "invalid argument");
  cudaStream_t s = as_stream(stream);
  const bool prezeroed = (act & I3D_STATS_PREZEROED) != 0;
  act &= 0xff;
  if (!prezeroed) I3D_CUDA(cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * F * I3D_STATS_STRIDE, s));
  if (M == 0) {
    if (zero_n > 0) I3D_CUDA(cudaMemsetAsync(zero_buf, 0, sizeof(float) * zero_n, s));
    return I3D_OK;
  }
  const bool v4 = can_vec4({Y, dO}, {F, ldy, ldd});
  const int V = v4 ? 4 : 1, FV = F / V;
  I3D_REQUIRE(FV <= kColThreads, "feature width too large");
  const size_t smem = sizeof(double) * 2 * V * kColThreads;
  const int grid = col_grid(M, FV);
  const bool two = rws_ok(rws, grid, 2 * F, sizeof(double));
  double* slots = two ? static_cast<double*>(rws->slots) : nullptr;
  unsigned* counter = two ? rws->counter : nullptr;
  if (v4)
    launch(bn_bwd_reduce_kernel<4>, grid, kColThreads, smem, s, dO, ldd, Y, ldy, M, F, act, save_mean_rstd,
                                                                      sums2, zero_buf, zero_n, m_valid, slots, counter);
  else
    launch(bn_bwd_reduce_kernel<1>, grid, kColThreads, smem, s, dO, ldd, Y, ldy, M, F, act, save_mean_rstd,
                                                                      sums2, zero_buf, zero_n, m_valid, slots, counter);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_bn_bwd_reduce(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act,
                      const float* save_mean_rstd, double* sums2, void* stream) {
  return i3d_bn_bwd_reduce_ex(dO, ldd, Y, ldy, M, F, act, save_mean_rstd, sums2, nullptr, 0, stream);
}

int i3d_bn_bwd_apply(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act, int has_bn,
                     int training, const float* save_mean_rstd, const float* gamma, const double* sums2, float* dY,
                     int lddy, float* dbias, float* dgamma, float* dbeta, void* stream) {
  return i3d_bn_bwd_apply_v(dO, ldd, Y, ldy, M, F, act, has_bn, training, save_mean_rstd, gamma, sums2, dY, lddy, dbias, 1,
                            dgamma, dbeta, nullptr, nullptr, stream);
}

int i3d_bn_bwd_apply_v(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act, int has_bn,
                       int training, const float* save_mean_rstd, const float* gamma, const double* sums2, float* dY,
                       int lddy, float* dbias, int dbias_stride, float* dgamma, float* dbeta, const int32_t* m_valid,
                       const i3d_reduce_ws* rws, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldy >= F && ldd >= F && lddy >= F && (M == 0 || (Y && dO && dY)) && dbias_stride >= 1,
              "invalid argument");
  I3D_REQUIRE(!has_bn || (save_mean_rstd && gamma && sums2 && dgamma && dbeta), "BN tensors missing");
  if (M == 0) return I3D_OK;
  const bool v4 = can_vec4({Y, dO, dY}, {F, ldy, ldd, lddy});
  const int V = v4 ? 4 : 1, FV = F / V;
  I3D_REQUIRE(FV <= kColThreads, "feature width too large");
  const size_t smem = sizeof(float) * V * kColThreads;
  cudaStream_t s = as_stream(stream);
  const int grid = col_grid(M, FV);
  const bool two = dbias && rws_ok(rws, grid, F, sizeof(float));
  float* slots = two ? static_cast<float*>(rws->slots) : nullptr;
  unsigned* counter = two ? rws->counter : nullptr;
  if (v4)
    launch(bn_bwd_apply_kernel<4>, grid, kColThreads, smem, s, dO, ldd, Y, ldy, M, F, act, has_bn, training,
                                                                     save_mean_rstd, gamma, sums2, dY, lddy, dbias,
                                                                     dgamma, dbeta, m_valid, slots, counter, dbias_stride);
  else
    launch(bn_bwd_apply_kernel<1>, grid, kColThreads, smem, s, dO, ldd, Y, ldy, M, F, act, has_bn, training,
                                                                     save_mean_rstd, gamma, sums2, dY, lddy, dbias,
                                                                     dgamma, dbeta, m_valid, slots, counter, dbias_stride);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_bn_bwd_fused_v(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act, int training,
                       const float* save_mean_rstd, const float* gamma, double* sums2, float* dY, int lddy,
                       float* dbias, int dbias_stride, float* dgamma, float* dbeta, float* zero_buf, int zero_n,
                       const int32_t* m_valid, unsigned int* barrier, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldy >= F && ldd >= F && lddy >= F && sums2 && save_mean_rstd && gamma && dgamma &&
                  dbeta && (M == 0 || (Y && dO && dY)) && dbias_stride >= 1 && zero_n >= 0 && (zero_n == 0 || zero_buf),
              "invalid argument");
  cudaStream_t s = as_stream(stream);
  const bool v4 = can_vec4({Y, dO, dY}, {F, ldy, ldd, lddy});
  const int V = v4 ? 4 : 1, FV = F / V;
  bool fused = barrier != nullptr && M > 0 && FV <= kColThreads;
  int grid = 0;
  size_t smem = 0;
  if (fused) {
    // I3D_BN_BWD=fused turns the one-launch path on.  Measured on B200 (batch 512, bench.py): 4.130 ms/step fused vs
    // 4.146 ms with the two kernels — the barrier costs what the kernel boundary cost under programmatic dependent
    // launch and only the second read of dO (an L2 hit) goes away — so the default stays the two-kernel path, which
    // has no spinning barrier.
    static int mode = -1;
    if (mode < 0) {
      const char* e = getenv("I3D_BN_BWD");
      mode = (e && e[0] == 'f') ? 1 : 0;
    }
    fused = mode == 1;
  }
  if (fused) {
    smem = sizeof(double) * 2 * V * kColThreads + sizeof(float) * (size_t)kFusedMaxRows * kColThreads * V;
    static int per_sm[2] = {-1, -1};       // resident CTAs per SM of the two instantiations (0 = unusable)
    int& nb = per_sm[v4 ? 1 : 0];
    if (nb < 0) {
      nb = 0;
      cudaError_t e = v4 ? cudaFuncSetAttribute(bn_bwd_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(bn_bwd_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      int n = 0;
      if (e == cudaSuccess)
        e = v4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bn_bwd_fused_kernel<4>, kColThreads, smem)
               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bn_bwd_fused_kernel<1>, kColThreads, smem);
      if (e == cudaSuccess) nb = n;
      else cudaGetLastError();
    }
    const int RP = kColThreads / FV;
    const int64_t cap = (int64_t)nb * sm_count();
    grid = col_grid(M, FV);
    if (grid > cap) grid = (int)cap;
    // every thread must be able to keep its rows of dO: ceil(M / (grid * RP)) <= kFusedMaxRows
    fused = grid >= 1 && (M + (int64_t)grid * RP - 1) / ((int64_t)grid * RP) <= kFusedMaxRows;
  }
  if (!fused) {
    if (int rc = i3d_bn_bwd_reduce_v(dO, ldd, Y, ldy, M, F, act, save_mean_rstd, sums2, zero_buf, zero_n, m_valid,
                                     nullptr, stream))
      return rc;
    return i3d_bn_bwd_apply_v(dO, ldd, Y, ldy, M, F, act & 0xff, 1, training, save_mean_rstd, gamma, sums2, dY, lddy,
                              dbias, dbias_stride, dgamma, dbeta, m_valid, nullptr, stream);
  }
  if (!(act & I3D_STATS_PREZEROED)) I3D_CUDA(cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * F * I3D_STATS_STRIDE, s));
  act &= 0xff;
  if (v4)
    launch(bn_bwd_fused_kernel<4>, grid, kColThreads, smem, s, dO, ldd, Y, ldy, M, F, act, training, save_mean_rstd,
           gamma, sums2, dY, lddy, dbias, dbias_stride, dgamma, dbeta, zero_buf, zero_n, m_valid, barrier);
  else
    launch(bn_bwd_fused_kernel<1>, grid, kColThreads, smem, s, dO, ldd, Y, ldy, M, F, act, training, save_mean_rstd,
           gamma, sums2, dY, lddy, dbias, dbias_stride, dgamma, dbeta, zero_buf, zero_n, m_valid, barrier);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_edge_gather_add(const float* P, int ldp, const int32_t* src, const int32_t* dst, const float* T, int ldt,
                        const int32_t* code, const float* bias, int64_t M, int F, float* Y, int ldy, double* col_stats,
                        int stats_act, const int32_t* m_valid, const i3d_reduce_ws* rws, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldp >= 2 * F && ldy >= F && (!T || ldt >= F) && (M == 0 || (P && src && dst && Y)),
              "invalid argument");
  cudaStream_t s = as_stream(stream);
  const bool prezeroed = (stats_act & I3D_STATS_PREZEROED) != 0;
  stats_act &= 0xff;
  if (col_stats && !prezeroed) I3D_CUDA(cudaMemsetAsync(col_stats, 0, sizeof(double) * 2 * F * I3D_STATS_STRIDE, s));
  if (M == 0) return I3D_OK;
  const bool v4 = can_vec4({P, T, bias, Y}, {F, ldp, ldy, T ? ldt : 0});
  const int V = v4 ? 4 : 1, FV = F / V;
  I3D_REQUIRE(FV <= kColThreads, "feature width too large");
  const size_t smem = sizeof(double) * 2 * V * kColThreads;
  const int RP = kColThreads / FV;
  int64_t need = (M + (int64_t)RP * 4 - 1) / ((int64_t)RP * 4);          // >= 4 rows per thread
  const int64_t cap = (int64_t)sm_count() * 3;
  const int grid = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  const bool two = col_stats && rws_ok(rws, grid, 2 * F, sizeof(double));
  double* slots = two ? static_cast<double*>(rws->slots) : nullptr;
  unsigned* counter = two ? rws->counter : nullptr;
  if (v4)
    launch(edge_gather_add_kernel<4>, grid, kColThreads, smem, s, P, ldp, src, dst, T, ldt, code, bias, M, F, Y,
           ldy, col_stats, stats_act, m_valid, slots, counter);
  else
    launch(edge_gather_add_kernel<1>, grid, kColThreads, smem, s, P, ldp, src, dst, T, ldt, code, bias, M, F, Y,
           ldy, col_stats, stats_act, m_valid, slots, counter);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_act_fwd(const float* x, int64_t n, int act, float* y, void* stream) {
  I3D_REQUIRE(n >= 0 && (n == 0 || (x && y)), "invalid argument");
  if (n == 0) return I3D_OK;
  launch(act_fwd_kernel, grid_for(n, 256), 256, 0, as_stream(stream), x, n, act, y);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_act_bwd(const float* gy, const float* x, int64_t n, int act, float* gx, void* stream) {
  I3D_REQUIRE(n >= 0 && (n == 0 || (gy && x && gx)), "invalid argument");
  if (n == 0) return I3D_OK;
  launch(act_bwd_kernel, grid_for(n, 256), 256, 0, as_stream(stream), gy, x, n, act, gx);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_colsum(const float* x, int ldx, int64_t M, int F, float* out, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && ldx >= F && out && (M == 0 || x), "invalid argument");
  cudaStream_t s = as_stream(stream);
  I3D_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * F, s));
  if (M == 0) return I3D_OK;
  const bool v4 = can_vec4({x}, {F, ldx});
  const int V = v4 ? 4 : 1, FV = F / V;
  I3D_REQUIRE(FV <= kColThreads, "feature width too large");
  const size_t smem = sizeof(float) * V * kColThreads;
  if (v4)
    launch(colsum_kernel<4>, col_grid(M, FV), kColThreads, smem, s, x, ldx, M, F, out);
  else
    launch(colsum_kernel<1>, col_grid(M, FV), kColThreads, smem, s, x, ldx, M, F, out);
  I3D_LAUNCHED();
  return I3D_OK;
}
}
