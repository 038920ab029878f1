// NTXent / NTXentMultiplePositives row kernels (commons/losses.py:143-155, 225-246) and the optimizer step.
// The similarity matrix itself comes from i3d_gemm (z1 z2^T); these kernels turn it into exp(sim/tau),
// the per-row positives / negatives sums, the loss rows, and in backward d loss / d dot plus the norm terms.
#include <initializer_list>

#include "i3d_vec.cuh"

namespace i3d {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

__global__ void row_norms_kernel(const float* __restrict__ z, int64_t R, int D, float* __restrict__ norms) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < R; r += nw) {
    float s = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float x = __ldg(z + r * D + c);
      s = fmaf(x, x, s);
    }
    s = warp_sum(s);
    if (lane == 0) norms[r] = sqrtf(s);
  }
}

// one CTA per local row i
__global__ void __launch_bounds__(256)
    ntxent_rows_fwd_kernel(float* __restrict__ P, int64_t B, int64_t Bc, int C, const float* __restrict__ n1,
                           const float* __restrict__ n2, int norm, float eps, float tau, int64_t row_offset,
                           float* __restrict__ rowstats, float* __restrict__ loss_rows) {
  pdl_grid_sync();
  __shared__ float sh[32];
  const int64_t ncol = Bc * C;
  for (int64_t i = blockIdx.x; i < B; i += gridDim.x) {
    float* row = P + i * ncol;
    const float a = norm ? n1[i] : 0.f;
    const int64_t pb = (row_offset + i) * C, pe = pb + C;
    float tot = 0.f, pos = 0.f;
    for (int64_t j = threadIdx.x; j < ncol; j += blockDim.x) {
      float s = row[j];
      if (norm) s = __fdiv_rn(s, __fadd_rn(__fmul_rn(a, __ldg(n2 + j)), eps));
      const float pv = expf(__fdiv_rn(s, tau));
      row[j] = pv;
      tot += pv;
      if (j >= pb && j < pe) pos += pv;
    }
    tot = block_sum(tot, sh);
    pos = block_sum(pos, sh);
    if (threadIdx.x == 0) {
      const float neg = tot - pos;
      rowstats[2 * i] = pos;
      rowstats[2 * i + 1] = neg;
      loss_rows[i] = -logf(pos / neg);
    }
  }
}

__global__ void __launch_bounds__(256)
    ntxent_rows_bwd_kernel(float* __restrict__ P, int64_t B, int64_t Bc, int C, const float* __restrict__ n1,
                           const float* __restrict__ n2, int norm, float eps, float tau, int64_t row_offset,
                           const float* __restrict__ rowstats, const float* __restrict__ gout, float inv_B,
                           float* __restrict__ dn1, float* __restrict__ dn2) {
  pdl_grid_sync();
  __shared__ float sh[32];
  const int64_t ncol = Bc * C;
  const float gl = gout[0] * inv_B;
  for (int64_t i = blockIdx.x; i < B; i += gridDim.x) {
    float* row = P + i * ncol;
    const float a = norm ? n1[i] : 0.f;
    const int64_t pb = (row_offset + i) * C, pe = pb + C;
    const float pos = rowstats[2 * i], neg = rowstats[2 * i + 1];
    const float c_pos = -gl / pos, c_neg = gl / neg;
    float acc1 = 0.f;
    for (int64_t j = threadIdx.x; j < ncol; j += blockDim.x) {
      const float pv = row[j];
      const float ds = ((j >= pb && j < pe) ? c_pos : c_neg) * pv / tau;
      float dd = ds;
      if (norm) {
        const float b = __ldg(n2 + j);
        const float den = a * b + eps;
        const float s = tau * logf(pv);  // sim recovered from p = exp(sim/tau)
        dd = ds / den;
        const float t = -ds * s / den;   // d/d den
        acc1 += t * b;
        atomicAdd(dn2 + j, t * a);
      }
      row[j] = dd;
    }
    if (norm) {
      acc1 = block_sum(acc1, sh);
      if (threadIdx.x == 0) dn1[i] = acc1;
    }
  }
}

__global__ void sum_scaled_kernel(const float* __restrict__ x, int64_t n, float scale, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float sh[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s * scale;
}

__global__ void norm_bwd_accum_kernel(const float* __restrict__ z, const float* __restrict__ norms,
                                      const float* __restrict__ dn, int64_t R, int D, float* __restrict__ dz) {
  pdl_grid_sync();
  const int64_t total = R * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / D;
    dz[t] += dn[r] * z[t] / norms[r];
  }
}

// hyper (host or device copy): [lr, beta1, beta2, eps, weight_decay, grad_scale] in fp64, exactly the Python
// floats torch.optim.Adam computes its bias corrections from.
struct AdamHyper {
  double lr, beta1, beta2, eps, weight_decay, grad_scale;
};

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, AdamHyper h, int64_t step,
                            const double* __restrict__ hyper_dev, const int64_t* __restrict__ step_dev) {
  pdl_grid_sync();
  __shared__ float sc[7];
  if (threadIdx.x == 0) {
    if (hyper_dev) {
      h.lr = hyper_dev[0], h.beta1 = hyper_dev[1], h.beta2 = hyper_dev[2];
      h.eps = hyper_dev[3], h.weight_decay = hyper_dev[4], h.grad_scale = hyper_dev[5];
    }
    if (step_dev) step = step_dev[0];
    const double bc1 = 1.0 - pow(h.beta1, (double)step);
    const double bc2 = 1.0 - pow(h.beta2, (double)step);
    sc[0] = (float)(h.lr / bc1);
    sc[1] = (float)sqrt(bc2);
    sc[2] = (float)h.beta1;
    sc[3] = (float)h.beta2;
    sc[4] = (float)h.eps;
    sc[5] = (float)h.weight_decay;
    sc[6] = (float)h.grad_scale;
  }
  __syncthreads();
  const float step_size = sc[0], bc2_sqrt = sc[1], beta1 = sc[2], beta2 = sc[3], eps = sc[4], wd = sc[5], gs = sc[6];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

// ------------------------------------------------------------------------------------------------
// Data-parallel optimizer step as ONE kernel over NVLink/NVSwitch multicast memory: gradient all-reduce + Adam +
// parameter broadcast (replaces NCCL all-reduce(s) followed by the Adam kernel; trainer/trainer.py:119-123 under DDP).
// The flat buffers [p | g | m | v] (n floats each) of every rank live in ONE symmetric allocation bound to a multicast
// address.  Rank r owns the slice [r*per, (r+1)*per) of the elements:
//   g_sum  = multimem.ld_reduce.add  over the slice of g     (the switch adds the 8 ranks' values: 1/8 of the buffer
//                                                             crosses this GPU's links instead of 2x the whole buffer)
//   Adam on the slice with the local copies of p, m, v       (replicas are identical)
//   multimem.st of the new p, m, v                           (the switch writes them into every rank's buffers)
// Cross-rank ordering (gradients complete before, stores landed after) is the caller's: a signal-pad barrier on the
// stream before and after this kernel (FusedAdam.step).  fp32 throughout; the sum order is the switch's.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256)
    adam_nvls_kernel(float* __restrict__ mc, float* __restrict__ local, int64_t n, int rank, int world, AdamHyper h,
                     int64_t step, const double* __restrict__ hyper_dev, const int64_t* __restrict__ step_dev) {
  pdl_grid_sync();
  __shared__ float sc[7];
  if (threadIdx.x == 0) {
    if (hyper_dev) {
      h.lr = hyper_dev[0], h.beta1 = hyper_dev[1], h.beta2 = hyper_dev[2];
      h.eps = hyper_dev[3], h.weight_decay = hyper_dev[4], h.grad_scale = hyper_dev[5];
    }
    if (step_dev) step = step_dev[0];
    const double bc1 = 1.0 - pow(h.beta1, (double)step);
    const double bc2 = 1.0 - pow(h.beta2, (double)step);
    sc[0] = (float)(h.lr / bc1), sc[1] = (float)sqrt(bc2), sc[2] = (float)h.beta1, sc[3] = (float)h.beta2;
    sc[4] = (float)h.eps, sc[5] = (float)h.weight_decay, sc[6] = (float)h.grad_scale;
  }
  __syncthreads();
  const float step_size = sc[0], bc2_sqrt = sc[1], beta1 = sc[2], beta2 = sc[3], eps = sc[4], wd = sc[5], gs = sc[6];
  const int64_t q = n >> 2;                                   // float4 elements per array (n is a multiple of 4)
  const int64_t per = (q + world - 1) / world;
  const int64_t lo = (int64_t)rank * per, hi = lo + per < q ? lo + per : q;
  const float4* p4 = reinterpret_cast<const float4*>(local);
  const float4* m4 = reinterpret_cast<const float4*>(local + 2 * n);
  const float4* v4 = reinterpret_cast<const float4*>(local + 3 * n);
  for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gsum = multimem_ld_reduce_add(mc + n + 4 * i);
    const float4 pv = p4[i], mv = m4[i], vv = v4[i];
    const float g[4] = {gsum.x, gsum.y, gsum.z, gsum.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
    const float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vvv[4] = {vv.x, vv.y, vv.z, vv.w};
    float po[4], mo[4], vo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gi = g[j] * gs;
      if (wd != 0.f) gi = fmaf(wd, pp[j], gi);
      mo[j] = mm[j] + (1.f - beta1) * (gi - mm[j]);
      vo[j] = vvv[j] * beta2 + (1.f - beta2) * gi * gi;
      const float denom = sqrtf(vo[j]) / bc2_sqrt + eps;
      po[j] = pp[j] - step_size * (mo[j] / denom);
    }
    multimem_st(mc + 4 * i, make_float4(po[0], po[1], po[2], po[3]));
    multimem_st(mc + 2 * n + 4 * i, make_float4(mo[0], mo[1], mo[2], mo[3]));
    multimem_st(mc + 3 * n + 4 * i, make_float4(vo[0], vo[1], vo[2], vo[3]));
  }
}

__global__ void add_i64_kernel(int64_t* x, int64_t d) {
  pdl_grid_sync(); x[0] += d; }

__global__ void multi_copy_kernel(const uint64_t* __restrict__ ptrs, const int64_t* __restrict__ off,
                                  const int64_t* __restrict__ len, const int64_t* __restrict__ stride,
                                  float* __restrict__ flat, int to_flat) {
  pdl_grid_sync();
  const int t = blockIdx.y;
  float* tp = reinterpret_cast<float*>(ptrs[t]);
  float* fp = flat + off[t];
  const int64_t n = len[t];
  const int64_t st = stride ? stride[t] : 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (to_flat) fp[i] = tp[i * st];
    else tp[i * st] = fp[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Contrastive metrics of the pre-training configs (trainer/metrics.py:240-334,444-463, pos_mask == None):
// all five read the same cosine-similarity matrix S = z1 z2^T / (|z1_i| |z2_j|); the reference rebuilds it once per
// metric (five einsums + masks over [B, B]), here one pass over the dot-product matrix produces per-row partials:
//   part[i] = (S_ii, sum_j S_ij, [pred_ii], #{j != i : !pred_ij})     pred_ij = ((S_ij + 1) / 2 > threshold)
// and a deterministic single-block reduction turns them into
//   out[0] positive_similarity = mean_i (S_ii + 1)/2                 out[1] negative_similarity = mean_i ((rowsum_i - S_ii)/(B-1) + 1)/2
//   out[2] true_positive_rate  = #pred_ii / B                        out[3] true_negative_rate  = #(!pred_ij, i != j) / (B (B-1))
//   out[4] contrastive_accuracy = (out[2] + out[3]) / 2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    contrastive_metric_rows_kernel(const float* __restrict__ dot, int64_t B, const float* __restrict__ n1,
                                   const float* __restrict__ n2, float threshold, float* __restrict__ part) {
  pdl_grid_sync();
  __shared__ float sh[32];
  for (int64_t i = blockIdx.x; i < B; i += gridDim.x) {
    const float* row = dot + i * B;
    const float a = n1[i];
    float rowsum = 0.f, diag = 0.f, tn = 0.f, tp = 0.f;
    for (int64_t j = threadIdx.x; j < B; j += blockDim.x) {
      const float s = __fdiv_rn(row[j], __fmul_rn(a, __ldg(n2 + j)));
      const bool pred = __fdiv_rn(__fadd_rn(s, 1.f), 2.f) > threshold;
      rowsum += s;
      if (j == i) {
        diag = s;
        tp = pred ? 1.f : 0.f;
      } else if (!pred) {
        tn += 1.f;
      }
    }
    rowsum = block_sum(rowsum, sh);
    diag = block_sum(diag, sh);
    tn = block_sum(tn, sh);
    tp = block_sum(tp, sh);
    if (threadIdx.x == 0) {
      part[4 * i] = diag, part[4 * i + 1] = rowsum, part[4 * i + 2] = tp, part[4 * i + 3] = tn;
    }
  }
}

__global__ void __launch_bounds__(1024)
    contrastive_metric_final_kernel(const float* __restrict__ part, int64_t B, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ double shd[32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const double d = part[4 * i], rs = part[4 * i + 1];
    acc[0] += (d + 1.0) * 0.5;
    acc[1] += B > 1 ? ((rs - d) / (double)(B - 1) + 1.0) * 0.5 : 0.0;
    acc[2] += part[4 * i + 2];
    acc[3] += part[4 * i + 3];
  }
  double tot[4];
  for (int k = 0; k < 4; ++k) {
    double v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = v;
    __syncthreads();
    v = 0.0;
    if (threadIdx.x < 32) {
      v = threadIdx.x < (blockDim.x >> 5) ? shd[threadIdx.x] : 0.0;
      v = warp_sum(v);
    }
    if (threadIdx.x == 0) tot[k] = v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double b = (double)B;
    out[0] = (float)(tot[0] / b);
    out[1] = (float)(tot[1] / b);
    out[2] = (float)(tot[2] / b);
    out[3] = B > 1 ? (float)(tot[3] / (b * (b - 1.0))) : 0.f;
    out[4] = 0.5f * (out[2] + out[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// The other four metrics the pre-training configs log (configs_clean/pre-train_QM9.yml): DimensionCovariance,
// BatchVariance, Alignment, Uniformity (trainer/metrics.py:161-176,212-230; cov_loss / uniformity_loss of
// commons/losses.py:946-964).  The reference builds a [D,D] covariance and a B(B-1)/2 pdist vector per metric and
// tensor; here tiles of those are reduced on the fly into a handful of fp64 accumulators:
//   acc[0],acc[1]  sum of squared off-diagonal covariance entries of x1, x2      (x - mean)^T (x - mean) / (B - 1)
//   acc[2],acc[3]  sum over columns of the unbiased standard deviation of x1, x2
//   acc[4]         sum_i ||x1_i - x2_i||^alpha
//   acc[5],acc[6]  sum_{i<j} exp(-t ||x_i - x_j||^2)   (exp evaluated in fp32 like the reference: it underflows to 0,
//                  and the metric to -inf, for unnormalised embeddings)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_add_f64(double v, double* __restrict__ target) {
  __shared__ double sh[32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(target, t);
  }
  __syncthreads();
}

// one thread per column: mean (kept for the covariance kernel) and unbiased std, two passes over the B rows
__global__ void metric_cols_kernel(const float* __restrict__ x, int64_t B, int D, double* __restrict__ mean,
                                   double* __restrict__ acc_std) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double sd = 0.0;
  if (c < D) {
    double s = 0.0;
    for (int64_t b = 0; b < B; ++b) s += (double)x[b * D + c];
    const double m = s / (double)B;
    double q = 0.0;
    for (int64_t b = 0; b < B; ++b) {
      const double d = (double)x[b * D + c] - m;
      q = fma(d, d, q);
    }
    mean[c] = m;
    sd = B > 1 ? sqrt(q / (double)(B - 1)) : 0.0;
  }
  block_add_f64(sd, acc_std);
}

// 32 x 32 tile of the covariance matrix per CTA (256 threads, 4 entries each), rows streamed through shared memory
__global__ void __launch_bounds__(256)
    metric_cov_kernel(const float* __restrict__ x, int64_t B, int D, const double* __restrict__ mean,
                      double* __restrict__ acc) {
  pdl_grid_sync();
  __shared__ float Xi[32][33], Xj[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float mi = i0 + tx < D ? (float)mean[i0 + tx] : 0.f, mj = j0 + tx < D ? (float)mean[j0 + tx] : 0.f;
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t b0 = 0; b0 < B; b0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t b = b0 + ty * 4 + q;
      Xi[ty * 4 + q][tx] = (b < B && i0 + tx < D) ? x[b * D + i0 + tx] - mi : 0.f;
      Xj[ty * 4 + q][tx] = (b < B && j0 + tx < D) ? x[b * D + j0 + tx] - mj : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < 32; ++r) {
      const float vj = Xj[r][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = fma((double)Xi[r][ty * 4 + q], (double)vj, a[q]);
    }
    __syncthreads();
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int gi = i0 + ty * 4 + q, gj = j0 + tx;
    if (gi < D && gj < D && gi != gj) {
      const double c = a[q] / (double)(B - 1);
      s = fma(c, c, s);
    }
  }
  block_add_f64(s, acc);
}

// 32 x 32 tile of row pairs per CTA (upper triangle): squared distances accumulated in fp32 in dimension order
__global__ void __launch_bounds__(256)
    metric_pair_kernel(const float* __restrict__ x, int64_t B, int D, float t, double* __restrict__ acc) {
  pdl_grid_sync();
  if (blockIdx.y < blockIdx.x) return;                        // i-tile <= j-tile only (whole CTA)
  __shared__ float Xi[32][33], Xj[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float d2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < D; c0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty * 4 + q;
      Xi[r][tx] = (i0 + r < B && c0 + tx < D) ? x[(i0 + r) * D + c0 + tx] : 0.f;
      Xj[r][tx] = (j0 + r < B && c0 + tx < D) ? x[(j0 + r) * D + c0 + tx] : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < 32; ++c) {
      const float vj = Xj[tx][c];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float d = Xi[ty * 4 + q][c] - vj;
        d2[q] = fmaf(d, d, d2[q]);
      }
    }
    __syncthreads();
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t gi = i0 + ty * 4 + q, gj = j0 + tx;
    if (gi < gj && gj < B) s += (double)expf(-t * d2[q]);
  }
  block_add_f64(s, acc);
}

// one warp per row: ||x1_i - x2_i||_2 ^ alpha
__global__ void metric_align_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int64_t B, int D,
                                    float alpha, double* __restrict__ acc) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  double v = 0.0;
  if (row < B) {
    float s = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = x1[row * D + c] - x2[row * D + c];
      s = fmaf(d, d, s);
    }
    s = warp_sum(s);
    if (lane == 0) v = (double)powf(sqrtf(s), alpha);
  }
  block_add_f64(v, acc);
}

__global__ void metric_final_kernel(const double* __restrict__ acc, int64_t B1, int64_t B2, int D,
                                    float* __restrict__ out) {
  pdl_grid_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  out[0] = (float)(acc[0] / D) + (float)(acc[1] / D);                          // DimensionCovariance
  out[1] = (float)(acc[2] / D) + (float)(acc[3] / D);                          // BatchVariance
  out[2] = (float)(acc[4] / (double)B1);                                       // Alignment
  const double p1 = 0.5 * (double)B1 * (double)(B1 - 1), p2 = 0.5 * (double)B2 * (double)(B2 - 1);
  out[3] = (logf((float)(acc[5] / p1)) + logf((float)(acc[6] / p2))) / 2.f;    // Uniformity
}

}  // namespace i3d

using namespace i3d;

extern "C" {

int i3d_row_norms(const float* z, int64_t R, int D, float* norms, void* stream) {
  I3D_REQUIRE(R >= 0 && D > 0 && (R == 0 || (z && norms)), "invalid argument");
  if (R == 0) return I3D_OK;
  launch(row_norms_kernel, grid_for(R * 32, 256), 256, 0, as_stream(stream), z, R, D, norms);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_ntxent_rows_fwd(float* P, int64_t B, int64_t Bc, int C, const float* n1, const float* n2, int norm,
                        float eps, float tau, int64_t row_offset, float* rowstats, float* loss_rows, void* stream) {
  I3D_REQUIRE(B >= 0 && Bc >= 0 && C >= 1 && tau > 0.f && row_offset >= 0 && row_offset + B <= Bc &&
                  (!norm || (n1 && n2)) && (B == 0 || (P && rowstats && loss_rows)),
              "invalid argument");
  if (B == 0) return I3D_OK;
  const int grid = (int)(B < 65535 ? B : 65535);
  launch(ntxent_rows_fwd_kernel, grid, 256, 0, as_stream(stream), P, B, Bc, C, n1, n2, norm, eps, tau, row_offset,
                                                              rowstats, loss_rows);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_sum_scaled(const float* x, int64_t n, float scale, float* out, void* stream) {
  I3D_REQUIRE(n >= 0 && out && (n == 0 || x), "invalid argument");
  launch(sum_scaled_kernel, 1, 1024, 0, as_stream(stream), x, n, scale, out);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_ntxent_rows_bwd(float* P, int64_t B, int64_t Bc, int C, const float* n1, const float* n2, int norm,
                        float eps, float tau, int64_t row_offset, const float* rowstats, const float* gout,
                        float inv_B, float* dn1, float* dn2, void* stream) {
  I3D_REQUIRE(B >= 0 && Bc >= 0 && C >= 1 && tau > 0.f && gout && (!norm || (n1 && n2 && dn1 && dn2)) &&
                  (B == 0 || (P && rowstats)),
              "invalid argument");
  if (B == 0) return I3D_OK;
  const int grid = (int)(B < 65535 ? B : 65535);
  launch(ntxent_rows_bwd_kernel, grid, 256, 0, as_stream(stream), P, B, Bc, C, n1, n2, norm, eps, tau, row_offset,
                                                              rowstats, gout, inv_B, dn1, dn2);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_norm_bwd_accum(const float* z, const float* norms, const float* dn, int64_t R, int D, float* dz,
                       void* stream) {
  I3D_REQUIRE(R >= 0 && D > 0 && (R == 0 || (z && norms && dn && dz)), "invalid argument");
  if (R == 0) return I3D_OK;
  launch(norm_bwd_accum_kernel, grid_for(R * D, 256), 256, 0, as_stream(stream), z, norms, dn, R, D, dz);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                  double eps, double weight_decay, double grad_scale, int64_t step, const double* hyper_dev,
                  const int64_t* step_dev, void* stream) {
  I3D_REQUIRE(n >= 0 && (step_dev || step >= 1) && (n == 0 || (p && g && m && v)), "invalid argument");
  if (n == 0) return I3D_OK;
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, grad_scale};
  launch(adam_kernel, grid_for(n, 256), 256, 0, as_stream(stream), p, g, m, v, n, h, step, hyper_dev, step_dev);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_adam_step_nvls(float* mc_base, float* local_base, int64_t n, int rank, int world, double lr, double beta1,
                       double beta2, double eps, double weight_decay, double grad_scale, int64_t step,
                       const double* hyper_dev, const int64_t* step_dev, void* stream) {
  I3D_REQUIRE(n >= 0 && (n & 3) == 0 && world >= 1 && rank >= 0 && rank < world && (step_dev || step >= 1) &&
                  (n == 0 || (mc_base && local_base)) && (reinterpret_cast<uintptr_t>(mc_base) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(local_base) & 15u) == 0,
              "invalid argument");
  if (n == 0) return I3D_OK;
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, grad_scale};
  const int64_t per = ((n >> 2) + world - 1) / world;
  launch(adam_nvls_kernel, grid_for(per, 256, 4), 256, 0, as_stream(stream), mc_base, local_base, n, rank, world, h, step,
         hyper_dev, step_dev);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_add_i64(int64_t* x, int64_t delta, void* stream) {
  I3D_REQUIRE(x != nullptr, "invalid argument");
  launch(add_i64_kernel, 1, 1, 0, as_stream(stream), x, delta);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_multi_copy(const uint64_t* ptrs, const int64_t* off, const int64_t* len, int T, float* flat, int to_flat,
                   void* stream) {
  return i3d_multi_copy_strided(ptrs, off, len, nullptr, T, flat, to_flat, stream);
}

int i3d_multi_copy_strided(const uint64_t* ptrs, const int64_t* off, const int64_t* len, const int64_t* stride, int T,
                           float* flat, int to_flat, void* stream) {
  I3D_REQUIRE(T >= 0 && T <= 65535 && (T == 0 || (ptrs && off && len && flat)), "invalid argument");
  if (T == 0) return I3D_OK;
  launch(multi_copy_kernel, dim3(16, T, 1), 256, 0, as_stream(stream), ptrs, off, len, stride, flat, to_flat);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_embedding_metrics(const float* x1, int64_t B1, const float* x2, int64_t B2, int D, float alpha, float t,
                          double* ws, float* out4, void* stream) {
  I3D_REQUIRE(x1 && x2 && ws && out4 && B1 >= 2 && B2 >= B1 && D >= 2 && D <= 65536, "invalid argument");
  cudaStream_t s = as_stream(stream);
  double* acc = ws;
  double* mean1 = ws + 8;
  double* mean2 = ws + 8 + D;
  I3D_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 8, s));
  const int tiles_d = (D + 31) / 32;
  launch(metric_cols_kernel, (D + 127) / 128, 128, 0, s, x1, B1, D, mean1, acc + 2);
  I3D_LAUNCHED();
  launch(metric_cols_kernel, (D + 127) / 128, 128, 0, s, x2, B2, D, mean2, acc + 3);
  I3D_LAUNCHED();
  launch(metric_cov_kernel, dim3(tiles_d, tiles_d), 256, 0, s, x1, B1, D, mean1, acc + 0);
  I3D_LAUNCHED();
  launch(metric_cov_kernel, dim3(tiles_d, tiles_d), 256, 0, s, x2, B2, D, mean2, acc + 1);
  I3D_LAUNCHED();
  const int t1 = (int)((B1 + 31) / 32), t2 = (int)((B2 + 31) / 32);
  I3D_REQUIRE(t2 <= 65535, "too many rows");
  launch(metric_pair_kernel, dim3(t1, t1), 256, 0, s, x1, B1, D, t, acc + 5);
  I3D_LAUNCHED();
  launch(metric_pair_kernel, dim3(t2, t2), 256, 0, s, x2, B2, D, t, acc + 6);
  I3D_LAUNCHED();
  launch(metric_align_kernel, (int)((B1 * 32 + 255) / 256), 256, 0, s, x1, x2, B1, D, alpha, acc + 4);
  I3D_LAUNCHED();
  launch(metric_final_kernel, 1, 32, 0, s, acc, B1, B2, D, out4);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_contrastive_metrics(const float* dot, int64_t B, const float* n1, const float* n2, float threshold,
                            float* part, float* out5, void* stream) {
  I3D_REQUIRE(B >= 1 && dot && n1 && n2 && part && out5, "invalid argument");
  const int grid = (int)(B < 65535 ? B : 65535);
  launch(contrastive_metric_rows_kernel, grid, 256, 0, as_stream(stream), dot, B, n1, n2, threshold, part);
  I3D_LAUNCHED();
  launch(contrastive_metric_final_kernel, 1, 1024, 0, as_stream(stream), part, B, out5);
  I3D_LAUNCHED();
  return I3D_OK;
}
}
