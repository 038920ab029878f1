// Tensor-core path of the segmented virtual-operand GEMM: tcgen05.mma (kind::tf32) with the accumulator in TMEM.
//
//   NT:  C[m, n] = bias[n] + sum_s sum_k  scale_s[m] * A_s[a_idx_s[m], k] * B_s[n, k]        (y = x W^T, dx = dy (W^T)^T)
//   TN:  C[m, n] (+)= sum_k  scale[k] * A[a_idx[k], m] * B[b_idx[k], n]                      (dW = dy^T x, split over K)
//
// Why a hand-written kernel instead of cuBLAS: the A operand is *virtual* — rows are gathered through the CSR
// edge lists (h[src], h[dst]) and scaled by the per-node degree scalers while they are staged, so neither
// torch.cat nor index_select ever touches HBM (models/pna.py:207,232,249).  TMA cannot express that staging, so
// A tiles are produced by the CTA's own threads (LDG -> hi/lo split -> STS in the canonical K-major SWIZZLE_128B
// layout, two k-blocks of register prefetch), published to the async proxy with fence.proxy.async, and consumed by
// tcgen05.mma issued by one thread.  The dense B operand (weights) is split into hi/lo ONCE per call by a small
// kernel and then streamed by TMA (cp.async.bulk.tensor, SWIZZLE_128B) straight into shared memory.
//
// Precision: the reference computes in fp32 (SGEMM).  Each fp32 operand is split into hi = tf32(x) and lo = x - hi
// and three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32 in TMEM ("3xTF32").  Measured against fp64
// (tests/gpu_cases.py::case_gemm_tc): 5e-7 .. 4e-6 of max|C| for K <= 600, 2e-5 at K = 2600 (the TMEM accumulator
// truncates, so the error grows with the number of accumulated MMAs); the fp32 SIMT backend gives 3e-7 .. 2e-6.
//
// Tile: 128 (M) x BN (N, multiple of 16, <= 256) x 32 (K = one 128-byte swizzle atom of tf32), 2 smem stages,
// 256 threads: all threads stage A, thread 0 issues TMA + MMAs, 8 warps drain TMEM through shared memory.
#include <cooperative_groups.h>
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "i3d_tc.cuh"

namespace i3d {

struct TcParams {
  i3d_gemm_seg seg[4];
  int n_seg;
  int64_t M;
  int N;
  float* C;
  int ldc;
  const float* bias;
  int accumulate;
  int kchunk;   // TN: K range per CTA (multiple of TC_BK)
  int splits;   // TN: gridDim.z
  double* stats;    // optional fused BatchNorm statistics [2N] (NT only)
  int stats_act;
  // TN over a degree plan (i3d_gemm_tn_chunked): CTA z reduces rows [tab[3z], +tab[3z+1]) of the virtual row order
  // into the output block of bucket tab[3z+2] (C + bucket * c_bucket_stride); always atomic, caller pre-zeroes C
  const int32_t* chunk_tab;
  int64_t c_bucket_stride;
  const int32_t* m_valid;   // optional device scalar: rows >= *m_valid are excluded from `stats`
  int cluster;              // TN: CTAs per thread-block cluster along the split-K (z) dimension, 1 = no cluster
};

// Split-K reduction inside a thread-block cluster (TN).  The `splits` partial tiles of one output block used to be
// reduced by float4 atomics from every CTA: ~74 CTAs finishing together and adding onto the same 128-byte sectors,
// which L2 serialises (the epilogue cost as much as the MMA loop).  With a cluster of CL CTAs along z every CTA drains
// its TMEM accumulator into its own shared memory, the cluster synchronises, and CTA r sums rows [r*128/CL, +128/CL) of
// all CL partial tiles through distributed shared memory (fixed order: deterministic per cluster) and issues the
// atomics for that slice only: CL times fewer atomics per output element.
template <int BN>
__device__ __forceinline__ void tn_cluster_epilogue(uint32_t tmem, float* ctile, bool has_acc, int64_t M, int N,
                                                    int64_t m0, int n0, float* __restrict__ C, int ldc,
                                                    const float* __restrict__ bias, int accumulate, bool atomic) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
  constexpr int LDT = BN + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  constexpr int HALF_COLS = BN / 2;
  const int c_begin = half * HALF_COLS;
  float* trow = ctile + (q * 32 + lane) * LDT;
  for (int c = c_begin; c < c_begin + HALF_COLS; c += 16) {
    float v[16];
    if (has_acc) {
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    const int lim = min(16, c_begin + HALF_COLS - c);
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      if (i < lim) *reinterpret_cast<float4*>(trow + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  cluster.sync();                                   // every partial tile of the cluster is in shared memory
  const float* peer[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) peer[r] = cluster.map_shared_rank(ctile, r < CL ? r : 0);
  const int rows_per = TC_BM / CL, rbeg = cr * rows_per;
  const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) && ((N & 3) == 0);
  constexpr int QUADS = BN / 4;
  for (int idx = tid; idx < rows_per * QUADS; idx += TC_THREADS) {
    const int rl = idx / QUADS, c = (idx - rl * QUADS) * 4;
    const int r = rbeg + rl;
    const int64_t gm = m0 + r;
    const int gn = n0 + c;
    if (gm >= M || gn >= N) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int pr = 0; pr < 8; ++pr) {
      if (pr < CL) {
        const float4 w = *reinterpret_cast<const float4*>(peer[pr] + r * LDT + c);
        v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
      }
    }
    float* out = C + gm * ldc + gn;
    if (vec) {
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + gn));
        v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
      }
      if (atomic) {
        atomicAdd(reinterpret_cast<float4*>(out), v);
      } else {
        if (accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(out);
          v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
        }
        *reinterpret_cast<float4*>(out) = v;
      }
    } else {
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (gn + i < N) {
          const float o = e[i] + (bias ? __ldg(bias + gn + i) : 0.f);
          if (atomic) atomicAdd(out + i, o);
          else out[i] = accumulate ? out[i] + o : o;
        }
      }
    }
  }
  cluster.sync();                                   // peers are done reading this CTA's tile
}

// =================================================================================================================
// Generic kernel: both operands staged by threads.  NT without a workspace, and TN (transposes while staging).
// =================================================================================================================
template <int MODE, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  pdl_grid_sync();
  using L = TcLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  float* tiles = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)TC_STAGES * L::STAGE);   // [STAGES] mma_done, [1] acc_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * BN;
  const int64_t M = p.M;
  const int N = p.N;
  float* Cout = p.C;
  int chunk_k0 = 0, chunk_rows = 0;
  if (MODE == I3D_GEMM_TN && p.chunk_tab) {
    chunk_k0 = __ldg(p.chunk_tab + 3 * blockIdx.z);
    chunk_rows = __ldg(p.chunk_tab + 3 * blockIdx.z + 1);
    if (chunk_rows <= 0) return;        // unused chunk of the degree plan (whole CTA, before any allocation)
    Cout += (int64_t)__ldg(p.chunk_tab + 3 * blockIdx.z + 2) * p.c_bucket_stride;
  }

  if (warp == 0) tmem_alloc(tmem_slot, L::TMEM_COLS);
  if (tid == 32) {
    for (int s = 0; s <= TC_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);

  // ---- k-block enumeration: (segment, k0) cursor used by the prefetcher -------------------------------------------
  int total = 0;
  int kbeg = 0, kend = 0;
  if (MODE == I3D_GEMM_NT) {
    for (int s = 0; s < p.n_seg; ++s) total += (p.seg[s].K + TC_BK - 1) / TC_BK;
  } else {
    kbeg = p.chunk_tab ? chunk_k0 : blockIdx.z * p.kchunk;
    kend = min(p.seg[0].K, kbeg + (p.chunk_tab ? chunk_rows : p.kchunk));
    total = kend > kbeg ? (kend - kbeg + TC_BK - 1) / TC_BK : 0;
  }
  int cur_seg = 0, cur_k0 = (MODE == I3D_GEMM_NT) ? 0 : kbeg;

  // NT: per-thread fixed A rows for the current segment (4 chunks per thread: row = i*32 + tid>>3, chunk j = tid&7)
  int64_t a_row[4];
  float a_sc[4];
  auto bind_segment = [&](int s) {
    if (MODE != I3D_GEMM_NT) return;
    const int32_t* __restrict__ a_idx = p.seg[s].a_idx;
    const float* __restrict__ scale = p.seg[s].scale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t gm = m0 + i * 32 + (tid >> 3);
      a_row[i] = -1;
      a_sc[i] = 1.f;
      if (gm < M) {
        a_row[i] = a_idx ? (int64_t)__ldg(a_idx + gm) : gm;
        if (scale) a_sc[i] = __ldg(scale + gm);
      }
    }
  };
  bind_segment(0);

  float4 va[4];
  float va_sc[4];              // scale of each staged A chunk, applied at store time (keeps the loads a prefetch)
  float4 vb[L::B_CHUNKS];

  // TN: the one k-row this thread stages in the k-block starting at k0 (indices resolved one block ahead)
  int64_t tn_arow = -1, tn_brow = 0;
  float tn_sc = 1.f;
  auto tn_bind = [&](int k0) {
    if (MODE != I3D_GEMM_TN) return;
    const int k = k0 + (warp & 3) * 8 + (lane >> 2);
    tn_arow = -1;
    tn_sc = 1.f;
    if (k < kend) {
      tn_arow = p.seg[0].a_idx ? (int64_t)__ldg(p.seg[0].a_idx + k) : (int64_t)k;
      tn_brow = p.seg[0].b_idx ? (int64_t)__ldg(p.seg[0].b_idx + k) : (int64_t)k;
      if (p.seg[0].scale) tn_sc = __ldg(p.seg[0].scale + k);
    }
  };
  tn_bind(cur_k0);

  // issue the global loads of the k-block under the cursor into registers (zero-filled outside the matrices)
  auto prefetch = [&]() {
    if (MODE == I3D_GEMM_NT) {
      const float* __restrict__ A = p.seg[cur_seg].A;
      const float* __restrict__ B = p.seg[cur_seg].B;
      const int lda = p.seg[cur_seg].lda, ldb = p.seg[cur_seg].ldb, K = p.seg[cur_seg].K;
      const int kc = cur_k0 + (tid & 7) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_row[i] >= 0 && kc < K) v = __ldg(reinterpret_cast<const float4*>(A + a_row[i] * lda + kc));
        va[i] = v;
        va_sc[i] = a_sc[i];
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int r = (tid + i * TC_THREADS) >> 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < BN && n0 + r < N && kc < K) v = __ldg(reinterpret_cast<const float4*>(B + (int64_t)(n0 + r) * ldb + kc));
        vb[i] = v;
      }
      cur_k0 += TC_BK;
      if (cur_k0 >= K && cur_seg + 1 < p.n_seg) {
        cur_seg += 1;
        cur_k0 = 0;
        bind_segment(cur_seg);
      }
    } else {
      // TN: tiles are transposed while staging.  One warp pass = 8 k-rows x 16 columns (float4 per lane); every
      // pass of a thread reads the SAME k-row (k = k0 + (warp&3)*8 + lane>>2), whose gather indices / scale were
      // loaded one k-block earlier (tn_arow / tn_brow / tn_sc) so that no data load waits on an index load.
      const float* __restrict__ A = p.seg[0].A;
      const float* __restrict__ B = p.seg[0].B;
      const int lda = p.seg[0].lda, ldb = p.seg[0].ldb;
      const int cg = (lane & 3) * 4;
      const bool kvalid = tn_arow >= 0 && tn_brow >= 0;     // negative gather index (padding row) = zero row
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ((warp + 8 * i) >> 2) * 16 + cg;       // 8 column blocks of 16
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvalid && m < M) v = __ldg(reinterpret_cast<const float4*>(A + tn_arow * lda + m));
        va[i] = v;
        va_sc[i] = tn_sc;
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int c = ((warp + 8 * i) >> 2) * 16 + cg;                // BN/16 column blocks
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvalid && c < BN && n0 + c < N) v = __ldg(reinterpret_cast<const float4*>(B + tn_brow * ldb + n0 + c));
        vb[i] = v;
      }
      cur_k0 += TC_BK;
      tn_bind(cur_k0);
    }
  };

  auto stage_store = [&](float* a_hi, float* a_lo, float* b_hi, float* b_lo) {
    if (MODE == I3D_GEMM_NT) {
      const int j = tid & 7;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = va[i];
        const float sc = va_sc[i];
        v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
        split_store4(a_hi, a_lo, sw128_off(i * 32 + (tid >> 3), j), v);
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int r = (tid + i * TC_THREADS) >> 3;
        if (r < BN) split_store4(b_hi, b_lo, sw128_off(r, j), vb[i]);
      }
    } else {
      const int kl = lane >> 2, cg = (lane & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pid = warp + 8 * i;
        const int kk = (pid & 3) * 8 + kl;               // k position inside the tile (0..31)
        const int r = (pid >> 2) * 16 + cg;              // tile row (= output row m)
        const float sc = va_sc[i];
        const float e[4] = {va[i].x * sc, va[i].y * sc, va[i].z * sc, va[i].w * sc};
#pragma unroll
        for (int q = 0; q < 4; ++q) split_store1(a_hi, a_lo, sw128_off(r + q, kk >> 2) + (kk & 3), e[q]);
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int pid = warp + 8 * i;
        const int kk = (pid & 3) * 8 + kl;
        const int r = (pid >> 2) * 16 + cg;
        if (r < BN) {
          const float e[4] = {vb[i].x, vb[i].y, vb[i].z, vb[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) split_store1(b_hi, b_lo, sw128_off(r + q, kk >> 2) + (kk & 3), e[q]);
        }
      }
    }
  };

  // (a second register set — k-blocks it+1 and it+2 in flight — was measured: no gain, 20.4 -> 21.1 us for the
  //  200 x 200 x E weight gradient; per k-block the 12 3xTF32 MMAs are ~0.7-0.9 us of the ~1.6 us, tools/tn_sweep.py)
  if (total > 0) prefetch();
  for (int it = 0; it < total; ++it) {
    const int st = it % TC_STAGES;
    const int use = it / TC_STAGES;
    float* a_hi = tiles + (size_t)st * L::STAGE;
    float* a_lo = a_hi + L::A_TILE;
    float* b_hi = a_lo + L::A_TILE;
    float* b_lo = b_hi + L::B_TILE;
    // the MMAs that last read this stage must have completed before it is overwritten
    if (use > 0) mbar_wait(&bars[st], (uint32_t)((use - 1) & 1));
    tc_fence_after();
    stage_store(a_hi, a_lo, b_hi, b_lo);
    fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
    if (it + 1 < total) prefetch();   // next k-block's global loads fly across the barrier and the MMA issue
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_kblock(tmem, a_hi, a_lo, b_hi, b_lo, idesc, it == 0);
      umma_commit(&bars[st]);    // frees this stage when the MMAs above are done (implies fence::before_thread_sync)
    }
  }
  if (total > 0) {
    if (tid == 0) umma_commit(&bars[TC_STAGES]);      // accumulator complete
    mbar_wait(&bars[TC_STAGES], 0);
    tc_fence_after();
  }
  const bool atomic = (MODE == I3D_GEMM_TN) && p.splits > 1;
  bool clustered = false;
  if constexpr (MODE == I3D_GEMM_TN) clustered = p.cluster > 1;
  if (clustered) {
    // (atomic unless this cluster is the only one working on the output block)
    tn_cluster_epilogue<BN>(tmem, tiles, total > 0, M, N, m0, n0, Cout, p.ldc,
                            (p.bias && blockIdx.z < (unsigned)p.cluster) ? p.bias : nullptr, p.accumulate,
                            p.splits > p.cluster);
  } else {
    tc_epilogue<BN>(tmem, tiles, total > 0, M, N, m0, n0, Cout, p.ldc,
                    (p.bias && !(atomic && blockIdx.z != 0)) ? p.bias : nullptr, p.accumulate, atomic,
                    MODE == I3D_GEMM_NT ? p.stats : nullptr, p.stats_act, nullptr, 0xffffffffu, p.m_valid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}
__global__ void tc_zero_block_kernel(float* __restrict__ C, int64_t M, int N, int ldc) {
  pdl_grid_sync();
  const int64_t total = M * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / N;
    C[m * ldc + (t - m * N)] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------- host side
template <typename Kern>
static int set_smem(Kern kern, size_t bytes, bool* configured) {
  if (*configured) return I3D_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("i3d_gemm(tc): cudaFuncSetAttribute -> %s", cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  *configured = true;
  return I3D_OK;
}

static int launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("i3d_gemm(tc): %s launch failed -> %s", what, cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  count_launch();
  return I3D_OK;
}

// cluster launch (TN split-K reduction through distributed shared memory): CL CTAs along z per cluster
template <typename... KArgs, typename... Args>
static inline void launch_cluster_z(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    int cl, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = 1, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = (unsigned)cl;
  ++n;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// CTAs of `kernel` (one per SM: launch bounds 1, ~170 KB of shared memory) that can be co-resident when they are
// launched as clusters of `cl`: clusters live inside one GPC, so a few SMs per GPC stay empty (0 = query failed)
template <typename Kern>
static int cluster_capacity(Kern kernel, size_t smem, int cl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(1, 1, (unsigned)cl), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = (unsigned)cl;
  cfg.attrs = attr, cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n * cl;
}

// I3D_TN_CLUSTER=c: largest cluster size tried for the split-K reduction of the weight-gradient GEMMs.  Default 1 (float4
// atomics from every CTA): measured on B200 at batch 512 (tools/gemm_bench.py, us per launch, cluster 1 / 2 / 4 / 8):
// dW 200x200 over K = E: 20.4 / 23.1 / 24.2 / 24.2; over K = N: 12.9 / 16.9 / 16.7 / 16.7; 200x800 over K = N:
// 30.5 / 34.8 / 34.8 / 34.8; whole step 4.20 / - / 4.27 / 4.30 ms.  The atomics are not what bounds this kernel.
static int tn_cluster_max() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("I3D_TN_CLUSTER");
    v = e ? atoi(e) : 1;
    if (v != 1 && v != 2 && v != 4 && v != 8) v = 1;
  }
  return v;
}

template <int MODE, int BN>
static int launch_generic(TcParams& p, cudaStream_t s) {
  using L = TcLayout<BN>;
  static bool configured = false;
  if (int rc = set_smem(gemm_tc_kernel<MODE, BN>, L::BYTES, &configured)) return rc;
  const int64_t gx = (p.M + TC_BM - 1) / TC_BM;
  const int gy = (p.N + BN - 1) / BN;
  int gz = 1;
  p.cluster = 1;
  if (MODE == I3D_GEMM_TN) {
    const int K = p.seg[0].K;
    const int64_t tiles = gx * gy;
    const int64_t max_splits = (K + 4 * TC_BK - 1) / (4 * TC_BK);   // at least 4 k-blocks per CTA
    int64_t want = 0;
    // largest cluster size that keeps (>= 85 % of) the split-K parallelism the machine offers without clusters
    static int cap[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
    int64_t plain = (sm_count() + tiles - 1) / tiles;
    if (plain > max_splits) plain = max_splits;
    for (int cl = tn_cluster_max(); cl > 1; cl >>= 1) {
      if (cap[cl] < 0) cap[cl] = cluster_capacity(gemm_tc_kernel<MODE, BN>, L::BYTES, cl);
      int64_t sp = (cap[cl] / tiles / cl) * cl;
      if (sp > max_splits) sp = (max_splits / cl) * cl;
      if (sp >= cl && sp * 100 >= plain * 85) {
        want = sp, p.cluster = cl;
        break;
      }
    }
    if (p.cluster == 1) {
      want = (sm_count() + tiles - 1) / tiles;          // ~one CTA per SM
      {
        // I3D_TN_SPLIT_DIV=d: 1/d of that (fewer, longer CTAs: less atomic traffic, SMs left to the main stream)
        static int div = -1;
        if (div < 0) {
          const char* e = getenv("I3D_TN_SPLIT_DIV");
          div = e ? atoi(e) : 1;
          if (div < 1) div = 1;
        }
        want = (want + div - 1) / div;
      }
      if (want > max_splits) want = max_splits;
      if (want < 1) want = 1;
    }
    int kchunk = (int)((K + want - 1) / want);
    kchunk = ((kchunk + TC_BK - 1) / TC_BK) * TC_BK;
    p.kchunk = kchunk;
    // (with a cluster the grid keeps `want` CTAs along z — a multiple of the cluster size — even when rounding the
    //  K range per CTA up to whole k-blocks leaves the last ones without work: they contribute zero tiles)
    p.splits = p.cluster > 1 ? (int)want : (K + kchunk - 1) / kchunk;
    gz = p.splits;
    if (p.splits > p.cluster && !p.accumulate) {
      launch(tc_zero_block_kernel, grid_for(p.M * p.N, 256), 256, 0, s, p.C, p.M, p.N, p.ldc);
      if (int rc = launched("zero")) return rc;
    }
    if (p.cluster > 1) {
      launch_cluster_z(gemm_tc_kernel<MODE, BN>, dim3((unsigned)gx, gy, gz), TC_THREADS, L::BYTES, s, p.cluster, p);
      return launched("gemm(tn, cluster)");
    }
  }
  launch(gemm_tc_kernel<MODE, BN>, dim3((unsigned)gx, gy, gz), TC_THREADS, L::BYTES, s, p);
  return launched("gemm");
}


bool gemm_ws_available();
static inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// problems the tensor-core kernels take: 16-byte aligned operands, float4-divisible extents (float4 staging)
bool gemm_tc_eligible(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs) {
  if (mode == I3D_GEMM_NT) {
    if (M < 256 || N < 16) return false;
    for (int s = 0; s < n_seg; ++s) {
      if (segs[s].K <= 0 || (segs[s].K & 3) || (segs[s].lda & 3) || (segs[s].ldb & 3) || !al16(segs[s].A) ||
          !al16(segs[s].B) || segs[s].b_idx)
        return false;
    }
    return true;
  }
  if (mode == I3D_GEMM_TN) {
    if (n_seg != 1 || segs[0].K < 512 || M < 16 || N < 16 || (M & 3) || (N & 3)) return false;
    return !(segs[0].lda & 3) && !(segs[0].ldb & 3) && al16(segs[0].A) && al16(segs[0].B);
  }
  return false;
}

// prepared B operands: only the A side, the shape and the TMA entry point decide
bool gemm_nt_prepared_ok(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs) {
  if (M < 256 || N < 16 || n_seg < 1 || n_seg > 4 || !gemm_ws_available()) return false;
  for (int s = 0; s < n_seg; ++s)
    if (segs[s].K <= 0 || (segs[s].K & 3) || (segs[s].lda & 3) || !al16(segs[s].A) || segs[s].b_idx) return false;
  return true;
}

int gemm_tn_ws(int64_t M, int N, const i3d_gemm_seg& sg, float* C, int ldc, int accumulate, cudaStream_t stream);
// true: MN-major (SWIZZLE_128B_BASE32B) warp-specialised TN kernel, i3d_gemm_backend(2).  Measured slower than the
// transposing generic kernel on the dW shapes of this path, so it is opt-in.
bool g_tn_ws = false;
size_t gemm_ws_bytes(int N, int n_seg, const i3d_gemm_seg* segs);
int gemm_ws_nt(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
               int accumulate, void* ws, double* stats, int stats_act, cudaStream_t stream, bool prepared,
               const int32_t* m_valid);

// bytes of scratch that let the NT kernel stream the B operand by TMA (hi + lo copies, K padded per segment)
size_t gemm_tc_ws_bytes(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs) {
  if (mode != I3D_GEMM_NT || !gemm_tc_eligible(mode, M, N, n_seg, segs)) return 0;
  return gemm_ws_bytes(N, n_seg, segs);
}

// pick the N tile: 208 covers the F=200 outputs of the PNA layers with one accumulator
template <typename F>
static int with_bn(int mode, int64_t M, int N, F&& f) {
  const int64_t gx = (M + TC_BM - 1) / TC_BM;
  if (N <= 32) return f(std::integral_constant<int, 32>());
  if (N <= 64) return f(std::integral_constant<int, 64>());
  if (N <= 112) return f(std::integral_constant<int, 112>());
  if (N <= 128) return f(std::integral_constant<int, 128>());
  // few row tiles (node-level GEMMs at batch 512: 72 tiles on 148 SMs): split N over two CTAs to fill the machine
  if (mode == I3D_GEMM_NT && N <= 208 && gx * 2 <= sm_count()) return f(std::integral_constant<int, 112>());
  if (N <= 208) return f(std::integral_constant<int, 208>());
  if ((N + 207) / 208 <= (N + 255) / 256) return f(std::integral_constant<int, 208>());   // same tiles, less padding
  return f(std::integral_constant<int, 256>());
}

// dW of a degree-bucketed GEMM: C[bucket] += A[a_idx[k], :]^T B[b_idx[k], :] over the chunks of the degree plan
int gemm_tn_chunked(int64_t M, int N, const i3d_gemm_seg& sg, float* C, int ldc, int64_t c_bucket_stride,
                    const int32_t* chunk_tab, int n_chunks, cudaStream_t stream) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.seg[0] = sg;
  p.n_seg = 1, p.M = M, p.N = N, p.C = C, p.ldc = ldc, p.bias = nullptr, p.accumulate = 1;
  p.kchunk = 0, p.splits = 2 /* atomic epilogue */, p.chunk_tab = chunk_tab, p.c_bucket_stride = c_bucket_stride;
  return with_bn(I3D_GEMM_TN, M, N, [&](auto bn) {
    constexpr int BN = decltype(bn)::value;
    using L = TcLayout<BN>;
    static bool configured = false;
    if (int rc = set_smem(gemm_tc_kernel<I3D_GEMM_TN, BN>, L::BYTES, &configured)) return rc;
    const int64_t gx = (M + TC_BM - 1) / TC_BM;
    const int gy = (N + BN - 1) / BN;
    launch(gemm_tc_kernel<I3D_GEMM_TN, BN>, dim3((unsigned)gx, gy, n_chunks), TC_THREADS, L::BYTES, stream, p);
    return launched("gemm(tn chunked)");
  });
}

int gemm_tc(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
            int accumulate, void* ws, size_t ws_bytes, double* stats, int stats_act, cudaStream_t stream,
            const int32_t* m_valid) {
  const size_t need = gemm_tc_ws_bytes(mode, M, N, n_seg, segs);
  if (mode == I3D_GEMM_NT && ws && need > 0 && ws_bytes >= need && gemm_ws_available())
    return gemm_ws_nt(M, N, n_seg, segs, C, ldc, bias, accumulate, ws, stats, stats_act, stream, false, m_valid);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.m_valid = m_valid;
  for (int s = 0; s < n_seg; ++s) p.seg[s] = segs[s];
  p.n_seg = n_seg, p.M = M, p.N = N, p.C = C, p.ldc = ldc, p.bias = bias, p.accumulate = accumulate;
  p.kchunk = 0, p.splits = 1, p.stats = stats, p.stats_act = stats_act;
  if (mode == I3D_GEMM_NT)
    return with_bn(mode, M, N, [&](auto bn) { return launch_generic<I3D_GEMM_NT, decltype(bn)::value>(p, stream); });
  if (g_tn_ws) return gemm_tn_ws(M, N, segs[0], C, ldc, accumulate, stream);
  return with_bn(mode, M, N, [&](auto bn) { return launch_generic<I3D_GEMM_TN, decltype(bn)::value>(p, stream); });
}

}  // namespace i3d
