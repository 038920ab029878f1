// Tensor-core path of the segmented virtual-operand GEMM: tcgen05.mma (kind::tf32) with the accumulator in TMEM.
//
//   NT:  C[m, n] = bias[n] + sum_s sum_k  scale_s[m] * A_s[a_idx_s[m], k] * B_s[n, k]        (y = x W^T, dx = dy (W^T)^T)
//   TN:  C[m, n] (+)= sum_k  scale[k] * A[a_idx[k], m] * B[b_idx[k], n]                      (dW = dy^T x, split over K)
//
// Why a hand-written kernel instead of cuBLAS: the operands are *virtual* — rows are gathered through the CSR
// edge lists (h[src], h[dst]) and scaled by the per-node degree scalers while they are staged, so neither
// torch.cat nor index_select ever touches HBM (models/pna.py:207,232,249).  TMA cannot express that staging, so
// operand tiles are produced by the CTA's own threads (LDG -> split -> STS in the canonical K-major SWIZZLE_128B
// layout; TN transposes while staging), published to the async proxy with fence.proxy.async, and consumed by
// tcgen05.mma issued by one thread.  The next k-block's global loads are issued before the barrier so that their
// latency overlaps the MMA issue and the wait for the stage to drain.
//
// Precision: the reference computes in fp32 (SGEMM).  Each fp32 operand is split into hi = tf32(x) and lo = x - hi
// and three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32 in TMEM ("3xTF32").  Measured against fp64
// (tests/gpu_cases.py::case_gemm_tc): 5e-7 .. 4e-6 of max|C| for K <= 600, 2e-5 at K = 2600 (the TMEM accumulator
// truncates, so the error grows with the number of accumulated MMAs); the fp32 SIMT backend gives 3e-7 .. 2e-6.
//
// Tile: 128 (M) x BN (N, multiple of 16, <= 256) x 32 (K = one 128-byte swizzle atom of tf32), 2 smem stages,
// 256 threads: all threads stage operands, thread 0 issues the MMAs, 8 warps drain TMEM (tcgen05.ld 32x32b).
#include "i3d_common.cuh"

namespace i3d {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 2, TC_THREADS = 256;

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
};

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout, sm_100):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 = 1024 B (8 rows x 128 B)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=b=TF32 [7,10)=[10,13)=2, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// float offset of 16-byte chunk j (0..7) of row r inside a [rows x 32 tf32] K-major SWIZZLE_128B tile
__device__ __forceinline__ int sw128_off(int r, int j) { return (r >> 3) * 256 + (r & 7) * 32 + ((j ^ (r & 7)) << 2); }

__device__ __forceinline__ void split_store4(float* hi_tile, float* lo_tile, int off, float4 v) {
  float4 h, l;
  h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
  l.x = v.x - h.x, l.y = v.y - h.y, l.z = v.z - h.z, l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}
__device__ __forceinline__ void split_store1(float* hi_tile, float* lo_tile, int off, float v) {
  const float h = to_tf32(v);
  hi_tile[off] = h;
  lo_tile[off] = v - h;
}

template <int BN>
struct TcLayout {
  static constexpr int A_TILE = TC_BM * TC_BK;                 // floats
  static constexpr int B_TILE = BN * TC_BK;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;       // a_hi, a_lo, b_hi, b_lo
  static constexpr size_t BYTES = (size_t)TC_STAGES * STAGE * 4 + 1024 /*align slack*/ + 64 /*barriers*/;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int B_CHUNKS = (BN * 8 + TC_THREADS - 1) / TC_THREADS;     // float4 per thread per k-block
};

template <int MODE, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
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
    kbeg = blockIdx.z * p.kchunk;
    kend = min(p.seg[0].K, kbeg + p.kchunk);
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
  float4 vb[L::B_CHUNKS];

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
        if (a_row[i] >= 0 && kc < K) {
          v = __ldg(reinterpret_cast<const float4*>(A + a_row[i] * lda + kc));
          const float sc = a_sc[i];
          v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
        }
        va[i] = v;
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
      // TN: tiles are transposed while staging.  One warp pass = 8 k-rows x 16 columns (float4 per lane).
      const float* __restrict__ A = p.seg[0].A;
      const float* __restrict__ B = p.seg[0].B;
      const int32_t* __restrict__ a_idx = p.seg[0].a_idx;
      const int32_t* __restrict__ b_idx = p.seg[0].b_idx;
      const float* __restrict__ scale = p.seg[0].scale;
      const int lda = p.seg[0].lda, ldb = p.seg[0].ldb;
      const int kl = lane >> 2, cg = (lane & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pid = warp + 8 * i;                    // 32 passes: 8 column blocks x 4 k blocks
        const int k = cur_k0 + (pid & 3) * 8 + kl;
        const int64_t m = m0 + (pid >> 2) * 16 + cg;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend && m < M) {
          const int64_t row = a_idx ? (int64_t)__ldg(a_idx + k) : (int64_t)k;
          v = __ldg(reinterpret_cast<const float4*>(A + row * lda + m));
          if (scale) {
            const float sc = __ldg(scale + k);
            v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
          }
        }
        va[i] = v;
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int pid = warp + 8 * i;                    // BN/16 column blocks x 4 k blocks
        const int k = cur_k0 + (pid & 3) * 8 + kl;
        const int c = (pid >> 2) * 16 + cg;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < BN && k < kend && n0 + c < N) {
          const int64_t row = b_idx ? (int64_t)__ldg(b_idx + k) : (int64_t)k;
          v = __ldg(reinterpret_cast<const float4*>(B + row * ldb + n0 + c));
        }
        vb[i] = v;
      }
      cur_k0 += TC_BK;
    }
  };

  auto stage_store = [&](float* a_hi, float* a_lo, float* b_hi, float* b_lo) {
    if (MODE == I3D_GEMM_NT) {
      const int j = tid & 7;
#pragma unroll
      for (int i = 0; i < 4; ++i) split_store4(a_hi, a_lo, sw128_off(i * 32 + (tid >> 3), j), va[i]);
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
        const float e[4] = {va[i].x, va[i].y, va[i].z, va[i].w};
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
      const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ++ks) {
        const uint32_t koff = ks * 32;                            // 8 tf32 = 32 bytes along K inside the swizzle atom
        const uint64_t dah = make_smem_desc(sa_hi + koff), dal = make_smem_desc(sa_lo + koff);
        const uint64_t dbh = make_smem_desc(sb_hi + koff), dbl = make_smem_desc(sb_lo + koff);
        umma_tf32(tmem, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        umma_tf32(tmem, dal, dbh, idesc, 1u);
        umma_tf32(tmem, dah, dbl, idesc, 1u);
      }
      umma_commit(&bars[st]);    // frees this stage when the MMAs above are done (implies fence::before_thread_sync)
    }
  }
  if (total > 0) {
    if (tid == 0) umma_commit(&bars[TC_STAGES]);      // accumulator complete
    mbar_wait(&bars[TC_STAGES], 0);
    tc_fence_after();
  }

  // ---- epilogue: TMEM -> registers -> global.  warp w owns lanes [32*(w&3), +32) and column half (w>>2) ----
  const bool atomic = (MODE == I3D_GEMM_TN) && p.splits > 1;
  const bool add_bias = p.bias && !(atomic && blockIdx.z != 0);
  const int q = warp & 3, half = warp >> 2;
  const int64_t row = m0 + q * 32 + lane;
  constexpr int HALF_COLS = BN / 2;
  const int c_begin = half * HALF_COLS;
  if (total > 0 || !atomic) {
    for (int c = c_begin; c < c_begin + HALF_COLS; c += 16) {
      float v[16];
      if (total > 0) {
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);    // warp-collective: no divergence around it
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      const int lim = min(16, c_begin + HALF_COLS - c);
      if (row < M) {
        float* out = p.C + row * p.ldc + n0 + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + c + i;
          if (i < lim && n < N) {
            float o = v[i] + (add_bias ? __ldg(p.bias + n) : 0.f);
            if (atomic) {
              atomicAdd(out + i, o);
            } else {
              if (p.accumulate) o += out[i];
              out[i] = o;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

__global__ void tc_zero_block_kernel(float* __restrict__ C, int64_t M, int N, int ldc) {
  const int64_t total = M * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / N;
    C[m * ldc + (t - m * N)] = 0.f;
  }
}

template <int MODE, int BN>
static int launch_tc(TcParams& p, cudaStream_t s) {
  using L = TcLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
    if (e != cudaSuccess) {
      set_error("i3d_gemm(tc): cudaFuncSetAttribute -> %s", cudaGetErrorString(e));
      return I3D_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t gx = (p.M + TC_BM - 1) / TC_BM;
  const int gy = (p.N + BN - 1) / BN;
  int gz = 1;
  if (MODE == I3D_GEMM_TN) {
    const int K = p.seg[0].K;
    int64_t want = (sm_count() + gx * gy - 1) / (gx * gy);          // ~one CTA per SM
    const int64_t max_splits = (K + 4 * TC_BK - 1) / (4 * TC_BK);   // at least 4 k-blocks per CTA
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    int kchunk = (int)((K + want - 1) / want);
    kchunk = ((kchunk + TC_BK - 1) / TC_BK) * TC_BK;
    p.kchunk = kchunk;
    p.splits = (K + kchunk - 1) / kchunk;
    gz = p.splits;
    if (p.splits > 1 && !p.accumulate) {
      tc_zero_block_kernel<<<grid_for(p.M * p.N, 256), 256, 0, s>>>(p.C, p.M, p.N, p.ldc);
      count_launch();
    }
  }
  gemm_tc_kernel<MODE, BN><<<dim3((unsigned)gx, gy, gz), TC_THREADS, L::BYTES, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("i3d_gemm(tc): launch failed -> %s", cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  count_launch();
  return I3D_OK;
}

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

template <int MODE>
static int dispatch_bn(TcParams& p, cudaStream_t stream) {
  const int N = p.N;
  const int64_t gx = (p.M + TC_BM - 1) / TC_BM;
  if (N <= 32) return launch_tc<MODE, 32>(p, stream);
  if (N <= 64) return launch_tc<MODE, 64>(p, stream);
  if (N <= 112) return launch_tc<MODE, 112>(p, stream);
  if (N <= 128) return launch_tc<MODE, 128>(p, stream);
  // few row tiles (node-level GEMMs at batch 512: 72 tiles on 148 SMs): split N over two CTAs to fill the machine
  if (MODE == I3D_GEMM_NT && N <= 208 && gx * 2 <= sm_count()) return launch_tc<MODE, 112>(p, stream);
  if (N <= 208) return launch_tc<MODE, 208>(p, stream);
  if ((N + 207) / 208 <= (N + 255) / 256) return launch_tc<MODE, 208>(p, stream);   // same tile count, less padding
  return launch_tc<MODE, 256>(p, stream);
}

int gemm_tc(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
            int accumulate, cudaStream_t stream) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < n_seg; ++s) p.seg[s] = segs[s];
  p.n_seg = n_seg;
  p.M = M;
  p.N = N;
  p.C = C;
  p.ldc = ldc;
  p.bias = bias;
  p.accumulate = accumulate;
  p.kchunk = 0;
  p.splits = 1;
  if (mode == I3D_GEMM_NT) return dispatch_bn<I3D_GEMM_NT>(p, stream);
  return dispatch_bn<I3D_GEMM_TN>(p, stream);
}

}  // namespace i3d
