// Tensor-core path of the segmented virtual-operand GEMM: tcgen05.mma (kind::tf32) with the accumulator in TMEM.
//
//   C[m, n] = bias[n] + sum_s sum_k  scale_s[m] * A_s[idx_s[m], k] * B_s[n, k]            (NT: y = x W^T)
//
// Why a hand-written kernel instead of cuBLAS: the A operand is *virtual* — rows are gathered through the CSR
// edge lists (h[src], h[dst]) and scaled by the per-node degree scalers while they are staged, so neither
// torch.cat nor index_select ever touches HBM (models/pna.py:207,232,249).  TMA cannot express that staging, so
// operand tiles are produced by the CTA's own threads (LDG -> split -> STS in the canonical K-major SWIZZLE_128B
// layout), published to the async proxy with fence.proxy.async, and consumed by tcgen05.mma issued by one thread.
//
// Precision: the reference computes in fp32 (SGEMM).  Each fp32 operand is split into hi = tf32(x) and lo = x - hi
// and three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32 in TMEM ("3xTF32"), which keeps the result at fp32
// rounding level (measured in tests/gpu_cases.py::case_gemm_tc) at one third of the tf32 tensor throughput.
//
// Tile: 128 (M) x BN (N <= 256, multiple of 16) x 32 (K, one 128-byte swizzle atom of tf32), 2 smem stages,
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

__device__ __forceinline__ void split_store(float* hi_tile, float* lo_tile, int off, float4 v) {
  float4 h, l;
  h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
  l.x = v.x - h.x, l.y = v.y - h.y, l.z = v.z - h.z, l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

template <int BN>
struct TcLayout {
  static constexpr int A_TILE = TC_BM * TC_BK;                 // floats
  static constexpr int B_TILE = BN * TC_BK;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;       // a_hi, a_lo, b_hi, b_lo
  static constexpr size_t BYTES = (size_t)TC_STAGES * STAGE * 4 + 1024 /*align slack*/ + 64 /*barriers*/;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_nt_kernel(const __grid_constant__ TcParams p) {
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

  int it = 0;
  for (int s = 0; s < p.n_seg; ++s) {
    const float* __restrict__ A = p.seg[s].A;
    const float* __restrict__ B = p.seg[s].B;
    const int32_t* __restrict__ a_idx = p.seg[s].a_idx;
    const float* __restrict__ scale = p.seg[s].scale;
    const int lda = p.seg[s].lda, ldb = p.seg[s].ldb, K = p.seg[s].K;
    // fixed rows per thread: A chunks c = tid + i*256 -> row = c>>3 (i*32 + tid>>3), j = tid&7
    int64_t a_row[4];
    float a_sc[4];
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
    const int j = tid & 7;
    for (int k0 = 0; k0 < K; k0 += TC_BK, ++it) {
      const int st = it % TC_STAGES;
      const int use = it / TC_STAGES;
      float* a_hi = tiles + (size_t)st * L::STAGE;
      float* a_lo = a_hi + L::A_TILE;
      float* b_hi = a_lo + L::A_TILE;
      float* b_lo = b_hi + L::B_TILE;
      // ---- global loads first (all in flight), zero-filled outside the matrix ----
      const int kc = k0 + j * 4;
      float4 va[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_row[i] >= 0 && kc < K) va[i] = __ldg(reinterpret_cast<const float4*>(A + a_row[i] * lda + kc));
      }
      constexpr int B_CHUNKS = (BN * 8 + TC_THREADS - 1) / TC_THREADS;
      float4 vb[B_CHUNKS];
#pragma unroll
      for (int i = 0; i < B_CHUNKS; ++i) {
        const int c = tid + i * TC_THREADS;
        const int r = c >> 3;
        vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < BN && n0 + r < N && kc < K)
          vb[i] = __ldg(reinterpret_cast<const float4*>(B + (int64_t)(n0 + r) * ldb + kc));
      }
      // ---- the MMAs that last read this stage must have completed before it is overwritten ----
      if (use > 0) mbar_wait(&bars[st], (uint32_t)((use - 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = va[i];
        const float sc = a_sc[i];
        v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
        split_store(a_hi, a_lo, sw128_off(i * 32 + (tid >> 3), j), v);
      }
#pragma unroll
      for (int i = 0; i < B_CHUNKS; ++i) {
        const int c = tid + i * TC_THREADS;
        const int r = c >> 3;
        if (r < BN) split_store(b_hi, b_lo, sw128_off(r, j), vb[i]);
      }
      fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
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
  }
  if (tid == 0) umma_commit(&bars[TC_STAGES]);      // accumulator complete
  mbar_wait(&bars[TC_STAGES], 0);
  tc_fence_after();

  // ---- epilogue: TMEM -> registers -> global.  warp w owns lanes [32*(w&3), +32) and column half (w>>2) ----
  const int q = warp & 3, half = warp >> 2;
  const int64_t row = m0 + q * 32 + lane;
  constexpr int HALF_COLS = BN / 2;                 // BN is a multiple of 32 or 208 (=2*104): chunks of 16, 8 tail
  const int c_begin = half * HALF_COLS;
  for (int c = c_begin; c < c_begin + HALF_COLS; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);    // warp-collective: no divergence around it
    const int lim = min(16, c_begin + HALF_COLS - c);
    if (row < M) {
      float* out = p.C + row * p.ldc + n0 + c;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = n0 + c + i;
        if (i < lim && n < N) {
          float o = v[i] + (p.bias ? __ldg(p.bias + n) : 0.f);
          if (p.accumulate) o += out[i];
          out[i] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

template <int BN>
static int launch_tc(const TcParams& p, cudaStream_t s) {
  using L = TcLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_nt_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
    if (e != cudaSuccess) {
      set_error("i3d_gemm(tc): cudaFuncSetAttribute -> %s", cudaGetErrorString(e));
      return I3D_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t gx = (p.M + TC_BM - 1) / TC_BM;
  const int gy = (p.N + BN - 1) / BN;
  gemm_tc_nt_kernel<BN><<<dim3((unsigned)gx, gy, 1), TC_THREADS, L::BYTES, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("i3d_gemm(tc): launch failed -> %s", cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  count_launch();
  return I3D_OK;
}

static inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// NT problems the tensor-core kernel takes: every segment 16-byte aligned with K % 4 == 0 (float4 staging)
bool gemm_tc_eligible(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs) {
  if (mode != I3D_GEMM_NT || M < 256 || N < 16) return false;
  for (int s = 0; s < n_seg; ++s) {
    if (segs[s].K <= 0 || (segs[s].K & 3) || (segs[s].lda & 3) || (segs[s].ldb & 3) || !al16(segs[s].A) ||
        !al16(segs[s].B) || segs[s].b_idx)
      return false;
  }
  return true;
}

int gemm_tc_nt(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
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
  if (N <= 32) return launch_tc<32>(p, stream);
  if (N <= 64) return launch_tc<64>(p, stream);
  if (N <= 128) return launch_tc<128>(p, stream);
  if (N <= 208) return launch_tc<208>(p, stream);
  return launch_tc<256>(p, stream);
}

}  // namespace i3d
