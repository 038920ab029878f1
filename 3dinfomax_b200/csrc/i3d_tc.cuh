// tcgen05 / TMEM / mbarrier / TMA building blocks shared by the tensor-core GEMM kernels (sm_100a inline PTX).
#pragma once
#include "i3d_common.cuh"

namespace i3d {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 2, TC_THREADS = 256;

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// TMA: 2-D tiled bulk tensor load global -> shared, completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
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
// (identical to what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B for a 128-byte inner box)
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

// issue the 3xTF32 MMAs of one 32-deep k-block: D += Ahi*Bhi + Alo*Bhi + Ahi*Blo  (4 k-steps of 8)
__device__ __forceinline__ void issue_kblock(uint32_t tmem, const float* a_hi, const float* a_lo, const float* b_hi,
                                             const float* b_lo, uint32_t idesc, bool first) {
  const uint64_t dah = make_smem_desc(smem_u32(a_hi)), dal = make_smem_desc(smem_u32(a_lo));
  const uint64_t dbh = make_smem_desc(smem_u32(b_hi)), dbl = make_smem_desc(smem_u32(b_lo));
#pragma unroll
  for (int ks = 0; ks < TC_BK / 8; ++ks) {
    const uint64_t koff = (uint64_t)(ks * 32 >> 4);          // 8 tf32 = 32 bytes along K inside the swizzle atom
    umma_tf32(tmem, dah + koff, dbh + koff, idesc, (!first || ks > 0) ? 1u : 0u);
    umma_tf32(tmem, dal + koff, dbh + koff, idesc, 1u);
    umma_tf32(tmem, dah + koff, dbl + koff, idesc, 1u);
  }
}

template <int BN>
struct TcLayout {
  static constexpr int A_TILE = TC_BM * TC_BK;                 // floats
  static constexpr int B_TILE = BN * TC_BK;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;       // a_hi, a_lo, b_hi, b_lo
  static constexpr size_t BYTES = (size_t)TC_STAGES * STAGE * 4 + 1024 /*align slack*/ + 128 /*barriers*/;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int B_CHUNKS = (BN * 8 + TC_THREADS - 1) / TC_THREADS;     // float4 per thread per k-block
};

// ---- epilogue shared by all tensor-core kernels -----------------------------------------------------------------
// TMEM -> registers (tcgen05.ld 32x32b: warp w owns lanes [32*(w&3), +32) and column half w>>2) -> row-major tile in
// shared memory (the operand stages are dead by now) -> coalesced float4 stores / vector atomics to global.
// phase timers of the epilogue (only the -DI3D_WS_DEBUG build of the warp-specialised NT kernel defines them)
#ifndef I3D_TC_EPI_MARK
#define I3D_TC_EPI_MARK(slot)
#endif

template <int BN, int NTHR = TC_THREADS>
__device__ __forceinline__ void tc_epilogue(uint32_t tmem, float* ctile, bool has_acc, int64_t M, int N, int64_t m0,
                                            int n0, float* __restrict__ C, int ldc, const float* __restrict__ bias,
                                            int accumulate, bool atomic, double* __restrict__ stats = nullptr,
                                            int stats_act = 0, const int32_t* __restrict__ row_map = nullptr,
                                            uint32_t tmem2 = 0xffffffffu,
                                            const int32_t* __restrict__ m_valid = nullptr) {
  // m_valid (shape-bucketed batches): output rows >= *m_valid are padding — stored like any row, but excluded from the
  // fused BatchNorm statistics
  // tmem2: optional second accumulator (same shape) that is added to the first while the tile is drained
  // row_map (degree-bucketed GEMMs): tile row r is written to output row row_map[m0 + r]; negative = padding row that
  // is neither stored nor counted in the statistics
  __shared__ int32_t s_row[TC_BM];
  if (row_map) {
    for (int r = threadIdx.x; r < TC_BM; r += blockDim.x) s_row[r] = (m0 + r < M) ? __ldg(row_map + m0 + r) : -1;
  }
  constexpr int LDT = BN + 4;               // row stride (floats): 16-byte aligned, conflict-free for per-row STS.128
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  constexpr int HALF_COLS = BN / 2;
  const int c_begin = half * HALF_COLS;
  float* trow = ctile + (q * 32 + lane) * LDT;
  for (int c = c_begin; tid < NTHR && c < c_begin + HALF_COLS; c += 16) {   // warp-uniform: NTHR is a multiple of 32
    float v[16];
    if (has_acc) {
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);    // warp-collective: no divergence around it
      if (tmem2 != 0xffffffffu) {
        float w[16];
        tmem_ld16(tmem2 + ((uint32_t)(q * 32) << 16) + (uint32_t)c, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    const int lim = min(16, c_begin + HALF_COLS - c);
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      if (i < lim) *reinterpret_cast<float4*>(trow + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  __syncthreads();
  I3D_TC_EPI_MARK(9);
  // (statistics and stores on disjoint threads at the same time — 96 store threads, the other 224 on the fp64 pipe — was
  //  measured: correct, slower: edge FC 25.0 -> 29.4 us, step 4.20 -> 4.31 ms; the store phase needs all the threads)
  if (stats) {
    // fused FCLayer statistics (models/base_layers.py:102-110): column sums of act(y) and act(y)^2 over the valid rows
    // of this tile, accumulated in fp64 (BatchNorm inputs with mean^2 >> var), one atomic pair per column and CTA
    const int64_t mv = m_valid ? (int64_t)__ldg(m_valid) : (int64_t)0x7fffffff;
    const int64_t mlim = (!row_map && mv < M) ? mv : M;
    const int rows = (int)(mlim - m0 < TC_BM ? (mlim - m0 > 0 ? mlim - m0 : 0) : TC_BM);
    for (int c = tid; tid < NTHR && c < BN; c += NTHR) {
      if (n0 + c >= N) continue;
      const float b = bias ? __ldg(bias + n0 + c) : 0.f;
      // four independent fp64 chains (the loop is bound by the DADD/DFMA dependency, not by the shared-memory reads).
      // Measured with the I3D_WS_DEBUG counters: 4.3 k cycles per 128 x 208 tile (F2F + DADD + DFMA per element on the
      // fp64 pipe).  Summing groups of 4 rows in fp32 first would cut that ~3x, but the squares must stay exact: Net3D's
      // BatchNorm inputs have mean^2 >> var, and fp32 partial sums of h^2 put ~1e-4 relative error on the variance.
      double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
      int r = 0;
      for (; r + 4 <= rows; r += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool valid = !row_map || (s_row[r + u] >= 0 && s_row[r + u] < mv);
          const double h = valid ? (double)act_apply(ctile[(r + u) * LDT + c] + b, stats_act) : 0.0;
          s1[u] += h;
          s2[u] = fma(h, h, s2[u]);
        }
      }
      for (; r < rows; ++r) {
        if (row_map && (s_row[r] < 0 || s_row[r] >= mv)) continue;
        const double h = (double)act_apply(ctile[r * LDT + c] + b, stats_act);
        s1[0] += h;
        s2[0] = fma(h, h, s2[0]);
      }
      atomicAdd(stats + (int64_t)(n0 + c) * I3D_STATS_STRIDE, (s1[0] + s1[1]) + (s1[2] + s1[3]));
      atomicAdd(stats + (int64_t)(N + n0 + c) * I3D_STATS_STRIDE, (s2[0] + s2[1]) + (s2[2] + s2[3]));
    }
  }
  I3D_TC_EPI_MARK(10);
  const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) && ((N & 3) == 0);
  constexpr int QUADS = BN / 4;
  for (int idx = tid; tid < NTHR && idx < TC_BM * QUADS; idx += NTHR) {
    const int r = idx / QUADS, c = (idx - r * QUADS) * 4;
    const int64_t gm = m0 + r;
    const int gn = n0 + c;
    if (gm >= M || gn >= N) continue;
    int64_t gm_out = gm;
    if (row_map) {
      gm_out = s_row[r];
      if (gm_out < 0) continue;
    }
    float4 v = *reinterpret_cast<const float4*>(ctile + r * LDT + c);
    float* out = C + gm_out * ldc + gn;
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
  I3D_TC_EPI_MARK(11);
}

}  // namespace i3d
