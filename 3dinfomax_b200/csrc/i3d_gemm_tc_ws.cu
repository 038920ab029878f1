// Warp-specialised NT tensor-core GEMM (the forward y = x W^T and, through W^T, dx = dy W).
//
//   C[m, n] = bias[n] + sum_s sum_k  scale_s[m] * A_s[a_idx_s[m], k] * B_s[n, k]
//
// Roles inside one 320-thread CTA (one CTA per SM, one 128 x BN output tile each):
//   warps 0-7  A stagers   : LDG (row-gathered, degree-scaled) -> tf32 hi/lo split -> STS into a K-major
//                            SWIZZLE_64B stage -> fence.proxy.async -> mbarrier arrive (a_full).  Two k-blocks of
//                            register prefetch; no CTA-wide barrier in the main loop.
//   warp  8    B producer  : one lane streams the pre-split weight tiles (hi, lo) with TMA
//                            (cp.async.bulk.tensor.2d, SWIZZLE_64B) up to STAGES k-blocks ahead (b_full, expect_tx).
//   warp  9    MMA issuer  : one lane waits a_full/b_full, issues 3xTF32 tcgen05.mma into the TMEM accumulator and
//                            tcgen05.commit's the stage back to the producers (mma_done).
//   epilogue   warps 0-7   : TMEM -> registers -> shared tile -> coalesced float4 stores.
//
// K is tiled in 16-float blocks (one 64-byte swizzle atom) so that 5 stages fit in shared memory: the B tiles of a
// k-block arrive ~2000 cycles after they are requested, a k-block of MMAs takes ~620 cycles, so the ring has to be
// at least 4 deep to keep the tensor pipe busy (ncu: 39 % tensor-active with the 2-stage kernel this one replaces).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace i3d {
__device__ unsigned long long g_ws_dbg[16];     // cycle counters of the -DI3D_WS_DEBUG build (see below)
__device__ long long g_ws_dbg_t0;
}
#ifdef I3D_WS_DEBUG
// epilogue phase marks of CTA (0,0), thread 0: cycles since the epilogue started at [9] tile in shared memory,
// [10] statistics done, [11] stores issued
#define I3D_TC_EPI_MARK(slot)                                                                     \
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)                                     \
  atomicAdd(&::i3d::g_ws_dbg[slot], (unsigned long long)(clock64() - ::i3d::g_ws_dbg_t0))
#endif
#include "i3d_tc.cuh"

namespace i3d {

constexpr int WS_BK = 16;                 // floats per k-block = one SWIZZLE_64B row
constexpr int WS_STAGER_THREADS = 256;    // warps 0-7
constexpr int WS_THREADS = 320;           // + TMA warp + MMA warp
#ifndef I3D_WS_EXP
#define I3D_WS_EXP 0      // timing experiments only (results are wrong): 1 no proxy fence, 2 no A loads, 3 no A stores
#endif
#ifndef I3D_WS_DUAL_ACC
#define I3D_WS_DUAL_ACC 0     // measured: no gain (55.7 vs 55.6 us on the K = 1000 posttrans GEMM), costs 2x TMEM columns
#endif
#ifndef I3D_WS_PREFETCH
#define I3D_WS_PREFETCH 2     // measured 2..8: no difference (the loop is MMA-issue bound), 2 keeps registers low
#endif
constexpr int WS_PREFETCH = I3D_WS_PREFETCH;   // k-blocks of A held in registers per stager thread

// Optional cycle accounting of the pipeline roles (build with -DI3D_WS_DEBUG; tools/gemm_bench.py --debug): CTA (0,0)
// adds the clock64() cycles each role spends blocked on each barrier.  [0] kernel total, [1] MMA waits a_full,
// [2] MMA waits b_full, [3] MMA issue, [4] TMA producer waits mma_done, [5] stager warp 0 waits mma_done,
// [6] stager warp 0 total main loop, [7] epilogue, [8] launches
#ifdef I3D_WS_DEBUG
#define WS_DBG_T0() const long long dbg_t0__ = clock64()
#define WS_DBG_ADD(slot) if (dbg_on) atomicAdd(&g_ws_dbg[slot], (unsigned long long)(clock64() - dbg_t0__))
#else
#define WS_DBG_T0()
#define WS_DBG_ADD(slot)
#endif

struct WsParams {
  CUtensorMap map_hi, map_lo;       // [N rows, Kpad_total cols] fp32, box = [BN rows x 16 cols], SWIZZLE_64B
  const float* A[4];
  const int32_t* a_idx[4];
  const float* scale[4];
  int32_t lda[4], K[4], kcol0[4];   // kcol0: first column of the segment inside the split buffers
  int n_seg;
  int64_t M;
  int N;
  float* C;
  int ldc;
  const float* bias;
  int accumulate;
  double* stats;                    // optional fused BatchNorm statistics [2N] (fp64), see tc_epilogue
  int stats_act;
  // B split (one launch for all segments)
  const float* Bsrc[4];
  int32_t ldb[4];
  // degree-bucketed mode (i3d_gemm_nt_bucketed): row tile t multiplies the weight block of bucket tile_bucket[t]
  // (rows [bucket * N, bucket * N + N) of the prepared B); row_map[m] is the output row of virtual row m, -1 = padding
  const int32_t* tile_bucket;
  const int32_t* row_map;
  int b_rows_total;                 // rows of the prepared B (n_buckets * N); 0 = plain mode (N rows)
  int b_pitch;                      // row pitch (floats) of the prepared B; 0 = the padded K extent
  const int32_t* m_valid;           // optional device scalar: output rows >= *m_valid are excluded from `stats`
};

// OCC = CTAs resident per SM.  OCC 1: deepest ring (5 stages at BN = 208).  OCC 2: two CTAs share an SM (2-3 stages
// each, the other CTA's MMAs cover this one's TMA latency) — used when there are more tiles than SMs, so that e.g.
// 152 row tiles run as one wave on 148 SMs instead of two.
template <int BN, int OCC = 1>
struct WsLayout {
  static constexpr int A_TILE = TC_BM * WS_BK;                 // floats (hi or lo)
  static constexpr int B_TILE = BN * WS_BK;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;       // a_hi, a_lo, b_hi, b_lo
  static constexpr int STAGE_BYTES = STAGE * 4;
  static constexpr int MAX_BYTES = OCC == 1 ? 220 * 1024 : 108 * 1024;
  static constexpr int STAGES = (MAX_BYTES / STAGE_BYTES) < 6 ? (MAX_BYTES / STAGE_BYTES) : 6;
  static constexpr int CTILE_BYTES = TC_BM * (BN + 4) * 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES > CTILE_BYTES ? STAGES * STAGE_BYTES : CTILE_BYTES;
  static constexpr size_t BYTES = (size_t)RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int ACC_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  // Optional second accumulator (-DI3D_WS_DUAL_ACC=1): hi*hi goes to the first, the two cross terms (lo*hi, hi*lo) to
  // the second, the epilogue adds them.  Written to test whether the ~160 cycles per 128 x N x 8 tf32 instruction seen
  // with the I3D_WS_DEBUG counters (same for N = 112 and 208) come from the accumulate dependency: they do not — two
  // independent chains run at the same pace, so the default stays one accumulator.
  static constexpr bool DUAL = (I3D_WS_DUAL_ACC != 0) && (2 * ACC_COLS * OCC <= 512);
  static constexpr int TMEM_COLS = DUAL ? 2 * ACC_COLS : ACC_COLS;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// float offset of 16-byte chunk j (0..3) of row r inside a [rows x 16 tf32] K-major SWIZZLE_64B tile
// (Swizzle<2,4,3>: byte-address bits [4,6) ^= bits [7,9); identical to CU_TENSOR_MAP_SWIZZLE_64B with a 64-byte box)
__device__ __forceinline__ int sw64_off(int r, int j) { return (r >> 3) * 128 + (r & 7) * 16 + ((j ^ ((r >> 1) & 3)) << 2); }

// K-major SWIZZLE_64B smem descriptor: SBO = 512 B (8 rows x 64 B), layout type 4
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

template <int BN, int OCC>
__global__ void __launch_bounds__(WS_THREADS, OCC) gemm_tc_nt_ws_kernel(const __grid_constant__ WsParams p) {
  pdl_grid_sync();
  using L = WsLayout<BN, OCC>;
  constexpr int S = L::STAGES;
  constexpr int PF = WS_PREFETCH;
  extern __shared__ uint8_t smem_raw[];
  float* tiles = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tiles) + L::RING_BYTES);
  uint64_t* mma_done = bars;            // [S] stage drained by the tensor core (count 1, tcgen05.commit)
  uint64_t* a_full = bars + S;          // [S] A tile staged (count 8: one arrival per stager warp)
  uint64_t* b_full = bars + 2 * S;      // [S] B tiles landed (count 1 + TMA transaction bytes)
  uint64_t* acc_done = bars + 3 * S;    // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * BN;
  const int64_t M = p.M;
  const int N = p.N;
  int b_row0 = n0;                      // first row of this CTA's B tile inside the prepared operand
  if (p.tile_bucket) {
    const int bucket = __ldg(p.tile_bucket + blockIdx.x);
    if (bucket < 0) return;             // unused tail tile of the degree plan (whole CTA, before any allocation)
    b_row0 += bucket * N;
  }
  const int32_t* __restrict__ row_map = p.row_map;
#ifdef I3D_WS_DEBUG
  const bool dbg_on = blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  const long long dbg_kernel_t0 = clock64();
#endif

  if (warp == 9) tmem_alloc(tmem_slot, L::TMEM_COLS);
  if (tid == 256) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&mma_done[s], 1);
      mbar_init(&a_full[s], 8);
      mbar_init(&b_full[s], 1);
    }
    mbar_init(acc_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.map_hi);
    tma_prefetch_desc(&p.map_lo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  int total = 0;
  for (int s = 0; s < p.n_seg; ++s) total += (p.K[s] + WS_BK - 1) / WS_BK;

  if (warp < 8) {
    // ======================================= A stagers =======================================================
    // chunk c = tid + i*256 (i < 2): row = c >> 2 = (tid >> 2) + 64 i, 16-byte chunk j = tid & 3
    int pf_seg = 0, pf_k0 = 0;
    int64_t a_row[2];
    float a_sc[2];
    auto bind_segment = [&](int s) {
      const int32_t* __restrict__ a_idx = p.a_idx[s];
      const float* __restrict__ scale = p.scale[s];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t gm = m0 + (tid >> 2) + 64 * i;
        a_row[i] = -1;
        a_sc[i] = 1.f;
        if (gm < M) {
          a_row[i] = a_idx ? (int64_t)__ldg(a_idx + gm) : ((row_map && __ldg(row_map + gm) < 0) ? -1 : gm);
          if (scale) a_sc[i] = __ldg(scale + gm);
        }
      }
    };
    bind_segment(0);
    // nothing in here may consume the loaded values (the scale is applied at store time): it is a prefetch
    auto prefetch = [&](float4 (&va)[2], float (&vs)[2]) {
      const float* __restrict__ A = p.A[pf_seg];
      const int lda = p.lda[pf_seg], K = p.K[pf_seg];
      const int kc = pf_k0 + (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#if I3D_WS_EXP != 2
        if (a_row[i] >= 0 && kc < K) v = __ldg(reinterpret_cast<const float4*>(A + a_row[i] * lda + kc));
#endif
        va[i] = v;
        vs[i] = a_sc[i];
      }
      pf_k0 += WS_BK;
      if (pf_k0 >= K && pf_seg + 1 < p.n_seg) {
        pf_seg += 1;
        pf_k0 = 0;
        bind_segment(pf_seg);
      }
    };
    auto body = [&](int it, float4 (&va)[2], float (&vs)[2]) {
      const int st = it % S;
      const int use = it / S;
      float* a_hi = tiles + (size_t)st * L::STAGE;
      float* a_lo = a_hi + L::A_TILE;
      if (use > 0) {
        WS_DBG_T0();
        mbar_wait(&mma_done[st], (uint32_t)((use - 1) & 1));
#ifdef I3D_WS_DEBUG
        if (warp == 0) { WS_DBG_ADD(5); }
#endif
      }
      const int j = tid & 3;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = va[i];
        const float sc = vs[i];
        v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
#if I3D_WS_EXP != 3
        split_store4(a_hi, a_lo, sw64_off((tid >> 2) + 64 * i, j), v);
#endif
      }
#if I3D_WS_EXP != 1
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core
#endif
      if (it + PF < total) prefetch(va, vs);     // refill with the k-block PF iterations ahead
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[st]);
    };
    // PF k-blocks of A are in flight per thread (registers).  ncu shows long_scoreboard as the top stall of the stager
    // warps, but they are not the bottleneck: with their loads removed they simply wait on mma_done instead
    // (-DI3D_WS_EXP=2), and PF = 2..8 gives the same kernel time
    float4 va[PF][2];
    float vs[PF][2];
#pragma unroll
    for (int q = 0; q < PF; ++q)
      if (total > q) prefetch(va[q], vs[q]);
    {
      WS_DBG_T0();
      for (int it = 0; it < total; it += PF) {
#pragma unroll
        for (int q = 0; q < PF; ++q)
          if (it + q < total) body(it + q, va[q], vs[q]);
      }
#ifdef I3D_WS_DEBUG
      if (warp == 0) { WS_DBG_ADD(6); }
#endif
    }
  } else if (warp == 8) {
    // ======================================= B producer (TMA) ================================================
    if (lane == 0) {
      int seg = 0, k0 = 0;
      for (int it = 0; it < total; ++it) {
        const int st = it % S;
        const int use = it / S;
        float* b_hi = tiles + (size_t)st * L::STAGE + 2 * L::A_TILE;
        float* b_lo = b_hi + L::B_TILE;
        if (use > 0) {
          WS_DBG_T0();
          mbar_wait(&mma_done[st], (uint32_t)((use - 1) & 1));
          WS_DBG_ADD(4);
        }
        const int kcol = p.kcol0[seg] + k0;
        mbar_expect_tx(&b_full[st], 2u * BN * WS_BK * 4u);
        tma_load_2d(b_hi, &p.map_hi, kcol, b_row0, &b_full[st]);
        tma_load_2d(b_lo, &p.map_lo, kcol, b_row0, &b_full[st]);
        k0 += WS_BK;
        if (k0 >= p.K[seg] && seg + 1 < p.n_seg) {
          seg += 1;
          k0 = 0;
        }
      }
    }
  } else {
    // ======================================= MMA issuer ======================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
      for (int it = 0; it < total; ++it) {
        const int st = it % S;
        const uint32_t par = (uint32_t)((it / S) & 1);
        float* a_hi = tiles + (size_t)st * L::STAGE;
        float* a_lo = a_hi + L::A_TILE;
        float* b_hi = a_lo + L::A_TILE;
        float* b_lo = b_hi + L::B_TILE;
        {
          WS_DBG_T0();
          mbar_wait(&a_full[st], par);
          WS_DBG_ADD(1);
        }
        {
          WS_DBG_T0();
          mbar_wait(&b_full[st], par);
          WS_DBG_ADD(2);
        }
        WS_DBG_T0();
        tc_fence_after();
        const uint64_t dah = make_smem_desc_sw64(smem_u32(a_hi)), dal = make_smem_desc_sw64(smem_u32(a_lo));
        const uint64_t dbh = make_smem_desc_sw64(smem_u32(b_hi)), dbl = make_smem_desc_sw64(smem_u32(b_lo));
#pragma unroll
        for (int ks = 0; ks < WS_BK / 8; ++ks) {
          const uint64_t koff = (uint64_t)(ks * 32 >> 4);      // 8 tf32 = 32 bytes along K inside the 64-byte row
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          if constexpr (L::DUAL) {
            umma_tf32(tmem, dah + koff, dbh + koff, idesc, acc);
            umma_tf32(tmem + L::ACC_COLS, dal + koff, dbh + koff, idesc, acc);
            umma_tf32(tmem + L::ACC_COLS, dah + koff, dbl + koff, idesc, 1u);
          } else {
            umma_tf32(tmem, dah + koff, dbh + koff, idesc, acc);
            umma_tf32(tmem, dal + koff, dbh + koff, idesc, 1u);
            umma_tf32(tmem, dah + koff, dbl + koff, idesc, 1u);
          }
        }
        umma_commit(&mma_done[st]);
        WS_DBG_ADD(3);
      }
      if (total > 0) umma_commit(acc_done);
    }
  }

  // ---- epilogue (warps 0-7 move the tile; all warps take part in the CTA barriers) ----
  if (total > 0) {
    mbar_wait(acc_done, 0);
    tc_fence_after();
  }
  __syncthreads();      // every role has left the operand ring before it is reused as the output tile
#ifdef I3D_WS_DEBUG
  const long long dbg_epi_t0 = clock64();
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_ws_dbg_t0 = dbg_epi_t0;
  __syncthreads();
#endif
  tc_epilogue<BN, WS_STAGER_THREADS>(tmem, tiles, total > 0, M, N, m0, n0, p.C, p.ldc, p.bias, p.accumulate, false,
                                     p.stats, p.stats_act, row_map, L::DUAL ? tmem + L::ACC_COLS : 0xffffffffu,
                                     p.m_valid);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, L::TMEM_COLS);
#ifdef I3D_WS_DEBUG
  if (dbg_on && warp == 0) {
    atomicAdd(&g_ws_dbg[0], (unsigned long long)(clock64() - dbg_kernel_t0));
    atomicAdd(&g_ws_dbg[7], (unsigned long long)(clock64() - dbg_epi_t0));
    atomicAdd(&g_ws_dbg[8], 1ull);
  }
#endif
}

// hi[n, kcol0_s + c] = tf32(B_s[n, c]),  lo = B_s - hi  for c < K_s;  zeros for K_s <= c < Kpad_s  (blockIdx.y = s)
struct WsSplitArgs {
  const float* B[4];
  int32_t ldb[4], K[4], kcol0[4], kpad[4];
  int N, ldo;
};
__global__ void ws_split_tf32_kernel(const WsSplitArgs a, float* __restrict__ hi, float* __restrict__ lo) {
  pdl_grid_sync();
  const int sg = blockIdx.y;
  const float* __restrict__ B = a.B[sg];
  const int K = a.K[sg], ldb = a.ldb[sg], col0 = a.kcol0[sg];
  const int quads = a.kpad[sg] >> 2;
  const int64_t total = (int64_t)a.N * quads;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(t / quads);
    const int c = (int)(t - (int64_t)n * quads) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < K) v = __ldg(reinterpret_cast<const float4*>(B + (int64_t)n * ldb + c));
    float4 h, l;
    h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
    l.x = v.x - h.x, l.y = v.y - h.y, l.z = v.z - h.z, l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi + (int64_t)n * a.ldo + col0 + c) = h;
    *reinterpret_cast<float4*>(lo + (int64_t)n * a.ldo + col0 + c) = l;
  }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*WsEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WsEncodeTiledFn ws_encode_tiled() {
  static WsEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WsEncodeTiledFn>(ptr);
  }
  return fn;
}

bool gemm_ws_available() { return ws_encode_tiled() != nullptr; }

// copies the 16 debug counters to `out` and clears them (all zero unless built with -DI3D_WS_DEBUG)
int gemm_ws_debug_read(unsigned long long* out) {
  if (cudaMemcpyFromSymbol(out, g_ws_dbg, sizeof(unsigned long long) * 16) != cudaSuccess) return I3D_ERR_CUDA;
  unsigned long long z[16] = {0};
  if (cudaMemcpyToSymbol(g_ws_dbg, z, sizeof(z)) != cudaSuccess) return I3D_ERR_CUDA;
  return I3D_OK;
}

static bool ws_make_b_map(CUtensorMap* map, float* base, int N, int ktot, int bn, int pitch = 0) {
  WsEncodeTiledFn enc = ws_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)(pitch > 0 ? pitch : ktot) * 4};
  const cuuint32_t box[2] = {(cuuint32_t)WS_BK, (cuuint32_t)bn};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int ws_kpad(int K) { return (K + 31) / 32 * 32; }   // segment columns padded to 32 (zeros)

size_t gemm_ws_bytes(int N, int n_seg, const i3d_gemm_seg* segs) {
  int64_t ktot = 0;
  for (int s = 0; s < n_seg; ++s) ktot += ws_kpad(segs[s].K);
  return (size_t)2 * (size_t)N * (size_t)ktot * sizeof(float) + 256;
}

template <int BN, int OCC = 1>
static int launch_ws(WsParams& p, float* hi, float* lo, int ktot, cudaStream_t s) {
  using L = WsLayout<BN, OCC>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_tc_nt_ws_kernel<BN, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
    if (e != cudaSuccess) {
      set_error("i3d_gemm(ws): cudaFuncSetAttribute -> %s", cudaGetErrorString(e));
      return I3D_ERR_CUDA;
    }
    configured = true;
  }
  const int b_rows = p.b_rows_total > 0 ? p.b_rows_total : p.N;
  if (!ws_make_b_map(&p.map_hi, hi, b_rows, ktot, BN, p.b_pitch) ||
      !ws_make_b_map(&p.map_lo, lo, b_rows, ktot, BN, p.b_pitch)) {
    set_error("i3d_gemm(ws): cuTensorMapEncodeTiled failed");
    return I3D_ERR_CUDA;
  }
  const int64_t gx = (p.M + TC_BM - 1) / TC_BM;
  const int gy = (p.N + BN - 1) / BN;
  launch(gemm_tc_nt_ws_kernel<BN, OCC>, dim3((unsigned)gx, gy, 1), WS_THREADS, L::BYTES, s, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("i3d_gemm(ws): launch failed -> %s", cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  count_launch();
  return I3D_OK;
}

// ---- batched preparation of many B operands (one launch per optimizer step) ---------------------------------
// item: one K-segment of one GEMM; 32x32 tiles over (n, padded k).  transposed items read B[c * ldb + n].
__global__ void __launch_bounds__(256) ws_prep_kernel(const i3d_prep_item* __restrict__ items, int n_items) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  // locate the item of this tile: items are sorted by tile0
  int lo_i = 0, hi_i = n_items - 1;
  const int t = blockIdx.x;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (items[mid].tile0 <= t) lo_i = mid; else hi_i = mid - 1;
  }
  const i3d_prep_item it = items[lo_i];
  const int ktiles = it.kpad >> 5;
  const int lt = t - it.tile0;
  const int n0 = (lt / ktiles) << 5, c0 = (lt % ktiles) << 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (it.transposed) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, n = n0 + tx;
      tile[ty + 8 * i][tx] = (c < it.K && n < it.N) ? __ldg(it.B + (int64_t)c * it.ldb + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i, c = c0 + tx;
      if (n < it.N) {
        const float v = tile[tx][ty + 8 * i];
        const float h = to_tf32(v);
        it.hi[(int64_t)n * it.ldo + it.col0 + c] = h;
        it.lo[(int64_t)n * it.ldo + it.col0 + c] = v - h;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i, c = c0 + tx;
      if (n < it.N) {
        const float v = c < it.K ? __ldg(it.B + (int64_t)n * it.ldb + c) : 0.f;
        const float h = to_tf32(v);
        it.hi[(int64_t)n * it.ldo + it.col0 + c] = h;
        it.lo[(int64_t)n * it.ldo + it.col0 + c] = v - h;
      }
    }
  }
}

int gemm_prep_describe(int N, int n_seg, const i3d_gemm_seg* segs, int transposed, void* ws, int tile0,
                       i3d_prep_item* out, int* tiles_out) {
  int ktot = 0;
  for (int s = 0; s < n_seg; ++s) ktot += ws_kpad(segs[s].K);
  float* hi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 127) & ~(uintptr_t)127);
  float* lo = hi + (size_t)N * ktot;
  int col0 = 0, tiles = 0;
  for (int s = 0; s < n_seg; ++s) {
    i3d_prep_item& it = out[s];
    it.B = segs[s].B, it.hi = hi, it.lo = lo;
    it.ldb = segs[s].ldb, it.N = N, it.K = segs[s].K, it.kpad = ws_kpad(segs[s].K), it.ldo = ktot, it.col0 = col0;
    it.transposed = transposed ? 1 : 0, it.tile0 = tile0 + tiles;
    col0 += it.kpad;
    tiles += ((N + 31) / 32) * (it.kpad / 32);
  }
  *tiles_out = tiles;
  return I3D_OK;
}

int gemm_prep_run(const i3d_prep_item* dev_items, int n_items, int total_tiles, cudaStream_t stream) {
  launch(ws_prep_kernel, total_tiles, 256, 0, stream, dev_items, n_items);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("i3d_gemm_prep_run: launch failed -> %s", cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  count_launch();
  return I3D_OK;
}

// tile shape by output width and number of row tiles
static int ws_dispatch(WsParams& p, float* hi, float* lo, int ktot, cudaStream_t stream) {
  const int64_t M = p.M;
  const int N = p.N;
  const int64_t gx = (M + TC_BM - 1) / TC_BM;
  const int sms = sm_count();
  // tuning override for tools/gemm_bench.py: I3D_WS_FORCE="<BN>,<OCC>" (only shapes with N <= 208)
  static int force_bn = -1, force_occ = 1;
  if (force_bn < 0) {
    const char* e = getenv("I3D_WS_FORCE");
    force_bn = 0;
    if (e && sscanf(e, "%d,%d", &force_bn, &force_occ) < 1) force_bn = 0;
  }
  if (force_bn > 0 && N > 64 && N <= 208) {
    if (force_bn == 112) return force_occ == 2 ? launch_ws<112, 2>(p, hi, lo, ktot, stream) : launch_ws<112>(p, hi, lo, ktot, stream);
    if (force_bn == 208) return force_occ == 2 ? launch_ws<208, 2>(p, hi, lo, ktot, stream) : launch_ws<208>(p, hi, lo, ktot, stream);
  }
  // narrow outputs (Net3D: 20 columns over ~160k edge rows = ~1260 row tiles): two CTAs per SM halve the waves
  if (N <= 32) return gx > sms ? launch_ws<32, 2>(p, hi, lo, ktot, stream) : launch_ws<32>(p, hi, lo, ktot, stream);
  if (N <= 64) return gx > sms ? launch_ws<64, 2>(p, hi, lo, ktot, stream) : launch_ws<64>(p, hi, lo, ktot, stream);
  if (N <= 112) return gx > sms ? launch_ws<112, 2>(p, hi, lo, ktot, stream) : launch_ws<112>(p, hi, lo, ktot, stream);
  if (N <= 128) return launch_ws<128>(p, hi, lo, ktot, stream);
  if (N <= 208) {
    // One accumulator covers the F = 200 outputs of a PNA layer.  Split N over two CTAs when the row tiles alone leave
    // half the SMs idle (node-level GEMMs: 72 row tiles at batch 512).
    if (2 * gx <= sms) return launch_ws<112>(p, hi, lo, ktot, stream);
    if (gx > sms) return launch_ws<208, 2>(p, hi, lo, ktot, stream);          // more tiles than SMs: co-resident pairs
    // between half and all of the SMs (node-level GEMMs at batch 512: 77 row tiles): 208-wide tiles would leave half the
    // machine idle, 112-wide ones need two waves at one CTA per SM -> two co-resident 112-wide CTAs per SM, one wave
    static int split = -1;
    if (split < 0) {
      const char* e = getenv("I3D_WS_SPLIT_MID");
      split = e ? atoi(e) : 0;      // measured neutral at batch 512 (4.390 vs 4.392 ms per step): off by default
    }
    if (split && 4 * gx <= 3 * (int64_t)sms + sms) return launch_ws<112, 2>(p, hi, lo, ktot, stream);
    return launch_ws<208>(p, hi, lo, ktot, stream);
  }
  if ((N + 207) / 208 <= (N + 255) / 256) {                                  // same tile count, less padding
    if (gx * ((N + 207) / 208) > sms) return launch_ws<208, 2>(p, hi, lo, ktot, stream);
    return launch_ws<208>(p, hi, lo, ktot, stream);
  }
  return launch_ws<256>(p, hi, lo, ktot, stream);
}

// NT GEMM through the warp-specialised kernel.  ws: device scratch of gemm_ws_bytes(...) for the hi/lo weight copies.
int gemm_ws_nt(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
               int accumulate, void* ws, double* stats, int stats_act, cudaStream_t stream, bool prepared,
               const int32_t* m_valid) {
  WsParams p;
  memset(&p, 0, sizeof(p));
  p.m_valid = m_valid;
  int ktot = 0;
  for (int s = 0; s < n_seg; ++s) {
    p.A[s] = segs[s].A, p.a_idx[s] = segs[s].a_idx, p.scale[s] = segs[s].scale;
    p.lda[s] = segs[s].lda, p.K[s] = segs[s].K, p.kcol0[s] = ktot;
    ktot += ws_kpad(segs[s].K);
  }
  p.n_seg = n_seg, p.M = M, p.N = N, p.C = C, p.ldc = ldc, p.bias = bias, p.accumulate = accumulate;
  p.stats = stats, p.stats_act = stats_act;
  float* hi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 127) & ~(uintptr_t)127);
  float* lo = hi + (size_t)N * ktot;
  if (!prepared) {
    WsSplitArgs a;
    memset(&a, 0, sizeof(a));
    int kmax = 0;
    for (int s = 0; s < n_seg; ++s) {
      a.B[s] = segs[s].B, a.ldb[s] = segs[s].ldb, a.K[s] = segs[s].K, a.kcol0[s] = p.kcol0[s];
      a.kpad[s] = ws_kpad(segs[s].K);
      kmax = a.kpad[s] > kmax ? a.kpad[s] : kmax;
    }
    a.N = N, a.ldo = ktot;
    launch(ws_split_tf32_kernel, dim3(grid_for((int64_t)N * (kmax / 4), 256), n_seg), 256, 0, stream, a, hi, lo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("i3d_gemm(ws): split launch failed -> %s", cudaGetErrorString(e));
      return I3D_ERR_CUDA;
    }
    count_launch();
  }
  return ws_dispatch(p, hi, lo, ktot, stream);
}

// Degree-bucketed NT GEMM (see i3d_degree_plan / i3d_posttrans_merge): M = virtual rows (multiple of 128), B = the
// prepared [n_buckets * N, ktot] hi/lo operands, output rows scattered through row_map.
int gemm_ws_nt_bucketed(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
                        const float* hi, const float* lo, int b_pitch, int n_buckets, const int32_t* tile_bucket,
                        const int32_t* row_map, double* stats, int stats_act, cudaStream_t stream,
                        const int32_t* m_valid) {
  WsParams p;
  memset(&p, 0, sizeof(p));
  p.m_valid = m_valid;
  int ktot = 0;
  for (int s = 0; s < n_seg; ++s) {
    p.A[s] = segs[s].A, p.a_idx[s] = segs[s].a_idx, p.scale[s] = segs[s].scale;
    p.lda[s] = segs[s].lda, p.K[s] = segs[s].K, p.kcol0[s] = ktot;
    ktot += ws_kpad(segs[s].K);
  }
  p.n_seg = n_seg, p.M = M, p.N = N, p.C = C, p.ldc = ldc, p.bias = bias, p.accumulate = 0;
  p.stats = stats, p.stats_act = stats_act;
  p.tile_bucket = tile_bucket, p.row_map = row_map, p.b_rows_total = n_buckets * N, p.b_pitch = b_pitch;
  return ws_dispatch(p, const_cast<float*>(hi), const_cast<float*>(lo), ktot, stream);
}

}  // namespace i3d
