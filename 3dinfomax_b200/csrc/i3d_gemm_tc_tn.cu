// Warp-specialised TN tensor-core GEMM: the weight gradients  dW[m, n] (+)= sum_k scale[k] * dY[a_idx[k], m] * X[b_idx[k], n].
//
// Both operands are activations stored [k rows][m or n contiguous]: exactly the *MN-major* operand form of
// tcgen05.mma (kind::tf32 allows it), so tiles are staged with coalesced float4 loads and float4 shared stores in
// the canonical MN-major layout — no transposition (the first TN kernel transposed with scalar stores and reached
// 10-13 % tensor-pipe activity, profiles/r01_*).  For 32-bit operands the only MN-major layout the tensor core accepts
// is SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom: Swizzle<2,5,2> over Shape<1024 bit, 4>):
// an atom is 4 k-rows x 128 bytes (32 tf32 along M/N); inside an atom the 32-byte chunk index is XOR-ed with the
// k-row index; atoms along M/N are LBO bytes apart, groups of 4 k-rows SBO bytes apart (one K=8 MMA reads two groups).
//
// M, N are a few hundred while K is the number of edges / nodes, so K is split over gridDim.z and the partial tiles
// are reduced with float4 vector atomics from the shared-memory staged epilogue.  Roles per 288-thread CTA:
// warps 0-7 stage A and B (gather indices / scales resolved one k-block ahead, two k-blocks of register prefetch),
// warp 8 issues the 3xTF32 MMAs; stages are handed over with mbarriers only.
#include "i3d_tc.cuh"

namespace i3d {

constexpr int TN_BK = 16;
constexpr int TN_STAGERS = 256;
constexpr int TN_THREADS = 288;

struct TnParams {
  const float* A;        // [K rows, M cols] (dY)
  const float* B;        // [K rows (gathered by b_idx), N cols] (x)
  const int32_t* a_idx;
  const int32_t* b_idx;
  const float* scale;    // per k
  int lda, ldb, K;
  int64_t M;
  int N;
  float* C;
  int ldc;
  int accumulate;
  int kchunk, splits;
};

template <int BN>
struct TnLayout {
  static constexpr int A_TILE = TN_BK * TC_BM;                 // floats (hi or lo): 4 k-groups x 4 atoms x 512 B
  static constexpr int B_TILE = TN_BK * BN;                    // 4 k-groups x BN/32 atoms x 512 B
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGE_BYTES = STAGE * 4;
  static constexpr int MAX_BYTES = 220 * 1024;
  static constexpr int STAGES = (MAX_BYTES / STAGE_BYTES) < 6 ? (MAX_BYTES / STAGE_BYTES) : 6;
  static constexpr int CTILE_BYTES = TC_BM * (BN + 4) * 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES > CTILE_BYTES ? STAGES * STAGE_BYTES : CTILE_BYTES;
  static constexpr size_t BYTES = (size_t)RING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int BQ = BN / 4;                            // float4 per k-row of the B tile
  static constexpr int B_CHUNKS = (TN_BK * BQ + TN_STAGERS - 1) / TN_STAGERS;
};

__device__ __forceinline__ void tn_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MN-major SWIZZLE_128B_BASE32B descriptor (layout type 1): LBO = bytes between atoms along M/N, SBO = bytes between
// 4-row k groups
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// instruction descriptor with both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
  return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);
}

// float offset of (k-row kk in [0,16), float4 index q along M/N) inside an MN-major tile with `atoms` 32-wide atoms:
// atom (kk>>2, q>>3) of 512 bytes, row r = kk&3 of 128 bytes, 32-byte chunk ((q&7)>>1) ^ r, 16-byte half q&1
__device__ __forceinline__ int mn_off(int kk, int q, int atoms) {
  const int r = kk & 3;
  return ((kk >> 2) * atoms + (q >> 3)) * 128 + r * 32 + (((((q & 7) >> 1) ^ r) << 3)) + ((q & 1) << 2);
}

template <int BN>
__global__ void __launch_bounds__(TN_THREADS, 1) gemm_tc_tn_ws_kernel(const __grid_constant__ TnParams p) {
  pdl_grid_sync();
  using L = TnLayout<BN>;
  constexpr int S = L::STAGES;
  constexpr int A_ATOMS = TC_BM / 32, B_ATOMS = BN / 32;
  extern __shared__ uint8_t smem_raw[];
  float* tiles = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tiles) + L::RING_BYTES);
  uint64_t* mma_done = bars;          // [S]
  uint64_t* full = bars + S;          // [S] count 8 (one arrival per stager warp)
  uint64_t* acc_done = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * BN;
  const int64_t M = p.M;
  const int N = p.N;

  if (warp == 8) tmem_alloc(tmem_slot, L::TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&mma_done[s], 1);
      mbar_init(&full[s], 8);
    }
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const int total = kend > kbeg ? (kend - kbeg + TN_BK - 1) / TN_BK : 0;

  if (warp < 8) {
    // ============================================ stagers =====================================================
    // A chunks: c = tid + 256 i (i < 2): k-row = c >> 5, float4 q = c & 31.  B chunks: c = tid + 256 i: k-row = c / BQ
    int a_kk[2], a_q[2], b_kk[L::B_CHUNKS], b_q[L::B_CHUNKS];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * TN_STAGERS;
      a_kk[i] = c >> 5;
      a_q[i] = c & 31;
    }
#pragma unroll
    for (int i = 0; i < L::B_CHUNKS; ++i) {
      const int c = tid + i * TN_STAGERS;
      b_kk[i] = c / L::BQ;
      b_q[i] = c - b_kk[i] * L::BQ;
    }
    // rows / scales of the k-block that the NEXT prefetch will read (resolved one block ahead)
    int64_t arow[2], brow[L::B_CHUNKS];
    float asc[2];
    int pf_k0 = kbeg;
    auto resolve = [&](int k0) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + a_kk[i];
        arow[i] = -1;
        asc[i] = 1.f;
        if (k < kend) {
          arow[i] = p.a_idx ? (int64_t)__ldg(p.a_idx + k) : (int64_t)k;
          if (p.scale) asc[i] = __ldg(p.scale + k);
        }
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int k = k0 + b_kk[i];
        brow[i] = -1;
        if (b_kk[i] < TN_BK && k < kend) brow[i] = p.b_idx ? (int64_t)__ldg(p.b_idx + k) : (int64_t)k;
      }
    };
    resolve(pf_k0);
    auto prefetch = [&](float4 (&va)[2], float (&vs)[2], float4 (&vb)[L::B_CHUNKS]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t m = m0 + 4 * a_q[i];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (arow[i] >= 0 && m < M) v = __ldg(reinterpret_cast<const float4*>(p.A + arow[i] * p.lda + m));
        va[i] = v;
        vs[i] = asc[i];
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i) {
        const int n = n0 + 4 * b_q[i];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (brow[i] >= 0 && n < N) v = __ldg(reinterpret_cast<const float4*>(p.B + brow[i] * p.ldb + n));
        vb[i] = v;
      }
      pf_k0 += TN_BK;
      resolve(pf_k0);         // index loads for the following block: complete long before they are needed
    };
    auto body = [&](int it, float4 (&va)[2], float (&vs)[2], float4 (&vb)[L::B_CHUNKS]) {
      const int st = it % S;
      const int use = it / S;
      float* a_hi = tiles + (size_t)st * L::STAGE;
      float* a_lo = a_hi + L::A_TILE;
      float* b_hi = a_lo + L::A_TILE;
      float* b_lo = b_hi + L::B_TILE;
      if (use > 0) mbar_wait(&mma_done[st], (uint32_t)((use - 1) & 1));
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = va[i];
        const float sc = vs[i];
        v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
        split_store4(a_hi, a_lo, mn_off(a_kk[i], a_q[i], A_ATOMS), v);
      }
#pragma unroll
      for (int i = 0; i < L::B_CHUNKS; ++i)
        if (b_kk[i] < TN_BK) split_store4(b_hi, b_lo, mn_off(b_kk[i], b_q[i], B_ATOMS), vb[i]);
      fence_proxy_async();
      if (it + 2 < total) prefetch(va, vs, vb);
      __syncwarp();
      if (lane == 0) tn_mbar_arrive(&full[st]);
    };
    float4 va0[2], va1[2], vb0[L::B_CHUNKS], vb1[L::B_CHUNKS];
    float vs0[2], vs1[2];
    if (total > 0) prefetch(va0, vs0, vb0);
    if (total > 1) prefetch(va1, vs1, vb1);
    for (int it = 0; it < total; it += 2) {
      body(it, va0, vs0, vb0);
      if (it + 1 < total) body(it + 1, va1, vs1, vb1);
    }
  } else {
    // ============================================ MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32_mn(TC_BM, BN);
      for (int it = 0; it < total; ++it) {
        const int st = it % S;
        const uint32_t par = (uint32_t)((it / S) & 1);
        float* a_hi = tiles + (size_t)st * L::STAGE;
        float* a_lo = a_hi + L::A_TILE;
        float* b_hi = a_lo + L::A_TILE;
        float* b_lo = b_hi + L::B_TILE;
        mbar_wait(&full[st], par);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TN_BK / 8; ++ks) {
          const uint32_t aoff = ks * 2 * A_ATOMS * 512, boff = ks * 2 * B_ATOMS * 512;   // two 4-row k groups per MMA
          const uint64_t dah = make_desc_mn_sw128(smem_u32(a_hi) + aoff, 512, A_ATOMS * 512);
          const uint64_t dal = make_desc_mn_sw128(smem_u32(a_lo) + aoff, 512, A_ATOMS * 512);
          const uint64_t dbh = make_desc_mn_sw128(smem_u32(b_hi) + boff, 512, B_ATOMS * 512);
          const uint64_t dbl = make_desc_mn_sw128(smem_u32(b_lo) + boff, 512, B_ATOMS * 512);
          umma_tf32(tmem, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(tmem, dal, dbh, idesc, 1u);
          umma_tf32(tmem, dah, dbl, idesc, 1u);
        }
        umma_commit(&mma_done[st]);
      }
      if (total > 0) umma_commit(acc_done);
    }
  }

  if (total > 0) {
    mbar_wait(acc_done, 0);
    tc_fence_after();
  }
  __syncthreads();
  const bool atomic = p.splits > 1;
  tc_epilogue<BN, TN_STAGERS>(tmem, tiles, total > 0, M, N, m0, n0, p.C, p.ldc, nullptr, p.accumulate, atomic);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, L::TMEM_COLS);
}

__global__ void tn_zero_block_kernel(float* __restrict__ C, int64_t M, int N, int ldc) {
  pdl_grid_sync();
  const int64_t total = M * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / N;
    C[m * ldc + (t - m * N)] = 0.f;
  }
}

template <int BN>
static int launch_tn(TnParams& p, cudaStream_t s) {
  using L = TnLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_tc_tn_ws_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
    if (e != cudaSuccess) {
      set_error("i3d_gemm(tn): cudaFuncSetAttribute -> %s", cudaGetErrorString(e));
      return I3D_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t gx = (p.M + TC_BM - 1) / TC_BM;
  const int gy = (p.N + BN - 1) / BN;
  int64_t want = (sm_count() + gx * gy - 1) / (gx * gy);                   // ~one CTA per SM
  const int64_t max_splits = (p.K + 8 * TN_BK - 1) / (8 * TN_BK);         // at least 8 k-blocks per CTA
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  int kchunk = (int)((p.K + want - 1) / want);
  kchunk = ((kchunk + TN_BK - 1) / TN_BK) * TN_BK;
  p.kchunk = kchunk;
  p.splits = (p.K + kchunk - 1) / kchunk;
  if (p.splits > 1 && !p.accumulate) {
    launch(tn_zero_block_kernel, grid_for(p.M * p.N, 256), 256, 0, s, p.C, p.M, p.N, p.ldc);
    count_launch();
  }
  launch(gemm_tc_tn_ws_kernel<BN>, dim3((unsigned)gx, gy, p.splits), TN_THREADS, L::BYTES, s, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("i3d_gemm(tn): launch failed -> %s", cudaGetErrorString(e));
    return I3D_ERR_CUDA;
  }
  count_launch();
  return I3D_OK;
}

int gemm_tn_ws(int64_t M, int N, const i3d_gemm_seg& sg, float* C, int ldc, int accumulate, cudaStream_t stream) {
  TnParams p;
  memset(&p, 0, sizeof(p));
  p.A = sg.A, p.B = sg.B, p.a_idx = sg.a_idx, p.b_idx = sg.b_idx, p.scale = sg.scale;
  p.lda = sg.lda, p.ldb = sg.ldb, p.K = sg.K;
  p.M = M, p.N = N, p.C = C, p.ldc = ldc, p.accumulate = accumulate;
  if (N <= 32) return launch_tn<32>(p, stream);
  if (N <= 64) return launch_tn<64>(p, stream);
  if (N <= 128) return launch_tn<128>(p, stream);
  if (N <= 224) return launch_tn<224>(p, stream);
  if ((N + 223) / 224 <= (N + 255) / 256) return launch_tn<224>(p, stream);     // same tile count, less padding
  return launch_tn<256>(p, stream);
}

}  // namespace i3d
