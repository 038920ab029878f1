// Segmented "virtual operand" GEMM, fp32 SIMT path.
//
//   C[m,n] (+)= bias[n] + sum_s sum_k  scale_s(.) * A_s(gather) * B_s(gather)
//
// The K dimension is a list of segments so the reference's torch.cat([h[src], h[dst], e]) (models/pna.py:249),
// torch.cat([h, agg, agg*amp, agg*att]) (models/pna.py:207,232) and the index_select gathers behind DGL's
// edges.src / edges.dst never hit HBM: the tile loader gathers rows and applies the per-row degree scaler
// while staging tiles into shared memory.
//
// Tiling: 128x64x16 CTA tile, 256 threads, 8x4 register micro-tile, double-buffered shared memory with
// register prefetch.  TN (weight-gradient) mode splits K over gridDim.z and reduces with fp32 atomics
// because M,N are a few hundred while K is the edge / node count.
#include "i3d_common.cuh"

namespace i3d {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;

struct GemmParams {
  i3d_gemm_seg seg[4];
  int n_seg;
  int64_t M;
  int N;
  float* C;
  int ldc;
  const float* bias;
  int accumulate;
  int kchunk;  // TN split-K chunk (multiple of BK)
  int splits;
};

__device__ __forceinline__ bool is_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(const __grid_constant__ GemmParams p) {
  pdl_grid_sync();
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t M = p.M;
  const int N = p.N;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < p.n_seg; ++s) {
    const float* __restrict__ A = p.seg[s].A;
    const float* __restrict__ B = p.seg[s].B;
    const int32_t* __restrict__ a_idx = p.seg[s].a_idx;
    const int32_t* __restrict__ b_idx = p.seg[s].b_idx;
    const float* __restrict__ scale = p.seg[s].scale;
    const int lda = p.seg[s].lda, ldb = p.seg[s].ldb;
    int kbeg = 0, kend = p.seg[s].K;
    if (MODE == I3D_GEMM_TN) {
      kbeg = blockIdx.z * p.kchunk;
      kend = min(kend, kbeg + p.kchunk);
    }
    const int ntile = (kend - kbeg + BK - 1) / BK;
    if (ntile <= 0) continue;
    const bool vecA = ((lda & 3) == 0) && is_al16(A);
    const bool vecB = ((ldb & 3) == 0) && is_al16(B);

    // per-thread fixed row bookkeeping for k-contiguous operands
    int64_t a_row[2] = {-1, -1};
    float a_sc[2] = {1.f, 1.f};
    int64_t b_row = -1;
    if (MODE != I3D_GEMM_TN) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t gm = m0 + ((tid + i * GEMM_THREADS) >> 2);
        if (gm < M) {
          a_row[i] = a_idx ? (int64_t)__ldg(a_idx + gm) : gm;
          if (scale) a_sc[i] = __ldg(scale + gm);
        }
      }
    }
    if (MODE == I3D_GEMM_NT) {
      const int gn = n0 + (tid >> 2);
      if (gn < N) b_row = gn;
    }

    float4 ra[2], rb;

    auto load_tile = [&](int k0) {
      // ---- A ----
      if (MODE != I3D_GEMM_TN) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int kq = ((tid + i * GEMM_THREADS) & 3) << 2;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a_row[i] >= 0) {
            const float* src = A + a_row[i] * lda + k0 + kq;
            if (vecA && k0 + kq + 3 < kend) {
              v = __ldg(reinterpret_cast<const float4*>(src));
            } else {
              if (k0 + kq + 0 < kend) v.x = __ldg(src + 0);
              if (k0 + kq + 1 < kend) v.y = __ldg(src + 1);
              if (k0 + kq + 2 < kend) v.z = __ldg(src + 2);
              if (k0 + kq + 3 < kend) v.w = __ldg(src + 3);
            }
            const float sc = a_sc[i];
            v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
          }
          ra[i] = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = tid + i * GEMM_THREADS;
          const int kk = c >> 5, mq = (c & 31) << 2;
          const int gk = k0 + kk;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          const int64_t row = gk < kend ? (a_idx ? (int64_t)__ldg(a_idx + gk) : (int64_t)gk) : -1;
          if (row >= 0) {                                  // negative gather index (padding row) = zero row
            const float* src = A + row * lda + m0 + mq;
            if (vecA && m0 + mq + 3 < M) {
              v = __ldg(reinterpret_cast<const float4*>(src));
            } else {
              if (m0 + mq + 0 < M) v.x = __ldg(src + 0);
              if (m0 + mq + 1 < M) v.y = __ldg(src + 1);
              if (m0 + mq + 2 < M) v.z = __ldg(src + 2);
              if (m0 + mq + 3 < M) v.w = __ldg(src + 3);
            }
            if (scale) {
              const float sc = __ldg(scale + gk);
              v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
            }
          }
          ra[i] = v;
        }
      }
      // ---- B ----
      if (MODE == I3D_GEMM_NT) {
        const int kq = (tid & 3) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b_row >= 0) {
          const float* src = B + b_row * ldb + k0 + kq;
          if (vecB && k0 + kq + 3 < kend) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (k0 + kq + 0 < kend) v.x = __ldg(src + 0);
            if (k0 + kq + 1 < kend) v.y = __ldg(src + 1);
            if (k0 + kq + 2 < kend) v.z = __ldg(src + 2);
            if (k0 + kq + 3 < kend) v.w = __ldg(src + 3);
          }
        }
        rb = v;
      } else {
        const int kk = tid >> 4, nq = (tid & 15) << 2;
        const int gk = k0 + kk;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t row = gk < kend ? ((MODE == I3D_GEMM_TN && b_idx) ? (int64_t)__ldg(b_idx + gk) : (int64_t)gk) : -1;
        if (row >= 0) {
          const float* src = B + row * ldb + n0 + nq;
          if (vecB && n0 + nq + 3 < N) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (n0 + nq + 0 < N) v.x = __ldg(src + 0);
            if (n0 + nq + 1 < N) v.y = __ldg(src + 1);
            if (n0 + nq + 2 < N) v.z = __ldg(src + 2);
            if (n0 + nq + 3 < N) v.w = __ldg(src + 3);
          }
        }
        rb = v;
      }
    };

    auto store_tile = [&](int buf) {
      if (MODE != I3D_GEMM_TN) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = tid + i * GEMM_THREADS;
          const int row = c >> 2, kq = (c & 3) << 2;
          As[buf][kq + 0][row] = ra[i].x;
          As[buf][kq + 1][row] = ra[i].y;
          As[buf][kq + 2][row] = ra[i].z;
          As[buf][kq + 3][row] = ra[i].w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = tid + i * GEMM_THREADS;
          const int kk = c >> 5, mq = (c & 31) << 2;
          *reinterpret_cast<float4*>(&As[buf][kk][mq]) = ra[i];
        }
      }
      if (MODE == I3D_GEMM_NT) {
        const int row = tid >> 2, kq = (tid & 3) << 2;
        Bs[buf][kq + 0][row] = rb.x;
        Bs[buf][kq + 1][row] = rb.y;
        Bs[buf][kq + 2][row] = rb.z;
        Bs[buf][kq + 3][row] = rb.w;
      } else {
        const int kk = tid >> 4, nq = (tid & 15) << 2;
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb;
      }
    };

    load_tile(kbeg);
    store_tile(0);
    __syncthreads();
    for (int t = 0; t < ntile; ++t) {
      const int cur = t & 1;
      if (t + 1 < ntile) load_tile(kbeg + (t + 1) * BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8 + 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (t + 1 < ntile) store_tile(cur ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const bool split = (MODE == I3D_GEMM_TN) && p.splits > 1;
  const int n = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias && !(split && blockIdx.z != 0)) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) bv[j] = __ldg(p.bias + n + j);
  }
  const bool vecC = ((p.ldc & 3) == 0) && is_al16(p.C) && (n + 3 < N);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
    float* c = p.C + m * p.ldc + n;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = acc[i][j] + bv[j];
    if (split) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) atomicAdd(c + j, o[j]);
    } else if (vecC) {
      float4 v = make_float4(o[0], o[1], o[2], o[3]);
      if (p.accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(c);
        v.x += old.x, v.y += old.y, v.z += old.z, v.w += old.w;
      }
      *reinterpret_cast<float4*>(c) = v;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) c[j] = p.accumulate ? c[j] + o[j] : o[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Small problems (M * N <= 64K outputs, K <= 4096): one thread per output element.  The tiled kernel above needs
// >= 128 x 64 outputs per CTA: the 60-row bond-feature tables of the factored edge layer (T = combo W_e^T, its two
// gradients) gave it 1-4 CTAs and ~20 us of pure latency each, 22 launches per step; here they spread over every SM.
// Operands are L1/L2 resident at these sizes, so the per-thread dot product streams float4s of its own B row (NT) or
// coalesced B columns (NN / TN) against a warp-broadcast A element.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) gemm_small_kernel(const __grid_constant__ GemmParams p) {
  pdl_grid_sync();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t M = p.M;
  const int N = p.N;
  if (t >= M * N) return;
  const int64_t m = t / N;
  const int n = (int)(t - m * N);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  for (int s = 0; s < p.n_seg; ++s) {
    const float* __restrict__ A = p.seg[s].A;
    const float* __restrict__ B = p.seg[s].B;
    const int32_t* __restrict__ a_idx = p.seg[s].a_idx;
    const int32_t* __restrict__ b_idx = p.seg[s].b_idx;
    const float* __restrict__ scale = p.seg[s].scale;
    const int lda = p.seg[s].lda, ldb = p.seg[s].ldb, K = p.seg[s].K;
    if (MODE == I3D_GEMM_TN) {
      for (int k = 0; k < K; ++k) {
        const int64_t ra = a_idx ? (int64_t)__ldg(a_idx + k) : (int64_t)k;
        const int64_t rb = b_idx ? (int64_t)__ldg(b_idx + k) : (int64_t)k;
        if (ra < 0 || rb < 0) continue;
        float a = __ldg(A + ra * lda + m);
        if (scale) a *= __ldg(scale + k);
        acc0 = fmaf(a, __ldg(B + rb * ldb + n), acc0);
      }
    } else {
      const int64_t ra = a_idx ? (int64_t)__ldg(a_idx + m) : m;
      if (ra < 0) continue;
      const float sc = scale ? __ldg(scale + m) : 1.f;
      const float* __restrict__ ap = A + ra * lda;
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
      if (MODE == I3D_GEMM_NT) {
        const float* __restrict__ bp = B + (int64_t)n * ldb;
        const bool vec = ((lda | ldb) & 3) == 0 && is_al16(A) && is_al16(B);
        int k = 0;
        if (vec) {
          for (; k + 4 <= K; k += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(ap + k));
            const float4 b = __ldg(reinterpret_cast<const float4*>(bp + k));
            p0 = fmaf(a.x, b.x, p0), p1 = fmaf(a.y, b.y, p1), p2 = fmaf(a.z, b.z, p2), p3 = fmaf(a.w, b.w, p3);
          }
        }
        for (; k < K; ++k) p0 = fmaf(__ldg(ap + k), __ldg(bp + k), p0);
      } else {
        int k = 0;
        for (; k + 4 <= K; k += 4) {
          p0 = fmaf(__ldg(ap + k), __ldg(B + (int64_t)k * ldb + n), p0);
          p1 = fmaf(__ldg(ap + k + 1), __ldg(B + (int64_t)(k + 1) * ldb + n), p1);
          p2 = fmaf(__ldg(ap + k + 2), __ldg(B + (int64_t)(k + 2) * ldb + n), p2);
          p3 = fmaf(__ldg(ap + k + 3), __ldg(B + (int64_t)(k + 3) * ldb + n), p3);
        }
        for (; k < K; ++k) p0 = fmaf(__ldg(ap + k), __ldg(B + (int64_t)k * ldb + n), p0);
      }
      const float part = (p0 + p1) + (p2 + p3);
      acc1 = fmaf(sc, part, acc1);
    }
  }
  float o = (acc0 + acc1) + (acc2 + acc3);
  if (p.bias) o += __ldg(p.bias + n);
  float* c = p.C + m * p.ldc + n;
  *c = p.accumulate ? *c + o : o;
}

__global__ void zero_block_kernel(float* __restrict__ C, int64_t M, int N, int ldc) {
  pdl_grid_sync();
  const int64_t total = M * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / N;
    C[m * ldc + (t - m * N)] = 0.f;
  }
}

}  // namespace i3d

namespace i3d {
bool gemm_tc_eligible(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs);
int gemm_tc(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
            int accumulate, void* ws, size_t ws_bytes, double* stats, int stats_act, cudaStream_t stream,
            const int32_t* m_valid);
size_t gemm_tc_ws_bytes(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs);

// out[c, r] = in[r, c]  (32x32 shared-memory tiles, coalesced on both sides)
__global__ void transpose_kernel(const float* __restrict__ in, int64_t rows, int cols, int ld_in,
                                 float* __restrict__ out, int ld_out) {
  pdl_grid_sync();
  __shared__ float t[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? in[r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(int64_t)c * ld_out + r] = t[threadIdx.x][i];
  }
}
static int g_gemm_backend = 0;   // 0: tensor cores (tcgen05) where eligible, fp32 SIMT otherwise; 1: SIMT only;
                                 // 2: as 0 but weight gradients (TN) through the MN-major descriptor kernel
extern bool g_tn_ws;
}  // namespace i3d

using namespace i3d;

extern "C" int i3d_transpose(const float* in, int64_t rows, int cols, int ld_in, float* out, int ld_out, void* stream) {
  I3D_REQUIRE(rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= rows && (rows * cols == 0 || (in && out)),
              "invalid argument");
  if (rows == 0 || cols == 0) return I3D_OK;
  const int64_t gx = (rows + 31) / 32;
  const int gy = (cols + 31) / 32;
  I3D_REQUIRE(gx < (1ll << 31) && gy <= 65535, "matrix too large");
  launch(transpose_kernel, dim3((unsigned)gx, gy), dim3(32, 8), 0, as_stream(stream), in, rows, cols, ld_in, out, ld_out);
  I3D_LAUNCHED();
  return I3D_OK;
}

extern "C" int i3d_gemm_backend(int backend) {
  const int old = g_gemm_backend;
  if (backend >= 0 && backend <= 2) g_gemm_backend = backend, g_tn_ws = backend == 2;
  return old;
}

extern "C" size_t i3d_gemm_ws_bytes(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs) {
  if (g_gemm_backend == 1 || !segs || n_seg < 1 || n_seg > 4 || mode < 0 || mode > 2) return 0;
  return gemm_tc_ws_bytes(mode, M, N, n_seg, segs);
}

namespace i3d {
bool gemm_nt_prepared_ok(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs);
int gemm_prep_describe(int N, int n_seg, const i3d_gemm_seg* segs, int transposed, void* ws, int tile0,
                       i3d_prep_item* out, int* tiles_out);
int gemm_prep_run(const i3d_prep_item* dev_items, int n_items, int total_tiles, cudaStream_t stream);
int gemm_ws_nt(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
               int accumulate, void* ws, double* stats, int stats_act, cudaStream_t stream, bool prepared,
               const int32_t* m_valid);
int gemm_ws_nt_bucketed(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
                        const float* hi, const float* lo, int b_pitch, int n_buckets, const int32_t* tile_bucket,
                        const int32_t* row_map, double* stats, int stats_act, cudaStream_t stream,
                        const int32_t* m_valid);
int gemm_tn_chunked(int64_t M, int N, const i3d_gemm_seg& sg, float* C, int ldc, int64_t c_bucket_stride,
                    const int32_t* chunk_tab, int n_chunks, cudaStream_t stream);
bool gemm_ws_available();
int gemm_ws_debug_read(unsigned long long* out);
}  // namespace i3d

extern "C" int i3d_gemm_debug_counters(unsigned long long* out16) {
  I3D_REQUIRE(out16 != nullptr, "out16 is null");
  return gemm_ws_debug_read(out16);
}

extern "C" int i3d_gemm_nt_bucketed(int64_t Mv, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                                    const float* bias, const float* b_hi, const float* b_lo, int b_pitch,
                                    int n_buckets, const int32_t* tile_bucket, const int32_t* row_map,
                                    double* col_stats, int stats_act, void* stream) {
  return i3d_gemm_nt_bucketed_v(Mv, N, n_seg, segs, C, ldc, bias, b_hi, b_lo, b_pitch, n_buckets, tile_bucket, row_map,
                                col_stats, stats_act, nullptr, stream);
}

extern "C" int i3d_gemm_nt_bucketed_v(int64_t Mv, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                                      const float* bias, const float* b_hi, const float* b_lo, int b_pitch,
                                      int n_buckets, const int32_t* tile_bucket, const int32_t* row_map,
                                      double* col_stats, int stats_act, const int32_t* m_valid, void* stream) {
  I3D_REQUIRE(Mv > 0 && (Mv % 128) == 0 && N >= 16 && (N & 3) == 0 && n_seg >= 1 && n_seg <= 4 && segs && C &&
                  ldc >= N && b_hi && b_lo && n_buckets >= 1 && n_buckets <= 16 && tile_bucket && row_map,
              "invalid argument");
  I3D_REQUIRE(gemm_ws_available(), "cuTensorMapEncodeTiled is not available (needs a CUDA 12 driver)");
  int ktot = 0;
  for (int s = 0; s < n_seg; ++s) ktot += (segs[s].K + 31) / 32 * 32;
  I3D_REQUIRE(b_pitch >= ktot && (b_pitch & 3) == 0, "b_pitch must cover the padded K extent and be a multiple of 4");
  for (int s = 0; s < n_seg; ++s)
    I3D_REQUIRE(segs[s].K > 0 && (segs[s].K & 3) == 0 && (segs[s].lda & 3) == 0 && segs[s].A && !segs[s].b_idx &&
                    !segs[s].scale && (reinterpret_cast<uintptr_t>(segs[s].A) & 15u) == 0,
                "segments must be 16-byte aligned, unscaled, with K a multiple of 4");
  const bool prezeroed = (stats_act & I3D_STATS_PREZEROED) != 0;
  stats_act &= 0xff;
  if (col_stats && !prezeroed) I3D_CUDA(cudaMemsetAsync(col_stats, 0, sizeof(double) * 2 * N * I3D_STATS_STRIDE, as_stream(stream)));
  return gemm_ws_nt_bucketed(Mv, N, n_seg, segs, C, ldc, bias, b_hi, b_lo, b_pitch, n_buckets, tile_bucket, row_map,
                             col_stats, stats_act, as_stream(stream), m_valid);
}

extern "C" int i3d_gemm_tn_chunked(int64_t M, int N, const i3d_gemm_seg* seg, float* C, int ldc,
                                   int64_t c_bucket_stride, const int32_t* chunk_tab, int n_chunks, void* stream) {
  I3D_REQUIRE(M >= 16 && N >= 16 && (M & 3) == 0 && (N & 3) == 0 && seg && C && ldc >= N && (ldc & 3) == 0 &&
                  chunk_tab && n_chunks >= 1 && n_chunks <= 65535 && c_bucket_stride >= 0 && (c_bucket_stride & 3) == 0,
              "invalid argument");
  I3D_REQUIRE(seg->K > 0 && seg->A && seg->B && (seg->lda & 3) == 0 && (seg->ldb & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(seg->A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(seg->B) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(C) & 15u) == 0,
              "operands must be 16-byte aligned with leading dimensions divisible by 4");
  return gemm_tn_chunked(M, N, *seg, C, ldc, c_bucket_stride, chunk_tab, n_chunks, as_stream(stream));
}

extern "C" int i3d_gemm_prep_describe(int N, int n_seg, const i3d_gemm_seg* segs, int transposed, void* ws, int tile0,
                                      i3d_prep_item* items_out, int* tiles_out) {
  I3D_REQUIRE(N > 0 && n_seg >= 1 && n_seg <= 4 && segs && ws && items_out && tiles_out && tile0 >= 0,
              "invalid argument");
  for (int s = 0; s < n_seg; ++s)
    I3D_REQUIRE(segs[s].K > 0 && segs[s].B && segs[s].ldb >= (transposed ? N : segs[s].K), "invalid segment");
  return gemm_prep_describe(N, n_seg, segs, transposed, ws, tile0, items_out, tiles_out);
}

extern "C" int i3d_gemm_prep_run(const i3d_prep_item* dev_items, int n_items, int total_tiles, void* stream) {
  I3D_REQUIRE(n_items >= 0 && total_tiles >= 0 && (n_items == 0 || dev_items), "invalid argument");
  if (n_items == 0 || total_tiles == 0) return I3D_OK;
  return gemm_prep_run(dev_items, n_items, total_tiles, as_stream(stream));
}

extern "C" int i3d_gemm_nt_prepared_ok(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs) {
  return (g_gemm_backend != 1 && segs && gemm_nt_prepared_ok(M, N, n_seg, segs)) ? 1 : 0;
}

extern "C" int i3d_gemm_nt_prepared(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                                    const float* bias, int accumulate, const void* ws, double* col_stats,
                                    int stats_act, void* stream) {
  return i3d_gemm_nt_prepared_v(M, N, n_seg, segs, C, ldc, bias, accumulate, ws, col_stats, stats_act, nullptr, stream);
}

extern "C" int i3d_gemm_nt_prepared_v(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                                      const float* bias, int accumulate, const void* ws, double* col_stats,
                                      int stats_act, const int32_t* m_valid, void* stream) {
  I3D_REQUIRE(M >= 0 && N >= 0 && n_seg >= 1 && n_seg <= 4 && segs && ldc >= N && ws, "invalid shape");
  I3D_REQUIRE(!col_stats || !accumulate, "col_stats needs accumulate == 0");
  if (M == 0 || N == 0) return I3D_OK;
  I3D_REQUIRE(C != nullptr, "C is null");
  I3D_REQUIRE(i3d_gemm_nt_prepared_ok(M, N, n_seg, segs), "shape not eligible for prepared operands");
  const bool prezeroed = (stats_act & I3D_STATS_PREZEROED) != 0;
  stats_act &= 0xff;
  if (col_stats && !prezeroed) I3D_CUDA(cudaMemsetAsync(col_stats, 0, sizeof(double) * 2 * N * I3D_STATS_STRIDE, as_stream(stream)));
  return gemm_ws_nt(M, N, n_seg, segs, C, ldc, bias, accumulate, const_cast<void*>(ws), col_stats, stats_act,
                    as_stream(stream), true, m_valid);
}

extern "C" int i3d_gemm(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                        const float* bias, int accumulate, void* stream) {
  return i3d_gemm_ex(mode, M, N, n_seg, segs, C, ldc, bias, accumulate, nullptr, 0, nullptr, 0, stream);
}

extern "C" int i3d_gemm_ex(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                           const float* bias, int accumulate, void* ws, size_t ws_bytes, double* col_stats,
                           int stats_act, void* stream) {
  return i3d_gemm_ex_v(mode, M, N, n_seg, segs, C, ldc, bias, accumulate, ws, ws_bytes, col_stats, stats_act, nullptr,
                       stream);
}

extern "C" int i3d_gemm_ex_v(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                             const float* bias, int accumulate, void* ws, size_t ws_bytes, double* col_stats,
                             int stats_act, const int32_t* m_valid, void* stream) {
  I3D_REQUIRE(mode >= 0 && mode <= 2, "mode must be NT, NN or TN");
  const bool prezeroed = (stats_act & I3D_STATS_PREZEROED) != 0;
  stats_act &= 0xff;
  I3D_REQUIRE(!col_stats || (mode == I3D_GEMM_NT && !accumulate), "col_stats needs NT mode without accumulate");
  I3D_REQUIRE(M >= 0 && N >= 0 && n_seg >= 1 && n_seg <= 4 && segs && ldc >= N, "invalid shape");
  if (M == 0 || N == 0) return I3D_OK;
  I3D_REQUIRE(C != nullptr, "C is null");
  I3D_REQUIRE(mode != I3D_GEMM_TN || n_seg == 1, "TN mode takes exactly one segment");
  for (int s = 0; s < n_seg; ++s) {
    I3D_REQUIRE(segs[s].K >= 0 && (segs[s].K == 0 || (segs[s].A && segs[s].B)), "segment operand is null");
    I3D_REQUIRE(mode == I3D_GEMM_TN || segs[s].b_idx == nullptr, "b_idx is only valid in TN mode");
  }
  if (g_gemm_backend != 1 && gemm_tc_eligible(mode, M, N, n_seg, segs)) {
    if (col_stats && !prezeroed) I3D_CUDA(cudaMemsetAsync(col_stats, 0, sizeof(double) * 2 * N * I3D_STATS_STRIDE, as_stream(stream)));
    return gemm_tc(mode, M, N, n_seg, segs, C, ldc, bias, accumulate, ws, ws_bytes, col_stats, stats_act,
                   as_stream(stream), m_valid);
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < n_seg; ++s) {
    I3D_REQUIRE(segs[s].K >= 0 && (segs[s].K == 0 || (segs[s].A && segs[s].B)), "segment operand is null");
    I3D_REQUIRE(mode == I3D_GEMM_TN || segs[s].b_idx == nullptr, "b_idx is only valid in TN mode");
    p.seg[s] = segs[s];
  }
  p.n_seg = n_seg;
  p.M = M;
  p.N = N;
  p.C = C;
  p.ldc = ldc;
  p.bias = bias;
  p.accumulate = accumulate;
  p.kchunk = 0;
  p.splits = 1;
  cudaStream_t s = as_stream(stream);
  const int64_t gx = (M + BM - 1) / BM;
  const int gy = (N + BN - 1) / BN;
  I3D_REQUIRE(gx < (1ll << 31) && gy <= 65535, "problem too large");
  int64_t ksum = 0;
  for (int sg = 0; sg < n_seg; ++sg) ksum += segs[sg].K;
  if (M * N <= 65536 && ksum <= 4096 && gx * gy < 32) {
    // too few 128 x 64 tiles to fill the machine: one thread per output element
    const int grid = (int)((M * N + 255) / 256);
    if (mode == I3D_GEMM_TN)
      launch(gemm_small_kernel<I3D_GEMM_TN>, grid, 256, 0, s, p);
    else if (mode == I3D_GEMM_NT)
      launch(gemm_small_kernel<I3D_GEMM_NT>, grid, 256, 0, s, p);
    else
      launch(gemm_small_kernel<I3D_GEMM_NN>, grid, 256, 0, s, p);
    I3D_LAUNCHED();
    if (col_stats && mode == I3D_GEMM_NT) return i3d_act_colstats_v(C, M, N, ldc, stats_act, col_stats, m_valid, nullptr, stream);
    return I3D_OK;
  }
  if (mode == I3D_GEMM_TN) {
    const int K = segs[0].K;
    const int64_t tiles = gx * gy;
    int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
    int64_t max_splits = (K + 63) / 64;
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    int kchunk = (int)((K + want - 1) / want);
    kchunk = ((kchunk + BK - 1) / BK) * BK;
    if (kchunk < BK) kchunk = BK;
    p.kchunk = kchunk;
    p.splits = K > 0 ? (K + kchunk - 1) / kchunk : 1;
    if (p.splits > 1 && !accumulate) {
      launch(zero_block_kernel, grid_for(M * N, 256), 256, 0, s, C, M, N, ldc);
      I3D_LAUNCHED();
    }
    launch(gemm_kernel<I3D_GEMM_TN>, dim3((unsigned)gx, gy, p.splits), GEMM_THREADS, 0, s, p);
  } else if (mode == I3D_GEMM_NT) {
    launch(gemm_kernel<I3D_GEMM_NT>, dim3((unsigned)gx, gy, 1), GEMM_THREADS, 0, s, p);
    if (col_stats) {
      I3D_LAUNCHED();
      return i3d_act_colstats_v(C, M, N, ldc, stats_act, col_stats, m_valid, nullptr, stream);
    }
  } else {
    launch(gemm_kernel<I3D_GEMM_NN>, dim3((unsigned)gx, gy, 1), GEMM_THREADS, 0, s, p);
  }
  I3D_LAUNCHED();
  return I3D_OK;
}
