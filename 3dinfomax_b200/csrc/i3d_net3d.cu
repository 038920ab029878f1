// Net3D element-wise kernels (hidden width 20 in every shipped config: pure HBM streams, no tensor cores):
// fourier distance encoding (commons/utils.py:103-110), the sigmoid soft-edge gate
// (models/net3d.py:117-118), node-embedding broadcast (models/net3d.py:61) and a plain add.
#include <initializer_list>

#include "i3d_vec.cuh"

namespace i3d {

__global__ void fourier_encode_kernel(const float* __restrict__ dist, const int32_t* __restrict__ perm, int64_t E,
                                      int k, float* __restrict__ out) {
  pdl_grid_sync();
  const int W = 2 * k + 1;
  const int64_t total = E * W;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / W;
    const int j = (int)(t - r * W);
    const float d = __ldg(dist + (perm ? (int64_t)perm[r] : r));
    float o;
    if (j == 2 * k) {
      o = d;
    } else {
      const int p = j < k ? j : j - k;
      const float x = d / (float)(1 << p);  // scales = 2**arange(k): exact power-of-two division
      o = j < k ? sinf(x) : cosf(x);
    }
    out[t] = o;
  }
}

// one thread per edge row; H is small (20) so a row is 80 contiguous bytes
__global__ void soft_gate_fwd_kernel(const float* __restrict__ msg, int64_t E, int H, const float* __restrict__ ws,
                                     const float* __restrict__ bs, float* __restrict__ m, float* __restrict__ w) {
  pdl_grid_sync();
  extern __shared__ float s_ws[];
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_ws[i] = ws[i];
  __syncthreads();
  const float b = bs[0];
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < E; r += (int64_t)gridDim.x * blockDim.x) {
    const float* p = msg + r * H;
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc = fmaf(__ldg(p + h), s_ws[h], acc);
    acc += b;
    const float g = 1.f / (1.f + expf(-acc));
    w[r] = g;
    float* o = m + r * H;
    for (int h = 0; h < H; ++h) o[h] = __ldg(p + h) * g;
  }
}

__global__ void soft_gate_bwd_kernel(const float* __restrict__ gm, const float* __restrict__ msg,
                                     const float* __restrict__ w, int64_t E, int H, const float* __restrict__ ws,
                                     float* __restrict__ gmsg, float* __restrict__ gws, float* __restrict__ gbs) {
  pdl_grid_sync();
  extern __shared__ float s_ws[];  // ws[H] then block partials gws[H], gbs
  float* s_acc = s_ws + H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_ws[i] = ws[i];
  for (int i = threadIdx.x; i <= H; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // every thread of a warp iterates the same number of times (warp-level reductions inside the loop)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (E + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t r = it * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool ok = r < E;
    float t = 0.f, g = 0.f;
    const float* pm = msg + (ok ? r : 0) * H;
    const float* pg = gm + (ok ? r : 0) * H;
    if (ok) {
      float gw = 0.f;
      for (int h = 0; h < H; ++h) gw = fmaf(__ldg(pg + h), __ldg(pm + h), gw);
      g = w[r];
      t = gw * g * (1.f - g);
      float* o = gmsg + r * H;
      for (int h = 0; h < H; ++h) o[h] = __ldg(pg + h) * g + t * s_ws[h];
    }
    for (int h = 0; h < H; ++h) {
      const float c = warp_sum(ok ? t * __ldg(pm + h) : 0.f);
      if (lane == 0) atomicAdd(&s_acc[h], c);
    }
    const float cb = warp_sum(t);
    if (lane == 0) atomicAdd(&s_acc[H], cb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(gws + i, s_acc[i]);
  if (threadIdx.x == 0) atomicAdd(gbs, s_acc[H]);
}

__global__ void broadcast_rows_kernel(const float* __restrict__ vec, int64_t M, int F, float* __restrict__ out) {
  pdl_grid_sync();
  const int64_t total = M * F;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    out[t] = __ldg(vec + (t % F));
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                           float* __restrict__ y) {
  pdl_grid_sync();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __fadd_rn(a[i], b[i]);
}

// y[m, :] = a[m, :] + b[m, :] for row-major matrices with their own leading dimensions (V floats per thread)
template <int V>
__global__ void add_rows_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int64_t M,
                                int F, float* __restrict__ y, int ldy) {
  pdl_grid_sync();
  const int FV = F / V;
  const int64_t total = M * FV;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / FV;
    const int c = (int)(t - m * FV) * V;
    if (V == 4) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(a + m * lda + c));
      const float4 w = __ldg(reinterpret_cast<const float4*>(b + m * ldb + c));
      *reinterpret_cast<float4*>(y + m * ldy + c) =
          make_float4(__fadd_rn(u.x, w.x), __fadd_rn(u.y, w.y), __fadd_rn(u.z, w.z), __fadd_rn(u.w, w.w));
    } else {
      y[m * ldy + c] = __fadd_rn(a[m * lda + c], b[m * ldb + c]);
    }
  }
}

}  // namespace i3d

using namespace i3d;

extern "C" {

int i3d_fourier_encode(const float* dist, const int32_t* perm, int64_t E, int k, float* out, void* stream) {
  I3D_REQUIRE(E >= 0 && k >= 0 && k <= 16 && (E == 0 || (dist && out)), "invalid argument");
  if (E == 0) return I3D_OK;
  launch(fourier_encode_kernel, grid_for(E * (2 * k + 1), 256), 256, 0, as_stream(stream), dist, perm, E, k, out);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_soft_gate_fwd(const float* msg, int64_t E, int H, const float* ws, const float* bs, float* m, float* w,
                      void* stream) {
  I3D_REQUIRE(E >= 0 && H > 0 && H <= 1024 && ws && bs && (E == 0 || (msg && m && w)), "invalid argument");
  if (E == 0) return I3D_OK;
  launch(soft_gate_fwd_kernel, grid_for(E, 128), 128, sizeof(float) * H, as_stream(stream), msg, E, H, ws, bs, m, w);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_soft_gate_bwd(const float* gm, const float* msg, const float* w, int64_t E, int H, const float* ws,
                      float* gmsg, float* gws, float* gbs, void* stream) {
  I3D_REQUIRE(E >= 0 && H > 0 && H <= 1024 && ws && gws && gbs && (E == 0 || (gm && msg && w && gmsg)),
              "invalid argument");
  if (E == 0) return I3D_OK;
  launch(soft_gate_bwd_kernel, grid_for(E, 128, 4), 128, sizeof(float) * (2 * H + 1), as_stream(stream), 
      gm, msg, w, E, H, ws, gmsg, gws, gbs);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_broadcast_rows(const float* vec, int64_t M, int F, float* out, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && vec && (M == 0 || out), "invalid argument");
  if (M == 0) return I3D_OK;
  launch(broadcast_rows_kernel, grid_for(M * F, 256), 256, 0, as_stream(stream), vec, M, F, out);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_add_rows(const float* a, int lda, const float* b, int ldb, int64_t M, int F, float* y, int ldy, void* stream) {
  I3D_REQUIRE(M >= 0 && F > 0 && lda >= F && ldb >= F && ldy >= F && (M == 0 || (a && b && y)), "invalid argument");
  if (M == 0) return I3D_OK;
  const bool v4 = !(F & 3) && !(lda & 3) && !(ldb & 3) && !(ldy & 3) &&
                  !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15u);
  if (v4)
    launch(add_rows_kernel<4>, grid_for(M * (F / 4), 256), 256, 0, as_stream(stream), a, lda, b, ldb, M, F, y, ldy);
  else
    launch(add_rows_kernel<1>, grid_for(M * F, 256), 256, 0, as_stream(stream), a, lda, b, ldb, M, F, y, ldy);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_add(const float* a, const float* b, int64_t n, float* y, void* stream) {
  I3D_REQUIRE(n >= 0 && (n == 0 || (a && b && y)), "invalid argument");
  if (n == 0) return I3D_OK;
  launch(add_kernel, grid_for(n, 256), 256, 0, as_stream(stream), a, b, n, y);
  I3D_LAUNCHED();
  return I3D_OK;
}
}
