// Shared helpers for the lib3dinfomax_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/i3d.h"

namespace i3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define I3D_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      ::i3d::set_error("%s: %s", __func__, msg);                 \
      return I3D_ERR_INVALID;                                    \
    }                                                            \
  } while (0)

#define I3D_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::i3d::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
      return I3D_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

#define I3D_LAUNCHED()                                                          \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      ::i3d::set_error("%s: launch failed -> %s", __func__, cudaGetErrorString(e__)); \
      return I3D_ERR_CUDA;                                                      \
    }                                                                           \
    ::i3d::count_launch();                                                      \
  } while (0)

// grid for a grid-stride elementwise kernel: enough CTAs to cover `work` items, capped at a
// multiple of the SM count (148 on B200) so the tail wave is full.
static inline int grid_for(int64_t work, int threads, int ctas_per_sm = 8) {
  int64_t need = (work + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float act_apply(float y, int act) {
  if (act == I3D_ACT_RELU) return y > 0.f ? y : 0.f;
  if (act == I3D_ACT_SILU) return y / (1.f + expf(-y));
  return y;
}
// d act(y) / dy
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == I3D_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == I3D_ACT_SILU) {
    float s = 1.f / (1.f + expf(-y));
    return s * (1.f + y * (1.f - s));
  }
  return 1.f;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace i3d
