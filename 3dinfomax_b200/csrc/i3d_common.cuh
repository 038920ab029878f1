// Shared helpers for the lib3dinfomax_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/i3d.h"

namespace i3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

bool pdl_enabled();   // programmatic dependent launch for every kernel of the library (I3D_PDL=0 turns it off)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Every kernel of the library starts with pdl_grid_sync() and is launched through launch(): with programmatic stream
// serialization the next kernel's CTAs are scheduled as soon as all CTAs of the running one have started, and block
// in griddepcontrol.wait until that grid has completed and its writes are visible.  The step is ~300 dependent
// launches of 5-50 us, so the launch/drain gap between them is a measurable share of it.
// Rule: a kernel never touches global memory before pdl_grid_sync(); kernels that are not ours (torch, NCCL, memset
// nodes) carry no attribute and serialise fully on both sides.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through cudaGetLastError()
}

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
  if (act == I3D_ACT_LEAKY_RELU) return y > 0.f ? y : 0.01f * y;
  return y;
}
// d act(y) / dy
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == I3D_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == I3D_ACT_SILU) {
    float s = 1.f / (1.f + expf(-y));
    return s * (1.f + y * (1.f - s));
  }
  if (act == I3D_ACT_LEAKY_RELU) return y > 0.f ? 1.f : 0.01f;
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
