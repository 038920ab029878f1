// Tiny fixed-width vector helper: 16-byte (float4) global accesses when the feature width,
// leading dimensions and base pointers allow it, scalar otherwise.  HBM-bound kernels in this
// library map one thread to one (row, column-group) pair so that a warp touches 512 contiguous bytes.
#pragma once
#include "i3d_common.cuh"

namespace i3d {

template <int V>
struct Vec {
  float v[V];
  __device__ __forceinline__ void load(const float* p) {
    if constexpr (V == 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = __ldg(p + i);
    }
  }
  // plain (coherent) load, e.g. from shared memory
  __device__ __forceinline__ void load_shared(const float* p) {
    if constexpr (V == 4) {
      float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = p[i];
    }
  }
  __device__ __forceinline__ void store(float* p) const {
    if constexpr (V == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) p[i] = v[i];
    }
  }
  // streaming store: written once, consumed by a later kernel
  __device__ __forceinline__ void store_cs(float* p) const {
    if constexpr (V == 4) {
      __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) p[i] = v[i];
    }
  }
  __device__ __forceinline__ void fill(float x) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = x;
  }
};

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// true when every listed pointer is 16-byte aligned and every listed width is a multiple of 4
static inline bool can_vec4(std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> widths) {
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return false;
  for (int64_t w : widths)
    if (w % 4 != 0) return false;
  return true;
}

}  // namespace i3d
