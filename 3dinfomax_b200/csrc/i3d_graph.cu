// Graph-structure kernels: stable CSR build (counting sort by destination), graph_ptr, degree scalers.
// Replaces DGL's host-side degree bucketing behind update_all (models/pna.py:206, models/net3d.py:109).
// Integer work: results are BIT EXACT against argsort(key, stable) (oracle/oracle.py: csr_reference).
#include <math.h>

#include "i3d_common.cuh"

namespace i3d {

template <typename K>
__global__ void csr_count_kernel(const K* __restrict__ key, int64_t E, int64_t N, int32_t* __restrict__ rowptr) {
  pdl_grid_sync();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = (int64_t)key[e];
    if (k >= 0 && k < N) atomicAdd(&rowptr[k + 1], 1);
  }
}

// in-place inclusive scan of a[0..n) by ONE CTA (chunked, carry kept in shared memory)
__global__ void scan_inplace_kernel(int32_t* __restrict__ a, int64_t n) {
  pdl_grid_sync();
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + tid;
    int32_t v = i < n ? a[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
      int32_t t = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int32_t incl = v + (w > 0 ? warp_tot[w - 1] : 0) + carry_s;
    if (i < n) a[i] = incl;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
}

template <typename K>
__global__ void csr_fill_kernel(const K* __restrict__ key, int64_t E, int64_t N, const int32_t* __restrict__ rowptr,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ eid) {
  pdl_grid_sync();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = (int64_t)key[e];
    if (k >= 0 && k < N) {
      int32_t pos = atomicAdd(&cursor[k], 1);
      eid[rowptr[k] + pos] = (int32_t)e;
    }
  }
}

// ascending edge id inside every row == stable sort.  Rows are short (bond graphs: D<=4..6;
// complete graphs: D = n-1 < 100) and arrive nearly sorted, so a per-row insertion sort is enough.
__global__ void csr_rowsort_kernel(const int32_t* __restrict__ rowptr, int64_t N, int32_t* __restrict__ eid) {
  pdl_grid_sync();
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < N; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = rowptr[v], e = rowptr[v + 1];
    for (int32_t i = b + 1; i < e; ++i) {
      int32_t x = eid[i];
      int32_t j = i - 1;
      while (j >= b && eid[j] > x) {
        eid[j + 1] = eid[j];
        --j;
      }
      eid[j + 1] = x;
    }
  }
}

template <typename K>
__global__ void csr_finalize_kernel(const K* __restrict__ key, const K* __restrict__ other, int64_t E,
                                    const int32_t* __restrict__ eid, int32_t* __restrict__ col,
                                    int32_t* __restrict__ rowid) {
  pdl_grid_sync();
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < E; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t e = eid[k];
    if (col) col[k] = (int32_t)other[e];
    if (rowid) rowid[k] = (int32_t)key[e];
  }
}

template <typename K>
static int csr_build_impl(const K* key, const K* other, int64_t E, int64_t N, int32_t* rowptr, int32_t* col,
                          int32_t* rowid, int32_t* eid, int32_t* cursor_ws, void* stream, const char* fn) {
  if (!(E >= 0 && N >= 0 && E < (1ll << 31) && N < (1ll << 31) - 1) || !rowptr || !cursor_ws ||
      (E > 0 && (!key || !eid)) || (col && !other)) {
    set_error("%s: invalid argument", fn);
    return I3D_ERR_INVALID;
  }
  cudaStream_t s = as_stream(stream);
  I3D_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(N + 1) * sizeof(int32_t), s));
  if (N > 0) I3D_CUDA(cudaMemsetAsync(cursor_ws, 0, (size_t)N * sizeof(int32_t), s));
  if (E > 0) {
    launch(csr_count_kernel<K>, grid_for(E, 256), 256, 0, s, key, E, N, rowptr);
    I3D_LAUNCHED();
  }
  if (N > 0) {
    launch(scan_inplace_kernel, 1, 1024, 0, s, rowptr + 1, N);
    I3D_LAUNCHED();
  }
  if (E > 0) {
    launch(csr_fill_kernel<K>, grid_for(E, 256), 256, 0, s, key, E, N, rowptr, cursor_ws, eid);
    I3D_LAUNCHED();
    launch(csr_rowsort_kernel, grid_for(N, 128), 128, 0, s, rowptr, N, eid);
    I3D_LAUNCHED();
    if (col || rowid) {
      launch(csr_finalize_kernel<K>, grid_for(E, 256), 256, 0, s, key, other, E, eid, col, rowid);
      I3D_LAUNCHED();
    }
  }
  return I3D_OK;
}

__global__ void counts_to_i32_kernel(const int64_t* __restrict__ counts, int64_t B, int32_t* __restrict__ ptr) {
  pdl_grid_sync();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= B; i += (int64_t)gridDim.x * blockDim.x)
    ptr[i] = i == 0 ? 0 : (int32_t)counts[i - 1];
}

__global__ void degree_scalers_kernel(const int32_t* __restrict__ rowptr, int64_t N, float* __restrict__ amp,
                                      float* __restrict__ att) {
  pdl_grid_sync();
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < N; v += (int64_t)gridDim.x * blockDim.x) {
    const int d = rowptr[v + 1] - rowptr[v];
    if (d <= 0) {
      amp[v] = 0.f;
      att[v] = 0.f;
    } else {
      // numpy evaluates np.log(D + 1) in float64; torch applies it as an fp32 scalar (models/pna.py:61-68)
      const double l = log((double)d + 1.0);
      amp[v] = (float)l;
      att[v] = (float)(1.0 / l);
    }
  }
}

}  // namespace i3d

extern "C" {

int i3d_csr_build(const int64_t* key, const int64_t* other, int64_t E, int64_t N, int32_t* rowptr, int32_t* col,
                  int32_t* rowid, int32_t* eid, int32_t* cursor_ws, void* stream) {
  return i3d::csr_build_impl<int64_t>(key, other, E, N, rowptr, col, rowid, eid, cursor_ws, stream, __func__);
}

int i3d_csr_build_i32(const int32_t* key, const int32_t* other, int64_t E, int64_t N, int32_t* rowptr, int32_t* col,
                      int32_t* rowid, int32_t* eid, int32_t* cursor_ws, void* stream) {
  return i3d::csr_build_impl<int32_t>(key, other, E, N, rowptr, col, rowid, eid, cursor_ws, stream, __func__);
}

int i3d_segment_ptr(const int64_t* counts, int64_t B, int32_t* ptr, void* stream) {
  I3D_REQUIRE(B >= 0 && ptr && (B == 0 || counts), "invalid argument");
  cudaStream_t s = i3d::as_stream(stream);
  i3d::launch(i3d::counts_to_i32_kernel, i3d::grid_for(B + 1, 256), 256, 0, s, counts, B, ptr);
  I3D_LAUNCHED();
  if (B > 0) {
    i3d::launch(i3d::scan_inplace_kernel, 1, 1024, 0, s, ptr + 1, B);
    I3D_LAUNCHED();
  }
  return I3D_OK;
}

int i3d_degree_scalers(const int32_t* rowptr, int64_t N, float* amp, float* att, void* stream) {
  I3D_REQUIRE(N >= 0 && rowptr && (N == 0 || (amp && att)), "invalid argument");
  if (N == 0) return I3D_OK;
  i3d::launch(i3d::degree_scalers_kernel, i3d::grid_for(N, 256), 256, 0, i3d::as_stream(stream), rowptr, N, amp, att);
  I3D_LAUNCHED();
  return I3D_OK;
}
}
