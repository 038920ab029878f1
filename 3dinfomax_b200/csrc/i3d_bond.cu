// Bond-feature tables of the factored edge layer, all PNA layers in one launch.
//
// The `e` K-segment of the edge MLP, e W_e^T (models/pna.py:249-252 with e = BondEncoder(edata['feat']),
// commons/mol_encoder.py:45-73), only takes n_codes = prod(5, 6, 2) = 60 distinct values per layer: T_l = combo W_e,l^T
// with combo[c] the embedding of feature combination c.  These are 60 x F x F products — three of them per layer and
// step (T_l, dW_e,l = dT_l^T combo, dcombo += dT_l W_e,l).  As separate GEMM launches each is pure latency (~15-25 us
// on 1-4 CTAs, 22 launches per step); here each of the three is ONE launch over all layers with the whole K extent
// staged in shared memory by coalesced loads (one round trip to L2), 8 outputs per thread.
#include "i3d_common.cuh"

namespace i3d {

constexpr int kBondMaxLayers = 16;
constexpr int kBondCodes = 64;          // rows of the staged combination tile (n_codes <= 64)

struct BondArgs {
  const float* W[kBondMaxLayers];       // [Fout, >= col0 + F] weights (ld = ldw)
  float* T[kBondMaxLayers];             // fwd: [n_codes, Fout] outputs
  const float* dT[kBondMaxLayers];      // bwd: [n_codes, Fout] (NULL: layer skipped)
  float* dW[kBondMaxLayers];            // bwd: weight-gradient matrices (ld = ldw), += on [:, col0 : col0 + F]
  int L, n_codes, F, Fout, ldw, col0;
};

// T_l[c, n] = sum_k combo[c, k] W_l[n, col0 + k].  grid (ceil(Fout / 32), L), 256 threads, 8 outputs per thread.
__global__ void __launch_bounds__(256) bond_tables_fwd_kernel(const __grid_constant__ BondArgs a,
                                                              const float* __restrict__ combo) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  const int K = a.F, P = K + 1;                  // pitch K + 1: lanes along n hit distinct banks (9 n + k mod 32)
  float* Cs = sm;                                // [64][P]
  float* Ws = sm + kBondCodes * P;               // [32][P]
  const int l = blockIdx.y, n0 = blockIdx.x * 32;
  const float* __restrict__ W = a.W[l];
  for (int t = threadIdx.x; t < kBondCodes * K; t += 256) {
    const int c = t / K, k = t - c * K;
    Cs[c * P + k] = c < a.n_codes ? __ldg(combo + (int64_t)c * K + k) : 0.f;
  }
  for (int t = threadIdx.x; t < 32 * K; t += 256) {
    const int n = t / K, k = t - n * K;
    Ws[n * P + k] = n0 + n < a.Fout ? __ldg(W + (int64_t)(n0 + n) * a.ldw + a.col0 + k) : 0.f;
  }
  __syncthreads();
  const int n = threadIdx.x & 31, c0 = (threadIdx.x >> 5) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* wp = Ws + n * P;
  for (int k = 0; k < K; ++k) {
    const float b = wp[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(Cs[(c0 + i) * P + k], b, acc[i]);     // warp-uniform address: broadcast
  }
  if (n0 + n < a.Fout) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i < a.n_codes) a.T[l][(int64_t)(c0 + i) * a.Fout + n0 + n] = acc[i];
  }
}

// dW_l[n, col0 + k] += sum_c dT_l[c, n] combo[c, k].  grid (ceil(F / 32), ceil(Fout / 64), L).
__global__ void __launch_bounds__(256) bond_tables_dw_kernel(const __grid_constant__ BondArgs a,
                                                             const float* __restrict__ combo) {
  pdl_grid_sync();
  const int l = blockIdx.z;
  if (!a.dT[l] || !a.dW[l]) return;
  __shared__ float Ds[kBondCodes][64 + 1];       // dT tile [c][n]
  __shared__ float Cs[kBondCodes][32 + 1];       // combo tile [c][k]
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 64;
  const float* __restrict__ dT = a.dT[l];
  for (int t = threadIdx.x; t < kBondCodes * 64; t += 256) {
    const int c = t >> 6, n = t & 63;
    Ds[c][n] = (c < a.n_codes && n0 + n < a.Fout) ? __ldg(dT + (int64_t)c * a.Fout + n0 + n) : 0.f;
  }
  for (int t = threadIdx.x; t < kBondCodes * 32; t += 256) {
    const int c = t >> 5, k = t & 31;
    Cs[c][k] = (c < a.n_codes && k0 + k < a.F) ? __ldg(combo + (int64_t)c * a.F + k0 + k) : 0.f;
  }
  __syncthreads();
  const int k = threadIdx.x & 31, nb = (threadIdx.x >> 5) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < a.n_codes; ++c) {
    const float b = Cs[c][k];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(Ds[c][nb + i], b, acc[i]);
  }
  if (k0 + k < a.F) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (n0 + nb + i < a.Fout) a.dW[l][(int64_t)(n0 + nb + i) * a.ldw + a.col0 + k0 + k] += acc[i];
  }
}

// dcombo[c, k] += sum_l sum_n dT_l[c, n] W_l[n, col0 + k].  grid (ceil(F / 32), L); fp32 atomics over the L layers.
__global__ void __launch_bounds__(256) bond_tables_dcombo_kernel(const __grid_constant__ BondArgs a,
                                                                 float* __restrict__ dcombo) {
  pdl_grid_sync();
  const int l = blockIdx.y;
  if (!a.dT[l]) return;
  extern __shared__ float sm[];
  const int NN = a.Fout, P = NN + 1;
  float* Ds = sm;                                // [64][P]   dT_l
  float* Ws = sm + kBondCodes * P;               // [Fout][33] W_l[:, col0 + k0 : +32]
  const int k0 = blockIdx.x * 32;
  const float* __restrict__ dT = a.dT[l];
  const float* __restrict__ W = a.W[l];
  for (int t = threadIdx.x; t < kBondCodes * NN; t += 256) {
    const int c = t / NN, n = t - c * NN;
    Ds[c * P + n] = c < a.n_codes ? __ldg(dT + (int64_t)c * NN + n) : 0.f;
  }
  for (int t = threadIdx.x; t < NN * 32; t += 256) {
    const int n = t >> 5, k = t & 31;
    Ws[n * 33 + k] = k0 + k < a.F ? __ldg(W + (int64_t)n * a.ldw + a.col0 + k0 + k) : 0.f;
  }
  __syncthreads();
  const int k = threadIdx.x & 31, c0 = (threadIdx.x >> 5) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n = 0; n < NN; ++n) {
    const float b = Ws[n * 33 + k];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(Ds[(c0 + i) * P + n], b, acc[i]);
  }
  if (k0 + k < a.F) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i < a.n_codes) atomicAdd(dcombo + (int64_t)(c0 + i) * a.F + k0 + k, acc[i]);
  }
}

static int bond_fill(BondArgs& a, int n_codes, int F, int L, const float* const* W, int ldw, int col0, int Fout) {
  if (!(n_codes >= 1 && n_codes <= kBondCodes && F >= 1 && F <= 512 && Fout >= 1 && Fout <= 512 && L >= 1 &&
        L <= kBondMaxLayers && W && ldw >= col0 + F && col0 >= 0))
    return -1;
  memset(&a, 0, sizeof(a));
  a.L = L, a.n_codes = n_codes, a.F = F, a.Fout = Fout, a.ldw = ldw, a.col0 = col0;
  for (int l = 0; l < L; ++l) {
    if (!W[l]) return -1;
    a.W[l] = W[l];
  }
  return 0;
}

}  // namespace i3d

using namespace i3d;

extern "C" {

int i3d_bond_tables_fwd(const float* combo, int n_codes, int F, int L, const float* const* W, int ldw, int col0,
                        int Fout, float* const* T, void* stream) {
  BondArgs a;
  I3D_REQUIRE(combo && T && bond_fill(a, n_codes, F, L, W, ldw, col0, Fout) == 0, "invalid argument");
  for (int l = 0; l < L; ++l) {
    I3D_REQUIRE(T[l] != nullptr, "T[l] is null");
    a.T[l] = T[l];
  }
  const size_t smem = sizeof(float) * (size_t)(kBondCodes + 32) * (F + 1);
  static size_t configured = 0;
  if (smem > configured) {
    I3D_CUDA(cudaFuncSetAttribute(bond_tables_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch(bond_tables_fwd_kernel, dim3((Fout + 31) / 32, L), 256, smem, as_stream(stream), a, combo);
  I3D_LAUNCHED();
  return I3D_OK;
}

int i3d_bond_tables_bwd(const float* combo, int n_codes, int F, int L, const float* const* W, int ldw, int col0,
                        int Fout, const float* const* dT, float* const* dW, float* dcombo, void* stream) {
  BondArgs a;
  I3D_REQUIRE(combo && dT && bond_fill(a, n_codes, F, L, W, ldw, col0, Fout) == 0, "invalid argument");
  bool any_dw = false;
  for (int l = 0; l < L; ++l) {
    a.dT[l] = dT[l];
    a.dW[l] = dW ? dW[l] : nullptr;
    any_dw = any_dw || (a.dT[l] && a.dW[l]);
  }
  cudaStream_t s = as_stream(stream);
  if (any_dw) {
    launch(bond_tables_dw_kernel, dim3((F + 31) / 32, (Fout + 63) / 64, L), 256, 0, s, a, combo);
    I3D_LAUNCHED();
  }
  if (dcombo) {
    const size_t smem = sizeof(float) * ((size_t)kBondCodes * (Fout + 1) + (size_t)Fout * 33);
    static size_t configured = 0;
    if (smem > configured) {
      I3D_CUDA(cudaFuncSetAttribute(bond_tables_dcombo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    launch(bond_tables_dcombo_kernel, dim3((F + 31) / 32, L), 256, smem, s, a, dcombo);
    I3D_LAUNCHED();
  }
  return I3D_OK;
}
}
