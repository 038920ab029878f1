// micro-benchmark: FP64 vs FP32 FMA issue rate per SM on this part (decides how column statistics are accumulated)
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_chain(T* out, int iters) {
  T a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T b = (T)1.0000001, c = (T)0.5;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
    a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T>
double run(const char* name) {
  T* out;
  cudaMalloc(&out, sizeof(T) * 148 * 4 * 1024);
  const int iters = 4096;
  fma_chain<T><<<148 * 4, 1024>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  cudaEventRecord(a);
  fma_chain<T><<<148 * 4, 1024>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double fma = 148.0 * 4 * 1024 * 8.0 * iters;
  printf("%s: %.3f ms, %.1f GFMA/s = %.2f TFLOP/s, %.1f FMA/clk/SM at 1.965 GHz\n", name, ms, fma / ms / 1e6,
         2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
  return ms;
}
int main() {
  run<float>("fp32");
  run<double>("fp64");
  return 0;
}
