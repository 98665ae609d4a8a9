// Micro-benchmark: instruction-rate of the GELU / SiLU epilogue math, scalar FFMA vs packed FFMA2 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I resshift_b200/csrc scripts/ubench/gelu_rate.cu -o resshift_b200/lib/ubench_gelu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace rs;

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
  float2 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = make_float2(seed + threadIdx.x * 1e-3f + j, seed - j * 0.37f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) { v[j].x = gelu_erf_f(v[j].x) + 0.5f; v[j].y = gelu_erf_f(v[j].y) + 0.5f; }
      if (MODE == 1) { v[j] = gelu_erf_f2(v[j]); v[j].x += 0.5f; v[j].y += 0.5f; }
      if (MODE == 2) { v[j].x = silu_f(v[j].x) + 0.5f; v[j].y = silu_f(v[j].y) + 0.5f; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j].x + v[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 4 * 512 * 4);
  const int iters = 2000;
  k<MODE><<<148 * 4, 512>>>(out, 10, 0.3f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 512>>>(out, iters, 0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double elems = 148.0 * 4 * 512 * 16 * iters;
  printf("%-14s %8.3f ms  %7.2f Gelem/s  -> %.2f elem/clk/SM @1.965GHz   (%s)\n", name, ms, elems / ms / 1e6,
         elems / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("gelu scalar");
  run<1>("gelu packed");
  run<2>("silu scalar");
  return 0;
}
