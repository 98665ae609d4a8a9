// Micro-benchmark 2: hardware floor of tcgen05.mma / tcgen05.commit issue from one thread — straight-line code
// (everything a compile-time constant), so the scalar instruction stream around the MMAs is minimal.
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace rs;

template <int N, int CE, int NC>
__global__ void __launch_bounds__(128) k(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc_dyn(&tmem_slot, 512u); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (lane == 0) {
      const uint64_t a0 = umma_desc_sw128(smem_u32(smem));
      const uint64_t b0 = umma_desc_sw128(smem_u32(smem) + 32768);
      t0 = clock64();
#pragma unroll 1
      for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base + (i & 1) * 256, a0 + ((i & 1) * 1024 + 2 * kk), b0 + ((i & 1) * 1024 + 2 * kk), idesc, 1u);
          if (CE && (i % CE) == CE - 1) {
#pragma unroll
            for (int c = 0; c < NC; ++c) umma_commit(&bar[1 + ((i / CE + c) & 3)]);
          }
        }
      }
      t1 = clock64();
      umma_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    if (lane == 0) { t2 = clock64(); out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_dyn(tmem_base, 512u); }
}

template <int N, int CE, int NC>
void run(long long* out) {
  cudaFuncSetAttribute(k<N, CE, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, CE, NC><<<1, 128, 100 * 1024>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("N=%3d  %d commit(s) every %d k-blocks: issue %6.1f cycles/k-block, complete %6.1f cycles/k-block (tensor nominal %d)  %s\n", N, NC, CE,
         (double)out[0] / 64, (double)out[1] / 64, 2 * N, cudaGetErrorString(e));
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  run<64, 0, 1>(out); run<128, 0, 1>(out); run<192, 0, 1>(out); run<256, 0, 1>(out);
  run<64, 1, 1>(out); run<128, 1, 1>(out); run<192, 1, 1>(out); run<256, 1, 1>(out);
  run<64, 2, 1>(out); run<128, 2, 1>(out); run<192, 2, 1>(out);
  run<64, 1, 2>(out); run<128, 1, 2>(out); run<192, 1, 2>(out);
  run<64, 2, 2>(out); run<128, 2, 2>(out); run<192, 2, 2>(out);
  return 0;
}
