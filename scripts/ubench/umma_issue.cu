// Micro-benchmark: cost of issuing tcgen05.mma (cta_group::1, M=128, K=16) back to back from one thread, for several N,
// with and without a tcgen05.commit after every 4 MMAs.  Operands are whatever is in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I resshift_b200/csrc scripts/ubench/umma_issue.cu -o resshift_b200/lib/ubench_umma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace rs;

__global__ void __launch_bounds__(128) k(long long* out, int N, int iters, int commit_every, int spread, int ncommit) {
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
      t0 = clock64();
      int nb = 0;
      for (int i = 0; i < iters; ++i) {
        const uint32_t sa = smem_u32(smem) + (spread ? (i & 1) * 16384 : 0);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + 32768);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + (spread ? (i & 1) * 256 : 0), adesc + 2 * kk, bdesc + 2 * kk, idesc, (i | kk) != 0 ? 1u : 0u);
        if (commit_every && (i % commit_every) == commit_every - 1) {
          for (int c = 0; c < ncommit; ++c) { umma_commit(&bar[1 + (nb & 3)]); ++nb; }
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

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 64;   // x 4 MMAs
  const int ces[4] = {0, 1, 2, 4};
  for (int nc = 1; nc <= 2; ++nc)
    for (int ci = 0; ci < 4; ++ci)
      for (int N = 64; N <= 256; N += 64) {
        const int ce = ces[ci];
        if (ce == 0 && nc == 2) continue;
        k<<<1, 128, 100 * 1024>>>(out, N, iters, ce, 1, nc);
        cudaError_t e = cudaDeviceSynchronize();
        printf("N=%3d  %d commit(s) every %d k-blocks (4 MMAs each): issue %6.1f cycles/MMA, complete %6.1f cycles/MMA (tensor nominal %d)  %s\n", N, nc, ce,
               (double)out[0] / (iters * 4), (double)out[1] / (iters * 4), N / 2, cudaGetErrorString(e));
      }
  return 0;
}
