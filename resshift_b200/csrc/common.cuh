// Common device/host helpers for the ResShift B200 (sm_100a) kernels.
// Raw PTX wrappers for mbarrier / TMA / tcgen05 (no CUTLASS dependency).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace rs {

// ----------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry point returns 0 or a negative code; the message is kept
// in a thread-local string readable through rs_last_error().
// ----------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define RS_CUDA_OK(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (void)cudaGetLastError(); /* clear the error so later calls report their own */            \
      return ::rs::fail(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
    }                                                                                          \
  } while (0)

#define RS_CHECK(cond, msg)                                                              \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      return ::rs::fail(-1, std::string("check failed: ") + #cond + " — " + (msg) + " (" + \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
    }                                                                                    \
  } while (0)

#ifdef __CUDACC__

// ----------------------------------------------------------------------------------------
// small device utilities
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: `pdl_trigger` lets the next kernel in the stream start its prologue
// early; `pdl_wait` blocks until every prerequisite grid has completed and its writes are visible.
// Every kernel of this library calls pdl_wait() before its first access to global memory.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// x * sigmoid(x): ex2 + approximate reciprocal (2 ulp), no IEEE-division fix-up sequence (the GroupNorm pass that
// applies it is bound by instruction issue / MUFU, not by memory)
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// exact-erf GELU (nn.GELU default, reference models/swin_transformer.py:18):  0.5 x (1 + erf(x / sqrt 2)).
// erf(z) = sign(z) (1 - 2^P(|z|)) with P a degree-7 minimax-style fit of log2(erfc) on [0, 4] (clamped beyond):
// |erf error| <= 4.3e-6, |GELU error| <= 6.4e-7 over all x — three orders below the fp16 rounding of the stored
// result — for one MUFU (ex2) and nine FMAs (the epilogues that apply it are instruction-bound).
__device__ __forceinline__ float gelu_erf_f(float v) {
  const float z = fminf(fabsf(v) * 0.70710678118654752f, 4.0f);
  float pz = fmaf(z, -2.177763781e-05f, 5.068330793e-04f);
  pz = fmaf(z, pz, -5.339398049e-03f);
  pz = fmaf(z, pz, 3.423144668e-02f);
  pz = fmaf(z, pz, -1.528908461e-01f);
  pz = fmaf(z, pz, -9.167589545e-01f);
  pz = fmaf(z, pz, -1.628154397e+00f);
  pz = fmaf(z, pz, 6.178960575e-06f);
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(pz));
  // 0.5 v (1 + sign(v) (1 - ex))  =  h + |h| - |h| ex   with h = v / 2
  const float h = 0.5f * v;
  return h + fmaf(-fabsf(h), ex, fabsf(h));
}

// Two GELUs at once on the packed fp32 pipe (fma.rn.f32x2, sm_100): same polynomial, half the FMA-pipe issue slots.
__device__ __forceinline__ float2 gelu_erf_f2(float2 v) {
  float2 z = __fmul2_rn(make_float2(fabsf(v.x), fabsf(v.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
  z.x = fminf(z.x, 4.0f); z.y = fminf(z.y, 4.0f);
  float2 pz = __ffma2_rn(z, make_float2(-2.177763781e-05f, -2.177763781e-05f), make_float2(5.068330793e-04f, 5.068330793e-04f));
  pz = __ffma2_rn(z, pz, make_float2(-5.339398049e-03f, -5.339398049e-03f));
  pz = __ffma2_rn(z, pz, make_float2(3.423144668e-02f, 3.423144668e-02f));
  pz = __ffma2_rn(z, pz, make_float2(-1.528908461e-01f, -1.528908461e-01f));
  pz = __ffma2_rn(z, pz, make_float2(-9.167589545e-01f, -9.167589545e-01f));
  pz = __ffma2_rn(z, pz, make_float2(-1.628154397e+00f, -1.628154397e+00f));
  pz = __ffma2_rn(z, pz, make_float2(6.178960575e-06f, 6.178960575e-06f));
  float2 ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(pz.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(pz.y));
  // 1 + sign(v) (1 - ex)  =  1 + s - s ex   with s = +-1
  const float2 s = make_float2(copysignf(1.0f, v.x), copysignf(1.0f, v.y));
  const float2 t = __ffma2_rn(make_float2(-s.x, -s.y), ex, make_float2(1.0f + s.x, 1.0f + s.y));
  return __fmul2_rn(__fmul2_rn(v, make_float2(0.5f, 0.5f)), t);
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Arrive on a barrier that lives in another CTA of the cluster (address from mapa_u32) with the default CTA-scope
// release.  The cluster-scope form (mbarrier.arrive.release.cluster) costs ~2000 cycles per arrival (it fences at
// cluster scope; profiles/r1_s21 vs r1_s22 timelines) and is not needed when the data the arrival publishes is consumed
// inside the ARRIVING thread's own CTA (its tensor core reading its own shared memory, made visible by
// fence.proxy.async) and the remote waiter only needs the count.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// Bounded spin: a pipeline bug must not hang the GPU box (that is a strike); after ~2 s of
// polling the kernel traps instead, which surfaces as a launch failure on the host.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFu) == 0 && global_timer_ns() - t0 > 2000000000ull) {
      printf("rs: mbarrier wait timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- CTA-pair (cta_group::2) variants: the pair's barrier lives in the leader CTA (cluster rank 0) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 16-byte load from the shared memory of another CTA of the cluster (address from mapa_u32)
__device__ __forceinline__ float4 ld_shared_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// smem -> global tile store (bulk async group), coordinates clip out-of-bounds elements
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the staged shared memory has been READ by every committed store (it may be reused / the CTA may exit); the global
// writes themselves complete asynchronously, at the latest at kernel completion
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// same, but the most recent committed store may still be reading (double-buffered staging)
__device__ __forceinline__ void tma_store_wait_read_keep1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 (5th-gen tensor cores, accumulators in TMEM)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_dst, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// cta_group::2 flavours (issued by the leader CTA of a pair; TMEM allocation by one warp of EACH CTA)
__device__ __forceinline__ void tmem_alloc_dyn_cg2(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn_cg2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs: 256 rows] (+)= A[128 rows from each CTA's smem] * B[N/2 rows from each CTA's smem]
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at this smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: 32 lanes x 32-bit, N consecutive columns per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// same, with the destination registers of the load as in/out operands: every use of v is ordered after the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle:
//   rows are 128 B (64 fp16) apart, 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100).
// Bit layout follows the PTX ISA "tcgen05 shared memory descriptor" (start addr [0,14), LBO [16,30),
// SBO [32,46), version [46,48), layout type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major); 16 B
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;            // version
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 accumulate, M x N.
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                                  // D format: F32
  d |= 0u << 7;                                  // A format: F16
  d |= 0u << 10;                                 // B format: F16
  d |= static_cast<uint32_t>(N >> 3) << 17;      // N / 8
  d |= static_cast<uint32_t>(M >> 4) << 24;      // M / 16
  return d;
}

#endif  // __CUDACC__

}  // namespace rs
