// GroupNorm statistics plumbing shared by every kernel that PRODUCES a tensor a GroupNorm will read.
//
// reference: GroupNorm32 = F.group_norm(x.float(), 32 groups, eps) (models/basic_ops.py:15-17; eps 1e-5 in the UNet,
// 1e-6 in the VQ-GAN: ldm/modules/diffusionmodules/model.py:47).  The reduction spans a whole image, so it cannot be a
// pure epilogue of one output tile.  Three stages, all deterministic (fixed summation order, no floating-point atomics):
//
//   1. every producer tile writes, per (image, tile slot, channel), the pair (mean, M2) of the fp16 values it STORED
//      (M2 = sum of squared deviations from that local mean).  Local sums are taken around a pivot (the first row of the
//      tile), so a channel with |mean| >> std loses nothing to cancellation (E[x^2] - mean^2 in fp32 does).
//   2. the LAST producer CTA to finish an image (an integer arrival counter per (GroupNorm, image) decides who that is;
//      the arithmetic does not depend on who) combines the pairs of each of the 32 groups — Chan et al.'s parallel
//      variance formula, again around a pivot — into gstat[image][group] = (mean, rstd).
//   3. consumers (gn_apply_kernel, the fused MLP's and the qkv GEMM's in-shared-memory operand transform) read the 32
//      pairs of their image and fold gamma / beta (/ FiLM) into a per-channel affine.
//
// Compared with round 1 (every consumer CTA re-reducing slots x C raw sums: 41 KB per CTA at 64x64, 0.5 MB at 256x256)
// the consumers' preamble is one 256-byte read, and the statistics are robust.
#pragma once

#include "common.cuh"

namespace rs {

// what a producer needs to know about ONE consuming GroupNorm (up to two per producer: a skip tensor feeds the next
// encoder block and, later, the decoder's concat GroupNorm)
struct GnSink {
  float* part;            // [N][slots][cstride][2] = (mean, M2) per image / tile slot / channel; nullptr: no statistics
  float* gstat;           // [N][32][2] = (group mean, group rstd), written by the last-arriving producer CTA
  unsigned int* counter;  // [N] channel-slots delivered so far (zeroed before every forward)
  int cstride;            // channel count of the consumer's tensor (this producer may cover only a slice of it)
  int coff;               // first channel of this producer's slice
  unsigned int expected;  // slots * cstride: the image is complete when the counter reaches it
  float eps;
};

#ifdef __CUDACC__

__device__ __forceinline__ float2 ldcg_f2(const float* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}

// counter += v with acquire-release semantics at GPU scope, by ONE thread after a CTA barrier: the barrier orders the
// other threads' pair stores before it (causality is cumulative over bar.sync), the release half publishes them, and —
// if this turns out to be the last arrival — the acquire half orders the finaliser's reads after every earlier arrival.
// (Each thread issuing __threadfence() instead costs MEMBAR.SC.GPU + CCTL.IVALL x 256 threads per tile: measured
// +50 % on the stats-bearing 3x3 layers, profiles/r2_s1_*.)
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// Combine two (count, mean, M2) triples with equal counts n each (Chan et al.): used for the two 64-row halves of a tile.
__device__ __forceinline__ void chan_merge_equal(float n, float m0, float q0, float m1, float q1, float& m, float& q) {
  m = 0.5f * (m0 + m1);
  const float d0 = m0 - m, d1 = m1 - m;
  q = q0 + q1 + n * (d0 * d0 + d1 * d1);
}

// Group statistics of image `n` from the per-(slot, channel) pairs: ONE warp (all 32 lanes) reduces kG groups at once
// (g0, g0 + gstep, ...) so that kG x 4 independent L2 loads are in flight per lane — the finaliser sits on the tail of a
// producer kernel and is pure load latency.  Per group: K = slots * cpg items of `ns` values each, single pass around
// the pivot item 0; lanes take items lane, lane + 32, ...; lane partials are combined with a fixed shuffle tree.
template <int kG>
__device__ __forceinline__ void gn_finalize_groups(const float* part, int n, int slots, int C, int g0, int gstep, int cpg, float ns,
                                                   float eps, float* gstat, int lane) {
  const float* base = part + (size_t)n * slots * C * 2;
  const int K = slots * cpg;
  float pivot[kG], s1[kG], s2[kG];
#pragma unroll
  for (int q = 0; q < kG; ++q) {
    const int g = g0 + q * gstep;
    pivot[q] = g < 32 ? ldcg_f2(base + (size_t)(g * cpg) * 2).x : 0.f;
    s1[q] = 0.f; s2[q] = 0.f;
  }
  for (int i0 = 0; i0 < K; i0 += 128) {
    float2 e[kG][4];
#pragma unroll
    for (int q = 0; q < kG; ++q) {
      const int g = g0 + q * gstep;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        if (g < 32 && i < K) {
          const int sl = i / cpg, c = i - sl * cpg;
          e[q][u] = ldcg_f2(base + ((size_t)sl * C + g * cpg + c) * 2);
        } else {
          e[q][u] = make_float2(pivot[q], 0.f);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kG; ++q)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d = e[q][u].x - pivot[q];
        s1[q] += d;
        s2[q] += fmaf(ns * d, d, e[q][u].y);
      }
  }
#pragma unroll
  for (int q = 0; q < kG; ++q) {
    float a = s1[q], b = s2[q];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    const int g = g0 + q * gstep;
    const float invK = 1.0f / (float)K;
    const float dm = a * invK;                    // group mean - pivot
    const float m2 = fmaxf(b - ns * (float)K * dm * dm, 0.f);
    const float var = m2 * invK / ns;
    if (lane == 0 && g < 32) {
      gstat[((size_t)n * 32 + g) * 2] = pivot[q] + dm;
      gstat[((size_t)n * 32 + g) * 2 + 1] = rsqrtf(var + eps);
    }
  }
}

// Arrival protocol, called by ALL `nthreads` threads of a producer's epilogue group (named barrier `bar_id`) AFTER they
// have written their (mean, M2) pairs for up to kMax (sink, image) pairs.  `img[i] < 0` = nothing delivered for entry i.
//   add[i]  = channels this CTA delivered for (sink[i], img[i]) in this tile slot
// The last arriver of an image reduces the image's 32 groups (warp w takes groups w, w + #warps, ... four at a time).
template <int kMax>
__device__ __forceinline__ void gn_arrive(const GnSink* const (&sink)[kMax], const int (&img)[kMax], const unsigned int (&add)[kMax],
                                          int slots, float ns, int tid, int nthreads, int bar_id, int* s_flag /* [kMax] shared */) {
  named_bar_sync(bar_id, nthreads);               // every thread's pair stores happen-before the arrival below
  if (tid < kMax) {
    int last = 0;
    if (sink[tid] != nullptr && img[tid] >= 0 && sink[tid]->gstat != nullptr) {
      const unsigned int old = atom_add_acq_rel_gpu(sink[tid]->counter + img[tid], add[tid]);
      last = (old + add[tid] == sink[tid]->expected) ? 1 : 0;
    }
    s_flag[tid] = last;
  }
  named_bar_sync(bar_id, nthreads);               // ... and the acquiring arrival happens-before the finaliser's reads
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    if (!s_flag[i]) continue;                     // uniform across the group
    const GnSink& s = *sink[i];                   // (pairs are read with ld.global.cg: L2, never a stale L1 line)
    const int cpg = s.cstride / 32;
    for (int g = warp; g < 32; g += 4 * nwarps) gn_finalize_groups<4>(s.part, img[i], slots, s.cstride, g, nwarps, cpg, ns, s.eps, s.gstat, lane);
  }
  named_bar_sync(bar_id, nthreads);               // s_flag may be rewritten by the next tile
}

// Deferred arrivals of a persistent producer: instead of one arrival (barrier + GPU-scope atomic round trip + barrier,
// ~2 us during which the eight epilogue warps idle) per output tile, thread 0 records (sink, image, channels) in a small
// shared list — merging repeats — and the CTA arrives ONCE for every entry after its last tile.
constexpr int kGnListCap = 64;
struct GnArriveList {
  int cnt;
  int over;                       // set by thread 0 when the list cannot take another tile's entries: arrive now instead
  short sink[kGnListCap];
  short img[kGnListCap];
  unsigned int add[kGnListCap];
  int flag[kGnListCap];
};
__device__ __forceinline__ void gn_list_add(GnArriveList* L, int d, int img, unsigned int add) {   // thread 0 only
  for (int j = 0; j < L->cnt; ++j)
    if (L->sink[j] == d && L->img[j] == img) { L->add[j] += add; return; }
  const int j = L->cnt++;
  L->sink[j] = (short)d; L->img[j] = (short)img; L->add[j] = add;
}
// all `nthreads` threads of the epilogue group, after the CTA's last tile (every pair store precedes the first barrier)
__device__ __forceinline__ void gn_list_arrive(GnArriveList* L, const GnSink& s0, const GnSink& s1, int slots, float ns, int tid,
                                               int nthreads, int bar_id) {
  named_bar_sync(bar_id, nthreads);
  const int cnt = L->cnt;
  for (int j = tid; j < cnt; j += nthreads) {
    const GnSink& s = L->sink[j] == 0 ? s0 : s1;
    const unsigned int old = atom_add_acq_rel_gpu(s.counter + L->img[j], L->add[j]);
    L->flag[j] = (old + L->add[j] == s.expected) ? 1 : 0;
  }
  named_bar_sync(bar_id, nthreads);
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  for (int j = 0; j < cnt; ++j) {
    if (!L->flag[j]) continue;
    const GnSink& s = L->sink[j] == 0 ? s0 : s1;
    const int cpg = s.cstride / 32;
    for (int g = warp; g < 32; g += 4 * nwarps) gn_finalize_groups<4>(s.part, L->img[j], slots, s.cstride, g, nwarps, cpg, ns, s.eps, s.gstat, lane);
  }
  named_bar_sync(bar_id, nthreads);
  if (tid == 0) L->cnt = 0;
}

// (mean, M2) of the `rows` fp16 values x[r * pitch_h] (r = 0 .. rows-1) read through `ld(r)` -> float2 (two adjacent
// columns), around the pivot ld(0).  Returns mean / M2 per column.
template <typename Ld>
__device__ __forceinline__ void pivot_stats2(Ld ld, int rows, float2& mean, float2& m2) {
  const float2 p = ld(0);
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
  for (int r = 1; r < rows; ++r) {
    const float2 v = ld(r);
    const float dx = v.x - p.x, dy = v.y - p.y;
    s1.x += dx; s1.y += dy;
    s2.x = fmaf(dx, dx, s2.x); s2.y = fmaf(dy, dy, s2.y);
  }
  const float inv = 1.0f / (float)rows;
  mean = make_float2(p.x + s1.x * inv, p.y + s1.y * inv);
  m2 = make_float2(fmaxf(s2.x - s1.x * s1.x * inv, 0.f), fmaxf(s2.y - s1.y * s1.y * inv, 0.f));
}

// Column statistics of a staged output tile: [BN / bc blocks][128 rows][bc columns] fp16, 16-byte units XOR-swizzled
// like the TMA box (Swizzle<B,4,3>: bc = 64 -> unit ^= row & 7, 32 -> (row >> 1) & 3, 16 -> (row >> 2) & 1).
// Thread `tid` (< BN) takes column pair (tid % (BN/2)) of the 64-row half (tid / (BN/2)) and writes
// wstat[half][col][2] = (mean, M2) for both columns.  Caller synchronises before reading wstat.
__device__ __forceinline__ void staged_tile_column_stats(const uint8_t* stage, int BN, int bc, int tid, float* wstat) {
  const int pairs = BN >> 1;
  if (tid >= 2 * pairs) return;
  const int half = tid / pairs, cp = tid - half * pairs;
  const int col = 2 * cp;
  const int blk = col / bc, cb = col - blk * bc;
  const int u0 = cb >> 3, inner = (cb & 7) * 2;
  const int pitch = 2 * bc;
  const uint8_t* base = stage + (size_t)blk * (128 * pitch) + (size_t)(half * 64) * pitch + inner;
  const int sh = (bc == 64) ? 0 : (bc == 32 ? 1 : 2);
  const int msk = (bc == 64) ? 7 : (bc == 32 ? 3 : 1);
  auto ld = [&](int r) {
    const int rr = half * 64 + r;                                      // row inside the tile (swizzle uses the tile row)
    const __half2 h = *reinterpret_cast<const __half2*>(base + (size_t)r * pitch + (((u0 ^ ((rr >> sh) & msk))) << 4));
    return __half22float2(h);
  };
  // unrolled by 8 rows (the swizzle pattern has period 8): loads of a group are independent
  const float2 p = ld(0);
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
  for (int r0 = 0; r0 < 64; r0 += 8) {
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ld(r0 + j);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dx = v[j].x - p.x, dy = v[j].y - p.y;
      s1.x += dx; s1.y += dy;
      s2.x = fmaf(dx, dx, s2.x); s2.y = fmaf(dy, dy, s2.y);
    }
  }
  const float inv = 1.0f / 64.0f;
  float* dst = wstat + ((size_t)half * BN + col) * 2;
  dst[0] = p.x + s1.x * inv; dst[1] = fmaxf(s2.x - s1.x * s1.x * inv, 0.f);
  dst[2] = p.y + s1.y * inv; dst[3] = fmaxf(s2.y - s1.y * s1.y * inv, 0.f);
}

// Final write of one tile's pairs from wstat[2 halves][BN][2] (as produced above, or by a plain [128][cw] column pass)
// into up to two sinks.  bn = images per tile (1: both halves belong to image n0 and are merged, ns = 128; 2: half h
// belongs to image n0 + h, ns = 64).  Thread `tid` of `nthreads` walks the tile's `ncols` valid columns.
__device__ __forceinline__ void write_tile_pairs(const float* wstat, int BN, int ncols, int col0, int bn, int n0, int Nimg, int slot,
                                                 int slots, const GnSink& s0, const GnSink& s1, int tid, int nthreads) {
  for (int cc = tid; cc < ncols; cc += nthreads) {
    const float m0 = wstat[((size_t)0 * BN + cc) * 2], q0 = wstat[((size_t)0 * BN + cc) * 2 + 1];
    const float m1 = wstat[((size_t)1 * BN + cc) * 2], q1 = wstat[((size_t)1 * BN + cc) * 2 + 1];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const GnSink& s = d == 0 ? s0 : s1;
      if (!s.part) continue;
      const size_t ch = (size_t)s.coff + col0 + cc;
      float* dst = s.part + (((size_t)n0 * slots + slot) * s.cstride + ch) * 2;
      if (bn == 1) {
        float m, q;
        chan_merge_equal(64.f, m0, q0, m1, q1, m, q);
        dst[0] = m; dst[1] = q;
      } else {
        dst[0] = m0; dst[1] = q0;
        if (n0 + 1 < Nimg) {
          float* dst1 = s.part + (((size_t)(n0 + 1) * slots + slot) * s.cstride + ch) * 2;
          dst1[0] = m1; dst1[1] = q1;
        }
      }
    }
  }
}

#endif  // __CUDACC__
}  // namespace rs
