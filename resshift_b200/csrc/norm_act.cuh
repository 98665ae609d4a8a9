// GroupNorm(32 groups, eps 1e-5, fp32 statistics) + FiLM + SiLU on NHWC fp16 views.
//
// reference: GroupNorm32 (models/basic_ops.py:15-17,89-96), its uses in ResBlock
// (models/unet.py:144-148,168-175,198-202: GN -> SiLU, and GN*(1+scale)+shift -> SiLU) and in
// SwinTransformerBlock (models/swin_transformer.py:248,279: plain GN), final head (unet.py:859-863).
//
// Two HBM/L2-bound passes:
//   gn_stats_kernel : per (image, channel) sum / sum-of-squares -> fp32 atomics into [N][C][2]
//   gn_apply_kernel : folds (mean, rstd, gamma, beta, FiLM) into a per-(image, channel) affine
//                     a*x+b in shared memory, then streams x -> y = act(a*x+b) with 128-bit accesses.
#pragma once

#include "common.cuh"

namespace rs {

struct GnStatsParams {
  const __half* x;          // view [N][HW][C], row stride ld
  long long sN;             // image stride (elements)
  int ld, C, HW, N;
  float* sums;              // [N][C][2], zeroed before the launch
  int rows_per_cta;
};

struct GnApplyParams {
  const __half* x; long long x_sN; int x_ld;
  __half* y; long long y_sN; int y_ld;
  int C, HW, N;
  const float* sums;        // [N][C][2]
  const float* gamma;       // [C]
  const float* beta;        // [C]
  const float* film;        // optional [N or 1][2*C] : scale = film[0:C], shift = film[C:2C]
  long long film_sN;        // 0 when the same timestep embedding is shared by the whole batch
  int silu;
  int rows_per_cta;
  float eps;
};

#ifdef __CUDACC__

// Each thread owns one 8-channel vector column and walks rows; per-channel partials are combined
// through shared-memory atomics, then one global atomic per (channel, moment) per CTA.
__global__ void __launch_bounds__(256) gn_stats_kernel(const GnStatsParams p) {
  extern __shared__ float s_acc[];   // [C][2]
  const int vecs = p.C >> 3;
  const int lanes = blockDim.x / vecs;          // row lanes
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int vec = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  if (rl < lanes) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const int r0 = blockIdx.x * p.rows_per_cta;
    const int r1 = min(r0 + p.rows_per_cta, p.HW);
    const __half* base = p.x + n * p.sN + vec * 8;
    for (int r = r0 + rl; r < r1; r += lanes) {
      const uint4 raw = *reinterpret_cast<const uint4*>(base + (long long)r * p.ld);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_acc[2 * (vec * 8 + j)], s[j]);
      atomicAdd(&s_acc[2 * (vec * 8 + j) + 1], q[j]);
    }
  }
  __syncthreads();
  float* dst = p.sums + (long long)n * p.C * 2;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) atomicAdd(dst + i, s_acc[i]);
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyParams p) {
  extern __shared__ float s_ab[];    // a[C], b[C]
  float* s_a = s_ab;
  float* s_b = s_ab + p.C;
  const int n = blockIdx.y;
  const int cpg = p.C / 32;
  const float inv_cnt = 1.0f / (float)(cpg * p.HW);
  const float* sums = p.sums + (long long)n * p.C * 2;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    float s = 0.f, q = 0.f;
    for (int j = 0; j < cpg; ++j) { s += sums[2 * (g0 + j)]; q += sums[2 * (g0 + j) + 1]; }
    const float mean = s * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    float a = rstd * p.gamma[c];
    float b = p.beta[c] - mean * a;
    if (p.film) {
      const float* f = p.film + n * p.film_sN;
      const float sc = 1.0f + f[c];
      a *= sc;
      b = b * sc + f[p.C + c];
    }
    s_a[c] = a; s_b[c] = b;
  }
  __syncthreads();
  const int vecs = p.C >> 3;
  const int r0 = blockIdx.x * p.rows_per_cta;
  const int r1 = min(r0 + p.rows_per_cta, p.HW);
  const long long total = (long long)(r1 - r0) * vecs;
  const __half* xb = p.x + n * p.x_sN;
  __half* yb = p.y + n * p.y_sN;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = r0 + (int)(i / vecs);
    const int c = (int)(i % vecs) * 8;
    const uint4 raw = *reinterpret_cast<const uint4*>(xb + (long long)r * p.x_ld + c);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      f.x = fmaf(f.x, s_a[c + 2 * j], s_b[c + 2 * j]);
      f.y = fmaf(f.y, s_a[c + 2 * j + 1], s_b[c + 2 * j + 1]);
      if (p.silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
      oh[j] = __floats2half2_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(yb + (long long)r * p.y_ld + c) = o;
  }
}

#endif
}  // namespace rs
