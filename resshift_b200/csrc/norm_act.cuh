// GroupNorm(32 groups, eps 1e-5, fp32 statistics) + FiLM + SiLU on NHWC fp16 views.
//
// reference: GroupNorm32 (models/basic_ops.py:15-17,89-96), its uses in ResBlock
// (models/unet.py:144-148,168-175,198-202: GN -> SiLU, and GN*(1+scale)+shift -> SiLU) and in
// SwinTransformerBlock (models/swin_transformer.py:248,279: plain GN), final head (unet.py:859-863).
//
// Statistics are kept as DETERMINISTIC partial sums  part[N][slots][C][2]  (sum, sum of squares per image,
// per row-slot, per channel; no atomics anywhere), produced either
//   * by the epilogue of the conv/GEMM kernel that wrote the tensor (conv_gemm.cuh, one slot per 128-pixel tile), or
//   * by gn_stats_kernel below (one slot per CTA) for tensors that have no fusable producer.
// gn_apply_kernel folds (slots -> group mean / rstd, gamma, beta, FiLM) into a per-(image, channel) affine
// a*x+b in shared memory, then streams x -> y = act(a*x+b) with 128-bit accesses.
#pragma once

#include "common.cuh"

namespace rs {

struct GnStatsParams {
  const __half* x;          // view [N][HW][C], row stride ld
  long long sN;             // image stride (elements)
  int ld, C, HW, N;
  float* part;              // [N][slots][C][2]
  int slots;
  int rows_per_slot;
};

struct GnApplyParams {
  const __half* x; long long x_sN; int x_ld;
  __half* y; long long y_sN; int y_ld;
  int C, HW, N;
  const float* part;        // [N][slots][C][2]
  int slots;
  const float* gamma;       // [C]
  const float* beta;        // [C]
  const float* film;        // optional [N or 1][2*C] : scale = film[0:C], shift = film[C:2C]
  long long film_sN;        // 0 when the same timestep embedding is shared by the whole batch
  int silu;
  int rows_per_cta;
  float eps;
  int Cs;                   // channels per CTA (blockIdx.z selects the slice; a multiple of 8 and of C/32): small tensors
                            // are split over channels as well as rows so that every SM gets a CTA
};

#ifdef __CUDACC__

// One CTA per (slot, image).  Each thread owns one 8-channel vector column and walks rows (4 loads in flight);
// row-lanes are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(256) gn_stats_kernel(const GnStatsParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_red[];   // [lanes][C][2]
  const int vecs = p.C >> 3;
  const int lanes = blockDim.x / vecs;
  const int n = blockIdx.y, slot = blockIdx.x;
  const int vec = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  if (rl < lanes) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const int r0 = slot * p.rows_per_slot;
    const int r1 = min(r0 + p.rows_per_slot, p.HW);
    const __half* base = p.x + n * p.sN + vec * 8;
    int r = r0 + rl;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint4*>(base + (long long)(r + u * lanes) * p.ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          s[2 * j] += f.x; q[2 * j] += f.x * f.x;
          s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
        }
      }
    }
    for (; r < r1; r += lanes) {
      const uint4 raw = *reinterpret_cast<const uint4*>(base + (long long)r * p.ld);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
    float* dst = s_red + ((size_t)rl * p.C + vec * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dst[2 * j] = s[j]; dst[2 * j + 1] = q[j]; }
  }
  __syncthreads();
  float* out = p.part + ((size_t)n * p.slots + slot) * p.C * 2;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += s_red[(size_t)l * p.C * 2 + i];
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const GnApplyParams p) {
  pdl_trigger();
  extern __shared__ float s_ab[];    // a[Cs], b[Cs], gamma[Cs], beta[Cs], mean[32], rstd[32]  (this CTA's channel slice)
  const int Cs = p.Cs;
  const int c0 = blockIdx.z * Cs;    // first channel of the slice (group aligned)
  float* s_a = s_ab;
  float* s_b = s_ab + Cs;
  float* s_g = s_ab + 2 * Cs;
  float* s_be = s_ab + 3 * Cs;
  float* s_mean = s_ab + 4 * Cs;
  float* s_rstd = s_mean + 32;
  const int n = blockIdx.y;
  const int cpg = p.C / 32;
  // layer parameters do not depend on the producing kernel: fetch them while it drains
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) { s_g[c] = __ldg(p.gamma + c0 + c); s_be[c] = __ldg(p.beta + c0 + c); }
  pdl_wait();
  // per-channel totals over the slots (independent loads, fixed order), staged in s_a / s_b together with the FiLM
  // pair (kept in registers: a thread owns the same channels in both passes); then per-group mean / rstd in a fixed
  // order over the group's channels
  float f_sc[8], f_sh[8];            // C <= 2048 -> at most 8 channels per thread
  {
    const float* part = p.part + (size_t)n * p.slots * p.C * 2 + (size_t)c0 * 2;
    const float* f = p.film ? p.film + n * p.film_sN + c0 : nullptr;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int c = threadIdx.x + it * 256;
      if (c >= Cs) break;
      f_sc[it] = f ? 1.0f + f[c] : 1.0f;
      f_sh[it] = f ? f[p.C + c] : 0.0f;
      float s = 0.f, q = 0.f;
      int sl = 0;
      // 16 independent loads in flight per thread (a 64x64 layer has 32 slots: two round trips instead of eight —
      // this dependent chain, not bandwidth, is what the small GroupNorm launches wait for); summed in slot order
      for (; sl + 16 <= p.slots; sl += 16) {
        float2 e[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) e[u] = *reinterpret_cast<const float2*>(part + ((size_t)(sl + u) * p.C + c) * 2);
#pragma unroll
        for (int u = 0; u < 16; ++u) { s += e[u].x; q += e[u].y; }
      }
      for (; sl + 4 <= p.slots; sl += 4) {
        float2 e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = *reinterpret_cast<const float2*>(part + ((size_t)(sl + u) * p.C + c) * 2);
#pragma unroll
        for (int u = 0; u < 4; ++u) { s += e[u].x; q += e[u].y; }
      }
      for (; sl < p.slots; ++sl) {
        const float2 e = *reinterpret_cast<const float2*>(part + ((size_t)sl * p.C + c) * 2);
        s += e.x; q += e.y;
      }
      s_a[c] = s; s_b[c] = q;
    }
    __syncthreads();
    if (threadIdx.x < Cs / cpg) {
      const int g = threadIdx.x;         // group index inside the slice
      float s = 0.f, q = 0.f;
      for (int j = 0; j < cpg; ++j) { s += s_a[g * cpg + j]; q += s_b[g * cpg + j]; }
      const float inv_cnt = 1.0f / (float)((long long)cpg * p.HW);
      const float mean = s * inv_cnt;
      const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
      s_mean[g] = mean; s_rstd[g] = rsqrtf(var + p.eps);
    }
  }
  __syncthreads();
  {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int c = threadIdx.x + it * 256;
      if (c >= Cs) break;
      const int g = c / cpg;
      float a = s_rstd[g] * s_g[c];
      float b = s_be[c] - s_mean[g] * a;
      a *= f_sc[it];
      b = b * f_sc[it] + f_sh[it];
      s_a[c] = a; s_b[c] = b;
    }
  }
  __syncthreads();
  const int vecs = Cs >> 3;
  const int lanes = blockDim.x / vecs;
  const int vec = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  if (rl >= lanes) return;
  const int c = vec * 8;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = s_a[c + j]; b[j] = s_b[c + j]; }
  const int r0 = blockIdx.x * p.rows_per_cta;
  const int r1 = min(r0 + p.rows_per_cta, p.HW);
  const __half* xb = p.x + n * p.x_sN + c0 + c;
  __half* yb = p.y + n * p.y_sN + c0 + c;
  auto one = [&](const uint4& raw) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      f.x = fmaf(f.x, a[2 * j], b[2 * j]);
      f.y = fmaf(f.y, a[2 * j + 1], b[2 * j + 1]);
      if (p.silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
      oh[j] = __floats2half2_rn(f.x, f.y);
    }
    return o;
  };
  int r = r0 + rl;
  for (; r + 7 * lanes < r1; r += 8 * lanes) {
    uint4 raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) raw[u] = *reinterpret_cast<const uint4*>(xb + (long long)(r + u * lanes) * p.x_ld);
#pragma unroll
    for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(yb + (long long)(r + u * lanes) * p.y_ld) = one(raw[u]);
  }
  for (; r + 3 * lanes < r1; r += 4 * lanes) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint4*>(xb + (long long)(r + u * lanes) * p.x_ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(yb + (long long)(r + u * lanes) * p.y_ld) = one(raw[u]);
  }
  for (; r < r1; r += lanes) {
    const uint4 raw = *reinterpret_cast<const uint4*>(xb + (long long)r * p.x_ld);
    *reinterpret_cast<uint4*>(yb + (long long)r * p.y_ld) = one(raw);
  }
}

#endif
}  // namespace rs
