// Fused (shifted-)window multi-head self-attention core for 8x8 windows, head_dim 32.
//
// reference: WindowAttention.forward (models/swin_transformer.py:114-145) together with the
// data movement around it in SwinTransformerBlock.forward (:251-275): torch.roll(-s), window_partition,
// q*scale, q@k^T, + relative-position bias, + shift mask, softmax, @v, window_reverse, torch.roll(+s).
// All of the movement is address arithmetic here: token (r, c) of window (wy, wx) of image n lives at
// pixel ((wy*8 + r + s) % H, (wx*8 + c + s) % W) of the un-shifted NHWC tensor, for reads and writes.
//
//   qkv : [N*H*W, 3*E] fp16, channel = which*E + head*32 + d      (output of the qkv GEMM, bias included)
//   out : [N*H*W, E]   fp16, channel = head*32 + d                (input of the proj GEMM)
//   bias: [heads][64][64] fp32, relative_position_bias_table gathered by relative_position_index
//   mask: generated on the fly; reproduces the reference's calculate_mask (:214-236) including its
//         axis quirks (see resshift_b200/arch.py::shifted_window_mask): label(token) = region(wy*8 + c).
//
// One CTA = one (window, head); 4 warps x 16 query rows; QK^T and PV on mma.sync m16n8k16 with the
// score tile kept in registers (C-fragment of QK^T is reused as the A-fragment of PV).
#pragma once

#include "common.cuh"

namespace rs {

struct WinAttnParams {
  const __half* qkv; int qkv_ld;
  __half* out; int out_ld;
  const float* bias;        // [heads][64][64]
  int N, H, W, heads, E;
  int shift;                // 0 or 4
  float scale;              // head_dim^-0.5
  int use_simt;
};

#ifdef __CUDACC__

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// region label of a token for the shifted-window mask (reference quirk: depends on wy and the token COLUMN)
__device__ __forceinline__ int swin_label(int wy, int c, int H, int shift) {
  const int y = wy * 8 + c;
  return (y < H - 8) ? 0 : ((y < H - shift) ? 1 : 2);
}

constexpr int kAttnPad = 40;   // halves per smem row (32 + 8 pad: conflict-free 32-bit fragment reads)

__global__ void __launch_bounds__(128) window_attn_kernel(const WinAttnParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) __half sQ[64 * kAttnPad];
  __shared__ __align__(16) __half sK[64 * kAttnPad];
  __shared__ __align__(16) __half sVt[32 * 72];        // V transposed: [d][token], 64 + 8 pad
  __shared__ int sPix[64];

  const int head = blockIdx.y;
  const int nWx = p.W >> 3, nWy = p.H >> 3;
  int win = blockIdx.x;
  const int wx = win % nWx; win /= nWx;
  const int wy = win % nWy; win /= nWy;
  const int n = win;

  if (threadIdx.x < 64) {
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    const int y = (wy * 8 + r + p.shift) % p.H;
    const int x = (wx * 8 + c + p.shift) % p.W;
    sPix[threadIdx.x] = (n * p.H + y) * p.W + x;
  }
  __syncthreads();
  // stage q, k, v of this (window, head): 64 tokens x 3 x 32 halves = 64 x 3 x 4 uint4
  for (int i = threadIdx.x; i < 64 * 12; i += blockDim.x) {
    const int tok = i / 12, rem = i % 12, which = rem >> 2, part = rem & 3;
    const __half* src = p.qkv + (long long)sPix[tok] * p.qkv_ld + which * p.E + head * 32 + part * 8;
    const uint4 raw = *reinterpret_cast<const uint4*>(src);
    if (which == 0) {
      *reinterpret_cast<uint4*>(&sQ[tok * kAttnPad + part * 8]) = raw;
    } else if (which == 1) {
      *reinterpret_cast<uint4*>(&sK[tok * kAttnPad + part * 8]) = raw;
    } else {
      const __half* hv = reinterpret_cast<const __half*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) sVt[(part * 8 + j) * 72 + tok] = hv[j];
    }
  }
  __syncthreads();

  const float* bias = p.bias + (long long)head * 64 * 64;

  if (p.use_simt) {
    // ---- plain fp32 path (debug cross-check): one thread per query row ----
    if (threadIdx.x < 64) {
      const int i = threadIdx.x;
      float q[32], sc[64];
      for (int d = 0; d < 32; ++d) q[d] = __half2float(sQ[i * kAttnPad + d]) * p.scale;
      const int li = p.shift ? swin_label(wy, i & 7, p.H, p.shift) : 0;
      float mx = -1e30f;
      for (int j = 0; j < 64; ++j) {
        float s = 0.f;
        for (int d = 0; d < 32; ++d) s = fmaf(q[d], __half2float(sK[j * kAttnPad + d]), s);
        s += bias[i * 64 + j];
        if (p.shift && swin_label(wy, j & 7, p.H, p.shift) != li) s += -100.0f;
        sc[j] = s; mx = fmaxf(mx, s);
      }
      float sum = 0.f;
      for (int j = 0; j < 64; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
      const float inv = 1.0f / sum;
      __half* dst = p.out + (long long)sPix[i] * p.out_ld + head * 32;
      for (int d = 0; d < 32; ++d) {
        float o = 0.f;
        for (int j = 0; j < 64; ++j) o = fmaf(sc[j], __half2float(sVt[d * 72 + j]), o);
        dst[d] = __float2half_rn(o * inv);
      }
    }
    return;
  }

  // ---- tensor-core path ----
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16 + g;          // this lane's rows: row0 and row0 + 8

  // Q fragments for the two k-steps (d 0..15, 16..31)
  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int d = ks * 16 + 2 * t;
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(&sQ[row0 * kAttnPad + d]);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(&sQ[(row0 + 8) * kAttnPad + d]);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(&sQ[row0 * kAttnPad + d + 8]);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(&sQ[(row0 + 8) * kAttnPad + d + 8]);
  }
  float s[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int key = nt * 8 + g, d = ks * 16 + 2 * t;
      uint32_t kb[2];
      kb[0] = *reinterpret_cast<const uint32_t*>(&sK[key * kAttnPad + d]);
      kb[1] = *reinterpret_cast<const uint32_t*>(&sK[key * kAttnPad + d + 8]);
      mma_16816(s[nt], qa[ks], kb);
    }
  }
  // scale, bias, mask; row-wise softmax (each row is spread over the 4 lanes of a quad)
  const int la = p.shift ? swin_label(wy, row0 & 7, p.H, p.shift) : 0;     // (row0+8)&7 == row0&7
  float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = nt * 8 + 2 * t + e;
      float m = 0.f;
      if (p.shift && swin_label(wy, col & 7, p.H, p.shift) != la) m = -100.0f;
      s[nt][e] = s[nt][e] * p.scale + bias[row0 * 64 + col] + m;
      s[nt][2 + e] = s[nt][2 + e] * p.scale + bias[(row0 + 8) * 64 + col] + m;
      mx0 = fmaxf(mx0, s[nt][e]);
      mx1 = fmaxf(mx1, s[nt][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[nt][e] = __expf(s[nt][e] - mx0); sum0 += s[nt][e];
      s[nt][2 + e] = __expf(s[nt][2 + e] - mx1); sum1 += s[nt][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  // O = P V : k = keys (4 steps of 16), n = d (4 tiles of 8)
  float o[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int d = dt * 8 + g, key = kk * 16 + 2 * t;
      uint32_t vb[2];
      vb[0] = *reinterpret_cast<const uint32_t*>(&sVt[d * 72 + key]);
      vb[1] = *reinterpret_cast<const uint32_t*>(&sVt[d * 72 + key + 8]);
      mma_16816(o[dt], pa, vb);
    }
  }
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  __half* d0 = p.out + (long long)sPix[row0] * p.out_ld + head * 32;
  __half* d1 = p.out + (long long)sPix[row0 + 8] * p.out_ld + head * 32;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    const int d = dt * 8 + 2 * t;
    *reinterpret_cast<__half2*>(d0 + d) = __floats2half2_rn(o[dt][0] * inv0, o[dt][1] * inv0);
    *reinterpret_cast<__half2*>(d1 + d) = __floats2half2_rn(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

#endif
}  // namespace rs
