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
// window_attn_kernel: one CTA per window, 4 warps x 16 query rows, looping over the heads with a
// double-buffered cp.async pipeline (head h+1 streams in while head h is computed).  QK^T and PV run on
// mma.sync m16n8k16 with the score tile kept in registers (the C fragment of QK^T is the A fragment of PV);
// V is read through ldmatrix.trans; the 64 x E output tile is staged in shared memory and written as
// full 2*E-byte rows.  window_attn_simt_kernel is a plain fp32 version kept as a cross-check (RS_ATTN_IMPL=simt).
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
  int hpc;                  // heads per CTA (grid.y = heads / hpc): fewer heads per CTA when there are few windows
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
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// two 8x8 b16 matrices, transposed on load: the B fragment (k x n, "col") of m16n8k16 from a row-major [k][n] tile
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* row_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(smem_u32(row_addr)));
}

// region label of a token for the shifted-window mask (reference quirk: depends on wy and the token COLUMN)
__device__ __forceinline__ int swin_label(int wy, int c, int H, int shift) {
  const int y = wy * 8 + c;
  return (y < H - 8) ? 0 : ((y < H - shift) ? 1 : 2);
}

constexpr int kAttnPad = 40;   // halves per smem row (32 + 8 pad): conflict-free fragment reads and ldmatrix rows

__global__ void __launch_bounds__(128) window_attn_kernel(const WinAttnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t attn_smem[];
  // [2 buffers][q | k | v][64][kAttnPad] halves, then the output tile [64][E + 8] halves, then the pixel table
  __half* sbuf = reinterpret_cast<__half*>(attn_smem);
  const int buf_halves = 3 * 64 * kAttnPad;
  const int opitch = p.hpc * 32 + 8;
  const int head0 = blockIdx.y * p.hpc;
  __half* sOut = sbuf + 2 * buf_halves;
  int* sPix = reinterpret_cast<int*>(sOut + 64 * opitch);

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

  auto stage_head = [&](int head, int buf) {
    __half* dst = sbuf + buf * buf_halves;
    for (int i = threadIdx.x; i < 64 * 12; i += 128) {
      const int tok = i / 12, rem = i - tok * 12, which = rem >> 2, part = rem & 3;
      const __half* src = p.qkv + (long long)sPix[tok] * p.qkv_ld + which * p.E + head * 32 + part * 8;
      cp_async_16(dst + which * 64 * kAttnPad + tok * kAttnPad + part * 8, src);
    }
    cp_async_commit();
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16 + g;          // this lane's rows: row0 and row0 + 8
  const int la = p.shift ? swin_label(wy, row0 & 7, p.H, p.shift) : 0;     // (row0 + 8) & 7 == row0 & 7

  stage_head(head0, 0);
  for (int hi = 0; hi < p.hpc; ++hi) {
    const int head = head0 + hi;
    const int buf = hi & 1;
    // this lane's 32 relative-position-bias values (a load-time table, independent of the staged q/k/v): issued before
    // waiting for the tile so the L2 round trip hides under the cp.async wait and the QK^T MMAs
    const float* bias = p.bias + (long long)head * 64 * 64;
    float2 bv0[8], bv1[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bv0[nt] = __ldg(reinterpret_cast<const float2*>(bias + row0 * 64 + nt * 8 + 2 * t));
      bv1[nt] = __ldg(reinterpret_cast<const float2*>(bias + (row0 + 8) * 64 + nt * 8 + 2 * t));
    }
    if (hi + 1 < p.hpc) { stage_head(head + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const __half* sQ = sbuf + buf * buf_halves;
    const __half* sK = sQ + 64 * kAttnPad;
    const __half* sV = sK + 64 * kAttnPad;

    // S = Q K^T : Q fragments for the two k-steps (d 0..15, 16..31)
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
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      const float2 b0 = bv0[nt], b1 = bv1[nt];
      float m0 = 0.f, m1 = 0.f;
      if (p.shift) {
        if (swin_label(wy, col & 7, p.H, p.shift) != la) m0 = -100.0f;
        if (swin_label(wy, (col + 1) & 7, p.H, p.shift) != la) m1 = -100.0f;
      }
      s[nt][0] = s[nt][0] * p.scale + b0.x + m0;
      s[nt][1] = s[nt][1] * p.scale + b0.y + m1;
      s[nt][2] = s[nt][2] * p.scale + b1.x + m0;
      s[nt][3] = s[nt][3] * p.scale + b1.y + m1;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
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

    // O = P V : k = keys (4 steps of 16), n = d (4 tiles of 8); V[key][d] row-major, read transposed
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
        // lanes 0..15 name the 16 key rows of this k-step (lanes 16..31 are ignored by .x2 but must be valid)
        uint32_t vb[2];
        ldmatrix_x2_trans(vb, &sV[(kk * 16 + (lane & 15)) * kAttnPad + dt * 8]);
        mma_16816(o[dt], pa, vb);
      }
    }
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int d = hi * 32 + dt * 8 + 2 * t;
      *reinterpret_cast<__half2*>(&sOut[row0 * opitch + d]) = __floats2half2_rn(o[dt][0] * inv0, o[dt][1] * inv0);
      *reinterpret_cast<__half2*>(&sOut[(row0 + 8) * opitch + d]) = __floats2half2_rn(o[dt][2] * inv1, o[dt][3] * inv1);
    }
    __syncthreads();      // everyone done with this head's buffer before it is refilled two iterations later
  }
  // write this CTA's 64 x (hpc*32) slice as contiguous row segments
  const int units = p.hpc * 4;
  for (int i = threadIdx.x; i < 64 * units; i += 128) {
    const int tok = i / units, u = i - tok * units;
    *reinterpret_cast<uint4*>(p.out + (long long)sPix[tok] * p.out_ld + head0 * 32 + u * 8) =
        *reinterpret_cast<const uint4*>(&sOut[tok * opitch + u * 8]);
  }
}

// Plain fp32 cross-check: one CTA per (window, head), one thread per query row.
__global__ void __launch_bounds__(64) window_attn_simt_kernel(const WinAttnParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __half sK[64 * 32];
  __shared__ __half sV[64 * 32];
  const int head = blockIdx.y;
  const int nWx = p.W >> 3, nWy = p.H >> 3;
  int win = blockIdx.x;
  const int wx = win % nWx; win /= nWx;
  const int wy = win % nWy; win /= nWy;
  const int n = win;
  const int i = threadIdx.x;
  const int y = (wy * 8 + (i >> 3) + p.shift) % p.H, x = (wx * 8 + (i & 7) + p.shift) % p.W;
  const long long pix = ((long long)n * p.H + y) * p.W + x;
  const __half* row = p.qkv + pix * p.qkv_ld + head * 32;
  float q[32];
  for (int d = 0; d < 32; ++d) {
    q[d] = __half2float(row[d]) * p.scale;
    sK[i * 32 + d] = row[p.E + d];
    sV[i * 32 + d] = row[2 * p.E + d];
  }
  __syncthreads();
  const float* bias = p.bias + (long long)head * 64 * 64;
  const int li = p.shift ? swin_label(wy, i & 7, p.H, p.shift) : 0;
  float sc[64];
  float mx = -1e30f;
  for (int j = 0; j < 64; ++j) {
    float s = 0.f;
    for (int d = 0; d < 32; ++d) s = fmaf(q[d], __half2float(sK[j * 32 + d]), s);
    s += bias[i * 64 + j];
    if (p.shift && swin_label(wy, j & 7, p.H, p.shift) != li) s += -100.0f;
    sc[j] = s; mx = fmaxf(mx, s);
  }
  float sum = 0.f;
  for (int j = 0; j < 64; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.0f / sum;
  __half* dst = p.out + pix * p.out_ld + head * 32;
  for (int d = 0; d < 32; ++d) {
    float o = 0.f;
    for (int j = 0; j < 64; ++j) o = fmaf(sc[j], __half2float(sV[j * 32 + d]), o);
    dst[d] = __float2half_rn(o * inv);
  }
}

#endif
}  // namespace rs
