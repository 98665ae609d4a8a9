// Implicit-GEMM convolution / linear layer on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
//   D[M = N*H*W pixels, Cout] = sum over taps, channels  A[pixel + tap offset, c] * Wt[cout, tap, c]
//
// * activations are NHWC fp16 *views* (channel count C, row stride ld >= C): a 4-D TMA tensor map
//   {C, W, H, N} per source.  One CTA computes a 128-pixel x BN-channel output tile; the 128 pixels
//   are a (bw x bh x bn) box in (W, H, N), so the tile of a 3x3 tap is the same box shifted by
//   (dw, dh) — TMA's out-of-bounds zero fill implements the conv padding, and the box lands in
//   shared memory exactly as the K-major, 128-byte-swizzled operand tcgen05.mma expects.
// * stride-2 convs read the four (row, column)-parity sub-grids of the input through four strided
//   tensor maps over the same buffer (no copy); each tap names its map.
// * weights are [Cout][tap][Cin_pad] fp16 (K-major), one 2-D tensor map.
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//   warps 2..5 = epilogue (TMEM -> registers -> bias / activation / residual -> global).
// * covers reference call sites: every nn.Conv2d / nn.Linear inside UNetModelSwin.forward
//   (reference models/unet.py:147,173,184,707,862,69,99-101; models/swin_transformer.py:22-24,105-107,480,515).
#pragma once

#include "common.cuh"

namespace rs {

constexpr int kConvBM = 128;      // pixels per tile (UMMA M)
constexpr int kConvBK = 64;       // channels per k-block (128 B rows, SWIZZLE_128B)
constexpr int kConvThreads = 192;
constexpr int kMaxTaps = 9;
constexpr int kMaxSrc = 4;

enum ConvAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

struct ConvParams {
  CUtensorMap tmA[kMaxSrc];
  CUtensorMap tmB;
  int num_taps;
  int tap_src[kMaxTaps];
  int tap_dh[kMaxTaps];
  int tap_dw[kMaxTaps];
  int kchunks;           // ceil(Cin / 64)
  int w_tap_stride;      // K offset between consecutive taps in the weight matrix (= Cin_pad)
  int bw, bh, bn;        // pixel box of a tile, bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_n;
  int Wout, Hout, Nimg;
  int BN, n_tiles, Cout;
  int stages;
  int tmem_cols;
  // epilogue
  const float* bias;                 // [Cout] fp32 or nullptr
  const __half* residual;            // optional, same pixel grid as the output
  long long res_sN, res_sH, res_sW;  // strides in elements
  __half* out;                       // NHWC fp16 view (may be nullptr when out_f32 is set)
  long long out_sN, out_sH, out_sW;
  float* out_f32_nchw;               // optional fp32 NCHW output [Nimg, Cout, Hout, Wout]
  int act;
};

#ifdef __CUDACC__

__global__ void __launch_bounds__(kConvThreads, 2) conv_gemm_sm100_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B BN*128 B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = kConvBM * kConvBK * 2;
  const int b_bytes = p.BN * kConvBK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int n_tile = blockIdx.x % p.n_tiles;
  int mt = blockIdx.x / p.n_tiles;
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  const int tn = mt;
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
  const int num_kb = p.num_taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kMaxSrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_dyn(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / p.kchunks;
        const int kc = kb - tap * p.kchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
        tma_load_4d(sa, &p.tmA[p.tap_src[tap]], &full_bar[stage], kc * kConvBK, w0 + p.tap_dw[tap],
                    h0 + p.tap_dh[tap], n0);
        tma_load_2d(sb, &p.tmB, &full_bar[stage], tap * p.w_tap_stride + kc * kConvBK, n_tile * p.BN);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_f16(kConvBM, p.BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + a_bytes);
#pragma unroll
        for (int k = 0; k < kConvBK / 16; ++k) {
          // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);                 // frees this smem stage when the MMAs retire
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
    const int quad = warp & 3;                       // warps 2,3,4,5 -> quadrants 2,3,0,1
    const int r = quad * 32 + lane;                  // row of the tile == TMEM lane
    const int lw = r % p.bw;
    const int lh = (r / p.bw) % p.bh;
    const int ln = r / (p.bw * p.bh);
    const int w = w0 + lw, h = h0 + lh, n = n0 + ln;
    const bool row_ok = (w < p.Wout) && (h < p.Hout) && (n < p.Nimg);
    const int col0 = n_tile * p.BN;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    __half* orow = p.out ? p.out + n * p.out_sN + h * p.out_sH + w * p.out_sW : nullptr;
    const __half* rrow = p.residual ? p.residual + n * p.res_sN + h * p.res_sH + w * p.res_sW : nullptr;

    for (int c = 0; c < p.BN; c += 16) {
      uint32_t v[16];
      __syncwarp();   // tcgen05.ld is warp-collective: reconverge after the divergent tail of the last chunk
      tmem_ld16(trow + c, v);
      tmem_ld_wait();
      const int col = col0 + c;
      if (!row_ok || col >= p.Cout) continue;
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
      const bool full = (col + 16 <= p.Cout);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full || col + j < p.Cout) f[j] += __ldg(p.bias + col + j);
      }
      if (p.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = gelu_erf_f(f[j]);
      } else if (p.act == ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = silu_f(f[j]);
      }
      if (rrow) {
        if (full) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(rrow + col);
          const uint4 r1 = *reinterpret_cast<const uint4*>(rrow + col + 8);
          const __half2* h0p = reinterpret_cast<const __half2*>(&r0);
          const __half2* h1p = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __half22float2(h0p[j]);
            const float2 b = __half22float2(h1p[j]);
            f[2 * j] += a.x; f[2 * j + 1] += a.y;
            f[8 + 2 * j] += b.x; f[8 + 2 * j + 1] += b.y;
          }
        } else {
          for (int j = 0; j < 16 && col + j < p.Cout; ++j) f[j] += __half2float(rrow[col + j]);
        }
      }
      if (p.out_f32_nchw) {
        for (int j = 0; j < 16 && col + j < p.Cout; ++j)
          p.out_f32_nchw[(((long long)n * p.Cout + (col + j)) * p.Hout + h) * p.Wout + w] = f[j];
      }
      if (orow) {
        if (full) {
          uint4 o0, o1;
          __half2* q0 = reinterpret_cast<__half2*>(&o0);
          __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            q0[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
            q1[j] = __floats2half2_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
          }
          *reinterpret_cast<uint4*>(orow + col) = o0;
          *reinterpret_cast<uint4*>(orow + col + 8) = o1;
        } else {
          for (int j = 0; j < 16 && col + j < p.Cout; ++j) orow[col + j] = __float2half_rn(f[j]);
        }
      }
    }
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Plain SIMT implementation of the same operator (debug / cross-check path, selected with
// RS_CONV_IMPL=simt).  Same ConvParams epilogue fields; sources passed as raw views.
// ------------------------------------------------------------------------------------------------
struct ConvSimtSrc {
  const __half* ptr[kMaxSrc];
  long long sN[kMaxSrc], sH[kMaxSrc], sW[kMaxSrc];
  int H[kMaxSrc], W[kMaxSrc];
  int C;
  const __half* wt;     // [Cout][taps][w_tap_stride]
};

__global__ void conv_simt_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ ConvSimtSrc s) {
  // one warp per output pixel, lanes stride over output channels
  const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long npix = (long long)p.Nimg * p.Hout * p.Wout;
  if (pix >= npix) return;
  const int w = (int)(pix % p.Wout);
  const int h = (int)((pix / p.Wout) % p.Hout);
  const int n = (int)(pix / ((long long)p.Wout * p.Hout));
  const int ktot = p.num_taps * p.w_tap_stride;
  for (int co = lane; co < p.Cout; co += 32) {
    float acc = 0.f;
    for (int t = 0; t < p.num_taps; ++t) {
      const int src = p.tap_src[t];
      const int hh = h + p.tap_dh[t], ww = w + p.tap_dw[t];
      if (hh < 0 || ww < 0 || hh >= s.H[src] || ww >= s.W[src]) continue;
      const __half* a = s.ptr[src] + n * s.sN[src] + hh * s.sH[src] + ww * s.sW[src];
      const __half* wr = s.wt + (long long)co * ktot + t * p.w_tap_stride;
      for (int c = 0; c < s.C; c += 2) {
        const float2 av = __half22float2(*reinterpret_cast<const __half2*>(a + c));
        const float2 wv = __half22float2(*reinterpret_cast<const __half2*>(wr + c));
        acc = fmaf(av.x, wv.x, acc);
        acc = fmaf(av.y, wv.y, acc);
      }
    }
    if (p.bias) acc += p.bias[co];
    if (p.act == ACT_GELU) acc = gelu_erf_f(acc);
    else if (p.act == ACT_SILU) acc = silu_f(acc);
    if (p.residual) acc += __half2float(p.residual[n * p.res_sN + h * p.res_sH + w * p.res_sW + co]);
    if (p.out_f32_nchw) p.out_f32_nchw[(((long long)n * p.Cout + co) * p.Hout + h) * p.Wout + w] = acc;
    if (p.out) p.out[n * p.out_sN + h * p.out_sH + w * p.out_sW + co] = __float2half_rn(acc);
  }
}

#endif  // __CUDACC__

}  // namespace rs
