// Small kernels of the VQ-GAN bookends and the image I/O edges (everything that is not a conv / GroupNorm / GEMM):
// row softmax of the single-head attention, nearest-codebook quantisation, the two tiny 1x1 convs around the quantiser,
// torch-compatible bicubic upsampling, uint8 <-> [-1, 1] conversion with mask blending, overlap-average tile scatter.
#pragma once

#include "common.cuh"

namespace rs {

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------------------
// softmax over the rows of S [rows][cols] fp16 (row stride ld), in place:  P = softmax(scale * S)
// reference: AttnBlock.forward, ldm/modules/diffusionmodules/model.py:190-192 (w_ * c^-0.5, softmax over keys).
// One CTA per row; each thread keeps its (at most 32) elements in registers between the passes.
// ------------------------------------------------------------------------------------------------
struct SoftmaxParams {
  __half* s; long long ld; int rows, cols; float scale;
};
__global__ void __launch_bounds__(256) softmax_rows_kernel(const SoftmaxParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[8];
  __half* row = p.s + (long long)blockIdx.x * p.ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kMaxVec = 4;                      // 4 x 8 halves per thread: cols <= 8192
  uint4 raw[kMaxVec];
  float v[kMaxVec][8];
  const int nvec = p.cols >> 3;
  float mx = -3.0e38f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int u = tid + i * 256;
    if (u < nvec) {
      raw[i] = *reinterpret_cast<const uint4*>(row + (long long)u * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[i][2 * j] = f.x * p.scale; v[i][2 * j + 1] = f.y * p.scale;
        mx = fmaxf(mx, fmaxf(v[i][2 * j], v[i][2 * j + 1]));
      }
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int u = tid + i * 256;
    if (u < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[i][j] = __expf(v[i][j] - mx); sum += v[i][j]; }
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += s_red[w];          // fixed order
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int u = tid + i * 256;
    if (u < nvec) {
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(v[i][2 * j] * inv, v[i][2 * j + 1] * inv);
      *reinterpret_cast<uint4*>(row + (long long)u * 8) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// y[n, co, hw] = b[co] + sum_ci w[co, ci] * x[n, ci, hw]   (fp32 NCHW in and out, C <= 8): quant_conv of the encoder
// reference: VQModelTorch.encode, ldm/models/autoencoder.py:28-31 (Conv2d(z_channels, embed_dim, 1)).
// ------------------------------------------------------------------------------------------------
struct PointwiseParams {
  const float* x; float* y; const __half* w; int w_ld; const float* b; int Cin, Cout, N, HW;
};
__global__ void pointwise_conv_f32_kernel(const PointwiseParams p) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.N * p.HW) return;
  const int n = (int)(i / p.HW), hw = (int)(i % p.HW);
  float xin[8];
  for (int c = 0; c < p.Cin; ++c) xin[c] = p.x[((long long)n * p.Cin + c) * p.HW + hw];
  for (int co = 0; co < p.Cout; ++co) {
    float acc = p.b[co];
    for (int c = 0; c < p.Cin; ++c) acc = fmaf(__half2float(p.w[co * p.w_ld + c]), xin[c], acc);
    p.y[((long long)n * p.Cout + co) * p.HW + hw] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// VectorQuantizer2.forward (reference ldm/modules/vqvae/quantize.py:271-284) fused with post_quant_conv
// (ldm/models/autoencoder.py:33-38) and the layout change the decoder's first conv wants:
//   idx = argmin_j ( |z|^2 + |e_j|^2 - 2 z.e_j )   (first minimum, fp32)
//   out[pix, :] = post_quant_conv(e_idx)  as NHWC fp16 padded to Cpad channels.
// One thread per latent position; the codebook streams through shared memory in chunks read by the whole CTA.
// ------------------------------------------------------------------------------------------------
struct QuantizeParams {
  const float* z;            // [N, E, HW] fp32
  const float* codebook;     // [n_e, E] fp32
  int n_e, E, N, HW;
  int quantize;              // 0: force_not_quantize (z passes through)
  const __half* pw; int pw_ld; const float* pb; int Cz;   // post_quant_conv [Cz][E] (fp16, row stride pw_ld), bias [Cz]
  __half* out; int Cpad;     // [N*HW, Cpad]
  int* idx_out;              // optional [N, HW]
};
__global__ void __launch_bounds__(256) vq_quantize_kernel(const QuantizeParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_code[];               // [chunk][E + 1]: code, |e|^2
  constexpr int kChunk = 1024;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < (long long)p.N * p.HW;
  const int n = live ? (int)(i / p.HW) : 0, hw = live ? (int)(i % p.HW) : 0;
  float zv[8];
  float zz = 0.f;
  for (int c = 0; c < p.E; ++c) { zv[c] = live ? p.z[((long long)n * p.E + c) * p.HW + hw] : 0.f; }
  for (int c = 0; c < p.E; ++c) zz += zv[c] * zv[c];          // torch.sum(z ** 2, dim=1): sequential over E
  int best = 0;
  if (p.quantize) {
    float bestd = 3.0e38f;
    const int stride = p.E + 1;
    for (int j0 = 0; j0 < p.n_e; j0 += kChunk) {
      const int cnt = min(kChunk, p.n_e - j0);
      __syncthreads();
      for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
        float ee = 0.f;
        for (int c = 0; c < p.E; ++c) { const float e = p.codebook[(long long)(j0 + t) * p.E + c]; s_code[t * stride + c] = e; ee += e * e; }
        s_code[t * stride + p.E] = ee;
      }
      __syncthreads();
      for (int t = 0; t < cnt; ++t) {
        const float* e = s_code + t * stride;                  // broadcast reads
        float dot = 0.f;
        for (int c = 0; c < p.E; ++c) dot = fmaf(zv[c], e[c], dot);
        const float d = (zz + e[p.E]) - 2.0f * dot;
        if (d < bestd) { bestd = d; best = j0 + t; }           // strict <: the first minimum wins, like torch.argmin
      }
    }
  }
  if (!live) return;
  float q[8];
  for (int c = 0; c < p.E; ++c) q[c] = p.quantize ? p.codebook[(long long)best * p.E + c] : zv[c];
  if (p.idx_out) p.idx_out[i] = p.quantize ? best : -1;
  __half* o = p.out + i * p.Cpad;
  int co = 0;
  for (; co < p.Cz; ++co) {
    float acc = p.pb[co];
    for (int c = 0; c < p.E; ++c) acc = fmaf(__half2float(p.pw[co * p.pw_ld + c]), q[c], acc);
    o[co] = __float2half_rn(acc);
  }
  for (; co < p.Cpad; ++co) o[co] = __float2half_rn(0.f);
}

// ------------------------------------------------------------------------------------------------
// Bicubic upsampling by an integer factor, fp32 NCHW, identical to F.interpolate(mode='bicubic', align_corners=False)
// (A = -0.75, source index (dst + 0.5) / sf - 0.5, border pixels clamped) — the pre-upsample of encode_first_stage
// (reference models/gaussian_diffusion.py:503-504).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_weights(float t, float (&w)[4]) {
  const float A = -0.75f;
  // same polynomial forms as ATen's cubic_convolution1 / cubic_convolution2 (UpSample.h)
  auto c1 = [&](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };
  auto c2 = [&](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };
  w[0] = c2(t + 1.f); w[1] = c1(t); w[2] = c1(1.f - t); w[3] = c2(2.f - t);
}
struct BicubicParams { const float* x; float* y; int NC, H, W, sf; };
__global__ void bicubic_upsample_kernel(const BicubicParams p) {
  pdl_trigger();
  pdl_wait();
  const int OW = p.W * p.sf, OH = p.H * p.sf;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.NC * OH * OW) return;
  const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
  const long long nc = i / ((long long)OW * OH);
  const float scale = 1.0f / (float)p.sf;
  const float sx = scale * ((float)ox + 0.5f) - 0.5f, sy = scale * ((float)oy + 0.5f) - 0.5f;
  const float fx = floorf(sx), fy = floorf(sy);
  const int ix = (int)fx, iy = (int)fy;
  float wx[4], wy[4];
  cubic_weights(sx - fx, wx);
  cubic_weights(sy - fy, wy);
  const float* src = p.x + nc * (long long)p.H * p.W;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), p.H - 1);
    float r = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int xx = min(max(ix - 1 + b, 0), p.W - 1);
      r += src[(long long)yy * p.W + xx] * wx[b];
    }
    acc += r * wy[a];
  }
  p.y[i] = acc;
}

__global__ void zero_u32_kernel(unsigned int* ptr, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ptr[i] = 0u;
}

// ------------------------------------------------------------------------------------------------
// Image I/O edges (reference sampler.py:218-223,286 + utils/util_image.py:216-273 tensor2img / imwrite):
//   ingest:  uint8 HWC (RGB, 1 or 3 channels) -> fp32 NCHW in [-1, 1]      ((v / 255 - 0.5) / 0.5)
//   emit:    fp32 NCHW in [-1, 1] -> clamp -> * 0.5 + 0.5 -> optional mask-back blend with the LQ image -> uint8 HWC
//            (round(v * 255), as tensor2img does: (x * 255).round() of the [0, 1]-clamped value), RGB or BGR order
// ------------------------------------------------------------------------------------------------
struct IngestParams { const uint8_t* src; float* dst; int N, H, W, C; };
__global__ void ingest_u8_kernel(const IngestParams p) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over N*H*W pixels
  if (i >= (long long)p.N * p.H * p.W) return;
  const long long n = i / ((long long)p.H * p.W), hw = i % ((long long)p.H * p.W);
  for (int c = 0; c < p.C; ++c) {
    const float v = (float)p.src[i * p.C + c] / 255.0f;
    p.dst[(n * p.C + c) * (long long)p.H * p.W + hw] = (v - 0.5f) / 0.5f;
  }
}
struct EmitParams {
  const float* sr;           // [N, 3, H, W] in [-1, 1] (un-clamped)
  const float* lq;           // optional [N, 3, H, W] in [-1, 1]: mask-back source
  const float* mask;         // optional [N, 1, H, W] in [-1, 1] (1 = unknown area: keep the model output there)
  uint8_t* dst;              // [N, H, W, 3]
  int N, H, W, bgr;
};
__global__ void emit_u8_kernel(const EmitParams p) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)p.H * p.W;
  if (i >= (long long)p.N * HW) return;
  const long long n = i / HW, hw = i % HW;
  const float m = p.mask ? p.mask[n * HW + hw] * 0.5f + 0.5f : 1.0f;
  for (int c = 0; c < 3; ++c) {
    float v = fminf(fmaxf(p.sr[(n * 3 + c) * HW + hw], -1.0f), 1.0f) * 0.5f + 0.5f;
    if (p.mask) v = v * m + (p.lq[(n * 3 + c) * HW + hw] * 0.5f + 0.5f) * (1.0f - m);
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    p.dst[i * 3 + (p.bgr ? 2 - c : c)] = (uint8_t)__float2int_rn(v * 255.0f);
  }
}

// ------------------------------------------------------------------------------------------------
// Overlap-average of tiled results (reference ImageSpliterTh.update / gather, utils/util_image.py:962-979): every output
// pixel is the mean of the tiles that cover it.  Gather form (no atomics, deterministic): one thread per output pixel
// walks the tile grid; tiles are [T, N, C, th, tw] with tile t = ty * ntx + tx at (ys[ty], xs[tx]) in output pixels.
// ------------------------------------------------------------------------------------------------
struct TileGatherParams {
  const float* tiles; float* out;
  int N, C, H, W;            // output
  int th, tw, nty, ntx;
  const int* ys; const int* xs;   // [nty], [ntx] tile origins (output pixels), ascending
};
__global__ void tile_gather_kernel(const TileGatherParams p) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)p.H * p.W;
  if (i >= (long long)p.N * p.C * HW) return;
  const int x = (int)(i % p.W), y = (int)((i / p.W) % p.H);
  const long long nc = i / HW;
  const long long tstride = (long long)p.N * p.C * p.th * p.tw;
  float acc = 0.f;
  int cnt = 0;
  for (int ty = 0; ty < p.nty; ++ty) {
    const int ly = y - p.ys[ty];
    if (ly < 0 || ly >= p.th) continue;
    for (int tx = 0; tx < p.ntx; ++tx) {
      const int lx = x - p.xs[tx];
      if (lx < 0 || lx >= p.tw) continue;
      acc += p.tiles[(long long)(ty * p.ntx + tx) * tstride + (nc * p.th + ly) * p.tw + lx];
      ++cnt;
    }
  }
  p.out[i] = acc / (float)cnt;
}

#endif
}  // namespace rs
