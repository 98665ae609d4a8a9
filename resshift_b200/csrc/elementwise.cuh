// Small HBM/latency-bound kernels around the GEMMs: input packing, nearest upsample, timestep
// embedding MLP, the residual-shift sampling update, and weight repacking.
#pragma once

#include "common.cuh"

namespace rs {

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------------------
// UNet input: cat([x * in_scale, lq], dim=1) in NCHW fp32  ->  NHWC fp16 with channels padded to Cpad.
// reference: _scale_input (models/gaussian_diffusion.py:598-603) + th.cat (models/unet.py:882).
// The LQ part is either raw NCHW fp32 (realsr: feature_extractor is Identity, unet.py:689-691) or an
// NHWC fp16 feature map produced by the feature extractor.
// ------------------------------------------------------------------------------------------------
struct PackInputParams {
  const float* x; int Cx;             // [N, Cx, H, W] fp32
  const float* scale_tab; int scale_idx;   // optional per-step table; value 1/sqrt(eta*kappa^2+1)
  const float* lq_nchw; int Cl;       // [N, Cl, H, W] fp32 or nullptr
  const __half* lq_nhwc; int lq_ld;   // [N*H*W, Cl] fp16 or nullptr
  __half* out; int Cpad;              // [N*H*W, Cpad]
  int N, HW;
};

__global__ void pack_input_kernel(const PackInputParams p) {
  pdl_trigger();
  pdl_wait();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.N * p.HW;
  if (pix >= total) return;
  const int n = (int)(pix / p.HW);
  const int hw = (int)(pix % p.HW);
  const float sc = p.scale_tab ? p.scale_tab[p.scale_idx] : 1.0f;
  __half* o = p.out + pix * p.Cpad;
  int c = 0;
  for (; c < p.Cx; ++c) o[c] = __float2half_rn(p.x[((long long)n * p.Cx + c) * p.HW + hw] * sc);
  if (p.lq_nchw) {
    for (int j = 0; j < p.Cl; ++j, ++c) o[c] = __float2half_rn(p.lq_nchw[((long long)n * p.Cl + j) * p.HW + hw]);
  } else if (p.lq_nhwc) {
    for (int j = 0; j < p.Cl; ++j, ++c) o[c] = p.lq_nhwc[pix * p.lq_ld + j];
  }
  for (; c < p.Cpad; ++c) o[c] = __float2half_rn(0.f);
}

// NCHW fp32 image (+ optional mask) -> NHWC fp16, channels padded: input of the feature extractor
// (reference models/unet.py:876-881: th.cat([lq, mask], dim=1)).
struct PackImageParams {
  const float* a; int Ca;
  const float* b; int Cb;
  __half* out; int Cpad;
  int N, HW;
};
__global__ void pack_image_kernel(const PackImageParams p) {
  pdl_trigger();
  pdl_wait();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)p.N * p.HW) return;
  const int n = (int)(pix / p.HW);
  const int hw = (int)(pix % p.HW);
  __half* o = p.out + pix * p.Cpad;
  int c = 0;
  for (int j = 0; j < p.Ca; ++j, ++c) o[c] = __float2half_rn(p.a[((long long)n * p.Ca + j) * p.HW + hw]);
  for (int j = 0; j < p.Cb; ++j, ++c) o[c] = __float2half_rn(p.b[((long long)n * p.Cb + j) * p.HW + hw]);
  for (; c < p.Cpad; ++c) o[c] = __float2half_rn(0.f);
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample, NHWC fp16 (reference Upsample.forward, models/unet.py:71-81)
// ------------------------------------------------------------------------------------------------
struct UpsampleParams {
  const __half* x; long long x_sN; int x_ld;   // [N, H, W, C] view
  __half* y;                                   // [N, 2H, 2W, C] contiguous
  int N, H, W, C;
};
__global__ void upsample2x_kernel(const UpsampleParams p) {
  pdl_trigger();
  pdl_wait();
  const int vecs = p.C >> 3;
  const long long total = (long long)p.N * (2 * p.H) * (2 * p.W) * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long q = i / vecs;
    const int ox = (int)(q % (2 * p.W)); q /= 2 * p.W;
    const int oy = (int)(q % (2 * p.H)); q /= 2 * p.H;
    const int n = (int)q;
    const uint4 raw = *reinterpret_cast<const uint4*>(p.x + n * p.x_sN + ((long long)(oy >> 1) * p.W + (ox >> 1)) * p.x_ld + v * 8);
    *reinterpret_cast<uint4*>(p.y + (((long long)n * 2 * p.H + oy) * 2 * p.W + ox) * p.C + v * 8) = raw;
  }
}

// ------------------------------------------------------------------------------------------------
// Timestep path (reference timestep_embedding, models/basic_ops.py:99-117; time_embed, models/unet.py:683-687;
// ResBlock.emb_layers = SiLU -> Linear, models/unet.py:161-167).  Tiny: one warp per output element.
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half_dim = dim / 2;
  if (i >= B * half_dim) return;
  const int b = i / half_dim, k = i % half_dim;
  const float freq = expf(-logf(10000.0f) * (float)k / (float)half_dim);
  const float a = t[b] * freq;
  out[(long long)b * dim + k] = cosf(a);
  out[(long long)b * dim + half_dim + k] = sinf(a);
  if ((dim & 1) && k == 0) out[(long long)b * dim + dim - 1] = 0.f;
}

// out[b, o] = bias[o] + sum_k act(x[b, k]) * W[o, k];  W fp16 row-major [O, K], x/out fp32.
__global__ void linear_small_kernel(const float* __restrict__ x, const __half* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ out, int B, int K, int O,
                                    int silu_in, int silu_out) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * O) return;
  const int b = warp / O, o = warp % O;
  const float* xr = x + (long long)b * K;
  const __half* wr = W + (long long)o * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float xv = xr[k];
    if (silu_in) xv = silu_f(xv);
    acc = fmaf(xv, __half2float(wr[k]), acc);
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    acc += bias[o];
    if (silu_out) acc = silu_f(acc);
    out[(long long)b * O + o] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Residual-shift sampling update (reference p_sample, models/gaussian_diffusion.py:332-365, with
// q_posterior_mean_variance :210-232):
//     x_{t-1} = coef1[t] * x_t + coef2[t] * x0_pred + [t != 0] * std[t] * noise
// fp32 NCHW in/out.  When `next_in` is set it also emits the NEXT denoiser input
// cat([x_{t-1} * in_scale[t-1], lq]) as NHWC fp16 (fusing _scale_input + th.cat + layout change).
// ------------------------------------------------------------------------------------------------
struct PSampleParams {
  const float* x_t;       // [N, C, HW]
  const float* x0;        // [N, C, HW]
  const float* noise;     // [N, C, HW]
  float* x_next;          // [N, C, HW]
  const float* coef1; const float* coef2; const float* stdv; const float* in_scale;   // [T] fp32 tables
  int t;                  // schedule index of THIS step (T-1 .. 0)
  int N, C, HW;
  __half* next_in; int next_cpad;     // optional: [N*HW, next_cpad]; channels [0, C) are written here
};
__global__ void p_sample_kernel(const PSampleParams p) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.N * p.C * p.HW;
  if (i >= total) return;
  const float c1 = p.coef1[p.t], c2 = p.coef2[p.t];
  const float sd = p.t != 0 ? p.stdv[p.t] : 0.f;
  float v = c1 * p.x_t[i] + c2 * p.x0[i];
  if (p.t != 0) v += sd * p.noise[i];
  p.x_next[i] = v;
  if (p.next_in && p.t > 0) {
    const int hw = (int)(i % p.HW);
    const int c = (int)((i / p.HW) % p.C);
    const int n = (int)(i / ((long long)p.HW * p.C));
    p.next_in[((long long)n * p.HW + hw) * p.next_cpad + c] = __float2half_rn(v * p.in_scale[p.t - 1]);
  }
}

// prior_sample (reference models/gaussian_diffusion.py:517-529): x_T = z_y + kappa*sqrt_eta_T * noise
__global__ void prior_sample_kernel(const float* __restrict__ zy, const float* __restrict__ noise,
                                    float* __restrict__ out, float coef, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = zy[i] + coef * noise[i];
}

// ------------------------------------------------------------------------------------------------
// Weight repacking (load time): fp32 OIHW -> fp16 [O][kh][kw][Ipad]; fp32 [O, I] -> fp16 [O][Ipad]
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ src, __half* __restrict__ dst, int O, int I,
                                        int KH, int KW, int Ipad) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)O * KH * KW * Ipad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ipad);
    long long q = i / Ipad;
    const int kw = (int)(q % KW); q /= KW;
    const int kh = (int)(q % KH); q /= KH;
    const int o = (int)q;
    float v = 0.f;
    if (c < I) v = src[(((long long)o * I + c) * KH + kh) * KW + kw];
    dst[i] = __float2half_rn(v);
  }
}

// relative_position_bias_table [(2w-1)^2, heads] -> dense [heads][64][64] fp32 (window 8)
// (reference models/swin_transformer.py:93-103 for the index, :127-130 for the gather)
__global__ void expand_relpos_kernel(const float* __restrict__ table, float* __restrict__ dst, int heads) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= heads * 64 * 64) return;
  const int h = i / 4096, r = (i / 64) % 64, c = i % 64;
  const int dy = (r >> 3) - (c >> 3) + 7, dx = (r & 7) - (c & 7) + 7;
  dst[i] = table[(dy * 15 + dx) * heads + h];
}

__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

#endif
}  // namespace rs
