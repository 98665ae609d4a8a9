// Small HBM/latency-bound kernels around the GEMMs: input packing, nearest upsample, timestep
// embedding MLP, the residual-shift sampling update, and weight repacking.
#pragma once

#include "common.cuh"
#include "gn_stats.cuh"

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
  unsigned int* zero_ptr; int zero_n; // GroupNorm arrival counters of the forward that follows (gn_stats.cuh): reset here
};

__global__ void pack_input_kernel(const PackInputParams p) {
  pdl_trigger();
  pdl_wait();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < p.zero_n) p.zero_ptr[pix] = 0u;
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
  unsigned int* zero_ptr; int zero_n; // GroupNorm arrival counters of the NEXT denoiser forward: reset here
};
__global__ void p_sample_kernel(const PSampleParams p) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.zero_n) p.zero_ptr[i] = 0u;
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
// Split-K finish: out = act(sum_s partial[s] + bias) + residual, fp16 NHWC view (or fp32 NCHW), plus the
// GroupNorm partial statistics of the result.  One CTA per (128-pixel slot, image) — the same slots the
// conv epilogue would have produced; splits are summed in a fixed order, statistics reduced in a fixed tree.
// ------------------------------------------------------------------------------------------------
struct SplitKReduceParams {
  const float* partial;     // [S][N*HW][C]
  int S, N, HW, C;
  const float* bias;
  const __half* residual; long long res_sN; int res_ld;
  __half* out; long long out_sN; int out_ld;
  int act;
  int rows_per_slot, slots;
  int cols_per_cta;         // multiple of 8; grid.z = ceil(C / cols_per_cta)
  GnSink sink[2];           // fused GroupNorm statistics of the result (gn_stats.cuh)
};

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const __grid_constant__ SplitKReduceParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_red[];       // [lanes][cols_per_cta][3] = (rows, mean, M2) per row-lane and column; + 4 flags
  const int c_begin = blockIdx.z * p.cols_per_cta;
  const int ccols = min(p.cols_per_cta, p.C - c_begin);
  const int vecs = ccols >> 3;
  const int lanes = blockDim.x / vecs;
  const int n = blockIdx.y, slot = blockIdx.x;
  const int vec = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  const long long npix = (long long)p.N * p.HW;
  const bool want_stats = p.sink[0].part != nullptr;
  if (rl < lanes) {
    // statistics of the stored values around a pivot (this thread's first row): no cancellation for |mean| >> std
    float pv[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { pv[j] = 0.f; s1[j] = 0.f; s2[j] = 0.f; }
    int cnt = 0;
    const int r0 = slot * p.rows_per_slot, r1 = min(r0 + p.rows_per_slot, p.HW);
    const int c = c_begin + vec * 8;
    float bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bs[j] = p.bias ? p.bias[c + j] : 0.f;
    for (int r = r0 + rl; r < r1; r += lanes) {
      const long long pix = (long long)n * p.HW + r;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int s = 0; s < p.S; ++s) {
        const float4* src = reinterpret_cast<const float4*>(p.partial + ((long long)s * npix + pix) * p.C + c);
        const float4 a = src[0], b = src[1];
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = acc[j] + bs[j];
        if (p.act == 1) v = gelu_erf_f(v); else if (p.act == 2) v = silu_f(v);
        acc[j] = v;
      }
      if (p.residual) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p.residual + n * p.res_sN + (long long)r * p.res_ld + c);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      }
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
      *reinterpret_cast<uint4*>(p.out + n * p.out_sN + (long long)r * p.out_ld + c) = o;
      if (want_stats) {
        float st[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(oh[j]); st[2 * j] = f.x; st[2 * j + 1] = f.y; }
        if (cnt == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = st[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = st[j] - pv[j]; s1[j] += d; s2[j] = fmaf(d, d, s2[j]); }
        }
        ++cnt;
      }
    }
    if (want_stats) {
      float* dst = s_red + ((size_t)rl * ccols + vec * 8) * 3;
      const float inv = cnt ? 1.0f / (float)cnt : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dst[3 * j] = (float)cnt;
        dst[3 * j + 1] = pv[j] + s1[j] * inv;
        dst[3 * j + 2] = fmaxf(s2[j] - s1[j] * s1[j] * inv, 0.f);
      }
    }
  }
  if (!want_stats) return;
  __syncthreads();
  // merge the row-lanes of each column in lane order (Chan et al.), then deliver the slot's pair to the sinks
  for (int i = threadIdx.x; i < ccols; i += blockDim.x) {
    float cn = 0.f, mean = 0.f, m2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      const float* e = s_red + ((size_t)l * ccols + i) * 3;
      const float nb = e[0];
      if (nb == 0.f) continue;
      const float tot = cn + nb, d = e[1] - mean;
      mean += d * (nb / tot);
      m2 += e[2] + d * d * (cn * nb / tot);
      cn = tot;
    }
    const int ch = c_begin + i;
#pragma unroll
    for (int d = 0; d < 2; ++d)
      if (p.sink[d].part) {
        float* dst = p.sink[d].part + (((size_t)n * p.slots + slot) * p.sink[d].cstride + p.sink[d].coff + ch) * 2;
        dst[0] = mean; dst[1] = m2;
      }
  }
  if (!(p.sink[0].gstat || p.sink[1].gstat)) return;
  int* s_flag = reinterpret_cast<int*>(s_red + (size_t)lanes * ccols * 3);
  const GnSink* const sk[2] = {&p.sink[0], p.sink[1].part ? &p.sink[1] : nullptr};
  const int im[2] = {n, n};
  const unsigned int ad[2] = {(unsigned)ccols, (unsigned)ccols};
  gn_arrive<2>(sk, im, ad, p.slots, (float)p.rows_per_slot, threadIdx.x, blockDim.x, 1, s_flag);
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
