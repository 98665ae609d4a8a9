// Fused Swin MLP:  out = residual + fc2( GELU( fc1(x) ) )  in ONE kernel, hidden activations never leave the SM.
//
// reference: Mlp.forward (models/swin_transformer.py:27-33: 1x1 conv E -> 4E, exact-erf GELU, 1x1 conv 4E -> E) and the
// residual add around it in SwinTransformerBlock.forward (:279).  Unfused this is two GEMMs with a [pixels, 4E] fp16
// intermediate (100 MB at batch 16, 64x64) written to and re-read from HBM; here a CTA owns a 128-pixel tile and walks
// the hidden dimension in chunks of 128:
//
//     acc1[128 x 128]  = X[128 x E] . W1_j^T            (tcgen05, TMEM, double buffered over chunks j)
//     H_j              = GELU(acc1 + b1_j)  -> fp16, written to shared memory AS THE NEXT MMA's A OPERAND
//                                              (K-major, 128-byte swizzle: exactly what a TMA load would have produced)
//     acc2[128 x E]   += H_j . W2_j^T                    (tcgen05, TMEM, accumulates over all chunks)
//
// then the usual staged epilogue (+ b2, + residual through a TMA load, fp16, TMA store, fused GroupNorm partials).
// Warp roles: warp 0 = TMA producers (lane 0: X once, then the W1_j tiles; lane 1: the W2_j tiles; one ring each),
// warp 1 = MMA issuer, warps 2..17 = GELU stage + final epilogue (each warp owns a 32-row x 16-column slice of every
// chunk and signals the MMA warp on its own: no CTA-wide barrier in the loop).  MMA1 of chunk j+1 overlaps the GELU
// stage of chunk j.  CTAs of a cluster (4 when the grid allows) each fetch a quarter of every weight tile and multicast it.
#pragma once

#include "common.cuh"
#include "conv_gemm.cuh"

namespace rs {

constexpr int kMlpHc = 128;          // hidden columns per chunk (64-column chunks make the MMA issue rate the bottleneck)
constexpr int kMlpEpiWarps = 16;     // four warps per TMEM lane quadrant: the GELU stage is instruction-bound
constexpr int kMlpThreads = 96 + 32 * kMlpEpiWarps;   // warps 0 / 2: TMA producers, warp 1: MMA issuer, warps 3..18: GELU + epilogue

struct MlpParams {
  CUtensorMap tmX, tmW1, tmW2, tmOut, tmRes;
  const float* bias1;                // [Hd]
  const float* bias2;                // [E]
  int E, Hd;                         // E % 64 == 0, E <= 256;  Hd % 64 == 0
  int ring1, ring2;                  // ring depths: fc1 weight tiles (kMlpHc x 64), fc2 weight tiles (E x 64)
  int cluster;                       // CTAs per cluster sharing (multicasting) the weight tiles: 1, 2 or 4
  int bw, bh, bn, tiles_w, tiles_h;
  int Wout, Hout, Nimg;
  int has_res;
  float* gn_part[2]; int gn_cstride[2]; int gn_coff[2]; int gn_slots;
  long long* dbg;                    // optional: CTA 0 writes a clock64 timeline [64 chunks][8] (profiling aid)
};

#ifdef __CUDACC__

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_fused_sm100_kernel(const __grid_constant__ MlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kx = p.E >> 6;                         // k-blocks of the first GEMM
  const int chunks = p.Hd / kMlpHc;
  constexpr int kTile = kConvBM * kConvBK * 2;     // 16 KB: 128 rows x 64 fp16
  constexpr int kW1 = kMlpHc * kConvBK * 2;        // fc1 weight tile: kMlpHc rows x 64 fp16
  constexpr int kHT = kMlpHc / kConvBK;            // 64-column tiles per hidden chunk (= k-blocks of the second GEMM)
  uint8_t* sX = smem;                              // kx tiles
  const int slot2 = p.E * 128;                     // fc2 weight tile: E rows x 64 fp16
  uint8_t* sW1 = sX + (size_t)kx * kTile;          // ring1 x kW1     (two rings, a producer warp each: neither weight
  uint8_t* sW2 = sW1 + (size_t)p.ring1 * kW1;      // ring2 x slot2    stream ever waits behind the other one's slots)
  uint8_t* sH = sW2 + (size_t)p.ring2 * slot2;     // 2 buffers x kHT tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + 2 * kHT * kTile);
  uint64_t* w1_full = bars;
  uint64_t* w1_empty = w1_full + p.ring1;
  uint64_t* w2_full = w1_empty + p.ring1;
  uint64_t* w2_empty = w2_full + p.ring2;
  uint64_t* x_full = w2_empty + p.ring2;
  uint64_t* acc1_full = x_full + 1;                // [2]
  uint64_t* acc1_empty = acc1_full + 2;            // [2]
  uint64_t* h_full = acc1_empty + 2;               // [2]
  uint64_t* h_empty = h_full + 2;                  // [2]
  uint64_t* acc2_full = h_empty + 2;
  uint64_t* res_bar = acc2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  float* s_b1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [Hd] fc1 bias
  float* s_b2 = s_b1 + p.Hd;                                                          // [E]  fc2 bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
  const long long t_start = clock64();
  // weight tiles are shared by the CTAs of a cluster: each CTA fetches 1/CS of every tile and multicasts it to all
  // (every CTA re-streams both weight matrices for its 128 pixels — without sharing the L2 -> SM traffic of the 148
  // concurrent CTAs, all on the same lines, is what bounds the kernel)
  const int CS = p.cluster;
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << CS) - 1);
  int mt = blockIdx.x;
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = mt * p.bn;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmOut); if (p.has_res) tma_prefetch_desc(&p.tmRes);
    for (int s = 0; s < p.ring1; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], CS); }   // a slot is free when
    for (int s = 0; s < p.ring2; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], CS); }   // EVERY CTA has read it
    mbar_init(x_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1); mbar_init(&h_empty[b], 1);                         // tcgen05.commit
      mbar_init(&acc1_empty[b], kMlpEpiWarps); mbar_init(&h_full[b], kMlpEpiWarps);   // one arrival per GELU warp
    }
    mbar_init(acc2_full, 1); mbar_init(res_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc_dyn(tmem_slot, 512u); tmem_relinquish(); }
  tc_fence_before();
  if (CS > 1) cluster_sync_all(); else __syncthreads();      // peers' barriers exist before any multicast / remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc2 = tmem_base + 2 * kMlpHc;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer 1: X once, then the fc1 weight stream =====================
    // (the two weight streams have a warp each: a lane blocked in mbarrier.try_wait stalls its whole warp)
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, (uint32_t)(kx * kTile));
      for (int kb = 0; kb < kx; ++kb) tma_load_4d(sX + (size_t)kb * kTile, &p.tmX, x_full, kb * kConvBK, w0, h0, n0);
      int st1 = 0; uint32_t ph1 = 0;
      const int rows = kMlpHc / CS;                // rows of every tile this CTA fetches
      for (int j = 0; j < chunks; ++j)
        for (int kb = 0; kb < kx; ++kb) {
          mbar_wait(&w1_empty[st1], ph1 ^ 1);
          mbar_arrive_expect_tx(&w1_full[st1], (uint32_t)kW1);
          uint8_t* dst = sW1 + (size_t)st1 * kW1 + (size_t)rank * rows * 128;
          if (CS > 1) tma_load_2d_mc(dst, &p.tmW1, &w1_full[st1], kb * kConvBK, j * kMlpHc + (int)rank * rows, cmask);
          else tma_load_2d(dst, &p.tmW1, &w1_full[st1], kb * kConvBK, j * kMlpHc);
          if (++st1 == p.ring1) { st1 = 0; ph1 ^= 1; }
        }
    }
  } else if (warp == 2) {
    // ===================== TMA producer 2: the fc2 weight stream =====================
    if (lane == 0) {
      int st2 = 0; uint32_t ph2 = 0;
      const int rows = p.E / CS;
      for (int j = 0; j < chunks; ++j)
        for (int t = 0; t < kHT; ++t) {
          mbar_wait(&w2_empty[st2], ph2 ^ 1);
          mbar_arrive_expect_tx(&w2_full[st2], (uint32_t)slot2);
          uint8_t* dst = sW2 + (size_t)st2 * slot2 + (size_t)rank * rows * 128;
          if (CS > 1) tma_load_2d_mc(dst, &p.tmW2, &w2_full[st2], j * kMlpHc + t * kConvBK, (int)rank * rows, cmask);
          else tma_load_2d(dst, &p.tmW2, &w2_full[st2], j * kMlpHc + t * kConvBK, 0);
          if (++st2 == p.ring2) { st2 = 0; ph2 ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // order: GEMM1(0), GEMM1(1), then per chunk j: GEMM2(j) as soon as H_j is written, then GEMM1(j + 2) into the
    // accumulator buffer GELU(j) has just released — the GELU warps always find the next chunk's acc1 ready.
    const uint32_t idesc1 = umma_idesc_f16(kConvBM, kMlpHc);
    const uint32_t idesc2 = umma_idesc_f16(kConvBM, p.E);
    int st1 = 0, st2 = 0; uint32_t ph1 = 0, ph2 = 0;
    auto gemm1 = [&](int c) {
      const int b = c & 1;
      mbar_wait(&acc1_empty[b], ((c >> 1) & 1) ^ 1);
      tc_fence_after();
      if (dbg && lane == 0) dbg[c * 8 + 0] = clock64() - t_start;
      for (int kb = 0; kb < kx; ++kb) {
        mbar_wait(&w1_full[st1], ph1);
        tc_fence_after();
        if (dbg && lane == 0 && kb == kx - 1) dbg[c * 8 + 1] = clock64() - t_start;
        if (lane == 0) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sX + (size_t)kb * kTile));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sW1 + (size_t)st1 * kW1));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + b * kMlpHc, adesc + 2 * k, bdesc + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
          if (CS > 1) umma_commit_mc(&w1_empty[st1], cmask); else umma_commit(&w1_empty[st1]);
          if (kb == kx - 1) umma_commit(&acc1_full[b]);
        }
        __syncwarp();
        if (++st1 == p.ring1) { st1 = 0; ph1 ^= 1; }
      }
    };
    mbar_wait(x_full, 0);
    gemm1(0);
    if (chunks > 1) gemm1(1);
    for (int j = 0; j < chunks; ++j) {
      const int b = j & 1;
      mbar_wait(&h_full[b], (j >> 1) & 1);
      tc_fence_after();
      if (dbg && lane == 0) dbg[j * 8 + 2] = clock64() - t_start;
      for (int t = 0; t < kHT; ++t) {
        mbar_wait(&w2_full[st2], ph2);
        tc_fence_after();
        if (dbg && lane == 0 && t == kHT - 1) dbg[j * 8 + 3] = clock64() - t_start;
        if (lane == 0) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sH + (size_t)(b * kHT + t) * kTile));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sW2 + (size_t)st2 * slot2));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tm_acc2, adesc + 2 * k, bdesc + 2 * k, idesc2, (j | t | k) != 0 ? 1u : 0u);
          if (CS > 1) umma_commit_mc(&w2_empty[st2], cmask); else umma_commit(&w2_empty[st2]);
          if (t == kHT - 1) {
            umma_commit(&h_empty[b]);
            if (j == chunks - 1) umma_commit(acc2_full);
          }
        }
        __syncwarp();
        if (++st2 == p.ring2) { st2 = 0; ph2 ^= 1; }
      }
      if (j + 2 < chunks) gemm1(j + 2);
    }
  } else if (warp >= 3) {
    // ===================== GELU stage + final epilogue (16 warps) =====================
    // warp -> TMEM lane quadrant (warp % 4) and one 16-column slice of every 64 hidden columns; warps run
    // independently (per-warp mbarrier arrivals, no CTA-wide barrier inside the chunk loop)
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read (warp id % 4)
    const int cpar = (warp - 3) >> 2;              // 0..3
    const int r = quad * 32 + lane;
    const int etid = threadIdx.x - 96;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    for (int i = etid; i < p.Hd; i += 32 * kMlpEpiWarps) s_b1[i] = __ldg(p.bias1 + i);
    for (int i = etid; i < p.E; i += 32 * kMlpEpiWarps) s_b2[i] = __ldg(p.bias2 + i);
    named_bar_sync(1, 32 * kMlpEpiWarps);
    const int c16 = cpar * 16;
    const int u0 = c16 >> 3;                       // 16-byte unit of this warp's first 8 columns inside a 128-byte row
    for (int j = 0; j < chunks; ++j) {
      const int b = j & 1;
      mbar_wait(&acc1_full[b], (j >> 1) & 1);
      if (dbg && etid == 0) dbg[j * 8 + 4] = clock64() - t_start;
      mbar_wait(&h_empty[b], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      if (dbg && etid == 0) dbg[j * 8 + 5] = clock64() - t_start;
      uint32_t v[kHT][16];
#pragma unroll
      for (int t = 0; t < kHT; ++t) tmem_ld16(tmem_base + lane_base + b * kMlpHc + t * kConvBK + c16, v[t]);   // all in flight
#pragma unroll
      for (int t = 0; t < kHT; ++t) tmem_ld_wait16(v[t]);
#pragma unroll
      for (int t = 0; t < kHT; ++t) {
        const float4* bp = reinterpret_cast<const float4*>(s_b1 + j * kMlpHc + t * kConvBK + c16);
        __half2 q[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = bp[i];
          // (scalar FFMA: the packed fma.rn.f32x2 form issues at half rate — scripts/ubench/gelu_rate.cu — and buys nothing)
          q[2 * i] = __floats2half2_rn(gelu_erf_f(__uint_as_float(v[t][4 * i]) + b4.x), gelu_erf_f(__uint_as_float(v[t][4 * i + 1]) + b4.y));
          q[2 * i + 1] = __floats2half2_rn(gelu_erf_f(__uint_as_float(v[t][4 * i + 2]) + b4.z), gelu_erf_f(__uint_as_float(v[t][4 * i + 3]) + b4.w));
        }
        uint8_t* row = sH + (size_t)(b * kHT + t) * kTile + r * 128;
        *reinterpret_cast<uint4*>(row + (((u0) ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(&q[0]);
        *reinterpret_cast<uint4*>(row + (((u0 + 1) ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(&q[4]);
      }
      if (dbg && etid == 0) dbg[j * 8 + 6] = clock64() - t_start;
      fence_proxy_async_smem();          // H_j will be read by the tensor core through the async proxy
      tc_fence_before();                 // this warp's tcgen05.ld of acc1[b] is complete (wait::ld above)
      __syncwarp();
      if (lane == 0) { mbar_arrive(&acc1_empty[b]); mbar_arrive(&h_full[b]); }
      if (dbg && etid == 0) dbg[j * 8 + 7] = clock64() - t_start;
    }

    // ---- final epilogue: acc2 + bias2 (+ residual) -> fp16 -> TMA store, GroupNorm partials ----
    mbar_wait(acc2_full, 0);
    tc_fence_after();
    if (dbg && etid == 0) dbg[63 * 8 + 0] = clock64() - t_start;
    const int lw = r % p.bw, lh = (r / p.bw) % p.bh, ln = r / (p.bw * p.bh);
    const bool row_ok = (w0 + lw < p.Wout) && (h0 + lh < p.Hout) && (n0 + ln < p.Nimg);
    const int nblk = p.E >> 6;
    uint8_t* sblk = sX;                                  // X is dead: every MMA has retired
    float* wsum = reinterpret_cast<float*>(sH);          // [4 quads][E][2]
    if (p.has_res) {
      if (etid == 0) {
        mbar_arrive_expect_tx(res_bar, (uint32_t)(nblk * kTile));
        for (int bq = 0; bq < nblk; ++bq) tma_load_4d(sblk + (size_t)bq * kTile, &p.tmRes, res_bar, bq * 64, w0, h0, n0);
      }
      mbar_wait(res_bar, 0);
    }
    const bool want_stats = p.gn_part[0] != nullptr;
    const uint32_t trow2 = tm_acc2 + lane_base;
    for (int c = cpar * 16; c < p.E; c += 64) {
      uint32_t v[16];
      tmem_ld16(trow2 + c, v);
      tmem_ld_wait();
      float f[16];
      {
        const float4* bp = reinterpret_cast<const float4*>(s_b2 + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = bp[i];
          f[4 * i] = __uint_as_float(v[4 * i]) + b4.x; f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
          f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z; f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
        }
      }
      uint8_t* brow = sblk + (size_t)(c >> 6) * kTile + r * 128;
      const int u0 = (c & 63) >> 3;
      uint4* a0 = reinterpret_cast<uint4*>(brow + (((u0) ^ (r & 7)) << 4));
      uint4* a1 = reinterpret_cast<uint4*>(brow + (((u0 + 1) ^ (r & 7)) << 4));
      if (p.has_res) {
        const uint4 r0 = *a0, r1 = *a1;
        const __half2* h0p = reinterpret_cast<const __half2*>(&r0);
        const __half2* h1p = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x0 = __half22float2(h0p[i]);
          const float2 x1 = __half22float2(h1p[i]);
          f[2 * i] += x0.x; f[2 * i + 1] += x0.y;
          f[8 + 2 * i] += x1.x; f[8 + 2 * i + 1] += x1.y;
        }
      }
      uint4 o0, o1;
      __half2* q0 = reinterpret_cast<__half2*>(&o0);
      __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q0[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        q1[i] = __floats2half2_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
      }
      *a0 = o0; *a1 = o1;
      if (want_stats) {
        float sv[16], sq[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x0 = __half22float2(q0[i]);
          const float2 x1 = __half22float2(q1[i]);
          sv[2 * i] = x0.x; sv[2 * i + 1] = x0.y; sv[8 + 2 * i] = x1.x; sv[8 + 2 * i + 1] = x1.y;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { if (!row_ok) sv[i] = 0.f; sq[i] = sv[i] * sv[i]; }
#pragma unroll
        for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
          const bool upper = (lane & bit) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send_s = upper ? sv[i] : sv[i + half];
            const float keep_s = upper ? sv[i + half] : sv[i];
            sv[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
            const float send_q = upper ? sq[i] : sq[i + half];
            const float keep_q = upper ? sq[i + half] : sq[i];
            sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
          }
        }
        sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
        sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 1);
        if ((lane & 1) == 0) {
          const int cidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          wsum[((size_t)quad * p.E + c + cidx) * 2] = sv[0];
          wsum[((size_t)quad * p.E + c + cidx) * 2 + 1] = sq[0];
        }
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 32 * kMlpEpiWarps);
    if (etid == 0) {
      for (int bq = 0; bq < nblk; ++bq) tma_store_4d(&p.tmOut, sblk + (size_t)bq * kTile, bq * 64, w0, h0, n0);
      tma_store_commit();
    }
    if (want_stats && n0 < p.Nimg) {
      const int slot = th * p.tiles_w + tw;
      for (int cc = etid; cc < p.E; cc += 32 * kMlpEpiWarps) {
        const float s0 = wsum[((size_t)0 * p.E + cc) * 2], q0s = wsum[((size_t)0 * p.E + cc) * 2 + 1];
        const float s1 = wsum[((size_t)1 * p.E + cc) * 2], q1s = wsum[((size_t)1 * p.E + cc) * 2 + 1];
        const float s2 = wsum[((size_t)2 * p.E + cc) * 2], q2s = wsum[((size_t)2 * p.E + cc) * 2 + 1];
        const float s3 = wsum[((size_t)3 * p.E + cc) * 2], q3s = wsum[((size_t)3 * p.E + cc) * 2 + 1];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          float* part = p.gn_part[d];
          if (!part) continue;
          const size_t ch = (size_t)p.gn_coff[d] + cc;
          if (p.bn == 1) {
            float* dst = part + (((size_t)n0 * p.gn_slots + slot) * p.gn_cstride[d] + ch) * 2;
            dst[0] = (s0 + s1) + (s2 + s3);
            dst[1] = (q0s + q1s) + (q2s + q3s);
          } else {
            float* dst = part + (((size_t)n0 * p.gn_slots + slot) * p.gn_cstride[d] + ch) * 2;
            dst[0] = s0 + s1; dst[1] = q0s + q1s;
            if (n0 + 1 < p.Nimg) {
              float* dst1 = part + (((size_t)(n0 + 1) * p.gn_slots + slot) * p.gn_cstride[d] + ch) * 2;
              dst1[0] = s2 + s3; dst1[1] = q2s + q3s;
            }
          }
        }
      }
    }
    if (etid == 0) tma_store_wait_read();
    if (dbg && etid == 0) dbg[63 * 8 + 1] = clock64() - t_start;
  }

  tc_fence_before();
  if (CS > 1) cluster_sync_all(); else __syncthreads();      // no CTA retires while a peer can still signal its barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc_dyn(tmem_base, 512u); }
}

#endif
}  // namespace rs
