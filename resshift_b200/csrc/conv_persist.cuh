// Persistent variant of the implicit-GEMM conv kernel (conv_gemm.cuh) for layers with several output tiles per SM.
//
// One CTA (or CTA pair) per SM walks a strided list of output tiles.  The accumulator is DOUBLE BUFFERED in tensor
// memory (2 x BN columns) and the epilogue has its own staging buffer, so while the eight epilogue warps drain tile i
// (tcgen05.ld -> bias / activation / residual -> fp16 -> shared -> TMA store, GroupNorm partials) the producer and the
// MMA issuer are already streaming tile i+1: barrier set-up, TMEM allocation, the first operand round trip, the
// epilogue and the store drain are paid once per CTA instead of once per tile (in the one-tile-per-CTA kernel they are
// ~40 % of a 64x64 3x3 layer: profiles/r1_s25_conv_timeline_uniform_issue.log).
// Same ConvParams, same tile / operand / epilogue conventions and bit-identical results as conv_gemm_sm100_kernel
// (staged TMA-store epilogue only: fp16 NHWC output, no split-K, one sub-tile per CTA).
#pragma once

#include "conv_gemm.cuh"

namespace rs {

#ifdef __CUDACC__

template <int kCG>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_persist_sm100_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = kConvBM * kConvBK * 2;
  const int b_rows = kCG == 2 ? p.BN / 2 : p.BN;             // weight rows staged by THIS CTA
  const int b_bytes = b_rows * kConvBK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  const size_t stage_sz = (size_t)p.BN * kConvBM * 2;
  uint8_t* s_stage0 = smem + (size_t)p.stages * stage_bytes;                // 2 x [BN/bc blocks][128 rows][bc] fp16 (swizzled):
                                                                            // tile i stages while tile i-1's TMA store drains
  float* wsum = reinterpret_cast<float*>(s_stage0 + 2 * stage_sz);          // [4 quads][BN][2]
  float* s_bias0 = wsum + (size_t)4 * p.BN * 2;                             // 2 x [BN]: this tile's / the next tile's bias
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias0 + 2 * p.BN);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* acc_full = empty_bar + p.stages;       // [2] per CTA (multicast commit in pair mode)
  uint64_t* acc_empty = acc_full + 2;              // [2] leader's: one arrival per epilogue warp (of both CTAs)
  uint64_t* res_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  GnArriveList* alist = reinterpret_cast<GnArriveList*>(tmem_slot + 2);   // deferred GroupNorm arrivals (gn_stats.cuh)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 8 : nullptr;
  if (dbg && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    dbg[0] = global_timer_ns(); dbg[7] = smid;
  }
  const uint32_t rank = kCG == 2 ? cluster_ctarank() : 0;    // leader = rank 0
  const int worker = kCG == 2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_workers = kCG == 2 ? (gridDim.x >> 1) : gridDim.x;
  const int num_kb = p.num_taps * p.kchunks;
  // work unit -> (channel tile, this CTA's 128-pixel tile origin)
  auto unit_tile = [&](int u, int& n_tile, int& tw_, int& th_, int& w0_, int& h0_, int& n0_) {
    n_tile = u % p.n_tiles;
    int mt = kCG == 2 ? (u / p.n_tiles) * 2 + (int)rank : (u / p.n_tiles);
    tw_ = mt % p.tiles_w; mt /= p.tiles_w;
    th_ = mt % p.tiles_h; mt /= p.tiles_h;
    w0_ = tw_ * p.bw; h0_ = th_ * p.bh; n0_ = mt * p.bn;
  };

  if (warp == kConvTmaWarp && lane == 0) {
    for (int s = 0; s < kMaxSrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    if (p.tma_res) tma_prefetch_desc(&p.tmRes);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kConvEpiWarps * kCG); }
    mbar_init(res_bar, 1);
    mbar_fence_init();
  }
  if (warp == kConvMmaWarp) {
    if constexpr (kCG == 2) { tmem_alloc_dyn_cg2(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish_cg2(); }
    else { tmem_alloc_dyn(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  if (dbg && threadIdx.x == 0) dbg[1] = global_timer_ns();

  if (warp == kConvTmaWarp) {
    // ===================== TMA producer (runs ahead across tile boundaries) =====================
    const bool el = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = (uint32_t)(kCG == 2 ? 2 * stage_bytes : stage_bytes);
    for (int u = worker; u < p.num_units; u += num_workers) {
      int n_tile, tw, th, w0, h0, n0;
      unit_tile(u, n_tile, tw, th, w0, h0, n0);
      int tap = 0, kc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        const CUtensorMap* ma = &p.tmA[p.tap_src[tap]];
        const int dw = p.tap_dw[tap], dh = p.tap_dh[tap];
        const int kcol = kc * kConvBK, wcol = tap * p.w_tap_stride + kc * kConvBK;
        if constexpr (kCG == 2) {
          const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (el) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d_cg2(sa, ma, lead_bar, kcol, w0 + dw, h0 + dh, n0);
            tma_load_2d_cg2(sb, &p.tmB, lead_bar, wcol, n_tile * p.BN + (int)rank * b_rows);
          }
        } else {
          if (el) {
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(sa, ma, &full_bar[stage], kcol, w0 + dw, h0 + dh, n0);
            tma_load_2d(sb, &p.tmB, &full_bar[stage], wcol, n_tile * p.BN);
          }
        }
        if (++kc == p.kchunks) { kc = 0; ++tap; }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kConvMmaWarp && rank == 0) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    const bool el = elect_one();
    const uint32_t idesc = umma_idesc_f16(kCG == 2 ? 2 * kConvBM : kConvBM, p.BN);
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int i = 0;
    for (int u = worker; u < p.num_units; u += num_workers, ++i) {
      const int b = i & 1;
      mbar_wait(&acc_empty[b], ((i >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(b * p.BN);
      int kc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)stage * (uint32_t)stage_bytes;
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + (uint32_t)a_bytes);
        const uint32_t acc0 = kb != 0 ? 1u : 0u;
        const int ksteps = (kc == p.kchunks - 1) ? p.tail_k16 : kConvBK / 16;   // skip the zero-filled tail of a partial chunk
        if (++kc == p.kchunks) kc = 0;
        if constexpr (kCG == 2) {
          if (el) {
            umma_f16_cg2(d, adesc, bdesc, idesc, acc0);
#pragma unroll
            for (int k = 1; k < kConvBK / 16; ++k) if (k < ksteps) umma_f16_cg2(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
            umma_commit_cg2(&empty_bar[stage], 3);
            if (kb == num_kb - 1) umma_commit_cg2(&acc_full[b], 3);
          }
        } else {
          if (el) {
            umma_f16(d, adesc, bdesc, idesc, acc0);
#pragma unroll
            for (int k = 1; k < kConvBK / 16; ++k) if (k < ksteps) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
            umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&acc_full[b]);
          }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < kConvEpiWarps) {
    // ===================== epilogue (8 warps: lane quadrant = warp % 4, column parity = warp / 4) ==========
    const int quad = warp & 3;
    const int cpar = warp >> 2;
    const int r = quad * 32 + lane;                  // row of the tile == TMEM lane
    const int lw = r % p.bw;
    const int lh = (r / p.bw) % p.bh;
    const int ln = r / (p.bw * p.bh);
    const int etid = threadIdx.x;
    const int bc = p.epi_bc;                         // columns per staging block
    const int nblk = p.BN / bc;
    const int blk_bytes = kConvBM * bc * 2;
    const int bshift = (bc == 64) ? 6 : (bc == 32 ? 5 : 4);
    const int swz = (bc == 64) ? (r & 7) : (bc == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
    const bool want_stats = p.sink[0].part != nullptr;
    const uint32_t lead_acc_empty = kCG == 2 ? mapa_u32(smem_u32(acc_empty), 0) : smem_u32(acc_empty);
    int i = 0;
    auto load_bias = [&](int u, float* dst) {
      const int c0 = (u % p.n_tiles) * p.BN;
      for (int c = etid; c < p.BN; c += 32 * kConvEpiWarps) dst[c] = (p.bias && c0 + c < p.Cout) ? __ldg(p.bias + c0 + c) : 0.f;
    };
    if (etid == 0) { alist->cnt = 0; alist->over = 0; }
    if (worker < p.num_units) load_bias(worker, s_bias0);          // visible after the first tile's opening barrier
    for (int u = worker; u < p.num_units; u += num_workers, ++i) {
      const int b = i & 1;
      int n_tile, tw, th, w0, h0, n0;
      unit_tile(u, n_tile, tw, th, w0, h0, n0);
      const int col0 = n_tile * p.BN;
      uint8_t* s_stage = s_stage0 + (size_t)b * stage_sz;
      const float* s_bias = s_bias0 + (size_t)b * p.BN;
      // this staging buffer was last used by tile i-2: its TMA store has read it (tile i-1's may still be draining);
      // the opening barrier also publishes this tile's bias (loaded during tile i-1) and frees wsum / the other bias slot
      if (etid == 0 && i >= 2) tma_store_wait_read_keep1();
      named_bar_sync(1, 32 * kConvEpiWarps);
      if (u + num_workers < p.num_units) load_bias(u + num_workers, s_bias0 + (size_t)(b ^ 1) * p.BN);
      if (p.tma_res && etid == 0) {
        mbar_arrive_expect_tx(res_bar, (uint32_t)(nblk * blk_bytes));
        for (int bq = 0; bq < nblk; ++bq)
          tma_load_4d(s_stage + (size_t)bq * blk_bytes, &p.tmRes, res_bar, col0 + bq * bc, w0, h0, n0);
      }
      mbar_wait(&acc_full[b], (i >> 1) & 1);
      tc_fence_after();
      if (p.tma_res) mbar_wait(res_bar, i & 1);

      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (uint32_t)(b * p.BN);
      uint32_t vn[16];
      if (cpar * 16 < p.BN) tmem_ld16(trow + cpar * 16, vn);
      for (int c = cpar * 16; c < p.BN; c += 32) {
        tmem_ld_wait16(vn);
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(vn[j]);
        if (c + 32 < p.BN) tmem_ld16(trow + c + 32, vn);     // next chunk's TMEM read overlaps this chunk's arithmetic
        {
          const float4* bp = reinterpret_cast<const float4*>(s_bias + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = bp[j];
            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
          }
        }
        if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = gelu_erf_f(f[j]);
        } else if (p.act == ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = silu_f(f[j]);
        }
        uint8_t* brow = s_stage + (size_t)(c >> bshift) * blk_bytes + r * (2 * bc);
        const int u0 = (c & (bc - 1)) >> 3;
        uint4* a0 = reinterpret_cast<uint4*>(brow + (((u0) ^ swz) << 4));
        uint4* a1 = reinterpret_cast<uint4*>(brow + (((u0 + 1) ^ swz) << 4));
        if (p.tma_res) {
          const uint4 r0 = *a0, r1 = *a1;
          const __half2* h0p = reinterpret_cast<const __half2*>(&r0);
          const __half2* h1p = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x0 = __half22float2(h0p[j]);
            const float2 x1 = __half22float2(h1p[j]);
            f[2 * j] += x0.x; f[2 * j + 1] += x0.y;
            f[8 + 2 * j] += x1.x; f[8 + 2 * j + 1] += x1.y;
          }
        }
        uint4 o0, o1;
        __half2* q0 = reinterpret_cast<__half2*>(&o0);
        __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          q0[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
          q1[j] = __floats2half2_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
        }
        *a0 = o0; *a1 = o1;
        // statistics of the values as stored, while the tile drains (gn_stats.cuh)
        if (want_stats) warp_chunk_stats(o0, o1, lane, wsum + ((size_t)quad * p.BN + c) * 2);
      }
      // this warp has read its share of the accumulator: hand the buffer back to the MMA issuer (tile i + 2)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCG == 2) mbar_arrive_remote(lead_acc_empty + b * 8); else mbar_arrive(&acc_empty[b]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 32 * kConvEpiWarps);
      if (etid == 0) {
        for (int bq = 0; bq < nblk; ++bq)
          if (col0 + bq * bc < p.Cout) tma_store_4d(&p.tmOut, s_stage + (size_t)bq * blk_bytes, col0 + bq * bc, w0, h0, n0);
        tma_store_commit();
      }
      if (want_stats && n0 < p.Nimg) {                               // (else: padding tile of an odd pair)
        const int ncols = min(p.BN, p.Cout - col0);
        const int slot = th * p.tiles_w + tw;
        write_quad_pairs(wsum, p.BN, ncols, col0, p.bn, n0, p.Nimg, slot, p.gn_slots, p.sink[0], p.sink[1], etid, 32 * kConvEpiWarps);
        if (etid == 0) {                                             // producer-side finalisation: arrivals are batched
#pragma unroll
          for (int d = 0; d < 2; ++d)
            if (p.sink[d].part && p.sink[d].gstat) {
              gn_list_add(alist, d, n0, (unsigned)ncols);
              if (p.bn == 2 && n0 + 1 < p.Nimg) gn_list_add(alist, d, n0 + 1, (unsigned)ncols);
            }
        }
      }
    }
    if (want_stats && (p.sink[0].gstat || p.sink[1].gstat)) gn_list_arrive(alist, p.sink[0], p.sink[1], p.gn_slots, 128.0f / (float)p.bn, etid, 32 * kConvEpiWarps, 1);
    if (etid == 0) tma_store_wait_read();
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[5] = global_timer_ns();
  if (warp == kConvMmaWarp) {
    tc_fence_after();
    if constexpr (kCG == 2) tmem_dealloc_dyn_cg2(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
  }
}

#endif
}  // namespace rs
