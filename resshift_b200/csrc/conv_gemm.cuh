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
// * warp roles: warps 0..7 = epilogue (TMEM -> registers -> bias / activation / residual -> shared -> TMA store),
//   warp 8 = TMA producer, warp 9 = TMEM allocator + single-thread MMA issuer.
// * covers reference call sites: every nn.Conv2d / nn.Linear inside UNetModelSwin.forward
//   (reference models/unet.py:147,173,184,707,862,69,99-101; models/swin_transformer.py:22-24,105-107,480,515).
#pragma once

#include "common.cuh"
#include "gn_stats.cuh"

namespace rs {

constexpr int kConvBM = 128;      // pixels per tile (UMMA M)
constexpr int kConvBK = 64;       // channels per k-block (128 B rows, SWIZZLE_128B)
constexpr int kConvEpiWarps = 8;   // two warps per TMEM lane quadrant, interleaved over the 16-column chunks
constexpr int kConvThreads = 64 + 32 * kConvEpiWarps;
// Warp roles: the warp scheduler arbitrates highest-warp-id-first (B300_MICROARCH.md), so the two single-thread
// control warps sit ABOVE the epilogue warps — a busy epilogue (this CTA's or, through shared issue slots, simply more
// eligible warps) must never delay an MMA issue or a TMA refill.
constexpr int kConvTmaWarp = kConvEpiWarps;        // warp 8
constexpr int kConvMmaWarp = kConvEpiWarps + 1;    // warp 9
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
  int tail_k16;          // 16-channel MMA steps of the LAST 64-channel chunk that hold real channels (1..4): the zero-filled
                         // remainder of a partial chunk (Cin = 160 -> 32 of 64, Cin = 8 -> 8 of 64) is not multiplied
  int w_tap_stride;      // K offset between consecutive taps in the weight matrix (= Cin_pad)
  int bw, bh, bn;        // pixel box of a tile, bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_n;
  int Wout, Hout, Nimg;
  int BN, n_tiles, Cout;
  int stages;
  int tmem_cols;
  int cg;                // 1: one CTA per tile;  2: CTA pair (cluster of 2, tcgen05 cta_group::2): a 256-pixel x BN tile,
                         //    each CTA stages its own 128 pixels of A and HALF of the weight tile (halves the smem traffic of B)
  int splitk;            // > 1: the K loop is cut into `splitk` ranges, one CTA (or pair) each; the epilogue then only
                         //      stores fp32 partial accumulators to `partial` and splitk_reduce_kernel finishes the layer
  float* partial;        // [splitk][Nimg*Hout*Wout pixels][Cout] fp32
  int splitk_cluster;    // 1 (pair mode only): the S K-ranges of a tile are ONE cluster of 2*S CTAs; partial tiles stay in
                         //    shared memory and are summed through distributed shared memory (no scratch, no second kernel)
  int msub;              // 128-pixel sub-tiles per CTA (1 or 2): two sub-tiles share every weight tile (fewer operand bytes per MMA)
  // epilogue
  const float* bias;                 // [Cout] fp32 or nullptr
  const __half* residual;            // optional, same pixel grid as the output
  long long res_sN, res_sH, res_sW;  // strides in elements
  __half* out;                       // NHWC fp16 view (may be nullptr when out_f32 is set)
  long long out_sN, out_sH, out_sW;
  float* out_f32_nchw;               // optional fp32 NCHW output [Nimg, Cout, Hout, Wout]
  int act;
  unsigned long long* dbg;           // optional per-CTA timeline (8 x u64 per CTA, globaltimer ns), profiling aid
  // staged epilogue: results go to shared memory (128B-swizzled 64-column blocks) and leave through TMA stores;
  // the residual tile arrives the same way through a TMA load
  CUtensorMap tmOut, tmRes;
  int tma_out;                       // 1: staged epilogue (fp16 NHWC output); 0: direct per-thread stores
  int tma_res;
  int epi_bc;                        // staging block width in columns: 64 / 32 / 16 (swizzle 128B / 64B / 32B),
                                     // the largest that divides BN so a block never spills into the next channel tile
  // fused GroupNorm statistics of the OUTPUT for up to two consumers (gn_stats.cuh): per image / 128-pixel tile slot /
  // channel the pair (mean, M2) of the stored fp16 values; the last CTA to finish an image reduces its 32 groups
  GnSink sink[2];
  int gn_slots;
  // persistent variant (conv_persist.cuh): CTAs (pairs) walk work units u = worker, worker + #workers, ...
  int persist;
  int num_units;         // (pixel tiles or tile pairs) x channel tiles
};

#ifdef __CUDACC__

// kCG = 1: one CTA per tile (no cluster);  kCG = 2: CTA pair, launched with cluster dimension 2.  Two instantiations
// because a kernel that contains cta_group::2 instructions must be launched as a cluster.
template <int kCG>
__global__ void __launch_bounds__(kConvThreads, 2) conv_gemm_sm100_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B BN*128 B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = kConvBM * kConvBK * 2;                 // one sub-tile of A
  const int b_rows = kCG == 2 ? p.BN / 2 : p.BN;             // weight rows staged by THIS CTA
  const int b_bytes = b_rows * kConvBK * 2;
  const int stage_bytes = p.msub * a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint64_t* res_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256);   // [BN] bias of this channel tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 8 : nullptr;
  if (dbg && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    dbg[0] = global_timer_ns(); dbg[7] = smid;
  }

  // tile coordinates: CTA -> (channel tile, msub consecutive 128-pixel tiles)
  // pair mode: the cluster is one CTA pair — or, with splitk_cluster, the S pairs (K ranges) of one tile: cluster rank =
  // 2 * split + (rank inside the pair); the pair's leader is the even rank
  const uint32_t crank = kCG == 2 ? cluster_ctarank() : 0;
  const uint32_t rank = crank & 1;                            // leader = rank 0
  const uint32_t lead = crank & ~1u;                          // cluster rank of this pair's leader
  const uint16_t pair_mask = (uint16_t)(3u << lead);          // multicast mask of this pair
  int unit = kCG == 2 ? (blockIdx.x >> 1) : blockIdx.x;       // work unit: a CTA, or a CTA pair
  const int split = unit % p.splitk;                          // which K range (fastest index: splits of a tile run together)
  unit /= p.splitk;
  const int n_tile = unit % p.n_tiles;
  const int mt0 = kCG == 2 ? (unit / p.n_tiles) * 2 + (int)rank : (unit / p.n_tiles) * p.msub;
  const int total_kb = p.num_taps * p.kchunks;
  const int kb_begin = (int)((long long)total_kb * split / p.splitk);
  const int kb_end = (int)((long long)total_kb * (split + 1) / p.splitk);
  const int num_kb = kb_end - kb_begin;
  auto tile_origin = [&](int sub, int& tw_, int& th_, int& w0_, int& h0_, int& n0_) {
    int mt = mt0 + sub;
    tw_ = mt % p.tiles_w; mt /= p.tiles_w;
    th_ = mt % p.tiles_h; mt /= p.tiles_h;
    w0_ = tw_ * p.bw; h0_ = th_ * p.bh; n0_ = mt * p.bn;
  };

  if (warp == kConvTmaWarp && lane == 0) {
    for (int s = 0; s < kMaxSrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
    if (p.tma_out) tma_prefetch_desc(&p.tmOut);
    if (p.tma_res) tma_prefetch_desc(&p.tmRes);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(res_bar, 1);
    mbar_fence_init();
  }
  if (warp == kConvMmaWarp) {
    if constexpr (kCG == 2) { tmem_alloc_dyn_cg2(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish_cg2(); }
    else { tmem_alloc_dyn(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();     // peer barriers must be initialised before remote arrivals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_trigger();
  pdl_wait();
  if (dbg && threadIdx.x == 0) dbg[1] = global_timer_ns();

  // The two control warps run their loops WARP-UNIFORMLY (every lane computes the same addresses / descriptors, which
  // the compiler keeps in uniform registers) and only the asynchronous instructions themselves are predicated on one
  // elected lane.  Written as `if (lane == 0) { loop }` instead, every tcgen05.mma / commit / TMA issue is wrapped in
  // an elect-and-broadcast loop and the scalar instruction stream (~450 cycles per k-block, scripts/ubench/
  // umma_issue*.cu) — not the tensor pipe — sets the pace of the main loop.
  if (warp == kConvTmaWarp) {
    // ===================== TMA producer =====================
    const bool el = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int tws[2], ths[2], w0s[2], h0s[2], n0s[2];
    for (int sub = 0; sub < p.msub; ++sub) tile_origin(sub, tws[sub], ths[sub], w0s[sub], h0s[sub], n0s[sub]);
    int tap = kb_begin / p.kchunks;
    int kc = kb_begin - tap * p.kchunks;
    const uint32_t tx_bytes = (uint32_t)(kCG == 2 ? 2 * stage_bytes : stage_bytes);
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + (size_t)stage * stage_bytes;
      uint8_t* sb = sa + p.msub * a_bytes;
      const CUtensorMap* ma = &p.tmA[p.tap_src[tap]];
      const int dw = p.tap_dw[tap], dh = p.tap_dh[tap];
      const int kcol = kc * kConvBK, wcol = tap * p.w_tap_stride + kc * kConvBK;
      if constexpr (kCG == 2) {
        // both CTAs' loads complete on the LEADER's full barrier; only the leader arms it (with the bytes of both)
        const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), lead);
        if (el) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_4d_cg2(sa, ma, lead_bar, kcol, w0s[0] + dw, h0s[0] + dh, n0s[0]);
          tma_load_2d_cg2(sb, &p.tmB, lead_bar, wcol, n_tile * p.BN + (int)rank * b_rows);
        }
      } else {
        if (el) {
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          for (int sub = 0; sub < p.msub; ++sub)
            tma_load_4d(sa + sub * a_bytes, ma, &full_bar[stage], kcol, w0s[sub] + dw, h0s[sub] + dh, n0s[sub]);
          tma_load_2d(sb, &p.tmB, &full_bar[stage], wcol, n_tile * p.BN);
        }
      }
      if (++kc == p.kchunks) { kc = 0; ++tap; }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == kConvMmaWarp && rank == 0) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    const bool el = elect_one();
    const uint32_t idesc = umma_idesc_f16(kCG == 2 ? 2 * kConvBM : kConvBM, p.BN);
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t b_off = (uint32_t)(p.msub * a_bytes);
    int stage = 0;
    uint32_t phase = 0;
    int kc = kb_begin % p.kchunks;                    // 64-channel chunk index inside the current tap
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (dbg && el && kb == 0) dbg[2] = global_timer_ns();
      const uint32_t sa = smem0 + (uint32_t)stage * (uint32_t)stage_bytes;
      const uint64_t adesc = umma_desc_sw128(sa);
      const uint64_t bdesc = umma_desc_sw128(sa + b_off);
      const uint32_t acc0 = kb != 0 ? 1u : 0u;
      const int ksteps = (kc == p.kchunks - 1) ? p.tail_k16 : kConvBK / 16;   // skip the zero-filled tail of a partial chunk
      if (++kc == p.kchunks) kc = 0;
      if constexpr (kCG == 2) {
        if (el) {
          umma_f16_cg2(tmem_base, adesc, bdesc, idesc, acc0);
#pragma unroll
          for (int k = 1; k < kConvBK / 16; ++k) if (k < ksteps) umma_f16_cg2(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
          umma_commit_cg2(&empty_bar[stage], pair_mask);                    // frees the stage in BOTH CTAs
          if (kb == num_kb - 1) umma_commit_cg2(tmem_full_bar, pair_mask);  // accumulators of both CTAs complete
        }
      } else {
        if (el) {
          // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc, bdesc, idesc, acc0);
#pragma unroll
          for (int k = 1; k < kConvBK / 16; ++k) if (k < ksteps) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
          if (p.msub == 2) {
            const uint64_t adesc1 = umma_desc_sw128(sa + a_bytes);
            umma_f16(tmem_base + p.BN, adesc1, bdesc, idesc, acc0);
#pragma unroll
            for (int k = 1; k < kConvBK / 16; ++k) if (k < ksteps) umma_f16(tmem_base + p.BN, adesc1 + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);                 // frees this smem stage when the MMAs retire
          if (kb == num_kb - 1) umma_commit(tmem_full_bar);  // accumulator complete
        }
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (dbg && el) dbg[3] = global_timer_ns();
  } else if (warp < kConvEpiWarps) {
    // ===================== epilogue (8 warps: lane quadrant = warp % 4, column parity = warp / 4) ==========
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int cpar = warp >> 2;                      // which half of the 16-column chunks this warp handles
    const int r = quad * 32 + lane;                  // row of a sub-tile == TMEM lane
    const int lw = r % p.bw;
    const int lh = (r / p.bw) % p.bh;
    const int ln = r / (p.bw * p.bh);
    const int col0 = n_tile * p.BN;
    const int etid = threadIdx.x;                    // 0..255 among epilogue threads

    // the channel tile's bias goes to shared memory while the main loop runs (broadcast reads in the epilogue)
    for (int i = etid; i < p.BN; i += 32 * kConvEpiWarps)
      s_bias[i] = (p.bias && col0 + i < p.Cout) ? __ldg(p.bias + col0 + i) : 0.f;
    named_bar_sync(1, 32 * kConvEpiWarps);

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (dbg && etid == 0) dbg[4] = global_timer_ns();

    if (p.splitk > 1 && p.splitk_cluster) {
      // ---------- cluster split-K, step 1: this K range's fp32 partial tile -> own shared memory ----------
      // (the operand ring is free: every MMA has retired; row pitch BN*4 + 16 B spreads the rows over the banks)
      const int pitch = p.BN * 4 + 16;
      const uint32_t drow = smem_u32(smem) + (uint32_t)(r * pitch);
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c = cpar * 16; c < p.BN; c += 32) {
        uint32_t v[16];
        tmem_ld16(trow + c, v);
        tmem_ld_wait16(v);
#pragma unroll
        for (int j = 0; j < 4; ++j) st_shared_v4(drow + (uint32_t)(c * 4 + j * 16), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    } else if (p.splitk > 1) {
      // ---------- split-K: raw fp32 partial sums, finished by splitk_reduce_kernel ----------
      int tw, th, w0, h0, n0;
      tile_origin(0, tw, th, w0, h0, n0);
      const int w = w0 + lw, h = h0 + lh, n = n0 + ln;
      const bool row_ok = (w < p.Wout) && (h < p.Hout) && (n < p.Nimg);
      const long long pix = ((long long)n * p.Hout + h) * p.Wout + w;
      const long long npix = (long long)p.Nimg * p.Hout * p.Wout;
      float* prow = p.partial + ((long long)split * npix + pix) * p.Cout;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c = cpar * 16; c < p.BN; c += 32) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(trow + c, v);
        tmem_ld_wait();
        const int col = col0 + c;
        if (!row_ok || col >= p.Cout) continue;
        if (col + 16 <= p.Cout) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(prow + col + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          for (int j = 0; j < 16 && col + j < p.Cout; ++j) prow[col + j] = __uint_as_float(v[j]);
        }
      }
    } else if (p.tma_out) {
      // ---------- staged epilogue ----------
      // All operand stages are free now (every MMA that read them has retired), so the pipeline smem is reused.
      // Per sub-tile: BN/bc blocks of [128 rows x bc columns] fp16, swizzled like the TMA box — first the residual
      // tile lands there (TMA load), then the finished outputs are written in place, then one TMA store per block.
      // After the blocks of all sub-tiles: per-warp GroupNorm partials [msub][4 quads][BN][2].
      const int bc = p.epi_bc;                       // columns per staging block
      const int nblk = p.BN / bc;
      const int blk_bytes = kConvBM * bc * 2;
      const int sub_bytes = nblk * blk_bytes;
      const int bshift = (bc == 64) ? 6 : (bc == 32 ? 5 : 4);
      // Swizzle<B,4,3>: the 16-byte unit index is XORed with address bits [7, 7+B); row pitch is 2*bc bytes
      const int swz = (bc == 64) ? (r & 7) : (bc == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
      float* wsum_all = reinterpret_cast<float*>(smem + (size_t)p.msub * sub_bytes);
      if (p.tma_res) {
        if (etid == 0) {
          mbar_arrive_expect_tx(res_bar, (uint32_t)(p.msub * sub_bytes));
          for (int sub = 0; sub < p.msub; ++sub) {
            int tw, th, w0, h0, n0;
            tile_origin(sub, tw, th, w0, h0, n0);
            for (int b = 0; b < nblk; ++b)
              tma_load_4d(smem + (size_t)sub * sub_bytes + (size_t)b * blk_bytes, &p.tmRes, res_bar, col0 + b * bc, w0, h0, n0);
          }
        }
        mbar_wait(res_bar, 0);
      }
      const bool want_stats = p.sink[0].part != nullptr;
      for (int sub = 0; sub < p.msub; ++sub) {
        int tw, th, w0, h0, n0;
        tile_origin(sub, tw, th, w0, h0, n0);
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + sub * p.BN;
        uint8_t* sblk = smem + (size_t)sub * sub_bytes;
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
          // this thread's two 16-byte units inside the swizzled block
          uint8_t* brow = sblk + (size_t)(c >> bshift) * blk_bytes + r * (2 * bc);
          const int u0 = (c & (bc - 1)) >> 3;                  // 16-byte unit index of columns c..c+7 inside the block
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
          if (want_stats) warp_chunk_stats(o0, o1, lane, wsum_all + ((size_t)(sub * 4 + quad) * p.BN + c) * 2);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 32 * kConvEpiWarps);
      if (dbg && etid == 0) dbg[6] = global_timer_ns();
      if (etid == 0) {
        for (int sub = 0; sub < p.msub; ++sub) {
          int tw, th, w0, h0, n0;
          tile_origin(sub, tw, th, w0, h0, n0);
          for (int b = 0; b < nblk; ++b)
            if (col0 + b * bc < p.Cout)
              tma_store_4d(&p.tmOut, smem + (size_t)sub * sub_bytes + (size_t)b * blk_bytes, col0 + b * bc, w0, h0, n0);
        }
        tma_store_commit();
      }
      if (want_stats) {
        int* s_flag = reinterpret_cast<int*>(wsum_all + (size_t)p.msub * 8 * p.BN);
        for (int sub = 0; sub < p.msub; ++sub) {
          int tw, th, w0, h0, n0;
          tile_origin(sub, tw, th, w0, h0, n0);
          const int ncols = min(p.BN, p.Cout - col0);
          const int slot = th * p.tiles_w + tw;
          if (n0 < p.Nimg)                                           // (else: padding tile of an odd pair)
            write_quad_pairs(wsum_all + (size_t)sub * 8 * p.BN, p.BN, ncols, col0, p.bn, n0, p.Nimg, slot, p.gn_slots, p.sink[0], p.sink[1],
                             etid, 32 * kConvEpiWarps);
          if (p.sink[0].gstat || p.sink[1].gstat) {                  // producer-side finalisation (large tensors only)
            const GnSink* const sk[4] = {&p.sink[0], &p.sink[0], p.sink[1].part ? &p.sink[1] : nullptr, p.sink[1].part ? &p.sink[1] : nullptr};
            const int n1 = (p.bn == 2 && n0 + 1 < p.Nimg) ? n0 + 1 : -1;
            const int im[4] = {n0 < p.Nimg ? n0 : -1, n0 < p.Nimg ? n1 : -1, n0 < p.Nimg ? n0 : -1, n0 < p.Nimg ? n1 : -1};
            const unsigned int ad[4] = {(unsigned)ncols, (unsigned)ncols, (unsigned)ncols, (unsigned)ncols};
            gn_arrive<4>(sk, im, ad, p.gn_slots, 128.0f / (float)p.bn, etid, 32 * kConvEpiWarps, 1, s_flag);
          }
        }
      }
      if (etid == 0) tma_store_wait_read();
    } else {
      // ---------- direct epilogue (fp32 NCHW model head, or RS_CONV_EPI=direct) ----------
      for (int sub = 0; sub < p.msub; ++sub) {
        int tw, th, w0, h0, n0;
        tile_origin(sub, tw, th, w0, h0, n0);
        const int w = w0 + lw, h = h0 + lh, n = n0 + ln;
        const bool row_ok = (w < p.Wout) && (h < p.Hout) && (n < p.Nimg);
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + sub * p.BN;
        __half* orow = p.out ? p.out + n * p.out_sN + h * p.out_sH + w * p.out_sW : nullptr;
        const __half* rrow = p.residual ? p.residual + n * p.res_sN + h * p.res_sH + w * p.res_sW : nullptr;
        for (int c = cpar * 16; c < p.BN; c += 32) {
          uint32_t v[16];
          __syncwarp();   // tcgen05.ld is warp-collective: reconverge after the divergent tail of the last chunk
          tmem_ld16(trow + c, v);
          tmem_ld_wait();
          const int col = col0 + c;
          if (!row_ok || col >= p.Cout) continue;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col + j < p.Cout) f[j] += __ldg(p.bias + col + j);
          }
          if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = gelu_erf_f(f[j]);
          } else if (p.act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = silu_f(f[j]);
          }
          if (rrow)
            for (int j = 0; j < 16 && col + j < p.Cout; ++j) f[j] += __half2float(rrow[col + j]);
          if (p.out_f32_nchw)
            for (int j = 0; j < 16 && col + j < p.Cout; ++j)
              p.out_f32_nchw[(((long long)n * p.Cout + (col + j)) * p.Hout + h) * p.Wout + w] = f[j];
          if (orow)
            for (int j = 0; j < 16 && col + j < p.Cout; ++j) orow[col + j] = __float2half_rn(f[j]);
        }
      }
    }
  }

  if constexpr (kCG == 2) {
    if (p.splitk > 1 && p.splitk_cluster) {
      // ---------- cluster split-K, step 2: sum the S partial tiles through distributed shared memory ----------
      // K range `split` finishes columns [split * BN/S, (split + 1) * BN/S) of the tile for its CTA's 128 pixels:
      // fixed summation order (range 0, 1, ...), then the usual epilogue work (bias / activation / residual / fp16
      // store / GroupNorm partials of the stored values).  Deterministic, no global scratch, no second kernel.
      cluster_sync_all();                                   // every K range's partial tile is in its CTA's shared memory
      if (warp < kConvEpiWarps) {
        const int etid = threadIdx.x;
        const int S = p.splitk, cw = p.BN / S, upr = cw >> 3;
        const int pitch = p.BN * 4 + 16;
        const int cbase = split * cw;
        const int col0 = n_tile * p.BN;
        __half* s_out = reinterpret_cast<__half*>(smem + (size_t)kConvBM * pitch);        // [128][cw] stored values
        float* s_col = reinterpret_cast<float*>(s_out + (size_t)kConvBM * cw);            // [2 halves][cw][2]
        int tw, th, w0, h0, n0;
        tile_origin(0, tw, th, w0, h0, n0);
        const uint32_t dump0 = smem_u32(smem);
        for (int u = etid; u < kConvBM * upr; u += 32 * kConvEpiWarps) {
          const int rr = u / upr, cu = u - rr * upr;
          const int ct = cbase + cu * 8;                    // column inside the BN tile
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int sp = 0; sp < S; ++sp) {
            const uint32_t a = mapa_u32(dump0 + (uint32_t)(rr * pitch + ct * 4), (uint32_t)(sp * 2) + rank);
            const float4 x0 = ld_shared_cluster_f4(a), x1 = ld_shared_cluster_f4(a + 16);
            acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w;
            acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
          }
          const int lw = rr % p.bw, lh = (rr / p.bw) % p.bh, ln = rr / (p.bw * p.bh);
          const int w = w0 + lw, h = h0 + lh, n = n0 + ln;
          const int col = col0 + ct;
          const bool ok = (w < p.Wout) && (h < p.Hout) && (n < p.Nimg) && (col < p.Cout);   // Cout % 8 == 0
          uint4 o = make_uint4(0, 0, 0, 0);
          if (ok) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] += __ldg(p.bias + col + j);
            }
            if (p.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = gelu_erf_f(acc[j]);
            } else if (p.act == ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = silu_f(acc[j]);
            }
            if (p.residual) {
              const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + n * p.res_sN + h * p.res_sH + w * p.res_sW + col);
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(rh[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
            }
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
            *reinterpret_cast<uint4*>(p.out + n * p.out_sN + h * p.out_sH + w * p.out_sW + col) = o;
          }
          *reinterpret_cast<uint4*>(s_out + (size_t)rr * cw + cu * 8) = o;                  // zeros outside the tensor
        }
        if (p.sink[0].part != nullptr) {
          named_bar_sync(1, 32 * kConvEpiWarps);
          // (mean, M2) of the stored values per column over the two 64-row halves (pivot = first row, rows in order)
          for (int t = etid; t < 2 * cw; t += 32 * kConvEpiWarps) {
            const int half = t / cw, c = t - half * cw;
            const __half* col = s_out + (size_t)(half * 64) * cw + c;
            const float pv = __half2float(col[0]);
            float s1 = 0.f, s2 = 0.f;
            for (int rr = 1; rr < 64; ++rr) {
              const float d = __half2float(col[(size_t)rr * cw]) - pv;
              s1 += d; s2 = fmaf(d, d, s2);
            }
            s_col[(half * cw + c) * 2] = pv + s1 * (1.0f / 64.0f);
            s_col[(half * cw + c) * 2 + 1] = fmaxf(s2 - s1 * s1 * (1.0f / 64.0f), 0.f);
          }
          named_bar_sync(1, 32 * kConvEpiWarps);
          const int slot = th * p.tiles_w + tw;
          const int ncols = max(0, min(cw, p.Cout - (col0 + cbase)));
          if (n0 < p.Nimg)
            write_tile_pairs(s_col, cw, ncols, col0 + cbase, p.bn, n0, p.Nimg, slot, p.gn_slots, p.sink[0], p.sink[1], etid, 32 * kConvEpiWarps);
          if (p.sink[0].gstat || p.sink[1].gstat) {
            int* s_flag = reinterpret_cast<int*>(s_col + 4 * cw);
            const GnSink* const sk[4] = {&p.sink[0], &p.sink[0], p.sink[1].part ? &p.sink[1] : nullptr, p.sink[1].part ? &p.sink[1] : nullptr};
            const int n1 = (p.bn == 2 && n0 + 1 < p.Nimg) ? n0 + 1 : -1;
            const int im[4] = {n0 < p.Nimg ? n0 : -1, n0 < p.Nimg ? n1 : -1, n0 < p.Nimg ? n0 : -1, n0 < p.Nimg ? n1 : -1};
            const unsigned int ad[4] = {(unsigned)ncols, (unsigned)ncols, (unsigned)ncols, (unsigned)ncols};
            gn_arrive<4>(sk, im, ad, p.gn_slots, 128.0f / (float)p.bn, etid, 32 * kConvEpiWarps, 1, s_flag);
          }
        }
      }
    }
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();    // pair: neither CTA may retire while the other can still
                                                               // touch its smem / barriers / TMEM
  if (dbg && threadIdx.x == 0) dbg[5] = global_timer_ns();
  if (warp == kConvMmaWarp) {
    tc_fence_after();
    if constexpr (kCG == 2) tmem_dealloc_dyn_cg2(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
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
  pdl_trigger();
  pdl_wait();
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
