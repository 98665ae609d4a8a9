// The attention half of a Swin block as ONE tcgen05 kernel (the tensor-core version of swin_attn_fused.cuh):
//
//     y = x + proj( window_attention( qkv( norm1(x) ) ) )            (+ GroupNorm statistics of y for norm2)
//
// reference: SwinTransformerBlock.forward, models/swin_transformer.py:246-275 (norm1 = GroupNorm32, torch.roll,
// window_partition, WindowAttention.forward :114-145 incl. qkv / relative-position bias / shift mask / softmax / proj,
// window_reverse, roll back, residual).  Same arithmetic and the same intermediate roundings as the unfused path and as
// swin_attn_fused_kernel (fp16 n1, fp16 q / k / v with bias, softmax in fp32, un-normalised fp16 P, fp16 O, one final
// rounding of y) — but every GEMM runs on tcgen05 with the accumulators in TMEM:
//
//   * a persistent CTA owns a run of consecutive window PAIRS; a tile = TWO 8x8 windows = 128 token rows = one UMMA M.  The
//     rows are gathered with cp.async straight into their K-major 128B-swizzled operand positions (cyclic shift and window
//     partition are address arithmetic; the next tile's rows already fly under this tile's epilogue), normalised in place
//     with the per-image affine of norm1 (kept across tiles, recomputed when the image changes) and fenced to the async
//     proxy: exactly what a TMA load of a normalised tensor would have produced;
//   * the heads are walked in GROUPS of two (64 channels): [Q_g | K_g | V_g] = Xn . W_g^T is one N = 192 accumulator
//     (weight tiles stream through a three-slot TMA ring: a whole group in flight), drained to shared memory as fp16
//     operands — Q_g / K_g row-major (K-major A / B operands of QK^T), V_g TRANSPOSED (the K-major B operand of PV wants
//     [head_dim][token]);
//   * per head: S = Q_h K_h^T as two N = 64 MMAs (the keys of window 0, the keys of window 1): all 128 rows are multiplied
//     against each window's keys and each row simply reads the 64 columns of ITS window — block-diagonal attention
//     without M = 64 instructions.  Softmax: one thread per query row (no shuffles), bias from a 225-entry table per head
//     in shared memory, P written as the next A operand.  O = P V the same way (two N = 32 MMAs into the TMEM columns S
//     just vacated), scaled by 1 / rowsum and written into the A operand of the projection (O_0 in its own tile, the
//     later groups' O into the by then dead Xn tile);
//   * pipelining inside a tile: the next group's QKV GEMM runs under this group's softmax, its Q / K drain under this
//     group's PV round trip, the projection's first k-blocks under the last group's softmax;
//   * y = O W_proj^T + b + x: N = E accumulator, + raw x (cp.async into the dead Q / K / V^T region while the projection
//     runs), one rounding, staged in place, written as full token rows; (mean, M2) per (window, channel) for the norm2
//     that follows — for all but a CTA's last tile computed while the next tile waits for its first GEMM.
//
// TMEM columns: [0, 192) the group's QKV accumulator, later y; [192, 320) and [320, 448) S of the two heads (O on top).
// What bounds it: TMEM -> register reads at ~64 B/clk per SM (678 KB per tile) and the serial MMA
// round trips of a tile; levels with fewer than ~100 window pairs keep the four-launch form (engine.cu).
//
// Warp roles: warps 0-7 workers (warp w owns TMEM lane quadrant w % 4 = token rows [32 (w % 4), +32); warps 0-3 take
// the first head of a group, warps 4-7 the second), warp 8 TMA producer (weights), warp 9 TMEM allocation + MMA issue.
#pragma once

#include "common.cuh"
#include "gn_stats.cuh"
#include "swin_attn_fused.cuh"

namespace rs {

struct SwinTcParams {
  CUtensorMap tmWqkv;                   // {E, 3E} fp16, box {64, 64}
  CUtensorMap tmWproj;                  // {E, E} fp16, box {64, E}
  SwinAttnParams a;
  long long* dbg;                       // optional: CTA 0 writes clock64 stamps of its first two tiles [2][64]
};

constexpr int kTcWorkers = 256;
constexpr int kTcThreads = 320;
constexpr int kTcTmaWarp = 8, kTcMmaWarp = 9;
constexpr int kTcSlotBytes = 24576;     // one weight tile: 192 rows x 64 fp16

template <int kE>
struct SwinTcSmem {
  static constexpr int kKB = kE / 64;                       // 64-channel k-blocks = head groups
  static constexpr int kSlots = kE == 192 ? 3 : 4;          // three slots = one whole group of weight tiles in flight
  static constexpr int off_xn = 0;                          // A operand of the QKV GEMMs; later O_1.. (k-blocks 1..) of the
                                                            // projection's A operand; y staging in the epilogue
  static constexpr int off_o = off_xn + kKB * 16384;        // O_0: k-block 0 of the projection's A operand (Xn is still live
                                                            // when group 0 finishes; the later groups' O go into the dead Xn)
  static constexpr int off_q = off_o + 16384;               // [128 tokens][64 ch]  (2 heads); Q / K / V^T later hold the residual
  static constexpr int off_k = off_q + 16384;               // [128 tokens][64 ch]
  static constexpr int off_vt = off_k + 16384;              // 2 token blocks x [64 ch][64 tokens]
  static constexpr int off_p = off_vt + 16384;              // 2 x [128 rows][64 keys]; norm1 scratch / statistics scratch
  static constexpr int off_ring = off_p + 32768;
  static constexpr int off_bars = off_ring + kSlots * kTcSlotBytes;
  static constexpr int off_pix = off_bars + 256;
  static constexpr int off_rpb = off_pix + 2 * 512;         // relative-position bias, compact: [heads][225] fp32
  static constexpr int off_ab = off_rpb + ((kE / 32) * 225 * 4 + 15) / 16 * 16;   // norm1 affine [2 windows][kE][2] fp32: kept
                                                            // across the CTA's tiles (recomputed when the image changes)
  // + slack for aligning the base.  The extern array is declared __align__(1024), which the toolchain honours by rounding
  // the (empty) static segment up to 1024 bytes — those count against the 227 KB limit — so the pad is 0 in practice;
  // the kernel traps if it ever exceeds the slack.
  static constexpr int kSlack = 256;
  static constexpr int total = off_ab + 2 * kE * 2 * 4 + kSlack;
  static_assert(total + 1024 <= 227 * 1024, "shared memory budget (dynamic + the 1024-byte static alignment segment)");
};

#ifdef __CUDACC__

// 16 fp32 accumulator values + 16 biases -> 16 fp16 in two 16-byte units
__device__ __forceinline__ void tc_pack16(const uint32_t (&v)[16], const float4 (&b)[4], uint4& o0, uint4& o1) {
  __half2* q0 = reinterpret_cast<__half2*>(&o0);
  __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float f0 = __uint_as_float(v[4 * i]) + b[i].x, f1 = __uint_as_float(v[4 * i + 1]) + b[i].y;
    const float f2 = __uint_as_float(v[4 * i + 2]) + b[i].z, f3 = __uint_as_float(v[4 * i + 3]) + b[i].w;
    if (i < 2) { q0[2 * i] = __floats2half2_rn(f0, f1); q0[2 * i + 1] = __floats2half2_rn(f2, f3); }
    else { q1[2 * (i - 2)] = __floats2half2_rn(f0, f1); q1[2 * (i - 2) + 1] = __floats2half2_rn(f2, f3); }
  }
}
__device__ __forceinline__ float tc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// warp-uniform NON-BLOCKING test of an mbarrier phase (every lane observes the completion); try_wait would suspend the
// warp for its time-out on each unready barrier and make the issuer's polling loop react late
__device__ __forceinline__ bool mbar_test_all(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return __all_sync(0xffffffffu, ok != 0 ? 1 : 0) != 0;
}

// all worker threads: deliver `add` channel-slots of image `img` to the sinks' arrival counters; the last arriver of an
// image reduces its pairs to the 32 group (mean, rstd) (gn_stats.cuh)
__device__ __noinline__ void swin_tc_arrive(const SwinAttnParams& p, int img, unsigned int add, int slots, int tid, int* s_flag) {
  const GnSink* const sk[2] = {&p.sink[0], (p.sink[1].part && p.sink[1].gstat) ? &p.sink[1] : nullptr};
  const int im[2] = {img, img};
  const unsigned int ad[2] = {add, add};
  gn_arrive<2>(sk, im, ad, slots, 64.0f, tid, kTcWorkers, 1, s_flag);
}

template <int kE>
__global__ void __launch_bounds__(kTcThreads, 1) swin_attn_tc_kernel(const __grid_constant__ SwinTcParams prm) {
  using L = SwinTcSmem<kE>;
  constexpr int kKB = L::kKB;
  constexpr int kG = kKB;                                    // head groups (two heads each)
  constexpr int kSlots = L::kSlots;
  constexpr int kUnits = kE / 8;                             // 16-byte units per token row
  constexpr uint32_t kTmS0 = 192, kTmS1 = 320;               // TMEM columns: [0,192) QKV_g / y; S_j at 192 + 128 j (O_j on top of S_j)
  const SwinAttnParams& p = prm.a;

  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  // (offset arithmetic on the __shared__ array itself: an integer round trip would lose the address space and turn every
  //  plain access below into a generic LD / ST)
  const uint32_t base_pad = (1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u;
  if (base_pad > (uint32_t)L::kSlack) __trap();
  uint8_t* smem = tc_smem_raw + base_pad;
  const uint32_t sb = smem_u32(smem);
  const uint32_t sXn = sb + L::off_xn, sO = sb + L::off_o, sQ = sb + L::off_q, sK = sb + L::off_k, sVT = sb + L::off_vt,
                 sP = sb + L::off_p, sRing = sb + L::off_ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bars);
  uint64_t* w_full = bars;                 // [4]
  uint64_t* w_empty = bars + 4;            // [4]
  uint64_t* xn_full = bars + 8;            // 8 worker warps
  uint64_t* acc_full = bars + 9;           // commit
  uint64_t* qkv_drained = bars + 10;       // 8 worker warps
  uint64_t* s_full = bars + 11;            // [2] commit
  uint64_t* p_full = bars + 13;            // [2] 4 warps of the head
  uint64_t* o_full = bars + 15;            // [2] commit
  uint64_t* o_ready = bars + 17;           // 8 worker warps
  uint64_t* y_full = bars + 18;            // commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  int* s_flag = reinterpret_cast<int*>(bars + 22);          // [2] "this CTA delivered the image's last pairs" (gn_arrive)
  int* sPix2 = reinterpret_cast<int*>(smem + L::off_pix);   // [2][128] token -> pixel row, or -1 (this tile's / the next tile's)
  float* sRpb = reinterpret_cast<float*>(smem + L::off_rpb);   // [heads][225]: bias(i, j) = sRpb[h][(yi - yj + 7) * 15 + (xi - xj + 7)]
  float* sAB = reinterpret_cast<float*>(smem + L::off_ab);  // [2 windows][kE][2] affine of norm1 (persists across tiles)
  // scratch inside the P buffers (dead at the start of a tile): only while the affine is being derived
  float* sCh = reinterpret_cast<float*>(smem + L::off_p);   // [2][kE][2] per-channel (mean, M2)
  float* sMR = sCh + 2 * kE * 2;                            // [2][32][2] group (mean, rstd)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nWx = p.W >> 3, nWy = p.H >> 3, nW = nWx * nWy;
  const int HW = p.H * p.W;
  const int num_pairs = (p.total_windows + 1) >> 1;
  const int pair_begin = (int)(((long long)blockIdx.x * num_pairs) / gridDim.x);
  const int pair_end = (int)(((long long)(blockIdx.x + 1) * num_pairs) / gridDim.x);
  long long* dbg = (prm.dbg && blockIdx.x == 0) ? prm.dbg : nullptr;
  const long long t_start = clock64();

  if (warp == kTcTmaWarp && lane == 0) {
    tma_prefetch_desc(&prm.tmWqkv); tma_prefetch_desc(&prm.tmWproj);
    for (int s = 0; s < 4; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(xn_full, 8); mbar_init(acc_full, 1); mbar_init(qkv_drained, 8);
    for (int j = 0; j < 2; ++j) { mbar_init(&s_full[j], 1); mbar_init(&p_full[j], 4); mbar_init(&o_full[j], 1); }
    mbar_init(o_ready, 8); mbar_init(y_full, 1);
    mbar_fence_init();
  }
  if (warp == kTcMmaWarp) { tmem_alloc_dyn(tmem_slot, 512u); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == kTcTmaWarp) {
    // ===================== TMA producer: the weight stream (static data: no dependency on the previous kernel) =====
    const bool el = elect_one();
    int slot = 0; uint32_t ph = 0;
    for (int pair = pair_begin; pair < pair_end; ++pair) {
      for (int t = 0; t < kG * kKB + kKB; ++t) {
        mbar_wait(&w_empty[slot], ph ^ 1);
        if (el) {
          uint8_t* dst = smem + L::off_ring + (size_t)slot * kTcSlotBytes;
          if (t < kG * kKB) {
            const int g = t / kKB, kb = t - g * kKB;
            mbar_arrive_expect_tx(&w_full[slot], 3 * 8192);
#pragma unroll
            for (int which = 0; which < 3; ++which)
              tma_load_2d(dst + which * 8192, &prm.tmWqkv, &w_full[slot], kb * 64, which * kE + g * 64);
          } else {
            const int kb = t - kG * kKB;
            mbar_arrive_expect_tx(&w_full[slot], (uint32_t)(kE * 128));
            tma_load_2d(dst, &prm.tmWproj, &w_full[slot], kb * 64, 0);
          }
        }
        if (++slot == kSlots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == kTcMmaWarp) {
    // ===================== MMA issuer =====================
    const bool el = elect_one();
    const uint32_t idesc_qkv = umma_idesc_f16(128, 192), idesc_s = umma_idesc_f16(128, 64), idesc_pv = umma_idesc_f16(128, 32),
                   idesc_proj = umma_idesc_f16(128, kE);
    int slot = 0; uint32_t ph = 0;
    uint32_t n_tile = 0, n_grp = 0;
    // one k-block of a QKV_g (N = 192) or projection (N = E) GEMM out of the weight ring
    auto issue_kb = [&](uint32_t a_tile, int kb, uint32_t idesc, uint64_t* done_bar) {
      const uint64_t adesc = umma_desc_sw128(a_tile);
      const uint64_t bdesc = umma_desc_sw128(sRing + (uint32_t)slot * kTcSlotBytes);
      if (el) {
        umma_f16(tmem_base, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        umma_commit(&w_empty[slot]);
        if (kb == kKB - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++slot == kSlots) { slot = 0; ph ^= 1; }
    };
    auto issue_pv = [&](int j) {
      if (el) {
        const uint64_t pd = umma_desc_sw128(sP + (uint32_t)j * 16384);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint64_t vd = umma_desc_sw128(sVT + (uint32_t)w * 8192 + (uint32_t)j * 4096);
          const uint32_t d = tmem_base + (j ? kTmS1 : kTmS0) + 32 * w;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, pd + 2 * k, vd + 2 * k, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[j]);
      }
      __syncwarp();
    };
    for (int pair = pair_begin; pair < pair_end; ++pair) {
      mbar_wait(xn_full, n_tile & 1);
      tc_fence_after();
      if (dbg && el && n_tile < 2) dbg[64 * n_tile + 32] = clock64() - t_start;
      for (int kb = 0; kb < kKB; ++kb) {
        mbar_wait(&w_full[slot], ph);
        tc_fence_after();
        issue_kb(sXn + (uint32_t)kb * 16384, kb, idesc_qkv, acc_full);
      }
      for (int g = 0; g < kG; ++g) {
        mbar_wait(qkv_drained, n_grp & 1);
        tc_fence_after();
        if (dbg && el && n_tile < 2) dbg[64 * n_tile + 33 + g * 4] = clock64() - t_start;
        // S_j = Q_h K_h^T for the two heads of the group: per window (key block) one N = 64 accumulator
        if (el) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t qd = umma_desc_sw128(sQ) + 4 * j;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const uint64_t kd = umma_desc_sw128(sK + (uint32_t)w * 8192) + 4 * j;
              const uint32_t d = tmem_base + (j ? kTmS1 : kTmS0) + 64 * w;
              umma_f16(d, qd, kd, idesc_s, 0u);
              umma_f16(d, qd + 2, kd + 2, idesc_s, 1u);
            }
            umma_commit(&s_full[j]);
          }
        }
        __syncwarp();
        // the next group's QKV GEMM runs under this group's softmax; PV_j goes out as soon as P_j is written — whichever
        // of the two is ready first (a k-block waiting for its weight tile must not hold back a finished head)
        int kb_next = (g + 1 < kG) ? 0 : kKB;
        // last group: the projection's k-blocks 0 .. kKB-2 go out early as well; only the last one waits for O of this group.
        // O of the earlier groups is written AFTER the worker's arrival on qkv_drained (the arrival only says "Q / K in place,
        // O has left TMEM"); what publishes those stores is the worker's next fence + arrival — p_full of this group.  So
        // the early k-blocks follow both PVs of this group (found by compute-sanitizer: results changed from run to run
        // when the issuer got ahead of the O_1 stores)
        int pj_next = 0;
        const int pj_early = (g + 1 < kG) ? 0 : kKB - 1;
        uint32_t pv_done = 0;
        uint32_t spins = 0;
        const uint64_t t0 = global_timer_ns();
        while (kb_next < kKB || pj_next < pj_early || pv_done != 3u) {
          bool progressed = false;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (!(pv_done & (1u << j)) && mbar_test_all(&p_full[j], n_grp & 1)) {
              tc_fence_after();
              if (dbg && el && n_tile < 2) dbg[64 * n_tile + 34 + g * 4 + j] = clock64() - t_start;
              issue_pv(j);
              pv_done |= 1u << j;
              progressed = true;
            }
          }
          if (kb_next < kKB && mbar_test_all(&w_full[slot], ph)) {
            tc_fence_after();
            issue_kb(sXn + (uint32_t)kb_next * 16384, kb_next, idesc_qkv, acc_full);
            ++kb_next;
            progressed = true;
          }
          if (pj_next < pj_early && pv_done == 3u && mbar_test_all(&w_full[slot], ph)) {
            tc_fence_after();
            issue_kb(pj_next == 0 ? sO : sXn + (uint32_t)pj_next * 16384, pj_next, idesc_proj, y_full);
            ++pj_next;
            progressed = true;
          }
          if (!progressed) {
            // back off: a tight test_wait loop would take issue slots from the two worker warps of this scheduler
            __nanosleep(40);
            if ((++spins & 0xFFu) == 0 && global_timer_ns() - t0 > 2000000000ull) {
              if (lane == 0) printf("rs: swin tc issuer timeout (block %d group %d)\n", blockIdx.x, g);
              __trap();
            }
          }
        }
        ++n_grp;
      }
      mbar_wait(o_ready, n_tile & 1);
      tc_fence_after();
      if (dbg && el && n_tile < 2) dbg[64 * n_tile + 46] = clock64() - t_start;
      for (int kb = kKB - 1; kb < kKB; ++kb) {               // (k-blocks 0 .. kKB-2 went out during the last group)
        mbar_wait(&w_full[slot], ph);
        tc_fence_after();
        issue_kb(kb == 0 ? sO : sXn + (uint32_t)kb * 16384, kb, idesc_proj, y_full);
      }
      if (dbg && el && n_tile < 2) dbg[64 * n_tile + 47] = clock64() - t_start;
      ++n_tile;
    }
  } else {
    // ===================== workers (8 warps) =====================
    const int quad = warp & 3, hf = warp >> 2;
    const int r = quad * 32 + lane;                          // this thread's token row / TMEM lane
    const int wi = r >> 6, ti = r & 63;                      // window of the pair, token inside the window
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tm_row = tmem_base + lane_base;
    const int rsw = r & 7;
    uint8_t* const pXn = smem + L::off_xn;                   // (plain C++ stores: the compiler may schedule them freely;
    uint8_t* const pO = smem + L::off_o;                     //  the phase-closing fences / arrivals are the ordering points)
    uint8_t* const pQ = smem + L::off_q;
    uint8_t* const pK = smem + L::off_k;
    uint8_t* const pVT = smem + L::off_vt;
    uint8_t* const pP = smem + L::off_p;
    uint32_t n_tile = 0, n_grp = 0;
    int cur_img[2] = {-1, -1};
    // gather / store geometry: a thread keeps ONE 16-byte unit (8 channels) and walks rows
    constexpr int kRowLanes = kTcWorkers / kUnits;           // 10 (E = 192: 240 threads active) / 32 (E = 64)
    constexpr int kIt = (64 + kRowLanes - 1) / kRowLanes;    // row iterations per window: 7 / 2
    const int gu = tid % kUnits, gr0 = tid / kUnits;
    const bool gact = gr0 < kRowLanes;
    const float kLog2e = 1.4426950408889634f;
    // pixel table of a window pair + gather of its 128 token rows (raw x) straight into their operand positions
    // (cp.async); the thread that copied a unit later normalises it: no barrier between copy and use
    auto issue_gather = [&](int pr, int* pix_tab) {
      if (tid < 128) {
        const int k = tid >> 6, tok = tid & 63;
        const int w2 = 2 * pr + k;
        int pix = -1;
        if (w2 < p.total_windows) {
          const int n = w2 / nW, rem = w2 % nW, wy = rem / nWx, wx = rem % nWx;
          const int yy = (wy * 8 + (tok >> 3) + p.shift) % p.H, xx = (wx * 8 + (tok & 7) + p.shift) % p.W;
          pix = (n * p.H + yy) * p.W + xx;
        }
        pix_tab[tid] = pix;
      }
      named_bar_sync(1, kTcWorkers);
      if (gact) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int it = 0; it < kIt; ++it) {
            const int lr = gr0 + it * kRowLanes;
            if (lr < 64) {
              const int row = k * 64 + lr;
              const int pix = pix_tab[row];
              uint8_t* dst = pXn + (gu >> 3) * 16384 + row * 128 + (((gu & 7) ^ (row & 7)) << 4);
              if (pix >= 0) cp_async_16(dst, p.x + (long long)pix * p.x_ld + gu * 8);
              else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
            }
          }
      }
      cp_async_commit();
    };
    // The dense [heads][64][64] bias is relative_position_bias_table gathered by relative_position_index (reference
    // models/swin_transformer.py:82-97,130-133): it depends on (yi - yj, xi - xj) only.  Keep the 225 distinct values per
    // head in shared memory (5.4 KB) — fetching each warp's 8 KB block of the dense table from L2 for every head of every
    // tile made all SMs hit the same few lines at the same time and cost ~3000 cycles per group (profiles/r2_s19).
    constexpr int kRpbPer = ((kE / 32) * 225 + kTcWorkers - 1) / kTcWorkers;
    float rpb[kRpbPer];                                      // requested now (static data), stored after the first gather is out
#pragma unroll
    for (int q = 0; q < kRpbPer; ++q) {
      const int t = tid + q * kTcWorkers;
      rpb[q] = 0.f;
      if (t < (kE / 32) * 225) {
        const int h = t / 225, d = t - h * 225;
        const int dy = d / 15 - 7, dx = d % 15 - 7;
        const int i = max(dy, 0) * 8 + max(dx, 0), j = max(-dy, 0) * 8 + max(-dx, 0);
        rpb[q] = __ldg(p.relbias + ((size_t)h * 64 + i) * 64 + j);
      }
    }
    // Producer-side finalisation of y's statistics (when the sinks carry gstat / counters): this CTA's windows are
    // consecutive, so the images it touches come in runs — the pairs delivered for the current image are counted and the
    // CTA arrives ONCE per (sink, image), when it moves on to the next image or after its last tile.  The last arriver of an
    // image reduces its pairs to the 32 group (mean, rstd) (gn_stats.cuh); different images end on different CTAs.
    const bool finalize = p.sink[0].part != nullptr && p.sink[0].gstat != nullptr;
    int pend_img = -1; unsigned int pend_add = 0;            // (identical in every worker thread)
    auto flush_arrival = [&]() {
      swin_tc_arrive(p, pend_img, pend_add, nW, tid, s_flag);          // (out of line: rare, and its registers stay out of the tile loop)
      pend_add = 0;
    };
    // Statistics of a finished tile: (mean, M2) per (window, channel) from the staged y tile (+ arrival bookkeeping).  For
    // all but a CTA's last tile this runs in the NEXT tile, while the workers would otherwise wait for its first QKV GEMM —
    // the staged tile (Q / K / V^T region) is overwritten only by that tile's first drain.
    const bool want_stats = p.sink[0].part != nullptr;
    uint8_t* const pStage = smem + L::off_q;                  // [128 rows][kE] fp16, 16-byte units XOR-swizzled by (row & 7)
    auto stats_pass = [&](int pr) {
      if (want_stats) {
        // (mean, M2) of the stored values per (window, channel): a second pass over the staged tile.  Thread = (window k,
        // unit u, sub): 16 rows x 8 channels, the four 16-row parts merged with Chan's formula through shuffles (sub = the
        // two low lane bits); fixed order.
        const int su = tid >> 2, sub = tid & 3;
        const bool sact = su < 2 * kUnits;
        const int k = sact ? su / kUnits : 0, u = sact ? su % kUnits : 0;
        float mean8[8], m28[8];
        {
          // (sum, sum of squares) over 16 fp16 values per column: q - s^2 / 16 in fp32 is benign for so few values (relative
          // error ~1e-4 of M2 even for |mean| = 60 std); every later combination is Chan's formula on (mean, M2) pairs
          float s1[8], s2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const int lr = 16 * sub + ((it + 2 * sub) & 15);          // (row & 7 differs between the four subs: fewer bank conflicts)
            const int row = k * 64 + lr;
            const uint4 raw = *reinterpret_cast<const uint4*>(pStage + row * (kE * 2) + ((u ^ (row & 7)) << 4));
            const __half2* hh = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(hh[j]);
              s1[2 * j] += f.x; s1[2 * j + 1] += f.y;
              s2[2 * j] = fmaf(f.x, f.x, s2[2 * j]); s2[2 * j + 1] = fmaf(f.y, f.y, s2[2 * j + 1]);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            mean8[j] = s1[j] * (1.0f / 16.0f);
            m28[j] = fmaxf(s2[j] - s1[j] * mean8[j], 0.f);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float m = mean8[j], q = m28[j];
          float mo = __shfl_xor_sync(0xffffffffu, m, 1), qo = __shfl_xor_sync(0xffffffffu, q, 1);
          float mm, qq;
          chan_merge_equal(16.f, (sub & 1) ? mo : m, (sub & 1) ? qo : q, (sub & 1) ? m : mo, (sub & 1) ? q : qo, mm, qq);
          mo = __shfl_xor_sync(0xffffffffu, mm, 2); qo = __shfl_xor_sync(0xffffffffu, qq, 2);
          chan_merge_equal(32.f, (sub & 2) ? mo : mm, (sub & 2) ? qo : qq, (sub & 2) ? mm : mo, (sub & 2) ? qq : qo, mean8[j], m28[j]);
        }
        const int w2 = 2 * pr + k;
        if (sact && sub == 0 && w2 < p.total_windows) {
          const int n = w2 / nW, slot = w2 % nW;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const GnSink& sk = p.sink[d];
            if (!sk.part) continue;
            float* dst = sk.part + (((size_t)n * nW + slot) * sk.cstride + sk.coff + u * 8) * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(mean8[2 * j], m28[2 * j], mean8[2 * j + 1], m28[2 * j + 1]);
          }
        }
      }
      if (finalize) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int w2 = 2 * pr + k;
          if (w2 >= p.total_windows) continue;
          const int n = w2 / nW;
          if (n != pend_img) {
            if (pend_add) flush_arrival();                   // (the pairs of THIS tile are not counted yet)
            pend_img = n;
          }
          pend_add += (unsigned int)kE;
        }
      }
    };
    pdl_wait();

    for (int pair = pair_begin; pair < pair_end; ++pair) {
      const bool stamp = dbg && tid == 0 && n_tile < 2;
      long long* const dbgw = dbg + 64 * (n_tile & 1);
      if (stamp) dbgw[0] = clock64() - t_start;
      // ---- geometry of the two windows ----
      int n_img[2], wy_[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int w2 = min(2 * pair + k, p.total_windows - 1);
        n_img[k] = w2 / nW;
        wy_[k] = (w2 % nW) / nWx;
      }
      int* const sPix = sPix2 + 128 * (n_tile & 1);
      if (n_tile == 0) {
        issue_gather(pair, sPix);                            // (later tiles: issued under the previous tile's epilogue)
#pragma unroll
        for (int q = 0; q < kRpbPer; ++q)
          if (tid + q * kTcWorkers < (kE / 32) * 225) sRpb[tid + q * kTcWorkers] = rpb[q];   // (visible after the barriers below)
      }
      // ---- norm1 affine of the windows' images (recomputed only when the image changes) ----
      const bool need_ab = (n_img[0] != cur_img[0]) || (n_img[1] != cur_img[1]);     // uniform
      if (need_ab) {
        constexpr int cpg = kE / 32;
        const int nset = (n_img[0] == n_img[1]) ? 1 : 2;     // both windows in one image: derive once, store twice
        float gam[(2 * kE + kTcWorkers - 1) / kTcWorkers], bet[(2 * kE + kTcWorkers - 1) / kTcWorkers];   // this thread's gamma / beta: in flight early
#pragma unroll
        for (int q = 0; q < (2 * kE + kTcWorkers - 1) / kTcWorkers; ++q) {
          const int idx = tid + q * kTcWorkers;
          const int c = idx % kE;
          gam[q] = __ldg(p.gamma + c); bet[q] = __ldg(p.beta + c);
        }
        if (p.gn_gstat) {
          if (tid < 64) {
            const int k = tid >> 5, gg = tid & 31;
            const float2 mr = ldcg_f2(p.gn_gstat + ((size_t)(k ? n_img[1] : n_img[0]) * 32 + gg) * 2);
            sMR[(k * 32 + gg) * 2] = mr.x; sMR[(k * 32 + gg) * 2 + 1] = mr.y;
          }
        } else {
          const float ns = (float)HW / (float)p.gn_slots;
          for (int idx = tid; idx < nset * kE; idx += kTcWorkers) {
            const int k = idx / kE, c = idx - k * kE;
            const float* pc = p.gn_part + (size_t)(k ? n_img[1] : n_img[0]) * p.gn_slots * kE * 2 + (size_t)c * 2;
            float2 mq;
            if (p.gn_slots <= 32) {
              // same arithmetic and order as gn_channel_from_pairs, but every slot's pair is requested before the first use
              float2 e[32];
#pragma unroll
              for (int sl = 0; sl < 32; ++sl) e[sl] = sl < p.gn_slots ? *reinterpret_cast<const float2*>(pc + (size_t)sl * kE * 2) : make_float2(0.f, 0.f);
              const float pivot = e[0].x;
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int sl = 0; sl < 32; ++sl)
                if (sl < p.gn_slots) { const float d = e[sl].x - pivot; s1 += d; s2 += fmaf(ns * d, d, e[sl].y); }
              const float dm = s1 / (float)p.gn_slots;
              mq = make_float2(pivot + dm, fmaxf(s2 - ns * (float)p.gn_slots * dm * dm, 0.f));
            } else {
              mq = gn_channel_from_pairs(pc, p.gn_slots, kE, ns);
            }
            sCh[(k * kE + c) * 2] = mq.x; sCh[(k * kE + c) * 2 + 1] = mq.y;
          }
          named_bar_sync(1, kTcWorkers);
          if (tid < 32 * nset) {
            const int k = tid >> 5, gg = tid & 31;
            float chp[2 * cpg];
#pragma unroll
            for (int j = 0; j < cpg; ++j) { chp[2 * j] = sCh[(k * kE + gg * cpg + j) * 2]; chp[2 * j + 1] = sCh[(k * kE + gg * cpg + j) * 2 + 1]; }
            const float2 mr = gn_group_from_channels(chp, cpg, (float)HW, p.eps);
            sMR[(k * 32 + gg) * 2] = mr.x; sMR[(k * 32 + gg) * 2 + 1] = mr.y;
            if (nset == 1) { sMR[((32 + gg)) * 2] = mr.x; sMR[((32 + gg)) * 2 + 1] = mr.y; }
          }
        }
        named_bar_sync(1, kTcWorkers);
#pragma unroll
        for (int q = 0; q < (2 * kE + kTcWorkers - 1) / kTcWorkers; ++q) {
          const int idx = tid + q * kTcWorkers;
          if (idx >= 2 * kE) break;
          const int k = idx / kE, c = idx - k * kE, gg = c / cpg;
          const float a = sMR[(k * 32 + gg) * 2 + 1] * gam[q];
          const float b = bet[q] - sMR[(k * 32 + gg) * 2] * a;
          sAB[(k * kE + c) * 2] = a; sAB[(k * kE + c) * 2 + 1] = b;
        }
        cur_img[0] = n_img[0]; cur_img[1] = n_img[1];
        named_bar_sync(1, kTcWorkers);
      }
      if (stamp) dbgw[1] = clock64() - t_start;
      // ---- normalise in place: unit gu of rows gr0, gr0 + kRowLanes, ... of each window ----
      cp_async_wait<0>();
      if (gact) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if ((2 * pair + k) >= p.total_windows) continue;   // (rows of a missing window stay zero)
          float4 ab[4];                                      // (a, b) of channels 2j, 2j + 1 of this unit, window k
#pragma unroll
          for (int j = 0; j < 4; ++j) ab[j] = *reinterpret_cast<const float4*>(sAB + ((size_t)k * kE + gu * 8) * 2 + 4 * j);
#pragma unroll
          for (int it = 0; it < kIt; ++it) {
            const int lr = gr0 + it * kRowLanes;
            if (lr < 64) {
              const int row = k * 64 + lr;
              uint4* ptr = reinterpret_cast<uint4*>(pXn + (gu >> 3) * 16384 + row * 128 + (((gu & 7) ^ (row & 7)) << 4));
              uint4 raw = *ptr;
              __half2* hh = reinterpret_cast<__half2*>(&raw);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(hh[j]);
                f.x = fmaf(f.x, ab[j].x, ab[j].y);
                f.y = fmaf(f.y, ab[j].z, ab[j].w);
                hh[j] = __floats2half2_rn(f.x, f.y);
              }
              *ptr = raw;
            }
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();                                     // (this warp's TMEM reads of the previous tile's y are complete)
      named_bar_sync(1, kTcWorkers);                         // the affine scratch (inside the P buffers) is dead from here on
      if (lane == 0) mbar_arrive(xn_full);
      if (stamp) dbgw[2] = clock64() - t_start;

      // shifted-window mask (reference quirk: the label depends on the window row and the token COLUMN, see
      // window_attn.cuh): mval[b] = -100 when key column b carries another region label than this query row
      float mval[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) mval[b] = 0.f;
      if (p.shift) {
        const int wyw = wi ? wy_[1] : wy_[0];
        const int la = swin_label(wyw, ti & 7, p.H, p.shift);
#pragma unroll
        for (int b = 0; b < 8; ++b) if (swin_label(wyw, b, p.H, p.shift) != la) mval[b] = -100.0f;
      }

      // Per group g:  drain Q_g / K_g (+ bias, fp16) -> S -> V_g^T stores, softmax -> P -> PV -> O_g.  The two MMA round trips
      // of a group (S, PV) are the serial part: the Q / K drain of group g + 1 (TMEM-bound, ~1000 cycles) is pulled into
      // the PV wait of group g.  warps 0-3: Q_g (4 chunks of 16 columns) + V columns [0, 32);  warps 4-7: K_g + V [32, 64).
      uint32_t vv[2][16];                                    // this thread's V values of the group being set up (raw accumulators)
      auto drain_qk = [&](int g) {
        const float4* bq4 = reinterpret_cast<const float4*>(p.bqkv + (hf ? kE : 0) + g * 64);
        uint32_t v[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16(tm_row + (uint32_t)(hf * 64 + c * 16), v[c]);
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld16(tm_row + (uint32_t)(128 + hf * 32 + c * 16), vv[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_wait16(v[c]);
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld_wait16(vv[c]);
        uint8_t* dstQK = (hf ? pK : pQ) + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o0, o1;
          float4 b4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) b4[i] = __ldg(bq4 + c * 4 + i);
          tc_pack16(v[c], b4, o0, o1);
          *reinterpret_cast<uint4*>(dstQK + (((2 * c) ^ rsw) << 4)) = o0;
          *reinterpret_cast<uint4*>(dstQK + (((2 * c + 1) ^ rsw) << 4)) = o1;
        }
        fence_proxy_async_smem();
        tc_fence_before();                                   // every TMEM read of the accumulator is complete
      };
      if (n_tile > 0) {                                      // the previous tile's statistics, while QKV_0 of this one runs
        stats_pass(pair - 1);
        named_bar_sync(1, kTcWorkers);                       // (every thread has read the staged tile before the first drain writes Q / K)
      }
      // group 0: last tile's MMAs have long read Q / K / V^T / P (both o_full of its last group were waited there)
      mbar_wait(acc_full, n_grp & 1);
      tc_fence_after();
      if (stamp) dbgw[3] = clock64() - t_start;
      drain_qk(0);
      __syncwarp();
      if (lane == 0) mbar_arrive(qkv_drained);               // S needs Q, K of all eight warps; the next QKV GEMM may start
      if (stamp) dbgw[4] = clock64() - t_start;

      for (int g = 0; g < kG; ++g) {
        // ---- V_g^T (needed by PV only; covered by p_full): element (channel cr, token r) -> token block r / 64, row cr,
        //      column r % 64 ----
        {
          const float* bv1 = p.bqkv + 2 * kE + g * 64 + hf * 32;
          uint8_t* vt_tok = pVT + wi * 8192 + (r & 7) * 2;
          const int tu = ti >> 3;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int cr = hf * 32 + c * 16 + i;
              const float f = __uint_as_float(vv[c][i]) + __ldg(bv1 + c * 16 + i);
              *reinterpret_cast<__half*>(vt_tok + cr * 128 + ((tu ^ (cr & 7)) << 4)) = __float2half_rn(f);
            }
          }
        }

        // ---- softmax of head h = 2 g + hf: this thread owns query row r (token ti of window wi) ----
        // bias(i, j) = table[(yi - yj + 7) * 15 + (xi - xj + 7)] = tab_i[-(15 yj + xj)]: a compile-time offset per key j
        const float* tab_i = sRpb + (2 * g + hf) * 225 + (ti >> 3) * 15 + (ti & 7) + 112;
        float bias[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) bias[j] = tab_i[-(15 * (j >> 3) + (j & 7))];
        const uint32_t tmS = tm_row + (hf ? kTmS1 : kTmS0);
        mbar_wait(&s_full[hf], n_grp & 1);
        tc_fence_after();
        if (stamp) dbgw[5 + g * 6] = clock64() - t_start;
        float s[64];
        {
          uint32_t sv[4][16];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld16(tmS + (uint32_t)(64 * wi + 16 * c), sv[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_wait16(sv[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[16 * c + i] = __uint_as_float(sv[c][i]);
        }
        // t = s * scale + bias (+ mask);  p = exp(t - max) = ex2(t * log2e - max * log2e);  four max / sum chains for ILP
        float mx4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (int j = 0; j < 64; ++j) { s[j] = fmaf(s[j], p.scale, bias[j]) + mval[j & 7]; mx4[j & 3] = fmaxf(mx4[j & 3], s[j]); }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float nmx = -mx * kLog2e;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t* prow = pP + hf * 16384 + r * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint4 q;
          uint32_t* qq = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float e0 = tc_ex2(fmaf(s[8 * u + 2 * i], kLog2e, nmx)), e1 = tc_ex2(fmaf(s[8 * u + 2 * i + 1], kLog2e, nmx));
            sum4[i] += e0 + e1;
            qq[i] = pack_h2(e0, e1);
          }
          *reinterpret_cast<uint4*>(prow + ((u ^ rsw) << 4)) = q;
        }
        const float inv = 1.0f / ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[hf]);
        if (stamp) dbgw[6 + g * 6] = clock64() - t_start;

        // ---- while PV runs: Q / K of the next group (S of BOTH heads of this group must have read Q / K) ----
        if (g + 1 < kG) {
          mbar_wait(acc_full, (n_grp + 1) & 1);
          mbar_wait(&s_full[hf ^ 1], n_grp & 1);
          tc_fence_after();
          drain_qk(g + 1);
        }

        // ---- O_h = (P V) / rowsum -> fp16, into the projection's A operand (k-block g, 64-byte half hf) ----
        mbar_wait(&o_full[hf], n_grp & 1);
        tc_fence_after();
        if (stamp) dbgw[7 + g * 6] = clock64() - t_start;
        {
          uint32_t ov[2][16];
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld16(tmS + (uint32_t)(32 * wi + 16 * c), ov[c]);
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld_wait16(ov[c]);
          if (g + 1 < kG) {
            // O_g has left TMEM (S of the next group lands on the same columns) and Q / K of the next group are in place
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qkv_drained);
          }
          // O_0 has its own tile; O_g (g >= 1) goes into k-block g of the Xn region, dead once the LAST group's QKV GEMM
          // has completed (it was issued under this group's softmax at the latest)
          if (g >= 1) mbar_wait(acc_full, (n_grp + (uint32_t)(kG - 1 - g)) & 1);
          uint8_t* orow = (g == 0 ? pO : pXn + g * 16384) + r * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 q0, q1;
            uint32_t* a0 = reinterpret_cast<uint32_t*>(&q0);
            uint32_t* a1 = reinterpret_cast<uint32_t*>(&q1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              a0[i] = pack_h2(__uint_as_float(ov[c][2 * i]) * inv, __uint_as_float(ov[c][2 * i + 1]) * inv);
              a1[i] = pack_h2(__uint_as_float(ov[c][8 + 2 * i]) * inv, __uint_as_float(ov[c][8 + 2 * i + 1]) * inv);
            }
            *reinterpret_cast<uint4*>(orow + (((4 * hf + 2 * c) ^ rsw) << 4)) = q0;
            *reinterpret_cast<uint4*>(orow + (((4 * hf + 2 * c + 1) ^ rsw) << 4)) = q1;
          }
        }
        if (stamp) dbgw[8 + g * 6] = clock64() - t_start;
        // the other head's PV has read V^T (and its P rows): the next group's V^T stores / the residual copies may follow
        mbar_wait(&o_full[hf ^ 1], n_grp & 1);
        ++n_grp;
      }
      // the residual (raw x row of this thread, its half of the columns) flies while the projection runs: cp.async into the
      // Q / K / V^T region — dead: BOTH heads' last PV has completed (waited at the end of the group loop)
      constexpr int kHalfCols = kE / 2;                       // columns per worker half
      constexpr int kCh = kHalfCols / 16;                     // 16-column chunks per thread: 6 / 2
      uint8_t* const pRes = pStage + r * (kE * 2);            // the residual, then (in place) the staged y tile
      const int mypix = sPix[r];
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_ready);
      if (mypix >= 0) {
#pragma unroll
        for (int u = 0; u < 2 * kCh; ++u)
          cp_async_16(pRes + (((hf * 2 * kCh + u) ^ rsw) << 4), p.x + (long long)mypix * p.x_ld + hf * kHalfCols + u * 8);
      }
      cp_async_commit();

      // ================= epilogue: y = acc + b + x -> fp16 (one rounding), full-row stores, statistics =================
      mbar_wait(y_full, n_tile & 1);
      tc_fence_after();
      cp_async_wait<0>();                                    // this thread's residual units (issued before the o_ready arrival) have landed
      if (stamp) dbgw[24] = clock64() - t_start;
      // every MMA of this tile has completed: the Xn tile is free — the next tile's rows start flying now, under the epilogue
      if (pair + 1 < pair_end) issue_gather(pair + 1, sPix2 + 128 * ((n_tile + 1) & 1));
#pragma unroll
      for (int c0 = 0; c0 < kCh; c0 += 3) {
        constexpr int kB = kCh < 3 ? kCh : 3;
        uint32_t v[kB][16];
#pragma unroll
        for (int c = 0; c < kB; ++c) tmem_ld16(tm_row + (uint32_t)(hf * kHalfCols + (c0 + c) * 16), v[c]);
#pragma unroll
        for (int c = 0; c < kB; ++c) tmem_ld_wait16(v[c]);
#pragma unroll
        for (int c = 0; c < kB; ++c) {
          const int col = hf * kHalfCols + (c0 + c) * 16;
          float f[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bproj + col) + i);
            f[4 * i] = __uint_as_float(v[c][4 * i]) + b4.x; f[4 * i + 1] = __uint_as_float(v[c][4 * i + 1]) + b4.y;
            f[4 * i + 2] = __uint_as_float(v[c][4 * i + 2]) + b4.z; f[4 * i + 3] = __uint_as_float(v[c][4 * i + 3]) + b4.w;
          }
          uint4 x0 = make_uint4(0, 0, 0, 0), x1 = make_uint4(0, 0, 0, 0);
          if (mypix >= 0) {
            x0 = *reinterpret_cast<const uint4*>(pRes + (((hf * 2 * kCh + 2 * (c0 + c)) ^ rsw) << 4));
            x1 = *reinterpret_cast<const uint4*>(pRes + (((hf * 2 * kCh + 2 * (c0 + c) + 1) ^ rsw) << 4));
          }
          const __half2* h0p = reinterpret_cast<const __half2*>(&x0);
          const __half2* h1p = reinterpret_cast<const __half2*>(&x1);
          uint4 o0, o1;
          __half2* q0 = reinterpret_cast<__half2*>(&o0);
          __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a0 = __half22float2(h0p[i]), a1 = __half22float2(h1p[i]);
            q0[i] = __floats2half2_rn(f[2 * i] + a0.x, f[2 * i + 1] + a0.y);
            q1[i] = __floats2half2_rn(f[8 + 2 * i] + a1.x, f[8 + 2 * i + 1] + a1.y);
          }
          // staged over the residual units this thread has just read (the Xn tile already receives the next tile's rows)
          *reinterpret_cast<uint4*>(pRes + (((hf * 2 * kCh + 2 * (c0 + c)) ^ rsw) << 4)) = o0;
          *reinterpret_cast<uint4*>(pRes + (((hf * 2 * kCh + 2 * (c0 + c) + 1) ^ rsw) << 4)) = o1;
        }
      }
      tc_fence_before();
      named_bar_sync(1, kTcWorkers);
      if (stamp) dbgw[25] = clock64() - t_start;
      // full token rows to global: unit gu of rows gr0, gr0 + kRowLanes, ...
      if (gact) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int it = 0; it < kIt; ++it) {
            const int lr = gr0 + it * kRowLanes;
            if (lr < 64) {
              const int row = k * 64 + lr;
              const int pix = sPix[row];
              if (pix >= 0)
                *reinterpret_cast<uint4*>(p.y + (long long)pix * p.y_ld + gu * 8) =
                    *reinterpret_cast<const uint4*>(pStage + row * (kE * 2) + ((gu ^ (row & 7)) << 4));
            }
          }
      }
      if (pair + 1 == pair_end) stats_pass(pair);            // (earlier tiles: under the next tile's first QKV GEMM)
      named_bar_sync(1, kTcWorkers);                         // staging / scratch / pixel table are free for the next tile
      if (stamp) dbgw[26] = clock64() - t_start;
      ++n_tile;
    }
    if (finalize && pend_add) flush_arrival();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) { tc_fence_after(); tmem_dealloc_dyn(tmem_base, 512u); }
}

#endif  // __CUDACC__
}  // namespace rs
