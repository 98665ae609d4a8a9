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
//   * a CTA owns TWO 8x8 windows = 128 token rows = one UMMA M.  The rows are gathered with plain 16-byte loads (cyclic
//     shift and window partition are address arithmetic), normalised in registers and written into shared memory
//     directly in the K-major 128B-swizzled operand layout (what a TMA load would have produced);
//   * the heads are walked in GROUPS of two (64 channels): [Q_g | K_g | V_g] = Xn . W_g^T is one N = 192 accumulator
//     (weight rows stream through a TMA ring), drained to shared memory as fp16 operands — Q_g / K_g row-major (K-major
//     A / B operands of QK^T), V_g TRANSPOSED (the K-major B operand of PV wants [head_dim][token]);
//   * per head: S = Q_h K_h^T as two N = 64 MMAs (the keys of window 0, the keys of window 1): all 128 rows are multiplied
//     against each window's keys and each row simply reads the 64 columns of ITS window — block-diagonal attention
//     without M = 64 instructions.  Softmax: one thread per query row (no shuffles), P written as the next A operand.
//     O = P V the same way (two N = 32 MMAs into the TMEM columns S just vacated), scaled by 1 / rowsum and written into
//     the A operand of the projection;
//   * while the 128 threads of a head do softmax the tensor pipe already runs the NEXT group's QKV GEMM;
//   * y = O W_proj^T + b + x: N = E accumulator, + raw x re-read from L2, one rounding, staged in shared memory, written as
//     full token rows; (mean, M2) per (window, channel) for the norm2 that follows.
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
  long long* dbg;                       // optional: CTA 0 writes clock64 stamps of its first tile [64]
};

constexpr int kTcWorkers = 256;
constexpr int kTcThreads = 320;
constexpr int kTcTmaWarp = 8, kTcMmaWarp = 9;
constexpr int kTcSlotBytes = 24576;     // one weight tile: 192 rows x 64 fp16

template <int kE>
struct SwinTcSmem {
  static constexpr int kKB = kE / 64;                       // 64-channel k-blocks = head groups
  static constexpr int kSlots = kE == 192 ? 2 : 4;
  static constexpr int off_xn = 0;                          // A operand of the QKV GEMMs; y staging in the epilogue
  static constexpr int off_o = off_xn + kKB * 16384;        // A operand of the projection
  static constexpr int off_q = off_o + kKB * 16384;         // [128 tokens][64 ch]  (2 heads)
  static constexpr int off_k = off_q + 16384;               // [128 tokens][64 ch]
  static constexpr int off_vt = off_k + 16384;              // 2 token blocks x [64 ch][64 tokens]
  static constexpr int off_p = off_vt + 16384;              // 2 x [128 rows][64 keys]; norm1 scratch / statistics scratch
  static constexpr int off_ring = off_p + 32768;
  static constexpr int off_bars = off_ring + kSlots * kTcSlotBytes;
  static constexpr int off_pix = off_bars + 256;
  static constexpr int total = off_pix + 512 + 1024;        // + slack for the 1024-byte alignment of the base
};

#ifdef __CUDACC__

// 16 fp32 accumulator values (+ 16 biases) -> 16 fp16 in two 16-byte units
__device__ __forceinline__ void tc_pack16(const uint32_t (&v)[16], const float* __restrict__ bias, uint4& o0, uint4& o1) {
  __half2* q0 = reinterpret_cast<__half2*>(&o0);
  __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + i);
    const float f0 = __uint_as_float(v[4 * i]) + b.x, f1 = __uint_as_float(v[4 * i + 1]) + b.y;
    const float f2 = __uint_as_float(v[4 * i + 2]) + b.z, f3 = __uint_as_float(v[4 * i + 3]) + b.w;
    if (i < 2) { q0[2 * i] = __floats2half2_rn(f0, f1); q0[2 * i + 1] = __floats2half2_rn(f2, f3); }
    else { q1[2 * (i - 2)] = __floats2half2_rn(f0, f1); q1[2 * (i - 2) + 1] = __floats2half2_rn(f2, f3); }
  }
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
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
  int* sPix = reinterpret_cast<int*>(smem + L::off_pix);    // [128] token -> pixel row, or -1
  // scratch inside the P buffers (dead at the start and at the end of a tile)
  float* sAB = reinterpret_cast<float*>(smem + L::off_p);   // [2 windows][kE][2] affine of norm1
  float* sCh = sAB + 2 * kE * 2;                            // [2][kE][2] per-channel (mean, M2)
  float* sMR = sCh + 2 * kE * 2;                            // [2][32][2] group (mean, rstd)
  float* sStat = reinterpret_cast<float*>(smem + L::off_p); // [4 quads][kE][2] (epilogue)

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
    auto issue_qkv = [&]() {
      for (int kb = 0; kb < kKB; ++kb) {
        mbar_wait(&w_full[slot], ph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(sXn + (uint32_t)kb * 16384);
        const uint64_t bdesc = umma_desc_sw128(sRing + (uint32_t)slot * kTcSlotBytes);
        if (el) {
          umma_f16(tmem_base, adesc, bdesc, idesc_qkv, kb != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_qkv, 1u);
          umma_commit(&w_empty[slot]);
          if (kb == kKB - 1) umma_commit(acc_full);
        }
        if (++slot == kSlots) { slot = 0; ph ^= 1; }
      }
    };
    for (int pair = pair_begin; pair < pair_end; ++pair) {
      mbar_wait(xn_full, n_tile & 1);
      tc_fence_after();
      if (dbg && el && n_tile == 0) dbg[32] = clock64() - t_start;
      issue_qkv();
      for (int g = 0; g < kG; ++g) {
        mbar_wait(qkv_drained, n_grp & 1);
        tc_fence_after();
        if (dbg && el && n_tile == 0) dbg[33 + g * 4] = clock64() - t_start;
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
        if (g + 1 < kG) issue_qkv();                         // the next group's GEMM runs under this group's softmax
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          mbar_wait(&p_full[j], n_grp & 1);
          tc_fence_after();
          if (dbg && el && n_tile == 0) dbg[34 + g * 4 + j] = clock64() - t_start;
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
        }
        ++n_grp;
      }
      mbar_wait(o_ready, n_tile & 1);
      tc_fence_after();
      if (dbg && el && n_tile == 0) dbg[46] = clock64() - t_start;
      for (int kb = 0; kb < kKB; ++kb) {
        mbar_wait(&w_full[slot], ph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(sO + (uint32_t)kb * 16384);
        const uint64_t bdesc = umma_desc_sw128(sRing + (uint32_t)slot * kTcSlotBytes);
        if (el) {
          umma_f16(tmem_base, adesc, bdesc, idesc_proj, kb != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_proj, 1u);
          umma_commit(&w_empty[slot]);
          if (kb == kKB - 1) umma_commit(y_full);
        }
        if (++slot == kSlots) { slot = 0; ph ^= 1; }
      }
      if (dbg && el && n_tile == 0) dbg[47] = clock64() - t_start;
      ++n_tile;
    }
  } else {
    // ===================== workers (8 warps) =====================
    const int quad = warp & 3, hf = warp >> 2;
    const int r = quad * 32 + lane;                          // this thread's token row / TMEM lane
    const int wi = r >> 6, ti = r & 63;                      // window of the pair, token inside the window
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tm_row = tmem_base + lane_base;
    const uint32_t rsw = (uint32_t)(r & 7);
    uint32_t n_tile = 0, n_grp = 0;
    int cur_img[2] = {-1, -1};
    pdl_wait();

    for (int pair = pair_begin; pair < pair_end; ++pair) {
      const bool stamp = dbg && tid == 0 && n_tile == 0;
      if (stamp) dbg[0] = clock64() - t_start;
      // ---- geometry of the two windows ----
      int n_img[2], wy_[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int w2 = min(2 * pair + k, p.total_windows - 1);
        n_img[k] = w2 / nW;
        wy_[k] = (w2 % nW) / nWx;
      }
      if (tid < 128) {
        const int k = tid >> 6, tok = tid & 63;
        const int w2 = 2 * pair + k;
        int pix = -1;
        if (w2 < p.total_windows) {
          const int n = w2 / nW, rem = w2 % nW, wy = rem / nWx, wx = rem % nWx;
          const int yy = (wy * 8 + (tok >> 3) + p.shift) % p.H, xx = (wx * 8 + (tok & 7) + p.shift) % p.W;
          pix = (n * p.H + yy) * p.W + xx;
        }
        sPix[tid] = pix;
      }
      named_bar_sync(1, kTcWorkers);
      // ---- gather the 128 token rows (raw x) into registers: the loads fly while the affine is derived ----
      constexpr int kPer = 128 * kUnits / kTcWorkers;        // 12 (E = 192) / 4 (E = 64)
      uint4 xr[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kTcWorkers;
        const int row = i / kUnits, unit = i - row * kUnits;
        const int pix = sPix[row];
        xr[k] = make_uint4(0, 0, 0, 0);
        if (pix >= 0) xr[k] = *reinterpret_cast<const uint4*>(p.x + (long long)pix * p.x_ld + unit * 8);
      }
      // ---- norm1 affine of the windows' images (recomputed only when the image changes) ----
      const bool need_ab = (n_img[0] != cur_img[0]) || (n_img[1] != cur_img[1]);     // uniform
      if (need_ab) {
        constexpr int cpg = kE / 32;
        if (p.gn_gstat) {
          if (tid < 64) {
            const int k = tid >> 5, gg = tid & 31;
            const float2 mr = ldcg_f2(p.gn_gstat + ((size_t)n_img[k] * 32 + gg) * 2);
            sMR[(k * 32 + gg) * 2] = mr.x; sMR[(k * 32 + gg) * 2 + 1] = mr.y;
          }
        } else {
          const float ns = (float)HW / (float)p.gn_slots;
          for (int idx = tid; idx < 2 * kE; idx += kTcWorkers) {
            const int k = idx / kE, c = idx - k * kE;
            const float2 mq = gn_channel_from_pairs(p.gn_part + (size_t)n_img[k] * p.gn_slots * kE * 2 + (size_t)c * 2, p.gn_slots, kE, ns);
            sCh[(k * kE + c) * 2] = mq.x; sCh[(k * kE + c) * 2 + 1] = mq.y;
          }
          named_bar_sync(1, kTcWorkers);
          if (tid < 64) {
            const int k = tid >> 5, gg = tid & 31;
            float chp[2 * cpg];
#pragma unroll
            for (int j = 0; j < cpg; ++j) { chp[2 * j] = sCh[(k * kE + gg * cpg + j) * 2]; chp[2 * j + 1] = sCh[(k * kE + gg * cpg + j) * 2 + 1]; }
            const float2 mr = gn_group_from_channels(chp, cpg, (float)HW, p.eps);
            sMR[(k * 32 + gg) * 2] = mr.x; sMR[(k * 32 + gg) * 2 + 1] = mr.y;
          }
        }
        named_bar_sync(1, kTcWorkers);
        for (int idx = tid; idx < 2 * kE; idx += kTcWorkers) {
          const int k = idx / kE, c = idx - k * kE, gg = c / cpg;
          const float a = sMR[(k * 32 + gg) * 2 + 1] * __ldg(p.gamma + c);
          const float b = __ldg(p.beta + c) - sMR[(k * 32 + gg) * 2] * a;
          sAB[(k * kE + c) * 2] = a; sAB[(k * kE + c) * 2 + 1] = b;
        }
        cur_img[0] = n_img[0]; cur_img[1] = n_img[1];
        named_bar_sync(1, kTcWorkers);
      }
      if (stamp) dbg[1] = clock64() - t_start;
      // ---- normalise, write the A operand (K-major, 128B swizzle) ----
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kTcWorkers;
        const int row = i / kUnits, unit = i - row * kUnits;
        const float* ab = sAB + ((size_t)(row >> 6) * kE + unit * 8) * 2;
        uint4 raw = xr[k];
        __half2* hh = reinterpret_cast<__half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 c4 = *reinterpret_cast<const float4*>(ab + 4 * j);      // (a, b) of channels 2j, 2j + 1
          float2 f = __half22float2(hh[j]);
          f.x = fmaf(f.x, c4.x, c4.y);
          f.y = fmaf(f.y, c4.z, c4.w);
          hh[j] = __floats2half2_rn(f.x, f.y);
        }
        if (sPix[row] < 0) raw = make_uint4(0, 0, 0, 0);
        st_shared_v4(sXn + (uint32_t)(unit >> 3) * 16384 + (uint32_t)row * 128 + ((((uint32_t)unit & 7) ^ ((uint32_t)row & 7)) << 4), raw.x, raw.y, raw.z, raw.w);
      }
      fence_proxy_async_smem();
      tc_fence_before();                                     // (this warp's TMEM reads of the previous tile's y are complete)
      named_bar_sync(1, kTcWorkers);                         // the affine scratch (inside the P buffers) is dead from here on
      if (lane == 0) mbar_arrive(xn_full);
      if (stamp) dbg[2] = clock64() - t_start;

      // shifted-window mask: bit b of mbits = key column b has a different region label than this row (reference quirk:
      // the label depends on the window row and the token COLUMN, see window_attn.cuh)
      uint32_t mbits = 0;
      if (p.shift) {
        const int la = swin_label(wy_[wi], ti & 7, p.H, p.shift);
#pragma unroll
        for (int b = 0; b < 8; ++b) if (swin_label(wy_[wi], b, p.H, p.shift) != la) mbits |= 1u << b;
      }

      for (int g = 0; g < kG; ++g) {
        // ---- drain [Q_g | K_g | V_g] (+ bias, fp16) into the attention operands ----
        mbar_wait(acc_full, n_grp & 1);
        if (n_grp > 0) { mbar_wait(&o_full[0], (n_grp - 1) & 1); mbar_wait(&o_full[1], (n_grp - 1) & 1); }   // last group's MMAs have read Q / K / V^T / P
        tc_fence_after();
        if (stamp) dbg[3 + g * 6] = clock64() - t_start;
        {
          // warps 0-3: Q_g (4 chunks of 16 columns) + V columns [0, 32);  warps 4-7: K_g + V columns [32, 64)
          const uint32_t dstQK = (hf ? sK : sQ) + (uint32_t)r * 128;
          const float* bqk = p.bqkv + (hf ? kE : 0) + g * 64;
          uint32_t v[4][16];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld16(tm_row + (uint32_t)(hf * 64 + c * 16), v[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_wait16(v[c]);
          uint32_t vv[2][16];
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld16(tm_row + (uint32_t)(128 + hf * 32 + c * 16), vv[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 o0, o1;
            tc_pack16(v[c], bqk + c * 16, o0, o1);
            st_shared_v4(dstQK + ((((uint32_t)(2 * c)) ^ rsw) << 4), o0.x, o0.y, o0.z, o0.w);
            st_shared_v4(dstQK + ((((uint32_t)(2 * c + 1)) ^ rsw) << 4), o1.x, o1.y, o1.z, o1.w);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld_wait16(vv[c]);
          // V transposed: element (channel cr, token r) -> token block r / 64, row cr, column r % 64
          const float* bv = p.bqkv + 2 * kE + g * 64 + hf * 32;
          const uint32_t vt_tok = sVT + (uint32_t)wi * 8192 + (uint32_t)(r & 7) * 2;
          const uint32_t tu = (uint32_t)(ti >> 3);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int cr = hf * 32 + c * 16 + i;
              const float f = __uint_as_float(vv[c][i]) + __ldg(bv + c * 16 + i);
              st_shared_u16(vt_tok + (uint32_t)cr * 128 + ((tu ^ ((uint32_t)cr & 7)) << 4), __half_as_ushort(__float2half_rn(f)));
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(qkv_drained);
        if (stamp) dbg[4 + g * 6] = clock64() - t_start;

        // ---- softmax of head h = 2 g + hf: this thread owns query row r (token ti of window wi) ----
        const int h = 2 * g + hf;
        float bias[64];
        {
          const float4* bp = reinterpret_cast<const float4*>(p.relbias + ((size_t)h * 64 + ti) * 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 b4 = __ldg(bp + i);
            bias[4 * i] = b4.x; bias[4 * i + 1] = b4.y; bias[4 * i + 2] = b4.z; bias[4 * i + 3] = b4.w;
          }
        }
        const uint32_t tmS = tm_row + (hf ? kTmS1 : kTmS0);
        mbar_wait(&s_full[hf], n_grp & 1);
        tc_fence_after();
        if (stamp) dbg[5 + g * 6] = clock64() - t_start;
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
        float mx = -1e30f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float m = ((mbits >> (j & 7)) & 1u) ? -100.0f : 0.f;
          s[j] = s[j] * p.scale + bias[j] + m;
          mx = fmaxf(mx, s[j]);
        }
        float sum = 0.f;
        const uint32_t prow = sP + (uint32_t)hf * 16384 + (uint32_t)r * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t q[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float e0 = __expf(s[8 * u + 2 * i] - mx), e1 = __expf(s[8 * u + 2 * i + 1] - mx);
            sum += e0; sum += e1;
            q[i] = pack_h2(e0, e1);
          }
          st_shared_v4(prow + ((((uint32_t)u) ^ rsw) << 4), q[0], q[1], q[2], q[3]);
        }
        const float inv = 1.0f / sum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[hf]);
        if (stamp) dbg[6 + g * 6] = clock64() - t_start;

        // ---- O_h = (P V) / rowsum -> fp16, into the projection's A operand (k-block g, 64-byte half hf) ----
        mbar_wait(&o_full[hf], n_grp & 1);
        tc_fence_after();
        if (stamp) dbg[7 + g * 6] = clock64() - t_start;
        {
          uint32_t ov[2][16];
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld16(tmS + (uint32_t)(32 * wi + 16 * c), ov[c]);
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld_wait16(ov[c]);
          const uint32_t orow = sO + (uint32_t)g * 16384 + (uint32_t)r * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t q[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = pack_h2(__uint_as_float(ov[c][2 * i]) * inv, __uint_as_float(ov[c][2 * i + 1]) * inv);
            st_shared_v4(orow + ((((uint32_t)(4 * hf + 2 * c)) ^ rsw) << 4), q[0], q[1], q[2], q[3]);
            st_shared_v4(orow + ((((uint32_t)(4 * hf + 2 * c + 1)) ^ rsw) << 4), q[4], q[5], q[6], q[7]);
          }
        }
        if (stamp) dbg[8 + g * 6] = clock64() - t_start;
        ++n_grp;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_ready);

      // ================= epilogue: y = acc + b + x -> fp16 (one rounding), statistics, full-row stores =================
      const int mypix = sPix[r];
      const bool want_stats = p.sink[0].part != nullptr;
      constexpr int kHalfCols = kE / 2;                       // columns per worker half
      mbar_wait(y_full, n_tile & 1);
      tc_fence_after();
      if (stamp) dbg[24] = clock64() - t_start;
#pragma unroll
      for (int c = 0; c < kHalfCols / 16; ++c) {
        const int col = hf * kHalfCols + c * 16;
        uint32_t v[16];
        tmem_ld16(tm_row + (uint32_t)col, v);
        uint4 x0 = make_uint4(0, 0, 0, 0), x1 = make_uint4(0, 0, 0, 0);
        if (mypix >= 0) {
          const uint4* xp = reinterpret_cast<const uint4*>(p.x + (long long)mypix * p.x_ld + col);
          x0 = xp[0]; x1 = xp[1];
        }
        tmem_ld_wait16(v);
        float f[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bproj + col) + i);
          f[4 * i] = __uint_as_float(v[4 * i]) + b4.x; f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
          f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z; f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
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
        const uint32_t yrow = sXn + (uint32_t)(col >> 6) * 16384 + (uint32_t)r * 128;
        const uint32_t u0 = (uint32_t)((col & 63) >> 3);
        st_shared_v4(yrow + ((u0 ^ rsw) << 4), o0.x, o0.y, o0.z, o0.w);
        st_shared_v4(yrow + (((u0 + 1) ^ rsw) << 4), o1.x, o1.y, o1.z, o1.w);
        if (want_stats) warp_chunk_stats(o0, o1, lane, sStat + ((size_t)quad * kE + col) * 2);
      }
      tc_fence_before();
      named_bar_sync(1, kTcWorkers);
      if (stamp) dbg[25] = clock64() - t_start;
      // full token rows to global
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kTcWorkers;
        const int row = i / kUnits, unit = i - row * kUnits;
        const int pix = sPix[row];
        if (pix >= 0) {
          const uint4 val = ld_shared_v4(sXn + (uint32_t)(unit >> 3) * 16384 + (uint32_t)row * 128 + ((((uint32_t)unit & 7) ^ ((uint32_t)row & 7)) << 4));
          *reinterpret_cast<uint4*>(p.y + (long long)pix * p.y_ld + unit * 8) = val;
        }
      }
      if (want_stats) {
        // merge the two 32-row quadrants of each window (Chan et al., equal counts) and deliver the window's pairs
        for (int idx = tid; idx < 2 * kE; idx += kTcWorkers) {
          const int k = idx / kE, c = idx - k * kE;
          const int w2 = 2 * pair + k;
          if (w2 >= p.total_windows) continue;
          const float m0 = sStat[((size_t)(2 * k) * kE + c) * 2], q0 = sStat[((size_t)(2 * k) * kE + c) * 2 + 1];
          const float m1 = sStat[((size_t)(2 * k + 1) * kE + c) * 2], q1 = sStat[((size_t)(2 * k + 1) * kE + c) * 2 + 1];
          float mm, qq;
          chan_merge_equal(32.f, m0, q0, m1, q1, mm, qq);
          const int n = w2 / nW, slot = w2 % nW;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const GnSink& sk = p.sink[d];
            if (!sk.part) continue;
            float* dst = sk.part + (((size_t)n * nW + slot) * sk.cstride + sk.coff + c) * 2;
            dst[0] = mm; dst[1] = qq;
          }
        }
      }
      named_bar_sync(1, kTcWorkers);                         // staging / scratch / pixel table are free for the next tile
      if (stamp) dbg[26] = clock64() - t_start;
      ++n_tile;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) { tc_fence_after(); tmem_dealloc_dyn(tmem_base, 512u); }
}

#endif  // __CUDACC__
}  // namespace rs
