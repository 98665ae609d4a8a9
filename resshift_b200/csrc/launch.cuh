// Host-side launch helpers: TMA descriptor encoding and one launcher per kernel family.
#pragma once

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "norm_act.cuh"
#include "window_attn.cuh"

namespace rs {

// ---- driver entry point for cuTensorMapEncodeTiled (the .so does not link libcuda) --------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// NHWC fp16 view descriptor used by the host code.
struct View {
  __half* ptr = nullptr;   // resolved at bind time
  int tens = -1;           // owning workspace tensor
  long long off = 0;       // element offset inside the tensor
  int N = 0, H = 0, W = 0, C = 0, ld = 0;
  long long sW() const { return ld; }
  long long sH() const { return (long long)W * ld; }
  long long sN() const { return (long long)H * W * ld; }
};

// 4-D activation map {C, W, H, N} with explicit element strides; box {64, bw, bh, bn}, 128B swizzle.
inline int encode_act_map(CUtensorMap* m, const __half* base, int C, int W, int H, int N, long long sW,
                          long long sH, long long sN, int bw, int bh, int bn) {
  PFN_encodeTiled enc = get_encode_tiled();
  RS_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)kConvBK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  RS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "activation base must be 16-byte aligned");
  RS_CHECK(sW % 8 == 0 && sH % 8 == 0 && sN % 8 == 0, "activation strides must be multiples of 16 bytes");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed with CUresult " + std::to_string((int)r));
  return 0;
}
// 2-D weight map {Ktot, Cout}, box {64, BN}.
inline int encode_weight_map(CUtensorMap* m, const __half* base, int Ktot, int Cout, int BN) {
  PFN_encodeTiled enc = get_encode_tiled();
  RS_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kConvBK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  RS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && Ktot % 8 == 0, "weight matrix must be 16-byte aligned");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with CUresult " + std::to_string((int)r));
  return 0;
}

inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}
inline bool env_is(const char* name, const char* val) {
  const char* v = std::getenv(name);
  return v && std::strcmp(v, val) == 0;
}

inline int pow2_floor_div(int x, int cap) {   // largest power of two dividing x, capped
  int p = 1;
  while (p * 2 <= cap && x % (p * 2) == 0) p *= 2;
  return p;
}

// Host description of one conv / linear layer instance.
struct ConvDesc {
  View in;                 // input view (for stride 2: the full-resolution input)
  int ksize = 1, stride = 1;
  const __half* wt = nullptr;   // [Cout][taps][ipad]
  int ipad = 0;
  const float* bias = nullptr;
  int Cout = 0;
  View out;                // NHWC fp16 output view (ptr may be null when out_f32 is used)
  bool has_out = true;
  View res; bool has_res = false;
  float* out_f32 = nullptr;
  int act = ACT_NONE;
  int bn_override = 0;
  // filled by finalize()
  ConvParams prm;
  ConvSimtSrc simt;
  int grid = 0; size_t smem = 0;
};

inline int conv_finalize(ConvDesc& d) {
  ConvParams& p = d.prm;
  std::memset(&p, 0, sizeof(p));
  const int Hin = d.in.H, Win = d.in.W;
  RS_CHECK(d.ksize == 1 || d.ksize == 3, "kernel size must be 1 or 3");
  RS_CHECK(d.stride == 1 || (d.stride == 2 && d.ksize == 3 && Hin % 2 == 0 && Win % 2 == 0), "unsupported stride");
  const int Hout = Hin / d.stride, Wout = Win / d.stride, N = d.in.N;
  p.Hout = Hout; p.Wout = Wout; p.Nimg = N; p.Cout = d.Cout;
  p.num_taps = d.ksize * d.ksize;
  p.kchunks = (d.in.C + kConvBK - 1) / kConvBK;
  p.w_tap_stride = d.ipad;
  RS_CHECK(d.ipad % 8 == 0 && d.ipad >= d.in.C, "weight channel padding");
  // pixel box
  p.bw = pow2_floor_div(Wout, kConvBM);
  p.bh = pow2_floor_div(Hout, kConvBM / p.bw);
  p.bn = kConvBM / (p.bw * p.bh);
  p.tiles_w = Wout / p.bw; p.tiles_h = Hout / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // channel tile
  const int cout16 = (d.Cout + 15) / 16 * 16;
  int BN = 0;
  const int forced = d.bn_override ? d.bn_override : env_int("RS_CONV_BN", 0);
  if (forced) {
    BN = std::min(forced, std::min(cout16, 256));
  } else {
    int smallest = 0;
    for (int cand = std::min(cout16, 256); cand >= 16; cand -= 16) {
      if (cout16 % cand) continue;
      smallest = cand;
      if (!BN && (long long)m_tiles * (cout16 / cand) >= 120) BN = cand;
      if (cand <= 32) break;
    }
    if (!BN) BN = smallest;
  }
  RS_CHECK(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid BN");
  p.BN = BN; p.n_tiles = (cout16 + BN - 1) / BN;
  int cols = 32; while (cols < BN) cols *= 2;
  p.tmem_cols = cols;
  const int stage_bytes = kConvBM * kConvBK * 2 + BN * kConvBK * 2;
  int occ = env_int("RS_CONV_OCC", 2);
  int stages = env_int("RS_CONV_STAGES", 0);
  if (!stages) {
    int budget = (occ >= 2 ? 110 : 220) * 1024 - 2048;
    stages = budget / stage_bytes;
    if (stages < 3) stages = (220 * 1024 - 2048) / stage_bytes;
    stages = std::max(2, std::min(stages, 8));
  }
  p.stages = stages;
  d.smem = (size_t)stages * stage_bytes + 1024 + 256;
  RS_CHECK(d.smem <= 227 * 1024, "shared memory budget exceeded");
  d.grid = m_tiles * p.n_tiles;
  // taps
  if (d.stride == 1) {
    int t = 0;
    for (int ky = 0; ky < d.ksize; ++ky)
      for (int kx = 0; kx < d.ksize; ++kx, ++t) {
        p.tap_src[t] = 0; p.tap_dh[t] = ky - d.ksize / 2; p.tap_dw[t] = kx - d.ksize / 2;
      }
  } else {
    int t = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx, ++t) {
        const int hp = (ky == 1) ? 0 : 1, wp = (kx == 1) ? 0 : 1;
        p.tap_src[t] = hp * 2 + wp; p.tap_dh[t] = (ky == 0) ? -1 : 0; p.tap_dw[t] = (kx == 0) ? -1 : 0;
      }
  }
  // epilogue
  p.bias = d.bias; p.act = d.act;
  if (d.has_res) {
    RS_CHECK(d.res.H == Hout && d.res.W == Wout && d.res.N == N && d.res.C >= d.Cout, "residual geometry");
    p.residual = d.res.ptr; p.res_sN = d.res.sN(); p.res_sH = d.res.sH(); p.res_sW = d.res.sW();
  }
  if (d.has_out) {
    RS_CHECK(d.out.H == Hout && d.out.W == Wout && d.out.N == N, "output geometry");
    p.out = d.out.ptr; p.out_sN = d.out.sN(); p.out_sH = d.out.sH(); p.out_sW = d.out.sW();
    RS_CHECK(d.out.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out.ptr) & 15) == 0, "output alignment");
  }
  p.out_f32_nchw = d.out_f32;
  // tensor maps + SIMT mirrors
  ConvSimtSrc& s = d.simt;
  std::memset(&s, 0, sizeof(s));
  s.C = d.in.C; s.wt = d.wt;
  const int nsrc = d.stride == 1 ? 1 : 4;
  for (int i = 0; i < kMaxSrc; ++i) {
    const int j = i < nsrc ? i : 0;
    const int hp = d.stride == 2 ? (j >> 1) : 0, wp = d.stride == 2 ? (j & 1) : 0;
    const __half* base = d.in.ptr + (long long)hp * d.in.sH() + (long long)wp * d.in.sW();
    const long long sW = d.in.sW() * d.stride, sH = d.in.sH() * d.stride, sN = d.in.sN();
    s.ptr[i] = base; s.sN[i] = sN; s.sH[i] = sH; s.sW[i] = sW; s.H[i] = Hout; s.W[i] = Wout;
    if (!env_is("RS_CONV_IMPL", "simt")) {
      int rc = encode_act_map(&p.tmA[i], base, d.in.C, Wout, Hout, N, sW, sH, sN, p.bw, p.bh, p.bn);
      if (rc) return rc;
    }
  }
  if (!env_is("RS_CONV_IMPL", "simt")) {
    int rc = encode_weight_map(&p.tmB, d.wt, p.num_taps * d.ipad, d.Cout, BN);
    if (rc) return rc;
  }
  return 0;
}

inline int conv_init() {   // once per process, outside any stream capture
  static bool attr_set = false;
  if (!attr_set) {
    RS_CUDA_OK(cudaFuncSetAttribute(conv_gemm_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  return 0;
}

inline int conv_launch(const ConvDesc& d, cudaStream_t st) {
  if (env_is("RS_CONV_IMPL", "simt")) {
    const long long npix = (long long)d.prm.Nimg * d.prm.Hout * d.prm.Wout;
    const int warps = 8;
    conv_simt_kernel<<<(unsigned)((npix + warps - 1) / warps), warps * 32, 0, st>>>(d.prm, d.simt);
  } else {
    conv_gemm_sm100_kernel<<<d.grid, kConvThreads, d.smem, st>>>(d.prm);
  }
  RS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- GroupNorm -------------------------------------------------------------------------------
struct GnDesc {
  View in, out;
  const float* gamma = nullptr; const float* beta = nullptr;
  const float* film = nullptr; long long film_sN = 0;   // resolved per launch for FiLM layers
  int film_off = -1;      // offset of this layer's [2C] slice inside an embedding row, or -1
  int silu = 0;
  float* sums = nullptr;  // [N][C][2]
};

inline int gn_launch(const GnDesc& g, cudaStream_t st) {
  const int C = g.in.C, HW = g.in.H * g.in.W, N = g.in.N;
  RS_CHECK(C % 32 == 0 && C % 8 == 0 && C <= 2048, "GroupNorm channel count");
  RS_CHECK(g.in.ld % 8 == 0 && g.out.ld % 8 == 0, "GroupNorm view alignment");
  // enough CTAs to fill the machine, at least 32 rows each
  int chunks = std::max(1, std::min((HW + 31) / 32, (148 * 4 + N - 1) / N));
  int rows = (HW + chunks - 1) / chunks;
  chunks = (HW + rows - 1) / rows;
  GnStatsParams sp{g.in.ptr, g.in.sN(), g.in.ld, C, HW, N, g.sums, rows};
  gn_stats_kernel<<<dim3(chunks, N), 256, 2 * C * sizeof(float), st>>>(sp);
  RS_CUDA_OK(cudaGetLastError());
  GnApplyParams ap{g.in.ptr, g.in.sN(), g.in.ld, g.out.ptr, g.out.sN(), g.out.ld, C, HW, N, g.sums,
                   g.gamma, g.beta, g.film, g.film_sN, g.silu, rows, 1e-5f};
  gn_apply_kernel<<<dim3(chunks, N), 256, 2 * C * sizeof(float), st>>>(ap);
  RS_CUDA_OK(cudaGetLastError());
  return 0;
}

inline int attn_launch(const View& qkv, const View& out, const float* bias, int heads, int E, int shift,
                       cudaStream_t st) {
  RS_CHECK(qkv.H % 8 == 0 && qkv.W % 8 == 0, "window attention needs H, W multiples of 8");
  RS_CHECK(E == heads * 32, "window attention kernel is specialised for head_dim 32");
  WinAttnParams p{qkv.ptr, qkv.ld, out.ptr, out.ld, bias, qkv.N, qkv.H, qkv.W, heads, E, shift,
                  0.17677669529663687f, env_is("RS_ATTN_IMPL", "simt") ? 1 : 0};
  window_attn_kernel<<<dim3(qkv.N * (qkv.H / 8) * (qkv.W / 8), heads), 128, 0, st>>>(p);
  RS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rs
