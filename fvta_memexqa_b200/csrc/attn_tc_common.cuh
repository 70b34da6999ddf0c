// What the tensor-core forward (attn_fwd_tc.cu) and backward (attn_bwd_tc.cu) share: the tile plan and
// the "split" stage that turns an fp32 context row into the two tensor-memory A operands
//   A_hi = the fp32 row as it is (kind::tf32 reads the upper 19 bits: no conversion work)
//   A_lo = fp16(x - tf32(x))    the 13 bits the tensor core drops
// and accumulates the row statistics on the way.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fvta {

constexpr int kTcJQP = 32;        // question positions per tile (N of the MMAs: 64 and 32)
constexpr int kTcTmemCols = 512;
constexpr int kTcMaxStages = 6;
constexpr float kMaskedWeight = 7.888609052210118e-31f;  // 2^-100
constexpr float kTanhScale = 2.885390081777927f;         // 2 * log2(e)

struct TcPlan {
  int TB;           // context positions per tile
  int rows;         // K * TB <= 128
  int K32;          // tf32 k-steps (8 features)  = ceil(D / 8)
  int K16;          // f16 k-steps (16 features)  = ceil(D / 16)
  int C32;          // 32-feature (128-byte) chunks of the tf32 B image
  int C16;          // 64-feature (128-byte) chunks of the f16 B image
  int stages;       // fp32 ring depth
  int stage_bytes;  // fp32 tile, rounded to 128 B
  int tiles;        // tiles per sample
  int slabs;        // CTAs per sample
  int col_ahi, w_ahi, col_alo, w_alo;   // TMEM columns of the two A buffers (buffer b at col + b * w)
  size_t smem;
};

// fixed: bytes of kernel-specific scratch next to the B images and coefficient vectors; stage_extra: bytes
// that come with every ring stage (per-row records of the tile); min_stages: ring depth to aim for.
inline bool tc_plan(const FvtaAttnDesc& d, TcPlan& p, size_t fixed_extra, int stage_extra, int min_stages) {
  if ((d.D & 3) != 0 || d.D > 128 || d.D < 8 || d.JQ > kTcJQP || d.K > 8) return false;
  const bool bounded = (d.simi == 4) || d.add_tanh;
  if (!bounded || d.time_warp_att) return false;
  p.K32 = (d.D + 7) / 8;
  p.K16 = (d.D + 15) / 16;
  p.C32 = (d.D + 31) / 32;
  p.C16 = (d.D + 63) / 64;
  p.w_ahi = 8 * p.K32;
  p.w_alo = 8 * p.K16;
  p.col_ahi = 128;
  p.col_alo = p.col_ahi + 2 * p.w_ahi;
  if (p.col_alo + 2 * p.w_alo > kTcTmemCols) return false;
  const int Dp = round_up(d.D, 4);
  const size_t fixed = (size_t)p.C32 * 8192 + (size_t)p.C16 * 4096 + (size_t)(3 * Dp + 2 * kTcJQP) * 4 + fixed_extra;
  const size_t budget = 227 * 1024;
  int best_tb = 0, best_st = 0;
#ifdef FVTA_TC_TB
  const int tb_max = FVTA_TC_TB;
#else
  const int tb_max = 128 / d.K;
#endif
  for (int TB = tb_max; TB >= 2; --TB) {
    const int sb = round_up(d.K * TB * d.D * 4, 128);
    int st = (int)((budget - fixed) / (size_t)(sb + stage_extra));
    if (st > kTcMaxStages) st = kTcMaxStages;
    if (st >= min_stages) { best_tb = TB; best_st = st; break; }
    if (st >= 2 && best_tb == 0) { best_tb = TB; best_st = st; }   // deepest ring is not reachable: take the widest tile
  }
  if (best_tb == 0) return false;
  p.TB = best_tb;
  p.stages = best_st;
  p.rows = d.K * p.TB;
  p.stage_bytes = round_up(p.rows * d.D * 4, 128);
  p.smem = fixed + (size_t)p.stages * (p.stage_bytes + stage_extra);
  p.tiles = (d.T + p.TB - 1) / p.TB;
#ifdef FVTA_TC_SLABS_CEIL
  const int want = (148 + d.N - 1) / d.N;
#else
  const int want = 148 / d.N;              // one CTA per SM, ONE wave: a second, nearly empty wave costs a whole CTA lifetime
#endif
  p.slabs = want < 1 ? 1 : (want > p.tiles ? p.tiles : want);
  return true;
}

// One split thread's share of a tile row: `NO` 8-feature groups starting at group o0 (compile-time when
// DT > 0: straight-line code, every bound folded).  Loads first, then statistics, residuals and the
// two TMEM stores per pair of groups.
// WITH_G (backward): also hg = gvec . h, the dot with the upstream gradient of g1_all.
template <bool HAS_LIN, bool WITH_G, int DT>
__device__ __forceinline__ void split_row(const float* __restrict__ hrow, const float* __restrict__ coef_s,
                                          const float* __restrict__ gvec_s, int Dp, int D, int o_begin, int o_end,
                                          int K32, uint32_t colA, uint32_t colL, float& zdot, float& quad, float& lin,
                                          float& hg) {
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 zdA = make_float2(0.f, 0.f), qdA = zdA, lnA = zdA, zdB = zdA, qdB = zdA, lnB = zdA, hgA = zdA;
  auto stats4 = [&](int d0, const float4& v, float2& zd, float2& qd, float2& ln) {
    const float4 cu = *reinterpret_cast<const float4*>(coef_s + d0);
    const float4 cq = *reinterpret_cast<const float4*>(coef_s + 2 * Dp + d0);
    const float2 a01 = make_float2(v.x, v.y), a23 = make_float2(v.z, v.w);
    const float2 s01 = __fmul2_rn(a01, a01), s23 = __fmul2_rn(a23, a23);
    zd = __ffma2_rn(make_float2(cu.x, cu.y), s01, zd);
    zd = __ffma2_rn(make_float2(cu.z, cu.w), s23, zd);
    qd = __ffma2_rn(make_float2(cq.x, cq.y), s01, qd);
    qd = __ffma2_rn(make_float2(cq.z, cq.w), s23, qd);
    if (HAS_LIN) {
      const float4 cl = *reinterpret_cast<const float4*>(coef_s + Dp + d0);
      ln = __ffma2_rn(make_float2(cl.x, cl.y), a01, ln);
      ln = __ffma2_rn(make_float2(cl.z, cl.w), a23, ln);
    }
    if (WITH_G) {
      const float4 cg = *reinterpret_cast<const float4*>(gvec_s + d0);
      hgA = __ffma2_rn(make_float2(cg.x, cg.y), a01, hgA);
      hgA = __ffma2_rn(make_float2(cg.z, cg.w), a23, hgA);
    }
  };
  // residual the tensor core drops, as two packed fp16: exact subtraction, fp16 rounding 2^-11 of it
  auto lo2 = [&](float a, float b) -> uint32_t {
    const float2 rr = __ffma2_rn(make_float2(tf32_part(a), tf32_part(b)), make_float2(-1.f, -1.f), make_float2(a, b));
    const __half2 p = __floats2half2_rn(rr.x, rr.y);
    return *reinterpret_cast<const uint32_t*>(&p);
  };
  auto pair = [&](int o, bool two) {
    const int d0 = o * 8;
    const float4 v0 = d0 < D ? *reinterpret_cast<const float4*>(hrow + d0) : zero4;
    const float4 v1 = d0 + 4 < D ? *reinterpret_cast<const float4*>(hrow + d0 + 4) : zero4;
    const float4 v2 = (two && d0 + 8 < D) ? *reinterpret_cast<const float4*>(hrow + d0 + 8) : zero4;
    const float4 v3 = (two && d0 + 12 < D) ? *reinterpret_cast<const float4*>(hrow + d0 + 12) : zero4;
    if (d0 < D) stats4(d0, v0, zdA, qdA, lnA);
    if (d0 + 4 < D) stats4(d0 + 4, v1, zdB, qdB, lnB);
    if (two && d0 + 8 < D) stats4(d0 + 8, v2, zdA, qdA, lnA);
    if (two && d0 + 12 < D) stats4(d0 + 12, v3, zdB, qdB, lnB);
    uint32_t hi[16], lo[8];
    hi[0] = __float_as_uint(v0.x); hi[1] = __float_as_uint(v0.y); hi[2] = __float_as_uint(v0.z); hi[3] = __float_as_uint(v0.w);
    hi[4] = __float_as_uint(v1.x); hi[5] = __float_as_uint(v1.y); hi[6] = __float_as_uint(v1.z); hi[7] = __float_as_uint(v1.w);
    hi[8] = __float_as_uint(v2.x); hi[9] = __float_as_uint(v2.y); hi[10] = __float_as_uint(v2.z); hi[11] = __float_as_uint(v2.w);
    hi[12] = __float_as_uint(v3.x); hi[13] = __float_as_uint(v3.y); hi[14] = __float_as_uint(v3.z); hi[15] = __float_as_uint(v3.w);
    lo[0] = lo2(v0.x, v0.y); lo[1] = lo2(v0.z, v0.w); lo[2] = lo2(v1.x, v1.y); lo[3] = lo2(v1.z, v1.w);
    lo[4] = lo2(v2.x, v2.y); lo[5] = lo2(v2.z, v2.w); lo[6] = lo2(v3.x, v3.y); lo[7] = lo2(v3.z, v3.w);
    // every condition below is warp-uniform (o and the half are): the stores are warp-collective
    if (two && o + 1 < K32) tmem_st16(colA + 8 * o, hi);
    else if (o < K32) tmem_st8(colA + 8 * o, hi);
    if (two) tmem_st8(colL + 4 * o, lo);
    else tmem_st4(colL + 4 * o, lo);
  };
  if (DT > 0) {
    constexpr int kOT = 2 * ((DT + 15) / 16), kHalf = kOT / 2;     // groups per thread half
#pragma unroll
    for (int j = 0; j < kHalf; j += 2) pair(o_begin + j, j + 1 < kHalf);
  } else {
#pragma unroll 1
    for (int o = o_begin; o < o_end; o += 2) pair(o, o + 1 < o_end);
  }
  const float2 zd2 = __fadd2_rn(zdA, zdB), qd2 = __fadd2_rn(qdA, qdB), ln2 = __fadd2_rn(lnA, lnB);
  zdot = zd2.x + zd2.y; quad = qd2.x + qd2.y; lin = ln2.x + ln2.y;
  hg = hgA.x + hgA.y;
}

}  // namespace fvta
