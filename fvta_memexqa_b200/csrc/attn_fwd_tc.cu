// Focal visual-text attention forward on the 5th-generation tensor cores (sm_100a only).
//
// Same math and outputs as attn_fwd.cu (model_v2.py:953-1009 time warp, :301-341 window, :210-298
// attention_3d, :23-48 softmax/softsel, utils.py:210-213 exp_mask), ONE pass over hall, ONE launch
// after the operand folding (the combine of the per-CTA partials runs in the last CTA of each sample).
//
// The question x context contraction P[t,j] = h[t,:] . qeff[j,:] -- what makes a CUDA-core version
// compute-bound at JQ/2 = 16 flop/B -- runs as tcgen05.mma with BOTH the accumulator and the A
// operand (the context rows) in tensor memory:
//
//   A_hi = the fp32 row as it is: kind::tf32 reads the upper 19 bits (no conversion work at all)
//   A_lo = fp16(x - tf32(x))      the 13 bits the tensor core drops, as a second operand
//   P    = A_hi.B_hi + A_hi.B_lo   (kind::tf32, N = 64: [B_hi ; B_lo] = tf32(q), q - tf32(q))
//        + A_lo.B_hi16             (kind::f16,  N = 32, accumulated into the B_lo columns)
//   ~21 mantissa bits per product, fp32 accumulation: the 1e-4 parity bar holds where single-pass
//   tf32 / bf16 would not (S = g.P + ... with |g| up to 7).
//
// Pipeline (one persistent CTA per SM, 7 warpgroups, every hand-over is an mbarrier; each role is
// bound by its own dependent-instruction latency, so every role gets as many warps as it takes to
// keep its time per tile under the HBM time of the tile; registers are moved between the roles
// with setmaxnreg):
//   warp 24       TMA      ring of fp32 context tiles [K modalities x TB positions x D]
//                          (cp.async.bulk.tensor.3d, complete_tx; rows past T are zero-filled)
//   warps 0-7     split    two threads per context row: row statistics (u.h^2, w_quad.h^2, w_lin.h),
//                          A_lo, and tcgen05.st of both operands into TMEM (lane = row)
//   warp 25       MMA      one elected lane: 13 tf32 + 7 f16 k-steps per tile at D = 100
//   warps 8-11,   softmax  two warpgroups on alternate tiles, one thread per row: gate, tcgen05.ld of
//         12-15            its 2 x 32 accumulator columns, logits (MUFU tanh), exp_mask as a min
//                          with a per-column cap, row max, softmax weight (fixed reference maximum
//                          1: the logits are bounded)
//   warps 16-23   wsum     weighted sum  o[k,:] += w.g.h  over the still resident fp32 tile (thread =
//                          one modality x one float4 of features), then the stage goes back to TMA
// The softmax over T needs no online rescaling: this kernel is used when the logits are bounded by
// 1 (tanh or cosine).  Masked rows get the weight 2^-100 instead of exp(-1e30 - 1) = 0: absorbed
// exactly by any unmasked weight, and it reproduces TF's uniform softmax over a fully masked row.
#include "common.cuh"
#include "attn_tile.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace fvta {

#ifdef FVTA_TC_TRACE
constexpr int kTraceTiles = 24;
#ifndef FVTA_TC_TRACE_T0
#define FVTA_TC_TRACE_T0 0
#endif
__device__ long long g_tc_trace[5][kTraceTiles][8];
#define TC_TRACE(role, i, ev)                                                                     \
  do { if (blockIdx.x == 0 && blockIdx.y == 0 && (i) >= FVTA_TC_TRACE_T0 && (i) < FVTA_TC_TRACE_T0 + kTraceTiles - 1 && \
           (threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 128 || threadIdx.x == kWarpSplit0 * 32 || threadIdx.x == kWarpSoft0 * 32 || threadIdx.x == kWarpSoft0 * 32 + 128)) \
         g_tc_trace[role][(i) - FVTA_TC_TRACE_T0][ev] = clock64(); } while (0)
__device__ __forceinline__ long long tc_globaltimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// kernel-level stamps of CTA (0,0), thread 0: slot 0 entry, 1 pipeline start, 2 pipeline end, 3 exit (cycles; +4: ns)
#define TC_STAMP(slot)                                                                            \
  do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == kWarpSplit0 * 32) {                \
         g_tc_trace[0][kTraceTiles - 1][slot] = clock64(); g_tc_trace[0][kTraceTiles - 1][4 + (slot)] = tc_globaltimer(); } } while (0)
#else
#define TC_TRACE(role, i, ev) do {} while (0)
#define TC_STAMP(slot) do {} while (0)
#endif

// warp 0 TMA, 1 MMA, 2-3 idle, 4-11 wsum, 12-19 softmax (2 x 4), 20-27 split.  The scheduler favours high warp
// ids: the split warps, which set the pace, come last.
constexpr int kTcThreads = 896;
#ifdef FVTA_TC_SPLIT_TOP
constexpr int kWarpTma = 0, kWarpMma = 1, kWarpWsum0 = 4, kWarpSoft0 = 12, kWarpSplit0 = 20;
#else
constexpr int kWarpTma = 0, kWarpMma = 1, kWarpWsum0 = 4, kWarpSplit0 = 12, kWarpSoft0 = 20;
#endif
constexpr int kRegSoftmax = 112, kRegWsum = 48, kRegIssue = 40;   // split keeps the launch value (72)
constexpr int kTcJQP = 32;        // question positions per tile (N of the MMAs: 64 and 32)
constexpr int kTcTmemCols = 512;
constexpr int kTcMaxStages = 6;
constexpr float kMaskedWeight = 7.888609052210118e-31f;  // 2^-100
constexpr float kTanhScale = 2.885390081777927f;         // 2 * log2(e)
constexpr int kTcStageExtra = 512 + 3 * 1024;   // per ring stage: wg [128] + row statistics [3][2][128]
#ifndef FVTA_TC_MIN_STAGES
#define FVTA_TC_MIN_STAGES 4
#endif

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

__host__ __device__ inline size_t tc_fixed_smem(const TcPlan& p, int Dp) {
  return (size_t)p.C32 * 8192 + (size_t)p.C16 * 4096 + (size_t)(3 * Dp + 2 * kTcJQP) * 4 + 5120 /*red*/ +
         1024 /*combine scratch*/ + 256 /*barriers*/;
}

static bool tc_make_plan(const FvtaAttnDesc& d, TcPlan& p) {
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
  const size_t fixed = tc_fixed_smem(p, round_up(d.D, 4));
  const size_t budget = 227 * 1024;
  int best_tb = 0, best_st = 0;
#ifdef FVTA_TC_TB
  const int tb_max = FVTA_TC_TB;
#else
  const int tb_max = 128 / d.K;
#endif
  for (int TB = tb_max; TB >= 2; --TB) {
    const int sb = round_up(d.K * TB * d.D * 4, 128);
    int st = (int)((budget - fixed) / (size_t)(sb + kTcStageExtra));
    if (st > kTcMaxStages) st = kTcMaxStages;
    if (st >= FVTA_TC_MIN_STAGES) { best_tb = TB; best_st = st; break; }
    if (st >= 2 && best_tb == 0) { best_tb = TB; best_st = st; }   // deepest ring is not reachable: take the widest tile
  }
  if (best_tb == 0) return false;
  p.TB = best_tb;
  p.stages = best_st;
  p.rows = d.K * p.TB;
  p.stage_bytes = round_up(p.rows * d.D * 4, 128);
  p.smem = fixed + (size_t)p.stages * (p.stage_bytes + kTcStageExtra);
  p.tiles = (d.T + p.TB - 1) / p.TB;
  const int want = (148 + d.N - 1) / d.N;  // one CTA per SM
  p.slabs = want < 1 ? 1 : (want > p.tiles ? p.tiles : want);
  return true;
}

bool attn_tc_covers(const FvtaAttnDesc& d) {
  TcPlan p;
  return tc_make_plan(d, p);
}

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the warp in hardware until the phase completes or an implementation time limit
// passes, so the loop around it rarely iterates and waiting warps do not compete for issue slots.
// (A suspend-time hint makes ptxas emit a NANOSLEEP.SYNCS loop that wakes on every barrier event of
// the CTA: measured 33 % of all issued instructions.)  Far more iterations than any real wait can
// need mean a pipeline bug: trap (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity))
    if (++spins > (1u << 24)) __trap();
}
// for the roles with slack (everything but the split warps, which set the pace): sleep between polls so
// that waiting warps leave the issue slots to the working ones (plain polling was 45 % of all issued
// instructions); the wake-up delay only adds pipeline latency, which the ring absorbs
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one lane of a converged warp; the branch it guards keeps warp-uniform values in uniform registers
// (a `lane == 0` branch makes ptxas re-broadcast every tcgen05.mma operand through an R2UR loop:
// ~110 cycles per MMA instead of ~10, tools/umma_probe.cu)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// D[tmem] (+)= A[tmem] . B[smem]^T, M = 128
__device__ __forceinline__ void umma_ts_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}

// K-major, 128B-swizzled operand: 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D = fp32, M = 128, A and B K-major; fmt 0 = fp16, 2 = tf32
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}

// MUFU.EX2 / MUFU.RCP; tanh(x) = 1 - 2 / (1 + 2^{2 log2(e) x}) built from them has |error| < 1e-6
// and never exceeds 1 in magnitude
__device__ __forceinline__ float exp2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// what kind::tf32 keeps of an fp32 operand (tools/umma_probe.cu: the tensor core truncates)
__device__ __forceinline__ float tf32_part(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- the fused kernel: grid (slabs, N) ----------------------------------------------------------
template <int NREG> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NREG)); }
template <int NREG> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NREG)); }

// One split thread's share of a tile row: `NO` 8-feature groups starting at group o0 (compile-time when
// DT > 0: straight-line code, every bound folded).  Loads first, then statistics, residuals and the
// two TMEM stores per pair of groups.
template <bool HAS_LIN, int DT>
__device__ __forceinline__ void split_row(const float* __restrict__ hrow, const float* __restrict__ coef_s, int Dp,
                                          int D, int o_begin, int o_end, int K32, uint32_t colA, uint32_t colL,
                                          float& zdot, float& quad, float& lin) {
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 zdA = make_float2(0.f, 0.f), qdA = zdA, lnA = zdA, zdB = zdA, qdB = zdA, lnB = zdA;
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
}

// DT: compile-time D (100, the reference hidden size 50) or 0 = any supported D at run time.
template <bool HAS_LIN, int DT>
__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc(
    const __grid_constant__ CUtensorMap tmap, FvtaAttnDesc d, TcPlan pl, const uint8_t* __restrict__ hall_mask,
    const uint8_t* __restrict__ hq_mask, const float* __restrict__ coef, const uint8_t* __restrict__ qB32,
    const uint8_t* __restrict__ qB16, const float* __restrict__ col, const float* __restrict__ zconst,
    float* __restrict__ a_logits, float* __restrict__ c_col, float* __restrict__ part_stat,
    float* __restrict__ part_o, int* __restrict__ counters, float* __restrict__ g1_all, float* __restrict__ stats,
    int Dp, int JQPg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  TC_STAMP(0);
  const int S = pl.stages;
  const int D = DT > 0 ? DT : d.D;
  uint8_t* b32_s = smem;                                   // tf32 [B_hi ; B_lo], swizzled
  uint8_t* b16_s = b32_s + pl.C32 * 8192;                  // f16 B_hi, swizzled
  uint8_t* ring = b16_s + pl.C16 * 4096;                   // fp32 tiles
  float* coef_s = reinterpret_cast<float*>(ring + (size_t)S * pl.stage_bytes);  // [3][Dp]
  float* col_s = coef_s + 3 * Dp;        // [32]  per-question-word term + bias
  float* cap_s = col_s + kTcJQP;         // [32]  +inf for a valid question word, -1e30 masked, -inf padding
  float* red_s = cap_s + kTcJQP;         // [2][5][128] slab statistics of the two softmax warpgroups
  float* cmb_s = red_s + 1280;           // [256] combine scratch
  float* wg_s = cmb_s + 256;             // [S][128]     softmax weight * gate
  float* zq_s = wg_s + S * 128;          // [S][3][2][128] half-row statistics: zdot, quad, lin
  uint64_t* bars = reinterpret_cast<uint64_t*>(zq_s + S * 768);
  uint64_t* full = bars;                              // [S]  TMA -> split
  uint64_t* empty = bars + kTcMaxStages;              // [S]  wsum -> TMA
  uint64_t* wg_ready = bars + 2 * kTcMaxStages;       // [S]  softmax -> wsum
  uint64_t* a_ready = bars + 3 * kTcMaxStages;        // [2]  split -> MMA (operands), softmax (row statistics)
  uint64_t* acc_full = a_ready + 2;                   // [2]  MMA -> softmax (accumulator), split (A buffers free)
  uint64_t* ep_done = a_ready + 4;                    // [2]  softmax -> MMA (accumulator drained)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 6);
  int* last_flag = reinterpret_cast<int*>(tmem_ptr + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x, n = blockIdx.y;
  const int tile_begin = (int)(((long long)slab * pl.tiles) / pl.slabs);
  const int tile_end = (int)(((long long)(slab + 1) * pl.tiles) / pl.slabs);
  const int ntile = tile_end - tile_begin;
  const int TB = pl.TB, rows = pl.rows;
  const int K32 = DT > 0 ? (DT + 7) / 8 : pl.K32, K16 = DT > 0 ? (DT + 15) / 16 : pl.K16;
  const int ws_warps = (d.K * (D >> 2) + 31) >> 5;      // weighted-sum warps with work (<= 8)

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ws_warps * 32); mbar_init(&wg_ready[s], 128); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(&a_ready[b2], 256); mbar_init(&acc_full[b2], 1); mbar_init(&ep_done[b2], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(kTcTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // per-sample constants
  {
    const uint4* s32 = reinterpret_cast<const uint4*>(qB32 + (size_t)n * pl.C32 * 8192);
    const uint4* s16 = reinterpret_cast<const uint4*>(qB16 + (size_t)n * pl.C16 * 4096);
    uint4* d32 = reinterpret_cast<uint4*>(b32_s);
    uint4* d16 = reinterpret_cast<uint4*>(b16_s);
    for (int i = threadIdx.x; i < pl.C32 * 512; i += blockDim.x) d32[i] = s32[i];
    for (int i = threadIdx.x; i < pl.C16 * 256; i += blockDim.x) d16[i] = s16[i];
    for (int i = threadIdx.x; i < 3 * Dp; i += blockDim.x) coef_s[i] = coef[i];
    // with tanh on, the logit is 1 - 2 / (1 + 2^(kTanhScale * S)): the scale is folded into the row
    // coefficients and this column term instead of costing a multiply per logit
    const float col_scale = (d.add_tanh && d.simi != 4) ? kTanhScale : 1.f;
    for (int j = threadIdx.x; j < kTcJQP; j += blockDim.x) {
      col_s[j] = j < JQPg ? col_scale * col[(size_t)n * JQPg + j] : 0.f;
      const bool qv = !d.has_mask || (j < d.JQ && hq_mask[(size_t)n * d.JQ + j]);
      cap_s[j] = j >= d.JQ ? -INFINITY : (qv ? INFINITY : FVTA_VERY_NEGATIVE);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  TC_STAMP(1);

  if (warp < kWarpWsum0) {
    reg_dec<kRegIssue>();
    if (warp == kWarpTma) {
      // ===== TMA producer (whole warp loops, one elected lane issues) =====
      for (int i = 0; i < ntile; ++i) {
        const int s = i % S;
        if (i >= S) mbar_wait_relaxed(&empty[s], ((i / S) - 1) & 1);
        TC_TRACE(0, i, 0);
        if (elect_one()) {
          mbar_expect_tx(&full[s], (uint32_t)(rows * D * 4));
          tma_load_3d(ring + (size_t)s * pl.stage_bytes, &tmap, &full[s], 0, (tile_begin + i) * TB, n * d.K);
        }
        __syncwarp();
      }
    } else if (warp == kWarpMma) {
      // ===== MMA issuer =====
      //   acc[:, 0:64]  = A_hi . [B_hi ; B_lo]^T   (kind::tf32)
      //   acc[:, 32:64] += A_lo . B_hi16^T          (kind::f16)       P = acc[:, 0:32] + acc[:, 32:64]
      // The whole warp runs the loop (uniform control flow keeps every operand in uniform registers:
      // a `lane == 0` branch costs ~110 cycles per MMA in R2UR broadcast loops); one elected lane issues.
      const uint32_t id32 = umma_idesc(2, 64), id16 = umma_idesc(0, 32);
      const uint32_t b32_u = smem_u32(b32_s), b16_u = smem_u32(b16_s);
      for (int i = 0; i < ntile; ++i) {
        const int b2 = i & 1;
        const uint32_t acc = tmem_base + (uint32_t)(b2 * 64);
        const uint32_t a_hi = tmem_base + (uint32_t)(pl.col_ahi + b2 * pl.w_ahi);
        const uint32_t a_lo = tmem_base + (uint32_t)(pl.col_alo + b2 * pl.w_alo);
        mbar_wait_relaxed(&a_ready[b2], (i >> 1) & 1);
        if (i >= 2) mbar_wait_relaxed(&ep_done[b2], ((i - 2) >> 1) & 1);   // softmax(i-2) has drained accumulator b2
        TC_TRACE(1, i, 0);
        tc_fence_after();
        if (elect_one()) {
          if (DT > 0) {
#pragma unroll
            for (int ks = 0; ks < (DT + 7) / 8; ++ks)
              umma_ts_tf32(acc, a_hi + 8 * ks, umma_desc_sw128(b32_u + (uint32_t)((ks >> 2) * 8192 + (ks & 3) * 32)), id32, ks > 0);
#pragma unroll
            for (int ks = 0; ks < (DT + 15) / 16; ++ks)
              umma_ts_f16(acc + 32, a_lo + 8 * ks, umma_desc_sw128(b16_u + (uint32_t)((ks >> 2) * 4096 + (ks & 3) * 32)), id16, 1u);
          } else {
#pragma unroll 4
            for (int ks = 0; ks < K32; ++ks)
              umma_ts_tf32(acc, a_hi + 8 * ks, umma_desc_sw128(b32_u + (uint32_t)((ks >> 2) * 8192 + (ks & 3) * 32)), id32, ks > 0);
#pragma unroll 4
            for (int ks = 0; ks < K16; ++ks)
              umma_ts_f16(acc + 32, a_lo + 8 * ks, umma_desc_sw128(b16_u + (uint32_t)((ks >> 2) * 4096 + (ks & 3) * 32)), id16, 1u);
          }
          umma_commit(&acc_full[b2]);
        }
        __syncwarp();
        TC_TRACE(1, i, 1);
      }
    }
  } else if (warp >= kWarpSplit0 && warp < kWarpSplit0 + 8) {
    // ===== split warps: threads (r, 0) and (r, 1) share context row r = k*TB + tt =====
    const int r = (threadIdx.x - kWarpSplit0 * 32) & 127, h = (threadIdx.x - kWarpSplit0 * 32) >> 7;
    const int roff = (r < rows ? r : 0) * D;        // idle lanes read row 0: their TMEM rows are never used
    const int OT = 2 * K16;                         // 8-feature groups ("octs") covering both operands
    const int o_begin = h ? OT / 2 : 0, o_end = h ? OT : OT / 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < ntile; ++i) {
      const int s = i % S, b2 = i & 1;
      TC_TRACE(2, i, 0);
      mbar_wait(&full[s], (i / S) & 1);
      if (i >= 2) {                                   // MMA(i-2) has consumed the A buffers b2
        mbar_wait(&acc_full[b2], ((i - 2) >> 1) & 1);
        tc_fence_after();
      }
      TC_TRACE(2, i, 1);
      const float* hrow = reinterpret_cast<const float*>(ring + (size_t)s * pl.stage_bytes) + roff;
      const uint32_t colA = lane_base + (uint32_t)(pl.col_ahi + b2 * pl.w_ahi);
      const uint32_t colL = lane_base + (uint32_t)(pl.col_alo + b2 * pl.w_alo);
      float zdot, quad, lin;
      split_row<HAS_LIN, DT>(hrow, coef_s, Dp, D, o_begin, o_end, K32, colA, colL, zdot, quad, lin);
      float* zq = zq_s + s * 768 + h * 128 + r;
      zq[0] = zdot; zq[256] = quad;
      if (HAS_LIN) zq[512] = lin;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a_ready[b2]);
      TC_TRACE(2, i, 2);
    }
  } else if (warp >= kWarpSoft0 && warp < kWarpSoft0 + 8) {
    // ===== softmax warpgroups: group p handles tiles i = p, p + 2, ...; thread r owns accumulator row r =====
    reg_inc<kRegSoftmax>();
    const int p = (warp - kWarpSoft0) >> 2;
    const int r = (threadIdx.x - kWarpSoft0 * 32) & 127;
    const bool active = r < rows;
    const int k = active ? r / TB : 0, tt = active ? r - k * TB : 0;
    const bool tanh_on = d.add_tanh && d.simi != 4;
    const bool cosine = d.simi == 4;
    const float zc = d.use_time_warp ? zconst[n] : 0.f;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p * 64);
    float l_acc = 0.f, smax = -INFINITY, ties = 0.f, am_t = 3.0e38f, am_j = 0.f;
    const float4* col4 = reinterpret_cast<const float4*>(col_s);
    const float4* cap4 = reinterpret_cast<const float4*>(cap_s);
    const uint8_t* mrow = hall_mask + ((size_t)n * d.K + k) * d.T;
    // the mask byte of the NEXT tile is fetched while this one is processed (a global load per tile
    // on the critical path costs ~1000 cycles)
    uint8_t mh_next = 1;
    {
      const int t0 = (tile_begin + p) * TB + tt;
      if (d.has_mask && active && p < ntile && t0 < d.T) mh_next = mrow[t0];
    }
    for (int i = p; i < ntile; i += 2) {
      const int s = i % S;
      const int t = (tile_begin + i) * TB + tt;
      const bool in_range = active && t < d.T;
      const bool mh = in_range && mh_next != 0;
      {
        const int t2 = t + 2 * TB;
        mh_next = 1;
        if (d.has_mask && active && i + 2 < ntile && t2 < d.T) mh_next = mrow[t2];
      }
      TC_TRACE(3, i, 0);
      mbar_wait_relaxed(&a_ready[p], (i >> 1) & 1);     // the split warps' row statistics are visible
      const float* zq = zq_s + s * 768;
      float g = 1.f;
      if (d.use_time_warp) {
        float z = zc;
        for (int kk = 0; kk < d.K; ++kk) z += zq[kk * TB + tt] + zq[128 + kk * TB + tt];
        const int tc = t < d.T ? t : d.T - 1;
        const float tz = 1.f - 2.f * rcp_approx(exp2f_approx(z * 2.885390081777927f) + 1.f);   // tanh(z)
        g = tz * window_count(tc, d.T, d.warp_type, d.window);
        if (c_col != nullptr && k == 0 && in_range) c_col[(size_t)n * d.T + t] = tz;
      }
      const int rr = active ? r : 0;
      const float quad = zq[256 + rr] + zq[384 + rr];
      const float lin = HAS_LIN ? zq[512 + rr] + zq[640 + rr] : 0.f;
      float c1, c0v;
      row_coefs(d, g, lin, quad, c1, c0v);
      const float p_c0 = cosine ? 0.f : c0v;
      const float s_scale = tanh_on ? kTanhScale : 1.f;
      mbar_wait_relaxed(&acc_full[p], (i >> 1) & 1);
      TC_TRACE(3, i, 1);
      tc_fence_after();
      float a[32];
      float m = -INFINITY;
      const float2 c1c1 = make_float2(c1 * s_scale, c1 * s_scale), c0c0 = make_float2(p_c0 * s_scale, p_c0 * s_scale);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t x0[16], x1[16];
        tmem_ld16_issue(taddr + 16 * hh, x0);
        tmem_ld16_issue(taddr + 32 + 16 * hh, x1);
        tmem_ld_wait();
        if (hh == 1) {           // both accumulator halves are in registers: the MMA warp may refill the buffer
          tc_fence_before();
          mbar_arrive(&ep_done[p]);
        }
        // branch-free logits.  |tanh| <= 1, so th + (-1e30) == -1e30 exactly: exp_mask is a min
        // with the per-column cap
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 cq = col4[(16 * hh + j) >> 2], kq = cap4[(16 * hh + j) >> 2];
          const float colr0 = (j & 2) ? cq.z : cq.x, colr1 = (j & 2) ? cq.w : cq.y;
          const float capr0 = (j & 2) ? kq.z : kq.x, capr1 = (j & 2) ? kq.w : kq.y;
          const float2 P = __fadd2_rn(make_float2(__uint_as_float(x0[j]), __uint_as_float(x0[j + 1])),
                                      make_float2(__uint_as_float(x1[j]), __uint_as_float(x1[j + 1])));
          float2 Sv = __fadd2_rn(__ffma2_rn(c1c1, P, c0c0), make_float2(colr0, colr1));
          if (tanh_on) {  // tanh(x) = 1 - 2 / (1 + e^{2x}), packed; Sv already carries the factor 2 log2(e)
            const float2 den = __fadd2_rn(make_float2(exp2f_approx(Sv.x), exp2f_approx(Sv.y)), make_float2(1.f, 1.f));
            Sv = __ffma2_rn(make_float2(rcp_approx(den.x), rcp_approx(den.y)), make_float2(-2.f, -2.f), make_float2(1.f, 1.f));
          }
          const float a0 = fminf(Sv.x, capr0), a1 = fminf(Sv.y, capr1);
          a[16 * hh + j] = a0; a[16 * hh + j + 1] = a1;
          m = fmaxf(m, fmaxf(a0, a1));
        }
      }
      if (__any_sync(0xffffffffu, !mh) && !mh) {   // masked context row: every logit is -1e30 (padding columns stay -inf)
        m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) { a[j] = fminf(cap_s[j], FVTA_VERY_NEGATIVE); m = fmaxf(m, a[j]); }
      }
      TC_TRACE(3, i, 2);
      // softmax weight first: the weighted-sum warps are the next stage of the pipeline
      float w = 0.f;
      if (in_range) {
        w = (m <= -1e29f) ? kMaskedWeight : exp2f_approx((m - 1.f) * 1.4426950408889634f);
        l_acc += w;
      }
      wg_s[s * 128 + r] = w * g;
      mbar_arrive(&wg_ready[s]);
      if (a_logits != nullptr && in_range) {
        float* arow = a_logits + (((size_t)n * d.K + k) * d.T + t) * d.JQ;
        if ((d.JQ & 3) == 0) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            if (j4 * 4 < d.JQ)
              *reinterpret_cast<float4*>(arow + j4 * 4) = make_float4(a[j4 * 4], a[j4 * 4 + 1], a[j4 * 4 + 2], a[j4 * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < d.JQ) arow[j] = a[j];
        }
      }
      // a new (or tied) running maximum is rare after the first tiles: keep the bookkeeping behind a
      // warp-uniform branch so it costs nothing when no lane needs it
      const bool newmax = in_range && m >= smax;
      if (__any_sync(0xffffffffu, newmax)) {
        if (newmax) {
          if (m > smax) { smax = m; ties = 0.f; am_t = 3.0e38f; }
          int cnt = 0, jm = -1;
#pragma unroll
          for (int j = 31; j >= 0; --j) {
            const bool hit = a[j] == m;
            cnt += hit ? 1 : 0;
            jm = hit ? j : jm;
          }
          ties += (float)cnt;
          if (jm >= 0 && (float)t < am_t) { am_t = (float)t; am_j = (float)jm; }
        }
      }
      TC_TRACE(3, i, 3);
    }
    // ---- slab statistics: per-modality reductions over the 2 x TB threads owning its rows ----
    float* rp = red_s + p * 640;
    rp[r] = l_acc; rp[128 + r] = smax; rp[256 + r] = ties; rp[384 + r] = am_t; rp[512 + r] = am_j;
    named_bar_sync(1, 256);
    if (p == 0 && r < d.K) {
      float l = 0.f, sm = -INFINITY, ti = 0.f, at = 3.0e38f, aj = 0.f;
      for (int pp = 0; pp < 2; ++pp)
        for (int t2 = 0; t2 < TB; ++t2) {
          const float* q = red_s + pp * 640 + r * TB + t2;
          l += q[0];
          const float v = q[128];
          if (v > sm) { sm = v; ti = 0.f; at = 3.0e38f; }
          if (v == sm) {
            ti += q[256];
            if (q[384] < at) { at = q[384]; aj = q[512]; }
          }
        }
      float* ps = part_stat + (((size_t)n * pl.slabs + slab) * d.K + r) * kPartW;
      ps[0] = 1.f; ps[1] = l; ps[2] = sm; ps[3] = ti; ps[4] = at; ps[5] = aj; ps[6] = 0.f; ps[7] = 0.f;
      __threadfence();
    }
  } else {
    // ===== wsum warps: thread u owns (modality k, features 4*d4 .. 4*d4+3): o += sum_tt (w*g)[k,tt] * h[k,tt,:] =====
    reg_dec<kRegWsum>();
    const int u = threadIdx.x - kWarpWsum0 * 32;
    const int D4 = D >> 2;
    const bool ws_on = u < d.K * D4;
    const int ws_k = ws_on ? u / D4 : 0, ws_d4 = ws_on ? u - ws_k * D4 : 0;
    const int ws_off = ws_k * TB * D + ws_d4 * 4;
    float2 oa = make_float2(0.f, 0.f), ob = oa;
    for (int i = 0; i < (warp - kWarpWsum0 < ws_warps ? ntile : 0); ++i) {      // warps without work skip the pipeline
      const int s = i % S;
      TC_TRACE(4, i, 0);
      mbar_wait_relaxed(&wg_ready[s], (i / S) & 1);
      TC_TRACE(4, i, 1);
      if (ws_on) {
        const float* hp = reinterpret_cast<const float*>(ring + (size_t)s * pl.stage_bytes) + ws_off;
        const float* wp = wg_s + s * 128 + ws_k * TB;
        int t2 = 0;
#pragma unroll 2
        for (; t2 + 4 <= TB; t2 += 4) {
          const float4 h0 = *reinterpret_cast<const float4*>(hp);
          const float4 h1 = *reinterpret_cast<const float4*>(hp + D);
          const float4 h2 = *reinterpret_cast<const float4*>(hp + 2 * D);
          const float4 h3 = *reinterpret_cast<const float4*>(hp + 3 * D);
          const float w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
          oa = __ffma2_rn(make_float2(w0, w0), make_float2(h0.x, h0.y), oa);
          ob = __ffma2_rn(make_float2(w0, w0), make_float2(h0.z, h0.w), ob);
          oa = __ffma2_rn(make_float2(w1, w1), make_float2(h1.x, h1.y), oa);
          ob = __ffma2_rn(make_float2(w1, w1), make_float2(h1.z, h1.w), ob);
          oa = __ffma2_rn(make_float2(w2, w2), make_float2(h2.x, h2.y), oa);
          ob = __ffma2_rn(make_float2(w2, w2), make_float2(h2.z, h2.w), ob);
          oa = __ffma2_rn(make_float2(w3, w3), make_float2(h3.x, h3.y), oa);
          ob = __ffma2_rn(make_float2(w3, w3), make_float2(h3.z, h3.w), ob);
          hp += 4 * D;
          wp += 4;
        }
        for (; t2 < TB; ++t2) {
          const float4 h0 = *reinterpret_cast<const float4*>(hp);
          const float w0 = wp[0];
          oa = __ffma2_rn(make_float2(w0, w0), make_float2(h0.x, h0.y), oa);
          ob = __ffma2_rn(make_float2(w0, w0), make_float2(h0.z, h0.w), ob);
          hp += D;
          wp += 1;
        }
      }
      mbar_arrive(&empty[s]);
      TC_TRACE(4, i, 2);
    }
    // one thread holds the whole slab sum of its (modality, feature group): no cross-thread reduction
    if (ws_on) {
      *reinterpret_cast<float4*>(part_o + (((size_t)n * pl.slabs + slab) * d.K + ws_k) * Dp + ws_d4 * 4) =
          make_float4(oa.x, oa.y, ob.x, ob.y);
      __threadfence();
    }
  }
  tc_fence_before();
  __syncthreads();
  TC_STAMP(2);
  if (warp == kWarpMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
  }

  // ===== combine: the last CTA of sample n merges the slab partials (model_v2.py:266-278:
  // softmax over T from the partial sums, softmax over the K modalities, g1_all) =====
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(&counters[n], 1);
    *last_flag = (ticket == pl.slabs - 1) ? 1 : 0;
  }
  __syncthreads();
  TC_STAMP(3);
  if (*last_flag == 0) return;
  __threadfence();
  {
    const int K = d.K, nsl = pl.slabs;
    float* L_s = cmb_s;          // [8]
    float* s_s = cmb_s + 8;      // [8]
    float* beta_s = cmb_s + 16;  // [8]
    const float* ps = part_stat + (size_t)n * nsl * K * kPartW;
    if (warp < K) {
      const int k = warp;
      float L = 0.f, sm = -INFINITY;
      for (int i = lane; i < nsl; i += 32) {
        const float* e = ps + ((size_t)i * K + k) * kPartW;
        L += __ldcg(e + 1);
        sm = fmaxf(sm, __ldcg(e + 2));
      }
      // fixed-order tree: the same value on every run
      L = warp_sum(L);
      sm = warp_max(sm);
      float ti = 0.f;
      int first = 0x7fffffff;
      for (int i = lane; i < nsl; i += 32) {
        const float* e = ps + ((size_t)i * K + k) * kPartW;
        if (__ldcg(e + 2) == sm) { ti += __ldcg(e + 3); first = min(first, i); }
      }
      ti = warp_sum(ti);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      if (lane == 0) {
        const float* ef = ps + ((size_t)(first < nsl ? first : 0) * K + k) * kPartW;
        float* st = stats + ((size_t)n * K + k) * kStatW;
        st[0] = 1.f; st[1] = L; st[2] = sm; st[3] = ti; st[5] = __ldcg(ef + 4); st[6] = __ldcg(ef + 5); st[7] = 0.f;
        L_s[k] = L; s_s[k] = sm;
      }
    }
    __syncthreads();
    if (threadIdx.x < K) {
      float smax = -INFINITY, den = 0.f;
      for (int kk = 0; kk < K; ++kk) smax = fmaxf(smax, s_s[kk]);
      for (int kk = 0; kk < K; ++kk) den += expf(s_s[kk] - smax);
      const float beta = expf(s_s[threadIdx.x] - smax) / den;
      beta_s[threadIdx.x] = beta;
      stats[((size_t)n * K + threadIdx.x) * kStatW + 4] = beta;
    }
    __syncthreads();
    // r_k = sum_slabs o / L_k; staged over the (now idle) ring so that g1_all adds K terms in order
    float* r_s = reinterpret_cast<float*>(ring);     // [K][Dp]
    const float* po = part_o + (size_t)n * nsl * K * Dp;
    for (int idx = threadIdx.x; idx < K * D; idx += blockDim.x) {
      const int k = idx / D, c = idx - k * D;
      float o = 0.f;
      for (int i = 0; i < nsl; ++i) o += __ldcg(po + ((size_t)i * K + k) * Dp + c);
      const float rv = o / L_s[k];
      r_s[k * Dp + c] = rv;
      stats[attn_stats_r_offset(d) + ((size_t)n * K + k) * D + c] = rv;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float gsum = 0.f;
      for (int k = 0; k < K; ++k) gsum = fmaf(beta_s[k], r_s[k * Dp + c], gsum);
      g1_all[(size_t)n * D + c] = gsum;
    }
    if (threadIdx.x == 0) counters[n] = 0;   // ready for the next call over this workspace
  }
}

#ifdef FVTA_TC_TRACE
extern "C" int fvta_tc_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 5 * kTraceTiles * 8);
}
#endif

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;   // resolved once; the same pointer for every caller
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Runs the tile kernel (+ fused combine) over a workspace whose folded operands are in place.
int attn_fwd_tc_launch(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall, const uint8_t* hall_mask,
                       const uint8_t* hq_mask, float* ws, float* g1_all, float* a_logits, float* c_col, float* stats,
                       cudaStream_t st) {
  TcPlan pl;
  if (!tc_make_plan(d, pl)) return FVTA_E_SHAPE;
  if (pl.slabs > L.npart) return FVTA_E_SHAPE;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_last_error("cuTensorMapEncodeTiled not found", cudaErrorUnknown); return FVTA_E_LAUNCH; }
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {(cuuint64_t)d.D, (cuuint64_t)d.T, (cuuint64_t)d.N * d.K};
  cuuint64_t gstr[2] = {(cuuint64_t)d.D * 4, (cuuint64_t)d.T * d.D * 4};
  cuuint32_t box[3] = {(cuuint32_t)d.D, (cuuint32_t)pl.TB, (cuuint32_t)d.K};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(hall), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed", cudaErrorInvalidValue); return FVTA_E_LAUNCH; }
  const bool has_lin = (d.simi == 1 || d.simi == 3);   // w_lin is identically zero otherwise
  auto kern = has_lin ? attn_fwd_tc<true, 0> : (d.D == 100 ? attn_fwd_tc<false, 100> : attn_fwd_tc<false, 0>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) { set_last_error("attn_fwd_tc smem attr", e); return FVTA_E_LAUNCH; }
  dim3 grid(pl.slabs, d.N);
  kern<<<grid, kTcThreads, pl.smem, st>>>(tmap, d, pl, hall_mask, hq_mask, ws + L.coef,
                                          reinterpret_cast<const uint8_t*>(ws + L.qB32),
                                          reinterpret_cast<const uint8_t*>(ws + L.qB16), ws + L.col, ws + L.zconst,
                                          a_logits, c_col, ws + L.part_stat, ws + L.part_o,
                                          reinterpret_cast<int*>(ws + L.counter), g1_all, stats, L.Dp, L.JQP);
  return check_launch("attn_fwd_tc");
}

}  // namespace fvta
