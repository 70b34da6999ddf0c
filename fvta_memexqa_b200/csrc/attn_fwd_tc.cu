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
#include "tc_ptx.cuh"
#include "attn_tc_common.cuh"

// A/B hook: FVTA_FWD_SPIN = 1 makes the waits on the accumulator loop (split -> MMA -> softmax -> MMA)
// spin instead of sleeping between polls, 2 the weighted-sum role's too.  Measured at C3-max: no
// difference (97.5 / 97.8 / 98.4 us) -- the wake-up delay of nanosleep is not what sets the tile period.
#ifndef FVTA_FWD_SPIN
#define FVTA_FWD_SPIN 0
#endif
#if FVTA_FWD_SPIN >= 1
#define TC_WAIT_HOT(bar, par) mbar_wait(bar, par)
#else
#define TC_WAIT_HOT(bar, par) mbar_wait_relaxed(bar, par)
#endif
#if FVTA_FWD_SPIN >= 2
#define TC_WAIT_WARM(bar, par) mbar_wait(bar, par)
#else
#define TC_WAIT_WARM(bar, par) mbar_wait_relaxed(bar, par)
#endif

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
constexpr int kTcStageExtra = 512 + 3 * 1024;   // per ring stage: wg [128] + row statistics [3][2][128]
#ifndef FVTA_TC_MIN_STAGES
#define FVTA_TC_MIN_STAGES 4
#endif

static bool tc_make_plan(const FvtaAttnDesc& d, TcPlan& p) {
  return tc_plan(d, p, /*fixed=*/5120 /*red*/ + 1024 /*combine scratch*/ + 512 /*barriers*/, kTcStageExtra, FVTA_TC_MIN_STAGES);
}

bool attn_tc_covers(const FvtaAttnDesc& d) {
  TcPlan p;
  return tc_make_plan(d, p);
}

// ---- the fused kernel: grid (slabs, N) ----------------------------------------------------------

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
  uint64_t* zq_ready = bars + 3 * kTcMaxStages;       // [S]  split -> softmax (row statistics).  Per STAGE, not per
                                                      //      TMEM buffer: a softmax group that lags cannot see this
                                                      //      barrier run two phases ahead (parity waits would alias)
  uint64_t* a_ready = bars + 4 * kTcMaxStages;        // [2]  split -> MMA (operands in tensor memory)
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
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1); mbar_init(&empty[s], ws_warps * 32); mbar_init(&wg_ready[s], 128); mbar_init(&zq_ready[s], 256);
    }
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
        TC_WAIT_HOT(&a_ready[b2], (i >> 1) & 1);
        if (i >= 2) TC_WAIT_HOT(&ep_done[b2], ((i - 2) >> 1) & 1);   // softmax(i-2) has drained accumulator b2
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
      float hg_unused;
      split_row<HAS_LIN, false, DT>(hrow, coef_s, nullptr, Dp, D, o_begin, o_end, K32, colA, colL, zdot, quad, lin, hg_unused);
      float* zq = zq_s + s * 768 + h * 128 + r;
      zq[0] = zdot; zq[256] = quad;
      if (HAS_LIN) zq[512] = lin;
      mbar_arrive(&zq_ready[s]);
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
#ifdef FVTA_AB_AREADY
      mbar_wait_relaxed(&a_ready[p], (i >> 1) & 1);
#else
      mbar_wait_relaxed(&zq_ready[s], (i / S) & 1);     // the split warps' row statistics are visible
#endif
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
      TC_WAIT_HOT(&acc_full[p], (i >> 1) & 1);
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
      TC_WAIT_WARM(&wg_ready[s], (i / S) & 1);
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
// Runs the tile kernel (+ fused combine) over a workspace whose folded operands are in place.
int attn_fwd_tc_launch(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall, const uint8_t* hall_mask,
                       const uint8_t* hq_mask, float* ws, float* g1_all, float* a_logits, float* c_col, float* stats,
                       cudaStream_t st) {
  TcPlan pl;
  if (!tc_make_plan(d, pl)) return FVTA_E_SHAPE;
  if (pl.slabs > L.npart) return FVTA_E_SHAPE;
  CUtensorMap tmap;
  if (!tc_make_tile_map(&tmap, hall, d.N, d.K, d.T, d.D, pl.TB)) {
    set_last_error("cuTensorMapEncodeTiled failed", cudaErrorInvalidValue);
    return FVTA_E_LAUNCH;
  }
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
