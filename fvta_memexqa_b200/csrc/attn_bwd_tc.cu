// Focal visual-text attention backward on the 5th-generation tensor cores (sm_100a only).
//
// Gradient of attn_fwd_tc.cu's forward (model_v2.py:953-1009 time warp, :210-298 attention_3d with both
// softmaxes and both reduce_max) w.r.t. the context tensor, the folded question
// operand and the coefficient vectors, in ONE pass over hall and ONE pass over d_hall.
//
// Two observations shape it:
//  1. The logits have to be recomputed (saving them costs 4.N.K.T.JQ bytes each way).  That is the same
//     contraction as the forward, so it runs the same way: the fp32 tile arrives by TMA, the "split"
//     warps put the rows into tensor memory as A_hi (raw fp32 = tf32) / A_lo (fp16 residual) and
//     tcgen05.mma forms P against the question images -- with the forward's exact instruction
//     sequence, so the recomputed logits equal the forward's bit for bit and equality with the saved
//     maxima is meaningful.
//  2. The gradient of the logits is SPARSE: d a[t, j] is non-zero only where a[t, j] is the row maximum
//     (softmax over T of max_j) or the sequence maximum (softmax over K of max_{t,j}); saturated tanh
//     entries have a zero derivative.  So  d h = dS . qeff  and  d qeff = dS^T . h  are not [rows x JQ x D]
//     contractions but ~one scaled row each: d h[t,:] += v . qeff[j*,:],  d qeff[j*,:] += v . h[t,:].
//     (The CUDA-core kernel attn_bwd.cu spends three dense JQ-wide contractions here.)
//
// Pipeline (same warp roles and mbarrier hand-overs as the forward):
//   TMA load -> split (+ h.g_hat) -> MMA -> row warpgroups (two, alternate tiles: logits, maxima, the few
//   non-zero dS of the row, gate backward, per-row record; rows are bucketed by their column j*)
//   -> update warps: (A) d qeff[j,:] += sum over bucket j of v.h  (thread = (j, float4), registers),
//                    (B) d h in place over the tile + the three coefficient-vector sums,
//                    TMA store of the tile into d_hall, stage back to the loader.
#include "common.cuh"
#include "attn_tile.cuh"
#include "tc_ptx.cuh"
#include "attn_tc_common.cuh"

#ifndef FVTA_BWD_SPIN
#define FVTA_BWD_SPIN 0
#endif
#if FVTA_BWD_SPIN >= 1
#define BT_WAIT_HOT(bar, par) mbar_wait(bar, par)
#else
#define BT_WAIT_HOT(bar, par) mbar_wait_relaxed(bar, par)
#endif
#if FVTA_BWD_SPIN >= 2
#define BT_WAIT_WARM(bar, par) mbar_wait(bar, par)
#else
#define BT_WAIT_WARM(bar, par) mbar_wait_relaxed(bar, par)
#endif

namespace fvta {

constexpr int kBtThreads = 896;
// warp 0 TMA, 1 MMA, 2-3 idle, 4-11 update, 12-19 split, 20-27 row warpgroups
constexpr int kBwTma = 0, kBwMma = 1, kBwUpd0 = 4, kBwSplit0 = 12, kBwRow0 = 20;
constexpr int kBtRegRow = 80, kBtRegUpd = 80, kBtRegIssue = 40, kBtRegLaunch = 72;   // split keeps the launch value
// setmaxnreg.inc waits until the pool holds the registers: a split that does not fit the file deadlocks
static_assert(128 * kBtRegIssue + 256 * (kBtRegLaunch + kBtRegRow + kBtRegUpd) <= 65536, "register file");
static_assert(kBtThreads * kBtRegLaunch <= 65536, "launch registers");
constexpr int kBtEnd = 0xffff;         // end of a column's entry list
constexpr int kBtStageExtra = 4 * 1024;   // per stage: half-row statistics [4][2][128]

#ifdef FVTA_TC_TRACE
constexpr int kBtTraceTiles = 24;
__device__ long long g_bt_trace[5][kBtTraceTiles][8];
__device__ int g_bt_count[8];
__device__ long long g_bt_iter[32];
#define BT_TRACE(role, i, ev)                                                                       \
  do { if (blockIdx.x == 0 && blockIdx.y == 0 && (i) < kBtTraceTiles && (threadIdx.x & 127) == 0)     \
         g_bt_trace[role][(i)][ev] = clock64(); } while (0)
#else
#define BT_TRACE(role, i, ev) do {} while (0)
#endif

// fixed scratch: gvec [128], d gate exchange [2 groups][2 phases][128], dcol [32], per-k statistics [8][12],
// per-row records [2 parities][128][8], entry lists (heads [2][32], links [2][3 x 128] u16), d zconst, barriers
constexpr int kBtFixedExtra = 512 + 2048 + 128 + 384 + 8192 + 256 + 1536 + 16 + 512;

static bool bt_make_plan(const FvtaAttnDesc& d, TcPlan& p) {
  return tc_plan(d, p, kBtFixedExtra, kBtStageExtra, 4);
}

bool attn_bwd_tc_covers(const FvtaAttnDesc& d) {
  TcPlan p;
  return bt_make_plan(d, p) && p.stages >= 2;
}

template <bool HAS_LIN, int DT>
__global__ void __launch_bounds__(kBtThreads, 1) attn_bwd_tc(
    const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out, FvtaAttnDesc d,
    TcPlan pl, const uint8_t* __restrict__ hall_mask, const uint8_t* __restrict__ hq_mask,
    const float* __restrict__ coef, const uint8_t* __restrict__ qB32, const uint8_t* __restrict__ qB16,
    const float* __restrict__ col, const float* __restrict__ zconst, const float* __restrict__ stats,
    const float* __restrict__ drow, const float* __restrict__ dg1, float* __restrict__ dqeffT,
    float* __restrict__ dcol, float* __restrict__ dcoef, float* __restrict__ dzc, int Dp, int JQPg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int S = pl.stages;
  const int D = DT > 0 ? DT : d.D;
  uint8_t* b32_s = smem;
  uint8_t* b16_s = b32_s + pl.C32 * 8192;
  uint8_t* ring = b16_s + pl.C16 * 4096;
  float* coef_s = reinterpret_cast<float*>(ring + (size_t)S * pl.stage_bytes);   // [3][Dp]
  float* col_s = coef_s + 3 * Dp;                 // [32] (prescaled like the forward's)
  float* cap_s = col_s + kTcJQP;                  // [32]
  float* gvec_s = cap_s + kTcJQP;                 // [128] upstream gradient d g1_all[n, :]
  float* dgx_s = gvec_s + 128;                    // [2 groups][2 phases][128] per-row d gate of a tile
  float* dcol_s = dgx_s + 512;                    // [32]
  float* kst_s = dcol_s + 32;                     // [8][12] per-modality statistics
  float* zq_s = kst_s + 96;                       // [S][4][2][128] half-row statistics: zdot, quad, lin, hg
  float4* rec_s = reinterpret_cast<float4*>(zq_s + S * 1024);   // [2][128][2] per-row records, by tile parity
  int* head_s = reinterpret_cast<int*>(rec_s + 2 * 256);        // [2][32] newest entry of column j
  uint16_t* next_s = reinterpret_cast<uint16_t*>(head_s + 64);  // [2][3 x 128] entry -> older entry of its column
  float* dzc_s = reinterpret_cast<float*>(next_s + 2 * 384);    // [1] (+3 pad)
  uint64_t* bars = reinterpret_cast<uint64_t*>(dzc_s + 4);
  uint64_t* full = bars;                              // [S]  TMA -> split
  uint64_t* empty = bars + kTcMaxStages;              // [S]  update -> TMA (tile stored, stage free)
  uint64_t* rec_ready = bars + 2 * kTcMaxStages;      // [S]  row -> update
  uint64_t* zq_ready = bars + 3 * kTcMaxStages;       // [S]  split -> row (row statistics; per stage: no parity aliasing
                                                      //      when a row group lags two tiles behind the split warps)
  uint64_t* a_ready = bars + 4 * kTcMaxStages;        // [2]  split -> MMA
  uint64_t* acc_full = a_ready + 2;                   // [2]  MMA -> row, split
  uint64_t* ep_done = a_ready + 4;                    // [2]  row -> MMA
  uint64_t* bk_free = a_ready + 6;                    // [2]  update -> row: buckets of this parity consumed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x, n = blockIdx.y;
  const int tile_begin = (int)(((long long)slab * pl.tiles) / pl.slabs);
  const int tile_end = (int)(((long long)(slab + 1) * pl.tiles) / pl.slabs);
  const int ntile = tile_end - tile_begin;
  const int TB = pl.TB, rows = pl.rows;
  const int K32 = DT > 0 ? (DT + 7) / 8 : pl.K32, K16 = DT > 0 ? (DT + 15) / 16 : pl.K16;
  const int D4 = D >> 2;
  const bool tanh_on = d.add_tanh && d.simi != 4;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&rec_ready[s], 128); mbar_init(&zq_ready[s], 256);
    }
    for (int b2 = 0; b2 < 2; ++b2) {
      mbar_init(&a_ready[b2], 256); mbar_init(&acc_full[b2], 1); mbar_init(&ep_done[b2], 128); mbar_init(&bk_free[b2], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kBwMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(kTcTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4* s32 = reinterpret_cast<const uint4*>(qB32 + (size_t)n * pl.C32 * 8192);
    const uint4* s16 = reinterpret_cast<const uint4*>(qB16 + (size_t)n * pl.C16 * 4096);
    uint4* d32 = reinterpret_cast<uint4*>(b32_s);
    uint4* d16 = reinterpret_cast<uint4*>(b16_s);
    for (int i = threadIdx.x; i < pl.C32 * 512; i += blockDim.x) d32[i] = s32[i];
    for (int i = threadIdx.x; i < pl.C16 * 256; i += blockDim.x) d16[i] = s16[i];
    for (int i = threadIdx.x; i < 3 * Dp; i += blockDim.x) coef_s[i] = coef[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) gvec_s[i] = i < D ? dg1[(size_t)n * D + i] : 0.f;
    const float col_scale = tanh_on ? kTanhScale : 1.f;
    for (int j = threadIdx.x; j < kTcJQP; j += blockDim.x) {
      col_s[j] = j < JQPg ? col_scale * col[(size_t)n * JQPg + j] : 0.f;
      const bool qv = !d.has_mask || (j < d.JQ && hq_mask[(size_t)n * d.JQ + j]);
      cap_s[j] = j >= d.JQ ? -INFINITY : (qv ? INFINITY : FVTA_VERY_NEGATIVE);
      dcol_s[j] = 0.f;
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) head_s[i] = kBtEnd;
    if (threadIdx.x == 0) dzc_s[0] = 0.f;
    if (threadIdx.x < d.K) {
      const int k = threadIdx.x;
      const float* st = stats + ((size_t)n * d.K + k) * kStatW;
      const float* dr = drow + ((size_t)n * d.K + k) * 4;
      float* o = kst_s + k * 12;
      o[0] = st[0]; o[1] = st[1]; o[2] = st[2]; o[3] = st[3]; o[4] = st[5]; o[5] = st[6];
      o[6] = dr[0]; o[7] = dr[1]; o[8] = dr[2];
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kBwUpd0) {
    reg_dec<kBtRegIssue>();
    if (warp == kBwTma) {
      for (int i = 0; i < ntile; ++i) {
        const int s = i % S;
        if (i >= S) mbar_wait_relaxed(&empty[s], ((i / S) - 1) & 1);
        BT_TRACE(0, i, 0);
        if (elect_one()) {
          mbar_expect_tx(&full[s], (uint32_t)(rows * D * 4));
          tma_load_3d(ring + (size_t)s * pl.stage_bytes, &tmap_in, &full[s], 0, (tile_begin + i) * TB, n * d.K);
        }
        __syncwarp();
      }
    } else if (warp == kBwMma) {
      const uint32_t id32 = umma_idesc(2, 64), id16 = umma_idesc(0, 32);
      const uint32_t b32_u = smem_u32(b32_s), b16_u = smem_u32(b16_s);
      for (int i = 0; i < ntile; ++i) {
        const int b2 = i & 1;
        const uint32_t acc = tmem_base + (uint32_t)(b2 * 64);
        const uint32_t a_hi = tmem_base + (uint32_t)(pl.col_ahi + b2 * pl.w_ahi);
        const uint32_t a_lo = tmem_base + (uint32_t)(pl.col_alo + b2 * pl.w_alo);
        BT_WAIT_HOT(&a_ready[b2], (i >> 1) & 1);
        BT_TRACE(1, i, 0);
        if (i >= 2) BT_WAIT_HOT(&ep_done[b2], ((i - 2) >> 1) & 1);
        BT_TRACE(1, i, 1);
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
      }
    }
  } else if (warp >= kBwSplit0 && warp < kBwSplit0 + 8) {
    // ===== split warps (as in the forward, plus the row's dot with the upstream gradient) =====
    const int r = (threadIdx.x - kBwSplit0 * 32) & 127, h = (threadIdx.x - kBwSplit0 * 32) >> 7;
    const int roff = (r < rows ? r : 0) * D;
    const int OT = 2 * K16;
    const int o_begin = h ? OT / 2 : 0, o_end = h ? OT : OT / 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < ntile; ++i) {
      const int s = i % S, b2 = i & 1;
      // every role of this kernel sleeps between polls: a spinning role starves the others of issue slots
      // (measured: the update warps, lowest in the scheduler's order, ran 4x slower next to spinning split warps)
      BT_TRACE(2, i, 0);
      mbar_wait_relaxed(&full[s], (i / S) & 1);
      BT_TRACE(2, i, 1);
      if (i >= 2) {
        mbar_wait_relaxed(&acc_full[b2], ((i - 2) >> 1) & 1);
        tc_fence_after();
      }
      BT_TRACE(2, i, 2);
      const float* hrow = reinterpret_cast<const float*>(ring + (size_t)s * pl.stage_bytes) + roff;
      const uint32_t colA = lane_base + (uint32_t)(pl.col_ahi + b2 * pl.w_ahi);
      const uint32_t colL = lane_base + (uint32_t)(pl.col_alo + b2 * pl.w_alo);
      float zdot, quad, lin, hg;
      split_row<HAS_LIN, true, DT>(hrow, coef_s, gvec_s, Dp, D, o_begin, o_end, K32, colA, colL, zdot, quad, lin, hg);
      float* zq = zq_s + s * 1024 + h * 128 + r;
      zq[0] = zdot; zq[256] = quad; zq[768] = hg;
      if (HAS_LIN) zq[512] = lin;
      mbar_arrive(&zq_ready[s]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a_ready[b2]);
      BT_TRACE(2, i, 3);
    }
  } else if (warp >= kBwRow0) {
    // ===== row warpgroups: group p handles tiles i = p, p + 2, ...; thread r owns context row r =====
    reg_inc<kBtRegRow>();
    const int p = (warp - kBwRow0) >> 2;
    const int r = (threadIdx.x - kBwRow0 * 32) & 127;
    const bool active = r < rows;
    const int k = active ? r / TB : 0, tt = active ? r - k * TB : 0;
    const bool cosine = d.simi == 4;
    const float zc = d.use_time_warp ? zconst[n] : 0.f;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p * 64);
    const float4* col4 = reinterpret_cast<const float4*>(col_s);
    const float4* cap4 = reinterpret_cast<const float4*>(cap_s);
    const uint8_t* mrow = hall_mask + ((size_t)n * d.K + k) * d.T;
    // per-modality statistics of the forward and the upstream gradients derived from them
    const float* ks = kst_s + k * 12;
    const float Mk = ks[0], sk = ks[2], tiesk = ks[3];
    const float rLk = ks[1] > 0.f ? 1.f / ks[1] : 0.f;
    const int amax_t = (int)ks[4], amax_j = (int)ks[5];
    const float dsk = ks[6], rdr = ks[7], beta = ks[8];
    const float ds_each = dsk / tiesk;
    const float s_scale = tanh_on ? kTanhScale : 1.f;
    int* head = head_s + p * 32;
    uint16_t* nxt = next_s + p * 384;
    float4* rec = rec_s + p * 256;
    float dzc_acc = 0.f;
    uint8_t mh_next = 1;
    {
      const int t0 = (tile_begin + p) * TB + tt;
      if (d.has_mask && active && p < ntile && t0 < d.T) mh_next = mrow[t0];
    }
    int use = 0;    // how many tiles this group has processed (parity of bk_free, phase of the d gate exchange)
    for (int i = p; i < ntile; i += 2, ++use) {
      const int s = i % S;
      const int t = (tile_begin + i) * TB + tt;
      const bool in_range = active && t < d.T;
      const bool mh = in_range && mh_next != 0;
      {
        const int t2 = t + 2 * TB;
        mh_next = 1;
        if (d.has_mask && active && i + 2 < ntile && t2 < d.T) mh_next = mrow[t2];
      }
      BT_TRACE(3, i, 0);
      mbar_wait_relaxed(&zq_ready[s], (i / S) & 1);
      BT_TRACE(3, i, 1);
      const float* zq = zq_s + s * 1024;
      float g = 1.f, tz = 0.f, cnt = 0.f;
      if (d.use_time_warp) {
        float z = zc;
        for (int kk = 0; kk < d.K; ++kk) z += zq[kk * TB + tt] + zq[128 + kk * TB + tt];
        const int tc = t < d.T ? t : d.T - 1;
        tz = 1.f - 2.f * rcp_approx(exp2f_approx(z * 2.885390081777927f) + 1.f);   // tanh(z), as the forward
        cnt = window_count(tc, d.T, d.warp_type, d.window);
        g = tz * cnt;
      }
      const int rr = active ? r : 0;
      const float quad = zq[256 + rr] + zq[384 + rr];
      const float lin = HAS_LIN ? zq[512 + rr] + zq[640 + rr] : 0.f;
      const float hg = zq[768 + rr] + zq[896 + rr];
      float c1, c0v;
      row_coefs(d, g, lin, quad, c1, c0v);
      const float p_c0 = cosine ? 0.f : c0v;
      const float c1s = c1 * s_scale, c0s = p_c0 * s_scale;
      const float2 c1c1 = make_float2(c1s, c1s), c0c0 = make_float2(c0s, c0s);
      // the previous tile of this group must have left the records and entry lists of this parity
      if (use >= 1) mbar_wait_relaxed(&bk_free[p], (use - 1) & 1);
      BT_TRACE(3, i, 2);
      BT_WAIT_HOT(&acc_full[p], (i >> 1) & 1);
      tc_fence_after();
      BT_TRACE(3, i, 3);

      // ---- the logits with the forward's instruction sequence; row maximum, its first column, ties.
      //      Straight-line selects: a branchy argmax costs 2x here (the pass is issue-bound). ----
      float m = -INFINITY, Pm = 0.f;
      int jm = 0, rowcnt = 0;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t x0[16], x1[16];
        tmem_ld16_issue(taddr + 16 * hh, x0);
        tmem_ld16_issue(taddr + 32 + 16 * hh, x1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 cq = col4[(16 * hh + j) >> 2], kq = cap4[(16 * hh + j) >> 2];
          const float colr0 = (j & 2) ? cq.z : cq.x, colr1 = (j & 2) ? cq.w : cq.y;
          const float capr0 = (j & 2) ? kq.z : kq.x, capr1 = (j & 2) ? kq.w : kq.y;
          const float2 P = __fadd2_rn(make_float2(__uint_as_float(x0[j]), __uint_as_float(x0[j + 1])),
                                      make_float2(__uint_as_float(x1[j]), __uint_as_float(x1[j + 1])));
          float2 Sv = __fadd2_rn(__ffma2_rn(c1c1, P, c0c0), make_float2(colr0, colr1));
          if (tanh_on) {
            const float2 den = __fadd2_rn(make_float2(exp2f_approx(Sv.x), exp2f_approx(Sv.y)), make_float2(1.f, 1.f));
            Sv = __ffma2_rn(make_float2(rcp_approx(den.x), rcp_approx(den.y)), make_float2(-2.f, -2.f), make_float2(1.f, 1.f));
          }
          float a0 = fminf(Sv.x, capr0), a1 = fminf(Sv.y, capr1);
          if (!mh) { a0 = fminf(capr0, FVTA_VERY_NEGATIVE); a1 = fminf(capr1, FVTA_VERY_NEGATIVE); }
          x0[j] = __float_as_uint(a0); x0[j + 1] = __float_as_uint(a1);        // logits
          x1[j] = __float_as_uint(P.x); x1[j + 1] = __float_as_uint(P.y);      // raw similarities
        }
        float mq = __uint_as_float(x0[0]);
#pragma unroll
        for (int j = 1; j < 16; ++j) mq = fmaxf(mq, __uint_as_float(x0[j]));
        int jh = 0, ch = 0;
        float Ph = 0.f;
#pragma unroll
        for (int j = 15; j >= 0; --j) {
          const bool eq = __uint_as_float(x0[j]) == mq;
          jh = eq ? j : jh;
          Ph = eq ? __uint_as_float(x1[j]) : Ph;
          ch += eq ? 1 : 0;
        }
        const bool gt = mq > m, eqm = mq == m;
        jm = gt ? 16 * hh + jh : jm;
        Pm = gt ? Ph : Pm;
        rowcnt = gt ? ch : (eqm ? rowcnt + ch : rowcnt);
        m = gt ? mq : m;
      }
      BT_TRACE(3, i, 6);

      // ---- upstream of the logits ----
      const float alpha = !in_range ? 0.f
                          : (sk <= -1e29f ? 1.f / (float)d.T : exp2f_approx((m - Mk) * 1.4426950408889634f) * rLk);
      const float e_up = beta * hg;                    // h . d r_k
      const float dm = alpha * (g * e_up - rdr);       // through the softmax over T
      float dg = alpha * e_up;                         // r_k = sum alpha * g * h
      const float dm_each = rowcnt > 0 ? dm / (float)rowcnt : 0.f;
      float dSsum = 0.f, dSP = 0.f;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;       // scalars, not an indexed array: the record stays in registers
      int j0 = 0, j1 = 0, j2 = 0, nrec = 0;
      // one non-zero d a[t, j]: the row sums it feeds; returns d P[t, j] (0: nothing to route)
      auto entry = [&](int j, float da, float P, float a) -> float {
        const float dS = tanh_on ? da * (1.f - a * a) : da;
        if (dS == 0.f || !(a > -1e29f)) return 0.f;     // saturated tanh / masked entry: no gradient
        dSsum += dS;
        dSP = fmaf(dS, P, dSP);
        if (!cosine) atomicAdd(dcol_s + j, dS);
        return c1 * dS;
      };
      // entry `slot` of this row joins the list of column j (the update warps walk the lists)
      auto push = [&](int slot, int j) {
        const int e = slot * 128 + r;
        nxt[e] = (uint16_t)atomicExch(head + j, e);
      };
      const bool want_g = in_range && tiesk == 1.f && t == amax_t;    // the unique sequence maximum sits in this row
      // fast path: a unique row maximum.  Everything else -- exact ties among unsaturated logits, the row
      // that holds the sequence maximum -- walks the columns once more, one at a time.
      const bool ties_m = in_range && rowcnt > 1 && dm != 0.f && !(tanh_on && fabsf(m) == 1.f);
      // (a saturated sequence maximum, |s_k| == 1 under tanh, has a zero derivative wherever it is attained)
      const bool ties_s = in_range && tiesk > 1.f && m == sk && dsk != 0.f && !(tanh_on && fabsf(sk) == 1.f);
      const bool slow = ties_m || ties_s || (want_g && dsk != 0.f);
      if (in_range && !slow && rowcnt == 1) {
        v0 = entry(jm, dm_each, Pm, m);
        if (v0 != 0.f) { j0 = jm; nrec = 1; push(0, jm); }
      }
      BT_TRACE(3, i, 7);
#ifdef FVTA_TC_TRACE
      { const unsigned bm = __ballot_sync(0xffffffffu, slow);
        if (lane == 0) { atomicAdd(&g_bt_count[0], 1); atomicAdd(&g_bt_count[1], bm != 0); atomicAdd(&g_bt_count[2], __popc(bm)); } }
#endif
      bool dense = false;     // more than three entries in this row: its d P vector stays in tensor memory
      if (__any_sync(0xffffffffu, slow)) {
        int ne = 0;
#pragma unroll 1
        for (int j = 0; j < kTcJQP; ++j) {
          __syncwarp();
          const uint32_t y0 = tmem_ld1(taddr + j), y1 = tmem_ld1(taddr + 32 + j);
          tmem_ld_wait();
          float dP = 0.f;
          if (slow) {
            const float P = __fadd_rn(__uint_as_float(y0), __uint_as_float(y1));
            float Sv = __fadd_rn(__fmaf_rn(c1s, P, c0s), col_s[j]);
            if (tanh_on) Sv = __fmaf_rn(rcp_approx(__fadd_rn(exp2f_approx(Sv), 1.f)), -2.f, 1.f);
            float a = fminf(Sv, cap_s[j]);
            if (!mh) a = fminf(cap_s[j], FVTA_VERY_NEGATIVE);
            float da = 0.f;
            if (a == m) da += dm_each;
            if (tiesk == 1.f ? (want_g && j == amax_j) : (a == sk)) da += ds_each;
            if (da != 0.f) dP = entry(j, da, P, a);
            if (dP != 0.f) {
              if (ne == 0) { v0 = dP; j0 = j; } else if (ne == 1) { v1 = dP; j1 = j; } else if (ne == 2) { v2 = dP; j2 = j; }
              ++ne;
            }
          }
          __syncwarp();
          tmem_st1(taddr + j, __float_as_uint(dP));     // column j has been read: it now holds d P[t, j]
        }
        tmem_st_wait();
        dense = slow && ne > 3;
        if (slow && !dense) {
          nrec = ne;
          if (ne > 0) push(0, j0);
          if (ne > 1) push(1, j1);
          if (ne > 2) push(2, j2);
        }
      }
      const bool dense_w = __any_sync(0xffffffffu, dense);
      if (!dense_w) {
        tc_fence_before();
        mbar_arrive(&ep_done[p]);
      }
      BT_TRACE(3, i, 4);

      float dlin = 0.f, dquad = 0.f;
      if (cosine) {
        const float vq = g * g * quad;
        if (vq > 1e-12f) {
          const float rs = rsqrtf(vq);
          dquad = dSP * (-0.5f) * g * g * g * rs * rs * rs;
        } else {
          dg = fmaf(dSP, 1e6f, dg);     // rsqrt(1e-12)
        }
      } else {
        dg += dSP + lin * dSsum + 2.f * g * quad * dSsum;
        dlin = g * dSsum;
        dquad = g * g * dSsum;
      }
      // ---- gate backward: d z[t] = (sum_k d g[k,t]) * cnt * (1 - tanh^2 z) ----
      float dz = 0.f;
      if (d.use_time_warp) {
        float* dgx = dgx_s + (p * 2 + (use & 1)) * 128;      // two phases: one barrier per tile is enough
        dgx[r] = in_range ? dg : 0.f;
        named_bar_sync(1 + p, 128);
        float dgt = 0.f;
        for (int kk = 0; kk < d.K; ++kk) dgt += dgx[kk * TB + tt];
        dz = in_range ? dgt * cnt * (1.f - tz * tz) : 0.f;
        if (k == 0) dzc_acc += dz;
      }
      const float hcoef = alpha * g * beta;
      if (dense_w) {
        // (rare: four or more exact ties among the unsaturated logits of one row, e.g. an all-zero context
        // row.)  The row's thread applies its own gradient, serially: d qeff and the coefficient sums go to
        // the global accumulators, d h[t, :] is written in place, and the update warps skip the row.
        float* trow = reinterpret_cast<float*>(ring + (size_t)s * pl.stage_bytes) + rr * D;
        float* dq_n = dqeffT + (size_t)n * Dp * JQPg;
        float* dc_n = dcoef + (size_t)n * 3 * Dp;
#pragma unroll 1
        for (int j = 0; j < kTcJQP; ++j) {
          __syncwarp();
          const float dP = __uint_as_float(tmem_ld1(taddr + j));
          tmem_ld_wait();
          if (dense && dP != 0.f && j < JQPg)
            for (int dd = 0; dd < D; ++dd) atomicAdd(dq_n + (size_t)dd * JQPg + j, dP * trow[dd]);
        }
        if (dense) {
          for (int dd = 0; dd < D; ++dd) {
            const float hv = trow[dd], hh = hv * hv;
            float o = fmaf(2.f * hv, fmaf(coef_s[2 * Dp + dd], dquad, coef_s[dd] * dz), hcoef * gvec_s[dd]);
            if (HAS_LIN) { o = fmaf(coef_s[Dp + dd], dlin, o); atomicAdd(dc_n + Dp + dd, dlin * hv); }
            trow[dd] = o;
            atomicAdd(dc_n + dd, dz * hh);
            atomicAdd(dc_n + 2 * Dp + dd, dquad * hh);
          }
        }
#pragma unroll 1
        for (int j = 0; j < kTcJQP; ++j) {
          __syncwarp();
          const float dP = __uint_as_float(tmem_ld1(taddr + j));
          tmem_ld_wait();
          if (dense && dP != 0.f)
            for (int dd = 0; dd < D; ++dd) {
              const int d4 = dd >> 2;
              const uint8_t* ph = b32_s + (d4 >> 3) * 8192 + j * 128 + (((d4 & 7) ^ (j & 7)) << 4) + (dd & 3) * 4;
              trow[dd] = fmaf(dP, *reinterpret_cast<const float*>(ph) + *reinterpret_cast<const float*>(ph + 4096), trow[dd]);
            }
        }
        __syncwarp();
        fence_async_smem();       // the row was written through the generic proxy; the TMA store reads it
        tc_fence_before();
        mbar_arrive(&ep_done[p]);
      }
      const uint32_t jbits = (uint32_t)j0 | ((uint32_t)j1 << 8) | ((uint32_t)j2 << 16) | ((uint32_t)(dense ? 0xff : nrec) << 24);
      rec[r * 2] = make_float4(v0, v1, v2, __uint_as_float(jbits));
      rec[r * 2 + 1] = make_float4(dlin, dquad, hcoef, dz);
      mbar_arrive(&rec_ready[s]);
      BT_TRACE(3, i, 5);
    }
    if (d.use_time_warp) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dzc_acc += __shfl_xor_sync(0xffffffffu, dzc_acc, o);
      if (lane == 0 && dzc_acc != 0.f) atomicAdd(dzc_s, dzc_acc);
    }
  } else {
    // ===== update warps =====
    reg_inc<kBtRegUpd>();
    const int u = threadIdx.x - kBwUpd0 * 32;          // 0..255
    // pass A ownership: pairs (column j, float4 group d4), idx = u + 256 q
    float4 accq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) accq[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int npair = kTcJQP * D4;
    // pass B ownership: float4 group d4 = u % D4, rows u / D4, u / D4 + nph, ...
    const int nph = 256 / D4;
    const bool b_on = u < nph * D4;
    const int bd4 = b_on ? u % D4 : 0, bph = u / D4;
    // (pass B runs on packed pairs -- FFMA2 -- throughout)
    float2 du01 = make_float2(0.f, 0.f), du23 = du01, dwl01 = du01, dwl23 = du01, dwq01 = du01, dwq23 = du01;
    float2 cu01, cu23, cq01, cq23, gv01, gv23;       // 2 u, 2 w_quad (the factor of d(h^2) folded in), g_hat
    {
      const float4 cu4 = *reinterpret_cast<const float4*>(coef_s + bd4 * 4);
      const float4 cq4 = *reinterpret_cast<const float4*>(coef_s + 2 * Dp + bd4 * 4);
      const float4 gv4 = *reinterpret_cast<const float4*>(gvec_s + bd4 * 4);
      cu01 = make_float2(2.f * cu4.x, 2.f * cu4.y); cu23 = make_float2(2.f * cu4.z, 2.f * cu4.w);
      cq01 = make_float2(2.f * cq4.x, 2.f * cq4.y); cq23 = make_float2(2.f * cq4.z, 2.f * cq4.w);
      gv01 = make_float2(gv4.x, gv4.y); gv23 = make_float2(gv4.z, gv4.w);
    }
    // question operand rows from the tf32 image (128-byte swizzled rows): q = hi + lo, float4 group d4 of row j
    const uint8_t* qbase = b32_s + (bd4 >> 3) * 8192 + ((bd4 & 7) << 4);
    auto q_ptr = [&](uint32_t j) -> const uint8_t* {
      return reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(qbase + j * 128) ^ (uintptr_t)((j & 7) << 4));
    };
    auto add_q = [&](uint32_t j, float v, float2& o01, float2& o23) {
      const uint8_t* ph = q_ptr(j);
      const float4 hi = *reinterpret_cast<const float4*>(ph);
      const float4 lo = *reinterpret_cast<const float4*>(ph + 4096);
      const float2 vv = make_float2(v, v);
      o01 = __ffma2_rn(vv, __fadd2_rn(make_float2(hi.x, hi.y), make_float2(lo.x, lo.y)), o01);
      o23 = __ffma2_rn(vv, __fadd2_rn(make_float2(hi.z, hi.w), make_float2(lo.z, lo.w)), o23);
    };
    for (int i = 0; i < ntile; ++i) {
      const int s = i % S, par = i & 1;
      BT_TRACE(4, i, 0);
      BT_WAIT_WARM(&rec_ready[s], (i / S) & 1);
      BT_TRACE(4, i, 1);
      float* tile = reinterpret_cast<float*>(ring + (size_t)s * pl.stage_bytes);
      const float4* rec = rec_s + par * 256;
      const int* head = head_s + par * 32;
      const uint16_t* nxt = next_s + par * 384;
      // ---- A: d qeff[j, 4 d4 ..] += sum over the entries listed under column j of v * h ----
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = u + 256 * q;
        if (idx < npair) {
          const int j = idx / D4, d4 = idx - j * D4;
          float4 a = accq[q];
          for (int e = head[j]; e != kBtEnd; e = nxt[e]) {
            const int rr = e & 127, slot = e >> 7;
            const float4 ra = rec[rr * 2];
            const float vv = slot == 0 ? ra.x : (slot == 1 ? ra.y : ra.z);
            const float4 h4 = *reinterpret_cast<const float4*>(tile + rr * D + d4 * 4);
            a.x = fmaf(vv, h4.x, a.x); a.y = fmaf(vv, h4.y, a.y); a.z = fmaf(vv, h4.z, a.z); a.w = fmaf(vv, h4.w, a.w);
          }
          accq[q] = a;
        }
      }
      BT_TRACE(4, i, 2);
      named_bar_sync(3, 256);           // every reader of h is done: the tile may be overwritten
      BT_TRACE(4, i, 3);
      // ---- B: d h in place, coefficient-vector sums.  Two rows per step: one warp issues a dependent
      //      instruction every ~5 cycles, so the role is bound by the length of its dependency chains, and
      //      two independent rows shorten that.  (Measured alternatives, cycles per 96-row tile: one row per
      //      step with the next row prefetched ~3900; this form 4300; all loads hoisted into two waves 5400;
      //      sixteen threads per row with two float4 groups each 7300.  Moving pass A into the row warpgroups
      //      -- they idle half of the time -- made the whole operator 407 us instead of 299: seven more float4
      //      accumulators do not fit next to the row pass at 96 registers and the spills land in both loops.) ----
      if (b_on) {
        auto row_math = [&](const float4& ra, const float4& rb, const float4& h4, float* hp, int nrec) {
          const uint32_t jb = __float_as_uint(ra.w);
          const bool skip = nrec == 0xff;     // no row here, or its own thread has written it (row warpgroups)
          const float2 h01 = make_float2(h4.x, h4.y), h23 = make_float2(h4.z, h4.w);
          const float dzv = skip ? 0.f : rb.w, dqv = skip ? 0.f : rb.y;
          const float2 dq2 = make_float2(dqv, dqv), dz2 = make_float2(dzv, dzv), hc2 = make_float2(rb.z, rb.z);
          // d h = 2 h (w_quad dquad + u dz) + g_hat hcoef (+ w_lin dlin) + sum_e v_e qeff[j_e]
          const float2 cc01 = __ffma2_rn(cq01, dq2, __fmul2_rn(cu01, dz2));
          const float2 cc23 = __ffma2_rn(cq23, dq2, __fmul2_rn(cu23, dz2));
          float2 o01 = __ffma2_rn(h01, cc01, __fmul2_rn(hc2, gv01));
          float2 o23 = __ffma2_rn(h23, cc23, __fmul2_rn(hc2, gv23));
          if (HAS_LIN) {
            const float4 cl4 = *reinterpret_cast<const float4*>(coef_s + Dp + bd4 * 4);
            const float dlv = skip ? 0.f : rb.x;
            const float2 dl2 = make_float2(dlv, dlv);
            o01 = __ffma2_rn(make_float2(cl4.x, cl4.y), dl2, o01);
            o23 = __ffma2_rn(make_float2(cl4.z, cl4.w), dl2, o23);
            dwl01 = __ffma2_rn(dl2, h01, dwl01);
            dwl23 = __ffma2_rn(dl2, h23, dwl23);
          }
          add_q(jb & 0x1f, ra.x, o01, o23);              // v0 = 0 (and j0 = 0) when the row has no entry
          if (nrec > 1 && !skip) {
            add_q((jb >> 8) & 0x1f, ra.y, o01, o23);
            add_q((jb >> 16) & 0x1f, ra.z, o01, o23);    // v2 = 0 when there are two
          }
          if (!skip) *reinterpret_cast<float4*>(hp) = make_float4(o01.x, o01.y, o23.x, o23.y);
          const float2 hh01 = __fmul2_rn(h01, h01), hh23 = __fmul2_rn(h23, h23);
          du01 = __ffma2_rn(dz2, hh01, du01);
          du23 = __ffma2_rn(dz2, hh23, du23);
          dwq01 = __ffma2_rn(dq2, hh01, dwq01);
          dwq23 = __ffma2_rn(dq2, hh23, dwq23);
        };
        for (int rr = bph; rr < rows; rr += 2 * nph) {
          float* hp0 = tile + rr * D + bd4 * 4;
          float* hp1 = hp0 + nph * D;
          const bool two = rr + nph < rows;
          const int r1 = two ? rr + nph : rr;
          const float4 ra0 = rec[rr * 2], rb0 = rec[rr * 2 + 1], h0 = *reinterpret_cast<const float4*>(hp0);
          const float4 ra1 = rec[r1 * 2], rb1 = rec[r1 * 2 + 1], h1 = *reinterpret_cast<const float4*>(two ? hp1 : hp0);
          const int n0 = (int)(__float_as_uint(ra0.w) >> 24);
          const int n1 = two ? (int)(__float_as_uint(ra1.w) >> 24) : 0xff;
          row_math(ra0, rb0, h0, hp0, n0);
          row_math(ra1, rb1, h1, hp1, n1);
        }
      }
      BT_TRACE(4, i, 4);
      fence_async_smem();               // the tile was written through the generic proxy; the TMA store reads it
      named_bar_sync(3, 256);
      BT_TRACE(4, i, 5);
      if (u < 32) {
        head_s[par * 32 + u] = kBtEnd;
        __syncwarp();
      }
      if (u == 0) {
        mbar_arrive(&bk_free[par]);
        tma_store_3d(&tmap_out, tile, 0, (tile_begin + i) * TB, n * d.K);
        tma_store_commit();
        // one store stays in flight: the stage of tile i - 1 is released once ITS store has read it
        if (i >= 1) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          mbar_arrive(&empty[(i - 1) % S]);
        }
      }
      BT_TRACE(4, i, 6);
    }
    if (u == 0 && ntile > 0) {
      tma_store_wait_read();
      mbar_arrive(&empty[(ntile - 1) % S]);
    }
    // ---- slab sums -> global accumulators ----
    float* dq_n = dqeffT + (size_t)n * Dp * JQPg;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = u + 256 * q;
      if (idx < npair) {
        const int j = idx / D4, d4 = idx - j * D4;
        if (j < JQPg) {
          float* dst = dq_n + (size_t)(d4 * 4) * JQPg + j;
          if (accq[q].x != 0.f) atomicAdd(dst, accq[q].x);
          if (accq[q].y != 0.f) atomicAdd(dst + JQPg, accq[q].y);
          if (accq[q].z != 0.f) atomicAdd(dst + 2 * JQPg, accq[q].z);
          if (accq[q].w != 0.f) atomicAdd(dst + 3 * JQPg, accq[q].w);
        }
      }
    }
    if (b_on) {
      float* dc = dcoef + (size_t)n * 3 * Dp + bd4 * 4;
      atomicAdd(dc, du01.x); atomicAdd(dc + 1, du01.y); atomicAdd(dc + 2, du23.x); atomicAdd(dc + 3, du23.y);
      if (HAS_LIN) { atomicAdd(dc + Dp, dwl01.x); atomicAdd(dc + Dp + 1, dwl01.y); atomicAdd(dc + Dp + 2, dwl23.x); atomicAdd(dc + Dp + 3, dwl23.y); }
      atomicAdd(dc + 2 * Dp, dwq01.x); atomicAdd(dc + 2 * Dp + 1, dwq01.y); atomicAdd(dc + 2 * Dp + 2, dwq23.x); atomicAdd(dc + 2 * Dp + 3, dwq23.y);
    }
    if (u == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every tile store has landed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
  }
  if (threadIdx.x < kTcJQP && threadIdx.x < JQPg && d.simi != 4) {
    const float sdc = dcol_s[threadIdx.x];
    if (sdc != 0.f) atomicAdd(dcol + (size_t)n * JQPg + threadIdx.x, sdc);
  }
  if (threadIdx.x == 0 && d.use_time_warp && dzc_s[0] != 0.f) atomicAdd(dzc + n, dzc_s[0]);
}

// Runs the tile kernel over a workspace whose folded operands AND tensor-core images are in place and
// whose accumulators (dqeffT .. dzc) are zeroed.  1 = shape not covered.
int attn_bwd_tc_try(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall, const uint8_t* hall_mask,
                    const uint8_t* hq_mask, float* ws, const float* stats, const float* dg1, float* d_hall,
                    cudaStream_t st) {
  TcPlan pl;
  if (!bt_make_plan(d, pl) || pl.stages < 2) return 1;
  CUtensorMap tin, tout;
  if (!tc_make_tile_map(&tin, hall, d.N, d.K, d.T, d.D, pl.TB) || !tc_make_tile_map(&tout, d_hall, d.N, d.K, d.T, d.D, pl.TB)) {
    set_last_error("cuTensorMapEncodeTiled failed", cudaErrorInvalidValue);
    return FVTA_E_LAUNCH;
  }
  const bool has_lin = (d.simi == 1 || d.simi == 3);
  auto kern = has_lin ? attn_bwd_tc<true, 0> : (d.D == 100 ? attn_bwd_tc<false, 100> : attn_bwd_tc<false, 0>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) { set_last_error("attn_bwd_tc smem attr", e); return FVTA_E_LAUNCH; }
  dim3 grid(pl.slabs, d.N);
  kern<<<grid, kBtThreads, pl.smem, st>>>(tin, tout, d, pl, hall_mask, hq_mask, ws + L.coef,
                                          reinterpret_cast<const uint8_t*>(ws + L.qB32),
                                          reinterpret_cast<const uint8_t*>(ws + L.qB16), ws + L.col, ws + L.zconst, stats,
                                          ws + L.drow, dg1, ws + L.dqeffT, ws + L.dcol, ws + L.dcoef, ws + L.dzc, L.Dp,
                                          L.JQP);
  return check_launch("attn_bwd_tc");
}

}  // namespace fvta

#ifdef FVTA_TC_TRACE
extern "C" int fvta_bt_iter_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, fvta::g_bt_iter, sizeof(long long) * 32);
}
extern "C" int fvta_bt_count_read(int* out) {
  return (int)cudaMemcpyFromSymbol(out, fvta::g_bt_count, sizeof(int) * 8);
}
extern "C" int fvta_bt_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, fvta::g_bt_trace, sizeof(long long) * 5 * fvta::kBtTraceTiles * 8);
}
#endif
