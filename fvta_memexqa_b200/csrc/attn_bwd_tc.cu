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

namespace fvta {

constexpr int kBtThreads = 896;
// warp 0 TMA, 1 MMA, 2-3 idle, 4-11 update, 12-19 split, 20-27 row warpgroups
constexpr int kBwTma = 0, kBwMma = 1, kBwUpd0 = 4, kBwSplit0 = 12, kBwRow0 = 20;
constexpr int kBtRegRow = 88, kBtRegUpd = 72, kBtRegIssue = 40;     // split keeps the launch value (72)
constexpr int kBtBucketCap = 48;       // entries per column bucket; the rest goes to the overflow list
constexpr int kBtOvCap = 64;
constexpr int kBtStageExtra = 4 * 1024 + 128 * 32;   // per stage: row statistics [4][2][128] + records [128][8]

struct BtOverflow { int r, j; float v; };

// fixed scratch: buckets 2 x (32 counts + 32 x cap u16), overflow 2 x (count + cap x 12 B), gvec [Dp<=128],
// dg exchange [2][128], dcol [32], per-k statistics [8][12], barriers
constexpr int kBtFixedExtra = 2 * (128 + 32 * kBtBucketCap * 2) + 2 * (16 + kBtOvCap * 12) + 512 + 1024 + 128 + 384 + 512 + 256;

static bool bt_make_plan(const FvtaAttnDesc& d, TcPlan& p) {
  return tc_plan(d, p, kBtFixedExtra, kBtStageExtra, 3);
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
  float* dgx_s = gvec_s + 128;                    // [2][128] per-row d gate of a tile (per row warpgroup)
  float* dcol_s = dgx_s + 256;                    // [32]
  float* kst_s = dcol_s + 32;                     // [8][12] per-modality statistics
  float* zq_s = kst_s + 96;                       // [S][4][2][128] half-row statistics: zdot, quad, lin, hg
  float4* rec_s = reinterpret_cast<float4*>(zq_s + S * 1024);   // [S][128][2] per-row records
  int* bcnt_s = reinterpret_cast<int*>(rec_s + S * 256);        // [2][32]
  uint16_t* blist_s = reinterpret_cast<uint16_t*>(bcnt_s + 64); // [2][32][cap]
  int* ovcnt_s = reinterpret_cast<int*>(blist_s + 2 * 32 * kBtBucketCap);   // [2] (+2 pad)
  BtOverflow* ov_s = reinterpret_cast<BtOverflow*>(ovcnt_s + 4);            // [2][cap]
  float* dzc_s = reinterpret_cast<float*>(ov_s + 2 * kBtOvCap);             // [1] (+3 pad)
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
    for (int i = threadIdx.x; i < 64; i += blockDim.x) bcnt_s[i] = 0;
    if (threadIdx.x < 2) ovcnt_s[threadIdx.x] = 0;
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
        mbar_wait_relaxed(&a_ready[b2], (i >> 1) & 1);
        if (i >= 2) mbar_wait_relaxed(&ep_done[b2], ((i - 2) >> 1) & 1);
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
      mbar_wait_relaxed(&full[s], (i / S) & 1);
      if (i >= 2) {
        mbar_wait_relaxed(&acc_full[b2], ((i - 2) >> 1) & 1);
        tc_fence_after();
      }
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
    const float Mk = ks[0], Lk = ks[1], sk = ks[2], tiesk = ks[3];
    const int amax_t = (int)ks[4], amax_j = (int)ks[5];
    const float dsk = ks[6], rdr = ks[7], beta = ks[8];
    const float ds_each = dsk / tiesk;
    int* bcnt = bcnt_s + p * 32;
    uint16_t* blist = blist_s + p * 32 * kBtBucketCap;
    uint8_t mh_next = 1;
    {
      const int t0 = (tile_begin + p) * TB + tt;
      if (d.has_mask && active && p < ntile && t0 < d.T) mh_next = mrow[t0];
    }
    int use = 0;    // how many tiles this group has processed (parity of bk_free)
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
      mbar_wait_relaxed(&zq_ready[s], (i / S) & 1);
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
      const float s_scale = tanh_on ? kTanhScale : 1.f;
      const float2 c1c1 = make_float2(c1 * s_scale, c1 * s_scale), c0c0 = make_float2(p_c0 * s_scale, p_c0 * s_scale);
      // the previous tile of this group must have left the buckets of this parity
      if (use >= 1) mbar_wait_relaxed(&bk_free[p], (use - 1) & 1);
      mbar_wait_relaxed(&acc_full[p], (i >> 1) & 1);
      tc_fence_after();

      // ---- logits with the forward's instruction sequence: row maximum, its first column, ties ----
      float m = -INFINITY, Pm = 0.f, Pg = 0.f, ag = 0.f;
      int jm = 0, rowcnt = 0;
      const bool want_g = in_range && tiesk == 1.f && t == amax_t;    // the unique sequence maximum sits in this row
      auto logits_pass = [&](auto&& visit) {
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
            visit(16 * hh + j, a0, P.x);
            visit(16 * hh + j + 1, a1, P.y);
          }
        }
      };
      logits_pass([&](int j, float a, float P) {
        if (a > m) { m = a; jm = j; Pm = P; rowcnt = 1; }
        else if (a == m) { rowcnt++; }
        if (want_g && j == amax_j) { Pg = P; ag = a; }
      });

      // ---- upstream of the logits ----
      const float alpha = !in_range ? 0.f : (sk <= -1e29f ? 1.f / (float)d.T : expf(m - Mk) / Lk);
      const float e_up = beta * hg;                    // h . d r_k
      const float dm = alpha * (g * e_up - rdr);       // through the softmax over T
      float dg = alpha * e_up;                         // r_k = sum alpha * g * h
      const float dm_each = rowcnt > 0 ? dm / (float)rowcnt : 0.f;
      float dSsum = 0.f, dSP = 0.f;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;       // scalars, not an indexed array: the record stays in registers
      int j0 = 0, j1 = 0, j2 = 0, nrec = 0;
      auto add_entry = [&](int j, float da, float P, float a) {
        const float dS = tanh_on ? da * (1.f - a * a) : da;
        if (dS == 0.f || !(a > -1e29f)) return;        // saturated tanh / masked entry: no gradient
        dSsum += dS;
        dSP = fmaf(dS, P, dSP);
        const float dP = c1 * dS;
        if (!cosine) atomicAdd(dcol_s + j, dS);
        if (dP == 0.f) return;
        if (nrec < 3) {
          const int pos = atomicAdd(bcnt + j, 1);
          if (pos < kBtBucketCap) {
            blist[j * kBtBucketCap + pos] = (uint16_t)(r | (nrec << 8));
            if (nrec == 0) { v0 = dP; j0 = j; } else if (nrec == 1) { v1 = dP; j1 = j; } else { v2 = dP; j2 = j; }
            ++nrec;
            return;
          }
        }
        const int o = atomicAdd(ovcnt_s + p, 1);
        if (o < kBtOvCap) { ov_s[p * kBtOvCap + o].r = r; ov_s[p * kBtOvCap + o].j = j; ov_s[p * kBtOvCap + o].v = dP; }
        else __trap();      // more exact ties than any real input produces
      };
      // fast path: a unique row maximum (+ the unique sequence maximum when it sits in this row)
      const bool ties_m = in_range && rowcnt > 1 && dm != 0.f && !(tanh_on && fabsf(m) == 1.f);
      // (a saturated sequence maximum, |s_k| == 1 under tanh, has a zero derivative wherever it is attained)
      const bool ties_s = in_range && tiesk > 1.f && m == sk && dsk != 0.f && !(tanh_on && fabsf(sk) == 1.f);
      const bool slow = ties_m || ties_s;
      if (in_range && !slow) {
        if (rowcnt == 1) add_entry(jm, dm_each, Pm, m);
        if (want_g) add_entry(amax_j, ds_each, Pg, ag);
      }
      if (__any_sync(0xffffffffu, slow)) {
        // exact ties among unsaturated logits (rare): walk the columns again and route per column
        logits_pass([&](int j, float a, float P) {
          if (!slow) return;
          float da = 0.f;
          if (a == m) da += dm_each;
          if (tiesk == 1.f ? (want_g && j == amax_j) : (a == sk)) da += ds_each;
          if (da != 0.f) add_entry(j, da, P, a);
        });
      }
      tc_fence_before();
      mbar_arrive(&ep_done[p]);

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
        dgx_s[p * 128 + r] = in_range ? dg : 0.f;
        named_bar_sync(1 + p, 128);
        float dgt = 0.f;
        for (int kk = 0; kk < d.K; ++kk) dgt += dgx_s[p * 128 + kk * TB + tt];
        dz = in_range ? dgt * cnt * (1.f - tz * tz) : 0.f;
        if (k == 0 && in_range && dz != 0.f) atomicAdd(dzc_s, dz);
        named_bar_sync(1 + p, 128);     // dgx is rewritten by this group's next tile
      }
      const float hcoef = alpha * g * beta;
      const uint32_t jbits = (uint32_t)j0 | ((uint32_t)j1 << 8) | ((uint32_t)j2 << 16) | ((uint32_t)nrec << 24);
      rec_s[(s * 128 + r) * 2] = make_float4(v0, v1, v2, __uint_as_float(jbits));
      rec_s[(s * 128 + r) * 2 + 1] = make_float4(dlin, dquad, hcoef, dz);
      mbar_arrive(&rec_ready[s]);
    }
  } else {
    // ===== update warps =====
    reg_dec<kBtRegUpd>();
    const int u = threadIdx.x - kBwUpd0 * 32;          // 0..255
    // pass A ownership: pairs (column j, float4 group d4), idx = u + 256 q
    float4 accq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) accq[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int npair = kTcJQP * D4;
    // pass B ownership: float4 group d4 = u % D4, rows u / D4, u / D4 + nph, ...
    const int nph = 256 / D4;
    const bool b_on = u < nph * D4;
    const int bd4 = u % D4, bph = u / D4;
    // (the coefficient vectors are re-read from shared memory per row: keeping them in registers next to
    // the seven float4 accumulators spills at the 72 registers this role gets)
    float4 du4 = make_float4(0.f, 0.f, 0.f, 0.f), dwl4 = du4, dwq4 = du4;
    const float* cu_p = coef_s + (b_on ? bd4 * 4 : 0);
    const float* gv_p = gvec_s + (b_on ? bd4 * 4 : 0);
    // question operand rows from the tf32 image: q = hi + lo, float4 group d4 of row j
    const uint32_t qoff = (uint32_t)((bd4 >> 3) * 8192), qunit = (uint32_t)(bd4 & 7);
    auto qrow4 = [&](int j) -> float4 {
      const uint8_t* ph = b32_s + qoff + j * 128 + ((qunit ^ (uint32_t)(j & 7)) << 4);
      const float4 hi = *reinterpret_cast<const float4*>(ph);
      const float4 lo = *reinterpret_cast<const float4*>(ph + 4096);
      return make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
    };
    for (int i = 0; i < ntile; ++i) {
      const int s = i % S, par = i & 1;
      mbar_wait_relaxed(&rec_ready[s], (i / S) & 1);
      float* tile = reinterpret_cast<float*>(ring + (size_t)s * pl.stage_bytes);
      const float4* rec = rec_s + s * 256;
      const int* bcnt = bcnt_s + par * 32;
      const uint16_t* blist = blist_s + par * 32 * kBtBucketCap;
      const int novf = min(ovcnt_s[par], kBtOvCap);
      const BtOverflow* ov = ov_s + par * kBtOvCap;
      // ---- A: d qeff[j, 4 d4 ..] += sum over the rows bucketed under column j of v * h ----
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = u + 256 * q;
        if (idx < npair) {
          const int j = idx / D4, d4 = idx - j * D4;
          const int ne = min(bcnt[j], kBtBucketCap);
          float4 a = accq[q];
          for (int e = 0; e < ne; ++e) {
            const int ref = blist[j * kBtBucketCap + e];
            const int rr = ref & 0xff, slot = ref >> 8;
            const float4 ra = rec[rr * 2];
            const float vv = slot == 0 ? ra.x : (slot == 1 ? ra.y : ra.z);
            const float4 h4 = *reinterpret_cast<const float4*>(tile + rr * D + d4 * 4);
            a.x = fmaf(vv, h4.x, a.x); a.y = fmaf(vv, h4.y, a.y); a.z = fmaf(vv, h4.z, a.z); a.w = fmaf(vv, h4.w, a.w);
          }
          for (int o = 0; o < novf; ++o)
            if (ov[o].j == j) {
              const float4 h4 = *reinterpret_cast<const float4*>(tile + ov[o].r * D + d4 * 4);
              const float vv = ov[o].v;
              a.x = fmaf(vv, h4.x, a.x); a.y = fmaf(vv, h4.y, a.y); a.z = fmaf(vv, h4.z, a.z); a.w = fmaf(vv, h4.w, a.w);
            }
          accq[q] = a;
        }
      }
      named_bar_sync(3, 256);           // every reader of h is done: the tile may be overwritten
      // ---- B: d h in place, coefficient-vector sums ----
      if (b_on) {
        for (int rr = bph; rr < rows; rr += nph) {
          const float4 ra = rec[rr * 2], rb = rec[rr * 2 + 1];
          const uint32_t jb = __float_as_uint(ra.w);
          const int nrec = (int)(jb >> 24);
          float* hp = tile + rr * D + bd4 * 4;
          const float4 h4 = *reinterpret_cast<const float4*>(hp);
          const float dlin = rb.x, dquad = rb.y, hcoef = rb.z, dz = rb.w;
          const float4 hh = make_float4(h4.x * h4.x, h4.y * h4.y, h4.z * h4.z, h4.w * h4.w);
          const float4 cu4 = *reinterpret_cast<const float4*>(cu_p);
          const float4 cq4 = *reinterpret_cast<const float4*>(cu_p + 2 * Dp);
          const float4 gv4 = *reinterpret_cast<const float4*>(gv_p);
          // d h = 2 h (w_quad dquad + u dz) + g_hat hcoef (+ w_lin dlin)
          const float4 cc = make_float4(2.f * fmaf(cq4.x, dquad, cu4.x * dz), 2.f * fmaf(cq4.y, dquad, cu4.y * dz),
                                        2.f * fmaf(cq4.z, dquad, cu4.z * dz), 2.f * fmaf(cq4.w, dquad, cu4.w * dz));
          float4 o = make_float4(fmaf(h4.x, cc.x, hcoef * gv4.x), fmaf(h4.y, cc.y, hcoef * gv4.y),
                                 fmaf(h4.z, cc.z, hcoef * gv4.z), fmaf(h4.w, cc.w, hcoef * gv4.w));
          if (HAS_LIN) {
            const float4 cl4 = *reinterpret_cast<const float4*>(cu_p + Dp);
            o.x = fmaf(cl4.x, dlin, o.x); o.y = fmaf(cl4.y, dlin, o.y); o.z = fmaf(cl4.z, dlin, o.z); o.w = fmaf(cl4.w, dlin, o.w);
          }
          if (nrec > 0) {
            const float4 q0 = qrow4((int)(jb & 0xff));
            o.x = fmaf(ra.x, q0.x, o.x); o.y = fmaf(ra.x, q0.y, o.y); o.z = fmaf(ra.x, q0.z, o.z); o.w = fmaf(ra.x, q0.w, o.w);
            if (nrec > 1) {
              const float4 q1 = qrow4((int)((jb >> 8) & 0xff));
              o.x = fmaf(ra.y, q1.x, o.x); o.y = fmaf(ra.y, q1.y, o.y); o.z = fmaf(ra.y, q1.z, o.z); o.w = fmaf(ra.y, q1.w, o.w);
              if (nrec > 2) {
                const float4 q2 = qrow4((int)((jb >> 16) & 0xff));
                o.x = fmaf(ra.z, q2.x, o.x); o.y = fmaf(ra.z, q2.y, o.y); o.z = fmaf(ra.z, q2.z, o.z); o.w = fmaf(ra.z, q2.w, o.w);
              }
            }
          }
          for (int oo = 0; oo < novf; ++oo)
            if (ov[oo].r == rr) {
              const float4 qx = qrow4(ov[oo].j);
              const float vv = ov[oo].v;
              o.x = fmaf(vv, qx.x, o.x); o.y = fmaf(vv, qx.y, o.y); o.z = fmaf(vv, qx.z, o.z); o.w = fmaf(vv, qx.w, o.w);
            }
          *reinterpret_cast<float4*>(hp) = o;
          du4.x = fmaf(dz, hh.x, du4.x); du4.y = fmaf(dz, hh.y, du4.y); du4.z = fmaf(dz, hh.z, du4.z); du4.w = fmaf(dz, hh.w, du4.w);
          if (HAS_LIN) { dwl4.x = fmaf(dlin, h4.x, dwl4.x); dwl4.y = fmaf(dlin, h4.y, dwl4.y); dwl4.z = fmaf(dlin, h4.z, dwl4.z); dwl4.w = fmaf(dlin, h4.w, dwl4.w); }
          dwq4.x = fmaf(dquad, hh.x, dwq4.x); dwq4.y = fmaf(dquad, hh.y, dwq4.y); dwq4.z = fmaf(dquad, hh.z, dwq4.z); dwq4.w = fmaf(dquad, hh.w, dwq4.w);
        }
      }
      fence_async_smem();               // the tile was written through the generic proxy; the TMA store reads it
      named_bar_sync(3, 256);
      if (u < 32) bcnt_s[par * 32 + u] = 0;
      if (u == 32) ovcnt_s[par] = 0;
      if (u < 64) named_bar_sync(4, 64);   // the two warps that reset the bucket state
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
      atomicAdd(dc, du4.x); atomicAdd(dc + 1, du4.y); atomicAdd(dc + 2, du4.z); atomicAdd(dc + 3, du4.w);
      if (HAS_LIN) { atomicAdd(dc + Dp, dwl4.x); atomicAdd(dc + Dp + 1, dwl4.y); atomicAdd(dc + Dp + 2, dwl4.z); atomicAdd(dc + Dp + 3, dwl4.w); }
      atomicAdd(dc + 2 * Dp, dwq4.x); atomicAdd(dc + 2 * Dp + 1, dwq4.y); atomicAdd(dc + 2 * Dp + 2, dwq4.z); atomicAdd(dc + 2 * Dp + 3, dwq4.w);
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
