// Focal visual-text attention forward on the 5th-generation tensor cores (sm_100a only).
//
// Same math, same outputs and the same tile-partial / combine protocol as attn_fwd.cu, but the
// question x context contraction P[t,j] = h[t,:] . qeff[j,:] -- the part that makes the FFMA
// version compute-bound at JQ/2 = 16 flop/B -- runs as tcgen05.mma tiles with the accumulator
// in TMEM, and the context tile arrives by TMA.  One pass over hall:
//
//   TMA warp      : 3-stage ring of fp32 context tiles  [K modalities x TB positions x D]
//                   (cp.async.bulk.tensor.3d, mbarrier complete_tx)
//   4 math warps  : one thread per context row.  pass 1: gate / row statistics and the split of
//                   the row into fp16 (hi, lo*2^11) operands written in the 128B-swizzled K-major
//                   layout UMMA reads;   epilogue: TMEM -> registers, logits, mask, row max,
//                   softmax weight;   pass 2: weighted sum of the (still resident) fp32 tile
//   MMA warp      : one elected thread issues, per 16-wide k-step,
//                     D[:, 0:64] += A_hi . [B_hi ; B_lo]^T   and   D[:,64:96] += A_lo . B_hi^T
//                   so P = D0 + (D1 + D2) * 2^-11 carries ~22 mantissa bits (fp32 accumulate):
//                   the 1e-4 parity bar holds where single-pass tf32/bf16 would not.
//
// The softmax over T needs no online rescaling here: this kernel is used when the logits are
// bounded by 1 (tanh or cosine), so exp(m - 1) cannot overflow and every tile uses the fixed
// reference maximum 1.  Masked rows get the weight 2^-100 instead of exp(-1e30 - 1) = 0: it is
// absorbed exactly by any unmasked weight and reproduces TF's uniform softmax over a fully
// masked sequence.
#include "common.cuh"
#include "attn_tile.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace fvta {

// Optional per-role timeline (build with -DFVTA_TC_TRACE, run tools/tc_trace.py): CTA (0,0) stores
// the SM clock at each pipeline event of its first kTraceTiles tiles into g_tc_trace.
#ifdef FVTA_TC_TRACE
constexpr int kTraceTiles = 24;
__device__ long long g_tc_trace[4][kTraceTiles][8];
#define TC_TRACE(role, i, ev)                                                                     \
  do { if (blockIdx.x == 0 && blockIdx.y == 0 && (i) < kTraceTiles && (threadIdx.x == 0 || threadIdx.x == 256 || threadIdx.x == 512 || threadIdx.x == 544)) \
         g_tc_trace[role][i][ev] = clock64(); } while (0)
#else
#define TC_TRACE(role, i, ev) do {} while (0)
#endif

int attn_combine_launch(const FvtaAttnDesc& d, const float* part_stat, const float* part_o, float* g1_all,
                        float* stats, int nt, int Dp, cudaStream_t st);
extern int g_attn_stages;

constexpr int kTcThreads = 576;  // warps 0-7 split, 8-15 softmax, warp 16 TMA, warp 17 MMA
constexpr int kTcStages = 3;
constexpr int kTcJQP = 32;       // question positions per tile (N of the MMA = 64 + 32)
constexpr int kTcTmemCols = 256; // 2 accumulator buffers x 128 columns (96 used)
constexpr float kMaskedWeight = 7.888609052210118e-31f;  // 2^-100

struct TcPlan {
  int TB;           // context positions per tile
  int rows;         // K * TB <= 128
  int KC;           // 64-wide fp16 K chunks (128-byte swizzle lines) = ceil(D / 64)
  int ksteps;       // 16-wide MMA k-steps = ceil(D / 16)
  int stage_bytes;  // fp32 tile, rounded to 128 B
  int tiles;        // tiles per sample
  int slabs;        // CTAs per sample
  size_t smem;      // dynamic shared memory
};

__host__ __device__ inline size_t tc_smem_bytes(int KC, int stage_bytes, int Dp) {
  size_t b = 0;
  b += (size_t)2 * KC * 16384;                      // A_hi, A_lo
  b += (size_t)KC * 8192;                           // [B_hi ; B_lo]
  b += (size_t)kTcStages * stage_bytes;             // fp32 ring
  b += (size_t)(3 * Dp + 2 * kTcJQP + 256 + 512 + 2 * 256 + 4 * 512) * sizeof(float);  // coef, col, cap, wg, mx, st, zq/red
  b += 256;                                         // barriers, tmem pointer
  return b;
}

static bool tc_make_plan(const FvtaAttnDesc& d, TcPlan& p) {
  if ((d.D & 3) != 0 || d.D > 128 || d.D < 8 || d.JQ > kTcJQP || d.K > 8) return false;
  const bool bounded = (d.simi == 4) || d.add_tanh;
  if (!bounded || d.time_warp_att) return false;
  p.KC = (d.D + 63) / 64;
  p.ksteps = (d.D + 15) / 16;
  for (int TB = 128 / d.K; TB >= 4; --TB) {
    p.TB = TB;
    p.rows = d.K * TB;
    p.stage_bytes = round_up(p.rows * d.D * 4, 128);
    p.smem = tc_smem_bytes(p.KC, p.stage_bytes, round_up(d.D, 4));
    if (p.smem <= 227 * 1024) break;
    p.TB = 0;
  }
  if (p.TB == 0) return false;
  p.tiles = (d.T + p.TB - 1) / p.TB;
  int want = (148 + d.N - 1) / d.N;  // one CTA per SM
  p.slabs = want < 1 ? 1 : (want > p.tiles ? p.tiles : want);
  return true;
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
// Waiting warps must not compete for issue slots with the warps they wait for (the scheduler
// favours high warp ids, and a try_wait spin loop is always ready): back off between polls.
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned sleep_ns = 40) {
  while (!mbar_try(bar, parity)) __nanosleep(sleep_ns);
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
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// issue only; pair with tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand: 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, A/B = fp16 K-major, D = fp32, M = 128
__device__ __forceinline__ uint32_t umma_idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }

// MUFU.EX2 / MUFU.RCP; tanh(x) = 1 - 2 / (1 + 2^{2 log2(e) x}) built from them has |error| < 1e-6
// and never exceeds 1 in magnitude
__device__ __forceinline__ float exp2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- B operand: [B_hi ; B_lo] as fp16, pre-swizzled, one image per sample ----------------------
__global__ void attn_prep_qb(FvtaAttnDesc d, const float* __restrict__ qeffT, __half* __restrict__ qB, int Dp,
                             int JQP, int KC) {
  const int n = blockIdx.x;
  const float* q = qeffT + (size_t)n * Dp * JQP;
  uint8_t* out = reinterpret_cast<uint8_t*>(qB) + (size_t)n * KC * 8192;
  for (int i = threadIdx.x; i < KC * 64 * 8; i += blockDim.x) {
    const int c = i & 7, row = (i >> 3) & 63, kc = i >> 9;
    const int j = row & 31;
    const bool lo = row >= 32;
    __half v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int dd = kc * 64 + c * 8 + e;
      const float x = (dd < d.D && j < d.JQ) ? q[(size_t)dd * JQP + j] : 0.f;
      const __half hi = __float2half_rn(x);
      v[e] = lo ? __float2half_rn(x - __half2float(hi)) : hi;
    }
    *reinterpret_cast<uint4*>(out + kc * 8192 + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
}

// ---- the fused kernel: grid (slabs, N) ----------------------------------------------------------
// 18 warps.  Two threads share one context row so that each scheduler has four math warps to
// interleave (the per-tile work is ~2.5k dependent-ish instructions per row):
//   warps 0-7   "split"   : pass 1 of tile i+1 -- thread (r, h) converts the 16-byte chunks c = h, h+2, ...
//                           of row r into the fp16 hi/lo operands and accumulates the row statistics
//   warps 8-15  "softmax" : epilogue + pass 2 of tile i -- thread (r, h) owns accumulator columns
//                           [16h, 16h+16) of TMEM lane r (warp & 3 selects the lane quadrant)
//   warp 16 TMA producer,  warp 17 MMA issuer (+ TMEM allocation)
template <bool HAS_LIN>
__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc(
    const __grid_constant__ CUtensorMap tmap, FvtaAttnDesc d, TcPlan pl, const uint8_t* __restrict__ hall_mask,
    const uint8_t* __restrict__ hq_mask, const float* __restrict__ coef, const __half* __restrict__ qB,
    const float* __restrict__ col, const float* __restrict__ zconst, float* __restrict__ a_logits,
    float* __restrict__ c_col, float* __restrict__ part_stat, float* __restrict__ part_o, int Dp, int JQPg) {
  // 1024-byte alignment is what the 128B swizzle pattern is defined against; keeping the pointers
  // derived from the __shared__ symbol (no integer round-trip) lets ptxas emit LDS/STS, not LD/ST
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + pl.KC * 16384;
  uint8_t* b_s = a_lo + pl.KC * 16384;
  uint8_t* ring = b_s + pl.KC * 8192;
  float* coef_s = reinterpret_cast<float*>(ring + kTcStages * pl.stage_bytes);  // [3][Dp]
  float* col_s = coef_s + 3 * Dp;      // [32]   per-question-word term + bias
  float* cap_s = col_s + kTcJQP;       // [32]   +inf for a valid question word, -1e30 masked, -inf padding
  float* wg_s = cap_s + kTcJQP;        // [2][128]     softmax weight * gate
  float* mx_s = wg_s + 256;            // [2][2][128]  half-row maxima
  float2* st_s = reinterpret_cast<float2*>(mx_s + 512);  // [2][128]  (c1, second coefficient) per row
  float4* zq_s = reinterpret_cast<float4*>(st_s + 256);  // [2][2][128] half-row (zdot, quad, lin, -)
  float* red_s = reinterpret_cast<float*>(zq_s);         // [5][256], aliases zq_s after the last tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(zq_s + 512);
  uint64_t* full = bars;               // [3]  TMA -> split, softmax
  uint64_t* empty = bars + 3;          // [3]  softmax -> TMA
  uint64_t* a_ready = bars + 6;        // [2]  split -> MMA, one per 64-feature operand chunk
  uint64_t* a_free = bars + 8;         // [2]  MMA -> split: the chunk's MMAs have read the operand
  uint64_t* acc_full = bars + 10;      // [2]  MMA -> softmax (accumulator ready)
  uint64_t* st_ready = bars + 12;      // [2]  split -> softmax (row coefficients)
  uint64_t* ep_done = bars + 14;       // [2]  softmax -> split (accumulator + coefficient buffer free)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x, n = blockIdx.y;
  const int tile_begin = (int)(((long long)slab * pl.tiles) / pl.slabs);
  const int tile_end = (int)(((long long)(slab + 1) * pl.tiles) / pl.slabs);
  const int ntile = tile_end - tile_begin;
  const int D = d.D, TB = pl.TB, rows = pl.rows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 256); }
    for (int b2 = 0; b2 < 2; ++b2) { mbar_init(&a_ready[b2], 256); mbar_init(&a_free[b2], 1); mbar_init(&acc_full[b2], 1); mbar_init(&st_ready[b2], 256); mbar_init(&ep_done[b2], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(kTcTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // per-sample constants
  {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(qB) + (size_t)n * pl.KC * 8192);
    uint4* dst = reinterpret_cast<uint4*>(b_s);
    for (int i = threadIdx.x; i < pl.KC * 512; i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < 3 * Dp; i += blockDim.x) coef_s[i] = coef[i];
    for (int j = threadIdx.x; j < kTcJQP; j += blockDim.x) {
      col_s[j] = j < JQPg ? col[(size_t)n * JQPg + j] : 0.f;
      const bool qv = !d.has_mask || (j < d.JQ && hq_mask[(size_t)n * d.JQ + j]);
      cap_s[j] = j >= d.JQ ? -INFINITY : (qv ? INFINITY : FVTA_VERY_NEGATIVE);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 16) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < ntile; ++i) {
        const int s = i % kTcStages;
        mbar_wait(&empty[s], ((i / kTcStages) & 1) ^ 1, 200);
        TC_TRACE(0, i, 0);
        mbar_expect_tx(&full[s], (uint32_t)(rows * D * 4));
        tma_load_3d(ring + s * pl.stage_bytes, &tmap, &full[s], 0, (tile_begin + i) * TB, n * d.K);
      }
    }
  } else if (warp == 17) {
    // ===== MMA issuer =====
    // per k-step:  D[:, 0:64] += A_hi . [B_hi ; B_lo]^T   and   D[:, 32:64] += A_lo . B_hi^T
    // so the epilogue adds two 32-column groups: P = D[:, 0:32] + D[:, 32:64]
    if (lane == 0) {
      const uint32_t idesc64 = umma_idesc_f16(64), idesc32 = umma_idesc_f16(32);
      const uint32_t a_hi_u = smem_u32(a_hi), a_lo_u = smem_u32(a_lo), b_u = smem_u32(b_s);
      for (int i = 0; i < ntile; ++i) {
        const uint32_t acc = tmem_base + (uint32_t)((i & 1) * 128);
        for (int kc = 0; kc < pl.KC; ++kc) {
          mbar_wait(&a_ready[kc], i & 1, 60);
          TC_TRACE(1, i, kc);
          tc_fence_after();
          const int ks_end = min(pl.ksteps, kc * 4 + 4);
          for (int ks = kc * 4; ks < ks_end; ++ks) {
            const uint32_t off = (uint32_t)(kc * 16384 + (ks & 3) * 32);
            const uint32_t boff = (uint32_t)(kc * 8192 + (ks & 3) * 32);
            umma_f16(acc, umma_desc_sw128(a_hi_u + off), umma_desc_sw128(b_u + boff), idesc64, ks > 0);
            umma_f16(acc + 32, umma_desc_sw128(a_lo_u + off), umma_desc_sw128(b_u + boff), idesc32, 1u);
          }
          umma_commit(&a_free[kc]);   // the split warps may overwrite this chunk for the next tile
        }
        umma_commit(&acc_full[i & 1]);
        TC_TRACE(1, i, 2);
      }
    }
  } else if (warp < 8) {
    // ===== split warps: threads (r, 0) and (r, 1) share context row r = k*TB + tt =====
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    const bool active = r < rows;
    const int k = active ? r / TB : 0, tt = active ? r - k * TB : 0;
    const int nch = pl.ksteps * 2;   // 16-byte (8 x fp16) chunks the MMA reads
    const int nfull = D >> 3;        // chunks holding 8 real features
    uint8_t* ahi_row = a_hi + r * 128;
    uint8_t* alo_row = a_lo + r * 128;
    const float zc = d.use_time_warp ? zconst[n] : 0.f;
    uint32_t sw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sw[j] = (uint32_t)(((h + 2 * j) ^ (r & 7)) << 4);
    for (int i = 0; i < ntile; ++i) {
      const int s = i % kTcStages, b2 = i & 1;
      TC_TRACE(2, i, 0);
      mbar_wait(&full[s], (i / kTcStages) & 1);
      if (i >= 2) mbar_wait(&ep_done[b2], ((i - 2) >> 1) & 1);             // softmax(i-2) released buffers b2
      TC_TRACE(2, i, 1);
      // packed fp32x2 arithmetic (FFMA2/FMUL2); two accumulator sets keep the dependent chains short
      float2 zdA = make_float2(0.f, 0.f), qdA = make_float2(0.f, 0.f), lnA = make_float2(0.f, 0.f);
      float2 zdB = make_float2(0.f, 0.f), qdB = make_float2(0.f, 0.f), lnB = make_float2(0.f, 0.f);
      const float* hrow = reinterpret_cast<const float*>(ring + s * pl.stage_bytes) + (size_t)r * D;
      const float2 neg1 = make_float2(-1.f, -1.f);
      auto split_store = [&](uint32_t off, const float4& v0, const float4& v1) {
        const __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
        const __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
        // residual x - fp16(x) is exact in fp32; below 2^-14 it lands on fp16 subnormals whose
        // 2^-25 absolute rounding is under the fp32 rounding of the operand itself
        const float2 r0 = __ffma2_rn(__half22float2(h0), neg1, make_float2(v0.x, v0.y));
        const float2 r1 = __ffma2_rn(__half22float2(h1), neg1, make_float2(v0.z, v0.w));
        const float2 r2 = __ffma2_rn(__half22float2(h2), neg1, make_float2(v1.x, v1.y));
        const float2 r3 = __ffma2_rn(__half22float2(h3), neg1, make_float2(v1.z, v1.w));
        const __half2 l0 = __floats2half2_rn(r0.x, r0.y), l1 = __floats2half2_rn(r1.x, r1.y);
        const __half2 l2 = __floats2half2_rn(r2.x, r2.y), l3 = __floats2half2_rn(r3.x, r3.y);
        uint4 ph, plo;
        ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
        ph.z = *reinterpret_cast<const uint32_t*>(&h2); ph.w = *reinterpret_cast<const uint32_t*>(&h3);
        plo.x = *reinterpret_cast<const uint32_t*>(&l0); plo.y = *reinterpret_cast<const uint32_t*>(&l1);
        plo.z = *reinterpret_cast<const uint32_t*>(&l2); plo.w = *reinterpret_cast<const uint32_t*>(&l3);
        *reinterpret_cast<uint4*>(ahi_row + off) = ph;
        *reinterpret_cast<uint4*>(alo_row + off) = plo;
      };
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
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kc = 0; kc < pl.KC; ++kc) {
        if (i >= 1) mbar_wait(&a_free[kc], (i - 1) & 1);   // MMA(i-1) has read this operand chunk
        TC_TRACE(2, i, 2 + 2 * kc);
        if (active) {
          const int c_end = min(nch, kc * 8 + 8);
          // this thread's chunks of the 64-feature group: c = 8 kc + h + 2 j, swizzled 16-byte slot sw[j]
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = kc * 8 + h + 2 * j;
            if (c < c_end) {
              const uint32_t off = (uint32_t)(kc * 16384) + sw[j];
              if (c < nfull) {
                const float4 v0 = *reinterpret_cast<const float4*>(hrow + c * 8);
                const float4 v1 = *reinterpret_cast<const float4*>(hrow + c * 8 + 4);
                stats4(c * 8, v0, zdA, qdA, lnA);
                stats4(c * 8 + 4, v1, zdB, qdB, lnB);
                split_store(off, v0, v1);
              } else {  // D % 8 == 4 leaves a half chunk; the rest is zero padding up to the k-step
                float4 v0 = zero4;
                if (c * 8 < D) { v0 = *reinterpret_cast<const float4*>(hrow + c * 8); stats4(c * 8, v0, zdA, qdA, lnA); }
                split_store(off, v0, zero4);
              }
            }
          }
        }
        TC_TRACE(2, i, 3 + 2 * kc);
        fence_async_smem();   // written through the generic proxy; UMMA reads via the async proxy
        mbar_arrive(&a_ready[kc]);
      }
      const float2 zd2 = __fadd2_rn(zdA, zdB), qd2 = __fadd2_rn(qdA, qdB), ln2 = __fadd2_rn(lnA, lnB);
      const float zdot = zd2.x + zd2.y, quad = qd2.x + qd2.y, lin = ln2.x + ln2.y;
      zq_s[(b2 * 2 + h) * 128 + r] = make_float4(zdot, quad, lin, 0.f);
      named_bar_sync(1, 256);
      TC_TRACE(2, i, 6);
      if (h == 0 && active) {
        const float4* zq0 = zq_s + (b2 * 2) * 128;
        const float4* zq1 = zq0 + 128;
        const float4 p0 = zq0[r], p1 = zq1[r];
        const int t = (tile_begin + i) * TB + tt;
        float g = 1.f;
        if (d.use_time_warp) {
          float z = zc;
          for (int kk = 0; kk < d.K; ++kk) z += zq0[kk * TB + tt].x + zq1[kk * TB + tt].x;
          const int tc = t < d.T ? t : d.T - 1;
          const float tz = 1.f - 2.f * rcp_approx(exp2f_approx(z * 2.885390081777927f) + 1.f);   // tanh(z)
          g = tz * window_count(tc, d.T, d.warp_type, d.window);
          if (c_col != nullptr && k == 0 && t < d.T) c_col[(size_t)n * d.T + t] = tz;
        }
        float c1, c0v;
        row_coefs(d, g, p0.z + p1.z, p0.y + p1.y, c1, c0v);
        // simi 1-3: c1 == g, second slot carries c0;  cosine: c0 == 0, second slot carries g
        st_s[b2 * 128 + r] = make_float2(c1, d.simi == 4 ? g : c0v);
      }
      mbar_arrive(&st_ready[b2]);
      TC_TRACE(2, i, 7);
    }
  } else {
    // ===== softmax warps: thread (r, h) owns columns [16h, 16h+16) of accumulator row r =====
    const int tid = threadIdx.x - 256;
    const int r = tid & 127, h = tid >> 7;
    const int wq = warp & 3;
    const bool active = r < rows;
    const int k = active ? r / TB : 0, tt = active ? r - k * TB : 0;
    const bool tanh_on = d.add_tanh && d.simi != 4;
    const bool cosine = d.simi == 4;
    // pass 2 mapping: thread u = part * D4 + d4 owns the float4 feature group d4 over the tile rows
    // [wr0, wr1) (rows are k-major: r = k * TB + tt).  A range never spans more than two modalities
    // (its length <= TB because K <= 256 / D4), hence two accumulators.
    const int D4 = D >> 2;
    const int nparts = 256 / D4;
    const bool ws_on = tid < nparts * D4;
    const int ws_d4 = tid % D4, ws_part = tid / D4;
    const int wr0 = ws_on ? (ws_part * rows) / nparts : 0;
    const int wr1 = ws_on ? ((ws_part + 1) * rows) / nparts : 0;
    const int ws_kb = (wr0 / TB + 1) * TB;                 // first row of the next modality
    const int ws_n0 = max(min(wr1, ws_kb) - wr0, 0);       // rows accumulated into o_acc[0]
    const int ws_n1 = (wr1 - wr0) - ws_n0;                 //                      into o_acc[1]
    const int ws_off = wr0 * D + ws_d4 * 4;
    float4 o_acc[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) o_acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    float l_acc = 0.f, smax = -INFINITY, ties = 0.f, am_t = 3.0e38f, am_j = 0.f;
    const float4* col4 = reinterpret_cast<const float4*>(col_s + 16 * h);
    const float4* cap4 = reinterpret_cast<const float4*>(cap_s + 16 * h);

    for (int i = 0; i < ntile; ++i) {
      const int s = i % kTcStages, b2 = i & 1;
      const int t = (tile_begin + i) * TB + tt;
      const bool in_range = active && t < d.T;
      // issue the mask load before blocking on the barriers
      uint8_t mh_u = 1;
      if (in_range && d.has_mask) mh_u = hall_mask[((size_t)n * d.K + k) * d.T + t];
      TC_TRACE(3, i, 0);
      mbar_wait(&st_ready[b2], (i >> 1) & 1);
      mbar_wait(&acc_full[b2], (i >> 1) & 1);
      TC_TRACE(3, i, 1);
      tc_fence_after();
      const float2 st = st_s[b2 * 128 + r];
      const float p_c1 = st.x, p_c0 = cosine ? 0.f : st.y, p_g = cosine ? st.y : st.x;
      const float* tile = reinterpret_cast<const float*>(ring + s * pl.stage_bytes);
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b2 * 128 + 16 * h);
      float P[16];
      {
        uint32_t x0[16], x1[16];
        tmem_ld16_issue(taddr, x0);
        tmem_ld16_issue(taddr + 32, x1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) P[j] = __uint_as_float(x0[j]) + __uint_as_float(x1[j]);
      }
      tc_fence_before();
      TC_TRACE(3, i, 2);
      const bool mh = in_range && mh_u != 0;
      // branch-free logits: P[] is overwritten by the masked logits a[].  |th| <= 1 here, so
      // th + (-1e30) == -1e30 exactly: exp_mask is a min with the per-column cap
      float m = -INFINITY;
      const float2 c1c1 = make_float2(p_c1, p_c1), c0c0 = make_float2(p_c0, p_c0);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float4 cq = col4[j >> 2], kq = cap4[j >> 2];
        const float colr0 = (j & 2) ? cq.z : cq.x, colr1 = (j & 2) ? cq.w : cq.y;
        const float capr0 = (j & 2) ? kq.z : kq.x, capr1 = (j & 2) ? kq.w : kq.y;
        float2 S = __fadd2_rn(__ffma2_rn(c1c1, make_float2(P[j], P[j + 1]), c0c0), make_float2(colr0, colr1));
        if (tanh_on) {  // tanh(x) = 1 - 2 / (1 + e^{2x}), packed
          const float2 e2 = __fmul2_rn(S, make_float2(2.885390081777927f, 2.885390081777927f));  // 2*log2(e)
          const float2 den = __fadd2_rn(make_float2(exp2f_approx(e2.x), exp2f_approx(e2.y)), make_float2(1.f, 1.f));
          S = __ffma2_rn(make_float2(rcp_approx(den.x), rcp_approx(den.y)), make_float2(-2.f, -2.f), make_float2(1.f, 1.f));
        }
        const float a0 = fminf(S.x, capr0);
        const float a1 = fminf(S.y, capr1);
        P[j] = a0; P[j + 1] = a1;
        m = fmaxf(m, fmaxf(a0, a1));
      }
      if (__any_sync(0xffffffffu, !mh) && !mh) {   // masked context row: every logit is -1e30 (padding columns stay -inf)
        m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) { P[j] = fminf(cap_s[16 * h + j], FVTA_VERY_NEGATIVE); m = fmaxf(m, P[j]); }
      }
      TC_TRACE(3, i, 3);
      mx_s[(b2 * 2 + h) * 128 + r] = m;
      if (a_logits != nullptr && in_range) {
        float* arow = a_logits + (((size_t)n * d.K + k) * d.T + t) * d.JQ + 16 * h;
        if ((d.JQ & 3) == 0) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            if (16 * h + j4 * 4 < d.JQ)
              *reinterpret_cast<float4*>(arow + j4 * 4) = make_float4(P[j4 * 4], P[j4 * 4 + 1], P[j4 * 4 + 2], P[j4 * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (16 * h + j < d.JQ) arow[j] = P[j];
        }
      }
      named_bar_sync(2, 256);
      TC_TRACE(3, i, 4);
      m = fmaxf(m, mx_s[(b2 * 2 + (h ^ 1)) * 128 + r]);   // full-row maximum
      float w = 0.f;
      // a new (or tied) running maximum is rare after the first tiles: keep the bookkeeping behind a
      // warp-uniform branch so it costs nothing when no lane needs it
      const bool newmax = in_range && m >= smax;
      if (__any_sync(0xffffffffu, newmax)) {
        if (newmax) {
          if (m > smax) { smax = m; ties = 0.f; am_t = 3.0e38f; }
          int cnt = 0, jm = -1;
#pragma unroll
          for (int j = 15; j >= 0; --j) {
            const bool hit = P[j] == m;
            cnt += hit ? 1 : 0;
            jm = hit ? j : jm;
          }
          ties += (float)cnt;
          if (jm >= 0 && (float)t < am_t) { am_t = (float)t; am_j = (float)(16 * h + jm); }
        }
      }
      if (in_range) {
        w = (m <= -1e29f) ? kMaskedWeight : expf(m - 1.f);
        if (h == 0) l_acc += w;
      }
      float* wg = wg_s + b2 * 128;
      if (h == 0) wg[r] = w * p_g;
      named_bar_sync(3, 256);
      TC_TRACE(3, i, 5);
      // pass 2: o[k][d4] += sum_rows (w*g) * h over this thread's row range, four rows per batch
      if (ws_on) {
        const float* hp = tile + ws_off;
        const float* wp = wg + wr0;
        auto wsum = [&](int nrow, float4& acc) {
          float2 oa = make_float2(acc.x, acc.y), ob = make_float2(acc.z, acc.w);
          int t2 = 0;
          for (; t2 + 4 <= nrow; t2 += 4) {
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
          for (; t2 < nrow; ++t2) {
            const float4 h0 = *reinterpret_cast<const float4*>(hp);
            const float w0 = wp[0];
            oa = __ffma2_rn(make_float2(w0, w0), make_float2(h0.x, h0.y), oa);
            ob = __ffma2_rn(make_float2(w0, w0), make_float2(h0.z, h0.w), ob);
            hp += D;
            wp += 1;
          }
          acc = make_float4(oa.x, oa.y, ob.x, ob.y);
        };
        wsum(ws_n0, o_acc[0]);
        TC_TRACE(3, i, 6);
        wsum(ws_n1, o_acc[1]);
      }
      mbar_arrive(&empty[s]);
      mbar_arrive(&ep_done[b2]);
      TC_TRACE(3, i, 7);
    }
    // ---- slab partials: per-modality reductions over the 2*TB threads owning its rows ----
    // (red_s aliases zq_s: the split warps finished their last tile before st_ready of that tile)
    red_s[tid] = l_acc; red_s[256 + tid] = smax; red_s[512 + tid] = ties; red_s[768 + tid] = am_t; red_s[1024 + tid] = am_j;
    named_bar_sync(2, 256);
    if (tid < d.K) {
      float l = 0.f, sm = -INFINITY, ti = 0.f, at = 3.0e38f, aj = 0.f;
      for (int hh = 0; hh < 2; ++hh)
        for (int t2 = 0; t2 < TB; ++t2) {
          const int rr = hh * 128 + tid * TB + t2;
          l += red_s[rr];
          const float v = red_s[256 + rr];
          if (v > sm) { sm = v; ti = 0.f; at = 3.0e38f; }
          if (v == sm) {
            ti += red_s[512 + rr];
            if (red_s[768 + rr] < at) { at = red_s[768 + rr]; aj = red_s[1024 + rr]; }
          }
        }
      float* ps = part_stat + (((size_t)n * pl.slabs + slab) * d.K + tid) * kPartW;
      ps[0] = 1.f; ps[1] = l; ps[2] = sm; ps[3] = ti; ps[4] = at; ps[5] = aj; ps[6] = 0.f; ps[7] = 0.f;
    }
    // the row ranges of one feature group live on nparts threads: stage every thread's two partial
    // sums in shared memory (over red_s, once the statistics above are out) and add them per
    // (modality, feature group) in a fixed order, so the result does not depend on scheduling
    named_bar_sync(2, 256);
    float4* stg = reinterpret_cast<float4*>(red_s);   // [256][2]
    if (ws_on) { stg[tid * 2] = o_acc[0]; stg[tid * 2 + 1] = o_acc[1]; }
    named_bar_sync(2, 256);
    for (int idx = tid; idx < d.K * D4; idx += 256) {
      const int k2 = idx / D4, d4 = idx - k2 * D4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int pp = 0; pp < nparts; ++pp) {
        const int q0 = (pp * rows) / nparts, q1 = ((pp + 1) * rows) / nparts;
        if (q1 <= q0) continue;
        const int klo = q0 / TB;
        int slot = -1;
        if (klo == k2) slot = 0;
        else if (klo + 1 == k2 && q1 > (klo + 1) * TB) slot = 1;
        if (slot >= 0) {
          const float4 o = stg[(pp * D4 + d4) * 2 + slot];
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
      }
      *reinterpret_cast<float4*>(part_o + (((size_t)n * pl.slabs + slab) * d.K + k2) * Dp + d4 * 4) = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
  }
}

#ifdef FVTA_TC_TRACE
extern "C" int fvta_tc_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 4 * kTraceTiles * 8);
}
#endif

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Returns FVTA_OK if it ran, 1 if this shape/flag combination is not covered (caller uses the
// FFMA kernel), or a negative error.
int attn_fwd_tc_try(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall, const uint8_t* hall_mask,
                    const uint8_t* hq_mask, float* ws, float* g1_all, float* a_logits, float* c_col, float* stats,
                    cudaStream_t st) {
  TcPlan pl;
  if (!tc_make_plan(d, pl)) return 1;
  if (pl.slabs > L.nt) return 1;  // partial buffers are sized for ceil(T/32) tiles per sample
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return 1;
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {(cuuint64_t)d.D, (cuuint64_t)d.T, (cuuint64_t)d.N * d.K};
  cuuint64_t gstr[2] = {(cuuint64_t)d.D * 4, (cuuint64_t)d.T * d.D * 4};
  cuuint32_t box[3] = {(cuuint32_t)d.D, (cuuint32_t)pl.TB, (cuuint32_t)d.K};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(hall), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return 1;
  __half* qB = reinterpret_cast<__half*>(ws + L.qB);
  int rc = FVTA_OK;
  if (g_attn_stages & 1) {
    attn_prep_qb<<<d.N, 256, 0, st>>>(d, ws + L.qeffT, qB, L.Dp, L.JQP, pl.KC);
    rc = check_launch("attn_prep_qb");
    if (rc) return rc;
  }
  if (!(g_attn_stages & 2)) return (g_attn_stages & 4) ? attn_combine_launch(d, ws + L.part_stat, ws + L.part_o, g1_all, stats, pl.slabs, L.Dp, st) : FVTA_OK;
  const bool has_lin = (d.simi == 1 || d.simi == 3);   // w_lin is identically zero otherwise
  auto kern = has_lin ? attn_fwd_tc<true> : attn_fwd_tc<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) { set_last_error("attn_fwd_tc smem attr", e); return FVTA_E_LAUNCH; }
  dim3 grid(pl.slabs, d.N);
  kern<<<grid, kTcThreads, pl.smem, st>>>(tmap, d, pl, hall_mask, hq_mask, ws + L.coef, qB, ws + L.col,
                                          ws + L.zconst, a_logits, c_col, ws + L.part_stat, ws + L.part_o, L.Dp,
                                          L.JQP);
  rc = check_launch("attn_fwd_tc");
  if (rc) return rc;
  if (!(g_attn_stages & 4)) return FVTA_OK;
  return attn_combine_launch(d, ws + L.part_stat, ws + L.part_o, g1_all, stats, pl.slabs, L.Dp, st);
}

}  // namespace fvta
