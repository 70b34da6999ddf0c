// Shared helpers for the FVTA sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/fvta_b200.h"

#define FVTA_VERY_NEGATIVE (-1e30f)  // utils.py:205

namespace fvta {

void set_last_error(const char* what, cudaError_t e);
void count_launch();

// Called once after every kernel launch of this library.
inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(what, e);
    return FVTA_E_LAUNCH;
  }
  return FVTA_OK;
}

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Row sums of time_indication_func's indicator matrix (model_v2.py:301-341): how many t2 a
// position t keeps.  type 1 all, 2 diagonal, 3 past (lower-tri), 4 future, 5 band(window).
__host__ __device__ inline float window_count(int t, int T, int warp_type, int window) {
  switch (warp_type) {
    case 1: return (float)T;
    case 2: return 1.f;
    case 3: return (float)(t + 1);
    case 4: return (float)(T - t);
    default: {
      int lo = t < window ? t : window;
      int hi = (T - 1 - t) < window ? (T - 1 - t) : window;
      return (float)(lo + hi + 1);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Attention workspace / stats layout (floats), shared by forward and backward.
// ---------------------------------------------------------------------------------------
constexpr int kAttnTT = 32;      // context positions per CTA tile (one warp lane per position)
constexpr int kAttnMaxDC = 128;  // D is streamed through shared memory in chunks of <= 128
constexpr int kStatW = 8;        // floats per (n,k) in stats: M, L, s, ties, beta, argmax t, argmax j, pad
constexpr int kPartW = 8;        // floats per tile partial: M, l, s, ties, argmax t, argmax j, pad

struct AttnLayout {
  int nt;    // tiles per sample = ceil(T / kAttnTT)
  int npart; // partial slots per sample the workspace holds: max(nt, CTAs per sample of the tcgen05 kernel)
  int JQP;   // JQ padded to the accumulator width the kernel is instantiated for
  int Dp;    // D padded to 4
  // forward scratch
  size_t coef;       // [3][Dp]      u, w_lin, w_quad
  size_t qeffT;      // [N][Dp][JQP] effective question operand, transposed (d-major)
  size_t col;        // [N][JQP]     per-question-word term + bias
  size_t zconst;     // [N]          K*((b_WH+lq).W_WC + b_WC)
  size_t part_stat;  // [N][npart][K][kPartW] tile partials: M, l, smax, ties, argmax (t, j)
  size_t part_o;     // [N][npart][K][Dp] tile partial weighted sums
  // backward scratch
  size_t dqeffT;     // [N][Dp][JQP]
  size_t dcol;       // [N][JQP]
  size_t dcoef;      // [N][3][Dp]   per-sample partial d(u, w_lin, w_quad)
  size_t dzc;        // [N]          d zconst
  size_t drow;       // [N][K][4]    per-(n,k) scalars the backward needs (ds, rdotdr, ...)
  size_t du;         // [Dp]         d u summed over samples
  size_t scal;       // [4]          sum_n dzc, d att_b
  size_t qB32;       // [N][ceil(D/32)][64 rows][128 B]: tf32 [B_hi ; B_lo] image of the tensor-core kernel
  size_t qB16;       // [N][ceil(D/64)][32 rows][128 B]: fp16 B_hi image
  size_t counter;    // [N] int: CTAs of the sample that have delivered their partials (self-resetting)
  size_t total;      // floats
};

__host__ __device__ inline int attn_jqp(int JQ) {
  // accumulator widths the kernels are instantiated for (in units of 4): 1,2,4,7,8,16
  if (JQ <= 4) return 4;
  if (JQ <= 8) return 8;
  if (JQ <= 16) return 16;
  if (JQ <= 28) return 28;
  if (JQ <= 32) return 32;
  if (JQ <= 64) return 64;
  return -1;
}

__host__ __device__ inline AttnLayout attn_layout(const FvtaAttnDesc& d) {
  AttnLayout L;
  L.nt = (d.T + kAttnTT - 1) / kAttnTT;
  {  // the tensor-core kernel runs min(tiles, ceil(148 / N)) CTAs per sample with tiles <= ceil(T / 4)
    const int per_sample = (148 + d.N - 1) / d.N, max_tiles = (d.T + 3) / 4;
    const int slabs = per_sample < max_tiles ? per_sample : max_tiles;
    L.npart = L.nt > slabs ? L.nt : slabs;
  }
  L.JQP = attn_jqp(d.JQ);
  L.Dp = round_up(d.D, 4);
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += round_up_sz(n, 4); return r; };
  L.coef = take((size_t)3 * L.Dp);
  L.qeffT = take((size_t)d.N * L.Dp * L.JQP);
  L.col = take((size_t)d.N * L.JQP);
  L.zconst = take((size_t)d.N);
  L.part_stat = take((size_t)d.N * L.npart * d.K * kPartW);
  L.part_o = take((size_t)d.N * L.npart * d.K * L.Dp);
  L.dqeffT = take((size_t)d.N * L.Dp * L.JQP);
  L.dcol = take((size_t)d.N * L.JQP);
  L.dcoef = take((size_t)d.N * 3 * L.Dp);
  L.dzc = take((size_t)d.N);
  L.drow = take((size_t)d.N * d.K * 4);
  L.du = take((size_t)L.Dp);
  L.scal = take(4);
  L.qB32 = take((size_t)d.N * ((d.D + 31) / 32) * 2048);
  L.qB16 = take((size_t)d.N * ((d.D + 63) / 64) * 1024);
  L.counter = take((size_t)d.N);
  L.total = o;
  return L;
}

// stats: [N][K][kStatW] then r [N][K][D]
__host__ __device__ inline size_t attn_stats_r_offset(const FvtaAttnDesc& d) {
  return (size_t)d.N * d.K * kStatW;
}

inline int attn_validate(const FvtaAttnDesc* d) {
  if (!d) return FVTA_E_BADARG;
  if (d->N <= 0 || d->K <= 0 || d->T <= 0 || d->D <= 0 || d->JQ <= 0) return FVTA_E_BADARG;
  if (d->simi < 1 || d->simi > 4) return FVTA_E_UNSUPPORTED;
  if (d->use_time_warp && (d->warp_type < 1 || d->warp_type > 5)) return FVTA_E_UNSUPPORTED;
  if (d->K > 8 || attn_jqp(d->JQ) < 0) return FVTA_E_SHAPE;  // block = K warps <= 256 threads
  return FVTA_OK;
}

// chunking of D through shared memory
struct ChunkPlan {
  int DC;      // chunk width (multiple of 4)
  int stride;  // smem row stride in floats, stride % 8 == 4 (conflict-free float4 row reads)
  int nchunk;
};
__host__ __device__ inline ChunkPlan chunk_plan(int Dp) {
  ChunkPlan p;
  p.DC = Dp < kAttnMaxDC ? Dp : kAttnMaxDC;
  p.stride = (p.DC % 8 == 4) ? p.DC : p.DC + 4;
  p.nchunk = (Dp + p.DC - 1) / p.DC;
  return p;
}

}  // namespace fvta
