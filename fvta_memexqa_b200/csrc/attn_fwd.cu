// Focal visual-text attention, forward (fp32, sm_100a).
//
// Replaces the TensorFlow ops of model_v2.py:953-1009 (time warp), :301-341
// (time_indication_func), :210-298 (attention_3d), :23-48 (softmax/softsel) and
// utils.py:210-213 (exp_mask) with three launches:
//
//   attn_prep_coef / attn_prep_q : O(D*D + N*JQ*D)  parameter folding
//   attn_tile_fwd                : one pass over hall -- gate, similarity, mask, tile softmax,
//                                  tile weighted sum                     <- the HBM-bound kernel
//   attn_combine_stats / _o      : merge tile partials, cross-sequence softmax, g1_all
//
// Math (as written in the reference, including its t1==t2 time-warp degeneracy):
//   z[n,t]  = sum_k u.h[n,k,t]^2 + K*((b_WH+lq[n]).W_WC + b_WC),  u = W_WH[:D] @ W_WC
//   g[n,t]  = tanh(z) * cnt(t)                      (1 when --use_time_warp is off)
//   S[t,j]  = c1*P[t,j] + c0 + col[j],  P = h . qeff_j
//             simi 1-3: c1 = g, c0 = g*(w_lin.h) + g^2*(w_quad.h^2);  simi 4: cosine
//   a       = (tanh)(S) + (1-mask)*(-1e30)
//   m[t]    = max_j a  (* g when --use_time_warp_att);  alpha = softmax_t m;  r_k = sum alpha*g*h
//   s_k     = max_{t,j} a;  beta = softmax_k s;  g1_all = sum_k beta_k r_k
#include "common.cuh"
#include "attn_tile.cuh"
#include "attn_tc_images.cuh"

namespace fvta {

// ---------------------------------------------------------------------------------------
// prep: coefficient vectors shared by all samples
// coef[0] = u (zeros without time warp), coef[1] = w_lin, coef[2] = w_quad (ones for cosine)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void prep_coef_block(const FvtaAttnDesc& d, const FvtaAttnParams& p, float* __restrict__ coef,
                                                int Dp, int block) {
  int warp = (block * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= Dp) return;
  int dd = warp;
  float u = 0.f, wl = 0.f, wq = 0.f;
  if (dd < d.D) {
    if (d.use_time_warp) {
      float acc = 0.f;
      for (int e = lane; e < d.D; e += 32) acc += p.WH_W[(size_t)dd * d.D + e] * p.WC_W[e];
      u = warp_sum(acc);
    }
    const int D = d.D;
    if (d.simi == 1) { wl = p.att_W[dd]; }
    else if (d.simi == 2) { wq = p.att_W[D + dd]; }
    else if (d.simi == 3) { wl = p.att_W[dd]; wq = p.att_W[2 * D + dd]; }
    else { wq = 1.f; }
  }
  if (lane == 0) {
    coef[dd] = u;
    coef[Dp + dd] = wl;
    coef[2 * Dp + dd] = wq;
  }
}

// ---------------------------------------------------------------------------------------
// prep: per-sample question operand (transposed, zero padded), column term, gate constant
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void prep_q_block(const FvtaAttnDesc& d, const FvtaAttnParams& p, const float* __restrict__ hq,
                                             const float* __restrict__ lq, float* __restrict__ qeffT,
                                             float* __restrict__ col, float* __restrict__ zconst, int Dp, int JQP,
                                             int n, float* scale /* smem [JQP]: cosine scale per question word */,
                                             float* qeff_s /* smem [Dp][JQP] or null */, bool write_global) {
  const int D = d.D, JQ = d.JQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int jy = 0, jstep = 1;
  const float* q = hq + (size_t)n * JQ * D;
  const float* wq_lin = nullptr;
  const float* wq_quad = nullptr;
  if (d.simi == 1) wq_lin = p.att_W + D;
  else if (d.simi == 2) wq_quad = p.att_W + D;
  else if (d.simi == 3) { wq_lin = p.att_W + D; wq_quad = p.att_W + 2 * D; }
  for (int j = jy + warp * jstep; j < JQP; j += nwarp * jstep) {
    float qq = 0.f, c = 0.f;
    if (j < JQ) {
      for (int e = lane; e < D; e += 32) {
        float v = q[(size_t)j * D + e];
        qq += v * v;
        if (wq_lin) c += wq_lin[e] * v;
        if (wq_quad) c += wq_quad[e] * v * v;
      }
    }
    qq = warp_sum(qq);
    c = warp_sum(c);
    if (lane == 0) {
      scale[j] = rsqrtf(fmaxf(qq, 1e-12f));  // tf.nn.l2_normalize epsilon
      float bias = (d.simi == 4) ? 0.f : p.att_b[0];
      if (write_global) col[(size_t)n * JQP + j] = (j < JQ && d.simi != 4) ? c + bias : 0.f;
    }
  }
  if (warp == 0 && jy == 0 && write_global) {
    float acc = 0.f;
    if (d.use_time_warp)
      for (int e = lane; e < D; e += 32) acc += (p.WH_b[e] + lq[(size_t)n * D + e]) * p.WC_W[e];
    acc = warp_sum(acc);
    if (lane == 0) zconst[n] = d.use_time_warp ? (float)d.K * (acc + p.WC_b[0]) : 0.f;
  }
  __syncthreads();
  float* out = qeffT + (size_t)n * Dp * JQP;
  const int nj = (JQP - jy + jstep - 1) / jstep;   // question words owned by this block
  for (int i = threadIdx.x; i < Dp * nj; i += blockDim.x) {
    int dd = i / nj, j = jy + (i - dd * nj) * jstep;
    float v = 0.f;
    if (dd < D && j < JQ) {
      float qv = q[(size_t)j * D + dd];
      if (d.simi == 1) v = qv * p.att_W[2 * D + dd];
      else if (d.simi == 2) v = qv * (p.att_W[dd] - 2.f * p.att_W[D + dd]);
      else if (d.simi == 3) v = qv * (p.att_W[3 * D + dd] - 2.f * p.att_W[2 * D + dd]);
      else v = qv * scale[j];
    }
    if (write_global) out[(size_t)dd * JQP + j] = v;
    if (qeff_s != nullptr) qeff_s[dd * JQP + j] = v;
  }
}

// ---------------------------------------------------------------------------------------
// the tile kernel: grid (nt, N), block K*32 threads, thread = (k = warp, tt = lane)
// ---------------------------------------------------------------------------------------
template <int JQ4>
__global__ void __launch_bounds__(256) attn_tile_fwd(
    FvtaAttnDesc d, const float* __restrict__ hall, const uint8_t* __restrict__ hall_mask,
    const uint8_t* __restrict__ hq_mask, const float* __restrict__ coef, const float* __restrict__ qeffT,
    const float* __restrict__ col, const float* __restrict__ zconst, float* __restrict__ a_logits,
    float* __restrict__ c_col, float* __restrict__ warp_h, float* __restrict__ part_stat,
    float* __restrict__ part_o, int nt, int Dp) {
  constexpr int JQP = 4 * JQ4;
  extern __shared__ __align__(16) float smem[];
  const ChunkPlan cp = chunk_plan(Dp);
  const int rows = d.K * kAttnTT;
  float* h_s = smem;                                  // [rows][stride]
  float* q_s = h_s + (size_t)rows * cp.stride;        // [DC][JQP]
  float* coef_s = q_s + (size_t)cp.DC * JQP;          // [kCoefRows][DC]
  float* zd_s = coef_s + kCoefRows * cp.DC;                   // [K][32] row gate dots
  float* wg_s = zd_s + rows;                          // [K][32] softmax weight * gate
  float* col_s = wg_s + rows;                         // [JQP]
  float* qm_s = col_s + JQP;                          // [JQP] question mask as 0/1

  const int tile = blockIdx.x, n = blockIdx.y;
  const int t0 = tile * kAttnTT;
  const int k = threadIdx.x >> 5, tt = threadIdx.x & 31;
  const int t = t0 + tt;
  const bool in_range = t < d.T;

  for (int j = threadIdx.x; j < JQP; j += blockDim.x) {
    col_s[j] = col[(size_t)n * JQP + j];
    qm_s[j] = (j < d.JQ && (!d.has_mask || hq_mask[(size_t)n * d.JQ + j])) ? 1.f : 0.f;
  }

  float acc[JQP];
#pragma unroll
  for (int j = 0; j < JQP; ++j) acc[j] = 0.f;
  float zdot = 0.f, lin = 0.f, quad = 0.f, hg_unused = 0.f;

  const float* qeffT_n = qeffT + (size_t)n * Dp * JQP;
  for (int c = 0; c < cp.nchunk; ++c) {
    const int c0 = c * cp.DC;
    const int DCc = min(cp.DC, Dp - c0);
    if (c > 0) __syncthreads();
    load_chunk(d, hall, n, t0, c0, DCc, cp.stride, h_s, qeffT_n, q_s, JQP, coef, coef_s, Dp, cp.DC);
    row_chunk_dot<JQ4, false>(h_s + (size_t)threadIdx.x * cp.stride, q_s, coef_s, DCc, cp.DC, acc, zdot, lin, quad, hg_unused);
  }

  // ---- gate -------------------------------------------------------------------------
  float g = 1.f;
  if (d.use_time_warp) {
    zd_s[threadIdx.x] = zdot;
    __syncthreads();
    float z = zconst[n];
    for (int kk = 0; kk < d.K; ++kk) z += zd_s[kk * kAttnTT + tt];
    float tz, cnt;
    g = gate_from_z(d, z, t, tz, cnt);
    if (c_col != nullptr && k == 0 && in_range) c_col[(size_t)n * d.T + t] = tz;
  }

  // ---- logits, row max --------------------------------------------------------------
  float c1, c0v;
  row_coefs(d, g, lin, quad, c1, c0v);
  const bool mh = in_range && (!d.has_mask || hall_mask[((size_t)n * d.K + k) * d.T + t]);
  const bool tanh_on = d.add_tanh && d.simi != 4;
  float m = -INFINITY;
  int rowcnt = 0, jm = 0;
  float* arow = (a_logits != nullptr && in_range)
                    ? a_logits + (((size_t)n * d.K + k) * d.T + t) * d.JQ : nullptr;
#pragma unroll
  for (int j = 0; j < JQP; ++j) {
    if (j < d.JQ) {
      float th;
      const float a = row_logit(c1, acc[j], c0v, col_s[j], tanh_on, d.has_mask, mh && qm_s[j] != 0.f, th);
      if (arow) arow[j] = a;
      if (a > m) { m = a; rowcnt = 1; jm = j; }
      else if (a == m) { rowcnt++; }
    }
  }

  // ---- tile statistics per modality (warp = k, lane = tt) ---------------------------
  float smax_in = in_range ? m : -INFINITY;
  float smax = warp_max(smax_in);
  float ties = warp_sum((in_range && m == smax) ? (float)rowcnt : 0.f);
  // position of the (first) maximum: lowest lane holding it
  const unsigned hit = __ballot_sync(0xffffffffu, in_range && m == smax);
  const int src = hit ? __ffs(hit) - 1 : 0;
  const float am_t = (float)(t0 + src);
  const float am_j = (float)__shfl_sync(0xffffffffu, jm, src);
  float ml = d.time_warp_att ? m * g : m;  // model_v2.py:269-275
  float mloc = in_range ? ml : -INFINITY;
  float M = warp_max(mloc);
  float w = in_range ? expf(ml - M) : 0.f;
  float l = warp_sum(w);
  wg_s[threadIdx.x] = w * g;
  if (tt == 0) {
    float* ps = part_stat + (((size_t)n * nt + tile) * d.K + k) * kPartW;
    ps[0] = M; ps[1] = l; ps[2] = smax; ps[3] = ties; ps[4] = am_t; ps[5] = am_j; ps[6] = 0.f; ps[7] = 0.f;
  }
  __syncwarp();

  // ---- tile weighted sum: o_k[:] = sum_tt w*g*h[k,tt,:]  (rows re-read through L2) ---
  const int tvalid = min(kAttnTT, d.T - t0);
  const float* hk = hall + (((size_t)n * d.K + k) * d.T + t0) * d.D;
  float* po = part_o + (((size_t)n * nt + tile) * d.K + k) * Dp;
  const float* wgk = wg_s + k * kAttnTT;
  if ((d.D & 3) == 0) {
    for (int c4 = tt; c4 < (d.D >> 2); c4 += 32) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < tvalid; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hk + (size_t)r * d.D + c4 * 4);
        const float ww = wgk[r];
        o.x = fmaf(ww, hv.x, o.x); o.y = fmaf(ww, hv.y, o.y);
        o.z = fmaf(ww, hv.z, o.z); o.w = fmaf(ww, hv.w, o.w);
      }
      *reinterpret_cast<float4*>(po + c4 * 4) = o;
    }
  } else {
    for (int c = tt; c < d.D; c += 32) {
      float o = 0.f;
      for (int r = 0; r < tvalid; ++r) o = fmaf(wgk[r], hk[(size_t)r * d.D + c], o);
      po[c] = o;
    }
  }
  if (warp_h != nullptr) {  // model.warp_h (visualisation only)
    float* whk = warp_h + (((size_t)n * d.K + k) * d.T + t0) * d.D;
    for (int r = 0; r < tvalid; ++r) {
      // gate of row r lives in lane r of this warp
      const float gr = __shfl_sync(0xffffffffu, g, r);
      for (int c = tt; c < d.D; c += 32) whk[(size_t)r * d.D + c] = gr * hk[(size_t)r * d.D + c];
    }
  }
}

// ---------------------------------------------------------------------------------------
// combine, stage 1: grid N, block K warps.  Per modality: global max / sum of the tile
// softmax partials, s_k and its tie count, then the cross-sequence softmax beta.
// ---------------------------------------------------------------------------------------
__global__ void attn_combine_stats(FvtaAttnDesc d, const float* __restrict__ part_stat,
                                   float* __restrict__ stats, int nt) {
  __shared__ float s_s[32];
  const int n = blockIdx.x;
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* ps = part_stat + (size_t)n * nt * d.K * kPartW;
  float M = -INFINITY, s = -INFINITY;
  for (int i = lane; i < nt; i += 32) {
    const float* e = ps + ((size_t)i * d.K + k) * kPartW;
    M = fmaxf(M, e[0]);
    s = fmaxf(s, e[2]);
  }
  M = warp_max(M);
  s = warp_max(s);
  float L = 0.f, ties = 0.f;
  int first = 0x7fffffff;  // lowest tile holding the maximum -> its recorded position
  for (int i = lane; i < nt; i += 32) {
    const float* e = ps + ((size_t)i * d.K + k) * kPartW;
    L += e[1] * expf(e[0] - M);
    if (e[2] == s) { ties += e[3]; first = min(first, i); }
  }
  L = warp_sum(L);
  ties = warp_sum(ties);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  const float* ef = ps + ((size_t)(first < nt ? first : 0) * d.K + k) * kPartW;
  const float am_t = ef[4], am_j = ef[5];
  if (lane == 0) s_s[k] = s;
  __syncthreads();
  float smax = -INFINITY;
  for (int kk = 0; kk < d.K; ++kk) smax = fmaxf(smax, s_s[kk]);
  float den = 0.f;
  for (int kk = 0; kk < d.K; ++kk) den += expf(s_s[kk] - smax);
  const float beta = expf(s - smax) / den;
  if (lane == 0) {
    float* st = stats + ((size_t)n * d.K + k) * kStatW;
    st[0] = M; st[1] = L; st[2] = s; st[3] = ties; st[4] = beta; st[5] = am_t; st[6] = am_j; st[7] = 0.f;
  }
}

// combine, stage 2: grid (ceil(D/32), N, K), block 32 x 8.  r_k[:] = sum_tiles o * exp(M_tile - M) / L
// (lane = feature, the 8 warps stride over the tiles), then g1_all += beta_k r_k (g1_all zeroed first).
__global__ void attn_combine_o(FvtaAttnDesc d, const float* __restrict__ part_stat,
                               const float* __restrict__ part_o, float* __restrict__ g1_all,
                               float* __restrict__ stats, int nt, int Dp) {
  __shared__ float red[8][33];
  const int n = blockIdx.y, k = blockIdx.z;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* ps = part_stat + (size_t)n * nt * d.K * kPartW;
  const float* st = stats + ((size_t)n * d.K + k) * kStatW;
  const float M = st[0], invL = 1.f / st[1], beta = st[4];
  float o = 0.f;
  if (c < d.D)
    for (int i = w; i < nt; i += 8) {
      const float sc = expf(ps[((size_t)i * d.K + k) * kPartW] - M);
      o = fmaf(part_o[(((size_t)n * nt + i) * d.K + k) * Dp + c], sc, o);
    }
  red[w][lane] = o;
  __syncthreads();
  if (w == 0 && c < d.D) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i][lane];
    r *= invL;
    stats[attn_stats_r_offset(d) + ((size_t)n * d.K + k) * d.D + c] = r;
    atomicAdd(g1_all + (size_t)n * d.D + c, beta * r);
  }
}

int attn_combine_launch(const FvtaAttnDesc& d, const float* part_stat, const float* part_o, float* g1_all,
                        float* stats, int nt, int Dp, cudaStream_t st) {
  attn_combine_stats<<<d.N, d.K * 32, 0, st>>>(d, part_stat, stats, nt);
  int rc = check_launch("attn_combine_stats");
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(g1_all, 0, (size_t)d.N * d.D * sizeof(float), st);
  if (e != cudaSuccess) { set_last_error("attn_combine memset", e); return FVTA_E_LAUNCH; }
  dim3 grid((d.D + 31) / 32, d.N, d.K);
  attn_combine_o<<<grid, 256, 0, st>>>(d, part_stat, part_o, g1_all, stats, nt, Dp);
  return check_launch("attn_combine_o");
}

size_t attn_tile_smem_bytes(const FvtaAttnDesc& d, int JQP, int Dp) {
  const ChunkPlan cp = chunk_plan(Dp);
  const int rows = d.K * kAttnTT;
  size_t fl = (size_t)rows * cp.stride + (size_t)cp.DC * JQP + kCoefRows * cp.DC + 2 * rows + 2 * JQP;
  return fl * sizeof(float);
}

template <int JQ4>
static int launch_tile_fwd(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall,
                           const uint8_t* hall_mask, const uint8_t* hq_mask, float* ws, float* a_logits,
                           float* c_col, float* warp_h, cudaStream_t st) {
  size_t smem = attn_tile_smem_bytes(d, 4 * JQ4, L.Dp);
  cudaError_t e = cudaFuncSetAttribute(attn_tile_fwd<JQ4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_last_error("attn_tile_fwd smem attr", e); return FVTA_E_LAUNCH; }
  dim3 grid(L.nt, d.N);
  attn_tile_fwd<JQ4><<<grid, d.K * kAttnTT, smem, st>>>(
      d, hall, hall_mask, hq_mask, ws + L.coef, ws + L.qeffT, ws + L.col, ws + L.zconst, a_logits, c_col,
      warp_h, ws + L.part_stat, ws + L.part_o, L.nt, L.Dp);
  return check_launch("attn_tile_fwd");
}

bool attn_tc_covers(const FvtaAttnDesc& d);

// Operand folding in ONE launch (the three small kernels it replaces cost 17 us of launch latency
// next to a 100 us tile kernel): blocks [0, N) fold sample n's question operand, column term and gate
// constant and -- for the tensor-core kernel -- build its B images and reset its CTA counter; the
// remaining blocks fold the coefficient vectors shared by all samples.
constexpr int kPrepSplit = 4;   // blocks per sample: each folds the question operand (cheap) and builds a quarter of the images

__global__ void __launch_bounds__(256) attn_prep_all(FvtaAttnDesc d, FvtaAttnParams p, const float* __restrict__ hq,
                                                      const float* __restrict__ lq, float* __restrict__ coef,
                                                      float* __restrict__ qeffT, float* __restrict__ col,
                                                      float* __restrict__ zconst, uint8_t* __restrict__ qB32,
                                                      uint8_t* __restrict__ qB16, int* __restrict__ counters, int Dp,
                                                      int JQP, int tc_images) {
  extern __shared__ float prep_sm[];
  float* scale = prep_sm;          // [64]
  float* qeff_s = prep_sm + 64;    // [Dp][JQP] (tensor-core path only)
  const int nsb = tc_images ? d.N * kPrepSplit : d.N;
  if ((int)blockIdx.x >= nsb) {
    prep_coef_block(d, p, coef, Dp, blockIdx.x - nsb);
    return;
  }
  const int n = tc_images ? blockIdx.x / kPrepSplit : blockIdx.x;
  const int part = tc_images ? blockIdx.x % kPrepSplit : 0;
  prep_q_block(d, p, hq, lq, qeffT, col, zconst, Dp, JQP, n, scale, tc_images ? qeff_s : nullptr, part == 0);
  if (!tc_images) return;
  __syncthreads();
  const int C32 = (d.D + 31) / 32, C16 = (d.D + 63) / 64;
  build_tc_images_block(d, qeff_s, qB32 + (size_t)n * C32 * 8192, qB16 + (size_t)n * C16 * 4096, Dp, JQP, part, kPrepSplit);
  if (threadIdx.x == 0 && part == 0) counters[n] = 0;
}

int attn_prep_launch(const FvtaAttnDesc& d, const AttnLayout& L, const FvtaAttnParams& p, const float* hq,
                     const float* lq, float* ws, cudaStream_t st, bool tc_images) {
  const int coef_blocks = (L.Dp * 32 + 255) / 256;
  const int sample_blocks = tc_images ? d.N * kPrepSplit : d.N;
  const size_t smem = (64 + (tc_images ? (size_t)L.Dp * L.JQP : 0)) * sizeof(float);
  attn_prep_all<<<sample_blocks + coef_blocks, 256, smem, st>>>(d, p, hq, lq, ws + L.coef, ws + L.qeffT, ws + L.col, ws + L.zconst,
                                                   reinterpret_cast<uint8_t*>(ws + L.qB32),
                                                   reinterpret_cast<uint8_t*>(ws + L.qB16),
                                                   reinterpret_cast<int*>(ws + L.counter), L.Dp, L.JQP, tc_images ? 1 : 0);
  return check_launch("attn_prep_all");
}

int attn_fwd_tc_launch(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall, const uint8_t* hall_mask,
                       const uint8_t* hq_mask, float* ws, float* g1_all, float* a_logits, float* c_col, float* stats,
                       cudaStream_t st);
bool attn_small_ok(const FvtaAttnDesc& d);

// The one place that decides which forward kernel serves a descriptor (fvta_attn_fwd and the
// fvta_attn_fwd_path query both call it).
int attn_pick_path(const FvtaAttnDesc& d, bool want_c_col, bool want_warp_h) {
  if (d.path != FVTA_PATH_AUTO && d.path != FVTA_PATH_FFMA) return FVTA_E_BADARG;
  if (d.path == FVTA_PATH_FFMA) return FVTA_PATH_FFMA;
  if (attn_small_ok(d) && !want_warp_h && !want_c_col) return FVTA_PATH_SMALL;
  if (!want_warp_h && attn_tc_covers(d)) return FVTA_PATH_TC;
  return FVTA_PATH_FFMA;
}

}  // namespace fvta

using namespace fvta;

extern "C" size_t fvta_attn_workspace_bytes(const FvtaAttnDesc* desc) {
  if (attn_validate(desc) != FVTA_OK) return 0;
  return attn_layout(*desc).total * sizeof(float);
}

extern "C" size_t fvta_attn_stats_floats(const FvtaAttnDesc* desc) {
  if (attn_validate(desc) != FVTA_OK) return 0;
  return (size_t)desc->N * desc->K * (kStatW + desc->D);
}

namespace fvta {
int attn_small_fwd(const FvtaAttnDesc& d, const float* hall, const uint8_t* hall_mask, const float* hq,
                   const uint8_t* hq_mask, const FvtaAttnParams& P, float* g1_all, float* a_logits, cudaStream_t st);
}  // namespace fvta

extern "C" int fvta_attn_fwd_path(const FvtaAttnDesc* desc, int want_c_col, int want_warp_h) {
  int rc = attn_validate(desc);
  if (rc) return rc;
  return attn_pick_path(*desc, want_c_col != 0, want_warp_h != 0);
}

extern "C" int fvta_attn_fwd(const FvtaAttnDesc* desc, const float* hall, const uint8_t* hall_mask,
                             const float* hq, const uint8_t* hq_mask, const float* lq,
                             const FvtaAttnParams* params, float* g1_all, float* a_logits, float* c_col,
                             float* warp_h, float* stats, void* workspace, size_t workspace_bytes,
                             void* stream) {
  int rc = attn_validate(desc);
  if (rc) return rc;
  const FvtaAttnDesc d = *desc;
  if (!hall || !hq || !params || !g1_all || !stats || !workspace) return FVTA_E_BADARG;
  if (d.has_mask && (!hall_mask || !hq_mask)) return FVTA_E_BADARG;
  if (d.use_time_warp && (!lq || !params->WH_W || !params->WH_b || !params->WC_W || !params->WC_b))
    return FVTA_E_BADARG;
  if (d.simi != 4 && (!params->att_W || !params->att_b)) return FVTA_E_BADARG;
  const AttnLayout L = attn_layout(d);
  if (workspace_bytes < L.total * sizeof(float)) return FVTA_E_WORKSPACE;
  if (attn_tile_smem_bytes(d, L.JQP, L.Dp) > 227 * 1024) return FVTA_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int path = attn_pick_path(d, c_col != nullptr, warp_h != nullptr);
  if (path < 0) return path;
  // question-attention shape (K = 1, JQ = 1, no time warp): one CTA per sample does everything
  if (path == FVTA_PATH_SMALL) return attn_small_fwd(d, hall, hall_mask, hq, hq_mask, *params, g1_all, a_logits, st);
  float* ws = (float*)workspace;
  if (!d.ws_from_fwd) {
    rc = attn_prep_launch(d, L, *params, hq, lq, ws, st, path == FVTA_PATH_TC);
    if (rc) return rc;
  }
  if (path == FVTA_PATH_TC)
    return attn_fwd_tc_launch(d, L, hall, hall_mask, hq_mask, ws, g1_all, a_logits, c_col, stats, st);
  switch (L.JQP) {
    case 4: rc = launch_tile_fwd<1>(d, L, hall, hall_mask, hq_mask, ws, a_logits, c_col, warp_h, st); break;
    case 8: rc = launch_tile_fwd<2>(d, L, hall, hall_mask, hq_mask, ws, a_logits, c_col, warp_h, st); break;
    case 16: rc = launch_tile_fwd<4>(d, L, hall, hall_mask, hq_mask, ws, a_logits, c_col, warp_h, st); break;
    case 28: rc = launch_tile_fwd<7>(d, L, hall, hall_mask, hq_mask, ws, a_logits, c_col, warp_h, st); break;
    case 32: rc = launch_tile_fwd<8>(d, L, hall, hall_mask, hq_mask, ws, a_logits, c_col, warp_h, st); break;
    case 64: rc = launch_tile_fwd<16>(d, L, hall, hall_mask, hq_mask, ws, a_logits, c_col, warp_h, st); break;
    default: return FVTA_E_SHAPE;
  }
  if (rc) return rc;
  return attn_combine_launch(d, ws + L.part_stat, ws + L.part_o, g1_all, stats, L.nt, L.Dp, st);
}
