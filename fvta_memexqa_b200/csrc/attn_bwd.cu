// Focal visual-text attention, backward (fp32, sm_100a), hand-derived.
//
//   attn_bwd_prep   : per (n,k) scalars from the saved softmax statistics
//   attn_tile_bwd   : recomputes gate + logits of a tile with the forward's instruction
//                     sequence, routes the gradient through both softmaxes / both maxima
//                     (ties split equally, like tf.reduce_max), writes d_hall, accumulates the
//                     question-side and coefficient gradients
//   attn_bwd_q      : d_hq, d_lq from the accumulated question-side gradients
//   attn_bwd_params : d att_W/b, d u            attn_bwd_warp : d WH_W/b, d WC_W/b
//
// The logits are recomputed rather than saved: saving a_logits costs 4*N*K*T*JQ bytes written
// and read back; recomputing costs one more pass of the (cheap) contraction.
#include "common.cuh"
#include "attn_tile.cuh"

namespace fvta {

int attn_prep_launch(const FvtaAttnDesc& d, const AttnLayout& L, const FvtaAttnParams& p, const float* hq,
                     const float* lq, float* ws, cudaStream_t st, bool tc_images);

// drow[n][k] = { ds_k, beta_k * (r_k . ghat), beta_k, 0 }
__global__ void attn_bwd_prep(FvtaAttnDesc d, const float* __restrict__ stats, const float* __restrict__ dg1,
                              float* __restrict__ drow) {
  __shared__ float dbeta_s[32];
  const int n = blockIdx.x;
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* r = stats + attn_stats_r_offset(d) + ((size_t)n * d.K + k) * d.D;
  float acc = 0.f;
  for (int c = lane; c < d.D; c += 32) acc = fmaf(r[c], dg1[(size_t)n * d.D + c], acc);
  acc = warp_sum(acc);
  if (lane == 0) dbeta_s[k] = acc;
  __syncthreads();
  if (lane == 0) {
    float dot = 0.f;
    for (int kk = 0; kk < d.K; ++kk) dot += stats[((size_t)n * d.K + kk) * kStatW + 4] * dbeta_s[kk];
    const float beta = stats[((size_t)n * d.K + k) * kStatW + 4];
    float* o = drow + ((size_t)n * d.K + k) * 4;
    o[0] = beta * (acc - dot);  // d s_k  (softmax over K)
    o[1] = beta * acc;          // sum_t alpha_t dalpha_t = r_k . d r_k
    o[2] = beta;
    o[3] = 0.f;
  }
}

template <int JQ4>
__global__ void __launch_bounds__(256) attn_tile_bwd(
    FvtaAttnDesc d, const float* __restrict__ hall, const uint8_t* __restrict__ hall_mask,
    const uint8_t* __restrict__ hq_mask, const float* __restrict__ coef, const float* __restrict__ qeffT,
    const float* __restrict__ col, const float* __restrict__ zconst, const float* __restrict__ stats,
    const float* __restrict__ drow, const float* __restrict__ dg1, float* __restrict__ d_hall,
    float* __restrict__ dqeffT, float* __restrict__ dcol, float* __restrict__ dcoef, float* __restrict__ dzc,
    int Dp) {
  constexpr int JQP = 4 * JQ4;
  extern __shared__ __align__(16) float smem[];
  const ChunkPlan cp = chunk_plan(Dp);
  const int rows = d.K * kAttnTT;
  float* h_s = smem;                                  // [rows][stride]
  float* q_s = h_s + (size_t)rows * cp.stride;        // [DC][JQP]
  float* coef_s = q_s + (size_t)cp.DC * JQP;          // [kCoefRows][DC]
  float* dS_s = coef_s + kCoefRows * cp.DC;           // [rows][JQP]
  float* zd_s = dS_s + (size_t)rows * JQP;            // [rows] gate dots, then per-row d g
  float* c1_s = zd_s + rows;                          // [rows]
  float* dlin_s = c1_s + rows;                        // [rows]
  float* dquad_s = dlin_s + rows;                     // [rows]
  float* dz_s = dquad_s + rows;                       // [32]
  float* col_s = dz_s + kAttnTT;                      // [JQP]
  float* qm_s = col_s + JQP;                          // [JQP]

  const int tile = blockIdx.x, n = blockIdx.y;
  const int t0 = tile * kAttnTT;
  const int k = threadIdx.x >> 5, tt = threadIdx.x & 31;
  const int t = t0 + tt;
  const bool in_range = t < d.T;
  const float* gvec = dg1 + (size_t)n * d.D;

  for (int j = threadIdx.x; j < JQP; j += blockDim.x) {
    col_s[j] = col[(size_t)n * JQP + j];
    qm_s[j] = (j < d.JQ && (!d.has_mask || hq_mask[(size_t)n * d.JQ + j])) ? 1.f : 0.f;
  }

  float acc[JQP];
#pragma unroll
  for (int j = 0; j < JQP; ++j) acc[j] = 0.f;
  float zdot = 0.f, lin = 0.f, quad = 0.f, hg = 0.f;

  const float* qeffT_n = qeffT + (size_t)n * Dp * JQP;
  for (int c = 0; c < cp.nchunk; ++c) {
    const int c0 = c * cp.DC;
    const int DCc = min(cp.DC, Dp - c0);
    if (c > 0) __syncthreads();
    load_chunk(d, hall, n, t0, c0, DCc, cp.stride, h_s, qeffT_n, q_s, JQP, coef, coef_s, Dp, cp.DC, gvec);
    row_chunk_dot<JQ4, true>(h_s + (size_t)threadIdx.x * cp.stride, q_s, coef_s, DCc, cp.DC, acc, zdot, lin,
                             quad, hg);
  }

  // ---- gate (same sequence as the forward) ---------------------------------------------
  float g = 1.f, tz = 0.f, cnt = 0.f;
  if (d.use_time_warp) {
    zd_s[threadIdx.x] = zdot;
    __syncthreads();
    float z = zconst[n];
    for (int kk = 0; kk < d.K; ++kk) z += zd_s[kk * kAttnTT + tt];
    g = gate_from_z(d, z, t, tz, cnt);
    __syncthreads();  // zd_s is reused below
  }
  float c1, c0v;
  row_coefs(d, g, lin, quad, c1, c0v);
  const bool mh = in_range && (!d.has_mask || hall_mask[((size_t)n * d.K + k) * d.T + t]);
  const bool tanh_on = d.add_tanh && d.simi != 4;

  // ---- recompute row max / ties ----------------------------------------------------------
  float m = -INFINITY;
  int rowcnt = 0;
#pragma unroll
  for (int j = 0; j < JQP; ++j) {
    if (j < d.JQ) {
      float th;
      const float a = row_logit(c1, acc[j], c0v, col_s[j], tanh_on, d.has_mask, mh && qm_s[j] != 0.f, th);
      if (a > m) { m = a; rowcnt = 1; }
      else if (a == m) { rowcnt++; }
    }
  }
  const float* st = stats + ((size_t)n * d.K + k) * kStatW;
  const float Mk = st[0], Lk = st[1], sk = st[2], tiesk = st[3];
  const int amax_t = (int)st[5], amax_j = (int)st[6];
  const float* dr = drow + ((size_t)n * d.K + k) * 4;
  const float dsk = dr[0], rdr = dr[1], beta = dr[2];
  const float ml = d.time_warp_att ? m * g : m;
  // a fully masked sequence has a uniform softmax over all T (the forward may encode that with
  // denormal-scale weights; do not divide them here)
  const float alpha = !in_range ? 0.f : (sk <= -1e29f ? 1.f / (float)d.T : expf(ml - Mk) / Lk);
  const float e = beta * hg;                 // h . d r_k
  const float dml = alpha * (g * e - rdr);   // softmax over T
  float dg = alpha * e;                      // r_k = sum alpha * g * h
  float dm = dml;
  if (d.time_warp_att) { dm = dml * g; dg = fmaf(dml, m, dg); }
  const float dm_each = dm / (float)rowcnt;
  const float ds_each = dsk / tiesk;

  // ---- d a -> d S; row sums --------------------------------------------------------------
  float dSsum = 0.f, dSP = 0.f;
  float* dSrow = dS_s + (size_t)threadIdx.x * JQP;
#pragma unroll
  for (int j = 0; j < JQP; ++j) {
    float dS = 0.f;
    if (j < d.JQ && in_range) {
      float th;
      const float a = row_logit(c1, acc[j], c0v, col_s[j], tanh_on, d.has_mask, mh && qm_s[j] != 0.f, th);
      float da = 0.f;
      if (a == m) da += dm_each;
      // a unique maximum is routed by the position the forward recorded (robust to the forward
      // and this recompute using different arithmetic); ties are split by value equality
      if (tiesk == 1.f ? (t == amax_t && j == amax_j) : (a == sk)) da += ds_each;
      dS = tanh_on ? da * (1.f - th * th) : da;
    }
    dSsum += dS;
    dSP = fmaf(dS, acc[j], dSP);
    dSrow[j] = dS;
    acc[j] = c1 * dS;  // d P
  }
  float dlin = 0.f, dquad = 0.f;
  if (d.simi == 4) {
    const float v = g * g * quad;
    if (v > 1e-12f) {
      const float rs = rsqrtf(v);
      dquad = dSP * (-0.5f) * g * g * g * rs * rs * rs;
    } else {
      dg = fmaf(dSP, 1e6f, dg);  // rsqrt(1e-12)
    }
  } else {
    dg += dSP + lin * dSsum + 2.f * g * quad * dSsum;
    dlin = g * dSsum;
    dquad = g * g * dSsum;
  }
  c1_s[threadIdx.x] = c1;
  dlin_s[threadIdx.x] = dlin;
  dquad_s[threadIdx.x] = dquad;
  float dz = 0.f;
  if (d.use_time_warp) {
    zd_s[threadIdx.x] = in_range ? dg : 0.f;
    __syncthreads();
    float dgt = 0.f;
    for (int kk = 0; kk < d.K; ++kk) dgt += zd_s[kk * kAttnTT + tt];
    dz = dgt * cnt * (1.f - tz * tz);
    if (k == 0) {
      dz_s[tt] = dz;
      const float s = warp_sum(in_range ? dz : 0.f);
      if (tt == 0) atomicAdd(dzc + n, s);
    }
  } else if (k == 0) {
    dz_s[tt] = 0.f;
  }
  const float hcoef = alpha * g * beta;  // d h~ = alpha * d r_k, times the gate
  __syncthreads();

  // ---- d col: column sums of d S over the tile's rows -----------------------------------
  if (d.simi != 4) {
    for (int j = threadIdx.x; j < d.JQ; j += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s += dS_s[(size_t)r * JQP + j];
      atomicAdd(dcol + (size_t)n * JQP + j, s);
    }
  }

  // ---- second sweep over D: question-side / coefficient gradients, then d_hall ----------
  float* dq_n = dqeffT + (size_t)n * Dp * JQP;
  float* dcoef_n = dcoef + (size_t)n * 3 * Dp;
  for (int c = 0; c < cp.nchunk; ++c) {
    const int c0 = c * cp.DC;
    const int DCc = min(cp.DC, Dp - c0);
    if (cp.nchunk > 1) {
      __syncthreads();
      load_chunk(d, hall, n, t0, c0, DCc, cp.stride, h_s, qeffT_n, q_s, JQP, coef, coef_s, Dp, cp.DC, gvec);
    }
    // d qeffT[dd][j] += sum_r h[r][dd] * c1[r] * dS[r][j]
    for (int p = threadIdx.x; p < DCc * JQ4; p += blockDim.x) {
      const int dd = p / JQ4, j4 = p - dd * JQ4;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < rows; ++r) {
        const float hc = h_s[(size_t)r * cp.stride + dd] * c1_s[r];
        const float4 ds = *reinterpret_cast<const float4*>(dS_s + (size_t)r * JQP + j4 * 4);
        o.x = fmaf(hc, ds.x, o.x); o.y = fmaf(hc, ds.y, o.y);
        o.z = fmaf(hc, ds.z, o.z); o.w = fmaf(hc, ds.w, o.w);
      }
      float* dst = dq_n + (size_t)(c0 + dd) * JQP + j4 * 4;
      atomicAdd(dst + 0, o.x); atomicAdd(dst + 1, o.y); atomicAdd(dst + 2, o.z); atomicAdd(dst + 3, o.w);
    }
    // d u, d w_lin, d w_quad
    for (int dd = threadIdx.x; dd < DCc; dd += blockDim.x) {
      float du = 0.f, dwl = 0.f, dwq = 0.f;
      for (int r = 0; r < rows; ++r) {
        const float h = h_s[(size_t)r * cp.stride + dd];
        const float h2 = h * h;
        du = fmaf(dz_s[r & (kAttnTT - 1)], h2, du);
        dwl = fmaf(dlin_s[r], h, dwl);
        dwq = fmaf(dquad_s[r], h2, dwq);
      }
      atomicAdd(dcoef_n + c0 + dd, du);
      atomicAdd(dcoef_n + Dp + c0 + dd, dwl);
      atomicAdd(dcoef_n + 2 * Dp + c0 + dd, dwq);
    }
    __syncthreads();
    // d h (in place over the tile), thread per row
    {
      float* hrow = h_s + (size_t)threadIdx.x * cp.stride;
      for (int cc = 0; cc < DCc; cc += 4) {
        float4 hv = *reinterpret_cast<float4*>(hrow + cc);
        float hx[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float h = hx[e2];
          const float4* qrow = reinterpret_cast<const float4*>(q_s + (size_t)(cc + e2) * JQP);
          float v = 0.f;
#pragma unroll
          for (int j4 = 0; j4 < JQ4; ++j4) {
            const float4 qv = qrow[j4];
            v = fmaf(acc[4 * j4 + 0], qv.x, v); v = fmaf(acc[4 * j4 + 1], qv.y, v);
            v = fmaf(acc[4 * j4 + 2], qv.z, v); v = fmaf(acc[4 * j4 + 3], qv.w, v);
          }
          v = fmaf(coef_s[cp.DC + cc + e2], dlin, v);
          v = fmaf(2.f * h * coef_s[2 * cp.DC + cc + e2], dquad, v);
          v = fmaf(hcoef, coef_s[3 * cp.DC + cc + e2], v);
          v = fmaf(2.f * h * coef_s[cc + e2], dz, v);
          hx[e2] = v;
        }
        *reinterpret_cast<float4*>(hrow + cc) = make_float4(hx[0], hx[1], hx[2], hx[3]);
      }
    }
    __syncthreads();
    // coalesced store of the tile's d_hall chunk
    if ((d.D & 3) == 0) {
      const int v4 = DCc >> 2;
      for (int i = threadIdx.x; i < rows * v4; i += blockDim.x) {
        int r = i / v4, c4 = i - r * v4;
        int kk = r / kAttnTT, t2 = t0 + (r - kk * kAttnTT);
        if (t2 < d.T)
          *reinterpret_cast<float4*>(d_hall + (((size_t)n * d.K + kk) * d.T + t2) * d.D + c0 + c4 * 4) =
              *reinterpret_cast<const float4*>(h_s + (size_t)r * cp.stride + c4 * 4);
      }
    } else {
      for (int i = threadIdx.x; i < rows * DCc; i += blockDim.x) {
        int r = i / DCc, c2 = i - r * DCc;
        int kk = r / kAttnTT, t2 = t0 + (r - kk * kAttnTT);
        if (t2 < d.T && c0 + c2 < d.D)
          d_hall[(((size_t)n * d.K + kk) * d.T + t2) * d.D + c0 + c2] = h_s[(size_t)r * cp.stride + c2];
      }
    }
  }
}

// d_hq[n][j][:] and d_lq[n][:] from the accumulated question-side gradients. grid N.
__global__ void attn_bwd_q(FvtaAttnDesc d, FvtaAttnParams p, const float* __restrict__ hq,
                           const float* __restrict__ dqeffT, const float* __restrict__ dcol,
                           const float* __restrict__ dzc, float* __restrict__ d_hq, float* __restrict__ d_lq,
                           int Dp, int JQP) {
  extern __shared__ float sm[];
  float* scale = sm;         // [JQP]
  float* proj = sm + JQP;    // [JQP]  sum_e dqeff[e][j] q[j][e] (cosine only)
  const int n = blockIdx.x;
  const int D = d.D, JQ = d.JQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const float* q = hq + (size_t)n * JQ * D;
  const float* dq = dqeffT + (size_t)n * Dp * JQP;
  if (d.simi == 4) {
    for (int j = warp; j < JQ; j += nwarp) {
      float qq = 0.f, pr = 0.f;
      for (int e = lane; e < D; e += 32) {
        const float v = q[(size_t)j * D + e];
        qq = fmaf(v, v, qq);
        pr = fmaf(dq[(size_t)e * JQP + j], v, pr);
      }
      qq = warp_sum(qq);
      pr = warp_sum(pr);
      if (lane == 0) {
        const float s = rsqrtf(fmaxf(qq, 1e-12f));
        scale[j] = s;
        proj[j] = (qq > 1e-12f) ? -s * s * s * pr : 0.f;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < JQ * D; i += blockDim.x) {
    const int j = i / D, dd = i - j * D;
    const float qv = q[i];
    const float dqe = dq[(size_t)dd * JQP + j];
    const float dc = dcol[(size_t)n * JQP + j];
    float v;
    if (d.simi == 1) v = dqe * p.att_W[2 * D + dd] + dc * p.att_W[D + dd];
    else if (d.simi == 2) v = dqe * (p.att_W[dd] - 2.f * p.att_W[D + dd]) + dc * 2.f * p.att_W[D + dd] * qv;
    else if (d.simi == 3)
      v = dqe * (p.att_W[3 * D + dd] - 2.f * p.att_W[2 * D + dd]) +
          dc * (p.att_W[D + dd] + 2.f * p.att_W[2 * D + dd] * qv);
    else v = dqe * scale[j] + proj[j] * qv;
    d_hq[(size_t)n * JQ * D + i] = v;
  }
  if (d_lq != nullptr) {
    for (int dd = threadIdx.x; dd < D; dd += blockDim.x)
      d_lq[(size_t)n * D + dd] = d.use_time_warp ? dzc[n] * (float)d.K * p.WC_W[dd] : 0.f;
  }
}

// One block per feature dd: reduce over samples and question words -> d att_W, d u[dd].
// Block 0 also produces d att_b and sum_n dzc.
__global__ void attn_bwd_params(FvtaAttnDesc d, const float* __restrict__ hq, const float* __restrict__ dqeffT,
                                const float* __restrict__ dcol, const float* __restrict__ dcoef,
                                const float* __restrict__ dzc, FvtaAttnGrads g, float* __restrict__ du,
                                float* __restrict__ scal, int Dp, int JQP) {
  __shared__ float red[6][32];
  const int dd = blockIdx.x;
  const int D = d.D, JQ = d.JQ;
  float a1 = 0.f, a2 = 0.f, a3 = 0.f, su = 0.f, sl = 0.f, sq = 0.f;
  for (int i = threadIdx.x; i < d.N * JQ; i += blockDim.x) {
    const int n = i / JQ, j = i - n * JQ;
    const float qv = hq[((size_t)n * JQ + j) * D + dd];
    const float dc = dcol[(size_t)n * JQP + j];
    a1 = fmaf(dqeffT[((size_t)n * Dp + dd) * JQP + j], qv, a1);
    a2 = fmaf(dc, qv, a2);
    a3 = fmaf(dc, qv * qv, a3);
  }
  for (int n = threadIdx.x; n < d.N; n += blockDim.x) {
    su += dcoef[((size_t)n * 3 + 0) * Dp + dd];
    sl += dcoef[((size_t)n * 3 + 1) * Dp + dd];
    sq += dcoef[((size_t)n * 3 + 2) * Dp + dd];
  }
  float vals[6] = {a1, a2, a3, su, sl, sq};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float v = warp_sum(vals[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[6];
    for (int i = 0; i < 6; ++i) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += red[i][w];
      tot[i] = s;
    }
    du[dd] = tot[3];
    if (g.att_W != nullptr) {
      if (d.simi == 1) { g.att_W[dd] = tot[4]; g.att_W[D + dd] = tot[1]; g.att_W[2 * D + dd] = tot[0]; }
      else if (d.simi == 2) { g.att_W[dd] = tot[0]; g.att_W[D + dd] = -2.f * tot[0] + tot[5] + tot[2]; }
      else if (d.simi == 3) {
        g.att_W[dd] = tot[4]; g.att_W[D + dd] = tot[1];
        g.att_W[2 * D + dd] = -2.f * tot[0] + tot[5] + tot[2]; g.att_W[3 * D + dd] = tot[0];
      }
    }
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    float sb = 0.f, sz = 0.f;
    for (int i = threadIdx.x; i < d.N * JQ; i += blockDim.x) sb += dcol[(size_t)(i / JQ) * JQP + (i % JQ)];
    for (int n = threadIdx.x; n < d.N; n += blockDim.x) sz += dzc[n];
    sb = warp_sum(sb);
    sz = warp_sum(sz);
    if (lane == 0) { red[0][warp] = sb; red[1][warp] = sz; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = 0.f, z = 0.f;
      for (int w = 0; w < nwarp; ++w) { b += red[0][w]; z += red[1][w]; }
      scal[0] = z;
      if (g.att_b != nullptr && d.simi != 4) g.att_b[0] = b;
    }
  }
}

// Time-warp parameter gradients from d u and d zconst (last line).
// grid: D blocks (one per column e of W_WH / entry of W_WC).
__global__ void attn_bwd_warp(FvtaAttnDesc d, FvtaAttnParams p, const float* __restrict__ lq,
                              const float* __restrict__ dzc, const float* __restrict__ du,
                              const float* __restrict__ scal, FvtaAttnGrads g) {
  __shared__ float red[2][32];
  const int e = blockIdx.x;
  const int D = d.D;
  const float wce = p.WC_W[e];
  // d W_WH[dd][e] = du[dd] * wc[e] for dd < D; rows D..2D multiply zeros -> zero gradient
  if (g.WH_W != nullptr) {
    for (int dd = threadIdx.x; dd < 2 * D; dd += blockDim.x)
      g.WH_W[(size_t)dd * D + e] = dd < D ? du[dd] * wce : 0.f;
  }
  float a = 0.f, b = 0.f;
  for (int dd = threadIdx.x; dd < D; dd += blockDim.x) a = fmaf(p.WH_W[(size_t)dd * D + e], du[dd], a);
  for (int n = threadIdx.x; n < d.N; n += blockDim.x) b = fmaf(dzc[n], lq[(size_t)n * D + e], b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, sb = 0.f;
    for (int w = 0; w < nwarp; ++w) { sa += red[0][w]; sb += red[1][w]; }
    const float K = (float)d.K, sz = scal[0];
    if (g.WC_W != nullptr) g.WC_W[e] = sa + K * (sb + sz * p.WH_b[e]);
    if (g.WH_b != nullptr) g.WH_b[e] = K * wce * sz;
    if (e == 0 && g.WC_b != nullptr) g.WC_b[0] = K * sz;
  }
}

size_t attn_tile_bwd_smem_bytes(const FvtaAttnDesc& d, int JQP, int Dp) {
  const ChunkPlan cp = chunk_plan(Dp);
  const int rows = d.K * kAttnTT;
  size_t fl = (size_t)rows * cp.stride + (size_t)cp.DC * JQP + kCoefRows * cp.DC + (size_t)rows * JQP +
              4 * rows + kAttnTT + 2 * JQP;
  return fl * sizeof(float);
}

template <int JQ4>
static int launch_tile_bwd(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall,
                           const uint8_t* hall_mask, const uint8_t* hq_mask, float* ws, const float* stats,
                           const float* dg1, float* d_hall, cudaStream_t st) {
  size_t smem = attn_tile_bwd_smem_bytes(d, 4 * JQ4, L.Dp);
  cudaError_t e = cudaFuncSetAttribute(attn_tile_bwd<JQ4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_last_error("attn_tile_bwd smem attr", e); return FVTA_E_LAUNCH; }
  dim3 grid(L.nt, d.N);
  attn_tile_bwd<JQ4><<<grid, d.K * kAttnTT, smem, st>>>(
      d, hall, hall_mask, hq_mask, ws + L.coef, ws + L.qeffT, ws + L.col, ws + L.zconst, stats, ws + L.drow, dg1,
      d_hall, ws + L.dqeffT, ws + L.dcol, ws + L.dcoef, ws + L.dzc, L.Dp);
  return check_launch("attn_tile_bwd");
}

}  // namespace fvta

using namespace fvta;

namespace fvta {
bool attn_bwd_tc_covers(const FvtaAttnDesc& d);
int attn_bwd_tc_try(const FvtaAttnDesc& d, const AttnLayout& L, const float* hall, const uint8_t* hall_mask,
                    const uint8_t* hq_mask, float* ws, const float* stats, const float* dg1, float* d_hall,
                    cudaStream_t st);
bool attn_small_ok(const FvtaAttnDesc& d);
size_t attn_small_ws_floats(const FvtaAttnDesc& d);
int attn_small_bwd(const FvtaAttnDesc& d, const float* hall, const uint8_t* hall_mask, const float* hq,
                   const uint8_t* hq_mask, const FvtaAttnParams& P, const float* dg1, float* d_hall, float* d_hq,
                   const FvtaAttnGrads& G, float* ws, cudaStream_t st);
}  // namespace fvta

extern "C" int fvta_attn_bwd(const FvtaAttnDesc* desc, const float* hall, const uint8_t* hall_mask,
                             const float* hq, const uint8_t* hq_mask, const float* lq,
                             const FvtaAttnParams* params, const float* stats, const float* d_g1_all,
                             float* d_hall, float* d_hq, float* d_lq, const FvtaAttnGrads* grads,
                             void* workspace, size_t workspace_bytes, void* stream) {
  int rc = attn_validate(desc);
  if (rc) return rc;
  const FvtaAttnDesc d = *desc;
  if (!hall || !hq || !params || !stats || !d_g1_all || !d_hall || !d_hq || !grads || !workspace)
    return FVTA_E_BADARG;
  if (d.has_mask && (!hall_mask || !hq_mask)) return FVTA_E_BADARG;
  if (d.use_time_warp && (!lq || !params->WH_W || !params->WH_b || !params->WC_W || !params->WC_b))
    return FVTA_E_BADARG;
  if (d.simi != 4 && (!params->att_W || !params->att_b)) return FVTA_E_BADARG;
  const AttnLayout L = attn_layout(d);
  if (workspace_bytes < L.total * sizeof(float)) return FVTA_E_WORKSPACE;
  if (attn_tile_bwd_smem_bytes(d, L.JQP, L.Dp) > 227 * 1024) return FVTA_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  // question-attention shape: fused per-sample kernels (attn_small.cu), nothing taken from the forward
  if (attn_small_ok(d) && d.path != FVTA_PATH_FFMA && workspace_bytes >= attn_small_ws_floats(d) * sizeof(float))
    return attn_small_bwd(d, hall, hall_mask, hq, hq_mask, *params, d_g1_all, d_hall, d_hq, *grads, ws, st);
  // The folded operands are recomputed (cheap) so the workspace need not survive between
  // the forward and the backward call.
  // tensor-core tile kernel (attn_bwd_tc.cu) when the shape / flags are covered; it needs the question
  // images next to the folded operands: ws_from_fwd == 2 says the forward left both
  const bool use_tc = d.path != FVTA_PATH_FFMA && attn_bwd_tc_covers(d);
  if (!d.ws_from_fwd || (use_tc && d.ws_from_fwd != 2)) {
    rc = attn_prep_launch(d, L, *params, hq, lq, ws, st, use_tc);
    if (rc) return rc;
  }
  cudaError_t e = cudaMemsetAsync(ws + L.dqeffT, 0, (L.drow - L.dqeffT) * sizeof(float), st);
  if (e != cudaSuccess) { set_last_error("attn_bwd memset", e); return FVTA_E_LAUNCH; }
  attn_bwd_prep<<<d.N, d.K * 32, 0, st>>>(d, stats, d_g1_all, ws + L.drow);
  rc = check_launch("attn_bwd_prep");
  if (rc) return rc;
  if (use_tc) {
    rc = attn_bwd_tc_try(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st);
    if (rc < 0) return rc;
  }
  if (!use_tc || rc == 1) switch (L.JQP) {
    case 4: rc = launch_tile_bwd<1>(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st); break;
    case 8: rc = launch_tile_bwd<2>(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st); break;
    case 16: rc = launch_tile_bwd<4>(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st); break;
    case 28: rc = launch_tile_bwd<7>(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st); break;
    case 32: rc = launch_tile_bwd<8>(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st); break;
    case 64: rc = launch_tile_bwd<16>(d, L, hall, hall_mask, hq_mask, ws, stats, d_g1_all, d_hall, st); break;
    default: return FVTA_E_SHAPE;
  }
  if (rc) return rc;
  attn_bwd_q<<<d.N, 256, 2 * L.JQP * sizeof(float), st>>>(d, *params, hq, ws + L.dqeffT, ws + L.dcol, ws + L.dzc,
                                                           d_hq, d_lq, L.Dp, L.JQP);
  rc = check_launch("attn_bwd_q");
  if (rc) return rc;
  attn_bwd_params<<<d.D, 128, 0, st>>>(d, hq, ws + L.dqeffT, ws + L.dcol, ws + L.dcoef, ws + L.dzc, *grads,
                                        ws + L.du, ws + L.scal, L.Dp, L.JQP);
  rc = check_launch("attn_bwd_params");
  if (rc) return rc;
  if (d.use_time_warp) {
    attn_bwd_warp<<<d.D, 128, 0, st>>>(d, *params, lq, ws + L.dzc, ws + L.du, ws + L.scal, *grads);
    rc = check_launch("attn_bwd_warp");
  }
  return rc;
}
