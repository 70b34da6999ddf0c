// Question attention (model_v2.py:1044 -> `attention`, :125-201): the context is the question
// itself (hinfo = hq [N,JQ,D]), the "question" is the single vector g1_all[n] -- in the terms of
// fvta_attn_fwd: K = 1, T = JQ_question <= 128 rows, JQ = 1, no time warp.  A few KB of work per
// sample: the general path (operand folding, tensor-core tile kernel, combine; in the backward
// prep / tile / question / parameter kernels) spends ~75 us on launch latency for it, so this
// shape gets one CTA per sample that does everything, forward in one launch, backward in two
// (the second adds the per-sample parameter-gradient partials in a fixed order).
//
// Same math as the general kernels:
//   simi 1-3:  S_t = h_t . (q * w_mul) + h_t . w_lin + h_t^2 . w_quad + (q . wq_lin + q^2 . wq_quad + b)
//   simi 4:    S_t = (h_t . q) / (max(|h_t|^2, 1e-12) max(|q|^2, 1e-12))^(1/2)
//   a_t = tanh(S_t) (simi 1-3 with add_tanh) + (valid ? 0 : -1e30);  alpha = softmax_t(a);
//   g = sum_t alpha_t h_t.   (The second-level softmax over K = 1 modality is the constant 1.)
// The backward recomputes the forward from its inputs: no statistics are exchanged.
#include "common.cuh"

namespace fvta {

constexpr int kSmallThreads = 128;
constexpr int kSmallMaxT = 128;

__host__ __device__ inline int small_dimw(int simi, int D) { return simi == 1 ? 3 * D : (simi == 2 ? 2 * D : (simi == 3 ? 4 * D : 0)); }

bool attn_small_ok(const FvtaAttnDesc& d) {
  if (d.K != 1 || d.JQ != 1 || d.use_time_warp || d.time_warp_att) return false;
  if (d.T > kSmallMaxT || (d.D & 3) != 0 || d.D > 1024) return false;
  return (size_t)(d.T + 6) * d.D * sizeof(float) + 9 * kSmallMaxT * sizeof(float) <= 96 * 1024;
}
size_t attn_small_smem(const FvtaAttnDesc& d) { return ((size_t)(d.T + 6) * d.D + 9 * kSmallMaxT) * sizeof(float); }
// floats of workspace the backward needs for the per-sample parameter-gradient partials
size_t attn_small_ws_floats(const FvtaAttnDesc& d) { return (size_t)d.N * (small_dimw(d.simi, d.D) + 1); }

__device__ __forceinline__ float block_sum128(float v, float* red) {   // blockDim.x == 128
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// Everything both directions need, left in shared memory:
//   h_s [T][D], q_s [D], weff_s [D] (q * w_mul, or q / |q| for cosine), wl_s [D], wq_s [D]
//   a_s [T] masked logits, th_s [T] activations, al_s [T] softmax, p_s [T] h.weff, c1_s [T] row scale,
//   h2_s [T] |h_t|^2 (cosine)
struct SmallSmem {
  float *h, *q, *weff, *wl, *wq, *aux0, *aux1, *a, *th, *al, *p, *c1, *h2, *da, *dS, *red;
};
__device__ __forceinline__ SmallSmem small_carve(float* sm, int T, int D) {
  SmallSmem s;
  s.h = sm; s.q = s.h + (size_t)T * D; s.weff = s.q + D; s.wl = s.weff + D; s.wq = s.wl + D;
  s.aux0 = s.wq + D; s.aux1 = s.aux0 + D;
  s.a = s.aux1 + D; s.th = s.a + kSmallMaxT; s.al = s.th + kSmallMaxT; s.p = s.al + kSmallMaxT;
  s.c1 = s.p + kSmallMaxT; s.h2 = s.c1 + kSmallMaxT; s.da = s.h2 + kSmallMaxT; s.dS = s.da + kSmallMaxT;
  s.red = s.dS + kSmallMaxT;
  return s;
}

// returns (|q|^2, 1/max(|q|, 1e-6)) for cosine, zeros otherwise
__device__ __forceinline__ float2 small_forward(const FvtaAttnDesc& d, const SmallSmem& s, const float* __restrict__ hall,
                                                const uint8_t* __restrict__ hall_mask,
                                                const uint8_t* __restrict__ hq_mask, const float* __restrict__ hq,
                                                const FvtaAttnParams& P, int n) {
  const int T = d.T, D = d.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* W = reinterpret_cast<const float*>(P.att_W);
  const bool tanh_on = d.add_tanh && d.simi != 4;
  // question vector and the folded coefficient vectors
  float colp = 0.f, qq = 0.f;
  for (int i = tid; i < D; i += kSmallThreads) {
    const float q = hq[(size_t)n * D + i];
    s.q[i] = q;
    qq = fmaf(q, q, qq);
    float w_mul = 0.f, w_lin = 0.f, w_quad = 0.f, wq_lin = 0.f, wq_quad = 0.f;
    if (d.simi == 1) { w_lin = W[i]; wq_lin = W[D + i]; w_mul = W[2 * D + i]; }
    else if (d.simi == 2) { w_mul = W[i] - 2.f * W[D + i]; w_quad = W[D + i]; wq_quad = W[D + i]; }
    else if (d.simi == 3) { w_lin = W[i]; wq_lin = W[D + i]; w_quad = W[2 * D + i]; wq_quad = W[2 * D + i]; w_mul = W[3 * D + i] - 2.f * W[2 * D + i]; }
    else { w_quad = 1.f; }
    s.weff[i] = q * w_mul;      // cosine: filled below once |q| is known
    s.wl[i] = w_lin;
    s.wq[i] = w_quad;
    colp += q * wq_lin + q * q * wq_quad;
  }
  // context tile
  const float4* src = reinterpret_cast<const float4*>(hall + (size_t)n * T * D);
  float4* dst = reinterpret_cast<float4*>(s.h);
  for (int i = tid; i < T * D / 4; i += kSmallThreads) dst[i] = src[i];
  float col = block_sum128(colp, s.red);
  float rq = 0.f, q2 = 0.f;
  if (d.simi == 4) {
    q2 = block_sum128(qq, s.red);
    rq = rsqrtf(fmaxf(q2, 1e-12f));
    col = 0.f;
    for (int i = tid; i < D; i += kSmallThreads) s.weff[i] = s.q[i] * rq;
  } else {
    col += reinterpret_cast<const float*>(P.att_b)[0];
  }
  __syncthreads();
  // logits: one warp per row
  const bool qv = !d.has_mask || hq_mask[n] != 0;
  for (int t = warp; t < T; t += kSmallThreads / 32) {
    const float4* hr = reinterpret_cast<const float4*>(s.h + (size_t)t * D);
    float p = 0.f, lin = 0.f, quad = 0.f;
    for (int c = lane; c < D / 4; c += 32) {
      const float4 h = hr[c];
      const float4 we = reinterpret_cast<const float4*>(s.weff)[c];
      const float4 wl = reinterpret_cast<const float4*>(s.wl)[c];
      const float4 wq = reinterpret_cast<const float4*>(s.wq)[c];
      p = fmaf(h.x, we.x, fmaf(h.y, we.y, fmaf(h.z, we.z, fmaf(h.w, we.w, p))));
      lin = fmaf(h.x, wl.x, fmaf(h.y, wl.y, fmaf(h.z, wl.z, fmaf(h.w, wl.w, lin))));
      quad = fmaf(h.x * h.x, wq.x, fmaf(h.y * h.y, wq.y, fmaf(h.z * h.z, wq.z, fmaf(h.w * h.w, wq.w, quad))));
    }
    p = warp_sum(p); lin = warp_sum(lin); quad = warp_sum(quad);
    if (lane == 0) {
      float c1 = 1.f, c0 = lin + quad;
      if (d.simi == 4) { c1 = rsqrtf(fmaxf(quad, 1e-12f)); c0 = 0.f; }
      const float S = fmaf(c1, p, c0) + col;
      const float th = tanh_on ? tanhf(S) : S;
      const bool valid = !d.has_mask || (qv && hall_mask[(size_t)n * T + t] != 0);
      s.th[t] = th;
      s.a[t] = d.has_mask ? th + (valid ? 0.f : 1.f) * FVTA_VERY_NEGATIVE : th;
      s.p[t] = p;
      s.c1[t] = c1;     // cosine: 1/|h_t|
      s.h2[t] = quad;   // cosine: |h_t|^2
    }
  }
  __syncthreads();
  // softmax over the T rows (warp 0)
  if (warp == 0) {
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s.a[t]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(s.a[t] - m); s.al[t] = e; l += e; }
    l = warp_sum(l);
    const float rl = 1.f / l;
    for (int t = lane; t < T; t += 32) s.al[t] *= rl;
  }
  __syncthreads();
  return make_float2(q2, rq);
}

__global__ void __launch_bounds__(kSmallThreads) attn_small_fwd_kernel(
    FvtaAttnDesc d, const float* __restrict__ hall, const uint8_t* __restrict__ hall_mask,
    const float* __restrict__ hq, const uint8_t* __restrict__ hq_mask, FvtaAttnParams P, float* __restrict__ g1_all,
    float* __restrict__ a_logits) {
  extern __shared__ __align__(16) float sm[];
  const SmallSmem s = small_carve(sm, d.T, d.D);
  const int n = blockIdx.x;
  small_forward(d, s, hall, hall_mask, hq_mask, hq, P, n);
  for (int i = threadIdx.x; i < d.D; i += kSmallThreads) {
    float r = 0.f;
    for (int t = 0; t < d.T; ++t) r = fmaf(s.al[t], s.h[(size_t)t * d.D + i], r);
    g1_all[(size_t)n * d.D + i] = r;
  }
  if (a_logits != nullptr)
    for (int t = threadIdx.x; t < d.T; t += kSmallThreads) a_logits[(size_t)n * d.T + t] = s.a[t];
}

// grid N.  part [N][dimW + 1]: this sample's contribution to d att_W (dimW) and d att_b (1).
__global__ void __launch_bounds__(kSmallThreads) attn_small_bwd_kernel(
    FvtaAttnDesc d, const float* __restrict__ hall, const uint8_t* __restrict__ hall_mask,
    const float* __restrict__ hq, const uint8_t* __restrict__ hq_mask, FvtaAttnParams P,
    const float* __restrict__ dg1, float* __restrict__ d_hall, float* __restrict__ d_hq, float* __restrict__ part) {
  extern __shared__ __align__(16) float sm[];
  const int T = d.T, D = d.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SmallSmem s = small_carve(sm, T, D);
  const int n = blockIdx.x;
  const float2 cr = small_forward(d, s, hall, hall_mask, hq_mask, hq, P, n);
  const float q2 = cr.x, rq = cr.y;
  const bool tanh_on = d.add_tanh && d.simi != 4;
  float* dg_s = s.aux0;   // [D] upstream gradient
  for (int i = tid; i < D; i += kSmallThreads) dg_s[i] = dg1[(size_t)n * D + i];
  __syncthreads();
  // d alpha_t = dg . h_t
  for (int t = warp; t < T; t += kSmallThreads / 32) {
    const float4* hr = reinterpret_cast<const float4*>(s.h + (size_t)t * D);
    float v = 0.f;
    for (int c = lane; c < D / 4; c += 32) {
      const float4 h = hr[c];
      const float4 g = reinterpret_cast<const float4*>(dg_s)[c];
      v = fmaf(h.x, g.x, fmaf(h.y, g.y, fmaf(h.z, g.z, fmaf(h.w, g.w, v))));
    }
    v = warp_sum(v);
    if (lane == 0) s.da[t] = v;
  }
  __syncthreads();
  // softmax backward, activation backward (the additive mask passes gradients through, as in TF)
  if (warp == 0) {
    float dot = 0.f;
    for (int t = lane; t < T; t += 32) dot = fmaf(s.al[t], s.da[t], dot);
    dot = warp_sum(dot);
    float s0 = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float da = s.al[t] * (s.da[t] - dot);
      const float dS = tanh_on ? da * (1.f - s.th[t] * s.th[t]) : da;
      s.dS[t] = dS;
      s0 += dS;
    }
    s0 = warp_sum(s0);
    if (lane == 0) s.red[4] = s0;
  }
  __syncthreads();
  const float s0 = s.red[4];
  const float* W = reinterpret_cast<const float*>(P.att_W);
  const int dimw = small_dimw(d.simi, D);
  float* pn = part + (size_t)n * (dimw + 1);
  float* dh_n = d_hall + (size_t)n * T * D;
  if (d.simi != 4) {
    for (int i = tid; i < D; i += kSmallThreads) {
      const float q = s.q[i], we = s.weff[i], wl = s.wl[i], wq = s.wq[i], g = dg_s[i];
      float A1 = 0.f, A2 = 0.f;
      for (int t = 0; t < T; ++t) {
        const float h = s.h[(size_t)t * D + i], dS = s.dS[t];
        A1 = fmaf(dS, h, A1);
        A2 = fmaf(dS * h, h, A2);
        dh_n[(size_t)t * D + i] = fmaf(s.al[t], g, dS * (we + wl + 2.f * h * wq));
      }
      const float qA1 = q * A1, sq = A2 - 2.f * q * A1 + q * q * s0;
      float w_mul, wq_lin = 0.f, wq_quad = 0.f;
      if (d.simi == 1) {
        w_mul = W[2 * D + i]; wq_lin = W[D + i];
        pn[i] = A1; pn[D + i] = q * s0; pn[2 * D + i] = qA1;
      } else if (d.simi == 2) {
        w_mul = W[i] - 2.f * W[D + i]; wq_quad = W[D + i];
        pn[i] = qA1; pn[D + i] = sq;
      } else {
        w_mul = W[3 * D + i] - 2.f * W[2 * D + i]; wq_lin = W[D + i]; wq_quad = W[2 * D + i];
        pn[i] = A1; pn[D + i] = q * s0; pn[2 * D + i] = sq; pn[3 * D + i] = qA1;
      }
      d_hq[(size_t)n * D + i] = fmaf(w_mul, A1, s0 * (wq_lin + 2.f * q * wq_quad));
    }
    if (tid == 0) pn[dimw] = s0;
  } else {
    // cosine: S_t = c1_t * (h_t . qhat),  c1_t = 1/|h_t|,  qhat = q * rq
    float* bv_s = s.aux1;   // [D] d qhat
    float bq = 0.f;
    for (int i = tid; i < D; i += kSmallThreads) {
      const float qh = s.weff[i], g = dg_s[i];
      float Bv = 0.f;
      for (int t = 0; t < T; ++t) {
        const float h = s.h[(size_t)t * D + i], dS = s.dS[t], c1 = s.c1[t];
        Bv = fmaf(dS * c1, h, Bv);
        // |h|^2 <= 1e-12: the norm is the clamp constant, no gradient through it
        const float shrink = (s.h2[t] > 1e-12f) ? s.p[t] * c1 * c1 * c1 * h : 0.f;
        dh_n[(size_t)t * D + i] = fmaf(s.al[t], g, dS * (c1 * qh - shrink));
      }
      bv_s[i] = Bv;
      bq = fmaf(Bv, s.q[i], bq);
    }
    const float bqs = block_sum128(bq, s.red);
    for (int i = tid; i < D; i += kSmallThreads) {
      const float shrink = (q2 > 1e-12f) ? rq * rq * rq * bqs * s.q[i] : 0.f;
      d_hq[(size_t)n * D + i] = rq * bv_s[i] - shrink;
    }
  }
}

// d att_W [dimW], d att_b [1] = sum over the samples' partials (overwrites, like the general path)
__global__ void attn_small_bwd_params(int N, int dimw, const float* __restrict__ part, float* __restrict__ gW,
                                      float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > dimw) return;
  float v = 0.f;
  for (int n = 0; n < N; ++n) v += part[(size_t)n * (dimw + 1) + i];
  if (i < dimw) { if (gW) gW[i] = v; }
  else if (gb) gb[0] = v;
}

int attn_small_fwd(const FvtaAttnDesc& d, const float* hall, const uint8_t* hall_mask, const float* hq,
                   const uint8_t* hq_mask, const FvtaAttnParams& P, float* g1_all, float* a_logits, cudaStream_t st) {
  const size_t smem = attn_small_smem(d);
  cudaError_t e = cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_last_error("attn_small_fwd smem attr", e); return FVTA_E_LAUNCH; }
  attn_small_fwd_kernel<<<d.N, kSmallThreads, smem, st>>>(d, hall, hall_mask, hq, hq_mask, P, g1_all, a_logits);
  return check_launch("attn_small_fwd_kernel");
}

int attn_small_bwd(const FvtaAttnDesc& d, const float* hall, const uint8_t* hall_mask, const float* hq,
                   const uint8_t* hq_mask, const FvtaAttnParams& P, const float* dg1, float* d_hall, float* d_hq,
                   const FvtaAttnGrads& G, float* ws, cudaStream_t st) {
  const size_t smem = attn_small_smem(d);
  cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_last_error("attn_small_bwd smem attr", e); return FVTA_E_LAUNCH; }
  attn_small_bwd_kernel<<<d.N, kSmallThreads, smem, st>>>(d, hall, hall_mask, hq, hq_mask, P, dg1, d_hall, d_hq, ws);
  int rc = check_launch("attn_small_bwd_kernel");
  if (rc || d.simi == 4) return rc;
  const int dimw = small_dimw(d.simi, d.D);
  attn_small_bwd_params<<<(dimw + 1 + 127) / 128, 128, 0, st>>>(d.N, dimw, ws, reinterpret_cast<float*>(G.att_W),
                                                               reinterpret_cast<float*>(G.att_b));
  return check_launch("attn_small_bwd_params");
}

}  // namespace fvta
