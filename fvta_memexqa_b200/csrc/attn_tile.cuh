// Device pieces shared by the forward tile kernel and the backward's recompute, so both
// evaluate the logits with the SAME instruction sequence (the backward routes gradient to
// arg-max entries by exact float equality with the forward's maxima).
#pragma once
#include "common.cuh"

namespace fvta {

constexpr int kCoefRows = 4;  // u, w_lin, w_quad, upstream gradient (backward only)

// Stream one D-chunk of the K*32-row context tile, the matching slice of qeffT and of the
// coefficient vectors (+ optionally the per-sample vector `gvec`, coefficient row 3) into
// shared memory.  Ends with a __syncthreads().
__device__ __forceinline__ void load_chunk(const FvtaAttnDesc& d, const float* __restrict__ hall, int n,
                                           int t0, int c0, int DCc, int stride, float* h_s,
                                           const float* __restrict__ qeffT_n, float* q_s, int JQP,
                                           const float* __restrict__ coef, float* coef_s, int Dp, int DC,
                                           const float* __restrict__ gvec = nullptr) {
  const int rows = d.K * kAttnTT;
  const int D = d.D;
  if ((D & 3) == 0) {
    const int v4 = DCc >> 2;
    for (int i = threadIdx.x; i < rows * v4; i += blockDim.x) {
      int r = i / v4, c = i - r * v4;
      int k = r / kAttnTT, tt = r - k * kAttnTT;
      int t = t0 + tt;
      float* dst = h_s + (size_t)r * stride + c * 4;
      if (t < d.T) {
        const float* src = hall + (((size_t)n * d.K + k) * d.T + t) * D + c0 + c * 4;
        cp_async16(dst, src);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else {
    for (int i = threadIdx.x; i < rows * DCc; i += blockDim.x) {
      int r = i / DCc, c = i - r * DCc;
      int k = r / kAttnTT, tt = r - k * kAttnTT;
      int t = t0 + tt;
      float v = 0.f;
      if (t < d.T && c0 + c < D) v = hall[(((size_t)n * d.K + k) * d.T + t) * D + c0 + c];
      h_s[(size_t)r * stride + c] = v;
    }
  }
  {
    const float* src = qeffT_n + (size_t)c0 * JQP;
    const int n4 = (DCc * JQP) >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(q_s + i * 4, src + i * 4);
  }
  for (int i = threadIdx.x; i < 3 * DCc; i += blockDim.x) {
    int w = i / DCc, c = i - w * DCc;
    coef_s[w * DC + c] = coef[(size_t)w * Dp + c0 + c];
  }
  if (gvec != nullptr) {
    for (int c = threadIdx.x; c < DCc; c += blockDim.x)
      coef_s[3 * DC + c] = (c0 + c < D) ? gvec[c0 + c] : 0.f;
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
}

// One thread owns one context row (k, t): P[j] for every question word plus the row
// statistics, accumulated over one D-chunk held in shared memory.
//   zdot = u.h^2, lin = w_lin.h, quad = w_quad.h^2, hg = gvec.h (WITH_G only)
template <int JQ4, bool WITH_G>
__device__ __forceinline__ void row_chunk_dot(const float* __restrict__ hrow, const float* __restrict__ q_s,
                                              const float* __restrict__ coef_s, int DCc, int DC,
                                              float (&acc)[4 * JQ4], float& zdot, float& lin, float& quad,
                                              float& hg) {
  constexpr int JQP = 4 * JQ4;
  for (int c = 0; c < DCc; c += 4) {
    const float4 hv = *reinterpret_cast<const float4*>(hrow + c);
    const float hx[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float h = hx[e];
      const float h2 = h * h;
      zdot = fmaf(coef_s[c + e], h2, zdot);
      lin = fmaf(coef_s[DC + c + e], h, lin);
      quad = fmaf(coef_s[2 * DC + c + e], h2, quad);
      if (WITH_G) hg = fmaf(coef_s[3 * DC + c + e], h, hg);
      const float4* qrow = reinterpret_cast<const float4*>(q_s + (size_t)(c + e) * JQP);
#pragma unroll
      for (int j4 = 0; j4 < JQ4; ++j4) {
        const float4 qv = qrow[j4];
        acc[4 * j4 + 0] = fmaf(h, qv.x, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(h, qv.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(h, qv.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(h, qv.w, acc[4 * j4 + 3]);
      }
    }
  }
}

// Gate of one position from the K row dots:  g = tanh(z) * cnt(t)   (model_v2.py:986-1007)
__device__ __forceinline__ float gate_from_z(const FvtaAttnDesc& d, float z, int t, float& tz, float& cnt) {
  tz = tanhf(z);
  cnt = window_count(t, d.T, d.warp_type, d.window);
  return tz * cnt;
}

// Row coefficients of S[t,j] = c1*P[t,j] + c0 + col[j]
__device__ __forceinline__ void row_coefs(const FvtaAttnDesc& d, float g, float lin, float quad, float& c1,
                                          float& c0) {
  if (d.simi == 4) {
    c1 = g * rsqrtf(fmaxf(g * g * quad, 1e-12f));  // l2_normalize of the gated row
    c0 = 0.f;
  } else {
    c1 = g;
    c0 = fmaf(g, lin, g * g * quad);
  }
}

// One logit: returns the masked value a; `th` receives the unmasked activation.
__device__ __forceinline__ float row_logit(float c1, float p, float c0, float colj, bool tanh_on,
                                           bool has_mask, bool valid, float& th) {
  const float S = fmaf(c1, p, c0) + colj;
  th = tanh_on ? tanhf(S) : S;
  float a = th;
  if (has_mask) a = a + (valid ? 0.f : 1.f) * FVTA_VERY_NEGATIVE;  // exp_mask in fp32
  return a;
}

}  // namespace fvta
