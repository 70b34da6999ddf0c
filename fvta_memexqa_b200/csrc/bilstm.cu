// Bidirectional LSTM reader over padded variable-length sequences (fp32, sm_100a).
//
// Replaces tf.nn.bidirectional_dynamic_rnn(BasicLSTMCell(d), sequence_length=...) at the eight
// call sites model_v2.py:694-823.  Semantics encoded:
//   gates = [x, h] @ kernel + bias, order i, j, f, o;  c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j);
//   h' = tanh(c')*sigmoid(o);  zero initial state;  outputs are zero past the length;
//   the final state is the state at the last valid step;  the backward direction walks the
//   valid prefix in reverse and writes each output at its original position.
//
// Forward = one input-projection GEMM per distinct weight set (all timesteps at once) +
// one persistent recurrent kernel: a CTA owns kSG sequences of one direction for all their
// timesteps, keeps W_h, h and c in shared memory and never leaves the SM between steps.
// Backward = the mirrored recurrent kernel (BPTT) producing d gates, then GEMMs for d_x and
// the weight gradients.
#include "common.cuh"
#include "bilstm.cuh"
#include <cstdlib>

namespace fvta {

constexpr int kSG = 8;          // sequences per CTA
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct LstmSmemPlan {
  bool w_in_smem;
  int wstride;  // 4d+1 when in smem (odd: conflict-free for both access patterns)
  size_t bytes;
};
__host__ __device__ inline LstmSmemPlan lstm_smem_plan(int d) {
  LstmSmemPlan p;
  size_t w = (size_t)d * (4 * d + 1) * sizeof(float);
  size_t rest = (size_t)kSG * (d + d + 4 * d + 4 * d) * sizeof(float) + 128;   // h/dh, c/dc, gates, partials, alignment
  p.w_in_smem = (w + rest) <= 200 * 1024;
  p.wstride = p.w_in_smem ? 4 * d + 1 : 4 * d;
  p.bytes = (p.w_in_smem ? w : 0) + rest;
  return p;
}

// Forward recurrence for d <= 64.  What bounds a step is the number of bytes shared memory returns
// to registers (a broadcast float4 of h costs a full 512-byte warp write-back), so a gate thread
// owns TWO gate columns (g and g + 2d) and one quarter of the hidden units: per unit it reads the
// two float4s h[e][0..7] once and issues eight packed FMAs against its 2 x HQ register-resident
// W_h entries.  The cell update runs one (unit, sequence) item per thread -- the cell state and the
// prefetched projected input of its four gates never leave its registers -- and adds the four
// quarter partials.
// Shared memory: h [4 HQ][kSG] (transposed: the sequences of a unit contiguous), gate partials
// [4][4d][kSG] (same orientation: stores and item-ordered loads are both contiguous).
// Optional per-step timeline of CTA (0,0) (build with -DFVTA_LSTM_TRACE, read with fvta_lstm_trace_read)
#ifdef FVTA_LSTM_TRACE
__device__ long long g_lstm_trace[32][8];
#define LSTM_TRACE(step, ev) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (step) < 32) g_lstm_trace[step][ev] = clock64(); } while (0)
#else
#define LSTM_TRACE(step, ev) do {} while (0)
#endif

template <int HQ>
__global__ void __launch_bounds__(HQ <= 13 ? 416 : 512, 2)
bilstm_rec_fwd_split(FvtaLstmDesc L, SeqAddr A, const int32_t* __restrict__ lens, const float* __restrict__ xp_fw,
                     const float* __restrict__ xp_bw, const float* __restrict__ wh_fw,
                     const float* __restrict__ wh_bw, float* __restrict__ out, float* __restrict__ last,
                     float* __restrict__ reserve) {
  static_assert(kSG == 8, "item layout below assumes 8 sequences per CTA");
  constexpr int HS = 4 * HQ;
  __shared__ __align__(16) float h_s[HS][kSG];
  __shared__ __align__(16) float g_s[4][4 * HS][kSG];
  __shared__ int len_s[kSG];
  const int d = L.d, G = 4 * d, G2 = 2 * d;
  const int dir = blockIdx.y;
  const float* wh_g = dir == 0 ? wh_fw : wh_bw;
  const float* xp = dir == 0 ? xp_fw : xp_bw;
  const int b0 = blockIdx.x * kSG;
  // ---- gate-phase role: (column pair cp, quarter) ----
  const bool gate = threadIdx.x < 4 * G2;
  const int qtr = gate ? threadIdx.x / G2 : 0;
  const int cp = threadIdx.x - qtr * G2;
  const int e0 = qtr * HQ;
  float w0[HQ], w1[HQ];
#pragma unroll
  for (int i = 0; i < HQ; ++i) {
    const bool ok = gate && e0 + i < d;
    w0[i] = ok ? wh_g[(size_t)(e0 + i) * G + cp] : 0.f;
    w1[i] = ok ? wh_g[(size_t)(e0 + i) * G + G2 + cp] : 0.f;
  }
  for (int i = threadIdx.x; i < HS * kSG; i += blockDim.x) (&h_s[0][0])[i] = 0.f;
  if (threadIdx.x < kSG) {
    const int b = b0 + threadIdx.x;
    len_s[threadIdx.x] = (b < L.B) ? min(lens[b], L.J) : 0;
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int q = 0; q < kSG; ++q) maxlen = max(maxlen, len_s[q]);
  // ---- cell-phase role: item (unit e, sequence s), s fastest ----
  const bool item = threadIdx.x < kSG * d;
  const int ie = threadIdx.x / kSG, is = threadIdx.x % kSG;
  const int ib = b0 + is;
  const int ilen = item ? len_s[is] : 0;
  const int itok = (item && ib < L.B) ? A.tok0(ib, L.J) : 0;
  const long long iorow = (item && ib < L.B) ? A.out0(ib, L.J) : -1;
  const float* xitem = xp + (size_t)itok * G + ie;
  float c = 0.f;
  float xg[4] = {0.f, 0.f, 0.f, 0.f};
  auto fetch_x = [&](int step) {
    if (step < ilen) {
      const float* r = xitem + (size_t)(dir == 0 ? step : ilen - 1 - step) * G;
      xg[0] = r[0]; xg[1] = r[d]; xg[2] = r[2 * d]; xg[3] = r[3 * d];
    }
  };
  fetch_x(0);
  const float* hrow = &h_s[e0][0];
  float* gdst0 = &g_s[qtr][gate ? cp : 0][0];
  float* gdst1 = &g_s[qtr][gate ? G2 + cp : 0][0];
  const float* gsrc = &g_s[0][item ? ie : 0][is];
  constexpr int kQStride = 4 * HS * kSG;   // floats between quarter partials

  for (int step = 0; step < maxlen; ++step) {
    LSTM_TRACE(step, 0);
    if (gate) {
      float2 a01 = make_float2(0.f, 0.f), a23 = a01, a45 = a01, a67 = a01;
      float2 c01 = a01, c23 = a01, c45 = a01, c67 = a01;
#pragma unroll
      for (int i = 0; i < HQ; ++i) {
        const float4 x0 = *reinterpret_cast<const float4*>(hrow + i * kSG);
        const float4 x1 = *reinterpret_cast<const float4*>(hrow + i * kSG + 4);
        const float2 u = make_float2(w0[i], w0[i]), v = make_float2(w1[i], w1[i]);
        a01 = __ffma2_rn(u, make_float2(x0.x, x0.y), a01);
        a23 = __ffma2_rn(u, make_float2(x0.z, x0.w), a23);
        a45 = __ffma2_rn(u, make_float2(x1.x, x1.y), a45);
        a67 = __ffma2_rn(u, make_float2(x1.z, x1.w), a67);
        c01 = __ffma2_rn(v, make_float2(x0.x, x0.y), c01);
        c23 = __ffma2_rn(v, make_float2(x0.z, x0.w), c23);
        c45 = __ffma2_rn(v, make_float2(x1.x, x1.y), c45);
        c67 = __ffma2_rn(v, make_float2(x1.z, x1.w), c67);
      }
      reinterpret_cast<float4*>(gdst0)[0] = make_float4(a01.x, a01.y, a23.x, a23.y);
      reinterpret_cast<float4*>(gdst0)[1] = make_float4(a45.x, a45.y, a67.x, a67.y);
      reinterpret_cast<float4*>(gdst1)[0] = make_float4(c01.x, c01.y, c23.x, c23.y);
      reinterpret_cast<float4*>(gdst1)[1] = make_float4(c45.x, c45.y, c67.x, c67.y);
    }
    LSTM_TRACE(step, 3);
    __syncthreads();
    LSTM_TRACE(step, 4);
    if (item && step < ilen) {
      const int t = dir == 0 ? step : ilen - 1 - step;
      float gv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* p = gsrc + (size_t)k * d * kSG;
        gv[k] = (p[0] + p[kQStride]) + (p[2 * kQStride] + p[3 * kQStride]) + xg[k];
      }
      fetch_x(step + 1);   // lands during the next gate phase
      const float ig = sigmoid_fast(gv[0]);
      const float jg = tanh_fast(gv[1]);
      const float fg = sigmoid_fast(gv[2] + 1.0f);  // forget_bias = 1.0 [TF]
      const float og = sigmoid_fast(gv[3]);
      const float hp = h_s[ie][is];
      c = fmaf(c, fg, ig * jg);
      const float h = tanh_fast(c) * og;
      h_s[ie][is] = h;
      if (iorow >= 0) out[((size_t)iorow + t) * 2 * d + dir * d + ie] = h;
      float* r = reserve + ((size_t)dir * A.ntok + itok + t) * kResW * d + ie;
      r[0] = ig; r[d] = jg; r[2 * d] = fg; r[3 * d] = og; r[4 * d] = c; r[5 * d] = hp;
    }
    LSTM_TRACE(step, 5);
    __syncthreads();
    LSTM_TRACE(step, 6);
  }
  if (item && ib < L.B) last[(size_t)ib * 2 * d + dir * d + ie] = h_s[ie][is];
}

#ifdef FVTA_LSTM_TRACE
}  // namespace fvta
extern "C" int fvta_lstm_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, fvta::g_lstm_trace, sizeof(long long) * 32 * 8);
}
namespace fvta {
#endif

// grid (ceil(B/kSG), 2); block: multiple of 32
__global__ void bilstm_rec_fwd(FvtaLstmDesc L, SeqAddr A, const int32_t* __restrict__ lens,
                               const float* __restrict__ xp_fw, const float* __restrict__ xp_bw,
                               const float* __restrict__ wh_fw, const float* __restrict__ wh_bw,
                               float* __restrict__ out, float* __restrict__ last, float* __restrict__ reserve) {
  extern __shared__ __align__(16) float sm[];
  const int d = L.d, G = 4 * d;
  const LstmSmemPlan plan = lstm_smem_plan(d);
  const int dir = blockIdx.y;
  const float* wh_g = dir == 0 ? wh_fw : wh_bw;   // [d][4d] rows In.. of the kernel
  const float* xp = dir == 0 ? xp_fw : xp_bw;     // [B*J][4d]
  float* w_s = sm;
  float* h_s = sm + (plan.w_in_smem ? (size_t)d * plan.wstride : 0);  // [kSG][d]
  float* c_s = h_s + kSG * d;                                         // [kSG][d]
  float* g_s = c_s + kSG * d;                                         // [kSG][4d]
  __shared__ int len_s[kSG];
  __shared__ int tok_s[kSG];
  __shared__ long long orow_s[kSG];
  const int b0 = blockIdx.x * kSG;
  if (plan.w_in_smem)
    for (int i = threadIdx.x; i < d * G; i += blockDim.x) w_s[(i / G) * plan.wstride + (i % G)] = wh_g[i];
  for (int i = threadIdx.x; i < kSG * d; i += blockDim.x) { h_s[i] = 0.f; c_s[i] = 0.f; }
  if (threadIdx.x < kSG) {
    const int b = b0 + threadIdx.x;
    len_s[threadIdx.x] = (b < L.B) ? min(lens[b], L.J) : 0;
    tok_s[threadIdx.x] = (b < L.B) ? A.tok0(b, L.J) : 0;
    orow_s[threadIdx.x] = (b < L.B) ? A.out0(b, L.J) : -1;
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int s = 0; s < kSG; ++s) maxlen = max(maxlen, len_s[s]);
  const float* W = plan.w_in_smem ? w_s : wh_g;
  const int ws = plan.wstride;

  for (int step = 0; step < maxlen; ++step) {
    // gate pre-activations: g_s[s][g] = xp[tok][g] + sum_e h[s][e] * Wh[e][g]
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      float acc[kSG];
#pragma unroll
      for (int s = 0; s < kSG; ++s) {
        const int len = len_s[s];
        const int t = dir == 0 ? step : len - 1 - step;
        acc[s] = (step < len) ? xp[((size_t)tok_s[s] + t) * G + g] : 0.f;
      }
      for (int e = 0; e < d; ++e) {
        const float w = W[(size_t)e * ws + g];
#pragma unroll
        for (int s = 0; s < kSG; ++s) acc[s] = fmaf(h_s[s * d + e], w, acc[s]);
      }
#pragma unroll
      for (int s = 0; s < kSG; ++s) g_s[s * G + g] = acc[s];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSG * d; i += blockDim.x) {
      const int s = i / d, e = i - s * d;
      const int len = len_s[s];
      if (step < len) {
        const int t = dir == 0 ? step : len - 1 - step;
        const float* gs = g_s + s * G;
        const float ig = sigmoidf_(gs[e]);
        const float jg = tanhf(gs[d + e]);
        const float fg = sigmoidf_(gs[2 * d + e] + 1.0f);  // forget_bias = 1.0 [TF]
        const float og = sigmoidf_(gs[3 * d + e]);
        const float hp = h_s[i];
        const float c = fmaf(c_s[i], fg, ig * jg);
        const float h = tanhf(c) * og;
        c_s[i] = c;
        h_s[i] = h;
        if (orow_s[s] >= 0) out[((size_t)orow_s[s] + t) * 2 * d + dir * d + e] = h;
        float* r = reserve + ((size_t)dir * A.ntok + tok_s[s] + t) * kResW * d;
        r[e] = ig; r[d + e] = jg; r[2 * d + e] = fg; r[3 * d + e] = og; r[4 * d + e] = c; r[5 * d + e] = hp;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < kSG * d; i += blockDim.x) {
    const int s = i / d, e = i - s * d;
    if (b0 + s < L.B) last[(size_t)(b0 + s) * 2 * d + dir * d + e] = h_s[i];
  }
}

// BPTT.  dgates [2][B*J][4d] must be zero-initialised (padded steps stay zero).
__global__ void bilstm_rec_bwd(FvtaLstmDesc L, SeqAddr A, const int32_t* __restrict__ lens,
                               const float* __restrict__ wh_fw, const float* __restrict__ wh_bw,
                               const float* __restrict__ reserve, const float* __restrict__ d_out,
                               const float* __restrict__ d_last, float* __restrict__ dgates) {
  extern __shared__ __align__(16) float sm[];
  const int d = L.d, G = 4 * d;
  const LstmSmemPlan plan = lstm_smem_plan(d);
  const int dir = blockIdx.y;
  const float* wh_g = dir == 0 ? wh_fw : wh_bw;
  float* w_s = sm;
  // every region starts on a 16-byte boundary (float4 reads of the transposed d gates)
  float* dh_s = sm + (plan.w_in_smem ? round_up_sz((size_t)d * plan.wstride, 4) : 0);  // [kSG][d]
  float* dc_s = dh_s + round_up(kSG * d, 4);                                          // [kSG][d]
  float* dg_s = dc_s + round_up(kSG * d, 4);                                          // [4d][kSG] + partials
  __shared__ int len_s[kSG];
  __shared__ int tok_s[kSG];
  __shared__ long long orow_s[kSG];
  const int b0 = blockIdx.x * kSG;
  if (plan.w_in_smem)
    for (int i = threadIdx.x; i < d * G; i += blockDim.x) w_s[(i / G) * plan.wstride + (i % G)] = wh_g[i];
  for (int i = threadIdx.x; i < kSG * d; i += blockDim.x) { dh_s[i] = 0.f; dc_s[i] = 0.f; }
  if (threadIdx.x < kSG) {
    const int b = b0 + threadIdx.x;
    len_s[threadIdx.x] = (b < L.B) ? min(lens[b], L.J) : 0;
    tok_s[threadIdx.x] = (b < L.B) ? A.tok0(b, L.J) : 0;
    orow_s[threadIdx.x] = (b < L.B) ? A.out0(b, L.J) : -1;
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int s = 0; s < kSG; ++s) maxlen = max(maxlen, len_s[s]);
  const float* W = plan.w_in_smem ? w_s : wh_g;
  const int ws = plan.wstride;
  float* dgd = dgates + (size_t)dir * A.ntok * G;

  // d gates of the current step live transposed in shared memory, dgT[g][kSG] (sequence index
  // contiguous), so the d h_prev product reads them as two broadcast float4s per g while W_h[e][g]
  // is read once per g for all kSG sequences: thread (e, quarter) accumulates one quarter of the g
  // range (= one gate block) for the kSG sequences, the four quarters are summed through shared
  // memory.  (One LDS per multiply-add, as in a (sequence, e) mapping, makes this loop LSU-bound.)
  float* dgT = dg_s;                        // [4d][kSG]
  float* part_s = dg_s + (size_t)G * kSG;   // [4][kSG][d]
  const int nq = (int)blockDim.x >= 4 * d ? 4 : ((int)blockDim.x >= 2 * d ? 2 : 1);   // quarters run in parallel
  for (int step = maxlen - 1; step >= 0; --step) {
    for (int i = threadIdx.x; i < kSG * d; i += blockDim.x) {
      const int s = i / d, e = i - s * d;
      const int len = len_s[s];
      float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
      if (step < len) {
        const int b = b0 + s;
        const int t = dir == 0 ? step : len - 1 - step;
        const float* r = reserve + ((size_t)dir * A.ntok + tok_s[s] + t) * kResW * d;
        const float ig = r[e], jg = r[d + e], fg = r[2 * d + e], og = r[3 * d + e], c = r[4 * d + e];
        float c_prev = 0.f;
        if (step > 0) {
          const int tp = dir == 0 ? t - 1 : t + 1;
          c_prev = reserve[((size_t)dir * A.ntok + tok_s[s] + tp) * kResW * d + 4 * d + e];
        }
        float dh = dh_s[i];
        if (d_out && orow_s[s] >= 0) dh += d_out[((size_t)orow_s[s] + t) * 2 * d + dir * d + e];
        if (d_last && step == len - 1) dh += d_last[(size_t)b * 2 * d + dir * d + e];
        const float tc = tanhf(c);
        const float dc = fmaf(dh * og, 1.f - tc * tc, dc_s[i]);
        dgo = dh * tc * og * (1.f - og);
        dgi = dc * jg * ig * (1.f - ig);
        dgj = dc * ig * (1.f - jg * jg);
        dgf = dc * c_prev * fg * (1.f - fg);
        dc_s[i] = dc * fg;
        float* o = dgd + ((size_t)tok_s[s] + t) * G;
        o[e] = dgi; o[d + e] = dgj; o[2 * d + e] = dgf; o[3 * d + e] = dgo;
      }
      dgT[(size_t)e * kSG + s] = dgi;
      dgT[(size_t)(d + e) * kSG + s] = dgj;
      dgT[(size_t)(2 * d + e) * kSG + s] = dgf;
      dgT[(size_t)(3 * d + e) * kSG + s] = dgo;
    }
    __syncthreads();
    // partial d h_prev over one part of the g range
    for (int i = threadIdx.x; i < nq * d; i += blockDim.x) {
      const int qtr = i / d, e = i - qtr * d;
      const int g0 = qtr * (G / nq), g1 = g0 + G / nq;
      const float* wr = W + (size_t)e * ws;
      float acc[kSG];
#pragma unroll
      for (int q = 0; q < kSG; ++q) acc[q] = 0.f;
      for (int g = g0; g < g1; ++g) {
        const float w = wr[g];
        const float4 a0 = *reinterpret_cast<const float4*>(dgT + (size_t)g * kSG);
        const float4 a1 = *reinterpret_cast<const float4*>(dgT + (size_t)g * kSG + 4);
        acc[0] = fmaf(a0.x, w, acc[0]); acc[1] = fmaf(a0.y, w, acc[1]);
        acc[2] = fmaf(a0.z, w, acc[2]); acc[3] = fmaf(a0.w, w, acc[3]);
        acc[4] = fmaf(a1.x, w, acc[4]); acc[5] = fmaf(a1.y, w, acc[5]);
        acc[6] = fmaf(a1.z, w, acc[6]); acc[7] = fmaf(a1.w, w, acc[7]);
      }
#pragma unroll
      for (int q = 0; q < kSG; ++q) part_s[((size_t)qtr * kSG + q) * d + e] = acc[q];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSG * d; i += blockDim.x) {
      const int s = i / d;
      if (step < len_s[s]) {
        float v = 0.f;
        for (int q = 0; q < nq; ++q) v += part_s[(size_t)q * kSG * d + i];
        dh_s[i] = v;
      }
    }
    __syncthreads();
  }
}

// BPTT for kSG * d <= blockDim.x (every thread owns ONE (unit e, sequence s) item, s fastest):
//   * the item's cell-state gradient lives in a register for the whole sequence;
//   * the saved gates / cell states / incoming d out of the NEXT step to process are fetched while
//     the d h_prev product of the current one runs (global latency off the per-step critical path);
//   * d gates go to shared memory transposed, dgT[g][kSG] (conflict-free with this item order).
//     The d h_prev product is bound by how many bytes shared memory has to return to registers
//     (every broadcast float4 costs a full 512-byte warp write-back), so a product thread owns TWO
//     hidden units and one of kParts slices of the gate range: per gate it reads two W_h scalars
//     and the two broadcast float4s of dgT[g] and issues eight packed FMAs.  The kParts partial
//     sums part[p][e][kSG] are added up by the next step's item threads -- two barriers per step.
// Shared memory: W_h [d][4d+1] | dgT [4d][kSG] | part [kParts][d][kSG].
constexpr int kParts = 8;

template <int BT>
__global__ void __launch_bounds__(BT, BT <= 256 ? 4 : (BT <= 416 ? 2 : 1))
bilstm_rec_bwd_item(FvtaLstmDesc L, SeqAddr A, const int32_t* __restrict__ lens,
                                    const float* __restrict__ wh_fw, const float* __restrict__ wh_bw,
                                    const float* __restrict__ reserve, const float* __restrict__ d_out,
                                    const float* __restrict__ d_last, float* __restrict__ dgates) {
  static_assert(kSG == 8, "item layout below assumes 8 sequences per CTA");
  extern __shared__ __align__(16) float sm[];
  const int d = L.d, G = 4 * d;
  const int wsd = G + 1;
  const int dir = blockIdx.y;
  const float* wh_g = dir == 0 ? wh_fw : wh_bw;
  float* w_s = sm;                                              // [d][4d+1]
  float* dgT = sm + round_up_sz((size_t)d * wsd, 4);            // [4d][kSG]
  float* part_s = dgT + (size_t)G * kSG;                        // [kParts][d][kSG]
  __shared__ int len_s[kSG];
  const int b0 = blockIdx.x * kSG;
  for (int i = threadIdx.x; i < d * G; i += blockDim.x) w_s[(i / G) * wsd + (i % G)] = wh_g[i];
  for (int i = threadIdx.x; i < kParts * d * kSG; i += blockDim.x) part_s[i] = 0.f;
  if (threadIdx.x < kSG) {
    const int b = b0 + threadIdx.x;
    len_s[threadIdx.x] = (b < L.B) ? min(lens[b], L.J) : 0;
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int q = 0; q < kSG; ++q) maxlen = max(maxlen, len_s[q]);

  // ---- item state ----
  const bool item = threadIdx.x < kSG * d;
  const int e = threadIdx.x / kSG, s = threadIdx.x % kSG;
  const int b = b0 + s;
  const int len = item ? len_s[s] : 0;
  const int tok0 = (item && b < L.B) ? A.tok0(b, L.J) : 0;
  const long long orow = (item && b < L.B) ? A.out0(b, L.J) : -1;
  const float* res = reserve + ((size_t)dir * A.ntok + tok0) * kResW * d + e;   // + t * kResW * d + slot * d
  const float* dout_p = (d_out && orow >= 0) ? d_out + (size_t)orow * 2 * d + dir * d + e : nullptr;
  float* dgd = dgates + ((size_t)dir * A.ntok + tok0) * G + e;
  const float* part_i = part_s + (size_t)e * kSG + s;
  float dc = 0.f;
  float ig = 0.f, jg = 0.f, fg = 0.f, og = 0.f, cc = 0.f, cp = 0.f, dout = 0.f;
  auto fetch = [&](int st) {   // saved values of recurrence step st (0 <= st < len)
    const int t = dir == 0 ? st : len - 1 - st;
    const float* r = res + (size_t)t * kResW * d;
    ig = r[0]; jg = r[d]; fg = r[2 * d]; og = r[3 * d]; cc = r[4 * d];
    cp = 0.f;
    if (st > 0) cp = res[(size_t)(dir == 0 ? t - 1 : t + 1) * kResW * d + 4 * d];
    dout = dout_p ? dout_p[(size_t)t * 2 * d] : 0.f;
    if (d_last && st == len - 1) dout += d_last[(size_t)b * 2 * d + dir * d + e];
  };
  if (item && maxlen - 1 < len) fetch(maxlen - 1);

  // ---- d h_prev product mapping: thread (part, unit pair) ----
  const int dh2 = (d + 1) >> 1;                     // units pe and pe + dh2
  const int gper = (G + kParts - 1) / kParts;
  const bool prod = threadIdx.x < kParts * dh2;
  const int pq = threadIdx.x / dh2, pe = threadIdx.x - pq * dh2;
  const int pe1 = pe + dh2 < d ? pe + dh2 : pe;     // odd d: the last pair repeats a unit, its second sum is dropped
  const int g0 = pq * gper, g1 = min(G, g0 + gper);
  const float* wr0 = w_s + (size_t)pe * wsd;
  const float* wr1 = w_s + (size_t)pe1 * wsd;

  for (int step = maxlen - 1; step >= 0; --step) {
    if (item) {
      float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
      if (step < len) {
        float dh = dout;
#pragma unroll
        for (int q = 0; q < kParts; ++q) dh += part_i[(size_t)q * d * kSG];
        const float tc = tanh_fast(cc);
        const float dcv = fmaf(dh * og, 1.f - tc * tc, dc);
        dgo = dh * tc * og * (1.f - og);
        dgi = dcv * jg * ig * (1.f - ig);
        dgj = dcv * ig * (1.f - jg * jg);
        dgf = dcv * cp * fg * (1.f - fg);
        dc = dcv * fg;
        const int t = dir == 0 ? step : len - 1 - step;
        float* o = dgd + (size_t)t * G;
        o[0] = dgi; o[d] = dgj; o[2 * d] = dgf; o[3 * d] = dgo;
      }
      dgT[(size_t)e * kSG + s] = dgi;
      dgT[(size_t)(d + e) * kSG + s] = dgj;
      dgT[(size_t)(2 * d + e) * kSG + s] = dgf;
      dgT[(size_t)(3 * d + e) * kSG + s] = dgo;
      if (step >= 1 && step - 1 < len) fetch(step - 1);   // lands while the product below runs
    }
    __syncthreads();
    if (prod) {
      float2 a01 = make_float2(0.f, 0.f), a23 = a01, a45 = a01, a67 = a01;
      float2 c01 = a01, c23 = a01, c45 = a01, c67 = a01;
#pragma unroll 5
      for (int g = g0; g < g1; ++g) {
        const float w0 = wr0[g], w1 = wr1[g];
        const float4 x0 = *reinterpret_cast<const float4*>(dgT + (size_t)g * kSG);
        const float4 x1 = *reinterpret_cast<const float4*>(dgT + (size_t)g * kSG + 4);
        const float2 u = make_float2(w0, w0), v = make_float2(w1, w1);
        a01 = __ffma2_rn(u, make_float2(x0.x, x0.y), a01);
        a23 = __ffma2_rn(u, make_float2(x0.z, x0.w), a23);
        a45 = __ffma2_rn(u, make_float2(x1.x, x1.y), a45);
        a67 = __ffma2_rn(u, make_float2(x1.z, x1.w), a67);
        c01 = __ffma2_rn(v, make_float2(x0.x, x0.y), c01);
        c23 = __ffma2_rn(v, make_float2(x0.z, x0.w), c23);
        c45 = __ffma2_rn(v, make_float2(x1.x, x1.y), c45);
        c67 = __ffma2_rn(v, make_float2(x1.z, x1.w), c67);
      }
      float4* dst = reinterpret_cast<float4*>(part_s + ((size_t)pq * d + pe) * kSG);
      dst[0] = make_float4(a01.x, a01.y, a23.x, a23.y);
      dst[1] = make_float4(a45.x, a45.y, a67.x, a67.y);
      if (pe1 != pe) {
        dst = reinterpret_cast<float4*>(part_s + ((size_t)pq * d + pe1) * kSG);
        dst[0] = make_float4(c01.x, c01.y, c23.x, c23.y);
        dst[1] = make_float4(c45.x, c45.y, c67.x, c67.y);
      }
    }
    __syncthreads();
  }
}

__global__ void add_inplace4(size_t n4, float4* __restrict__ a, const float4* __restrict__ b) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 x = a[i];
    const float4 y = b[i];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    a[i] = x;
  }
}

// out[n] += sum_m A[m][n]
__global__ void colsum_accumulate(int M, int N, const float* __restrict__ A, int lda, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  float acc = 0.f;
  if (n < N)
    for (int m = blockIdx.y * 8 + w; m < M; m += gridDim.y * 8) acc += A[(size_t)m * lda + n];
  red[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && n < N) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    atomicAdd(out + n, s);
  }
}

}  // namespace fvta

using namespace fvta;

static int lstm_validate(const FvtaLstmDesc* L) {
  if (!L) return FVTA_E_BADARG;
  if (L->B <= 0 || L->J <= 0 || L->In <= 0 || L->d <= 0) return FVTA_E_BADARG;
  return FVTA_OK;
}

// Forward over `ntok` packed token rows (dense layout: ntok = B*J).
//   workspace: xp_fw [ntok*4d], xp_bw [ntok*4d];   reserve: [2][ntok][6d]
namespace fvta {
int lstm_fwd_tc_try(const FvtaLstmDesc& L, const SeqAddr& A, const int32_t* lens, const float* xp_fw, const float* xp_bw,
                    const float* wh_fw, const float* wh_bw, float* out, float* last, float* reserve, cudaStream_t st);
}

static int lstm_fwd_impl(const FvtaLstmDesc& L, const SeqAddr& A, const float* x, const int32_t* lens,
                         const float* kernel_fw, const float* bias_fw, const float* kernel_bw,
                         const float* bias_bw, float* out, float* last, float* reserve, float* workspace,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int G = 4 * L.d;
  float* xp_fw = workspace;
  float* xp_bw = xp_fw + (size_t)A.ntok * G;
  // input projection for every token at once: xp = x @ kernel[:In] + bias
  int rc = fvta_gemm(A.ntok, G, L.In, x, L.In, 0, kernel_fw, G, 0, bias_fw, 0, xp_fw, G, 0, stream);
  if (rc) return rc;
  const bool shared = (kernel_bw == kernel_fw && bias_bw == bias_fw);
  if (!shared) {
    rc = fvta_gemm(A.ntok, G, L.In, x, L.In, 0, kernel_bw, G, 0, bias_bw, 0, xp_bw, G, 0, stream);
    if (rc) return rc;
  } else {
    xp_bw = xp_fw;
  }
  // padded steps of `reserve` feed 0 * h_prev products in the backward GEMMs: keep them finite
  cudaError_t e = cudaMemsetAsync(reserve, 0, (size_t)2 * A.ntok * kResW * L.d * sizeof(float), st);
  if (e != cudaSuccess) { set_last_error("bilstm memset", e); return FVTA_E_LAUNCH; }
  const LstmSmemPlan plan = lstm_smem_plan(L.d);
  e = cudaFuncSetAttribute(bilstm_rec_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.bytes);
  if (e != cudaSuccess) { set_last_error("bilstm_rec_fwd smem attr", e); return FVTA_E_LAUNCH; }
  int threads = round_up(G < 1024 ? G : 1024, 32);
  dim3 grid((L.B + kSG - 1) / kSG, 2);
  if (L.path == FVTA_LSTM_TC) {   // d <= 64: the recurrent product on tcgen05 (bilstm_tc.cu)
    const int rc_tc = lstm_fwd_tc_try(L, A, lens, xp_fw, xp_bw, kernel_fw + (size_t)L.In * G, kernel_bw + (size_t)L.In * G,
                                      out, last, reserve, st);
    return rc_tc == 1 ? FVTA_E_SHAPE : rc_tc;
  }
  // FVTA_LSTM_AUTO: the CUDA-core recurrence is the faster one at the reference sizes (a step is bound by
  // the cell update and the hand-over latency, not by the 4d x d product: profiles/README.md)
  if (L.d <= 64) {   // W_h columns in registers (50; 100 is the flag default): W_h columns in registers
    const int th = round_up(8 * L.d, 32);
    if (L.d <= 32)
      bilstm_rec_fwd_split<8><<<grid, th, 0, st>>>(L, A, lens, xp_fw, xp_bw, kernel_fw + (size_t)L.In * G,
                                                    kernel_bw + (size_t)L.In * G, out, last, reserve);
    else if (L.d <= 52)
      bilstm_rec_fwd_split<13><<<grid, th, 0, st>>>(L, A, lens, xp_fw, xp_bw, kernel_fw + (size_t)L.In * G,
                                                     kernel_bw + (size_t)L.In * G, out, last, reserve);
    else
      bilstm_rec_fwd_split<16><<<grid, th, 0, st>>>(L, A, lens, xp_fw, xp_bw, kernel_fw + (size_t)L.In * G,
                                                     kernel_bw + (size_t)L.In * G, out, last, reserve);
    return check_launch("bilstm_rec_fwd_split");
  }
  bilstm_rec_fwd<<<grid, threads, plan.bytes, st>>>(L, A, lens, xp_fw, xp_bw, kernel_fw + (size_t)L.In * G,
                                                   kernel_bw + (size_t)L.In * G, out, last, reserve);
  return check_launch("bilstm_rec_fwd");
}

//   workspace: d gates [2][ntok][4d]
__global__ void add_out4(size_t n4, float4* __restrict__ out, const float4* __restrict__ a, const float4* __restrict__ b) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}

// The three weight/input products of the backward (d kernel[In:], d x, d kernel[:In]) are independent once
// the recurrence has produced the gate gradients; each fills a fraction of the 148 SMs for 15-25 us.  They
// run as a fork/join over two auxiliary streams (created once per device, non-blocking; events carry the
// dependencies, so inside a captured step they become parallel branches of the graph).  Two pairs alternate
// between calls so that the text and the image reader, which the caller may run on different streams, do
// not queue behind each other.  The only state the library keeps.
struct LstmAux {
  cudaStream_t s[4];
  cudaEvent_t e[8];
  int next;
  int state;      // 0 = not tried, 1 = ready, -1 = unavailable (serial path)
};
static LstmAux g_lstm_aux[32];

static LstmAux* lstm_aux() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  LstmAux& a = g_lstm_aux[dev];
  if (a.state == 0) {
    a.state = 1;
    for (int i = 0; i < 4 && a.state == 1; ++i)
      if (cudaStreamCreateWithFlags(&a.s[i], cudaStreamNonBlocking) != cudaSuccess) a.state = -1;
    for (int i = 0; i < 8 && a.state == 1; ++i)
      if (cudaEventCreateWithFlags(&a.e[i], cudaEventDisableTiming) != cudaSuccess) a.state = -1;
    if (a.state != 1) cudaGetLastError();
  }
  return a.state == 1 ? &a : nullptr;
}

static int lstm_bwd_impl(const FvtaLstmDesc& L, const SeqAddr& A, const float* x, const int32_t* lens,
                         const float* kernel_fw, const float* kernel_bw, const float* reserve, const float* d_out,
                         const float* d_last, float* d_x, float* d_kernel_fw, float* d_bias_fw,
                         float* d_kernel_bw, float* d_bias_bw, float* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int G = 4 * L.d, NT = A.ntok;
  float* dg = workspace;
  cudaError_t e = cudaMemsetAsync(dg, 0, (size_t)2 * NT * G * sizeof(float), st);
  if (e != cudaSuccess) { set_last_error("bilstm bwd memset", e); return FVTA_E_LAUNCH; }
  const LstmSmemPlan plan = lstm_smem_plan(L.d);
  dim3 grid((L.B + kSG - 1) / kSG, 2);
  int rc;
  const size_t item_smem = (round_up_sz((size_t)L.d * (G + 1), 4) + (size_t)G * kSG + (size_t)kParts * L.d * kSG) * sizeof(float);
  if (kSG * L.d <= 1024 && item_smem <= 200 * 1024) {
    const int threads = round_up(kSG * L.d, 32);
    auto kern = threads <= 256 ? bilstm_rec_bwd_item<256> : (threads <= 416 ? bilstm_rec_bwd_item<416> : bilstm_rec_bwd_item<1024>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)item_smem);
    if (e != cudaSuccess) { set_last_error("bilstm_rec_bwd_item smem attr", e); return FVTA_E_LAUNCH; }
    kern<<<grid, threads, item_smem, st>>>(L, A, lens, kernel_fw + (size_t)L.In * G, kernel_bw + (size_t)L.In * G,
                                           reserve, d_out, d_last, dg);
    rc = check_launch("bilstm_rec_bwd_item");
  } else {
    e = cudaFuncSetAttribute(bilstm_rec_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.bytes);
    if (e != cudaSuccess) { set_last_error("bilstm_rec_bwd smem attr", e); return FVTA_E_LAUNCH; }
    int threads = round_up((kSG * L.d) < 1024 ? (kSG * L.d) : 1024, 32);
    bilstm_rec_bwd<<<grid, threads, plan.bytes, st>>>(L, A, lens, kernel_fw + (size_t)L.In * G,
                                                     kernel_bw + (size_t)L.In * G, reserve, d_out, d_last, dg);
    rc = check_launch("bilstm_rec_bwd");
  }
  if (rc) return rc;
  const float* dgs[2] = {dg, dg + (size_t)NT * G};
  const float* ress[2] = {reserve, reserve + (size_t)NT * kResW * L.d};
  if (kernel_fw == kernel_bw && d_kernel_fw == d_kernel_bw && d_bias_fw == d_bias_bw) {
    // One cell for both directions (the reference's case): x is common, so
    //   d_x = (dg_fw + dg_bw) @ Wx^T,   d kernel[:In] += x^T @ (dg_fw + dg_bw),
    // and h_prev / dg of the two directions are contiguous in memory, so
    //   d kernel[In:] += [h_prev_fw ; h_prev_bw]^T @ [dg_fw ; dg_bw]   is one product with K = 2*ntok.
    float* dsum = dg + (size_t)2 * NT * G;          // dg_fw + dg_bw (third slice of the workspace)
    const size_t n4 = ((size_t)NT * G) / 4;         // G = 4d: always a multiple of 4
    dim3 cg((G + 31) / 32, 64);
    LstmAux* ax = lstm_aux();
    if (ax == nullptr) {                            // no auxiliary streams: the same work in a row
      rc = fvta_gemm(L.d, G, 2 * NT, reserve + 5 * L.d, kResW * L.d, 1, dg, G, 0, nullptr, 0,
                     d_kernel_fw + (size_t)L.In * G, G, 1, stream);
      if (rc) return rc;
      add_out4<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(n4, reinterpret_cast<float4*>(dsum),
                                                            reinterpret_cast<const float4*>(dgs[0]),
                                                            reinterpret_cast<const float4*>(dgs[1]));
      rc = check_launch("add_out4");
      if (rc) return rc;
      rc = fvta_gemm(NT, L.In, G, dsum, G, 0, kernel_fw, G, 1, nullptr, 0, d_x, L.In, 0, stream);
      if (rc) return rc;
      rc = fvta_gemm(L.In, G, NT, x, L.In, 1, dsum, G, 0, nullptr, 0, d_kernel_fw, G, 1, stream);
      if (rc) return rc;
      colsum_accumulate<<<cg, 256, 0, st>>>(NT, G, dsum, G, d_bias_fw);
      return check_launch("colsum_accumulate");
    }
    const int pair = (ax->next++) & 1;
    cudaStream_t sa = ax->s[2 * pair], sb = ax->s[2 * pair + 1];
    cudaEvent_t* ev = ax->e + 4 * pair;
    bool ok = cudaEventRecord(ev[0], st) == cudaSuccess && cudaStreamWaitEvent(sa, ev[0], 0) == cudaSuccess;
    if (!ok) { set_last_error("bilstm bwd fork", cudaGetLastError()); return FVTA_E_LAUNCH; }
    // branch a: d kernel[In:] += [h_prev_fw ; h_prev_bw]^T @ [dg_fw ; dg_bw]
    rc = fvta_gemm(L.d, G, 2 * NT, reserve + 5 * L.d, kResW * L.d, 1, dg, G, 0, nullptr, 0,
                   d_kernel_fw + (size_t)L.In * G, G, 1, sa);
    int rc_a = rc;
    ok = cudaEventRecord(ev[1], sa) == cudaSuccess;
    // calling stream: dsum, then d x
    add_out4<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(n4, reinterpret_cast<float4*>(dsum),
                                                          reinterpret_cast<const float4*>(dgs[0]),
                                                          reinterpret_cast<const float4*>(dgs[1]));
    rc = check_launch("add_out4");
    ok = ok && cudaEventRecord(ev[2], st) == cudaSuccess && cudaStreamWaitEvent(sb, ev[2], 0) == cudaSuccess;
    int rc_b = FVTA_OK;
    if (ok && rc == FVTA_OK) {
      // branch b: d kernel[:In] += x^T @ dsum ; d bias += colsum(dsum)
      rc_b = fvta_gemm(L.In, G, NT, x, L.In, 1, dsum, G, 0, nullptr, 0, d_kernel_fw, G, 1, sb);
      if (rc_b == FVTA_OK) {
        colsum_accumulate<<<cg, 256, 0, sb>>>(NT, G, dsum, G, d_bias_fw);
        rc_b = check_launch("colsum_accumulate");
      }
      ok = cudaEventRecord(ev[3], sb) == cudaSuccess;
      if (rc == FVTA_OK) rc = fvta_gemm(NT, L.In, G, dsum, G, 0, kernel_fw, G, 1, nullptr, 0, d_x, L.In, 0, stream);
    }
    // join (always: a forked stream left dangling would poison a capture)
    cudaStreamWaitEvent(st, ev[1], 0);
    if (ok) cudaStreamWaitEvent(st, ev[3], 0);
    if (!ok) { set_last_error("bilstm bwd fork/join", cudaGetLastError()); return FVTA_E_LAUNCH; }
    if (rc_a) return rc_a;
    if (rc_b) return rc_b;
    return rc;
  }
  // d_x = dg_fw @ Wx_fw^T + dg_bw @ Wx_bw^T
  rc = fvta_gemm(NT, L.In, G, dgs[0], G, 0, kernel_fw, G, 1, nullptr, 0, d_x, L.In, 0, stream);
  if (rc) return rc;
  rc = fvta_gemm(NT, L.In, G, dgs[1], G, 0, kernel_bw, G, 1, nullptr, 0, d_x, L.In, 1, stream);
  if (rc) return rc;
  float* dks[2] = {d_kernel_fw, d_kernel_bw};
  float* dbs[2] = {d_bias_fw, d_bias_bw};
  for (int dir = 0; dir < 2; ++dir) {
    // d kernel[:In] += x^T @ dg ;  d kernel[In:] += h_prev^T @ dg ;  d bias += colsum(dg)
    rc = fvta_gemm(L.In, G, NT, x, L.In, 1, dgs[dir], G, 0, nullptr, 0, dks[dir], G, 1, stream);
    if (rc) return rc;
    rc = fvta_gemm(L.d, G, NT, ress[dir] + 5 * L.d, kResW * L.d, 1, dgs[dir], G, 0, nullptr, 0,
                   dks[dir] + (size_t)L.In * G, G, 1, stream);
    if (rc) return rc;
    dim3 cg((G + 31) / 32, 64);
    colsum_accumulate<<<cg, 256, 0, st>>>(NT, G, dgs[dir], G, dbs[dir]);
    rc = check_launch("colsum_accumulate");
    if (rc) return rc;
  }
  return FVTA_OK;
}

// workspace: xp_fw [B*J*4d], xp_bw [B*J*4d] (forward) or dgates [2][B*J*4d] + their sum (backward)
extern "C" size_t fvta_bilstm_workspace_bytes(const FvtaLstmDesc* L) {
  if (lstm_validate(L)) return 0;
  return (size_t)3 * L->B * L->J * 4 * L->d * sizeof(float);
}
extern "C" size_t fvta_bilstm_reserve_floats(const FvtaLstmDesc* L) {
  if (lstm_validate(L)) return 0;
  return (size_t)2 * L->B * L->J * kResW * L->d;
}

extern "C" int fvta_bilstm_fwd(const FvtaLstmDesc* desc, const float* x, const int32_t* lens,
                               const float* kernel_fw, const float* bias_fw, const float* kernel_bw,
                               const float* bias_bw, float* out, float* last, float* reserve, void* workspace,
                               size_t workspace_bytes, void* stream) {
  int rc = lstm_validate(desc);
  if (rc) return rc;
  const FvtaLstmDesc L = *desc;
  if (!x || !lens || !kernel_fw || !bias_fw || !kernel_bw || !bias_bw || !out || !last || !reserve || !workspace)
    return FVTA_E_BADARG;
  if (workspace_bytes < fvta_bilstm_workspace_bytes(desc)) return FVTA_E_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)L.B * L.J * 2 * L.d * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) { set_last_error("bilstm memset", e); return FVTA_E_LAUNCH; }
  SeqAddr A{nullptr, nullptr, L.B * L.J};
  return lstm_fwd_impl(L, A, x, lens, kernel_fw, bias_fw, kernel_bw, bias_bw, out, last, reserve,
                       (float*)workspace, stream);
}

extern "C" int fvta_bilstm_bwd(const FvtaLstmDesc* desc, const float* x, const int32_t* lens,
                               const float* kernel_fw, const float* kernel_bw, const float* out,
                               const float* reserve, const float* d_out, const float* d_last, float* d_x,
                               float* d_kernel_fw, float* d_bias_fw, float* d_kernel_bw, float* d_bias_bw,
                               void* workspace, size_t workspace_bytes, void* stream) {
  (void)out;
  int rc = lstm_validate(desc);
  if (rc) return rc;
  const FvtaLstmDesc L = *desc;
  if (!x || !lens || !kernel_fw || !kernel_bw || !reserve || !d_x || !d_kernel_fw || !d_bias_fw ||
      !d_kernel_bw || !d_bias_bw || !workspace)
    return FVTA_E_BADARG;
  if (workspace_bytes < fvta_bilstm_workspace_bytes(desc)) return FVTA_E_WORKSPACE;
  SeqAddr A{nullptr, nullptr, L.B * L.J};
  return lstm_bwd_impl(L, A, x, lens, kernel_fw, kernel_bw, reserve, d_out, d_last, d_x, d_kernel_fw, d_bias_fw,
                       d_kernel_bw, d_bias_bw, (float*)workspace, stream);
}

// ---- packed reader ------------------------------------------------------------------------------
static int reader_validate(const FvtaReaderDesc* R) {
  if (!R) return FVTA_E_BADARG;
  if (R->B <= 0 || R->ntok <= 0 || R->In <= 0 || R->d <= 0 || R->Jmax <= 0) return FVTA_E_BADARG;
  return FVTA_OK;
}
extern "C" size_t fvta_reader_workspace_bytes(const FvtaReaderDesc* R) {
  if (reader_validate(R)) return 0;
  return (size_t)3 * R->ntok * 4 * R->d * sizeof(float);
}
extern "C" size_t fvta_reader_reserve_floats(const FvtaReaderDesc* R) {
  if (reader_validate(R)) return 0;
  return (size_t)2 * R->ntok * kResW * R->d;
}

extern "C" int fvta_reader_fwd(const FvtaReaderDesc* desc, const float* x_tok, const int32_t* seq_tok_off,
                               const int32_t* seq_len, const int64_t* seq_out_row, const float* kernel,
                               const float* bias, float* out_arena, float* last, float* reserve, void* workspace,
                               size_t workspace_bytes, void* stream) {
  int rc = reader_validate(desc);
  if (rc) return rc;
  if (!x_tok || !seq_tok_off || !seq_len || !seq_out_row || !kernel || !bias || !out_arena || !last || !reserve ||
      !workspace)
    return FVTA_E_BADARG;
  if (workspace_bytes < fvta_reader_workspace_bytes(desc)) return FVTA_E_WORKSPACE;
  FvtaLstmDesc L{desc->B, desc->Jmax, desc->In, desc->d, desc->path};
  SeqAddr A{seq_tok_off, seq_out_row, desc->ntok};
  return lstm_fwd_impl(L, A, x_tok, seq_len, kernel, bias, kernel, bias, out_arena, last, reserve,
                       (float*)workspace, stream);
}

extern "C" int fvta_reader_bwd(const FvtaReaderDesc* desc, const float* x_tok, const int32_t* seq_tok_off,
                               const int32_t* seq_len, const int64_t* seq_out_row, const float* kernel,
                               const float* reserve, const float* d_out_arena, const float* d_last, float* d_x_tok,
                               float* d_kernel, float* d_bias, void* workspace, size_t workspace_bytes,
                               void* stream) {
  int rc = reader_validate(desc);
  if (rc) return rc;
  if (!x_tok || !seq_tok_off || !seq_len || !seq_out_row || !kernel || !reserve || !d_x_tok || !d_kernel ||
      !d_bias || !workspace)
    return FVTA_E_BADARG;
  if (workspace_bytes < fvta_reader_workspace_bytes(desc)) return FVTA_E_WORKSPACE;
  FvtaLstmDesc L{desc->B, desc->Jmax, desc->In, desc->d, desc->path};
  SeqAddr A{seq_tok_off, seq_out_row, desc->ntok};
  return lstm_bwd_impl(L, A, x_tok, seq_len, kernel, kernel, reserve, d_out_arena, d_last, d_x_tok, d_kernel,
                       d_bias, d_kernel, d_bias, (float*)workspace, stream);
}
