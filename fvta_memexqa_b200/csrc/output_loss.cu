// Output layer + softmax + cross-entropy, forward AND backward in one launch
// (replaces model_v2.py:1061-1083 `linear` over [gq, g1, gc, g1*gc, gq*gc] -> logits -> softmax,
// and :1088-1090 mean softmax-CE).  The loss is the end of the graph, so d loss / d logits is
// known the moment the logits are: one block per QA computes logits, yp, its CE term and the
// gradients w.r.t. gq, g1_all, the choice encodings and (atomically) W, b.
//
//   x[n,c,:]  = [gq, g1, gc, g1*gc, gq*gc]      (5 blocks of D)        W = (W0..W4), each [D]
//   logit[n,c]= sum_d gq*W0 + g1*W1 + gc*(W2 + g1*W3 + gq*W4) + b
//   loss      = (1/N) sum_n -sum_c y log softmax(logit)                (padded rows: y == 0 -> 0)
//   dlogit    = (p * sum_c y - y) / N
#include "common.cuh"

namespace fvta {

constexpr int kMaxChoice = 8;

__global__ void output_loss_kernel(int N, int C, int D, const float* __restrict__ gq, const float* __restrict__ g1,
                                   const float* __restrict__ gc, const float* __restrict__ W,
                                   const float* __restrict__ b, const uint8_t* __restrict__ y,
                                   float* __restrict__ logits, float* __restrict__ yp, float* __restrict__ loss,
                                   float* __restrict__ d_gq, float* __restrict__ d_g1, float* __restrict__ d_gc,
                                   float* __restrict__ dW, float* __restrict__ db, int want_grad) {
  __shared__ float red[kMaxChoice][33];
  __shared__ float dl_s[kMaxChoice];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float* q = gq + (size_t)n * D;
  const float* a = g1 + (size_t)n * D;
  float part[kMaxChoice];
#pragma unroll
  for (int c = 0; c < kMaxChoice; ++c) part[c] = 0.f;
  for (int d0 = threadIdx.x; d0 < D; d0 += blockDim.x) {
    const float qv = q[d0], av = a[d0];
    const float base = qv * W[d0] + av * W[D + d0];
    const float wc = W[2 * D + d0] + av * W[3 * D + d0] + qv * W[4 * D + d0];
#pragma unroll
    for (int c = 0; c < kMaxChoice; ++c)
      if (c < C) part[c] += base + gc[((size_t)n * C + c) * D + d0] * wc;
  }
#pragma unroll
  for (int c = 0; c < kMaxChoice; ++c) {
    const float v = warp_sum(part[c]);
    if (lane == 0) red[c][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lg[kMaxChoice], mx = -INFINITY, sy = 0.f;
    for (int c = 0; c < C; ++c) {
      float s = b[0];
      for (int w = 0; w < nwarp; ++w) s += red[c][w];
      lg[c] = s;
      mx = fmaxf(mx, s);
      logits[(size_t)n * C + c] = s;
      sy += y ? (float)y[(size_t)n * C + c] : 0.f;
    }
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(lg[c] - mx);
    const float lse = mx + logf(den);
    float ce = 0.f, dbs = 0.f;
    for (int c = 0; c < C; ++c) {
      const float p = expf(lg[c] - lse);
      yp[(size_t)n * C + c] = p;
      const float yc = y ? (float)y[(size_t)n * C + c] : 0.f;
      ce -= yc * (lg[c] - lse);
      const float dl = (p * sy - yc) / (float)N;
      dl_s[c] = dl;
      dbs += dl;
    }
    if (loss) atomicAdd(loss, ce / (float)N);
    if (want_grad && db) atomicAdd(db, dbs);
  }
  if (!want_grad) return;
  __syncthreads();
  for (int d0 = threadIdx.x; d0 < D; d0 += blockDim.x) {
    const float qv = q[d0], av = a[d0];
    const float w0 = W[d0], w1 = W[D + d0], w2 = W[2 * D + d0], w3 = W[3 * D + d0], w4 = W[4 * D + d0];
    const float wc = w2 + av * w3 + qv * w4;
    float sdl = 0.f, sdlc = 0.f;   // sum_c dl, sum_c dl*gc
#pragma unroll
    for (int c = 0; c < kMaxChoice; ++c)
      if (c < C) {
        const float dl = dl_s[c];
        const float g = gc[((size_t)n * C + c) * D + d0];
        sdl += dl;
        sdlc += dl * g;
        d_gc[((size_t)n * C + c) * D + d0] = dl * wc;
      }
    d_gq[(size_t)n * D + d0] = sdl * w0 + sdlc * w4;
    d_g1[(size_t)n * D + d0] = sdl * w1 + sdlc * w3;
    atomicAdd(dW + d0, sdl * qv);
    atomicAdd(dW + D + d0, sdl * av);
    atomicAdd(dW + 2 * D + d0, sdlc);
    atomicAdd(dW + 3 * D + d0, sdlc * av);
    atomicAdd(dW + 4 * D + d0, sdlc * qv);
  }
}

}  // namespace fvta

extern "C" int fvta_output_loss(int32_t N, int32_t C, int32_t D, const float* gq, const float* g1, const float* gc,
                                const float* W, const float* b, const uint8_t* y, float* logits, float* yp,
                                float* loss, float* d_gq, float* d_g1, float* d_gc, float* dW, float* db,
                                void* stream) {
  using namespace fvta;
  if (N <= 0 || C <= 0 || D <= 0 || !gq || !g1 || !gc || !W || !b || !logits || !yp) return FVTA_E_BADARG;
  if (C > kMaxChoice) return FVTA_E_SHAPE;
  const int want_grad = (d_gq && d_g1 && d_gc && dW && db) ? 1 : 0;
  output_loss_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(N, C, D, gq, g1, gc, W, b, y, logits, yp, loss, d_gq, d_g1,
                                                         d_gc, dW, db, want_grad);
  return check_launch("output_loss_kernel");
}
