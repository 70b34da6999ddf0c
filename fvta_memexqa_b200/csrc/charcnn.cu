// Character CNN over words (replaces conv1d, model_v2.py:52-70: 1 x height VALID conv over the
// embedded characters, + bias, ReLU, max over positions) for the undropped case.
//
// The conv output at a position depends only on the `height` character IDS in its window:
//   out[w,p,:] = bias + sum_h  char_emb[id[w,p+h]] @ filter[h]      = bias + sum_h table[id[w,p+h]][h][:]
// with table[:, h, :] = char_emb @ filter[h]  ([VC,cdim] x [cdim,cw], a few MFLOP, done by the caller
// as ONE product char_emb @ [filter[0] | ... | filter[height-1]], hence the [VC, height, cw] layout).
// So instead of a [words*positions, height*cdim] x [height*cdim, cw] GEMM (4 GFLOP per step at
// the reference shape, the largest FLOP term of the whole model) the layer is a gather-add over
// an L2-resident table: 5 loads per output.  Same function; only the fp32 summation order
// differs (per-tap partial sums are formed first).
//
// Backward routes d y through the max (first arg-max position; ReLU-dead outputs get nothing)
// into d table with atomics; the caller turns d table into d char_emb / d filter with the same
// tiny matrix products.
#include "common.cuh"

namespace fvta {

constexpr int kCnnWordsPerBlock = 8;

// grid ceil(ntok / 8), block 8 x 32: lane = group of 4 output channels, threadIdx.y = word
__global__ void charcnn_fwd_kernel(int ntok, int W, int VC, int cw, int height, const int32_t* __restrict__ ids,
                                   const float* __restrict__ table, const float* __restrict__ bias,
                                   float* __restrict__ y, int ldy, uint8_t* __restrict__ argpos, int vec) {
  __shared__ int id_s[kCnnWordsPerBlock][64];
  const int wl = threadIdx.y, lane = threadIdx.x;
  const int w = blockIdx.x * kCnnWordsPerBlock + wl;
  if (w < ntok)
    for (int i = lane; i < W; i += 32) {
      const int v = ids[(size_t)w * W + i];
      id_s[wl][i] = (v >= 0 && v < VC) ? v : 0;
    }
  __syncthreads();
  if (w >= ntok) return;
  const int npos = W - height + 1;
  if (!vec) {   // any cw / row stride: one output channel per lane
    for (int o = lane; o < cw; o += 32) {
      const float b = bias[o];
      float best = -INFINITY;
      int bp = 0;
      for (int p = 0; p < npos; ++p) {
        float a = b;
        for (int h = 0; h < height; ++h) a += table[((size_t)id_s[wl][p + h] * height + h) * cw + o];
        if (a > best) { best = a; bp = p; }
      }
      y[(size_t)w * ldy + o] = fmaxf(best, 0.f);
      argpos[(size_t)w * cw + o] = (unsigned char)bp;
    }
    return;
  }
  const int cw4 = cw >> 2;
  for (int o4 = lane; o4 < cw4; o4 += 32) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + o4 * 4);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int bp0 = 0, bp1 = 0, bp2 = 0, bp3 = 0;
    for (int p = 0; p < npos; ++p) {
      float4 a = b4;
      for (int h = 0; h < height; ++h) {
        const float4 t = *reinterpret_cast<const float4*>(table + ((size_t)id_s[wl][p + h] * height + h) * cw + o4 * 4);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      if (a.x > best.x) { best.x = a.x; bp0 = p; }
      if (a.y > best.y) { best.y = a.y; bp1 = p; }
      if (a.z > best.z) { best.z = a.z; bp2 = p; }
      if (a.w > best.w) { best.w = a.w; bp3 = p; }
    }
    // relu(max_p) == max_p relu
    const float4 out = make_float4(fmaxf(best.x, 0.f), fmaxf(best.y, 0.f), fmaxf(best.z, 0.f), fmaxf(best.w, 0.f));
    *reinterpret_cast<float4*>(y + (size_t)w * ldy + o4 * 4) = out;
    *reinterpret_cast<uchar4*>(argpos + (size_t)w * cw + o4 * 4) =
        make_uchar4((unsigned char)bp0, (unsigned char)bp1, (unsigned char)bp2, (unsigned char)bp3);
  }
}

__global__ void charcnn_bwd_kernel(int ntok, int W, int VC, int cw, int height, const int32_t* __restrict__ ids,
                                   const float* __restrict__ y, int ldy, const uint8_t* __restrict__ argpos,
                                   const float* __restrict__ dy, int lddy, float* __restrict__ d_table,
                                   float* __restrict__ d_bias) {
  __shared__ int id_s[kCnnWordsPerBlock][64];
  __shared__ float db_s[kCnnWordsPerBlock][129];
  const int wl = threadIdx.y, lane = threadIdx.x;
  const int w = blockIdx.x * kCnnWordsPerBlock + wl;
  if (w < ntok)
    for (int i = lane; i < W; i += 32) {
      const int v = ids[(size_t)w * W + i];
      id_s[wl][i] = (v >= 0 && v < VC) ? v : 0;
    }
  __syncthreads();
  for (int o = lane; o < cw; o += 32) {
    float g = 0.f;
    if (w < ntok && y[(size_t)w * ldy + o] > 0.f) {   // ReLU alive
      g = dy[(size_t)w * lddy + o];
      const int p = argpos[(size_t)w * cw + o];
      for (int h = 0; h < height; ++h) atomicAdd(d_table + ((size_t)id_s[wl][p + h] * height + h) * cw + o, g);
    }
    if (o < 128) db_s[wl][o] = g;
  }
  __syncthreads();
  if (wl == 0)
    for (int o = lane; o < cw && o < 128; o += 32) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kCnnWordsPerBlock; ++i) s += db_s[i][o];
      if (s != 0.f) atomicAdd(d_bias + o, s);
    }
}

// Channel-major backward: a block owns FOUR output channels and a slice of the words, and accumulates
// d table[:, :, 4 o4 .. 4 o4 + 3] in shared memory (8 copies against collisions between warps), so the
// 5 x ntok x cw scattered additions never leave the SM; the flush is VC x height x 4 global atomics per
// block, contended only between the slices.  (The word-major kernel above issues every addition as a
// global atomic on a 160 KB table: 38 us at C2, 7 % issue-active.)
constexpr int kCbThreads = 1024, kCbCopies = 8, kCbSlices = 5;

__global__ void __launch_bounds__(kCbThreads, 1) charcnn_bwd_chan4(
    int ntok, int W, int VC, int cw, int height, const int32_t* __restrict__ ids, const float* __restrict__ y, int ldy,
    const uint8_t* __restrict__ argpos, const float* __restrict__ dy, int lddy, float* __restrict__ d_table,
    float* __restrict__ d_bias) {
  extern __shared__ float acc[];                 // [copies][VC * height][4]
  __shared__ float db_s[4];
  const int o4 = blockIdx.x, slice = blockIdx.y;
  const int per = VC * height * 4;
  for (int i = threadIdx.x; i < kCbCopies * per; i += blockDim.x) acc[i] = 0.f;
  if (threadIdx.x < 4) db_s[threadIdx.x] = 0.f;
  __syncthreads();
  float* mine = acc + ((threadIdx.x >> 5) & (kCbCopies - 1)) * per;
  float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = slice * blockDim.x + threadIdx.x; t < ntok; t += gridDim.y * blockDim.x) {
    const float4 yv = *reinterpret_cast<const float4*>(y + (size_t)t * ldy + 4 * o4);
    const float4 g4 = *reinterpret_cast<const float4*>(dy + (size_t)t * lddy + 4 * o4);
    const uchar4 ap = *reinterpret_cast<const uchar4*>(argpos + (size_t)t * cw + 4 * o4);
    const float g[4] = {yv.x > 0.f ? g4.x : 0.f, yv.y > 0.f ? g4.y : 0.f, yv.z > 0.f ? g4.z : 0.f, yv.w > 0.f ? g4.w : 0.f};
    const int pp[4] = {ap.x, ap.y, ap.z, ap.w};
    db.x += g[0]; db.y += g[1]; db.z += g[2]; db.w += g[3];
    const int32_t* idr = ids + (size_t)t * W;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (g[c] == 0.f) continue;                 // dead ReLU / padding word
      for (int h = 0; h < height; ++h) {
        int v = idr[pp[c] + h];
        v = (v >= 0 && v < VC) ? v : 0;
        atomicAdd(mine + (v * height + h) * 4 + c, g[c]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    db.x += __shfl_xor_sync(0xffffffffu, db.x, o); db.y += __shfl_xor_sync(0xffffffffu, db.y, o);
    db.z += __shfl_xor_sync(0xffffffffu, db.z, o); db.w += __shfl_xor_sync(0xffffffffu, db.w, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (db.x != 0.f) atomicAdd(&db_s[0], db.x);
    if (db.y != 0.f) atomicAdd(&db_s[1], db.y);
    if (db.z != 0.f) atomicAdd(&db_s[2], db.z);
    if (db.w != 0.f) atomicAdd(&db_s[3], db.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    float sacc = 0.f;
#pragma unroll
    for (int k = 0; k < kCbCopies; ++k) sacc += acc[k * per + i];
    if (sacc != 0.f) atomicAdd(d_table + (size_t)(i >> 2) * cw + 4 * o4 + (i & 3), sacc);
  }
  if (threadIdx.x < 4 && db_s[threadIdx.x] != 0.f) atomicAdd(d_bias + 4 * o4 + threadIdx.x, db_s[threadIdx.x]);
}

}  // namespace fvta

extern "C" int fvta_charcnn_fwd(int32_t ntok, int32_t W, int32_t VC, int32_t cw, int32_t height,
                                const int32_t* ids, const float* table, const float* bias, float* y,
                                int32_t ldy, uint8_t* argpos, void* stream) {
  using namespace fvta;
  if (ntok < 0 || W <= 0 || VC <= 0 || cw <= 0 || height <= 0) return FVTA_E_BADARG;
  if (ntok == 0) return FVTA_OK;
  if (!ids || !table || !bias || !y || !argpos) return FVTA_E_BADARG;
  if (cw > 128 || W > 64 || W < height || W - height + 1 > 255) return FVTA_E_SHAPE;
  if (ldy < cw) return FVTA_E_BADARG;
  // float4 path when the channel count, the row stride and every base pointer allow it
  const int vec = ((cw & 3) == 0 && (ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(table) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(argpos) & 3) == 0) ? 1 : 0;
  dim3 block(32, kCnnWordsPerBlock);
  charcnn_fwd_kernel<<<(ntok + kCnnWordsPerBlock - 1) / kCnnWordsPerBlock, block, 0, (cudaStream_t)stream>>>(
      ntok, W, VC, cw, height, ids, table, bias, y, ldy, argpos, vec);
  return check_launch("charcnn_fwd_kernel");
}

extern "C" int fvta_charcnn_bwd(int32_t ntok, int32_t W, int32_t VC, int32_t cw, int32_t height,
                                const int32_t* ids, const float* y, int32_t ldy, const uint8_t* argpos,
                                const float* dy, int32_t lddy, float* d_table, float* d_bias, void* stream) {
  using namespace fvta;
  if (ntok < 0 || W <= 0 || VC <= 0 || cw <= 0 || height <= 0) return FVTA_E_BADARG;
  if (ntok == 0) return FVTA_OK;
  if (!ids || !y || !argpos || !dy || !d_table || !d_bias) return FVTA_E_BADARG;
  if (cw > 128 || W > 64 || W < height) return FVTA_E_SHAPE;
  if (ldy < cw || lddy < cw) return FVTA_E_BADARG;
  const size_t chan_smem = (size_t)kCbCopies * VC * height * 4 * sizeof(float);
  const bool vec = (cw & 3) == 0 && (ldy & 3) == 0 && (lddy & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(argpos) & 3) == 0;
  if (vec && chan_smem <= 200 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(charcnn_bwd_chan4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chan_smem);
    if (e != cudaSuccess) { set_last_error("charcnn_bwd_chan4 smem attr", e); return FVTA_E_LAUNCH; }
    const int slices = ntok >= kCbSlices * kCbThreads ? kCbSlices : (ntok + kCbThreads - 1) / kCbThreads;
    dim3 grid(cw / 4, slices < 1 ? 1 : slices);
    charcnn_bwd_chan4<<<grid, kCbThreads, chan_smem, (cudaStream_t)stream>>>(ntok, W, VC, cw, height, ids, y, ldy, argpos,
                                                                           dy, lddy, d_table, d_bias);
    return check_launch("charcnn_bwd_chan4");
  }
  dim3 block(32, kCnnWordsPerBlock);
  charcnn_bwd_kernel<<<(ntok + kCnnWordsPerBlock - 1) / kCnnWordsPerBlock, block, 0, (cudaStream_t)stream>>>(
      ntok, W, VC, cw, height, ids, y, ldy, argpos, dy, lddy, d_table, d_bias);
  return check_launch("charcnn_bwd_kernel");
}
