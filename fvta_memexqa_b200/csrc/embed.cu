// Token-embedding plumbing of the reader input (model_v2.py:524-620) and the bias / activation
// part of a linear layer's backward (model_v2.py:75-100), as single launches.
//
//   fvta_embed_rows      out[i,:] = id < VW ? table_new[id] : table_fixed[id - VW]   (:596 concat of
//                        the trainable and the fed GloVe table, then embedding_lookup), written
//                        with a row stride so it lands next to the char-CNN features (:617 concat)
//   fvta_embed_rows_bwd  d table_new[id] += dy[i,:] for id < VW (scatter-add, atomics)
//   fvta_bias_act_grad   dz = dy * act'(y),  db += column sums of dz
#include "common.cuh"

namespace fvta {

constexpr int kEmbRowsPerBlock = 8;

// grid ceil(ntok / 8), block (32, 8): one warp per token
__global__ void embed_rows_kernel(int ntok, int dim, int VW, int VF, const int64_t* __restrict__ ids,
                                  const float* __restrict__ table_new, const float* __restrict__ table_fixed,
                                  float* __restrict__ out, int ld) {
  const int w = blockIdx.x * kEmbRowsPerBlock + threadIdx.y;
  if (w >= ntok) return;
  const long long id = ids[w];
  const float* src;
  if (id < VW) {
    src = table_new + (size_t)(id < 0 ? 0 : id) * dim;
  } else {
    long long r = id - VW;
    if (r >= VF) r = VF - 1;
    src = table_fixed + (size_t)r * dim;
  }
  float* dst = out + (size_t)w * ld;
  if (((dim | ld) & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    for (int c = threadIdx.x; c < (dim >> 2); c += 32)
      reinterpret_cast<float4*>(dst)[c] = reinterpret_cast<const float4*>(src)[c];
  } else {
    for (int c = threadIdx.x; c < dim; c += 32) dst[c] = src[c];
  }
}

__global__ void embed_rows_bwd_kernel(int ntok, int dim, int VW, const int64_t* __restrict__ ids,
                                      const float* __restrict__ dy, int ld, float* __restrict__ d_table) {
  const int w = blockIdx.x * kEmbRowsPerBlock + threadIdx.y;
  if (w >= ntok) return;
  const long long id = ids[w];
  if (id >= VW) return;   // the fed table is not trained
  float* dst = d_table + (size_t)(id < 0 ? 0 : id) * dim;
  const float* src = dy + (size_t)w * ld;
  for (int c = threadIdx.x; c < dim; c += 32) atomicAdd(dst + c, src[c]);
}

// grid (ceil(N / 32), row blocks), block 256 = 8 row lanes x 32 columns
__global__ void bias_act_grad_kernel(int M, int N, const float* __restrict__ dy, int ld_dy,
                                     const float* __restrict__ y, int ld_y, int act, float* __restrict__ dz,
                                     int ld_dz, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (n < N)
    for (int m = blockIdx.y * 8 + w; m < M; m += gridDim.y * 8) {
      float g = dy[(size_t)m * ld_dy + n];
      if (act != 0) {
        const float yv = y[(size_t)m * ld_y + n];
        g = (act == 1) ? g * (1.f - yv * yv) : (yv > 0.f ? g : 0.f);
        dz[(size_t)m * ld_dz + n] = g;
      }
      acc += g;
    }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && n < N && db != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][lane];
    atomicAdd(db + n, s);
  }
}

}  // namespace fvta

using namespace fvta;

extern "C" int fvta_embed_rows(int32_t ntok, int32_t dim, int32_t VW, int32_t VF, const int64_t* ids,
                               const float* table_new, const float* table_fixed, float* out, int32_t ld_out,
                               void* stream) {
  if (ntok < 0 || dim <= 0 || VW < 0 || VF < 0 || ld_out < dim) return FVTA_E_BADARG;
  if (ntok == 0) return FVTA_OK;
  if (!ids || !out || (VW > 0 && !table_new) || (VF > 0 && !table_fixed) || VW + VF == 0) return FVTA_E_BADARG;
  if (VW == 0) table_new = table_fixed;     // never dereferenced with VW == 0 unless an id is negative
  if (VF == 0) { table_fixed = table_new; VF = VW; }
  dim3 block(32, kEmbRowsPerBlock);
  embed_rows_kernel<<<(ntok + kEmbRowsPerBlock - 1) / kEmbRowsPerBlock, block, 0, (cudaStream_t)stream>>>(
      ntok, dim, VW, VF, ids, table_new, table_fixed, out, ld_out);
  return check_launch("embed_rows_kernel");
}

extern "C" int fvta_embed_rows_bwd(int32_t ntok, int32_t dim, int32_t VW, const int64_t* ids, const float* dy,
                                   int32_t ld_dy, float* d_table_new, void* stream) {
  if (ntok < 0 || dim <= 0 || VW < 0 || ld_dy < dim) return FVTA_E_BADARG;
  if (ntok == 0 || VW == 0) return FVTA_OK;
  if (!ids || !dy || !d_table_new) return FVTA_E_BADARG;
  dim3 block(32, kEmbRowsPerBlock);
  embed_rows_bwd_kernel<<<(ntok + kEmbRowsPerBlock - 1) / kEmbRowsPerBlock, block, 0, (cudaStream_t)stream>>>(
      ntok, dim, VW, ids, dy, ld_dy, d_table_new);
  return check_launch("embed_rows_bwd_kernel");
}

extern "C" int fvta_bias_act_grad(int32_t M, int32_t N, const float* dy, int32_t ld_dy, const float* y,
                                  int32_t ld_y, int32_t act, float* dz, int32_t ld_dz, float* db, void* stream) {
  if (M < 0 || N <= 0 || ld_dy < N || act < 0 || act > 2) return FVTA_E_BADARG;
  if (M == 0) return FVTA_OK;
  if (!dy) return FVTA_E_BADARG;
  if (act != 0 && (!y || !dz || ld_y < N || ld_dz < N)) return FVTA_E_BADARG;
  if (act == 0 && !db) return FVTA_OK;      // nothing to do
  int rb = (M + 63) / 64;
  if (rb > 64) rb = 64;
  dim3 grid((N + 31) / 32, rb);
  bias_act_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, dy, ld_dy, y, ld_y, act, dz, ld_dz, db);
  return check_launch("bias_act_grad_kernel");
}
