// Tensor-core B operand images of the question (read by attn_fwd_tc.cu, built by the operand-folding
// kernel in attn_fwd.cu), pre-swizzled, one image per sample:
//   qB32: ceil(D/32) chunks x 64 rows x 128 B: rows 0-31 tf32(qeff[j]), rows 32-63 qeff[j] - tf32(qeff[j])
//   qB16: ceil(D/64) chunks x 32 rows x 128 B: fp16(qeff[j])
// 16-byte unit u of row n sits at unit (u ^ (n & 7)) of its 128-byte line (UMMA K-major, 128B swizzle).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace fvta {

__device__ __forceinline__ void build_tc_images_block(const FvtaAttnDesc& d, const float* __restrict__ qeffT_n,
                                                      uint8_t* __restrict__ qB32_n, uint8_t* __restrict__ qB16_n,
                                                      int Dp, int JQP, int part, int nparts) {
  // qeffT_n may live in shared memory; this block builds units i = part, part + nparts, ... of each image
  const int C32 = (d.D + 31) / 32, C16 = (d.D + 63) / 64;
  for (int i = part * blockDim.x + threadIdx.x; i < C32 * 64 * 8; i += nparts * blockDim.x) {
    const int u = i & 7, row = (i >> 3) & 63, c = i >> 9;
    const int j = row & 31;
    const bool lo = row >= 32;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int dd = c * 32 + u * 4 + e;
      const float x = (dd < d.D && j < d.JQ) ? qeffT_n[(size_t)dd * JQP + j] : 0.f;
      // round to nearest on the question side: the low 13 bits of `hi` are zero, so the tensor
      // core's own truncation leaves it alone, and x - hi is exact in fp32
      const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
      v[e] = lo ? x - hi : hi;
    }
    *reinterpret_cast<float4*>(qB32_n + c * 8192 + row * 128 + ((u ^ (row & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
  }
  for (int i = part * blockDim.x + threadIdx.x; i < C16 * 32 * 8; i += nparts * blockDim.x) {
    const int u = i & 7, row = (i >> 3) & 31, c = i >> 8;
    __half v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int dd = c * 64 + u * 8 + e;
      v[e] = __float2half_rn((dd < d.D && row < d.JQ) ? qeffT_n[(size_t)dd * JQP + row] : 0.f);
    }
    *reinterpret_cast<uint4*>(qB16_n + c * 4096 + row * 128 + ((u ^ (row & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
}

}  // namespace fvta
