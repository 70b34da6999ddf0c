// Pieces shared by the reader kernels (bilstm.cu: CUDA-core recurrences and BPTT; bilstm_tc.cu: the
// tcgen05 forward recurrence).
#pragma once
#include "common.cuh"

namespace fvta {

constexpr int kResW = 6;        // reserve floats per (dir, b, t) in units of d: i, j, f, o, c, h_prev

// Addressing of one sequence b: its tokens are rows tok0 .. tok0+len-1 of the packed arrays
// (projected inputs, reserve, d gates); its outputs go to rows out0 .. of the output arena
// (out0 < 0: outputs are not stored).  Dense [B,J] layout = (b*J, b*J).
struct SeqAddr {
  const int32_t* tok_off;   // [B] or null (dense)
  const int64_t* out_row;   // [B] or null (dense)
  int ntok;                 // rows of the packed arrays
  __device__ __forceinline__ int tok0(int b, int J) const { return tok_off ? tok_off[b] : b * J; }
  __device__ __forceinline__ long long out0(int b, int J) const { return out_row ? out_row[b] : (long long)b * J; }
};

// MUFU-based activations (ex2.approx / rcp.approx): |error| < 1e-6, an order of magnitude cheaper
// than expf / tanhf, which dominated the cell update
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx_(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcp_approx_(1.f + ex2_approx(2.885390081777927f * x)); }


}  // namespace fvta
