// Adadelta over one flat fp32 buffer (replaces tf.train.AdadeltaOptimizer.apply_gradients,
// reference trainer.py:16-27; TF recurrence).  HBM-bound: 4 reads +
// 3 writes of 4 B per parameter; float4-vectorised, grid-stride.
//
// Embedding tables (model_v2.py:534, :596-600) are read through tf.nn.embedding_lookup, so their
// gradients are IndexedSlices and TF applies sparse_apply_adadelta [TF]: rows that no id of the
// batch looked up keep var, accum AND accum_update unchanged (a dense update would decay both
// accumulators by rho).  `touch_index[i]` >= 0 names the entry of `touch` that says whether the
// row holding element i was looked up this step (> 0); -1 = ordinary dense variable.
#include "common.cuh"
#include "adadelta.cuh"

namespace fvta {

__global__ void adadelta_kernel(int64_t n, float* __restrict__ var, const float* __restrict__ grad,
                                float* __restrict__ accum, float* __restrict__ accum_update,
                                const float* __restrict__ mask, const int32_t* __restrict__ touch_index,
                                const float* __restrict__ touch, float lr, float rho, float eps, float gs) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(var)[i];
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    float4 a = reinterpret_cast<float4*>(accum)[i];
    float4 u = reinterpret_cast<float4*>(accum_update)[i];
    float4 m = mask ? reinterpret_cast<const float4*>(mask)[i] : make_float4(1.f, 1.f, 1.f, 1.f);
    if (touch_index != nullptr) {
      const int4 ti = reinterpret_cast<const int4*>(touch_index)[i];
      if (ti.x >= 0 && !(touch[ti.x] > 0.f)) m.x = 0.f;
      if (ti.y >= 0 && !(touch[ti.y] > 0.f)) m.y = 0.f;
      if (ti.z >= 0 && !(touch[ti.z] > 0.f)) m.z = 0.f;
      if (ti.w >= 0 && !(touch[ti.w] > 0.f)) m.w = 0.f;
    }
    adadelta_one(v.x, g.x, a.x, u.x, m.x, lr, rho, eps, gs);
    adadelta_one(v.y, g.y, a.y, u.y, m.y, lr, rho, eps, gs);
    adadelta_one(v.z, g.z, a.z, u.z, m.z, lr, rho, eps, gs);
    adadelta_one(v.w, g.w, a.w, u.w, m.w, lr, rho, eps, gs);
    reinterpret_cast<float4*>(var)[i] = v;
    reinterpret_cast<float4*>(accum)[i] = a;
    reinterpret_cast<float4*>(accum_update)[i] = u;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = var[i], a = accum[i], u = accum_update[i];
    float m = mask ? mask[i] : 1.f;
    if (touch_index != nullptr && touch_index[i] >= 0 && !(touch[touch_index[i]] > 0.f)) m = 0.f;
    adadelta_one(v, grad[i], a, u, m, lr, rho, eps, gs);
    var[i] = v; accum[i] = a; accum_update[i] = u;
  }
}

}  // namespace fvta

extern "C" int fvta_adadelta(int64_t n, float* var, const float* grad, float* accum, float* accum_update,
                             const float* update_mask, const int32_t* touch_index, const float* touch,
                             float lr, float rho, float eps, float grad_scale, void* stream) {
  if (n < 0 || (n > 0 && (!var || !grad || !accum || !accum_update))) return FVTA_E_BADARG;
  if ((touch_index == nullptr) != (touch == nullptr)) return FVTA_E_BADARG;
  if (n == 0) return FVTA_OK;
  const int threads = 256;
  int64_t want = (n / 4 + threads - 1) / threads;
  int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  fvta::adadelta_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(n, var, grad, accum, accum_update,
                                                                     update_mask, touch_index, touch, lr, rho, eps,
                                                                     grad_scale);
  return fvta::check_launch("adadelta_kernel");
}
