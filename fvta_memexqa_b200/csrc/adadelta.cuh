// The Adadelta recurrence of tf.train.AdadeltaOptimizer (trainer.py:16-27), shared by
// the plain optimizer kernel (adadelta.cu) and the fused all-reduce + optimizer kernel (p2p_adadelta.cu).
#pragma once

namespace fvta {

__device__ __forceinline__ void adadelta_one(float& v, float g, float& a, float& au, float mask, float lr,
                                             float rho, float eps, float gs) {
  g *= gs;
  const float na = rho * a + (1.f - rho) * g * g;
  const float upd = sqrtf(au + eps) * rsqrtf(na + eps) * g;
  const float nau = rho * au + (1.f - rho) * upd * upd;
  if (mask != 0.f) {
    a = na;
    v = v - lr * upd;
    au = nau;
  }
}

}  // namespace fvta
