// Entry points declared in include/fvta_b200.h whose kernels are not written yet.
// Each returns FVTA_E_UNSUPPORTED; this file shrinks to nothing as they land.
#include "common.cuh"

extern "C" size_t fvta_bilstm_workspace_bytes(const FvtaLstmDesc*) { return 0; }
extern "C" size_t fvta_bilstm_reserve_floats(const FvtaLstmDesc*) { return 0; }
extern "C" int fvta_bilstm_fwd(const FvtaLstmDesc*, const float*, const int32_t*, const float*, const float*,
                               const float*, const float*, float*, float*, float*, void*, size_t, void*) {
  return FVTA_E_UNSUPPORTED;
}
extern "C" int fvta_bilstm_bwd(const FvtaLstmDesc*, const float*, const int32_t*, const float*, const float*,
                               const float*, const float*, const float*, const float*, float*, float*, float*,
                               float*, float*, void*, size_t, void*) { return FVTA_E_UNSUPPORTED; }
extern "C" int fvta_gemm(int32_t, int32_t, int32_t, const float*, int32_t, int32_t, const float*, int32_t,
                         int32_t, const float*, int32_t, float*, int32_t, int32_t, void*) {
  return FVTA_E_UNSUPPORTED;
}
