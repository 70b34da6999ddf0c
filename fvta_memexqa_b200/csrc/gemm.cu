// fp32 GEMM with fused bias / activation epilogue (replaces `linear`, model_v2.py:75-100, and
// the matmuls inside BasicLSTMCell).  CUDA-core FFMA so that results carry full fp32
// accuracy (the parity bar is 1e-4 relative through long recurrences); 64x64x16 block tile,
// 4x4 register micro-tile, operands staged in shared memory with the K index outermost so
// the inner product reads are conflict-free float4s.
//
//   C[M,N] (ldc) = act( op(A)[M,K] @ op(B)[K,N] + bias[N] )        accumulate == 0
//   C[M,N]      += op(A) @ op(B)                                   accumulate != 0
// op(X) = X stored row-major [rows, ld] when trans == 0, or its transpose when trans != 0.
#include "common.cuh"

namespace fvta {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                   const float* __restrict__ B, int ldb,
                                                   const float* __restrict__ bias, int act,
                                                   float* __restrict__ C, int ldc, int accumulate,
                                                   int k_per_split) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4x4 micro-tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // split-K (accumulate mode only): blockIdx.z owns K range [kbeg, kend) and adds atomically
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const bool atomic_out = gridDim.z > 1;

  // each thread moves 4 elements of the A tile and 4 of the B tile per k-tile; the next k-tile is
  // fetched into registers while the current one is multiplied (these products are short and
  // latency-bound, not FLOP-bound)
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + q * 256;
    if (TA) { am[q] = i % BM; ak[q] = i / BM; } else { ak[q] = i % BK; am[q] = i / BK; }
    if (TB) { bk[q] = i % BK; bn[q] = i / BK; } else { bn[q] = i % BN; bk[q] = i / BN; }
  }
  float ar[4], br[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gm = m0 + am[q], gka = k0 + ak[q];
      ar[q] = (gm < M && gka < kend) ? (TA ? A[(size_t)gka * lda + gm] : A[(size_t)gm * lda + gka]) : 0.f;
      const int gn = n0 + bn[q], gkb = k0 + bk[q];
      br[q] = (gn < N && gkb < kend) ? (TB ? B[(size_t)gn * ldb + gkb] : B[(size_t)gkb * ldb + gn]) : 0.f;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) { As[ak[q]][am[q]] = ar[q]; Bs[bk[q]][bn[q]] = br[q]; }
  };
  if (kbeg < kend) {
    gload(kbeg);
    sstore();
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (more) {
      sstore();
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float* c = C + (size_t)gm * ldc + gn;
      if (accumulate) {
        if (atomic_out) atomicAdd(c, v);
        else *c += v;
      } else {
        if (bias) v += bias[gn];
        if (act == 1) v = tanhf(v);
        else if (act == 2) v = fmaxf(v, 0.f);
        *c = v;
      }
    }
  }
}

// bias + activation over a finished split-K sum
__global__ void gemm_epilogue(int M, int N, float* __restrict__ C, int ldc, const float* __restrict__ bias, int act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float v = C[(size_t)m * ldc + n];
  if (bias) v += bias[n];
  if (act == 1) v = tanhf(v);
  else if (act == 2) v = fmaxf(v, 0.f);
  C[(size_t)m * ldc + n] = v;
}

}  // namespace fvta

extern "C" int fvta_gemm(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t trans_a,
                         const float* B, int32_t ldb, int32_t trans_b, const float* bias, int32_t act,
                         float* C, int32_t ldc, int32_t accumulate, void* stream) {
  using namespace fvta;
  if (M < 0 || N < 0 || K < 0 || !C) return FVTA_E_BADARG;
  if (M == 0 || N == 0) return FVTA_OK;
  if (K > 0 && (!A || !B)) return FVTA_E_BADARG;
  if (act < 0 || act > 2) return FVTA_E_UNSUPPORTED;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  if (grid.y > 65535) return FVTA_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  // Weight-gradient GEMMs are short and fat (M, N ~ 200, K = tokens): split K over the SMs.
  // Short-and-fat products (weight gradients: M, N ~ 200, K = tokens; the image transform:
  // M = photos, K = 2537) leave most SMs idle with one CTA per output tile: split K over the SMs
  // and add the partial tiles atomically.  A non-accumulating call zeroes C first and applies
  // bias / activation in a second tiny kernel.
  int kps = K > 0 ? K : 1;
  bool split_epilogue = false;
  if (K >= 128) {
    const int tiles = (int)(grid.x * grid.y);
    int splits = (148 + tiles - 1) / tiles;          // about one wave; more only adds atomic traffic
    const int max_splits = (K + 63) / 64;            // at least four k-tiles per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits > 1) {
      kps = round_up((K + splits - 1) / splits, BK);
      grid.z = (K + kps - 1) / kps;
      if (!accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st);
        if (e != cudaSuccess) { set_last_error("gemm memset", e); return FVTA_E_LAUNCH; }
        split_epilogue = (bias != nullptr) || act != 0;
        accumulate = 1;
      }
    }
  }
  const float* kbias = split_epilogue ? nullptr : bias;
  const int kact = split_epilogue ? 0 : act;
  if (!trans_a && !trans_b) gemm_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, kbias, kact, C, ldc, accumulate, kps);
  else if (!trans_a && trans_b) gemm_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, kbias, kact, C, ldc, accumulate, kps);
  else if (trans_a && !trans_b) gemm_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, kbias, kact, C, ldc, accumulate, kps);
  else gemm_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, kbias, kact, C, ldc, accumulate, kps);
  int rc = check_launch("gemm_kernel");
  if (rc || !split_epilogue) return rc;
  gemm_epilogue<<<(M * N + 255) / 256, 256, 0, st>>>(M, N, C, ldc, bias, act);
  return check_launch("gemm_epilogue");
}
