// fp32 GEMM with fused bias / activation epilogue (replaces `linear`, model_v2.py:75-100, and
// the matmuls inside BasicLSTMCell).  64x64x16 block tile, operands staged in shared memory with
// the K index outermost.  Two inner products over the same staging:
//   gemm_mma_kernel  (default) tensor cores with split operands: every fp32 operand is split, as its
//                    fragment is read from shared memory, into tf32 hi + tf32 lo
//                    (x = hi + lo to ~22 bits) and the product is formed as
//                    lo.hi + hi.lo + hi.hi with fp32 accumulation (mma.sync m16n8k8) -- fp32-grade
//                    accuracy (the parity bar is 1e-4 relative through long recurrences) at a
//                    third of the CUDA-core instruction count.  These products are a few hundred
//                    rows/columns: launch- and latency-bound, far too small to feed a tcgen05
//                    pipeline (a 128 x N tile per SM would leave most SMs idle).
//   gemm_kernel      CUDA-core FFMA, 4x4 register micro-tile (FVTA_GEMM_FFMA=1; kept as the A/B
//                    reference).
//
//   C[M,N] (ldc) = act( op(A)[M,K] @ op(B)[K,N] + bias[N] )        accumulate == 0
//   C[M,N]      += op(A) @ op(B)                                   accumulate != 0
// op(X) = X stored row-major [rows, ld] when trans == 0, or its transpose when trans != 0.
#include "common.cuh"
#include <cstdlib>

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

// ---- tensor-core inner product with split tf32 operands ------------------------------------------
// Split with integer rounding: hi = x rounded to tf32 (add half an ulp to the bit pattern, clear the
// 13 low mantissa bits), x - hi is exact in fp32, lo = that residual rounded the same way, so
// x = hi + lo to ~22 bits with unbiased errors -- five instructions per element (cvt.rna.tf32 is a
// long multi-instruction sequence on sm_100).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Operand staging: a 4-stage cp.async ring of raw fp32 tiles (these products are a few hundred
// rows/columns with K ~ 200: what bounds them is the global-load round trip per k-tile, so the
// loads of three tiles are always in flight).  Each operand keeps its global orientation in shared
// memory, so a tile is 256 16-byte copies -- one per thread -- when the pointer and leading
// dimension allow (VEC = 4), or 1024 4-byte copies (VEC = 1).  Row strides are chosen so that the
// m16n8k8 fragment loads hit 32 distinct banks:
//   k contiguous   (A not transposed / B transposed): [64][20]  bank = 20 g + t  (+ const)
//   m/n contiguous (A transposed / B not transposed): [16][72]  bank =  8 t + g  (+ const)
constexpr int kStages = 4;
constexpr int kLdK = BK + 4;   // 20
constexpr int kLdM = BM + 8;   // 72
constexpr int kOpFloats = BM * kLdK > BK * kLdM ? BM * kLdK : BK * kLdM;   // 1280

__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(src_bytes) : "memory");
}

// One operand tile -> shared memory.  `rows` x `kk`: the operand is X[r][k] (r = m or n) stored
// either with k contiguous (KMAJ: X + r * ld + k) or with r contiguous (X + k * ld + r).
// `rowmap` (gather variant, null otherwise): the index that multiplies `ld` -- the storage row of X --
// is looked up, rowmap[row], so the operand is a set of rows of a larger table read in place (the image
// transform multiplies the batch's photo features image_emb_mat[pis], model_v2.py:634-645, without
// materialising the gathered [tokens, 2537] matrix).
template <bool KMAJ, int VEC>
__device__ __forceinline__ void stage_operand(float* dst, const float* __restrict__ X, int ld, int r0, int R, int k0,
                                              int kend, int tid, const int64_t* __restrict__ rowmap = nullptr) {
  if (VEC == 4) {
    if (KMAJ) {   // 64 rows x 4 chunks of 4 k
      const int r = tid >> 2, kq = (tid & 3) * 4;
      const int gr = r0 + r, gk = k0 + kq;
      int nb = (gr < R) ? (kend - gk) * 4 : 0;
      nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
      const size_t row = (rowmap != nullptr && nb) ? (size_t)rowmap[gr] : (size_t)gr;
      cp_async16(dst + r * kLdK + kq, nb ? X + row * ld + gk : X, nb);
    } else {      // 16 k x 16 chunks of 4 rows
      const int k = tid >> 4, rq = (tid & 15) * 4;
      const int gk = k0 + k, gr = r0 + rq;
      int nb = (gk < kend) ? (R - gr) * 4 : 0;
      nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
      const size_t row = (rowmap != nullptr && nb) ? (size_t)rowmap[gk] : (size_t)gk;
      cp_async16(dst + k * kLdM + rq, nb ? X + row * ld + gr : X, nb);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = tid + q * 256;
      if (KMAJ) {
        const int k = i & (BK - 1), r = i >> 4;
        const bool ok = (r0 + r < R) && (k0 + k < kend);
        const size_t row = (rowmap != nullptr && ok) ? (size_t)rowmap[r0 + r] : (size_t)(r0 + r);
        cp_async4(dst + r * kLdK + k, ok ? X + row * ld + k0 + k : X, ok ? 4 : 0);
      } else {
        const int r = i & (BM - 1), k = i >> 6;
        const bool ok = (r0 + r < R) && (k0 + k < kend);
        const size_t row = (rowmap != nullptr && ok) ? (size_t)rowmap[k0 + k] : (size_t)(k0 + k);
        cp_async4(dst + k * kLdM + r, ok ? X + row * ld + r0 + r : X, ok ? 4 : 0);
      }
    }
  }
}

// 8 warps = 2 (m) x 4 (n); a warp owns a 32 x 16 block = 2 x 2 m16n8 accumulator fragments
template <bool TA, bool TB, int VA, int VB>
__global__ void __launch_bounds__(256) gemm_mma_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                       const float* __restrict__ B, int ldb,
                                                       const float* __restrict__ bias, int act,
                                                       float* __restrict__ C, int ldc, int accumulate,
                                                       int k_per_split, const int64_t* __restrict__ a_rows) {
  static_assert(BM == 64 && BN == 64 && BK == 16, "warp layout below is written for 64x64x16");
  __shared__ __align__(16) float As[kStages][kOpFloats];
  __shared__ __align__(16) float Bs[kStages][kOpFloats];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const bool atomic_out = gridDim.z > 1;
  const int ntile = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

  auto issue = [&](int tile) {
    if (tile < ntile) {
      const int k0 = kbeg + tile * BK, st = tile % kStages;
      stage_operand<!TA, VA>(As[st], A, lda, m0, M, k0, kend, tid, a_rows);   // A[m][k]: k contiguous unless transposed
      stage_operand<TB, VB>(Bs[st], B, ldb, n0, N, k0, kend, tid);    // B[k][n]: k contiguous only if transposed
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int s2 = 0; s2 < kStages - 1; ++s2) issue(s2);

  for (int it = 0; it < ntile; ++it) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
    __syncthreads();               // tile `it` has landed for everyone; everyone is done with tile it-1
    issue(it + kStages - 1);       // refills the stage tile it-1 occupied
    const float* as = As[it % kStages];
    const float* bs = Bs[it % kStages];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ka = j * 8 + t, kb = ka + 4;
      uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = wm * 32 + mi * 16 + g;
        float x[4];
        if (!TA) { x[0] = as[r * kLdK + ka]; x[1] = as[(r + 8) * kLdK + ka]; x[2] = as[r * kLdK + kb]; x[3] = as[(r + 8) * kLdK + kb]; }
        else     { x[0] = as[ka * kLdM + r]; x[1] = as[ka * kLdM + r + 8];   x[2] = as[kb * kLdM + r]; x[3] = as[kb * kLdM + r + 8]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) split_tf32(x[q], ah[mi][q], al[mi][q]);
      }
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const int c = wn * 16 + ni * 8 + g;
        float x[2];
        if (TB) { x[0] = bs[c * kLdK + ka]; x[1] = bs[c * kLdK + kb]; }
        else    { x[0] = bs[ka * kLdM + c]; x[1] = bs[kb * kLdM + c]; }
#pragma unroll
        for (int q = 0; q < 2; ++q) split_tf32(x[q], bh[ni][q], bl[ni][q]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          mma_tf32(acc[mi][ni], al[mi], bh[ni]);   // small terms first
          mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
          mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
        }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // accumulator fragment: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gm = m0 + wm * 32 + mi * 16 + g + (q >> 1) * 8;
        const int gn = n0 + wn * 16 + ni * 8 + 2 * t + (q & 1);
        if (gm >= M || gn >= N) continue;
        float v = acc[mi][ni][q];
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

static int gemm_impl(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t trans_a,
                     const float* B, int32_t ldb, int32_t trans_b, const float* bias, int32_t act,
                     float* C, int32_t ldc, int32_t accumulate, void* stream, const int64_t* a_rows) {
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
  static const bool use_ffma = [] { const char* e = getenv("FVTA_GEMM_FFMA"); return e && e[0] == '1'; }();
#define FVTA_GEMM_ARGS M, N, K, A, lda, B, ldb, kbias, kact, C, ldc, accumulate, kps
  if (use_ffma && a_rows == nullptr) {
    if (!trans_a && !trans_b) gemm_kernel<false, false><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS);
    else if (!trans_a && trans_b) gemm_kernel<false, true><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS);
    else if (trans_a && !trans_b) gemm_kernel<true, false><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS);
    else gemm_kernel<true, true><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS);
  } else {
    // 16-byte copies need a 16-byte aligned base, a leading dimension that keeps rows aligned and
    // split-K boundaries on multiples of 4 (kps is a multiple of BK)
    const bool va = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0);
    const bool vb = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0);
#define FVTA_GEMM_V(TA_, TB_)                                                                          \
    do {                                                                                               \
      if (va && vb) gemm_mma_kernel<TA_, TB_, 4, 4><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS, a_rows);     \
      else if (va) gemm_mma_kernel<TA_, TB_, 4, 1><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS, a_rows);      \
      else if (vb) gemm_mma_kernel<TA_, TB_, 1, 4><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS, a_rows);      \
      else gemm_mma_kernel<TA_, TB_, 1, 1><<<grid, 256, 0, st>>>(FVTA_GEMM_ARGS, a_rows);              \
    } while (0)
    if (!trans_a && !trans_b) FVTA_GEMM_V(false, false);
    else if (!trans_a && trans_b) FVTA_GEMM_V(false, true);
    else if (trans_a && !trans_b) FVTA_GEMM_V(true, false);
    else FVTA_GEMM_V(true, true);
#undef FVTA_GEMM_V
  }
#undef FVTA_GEMM_ARGS
  int rc = check_launch("gemm_kernel");
  if (rc || !split_epilogue) return rc;
  gemm_epilogue<<<(M * N + 255) / 256, 256, 0, st>>>(M, N, C, ldc, bias, act);
  return check_launch("gemm_epilogue");
}

extern "C" int fvta_gemm(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t trans_a,
                         const float* B, int32_t ldb, int32_t trans_b, const float* bias, int32_t act,
                         float* C, int32_t ldc, int32_t accumulate, void* stream) {
  return gemm_impl(M, N, K, A, lda, trans_a, B, ldb, trans_b, bias, act, C, ldc, accumulate, stream, nullptr);
}

extern "C" int fvta_gemm_gather(int32_t M, int32_t N, int32_t K, const float* A_table, int32_t lda,
                                const int64_t* a_rows, int32_t trans_a, const float* B, int32_t ldb,
                                int32_t trans_b, const float* bias, int32_t act, float* C, int32_t ldc,
                                int32_t accumulate, void* stream) {
  if (!a_rows) return FVTA_E_BADARG;
  return gemm_impl(M, N, K, A_table, lda, trans_a, B, ldb, trans_b, bias, act, C, ldc, accumulate, stream, a_rows);
}
