// Hardware probe (development tool, not part of the library): pins down the tcgen05 behaviours the
// attention kernels rely on, on the actual B200:
//   1. kind::tf32 with A in TMEM (lane = row, one 32-bit column per k) and B in shared memory
//      (K-major, 128B swizzle): does the tensor core TRUNCATE or ROUND the 13 low mantissa bits?
//   2. kind::f16 with A in TMEM: two fp16 per column, element 2c in the low half?
//   3. both kinds accumulating into the same fp32 accumulator columns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe.bin tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// M = 128; fmt 0 = f16, 2 = tf32
__device__ __forceinline__ uint32_t idesc(int fmt, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}

constexpr int KD = 32;   // k extent (one 128-byte swizzle row of tf32, half a row of f16)
constexpr int NB = 32;   // N

// A32 [128][KD] fp32, B32 [NB][KD] fp32, A16 [128][KD] half, B16 [NB][KD] half
// out [3][128][NB]: tf32 product, f16 product, tf32 + f16 accumulated together
__global__ void __launch_bounds__(128, 1) probe(const float* A32, const float* B32, const __half* A16,
                                                const __half* B16, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* b32_s = smem;            // [NB][128 B]
  uint8_t* b16_s = smem + 4096;     // [NB][128 B] (first 64 B of each row used)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B operands, swizzled: 16-byte unit u of row n lands at unit (u ^ (n & 7))
  for (int i = tid; i < NB * 8; i += 128) {
    const int n = i >> 3, u = i & 7;
    *reinterpret_cast<uint4*>(b32_s + n * 128 + ((u ^ (n & 7)) << 4)) = *reinterpret_cast<const uint4*>(B32 + n * KD + u * 4);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (u < 4) v = *reinterpret_cast<const uint4*>(B16 + n * KD + u * 8);
    *reinterpret_cast<uint4*>(b16_s + n * 128 + ((u ^ (n & 7)) << 4)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  // columns: [0,32) acc tf32, [32,64) acc f16, [64,96) acc both, [96,128) A tf32, [128,144) A f16
  {
    uint32_t r[32];
    for (int k = 0; k < KD; ++k) r[k] = __float_as_uint(A32[tid * KD + k]);
    for (int c = 0; c < KD; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + 96 + c),
                   "r"(r[c]), "r"(r[c + 1]), "r"(r[c + 2]), "r"(r[c + 3]), "r"(r[c + 4]), "r"(r[c + 5]), "r"(r[c + 6]), "r"(r[c + 7]) : "memory");
    uint32_t p[16];
    for (int c = 0; c < 16; ++c) {
      const __half2 v = __halves2half2(A16[tid * KD + 2 * c], A16[tid * KD + 2 * c + 1]);   // low half = element 2c
      p[c] = *reinterpret_cast<const uint32_t*>(&v);
    }
    for (int c = 0; c < 16; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + 128 + c),
                   "r"(p[c]), "r"(p[c + 1]), "r"(p[c + 2]), "r"(p[c + 3]), "r"(p[c + 4]), "r"(p[c + 5]), "r"(p[c + 6]), "r"(p[c + 7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t i32 = idesc(2, NB), i16 = idesc(0, NB);
    for (int ks = 0; ks < KD / 8; ++ks) {   // tf32: 8 per k-step = 32 bytes, 8 TMEM columns
      const uint64_t bd = desc_sw128(smem_u32(b32_s) + ks * 32);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tb + 0), "r"(tb + 96 + ks * 8), "l"(bd), "r"(i32), "r"((uint32_t)(ks > 0)) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tb + 64), "r"(tb + 96 + ks * 8), "l"(bd), "r"(i32), "r"((uint32_t)(ks > 0)) : "memory");
    }
    for (int ks = 0; ks < KD / 16; ++ks) {  // f16: 16 per k-step = 32 bytes, 8 TMEM columns
      const uint64_t bd = desc_sw128(smem_u32(b16_s) + ks * 32);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tb + 32), "r"(tb + 128 + ks * 8), "l"(bd), "r"(i16), "r"((uint32_t)(ks > 0)) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tb + 64), "r"(tb + 128 + ks * 8), "l"(bd), "r"(i16), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int g = 0; g < 3; ++g) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(lane_base + g * 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < NB; ++j) out[(g * 128 + tid) * NB + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tb) : "memory");
}

// ---- timing: cycles per tcgen05.mma for the operand sources / shapes the attention kernels consider ----
// mode 0: A in TMEM (TS), 1: A in shared memory (SS).  fmt 0 = f16, 2 = tf32.  reps MMAs back to back into
// one accumulator, then commit; cycles from first issue to the commit's arrival.
__global__ void __launch_bounds__(128, 1) mma_rate(int mode, int fmt, int N, int reps, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_s = smem;              // 128 rows x 128 B
  uint8_t* b_s = smem + 16384;      // 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  {
    uint32_t r[8];
    for (int c = 0; c < 8; ++c) r[c] = 0x3c003c00u;
    const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 8)   // A region: columns 256..319
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + 256 + c),
                   "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t id = idesc(fmt, N);
    const uint32_t tb0 = tb;
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int ks = i & 3;
      const uint64_t bd = desc_sw128(smem_u32(b_s) + ks * 32);
      const uint32_t tb = tb0 + (uint32_t)((i % nacc) * 64);     // independent accumulators, round robin
      if (mode == 0) {
        if (fmt == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tb), "r"(tb0 + 256 + ks * 8), "l"(bd), "r"(id), "r"((uint32_t)(i >= nacc)) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tb), "r"(tb0 + 256 + ks * 8), "l"(bd), "r"(id), "r"((uint32_t)(i >= nacc)) : "memory");
      } else {
        const uint64_t ad = desc_sw128(smem_u32(a_s) + ks * 32);
        if (fmt == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tb), "l"(ad), "l"(bd), "r"(id), "r"((uint32_t)(i > 0)) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tb), "l"(ad), "l"(bd), "r"(id), "r"((uint32_t)(i > 0)) : "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

static void run_rates() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int reps = 256;
  printf("cycles per tcgen05.mma (M=128, %d back to back): issue loop / until commit arrives\n", reps);
  for (int mode = 0; mode < 2; ++mode)
    for (int fmt = 0; fmt <= 2; fmt += 2)
      for (int N = 32; N <= 64; N *= 2)
      for (int nacc = 1; nacc <= 4; nacc *= 2) {
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          mma_rate<<<1, 128, 49152>>>(mode, fmt, N, reps, nacc, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mma_rate failed: %s\n", cudaGetErrorString(e)); return; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("  A in %s  %-4s N=%-3d accumulators=%d  issue %.1f  complete %.1f\n", mode == 0 ? "TMEM" : "smem", fmt == 2 ? "tf32" : "f16", N, nacc,
               (double)h[0] / reps, (double)h[1] / reps);
      }
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rn_tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x0FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A(128 * KD), B(NB * KD);
  std::vector<__half> A16(128 * KD), B16(NB * KD);
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (size_t i = 0; i < A.size(); ++i) A16[i] = __float2half_rn(A[i] * 0.001f);
  for (size_t i = 0; i < B.size(); ++i) B16[i] = __float2half_rn(B[i]);
  float *dA, *dB, *dO; __half *dA16, *dB16;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 3 * 128 * NB * 4);
  cudaMalloc(&dA16, A16.size() * 2); cudaMalloc(&dB16, B16.size() * 2);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dA16, A16.data(), A16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB16, B16.data(), B16.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe<<<1, 128, 8192>>>(dA, dB, dA16, dB16, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  run_rates();
  std::vector<float> O(3 * 128 * NB);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  typedef float (*Cv)(float);
  Cv cvs[3] = {trunc_tf32, rn_tf32, rna_tf32};
  const char* nm[3] = {"trunc", "rne", "rna"};
  for (int ca = 0; ca < 3; ++ca)
    for (int cb = 0; cb < 3; ++cb) {
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NB; ++n) {
          double s = 0;
          for (int k = 0; k < KD; ++k) s += (double)cvs[ca](A[m * KD + k]) * (double)cvs[cb](B[n * KD + k]);
          worst = fmax(worst, fabs(s - O[m * NB + n]));
        }
      printf("tf32  A:%-5s B:%-5s  max|diff| = %.3e\n", nm[ca], nm[cb], worst);
    }
  {
    double worst = 0, worst_exact = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < NB; ++n) {
        double s = 0, sx = 0;
        for (int k = 0; k < KD; ++k) { s += (double)__half2float(A16[m * KD + k]) * (double)__half2float(B16[n * KD + k]); sx += (double)A[m * KD + k] * B[n * KD + k]; }
        worst = fmax(worst, fabs(s - O[(128 + m) * NB + n]));
        worst_exact = fmax(worst_exact, fabs(sx - O[m * NB + n]));
      }
    printf("f16   max|diff| vs fp16-exact products = %.3e   (tf32 vs fp32-exact inputs: %.3e)\n", worst, worst_exact);
  }
  {
    double worst = 0;
    for (int i = 0; i < 128 * NB; ++i) worst = fmax(worst, fabs((double)O[i] + (double)O[128 * NB + i] - (double)O[2 * 128 * NB + i]));
    printf("mixed accumulate: max|tf32 + f16 - both| = %.3e\n", worst);
  }
  return 0;
}
