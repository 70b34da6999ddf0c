// Forward LSTM recurrence of the reader on the 5th-generation tensor cores (north_star c).
//
// Replaces the per-step product h_{t-1} . W_h of BasicLSTMCell (model_v2.py:649-661; the eight
// bidirectional_dynamic_rnn call sites :694-823) for hidden sizes d <= 64.  A CTA owns 32 sequences of
// one direction for all their steps; the recurrent product runs TRANSPOSED so that the wide dimension
// is the MMA's M:
//
//   G^T [4d gates x 32 sequences] = W_h^T [4d x d] . h_{t-1}^T [d x 32]        per step
//
//   A (shared memory, static)  : W_h^T as two M = 128 tiles of fp16 (hi, lo) images, K-major, 128B swizzle
//                                tile 0: rows 0-63 gate i of unit r, rows 64-127 gate j of unit r - 64
//                                tile 1: rows 0-63 gate f,           rows 64-127 gate o
//   B (shared memory, per step): [h_hi ; h_lo], 64 rows (sequence) x 64 (unit) fp16, rewritten by the cell
//                                threads every step
//   D (tensor memory)          : acc[tile][:, 0:32]  = W_hi . h_hi^T  (+ the projected input x_t.W_x + b,
//                                stored there BEFORE the MMAs run: the accumulator is the prefetch buffer)
//                                acc[tile][:, 32:64] = W_hi . h_lo^T + W_lo . h_hi^T
//   fp16 hi/lo split of both operands: ~22 mantissa bits per product, fp32 accumulation -- the 1e-4 parity
//   bar through 25 recurrent steps holds (single-pass fp16/bf16 would not).
//
// 16 tcgen05.mma per step (2 tiles x 4 k-steps x {N = 64, N = 32}), issued by one elected lane; 16 cell
// warps read the gate pre-activations back with tcgen05.ld (lane = gate row, columns = sequences): thread
// (unit e, lower half) holds gates i and f, thread (unit e, upper half) gates j and o, which it hands over
// through shared memory; the lower thread keeps the cell state of its 8 sequences in registers.
#include "common.cuh"
#include "bilstm.cuh"
#include <cuda_fp16.h>

namespace fvta {

constexpr int kTcSeq = 32;                 // sequences per CTA (N of the MMAs)
constexpr int kTcLstmThreads = 544;        // warps 0-15 cell update, warp 16 MMA issue + TMEM allocation
constexpr int kLtMmaWarp = 16;
constexpr int kLtSPT = 8;                  // sequences per cell thread (4 column groups of 8)
constexpr int kTcLstmTmemCols = 256;       // acc[2 buffers][2 tiles][64 columns]

__device__ __forceinline__ uint32_t lt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool lt_mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(lt_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void lt_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!lt_mbar_try(bar, parity))
    if (++spins > (1u << 24)) __trap();     // a pipeline bug must fail the launch, not hang the device
}
__device__ __forceinline__ void lt_mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!lt_mbar_try(bar, parity)) {
    __nanosleep(32);
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ bool lt_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void lt_umma_ss_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint64_t lt_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, fp16 operands K-major, fp32 accumulate, M = 128
__device__ __forceinline__ uint32_t lt_idesc(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ void lt_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void lt_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void lt_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void lt_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}

// shared memory: A images 4 x 16 KB | B image 8 KB | exchange 2 x 64 x 32 floats | per-sequence metadata
constexpr int kLtSmemA = 4 * 16384, kLtSmemB = 8192, kLtSmemX = 2 * 64 * kTcSeq * 4;
constexpr int kLtSmemBytes = kLtSmemA + kLtSmemB + kLtSmemX + 1024;

// grid (ceil(B / 32), 2 directions)
__global__ void __launch_bounds__(kTcLstmThreads, 1)
lstm_rec_fwd_tc(FvtaLstmDesc L, SeqAddr A, const int32_t* __restrict__ lens, const float* __restrict__ xp_fw,
                const float* __restrict__ xp_bw, const float* __restrict__ wh_fw, const float* __restrict__ wh_bw,
                float* __restrict__ out, float* __restrict__ last, float* __restrict__ reserve) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_img = smem;                                   // [tile][hi/lo][128 rows][128 B]
  uint8_t* b_img = smem + kLtSmemA;                        // [64 rows][128 B]
  float* xch = reinterpret_cast<float*>(smem + kLtSmemA + kLtSmemB);   // [2][64 units][32 sequences]
  int* len_s = reinterpret_cast<int*>(smem + kLtSmemA + kLtSmemB + kLtSmemX);   // [32]
  int* tok_s = len_s + kTcSeq;                                                  // [32]
  long long* orow_s = reinterpret_cast<long long*>(tok_s + kTcSeq);             // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(orow_s + kTcSeq);
  uint64_t* b_ready = bars;            // cell -> MMA: h image and the next accumulator's initial value are in place
  uint64_t* acc_full = bars + 1;       // MMA -> cell
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);

  const int d = L.d, G = 4 * d;
  const int dir = blockIdx.y;
  const float* wh = dir == 0 ? wh_fw : wh_bw;      // [d][4d]: rows In.. of the LSTM kernel
  const float* xp = dir == 0 ? xp_fw : xp_bw;      // [ntok][4d] projected inputs incl. bias
  const int b0 = blockIdx.x * kTcSeq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- prologue: operand images, metadata, barriers, tensor memory ----
  // A element (tile, part, row r, k): gate = 2*tile + (r >= 64), unit e = r & 63, value W_h[k][gate*d + e]
  for (int i = threadIdx.x; i < 2 * 128 * 8; i += blockDim.x) {
    const int u = i & 7, r = (i >> 3) & 127, tile = i >> 10;
    const int gate = 2 * tile + (r >> 6), e = r & 63;
    __half hi[8], lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = u * 8 + q;
      const float w = (e < d && k < d) ? wh[(size_t)k * G + gate * d + e] : 0.f;
      hi[q] = __float2half_rn(w);
      lo[q] = __float2half_rn(w - __half2float(hi[q]));
    }
    const int off = r * 128 + ((u ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(a_img + (tile * 2 + 0) * 16384 + off) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(a_img + (tile * 2 + 1) * 16384 + off) = *reinterpret_cast<uint4*>(lo);
  }
  for (int i = threadIdx.x; i < kLtSmemB / 16; i += blockDim.x) reinterpret_cast<uint4*>(b_img)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < kTcSeq) {
    const int b = b0 + threadIdx.x;
    len_s[threadIdx.x] = (b < L.B) ? min(lens[b], L.J) : 0;
    tok_s[threadIdx.x] = (b < L.B) ? A.tok0(b, L.J) : 0;
    orow_s[threadIdx.x] = (b < L.B) ? A.out0(b, L.J) : -1;
  }
  if (threadIdx.x == 0) {
    lt_mbar_init(b_ready, 512);
    lt_mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLtMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(lt_smem_u32(tmem_ptr)),
                 "r"(kTcLstmTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;
  int steps = 0;
#pragma unroll 8
  for (int q = 0; q < kTcSeq; ++q) steps = max(steps, len_s[q]);

  if (warp == kLtMmaWarp) {
    // ===== MMA issuer: the whole warp loops, one elected lane issues =====
    const uint32_t id64 = lt_idesc(64), id32 = lt_idesc(32);
    const uint32_t a_u = lt_smem_u32(a_img), b_u = lt_smem_u32(b_img);
    for (int t = 0; t < steps; ++t) {
      lt_mbar_wait(b_ready, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lt_elect_one()) {
        const uint32_t acc = tmem_base + (uint32_t)((t & 1) * 128);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bd = lt_desc_sw128(b_u + ks * 32);
            // W_hi . [h_hi ; h_lo]^T -> columns 0-63;   W_lo . h_hi^T -> columns 32-63
            lt_umma_ss_f16(acc + tile * 64, lt_desc_sw128(a_u + (tile * 2 + 0) * 16384 + ks * 32), bd, id64);
            lt_umma_ss_f16(acc + tile * 64 + 32, lt_desc_sw128(a_u + (tile * 2 + 1) * 16384 + ks * 32), bd, id32);
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(lt_smem_u32(acc_full)) : "memory");
      }
      __syncwarp();
    }
  } else {
    // ===== cell warps: lane row r = 32 * (warp & 3) + lane, sequences 8 * (warp >> 2) .. + 7 =====
    const int r = 32 * (warp & 3) + lane;
    const int sh = warp >> 2;                    // column group: sequences 8 sh .. 8 sh + 7
    const bool primary = r < 64;                 // gates i (tile 0) and f (tile 1); the upper half has j and o
    const int e = r & 63;
    const bool unit_ok = e < d;
    const int g0 = (primary ? 0 : 1) * d + e, g1 = (primary ? 2 : 3) * d + e;    // columns of xp / gate blocks
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    float c[kLtSPT], hcur[kLtSPT];
#pragma unroll
    for (int s = 0; s < kLtSPT; ++s) { c[s] = 0.f; hcur[s] = 0.f; }

    // projected inputs of step `t` for this thread's two gate rows -> the first 32 columns of accumulator
    // buffer (t & 1); the other 32 columns start at zero
    auto load_xp = [&](int t, float (&x0)[kLtSPT], float (&x1)[kLtSPT]) {
#pragma unroll
      for (int s = 0; s < kLtSPT; ++s) {
        const int sq = kLtSPT * sh + s, len = len_s[sq];
        float v0 = 0.f, v1 = 0.f;
        if (unit_ok && t < len) {
          const float* row = xp + ((size_t)tok_s[sq] + (dir == 0 ? t : len - 1 - t)) * G;
          v0 = row[g0];
          v1 = row[g1];
        }
        x0[s] = v0; x1[s] = v1;
      }
    };
    auto store_xp = [&](int t, const float (&x0)[kLtSPT], const float (&x1)[kLtSPT]) {
      const uint32_t acc = lane_base + (uint32_t)((t & 1) * 128);
      uint32_t z[kLtSPT], v[kLtSPT];
#pragma unroll
      for (int s = 0; s < kLtSPT; ++s) z[s] = 0u;
#pragma unroll
      for (int s = 0; s < kLtSPT; ++s) v[s] = __float_as_uint(x0[s]);
      lt_st8(acc + kLtSPT * sh, v);
      lt_st8(acc + 32 + kLtSPT * sh, z);
#pragma unroll
      for (int s = 0; s < kLtSPT; ++s) v[s] = __float_as_uint(x1[s]);
      lt_st8(acc + 64 + kLtSPT * sh, v);
      lt_st8(acc + 96 + kLtSPT * sh, z);
    };
    {
      float x0[kLtSPT], x1[kLtSPT];
      load_xp(0, x0, x1);
      store_xp(0, x0, x1);
      if (steps > 1) { load_xp(1, x0, x1); store_xp(1, x0, x1); }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      lt_mbar_arrive(b_ready);
    }
    // exchange buffer, unit index fastest (consecutive lanes -> consecutive banks): xch[which][sequence][unit]
    float* xj = xch + (size_t)(kLtSPT * sh) * 64 + e;
    float* xo = xj + kTcSeq * 64;
    for (int t = 0; t < steps; ++t) {
      lt_mbar_wait(acc_full, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = lane_base + (uint32_t)((t & 1) * 128);
      float v0[kLtSPT], v1[kLtSPT];
      {
        uint32_t p0[kLtSPT], p1[kLtSPT], q0[kLtSPT], q1[kLtSPT];
        lt_ld8(acc + kLtSPT * sh, p0);
        lt_ld8(acc + 32 + kLtSPT * sh, p1);
        lt_ld8(acc + 64 + kLtSPT * sh, q0);
        lt_ld8(acc + 96 + kLtSPT * sh, q1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int s = 0; s < kLtSPT; ++s) {
          v0[s] = __uint_as_float(p0[s]) + __uint_as_float(p1[s]);
          v1[s] = __uint_as_float(q0[s]) + __uint_as_float(q1[s]);
        }
      }
      // the projected inputs of step t + 2 travel while this step's cell update runs
      float nx0[kLtSPT], nx1[kLtSPT];
      const bool more = t + 2 < steps;
      if (more) load_xp(t + 2, nx0, nx1);
      if (!primary) {
#pragma unroll
        for (int s = 0; s < kLtSPT; ++s) {
          const int sq = kLtSPT * sh + s, len = len_s[sq];
          const float jg = tanh_fast(v0[s]), og = sigmoid_fast(v1[s]);
          xj[s * 64] = jg;
          xo[s * 64] = og;
          if (unit_ok && t < len) {
            float* rs = reserve + ((size_t)dir * A.ntok + tok_s[sq] + (dir == 0 ? t : len - 1 - t)) * kResW * d + e;
            rs[d] = jg; rs[3 * d] = og;
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + sh) : "memory");     // the four warps of this column group
      if (primary) {
#pragma unroll
        for (int s = 0; s < kLtSPT; ++s) {
          const int sq = kLtSPT * sh + s, len = len_s[sq];
          if (unit_ok && t < len) {
            const int tt = dir == 0 ? t : len - 1 - t;
            const float ig = sigmoid_fast(v0[s]);
            const float fg = sigmoid_fast(v1[s] + 1.0f);            // forget_bias = 1.0 [TF]
            const float jg = xj[s * 64], og = xo[s * 64];
            const float hp = hcur[s];
            c[s] = fmaf(c[s], fg, ig * jg);
            const float h = tanh_fast(c[s]) * og;
            hcur[s] = h;
            if (orow_s[sq] >= 0) out[((size_t)orow_s[sq] + tt) * 2 * d + dir * d + e] = h;
            float* rs = reserve + ((size_t)dir * A.ntok + tok_s[sq] + tt) * kResW * d + e;
            rs[0] = ig; rs[2 * d] = fg; rs[4 * d] = c[s]; rs[5 * d] = hp;
            // B image: row sq holds h_hi, row 32 + sq holds h_lo; element k = e
            const __half hh = __float2half_rn(h);
            const __half hl = __float2half_rn(h - __half2float(hh));
            const int off = sq * 128 + (((e >> 3) ^ (sq & 7)) << 4) + (e & 7) * 2;
            *reinterpret_cast<__half*>(b_img + off) = hh;
            *reinterpret_cast<__half*>(b_img + 32 * 128 + off) = hl;
          }
        }
      }
      if (more) store_xp(t + 2, nx0, nx1);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + sh) : "memory");     // xch is rewritten next step
      if (t + 1 < steps) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // h image: generic-proxy writes, UMMA reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        lt_mbar_arrive(b_ready);
      }
    }
    if (primary && unit_ok) {
#pragma unroll
      for (int s = 0; s < kLtSPT; ++s) {
        const int b = b0 + kLtSPT * sh + s;
        if (b < L.B) last[(size_t)b * 2 * d + dir * d + e] = hcur[s];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kLtMmaWarp)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcLstmTmemCols) : "memory");
}

// 0 = launched, 1 = shape not covered (caller uses the CUDA-core recurrence), < 0 error
int lstm_fwd_tc_try(const FvtaLstmDesc& L, const SeqAddr& A, const int32_t* lens, const float* xp_fw, const float* xp_bw,
                    const float* wh_fw, const float* wh_bw, float* out, float* last, float* reserve, cudaStream_t st) {
  if (L.d > 64 || L.d < 1) return 1;
  cudaError_t e = cudaFuncSetAttribute(lstm_rec_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kLtSmemBytes);
  if (e != cudaSuccess) { set_last_error("lstm_rec_fwd_tc smem attr", e); return FVTA_E_LAUNCH; }
  dim3 grid((L.B + kTcSeq - 1) / kTcSeq, 2);
  lstm_rec_fwd_tc<<<grid, kTcLstmThreads, kLtSmemBytes, st>>>(L, A, lens, xp_fw, xp_bw, wh_fw, wh_bw, out, last, reserve);
  return check_launch("lstm_rec_fwd_tc");
}

}  // namespace fvta
