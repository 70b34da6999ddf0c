// Data-parallel gradient all-reduce FUSED with the Adadelta update, over NVLink peer memory
// (north_star d: "gradients fused into one buffer and allreduced over NVLink").
//
// The reference is single-device (model_v2.py:14); its optimizer step is trainer.py:16-27.  With one
// process per GPU each rank holds a flat gradient buffer (ParamStore.grad_buf: every trainable + the
// "row was looked up" flags of the sparse embedding tables).  Instead of ncclAllReduce (45-95 us of
// mostly latency on 1.8 MB) followed by a separate optimizer launch, ONE kernel per rank
//   A. copies its gradient into its half of a peer-mapped staging buffer (two halves alternate by step,
//      so a rank that runs ahead never overwrites what a peer is still reading),
//   B. crosses a flag barrier over NVLink (each rank stores the step number into every peer's flag
//      array, then waits until its own array shows every rank),
//   C. reads every rank's staging buffer through the NVLink mappings, sums the W copies of each element
//      in rank order (the same order on every rank: bitwise identical weights everywhere, which NCCL's
//      ring order does not give), scales by 1/W and applies Adadelta to its local replica.
// One-shot all-reduce: each GPU reads (W-1) x 1.8 MB, ~20 us at 8 GPUs; no second barrier, no all-gather.
// The W loads of an element are in flight together (see WT below): what bounds the phase is NVLink
// bandwidth, not W round trips in a row.
#include "common.cuh"
#include "adadelta.cuh"

namespace fvta {

constexpr int kP2pMaxWorld = 16;
constexpr int kP2pMaxTail = 8192;   // floats of "touch" flags reduced into shared memory by every block

__device__ __forceinline__ float4 ld_peer4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// (relaxed: ONE system-scope fence orders the staging writes before all W flag stores; a release per
// store would wait for the previous store's NVLink round trip each time)
__device__ __forceinline__ void st_flag_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// state: [0] last completed epoch, [1] blocks that finished phase A, [2] blocks that finished phase C
// WT: the world size when it is 2, 4 or 8 (the peer loads of an element are then issued together -- a loop
// with a runtime trip count waits for each NVLink round trip, ~2.5 us, before asking for the next), else 0.
template <int WT>
__global__ void __launch_bounds__(512, 1) p2p_allreduce_adadelta_kernel(
    int world, int rank, long long n, long long n_tail, const float* __restrict__ grad_local,
    float* const* __restrict__ stage_ptrs, uint32_t* const* __restrict__ flag_ptrs, uint32_t* __restrict__ state,
    float* __restrict__ var, float* __restrict__ accum, float* __restrict__ accum_update,
    const float* __restrict__ mask, const int32_t* __restrict__ touch_index, float lr, float rho, float eps) {
  __shared__ unsigned touch_s[kP2pMaxTail];      // != 0: some rank looked the row up
  __shared__ const float* stage_s[kP2pMaxWorld];
  const long long total = n + n_tail;                    // both multiples of 4
  const uint32_t epoch = state[0] + 1;
  const long long half = (long long)(epoch & 1u) * total;
  if (threadIdx.x < world) stage_s[threadIdx.x] = stage_ptrs[threadIdx.x] + half;
  __syncthreads();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  // ---- A: local gradient -> own staging half ----
  {
    float4* dst = reinterpret_cast<float4*>(stage_ptrs[rank] + half);
    const float4* src = reinterpret_cast<const float4*>(grad_local);
    for (long long i = tid; i < (total >> 2); i += nthr) dst[i] = src[i];
  }
  for (long long j = threadIdx.x; j < n_tail; j += blockDim.x) touch_s[j] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    // fences are cumulative: after the barrier, this thread's fence orders the whole block's staging
    // writes before its counter increment; the last block's system-scope fence then orders every
    // block's writes before the flag stores
    __threadfence();
    const uint32_t prev = atomicAdd(&state[1], 1u);
    if (prev == gridDim.x - 1) {                         // every block of this rank has published its slice
      state[1] = 0;
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_flag_sys(flag_ptrs[r] + rank, epoch);
    }
  }
  // ---- B: every rank has published ----
  if (threadIdx.x < world) {
    const uint32_t* f = flag_ptrs[rank] + threadIdx.x;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_flag_sys(f) - epoch) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) __trap();           // 20 s: a peer died; fail the launch instead of hanging
    }
  }
  __syncthreads();
  // ---- C: reduce + update ----
  // the "row was looked up" flags: only > 0 matters, so the W x n_tail loads run in parallel over the
  // block's threads (one NVLink round trip, not W in a row) and meet in shared memory with an OR
  {
    const long long t4 = n_tail >> 2;
    for (long long idx = threadIdx.x; idx < t4 * world; idx += blockDim.x) {
      const int r = (int)(idx / t4);
      const long long j4 = idx - (long long)r * t4;
      const float4 x = ld_peer4(reinterpret_cast<const float4*>(stage_s[r] + n) + j4);
      if (x.x > 0.f) atomicOr(&touch_s[4 * j4], 1u);
      if (x.y > 0.f) atomicOr(&touch_s[4 * j4 + 1], 1u);
      if (x.z > 0.f) atomicOr(&touch_s[4 * j4 + 2], 1u);
      if (x.w > 0.f) atomicOr(&touch_s[4 * j4 + 3], 1u);
    }
  }
  __syncthreads();
  const float gs = 1.f / (float)world;
  for (long long i = tid; i < (n >> 2); i += nthr) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (WT > 0) {
      float4 x[WT > 0 ? WT : 1];
#pragma unroll
      for (int r = 0; r < WT; ++r) x[r] = ld_peer4(reinterpret_cast<const float4*>(stage_s[r]) + i);
#pragma unroll
      for (int r = 0; r < WT; ++r) { g.x += x[r].x; g.y += x[r].y; g.z += x[r].z; g.w += x[r].w; }   // rank order
    } else {
      for (int r = 0; r < world; ++r) {
        const float4 x = ld_peer4(reinterpret_cast<const float4*>(stage_s[r]) + i);
        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
      }
    }
    float4 v = reinterpret_cast<float4*>(var)[i];
    float4 a = reinterpret_cast<float4*>(accum)[i];
    float4 u = reinterpret_cast<float4*>(accum_update)[i];
    float4 m = mask ? reinterpret_cast<const float4*>(mask)[i] : make_float4(1.f, 1.f, 1.f, 1.f);
    if (touch_index != nullptr) {
      const int4 ti = reinterpret_cast<const int4*>(touch_index)[i];
      if (ti.x >= 0 && touch_s[ti.x] == 0u) m.x = 0.f;
      if (ti.y >= 0 && touch_s[ti.y] == 0u) m.y = 0.f;
      if (ti.z >= 0 && touch_s[ti.z] == 0u) m.z = 0.f;
      if (ti.w >= 0 && touch_s[ti.w] == 0u) m.w = 0.f;
    }
    adadelta_one(v.x, g.x, a.x, u.x, m.x, lr, rho, eps, gs);
    adadelta_one(v.y, g.y, a.y, u.y, m.y, lr, rho, eps, gs);
    adadelta_one(v.z, g.z, a.z, u.z, m.z, lr, rho, eps, gs);
    adadelta_one(v.w, g.w, a.w, u.w, m.w, lr, rho, eps, gs);
    reinterpret_cast<float4*>(var)[i] = v;
    reinterpret_cast<float4*>(accum)[i] = a;
    reinterpret_cast<float4*>(accum_update)[i] = u;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t prev = atomicAdd(&state[2], 1u);
    if (prev == gridDim.x - 1) { state[2] = 0; __threadfence(); state[0] = epoch; }
  }
}

}  // namespace fvta

extern "C" int fvta_p2p_allreduce_adadelta(int32_t world, int32_t rank, int64_t n, int64_t n_tail,
                                           const float* grad_local, float* const* stage_ptrs_dev,
                                           uint32_t* const* flag_ptrs_dev, uint32_t* state, float* var,
                                           float* accum, float* accum_update, const float* update_mask,
                                           const int32_t* touch_index, float lr, float rho, float eps,
                                           void* stream) {
  if (world < 1 || world > fvta::kP2pMaxWorld || rank < 0 || rank >= world) return FVTA_E_BADARG;
  if (n <= 0 || (n & 3) || n_tail < 0 || (n_tail & 3) || n_tail > fvta::kP2pMaxTail) return FVTA_E_SHAPE;
  if (!grad_local || !stage_ptrs_dev || !flag_ptrs_dev || !state || !var || !accum || !accum_update) return FVTA_E_BADARG;
  if ((touch_index != nullptr) != (n_tail > 0)) return FVTA_E_BADARG;
  // every block must be resident at once (they meet at the flag barrier): one block per SM at most
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = ((n >> 2) + 511) / 512;
  int blocks = (int)(want < 1 ? 1 : (want > sms ? sms : want));
  auto kern = world == 2   ? fvta::p2p_allreduce_adadelta_kernel<2>
              : world == 4 ? fvta::p2p_allreduce_adadelta_kernel<4>
              : world == 8 ? fvta::p2p_allreduce_adadelta_kernel<8>
                           : fvta::p2p_allreduce_adadelta_kernel<0>;
  kern<<<blocks, 512, 0, (cudaStream_t)stream>>>(world, rank, n, n_tail, grad_local, stage_ptrs_dev, flag_ptrs_dev, state,
                                                 var, accum, accum_update, update_mask, touch_index, lr, rho, eps);
  return fvta::check_launch("p2p_allreduce_adadelta_kernel");
}
