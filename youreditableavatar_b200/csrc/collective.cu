// collective.cu — gradient exchange over NVSwitch: a two-shot all-reduce with in-switch reduction (NVLS).
//
// The only exchange step of the data-parallel path (parallel.py) is the sum of one flat fp32 gradient buffer over
// the ranks.  When that buffer lives in symmetric memory with a multicast mapping, rank r owns slice r of it:
//   multimem.ld_reduce  — the switch reads the slice from EVERY GPU, adds, returns the sum to rank r;
//   multimem.st         — rank r stores the sum through the multicast address, the switch writes it to every GPU.
// Each GPU moves 1/N of the buffer per direction and phase, no intermediate copies, one kernel.  The caller
// brackets the launch with cross-rank barriers (all local gradients written before / all broadcasts landed after);
// those come from the symmetric-memory handle (torch.distributed._symmetric_memory, plumbing).
// The reference has no multi-GPU path at all (one view per step, tetgs_texture/refine.py:54).
#include <algorithm>
#include "common.cuh"

namespace tgr {

constexpr int MM_UNROLL = 8;   // independent in-switch reductions in flight per thread (the round trip is ~2-3 us)

__device__ __forceinline__ float4 mm_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* p, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256) multimem_allreduce_f32_kernel(float* __restrict__ mc, uint64_t n4, int rank, int world) {
  const uint64_t per = (n4 + world - 1) / world;
  const uint64_t begin = per * (uint64_t)rank;
  const uint64_t end = begin + per < n4 ? begin + per : n4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (MM_UNROLL - 1) * stride < end; i += MM_UNROLL * stride) {
    float4 v[MM_UNROLL];
#pragma unroll
    for (int k = 0; k < MM_UNROLL; ++k) v[k] = mm_ld_reduce(mc + 4 * (i + k * stride));
#pragma unroll
    for (int k = 0; k < MM_UNROLL; ++k) mm_st(mc + 4 * (i + k * stride), v[k]);
  }
  for (; i < end; i += stride) mm_st(mc + 4 * i, mm_ld_reduce(mc + 4 * i));
}

}  // namespace tgr

extern "C" int tgr_multimem_allreduce_f32_capped(void* multicast_ptr, uint64_t n_floats, int32_t rank, int32_t world,
                                                 int32_t max_ctas, void* stream);
extern "C" int tgr_multimem_allreduce_f32(void* multicast_ptr, uint64_t n_floats, int32_t rank, int32_t world, void* stream) {
  return tgr_multimem_allreduce_f32_capped(multicast_ptr, n_floats, rank, world, 0, stream);
}

extern "C" int tgr_multimem_allreduce_f32_capped(void* multicast_ptr, uint64_t n_floats, int32_t rank, int32_t world,
                                                 int32_t max_ctas, void* stream) {
  using namespace tgr;
  if (!multicast_ptr || world <= 0 || rank < 0 || rank >= world) { set_error("multimem_allreduce: bad arguments"); return 1; }
  if ((n_floats & 3) != 0 || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) != 0) {
    set_error("multimem_allreduce: buffer must be 16-byte aligned and a multiple of 4 floats");
    return 1;
  }
  if (n_floats == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint64_t n4 = n_floats / 4;
  const uint64_t per = (n4 + world - 1) / world;
  // many small threads beat few unrolled ones here (N = 8, 236 MB: 0.56 ms with SMs x 8 CTAs of 256 threads,
  // 0.64 ms with an eighth of them and 8 reductions in flight per thread; NCCL: 0.64 ms)
  // max_ctas > 0 caps the grid: a slice exchanged WHILE the backward still computes the next range of Gaussians must
  // not take the SMs away from it (the in-switch round trip is latency, not SM work: MM_UNROLL requests per thread)
  const uint64_t cap = max_ctas > 0 ? (uint64_t)max_ctas : (uint64_t)NUM_SM * 8;
  const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((per + 255) / 256, cap));
  multimem_allreduce_f32_kernel<<<blocks, 256, 0, s>>>(static_cast<float*>(multicast_ptr), n4, rank, world);
  count_launch();
  return check_launch("multimem_allreduce", false, s);
}
