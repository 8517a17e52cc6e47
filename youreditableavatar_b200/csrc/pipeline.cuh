// pipeline.cuh — producer/consumer plumbing shared by the blend kernels: mbarrier helpers, the
// shared-memory ring geometry, footprint -> warp-block mask, and the producer warp's batch gather.
#pragma once
#include "common.cuh"

namespace tgr {

constexpr int BL_BATCH = 128;                // list entries per ring stage
constexpr int BL_CHUNKS = BL_BATCH / 32;     // 32-entry chunks (one ballot word each)
constexpr int BL_THREADS = 9 * 32;           // 8 consumer warps + 1 producer warp
static_assert(8 * BL_CHUNKS == 32, "produce_batch maps one (block, chunk) ballot word to each producer lane");

// ---- SM-affine work queues ------------------------------------------------------------------------------
// Tiles are sorted heaviest-first (tile_order_kernel).  All non-empty tiles of an avatar view fit in the first
// wave of CTAs, so WHICH SM a tile lands on decides the balance, and the hardware's CTA->SM placement is not
// a simple modulo (measured on B200: tools/cuda/smid_map.cu).  Instead of trusting blockIdx, every CTA reads
// %smid and pops the next rank of that SM's own queue; queue q holds the ranks dealt to it boustrophedon-wise
// (q, 2n-1-q, 2n+q, 4n-1-q, ...) so per-SM sums of a sorted list are even.  A CTA whose queue is drained
// steals from the others (warp-parallel scan of the counters), which also absorbs dynamic imbalance.
constexpr uint32_t NO_TILE = 0xffffffffu;

__device__ __forceinline__ uint32_t queue_rank(uint32_t q, uint32_t k, uint32_t nq) {
  return k * nq + ((k & 1u) ? (nq - 1u - q) : q);
}

// Called by warp 0 (all 32 lanes).  counters[nq] must be zero at kernel start.  Returns a rank < T or NO_TILE.
__device__ __forceinline__ uint32_t fetch_tile_rank(uint32_t* __restrict__ counters, uint32_t T, uint32_t nq, int lane) {
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const uint32_t q = smid % nq;
  uint32_t r = NO_TILE;
  if (lane == 0) {
    const uint32_t k = atomicAdd(&counters[q], 1u);
    r = queue_rank(q, k, nq);
    if (r >= T) r = NO_TILE;
  }
  r = __shfl_sync(0xffffffffu, r, 0);
  if (r != NO_TILE) return r;
  // own queue drained: steal.  Each lane inspects queues q+1+lane, q+1+lane+32, ...
  // (#CTAs == #ranks, so a CTA may only give up once every queue is observed drained)
  while (true) {
    uint32_t found = NO_TILE;
    for (uint32_t i = lane; i < nq && found == NO_TILE; i += 32) {
      const uint32_t qq = (q + 1u + i) % nq;
      const uint32_t k = *(volatile uint32_t*)&counters[qq];
      if (queue_rank(qq, k, nq) < T) found = qq;
    }
    const uint32_t have = __ballot_sync(0xffffffffu, found != NO_TILE);
    if (!have) return NO_TILE;  // every queue is drained
    const int src = __ffs(have) - 1;
    const uint32_t qq = __shfl_sync(0xffffffffu, found, src);
    if (lane == 0) {
      const uint32_t k = atomicAdd(&counters[qq], 1u);
      r = queue_rank(qq, k, nq);
      if (r >= T) r = NO_TILE;
    }
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r != NO_TILE) return r;
  }
}

// ---- mbarrier (shared::cta) ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- cp.async (LDGSTS) -----------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 8-bit mask of the warp blocks (bit = 2*row4 + col8) that the footprint [x-hx,x+hx]x[y-hy,y+hy] reaches.
// Pixel centres are integers (forward.cu:282), so pixel p is inside iff x-hx <= p <= x+hx.
__device__ __forceinline__ uint32_t block_mask(float x, float y, float hx, float hy, float tile_x0, float tile_y0) {
  if (!(hx >= 0.f)) return 0u;  // can never reach alpha >= 1/255 (or NaN extents)
  const float fx0 = ceilf(x - hx) - tile_x0, fx1 = floorf(x + hx) - tile_x0;
  const float fy0 = ceilf(y - hy) - tile_y0, fy1 = floorf(y + hy) - tile_y0;
  if (fx1 < 0.f || fy1 < 0.f || fx0 > 15.f || fy0 > 15.f || fx0 > fx1 || fy0 > fy1) return 0u;
  const int x0 = (int)fmaxf(fx0, 0.f), x1 = (int)fminf(fx1, 15.f);
  const int y0 = (int)fmaxf(fy0, 0.f), y1 = (int)fminf(fy1, 15.f);
  const uint32_t colm = ((x0 < 8) ? 1u : 0u) | ((x1 >= 8) ? 2u : 0u);
  const int r0 = y0 >> 2, r1 = y1 >> 2;
  uint32_t m = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r >= r0 && r <= r1) m |= colm << (2 * r);
  return m;
}

// ---- producer warp ------------------------------------------------------------------------------------
// Batch entry e maps to list position e (forward) or total-1-e (reverse, for the back-to-front replay).
// The producer is software-pipelined: ids of batch b+2 are prefetched into registers and the 16-byte
// cp.async gathers of batch b+1 are in flight while the footprints of batch b are classified.
__device__ __forceinline__ void prod_load_ids(const uint32_t* __restrict__ list, int total, int first, bool reverse,
                                              int lane, uint32_t (&ids)[BL_CHUNKS]) {
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    const int e = first + c * 32 + lane;
    ids[c] = 0xffffffffu;
    if (e < total) ids[c] = __ldg(list + (reverse ? (total - 1 - e) : e));
  }
}

// Records go global -> shared with cp.async (no register staging); one commit group per batch.
__device__ __forceinline__ void prod_issue(const uint32_t (&ids)[BL_CHUNKS], const float4* __restrict__ xy_ext,
                                           const float4* __restrict__ conic_opacity,
                                           const float4* __restrict__ rgb_depth, float4* s_xy, float4* s_co,
                                           float4* s_cd, uint32_t* s_id, int lane) {
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    if (ids[c] != 0xffffffffu) {
      const int j = c * 32 + lane;
      cp_async16(&s_xy[j], &xy_ext[ids[c]]);
      cp_async16(&s_co[j], &conic_opacity[ids[c]]);
      cp_async16(&s_cd[j], &rgb_depth[ids[c]]);
      if (s_id) s_id[j] = ids[c];
    }
  }
  cp_async_commit();
}

// Classifies the landed footprints of one batch against the eight warp blocks: s_ball[block][chunk].
__device__ __forceinline__ void prod_classify(int total, int first, const float4* s_xy, uint32_t (*s_ball)[BL_CHUNKS],
                                              float tile_x0, float tile_y0, int lane) {
  uint32_t keep = 0;
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    uint32_t mm = 0;
    if (first + c * 32 + lane < total) {
      const float4 g = s_xy[c * 32 + lane];
      mm = block_mask(g.x, g.y, g.z, g.w, tile_x0, tile_y0);
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (mm >> b) & 1u);
      if (lane == b * BL_CHUNKS + c) keep = bal;  // lane (b,c) holds the word for block b, chunk c
    }
  }
  if (lane < 8 * BL_CHUNKS) s_ball[lane / BL_CHUNKS][lane % BL_CHUNKS] = keep;
}

}  // namespace tgr
