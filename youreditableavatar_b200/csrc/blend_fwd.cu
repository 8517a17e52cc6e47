// blend_fwd.cu — front-to-back alpha blending, one CTA per 16x16 tile.
//
// Behavioural reference: renderCUDA, diff-gaussian-rasterization/cuda_rasterizer/forward.cu:261-374.
// The per-pair arithmetic (power, alpha = min(0.99, o*exp(power)), the 1/255 and 1e-4 thresholds, colour
// accumulation order) is evaluated with the same fp32 expressions so the image is bit-identical.
// Extras (new): depth = sum z_i alpha_i T_i, alpha = 1 - T_final.
//
// How it differs from the reference kernel (which stages 256 entries, __syncthreads, all 256 pixels
// evaluate all 256 entries, __syncthreads, repeat):
//   * warp specialisation + mbarrier ring.  9 warps: warp 8 is the PRODUCER — it gathers the tile's list in
//     batches of 128 entries (point_list -> per-Gaussian 16-byte records) into a 4-stage shared-memory ring
//     and signals full[stage]; warps 0..7 are CONSUMERS, each owning the 8x4 pixel block at
//     (8*(w&1), 4*(w>>1)) and signalling empty[stage] when done with a stage.  There is no __syncthreads in
//     the loop, so a consumer whose block is cheap runs ahead by up to 4 batches instead of waiting for the
//     slowest warp after every batch (ncu on the barrier version: >50% of issue slots stalled on barrier).
//   * footprint culling.  The producer tests every Gaussian's conservative footprint (half-extents of the
//     alpha >= 1/255 ellipse, computed once per Gaussian by the preprocess kernel) against the eight warp
//     blocks; __ballot_sync turns that into one 32-bit mask per (consumer, chunk of 32 entries).  A consumer
//     only evaluates entries whose footprint reaches its block; everything it skips would have failed the
//     reference's alpha < 1/255 test, so results are unchanged.
//   * the tile's maximum contributing list position is written out for the backward (tile_last).
#include "common.cuh"
#include "pipeline.cuh"

namespace tgr {

#ifdef TGR_MEASURE_STAGING
// Experiment (profiles/r02_staging_wait.txt): how long do the consumer warps of the forward wait for staged data?  That
// is the most any faster staging — a tile-ordered packed record stream moved by cp.async.bulk / TMA instead of 16-byte
// LDGSTS gathers — could win.  [0] cycles consumers spent blocked on full[], [1] all consumer cycles, [2] cycles the
// producer spent blocked on empty[] (back-pressure: data was ready earlier than needed), [3] all producer cycles.
__device__ unsigned long long g_staging_cycles[4];
#endif

// Shared-memory ring depth = how far the consumer warps of a tile may drift apart (a warp whose block is sparse runs
// ahead of the one with the dense block).  Measured on C3 x8: 4 stages 1.258 ms, 6 1.213, 7 1.208 (47.7 KB static, the
// most that fits without dynamic shared memory; still 4 CTAs/SM).
#ifndef TGR_FWD_STAGES
#define TGR_FWD_STAGES 7
#endif
constexpr int BL_STAGES = TGR_FWD_STAGES;



template <bool EXTRAS>
__global__ void __launch_bounds__(BL_THREADS, 4) blend_fwd_kernel(const __grid_constant__ RenderBatch rb, uint32_t num_queues) {
  constexpr int FG = 4;   // candidates evaluated together by a consumer lane group
  __shared__ __align__(16) float4 s_xy[BL_STAGES][BL_BATCH + 1];   // x, y, hx, hy   (+1: the PAD_ENTRY dummy)
  __shared__ __align__(16) float4 s_co[BL_STAGES][BL_BATCH + 1];   // conic xx, xy, yy, opacity
  __shared__ __align__(16) float4 s_cd[BL_STAGES][BL_BATCH + 1];   // r, g, b, depth
  __shared__ __align__(4) uint8_t s_list[8][SUB_GROUPS * LIST_BYTES];  // per consumer warp and lane group: candidates of the current batch
  __shared__ __align__(8) uint64_t s_full[BL_STAGES], s_empty[BL_STAGES];
  __shared__ uint32_t s_stop[BL_STAGES];
  __shared__ uint32_t s_done_warps;
  __shared__ uint32_t s_last[8];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Work item = one tile of one view.  Global rank r of the batch's queue -> view r % V, rank r / V in that view's
  // heaviest-first tile order, so the heaviest tiles of every view start first and one launch balances all views.
  __shared__ uint32_t s_rank;
  if (warp == 8) {
    // On the side (the producer warp idles here, DRAM is 11 % busy in this kernel): clear the packed 2-D gradient rows
    // the backward will reduce into — grid slice blockIdx.x / V of view blockIdx.x % V.  Replaces a 48 MB memset node
    // per view in front of the blend backward.
    const uint32_t V = (uint32_t)rb.V;
    const RenderView& zv = rb.v[blockIdx.x % V];
    const uint32_t n4 = (uint32_t)(((size_t)zv.P * GRAD_ACC) / 4);          // GRAD_ACC is a multiple of 4
    const uint32_t slices = gridDim.x / V;
    const uint32_t per = (n4 + slices - 1) / slices;
    const uint32_t lo = (blockIdx.x / V) * per, hi = min(lo + per, n4);
    float4* g4 = reinterpret_cast<float4*>(zv.grad_acc);
    for (uint32_t i = lo + lane; i < hi; i += 32) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == 0) {
    const uint32_t V = (uint32_t)rb.V;
    uint32_t r;
    while (true) {
      r = fetch_tile_rank(rb.queue_counters, V * rb.T_max, num_queues, lane);
      if (r == NO_TILE || r / V < rb.v[r % V].T) break;   // ranks beyond a smaller view's tile count are skipped
    }
    if (lane == 0) s_rank = r;
  }
  __syncthreads();
  if (s_rank == NO_TILE) return;
  const RenderView& rv = rb.v[s_rank % (uint32_t)rb.V];
  const int W = rv.W, H = rv.H;
  const uint2* __restrict__ ranges = rv.ranges;
  const uint32_t* __restrict__ point_list = rv.point_list;
  const float4* __restrict__ xy_ext = rv.xy_ext;
  const float4* __restrict__ conic_opacity = rv.conic_opacity;
  const float4* __restrict__ rgb_depth = rv.rgb_depth;
  const float* __restrict__ bg = rv.bg;
  float* __restrict__ final_T = rv.final_T;
  uint32_t* __restrict__ n_contrib = rv.n_contrib;
  uint32_t* __restrict__ tile_last = rv.tile_last;
  float* __restrict__ out_color = rv.out_color;
  float* __restrict__ out_depth = rv.out_depth;
  float* __restrict__ out_alpha = rv.out_alpha;
  const uint32_t* __restrict__ seg_base = rv.seg_base;
  float4* __restrict__ ckpt = rv.ckpt;
  float* __restrict__ ckpt_z = rv.ckpt_z;
  float4* __restrict__ final_state = rv.final_state;
  float* __restrict__ final_z = rv.final_z;
  const uint32_t tiles_x = (W + TILE - 1) / TILE;
  const uint32_t tile_id = rv.order_fwd[s_rank / (uint32_t)rb.V];
  const uint32_t tile_bx = tile_id % tiles_x, tile_by = tile_id / tiles_x;
  const uint2 range = ranges[tile_id];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + BL_BATCH - 1) / BL_BATCH;

  if (tid == 0) {
    for (int s = 0; s < BL_STAGES; ++s) {
      mbar_init(&s_full[s], 32);   // every producer lane's copies arrive (cp.async.mbarrier.arrive)
      mbar_init(&s_empty[s], 8);   // one arrival per consumer warp
      s_stop[s] = 0;
    }
    s_done_warps = 0;
  }
  if (tid < BL_STAGES) init_pad_record(s_xy[tid], s_co[tid], s_cd[tid]);
  __syncthreads();

#ifdef TGR_MEASURE_STAGING
  const long long t_begin = clock64();
  long long t_wait = 0;
#endif
  if (warp == 8) {
    // ======================= PRODUCER: a pure data mover =======================
    auto stop = [&](int stage) {
      if (*(volatile uint32_t*)&s_done_warps != 8u) return false;  // some pixel of the tile is still alive
      if (lane == 0) s_stop[stage] = 1;
      return true;
    };
    producer_loop<BL_STAGES, false, false>(point_list + range.x, total, rounds, xy_ext, conic_opacity, rgb_depth, s_xy,
                                             s_co, s_cd, nullptr, s_full, s_empty, lane, stop);
#ifdef TGR_MEASURE_STAGING
    if (lane == 0) atomicAdd(&g_staging_cycles[3], (unsigned long long)(clock64() - t_begin));
#endif
    return;
  }

  // ========================= CONSUMERS =========================
  int lx, ly, group;
  lane_pixel(warp, lane, lx, ly, group);
  const uint32_t px = tile_bx * TILE + lx;
  const uint32_t py = tile_by * TILE + ly;
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float2 pixf = {(float)px, (float)py};
  const float bx0 = (float)(tile_bx * TILE + (warp & 1) * 8);   // origin of this warp's 8x4 pixel block
  const float by0 = (float)(tile_by * TILE + (warp >> 1) * 4);

  bool done = !inside;
  bool warp_done = false;
  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float Dz = 0.f;

  const uint32_t ckpt_base = seg_base[tile_id];
  const int pix_in_tile = ly * TILE + lx;   // checkpoints are stored in tile raster order
  for (int b = 0; b < rounds; ++b) {
    const int stage = b % BL_STAGES;
    // checkpoint of the recurrence state at every SEG-th list position: lets the backward replay the
    // tile's list in independent segments (blend_bwd.cu)
    if (b > 0 && (b % (SEG / BL_BATCH)) == 0) {
      const size_t slot = ((size_t)ckpt_base + (size_t)(b / (SEG / BL_BATCH))) * TILE_PIX + pix_in_tile;
      ckpt[slot] = make_float4(T, C[0], C[1], C[2]);
      if (EXTRAS) ckpt_z[slot] = Dz;
    }
    if (!warp_done && __all_sync(0xffffffffu, done)) {
      warp_done = true;
      if (lane == 0) atomicAdd(&s_done_warps, 1u);
    }
#ifdef TGR_MEASURE_STAGING
    const long long t_w0 = clock64();
#endif
    mbar_wait(&s_full[stage], (b / BL_STAGES) & 1);
#ifdef TGR_MEASURE_STAGING
    t_wait += clock64() - t_w0;
#endif
    if (*(volatile uint32_t*)&s_stop[stage]) break;
    if (!warp_done) {
      const uint32_t base_pos = (uint32_t)(b * BL_BATCH);
      int longest;
      // sub-blocks all of whose pixels have terminated take no more candidates
      const uint32_t live = ~__ballot_sync(0xffffffffu, done);
      uint32_t group_live = 0;
#pragma unroll
      for (int g = 0; g < SUB_GROUPS; ++g) group_live |= ((live >> (g * SUB_LANES)) & ((1u << SUB_LANES) - 1u)) ? (1u << g) : 0u;
      const int ncand = cons_classify(0, min(BL_BATCH, total - b * BL_BATCH), s_xy[stage], s_list[warp], bx0, by0, lane, longest,
                                      group_live);
      const uint8_t* cand8 = s_list[warp] + group * LIST_BYTES;
      // Candidates are taken CAND_GROUP at a time (one 32-bit load = four batch-local indices): their loads,
      // power and exp are independent (ILP), only the transmittance update is a serial chain.  Profiling showed
      // the kernel bound by the single-warp latency of the heaviest tile, not by SM throughput.
#pragma unroll 1
      for (int i = 0; i < longest; i += FG) {
        const uint32_t packed = i < ncand ? *reinterpret_cast<const uint32_t*>(cand8 + i) : PAD_WORD;
        int j[FG];
        bool ok[FG];
        float alpha[FG], G[FG];
        float4 cd[FG];
#pragma unroll
        for (int k = 0; k < FG; ++k) {
          j[k] = (int)((packed >> (8 * k)) & 0xffu);
          const float4 g = s_xy[stage][j[k]];
          const float4 con_o = s_co[stage][j[k]];
          cd[k] = s_cd[stage][j[k]];
          const float2 d = {g.x - pixf.x, g.y - pixf.y};
          const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
          G[k] = expf(power);
          alpha[k] = min(0.99f, con_o.w * G[k]);
          ok[k] = (power <= 0.0f) && (alpha[k] >= 1.0f / 255.0f);
        }
#pragma unroll
        for (int k = 0; k < FG; ++k) {
          const float test_T = T * (1 - alpha[k]);
          const bool act = ok[k] && !done;
          const bool term = act && (test_T < 0.0001f);
          done = done || term;
          const bool blend = act && !term;
          if (blend) {
            C[0] += cd[k].x * alpha[k] * T;
            C[1] += cd[k].y * alpha[k] * T;
            C[2] += cd[k].z * alpha[k] * T;
            if (EXTRAS) Dz += cd[k].w * alpha[k] * T;
            T = test_T;
            last_contributor = base_pos + j[k] + 1;
          }
        }
        if (__all_sync(0xffffffffu, done)) break;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[stage]);
  }

  if (inside) {
    final_T[pix_id] = T;
    n_contrib[pix_id] = last_contributor;
    final_state[pix_id] = make_float4(T, C[0], C[1], C[2]);
    if (EXTRAS) final_z[pix_id] = Dz;
    const size_t HW = (size_t)H * W;
    out_color[0 * HW + pix_id] = C[0] + T * bg[0];
    out_color[1 * HW + pix_id] = C[1] + T * bg[1];
    out_color[2 * HW + pix_id] = C[2] + T * bg[2];
    if (EXTRAS) {
      out_depth[pix_id] = Dz;
      out_alpha[pix_id] = 1.0f - T;
    }
  }
#ifdef TGR_MEASURE_STAGING
  if (lane == 0) {
    atomicAdd(&g_staging_cycles[0], (unsigned long long)t_wait);
    atomicAdd(&g_staging_cycles[1], (unsigned long long)(clock64() - t_begin));
  }
#endif
  // tile-wide maximum of last_contributor: lets the backward start at the last useful list entry
  const uint32_t wl = __reduce_max_sync(0xffffffffu, inside ? last_contributor : 0u);
  if (lane == 0) s_last[warp] = wl;
  bar_sync_named(BAR_EPILOGUE, 256);  // the eight consumer warps only (the producer has left)
  if (tid == 0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = max(m, s_last[w]);
    tile_last[tile_id] = m;
  }
}

#ifdef TGR_MEASURE_STAGING
extern "C" int tgr_debug_staging_cycles(unsigned long long out[4], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_staging_cycles, sizeof(g_staging_cycles));
  cudaMemcpyFromSymbol(&out[2], g_producer_empty_wait, sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_staging_cycles, z, sizeof(z));
    cudaMemcpyToSymbol(g_producer_empty_wait, z, sizeof(unsigned long long));
  }
  return 0;
}
#endif

int launch_blend_fwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s) {
  if (int rc = launch_tile_order(rb, s)) return rc;
  // one CTA per (view, tile); the device-side queue hands out the work
  const dim3 grid(rb.T_max * (uint32_t)rb.V, 1, 1);
  if (extras) blend_fwd_kernel<true><<<grid, BL_THREADS, 0, s>>>(rb, num_queues());
  else blend_fwd_kernel<false><<<grid, BL_THREADS, 0, s>>>(rb, num_queues());
  count_launch();
  return check_launch("blend_fwd", debug, s);
}

}  // namespace tgr
