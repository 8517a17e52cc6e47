// blend_bwd.cu — gradient of the alpha blending; work unit = one 512-entry SEGMENT of one 16x16 tile's list.
//
// Behavioural reference: renderCUDA (backward), diff-gaussian-rasterization/cuda_rasterizer/backward.cu:399-557.
// Differences in *how* (results agree to fp32 rounding / atomics order; gated at rel-L2 1e-3 vs the reference):
//   * segment-parallel replay.  The forward checkpoints every pixel's recurrence state (T, C) at every 512th
//     list position (blend_fwd.cu) and stores the final state; a CTA replays one segment back-to-front starting
//     from "transmittance before the segment end" and "colour blended behind it" = C_final - C_checkpoint.
//     The reference walks the whole tile list in one CTA; here the ~585 non-empty tiles of an avatar view
//     become ~3-4 k independent units that the block scheduler balances dynamically, and the never-blended
//     tail beyond the tile's last contributor (tile_last) is not even staged;
//   * same producer warp / mbarrier ring as the forward;
//   * PAIR-CENTRIC consumers (pipeline.cuh): a warp owns two pixel rows of the tile, but its lanes stand for (entry,
//     pixel) candidates — the pixels of the entry's ellipse on the rows' live columns — and then for surviving pairs,
//     not for pixels.  The first version mapped pixels to lanes and walked the candidates in a loop: ncu showed its
//     gradient block running with 5.8 of 32 lanes active and 1.5 G warp instructions per 8-view batch (2.17 ms).
//     Here the alpha evaluation runs on 32 candidates per instruction and the gradient arithmetic + reductions on 32
//     surviving pairs per instruction; the per-pixel recurrences (transmittance, colour behind) live in shared
//     memory and are the only serialised part (1.08 G warp instructions, 1.69 ms);
//   * the backward evaluates exp / reciprocal with the hardware approximations (ex2.approx, rcp.approx) and this file
//     is compiled with fast-math (build.py): the gradient tolerance is relative L2 1e-3, measured ~1e-6.  (The forward
//     stays IEEE: its alpha >= 1/255 and T < 1e-4 decisions define n_contrib, which is compared bit for bit.)
//   * gradients are committed per pair with 16-byte VECTOR reductions (red.global.add.v4.f32) into one packed
//     accumulator row grad_acc[g][12] = {dmean2D.xy, dconic.xx/xy/yy, dopacity, drgb, dz, -, -};
//   * extras: gradients of the depth / alpha images flow through the same recurrence.
#include <algorithm>
#include "common.cuh"
#include "pipeline.cuh"

namespace tgr {

constexpr int BL_STAGES = 4;  // shared-memory ring depth = batches of a work unit: every batch of a unit has its own stage
static_assert(BL_STAGES * BL_BATCH >= SEG, "a work unit must fit the ring: pairs carried over batches refer to their stage");

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// per consumer warp
struct BwdWarpState {
  uint32_t hits[BL_BATCH];        // hit words of the current batch
  uint2 pairs[PB_CAP];            // {entry | pixel << 7, G bits}
  float2 ts[32];                  // per pixel: T (after the current entry), S = <dL/dpixel, blended behind> + T_final * tail
  float4 dpix[32];                // per pixel: dL/dcolour rgb, dL/ddepth
  int last[32];                   // per pixel: last contributor (1-based list position)
};

template <bool EXTRAS>
#ifndef TGR_BWD_MIN_CTAS
#define TGR_BWD_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(BL_THREADS, TGR_BWD_MIN_CTAS) blend_bwd_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.y];   // one launch serves every view of the batch
  const uint2* __restrict__ units = rv.units;
  const uint32_t* __restrict__ unit_count = rv.unit_count;
  const uint2* __restrict__ ranges = rv.ranges;
  const uint32_t* __restrict__ point_list = rv.point_list;
  const int W = rv.W, H = rv.H;
  const float* __restrict__ bg = rv.bg;
  const float4* __restrict__ xy_ext = rv.xy_ext;
  const float4* __restrict__ conic_opacity = rv.conic_opacity;
  const float4* __restrict__ rgb_depth = rv.rgb_depth;
  const float4* __restrict__ final_state = rv.final_state;
  const float* __restrict__ final_depth = rv.final_z;
  const uint32_t* __restrict__ n_contrib = rv.n_contrib;
  const uint32_t* __restrict__ tile_last = rv.tile_last;
  const uint32_t* __restrict__ seg_base = rv.seg_base;
  const float4* __restrict__ ckpt = rv.ckpt;
  const float* __restrict__ ckpt_z = rv.ckpt_z;
  const float* __restrict__ dL_dpix = rv.dL_dpix;
  const float* __restrict__ dL_ddepth = rv.dL_ddepth;
  const float* __restrict__ dL_dalpha_img = rv.dL_dalpha;
  float* __restrict__ grad_acc = rv.grad_acc;
  __shared__ uint32_t s_id[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_xy[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_co[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_cd[BL_STAGES][BL_BATCH + 1];
  extern __shared__ __align__(16) unsigned char s_dyn[];                 // 8 x BwdWarpState (static + dynamic > 48 KB)
  BwdWarpState* s_warp = reinterpret_cast<BwdWarpState*>(s_dyn);

  // ---- work unit = (tile, segment): list positions [seg*SEG, min((seg+1)*SEG, total)) of one tile -------
  if (blockIdx.x >= *unit_count) return;
  const uint2 unit = units[blockIdx.x];
  const uint32_t tile_id = unit.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tiles_x = (W + TILE - 1) / TILE;
  const uint32_t tile_bx = tile_id % tiles_x, tile_by = tile_id / tiles_x;
  const uint2 range = ranges[tile_id];
  const int total = (int)min(tile_last[tile_id], range.y - range.x);  // entries [0,total) can contribute
  const int seg_lo = (int)unit.y * SEG;
  const int seg_hi = min(seg_lo + SEG, total);                        // exclusive
  const int count = seg_hi - seg_lo;
  if (count <= 0) return;
  const int rounds = (count + BL_BATCH - 1) / BL_BATCH;

  __shared__ __align__(8) uint64_t s_full[BL_STAGES], s_empty[BL_STAGES];
  if (tid == 0) {
    for (int s = 0; s < BL_STAGES; ++s) {
      mbar_init(&s_full[s], 32);   // every producer lane's copies arrive (cp.async.mbarrier.arrive)
      mbar_init(&s_empty[s], 8);   // one arrival per consumer warp
    }
  }
  __syncthreads();

  if (warp == 8) {
    // ======================= PRODUCER: back-to-front gather of this segment =======================
    // segment-local list; batch entry e <-> list position count-1-e
    producer_loop<BL_STAGES, true, true>(point_list + range.x + seg_lo, count, rounds, xy_ext, conic_opacity, rgb_depth,
                                            s_xy, s_co, s_cd, s_id, s_full, s_empty, lane, [](int) { return false; });
    return;
  }

  // ========================= CONSUMERS =========================
  BwdWarpState& ws = s_warp[warp];
  const int bx0 = (int)(tile_bx * TILE);                    // pixel origin of this warp's 16x2 rows
  const int by0 = (int)(tile_by * TILE) + warp * 2;
  const float ddelx_dx = 0.5f * W;
  const float ddely_dy = 0.5f * H;
  int warp_last, my_last;
  {
    // lane l owns pixel (l & 15, l >> 4) of the two rows for the set-up of the per-pixel state
    const int lx = lane & 15, ly = lane >> 4;
    const uint32_t px = (uint32_t)(bx0 + lx), py = (uint32_t)(by0 + ly);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    const int pix_in_tile = (warp * 2 + ly) * TILE + lx;     // checkpoints are stored in tile raster order
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
    my_last = last_contributor;
    warp_last = __reduce_max_sync(0xffffffffu, last_contributor);

    // State of the forward recurrence at the END of this segment.  Behind-colour B = sum_{j >= seg_hi} c_j a_j T_j
    // = C_final - C(before seg_hi); T = transmittance before entry seg_hi.  If nothing contributes at or after
    // seg_hi (it is the tile's last needed segment) the end state is the final state.
    float T_final = 0.f, T = 0.f, B0 = 0.f, B1 = 0.f, B2 = 0.f, Bz = 0.f;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f, ddep = 0.f, dalp = 0.f;
    if (inside) {
      const float4 fs = final_state[pix_id];
      T_final = fs.x;
      T = fs.x;
      if (seg_hi < total) {
        const size_t slot = ((size_t)seg_base[tile_id] + (size_t)(seg_hi / SEG)) * TILE_PIX + pix_in_tile;
        const float4 ck = ckpt[slot];
        T = ck.x;
        B0 = fs.y - ck.y; B1 = fs.z - ck.z; B2 = fs.w - ck.w;
        if (EXTRAS) Bz = final_depth[pix_id] - ckpt_z[slot];
      }
      const size_t HW = (size_t)H * W;
      dp0 = dL_dpix[0 * HW + pix_id];
      dp1 = dL_dpix[1 * HW + pix_id];
      dp2 = dL_dpix[2 * HW + pix_id];
      if (EXTRAS) {
        if (dL_ddepth) ddep = dL_ddepth[pix_id];
        if (dL_dalpha_img) dalp = dL_dalpha_img[pix_id];
      }
    }
    // d(out)/dalpha_i through the final transmittance: colour gets +T_final*bg, the alpha image -T_final
    float tail = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    if (EXTRAS) tail -= dalp;
    // Only the inner product of "blended behind" with the pixel's upstream gradient enters dL/dalpha, so the
    // recurrence carries that scalar instead of the four channels; the final-transmittance term has the same
    // 1/(1-alpha) factor and is folded into its start value.
    float S = B0 * dp0 + B1 * dp1 + B2 * dp2 + T_final * tail;
    if (EXTRAS) S = fmaf(Bz, ddep, S);
    ws.ts[lane] = make_float2(T, S);
    ws.dpix[lane] = make_float4(dp0, dp1, dp2, ddep);
    ws.last[lane] = last_contributor;
  }
  __syncwarp();

  const uint32_t lt = (1u << lane) - 1u;
  int npairs = 0;      // pairs waiting in the ring (warp-uniform)
  int phead = 0;       // ring position of the oldest one

  // Commits the oldest min(32, npairs) pairs of the ring: ordered update of the per-pixel recurrences, then the
  // gradient arithmetic and the reductions at full width.
  auto commit = [&]() {
    const int n = min(npairs, 32);
    const bool act = lane < n;
    uint32_t pw = 0;
    float G = 0.f;
    if (act) {
      const uint2 pr = ws.pairs[(phead + lane) & (PB_CAP - 1)];
      pw = pr.x;
      G = __uint_as_float(pr.y);
    }
    const int j = (int)(pw & 127u);
    const int pix = (int)((pw >> 7) & 31u);
    const int stage = (int)((pw >> 12) & 3u);   // the ring stage the pair's records sit in (pairs are carried over batches)
    const float4 con_o = s_co[stage][j];
    const float4 cd = s_cd[stage][j];
    const float alpha = min(0.99f, con_o.w * G);
    const float rinv = rcp_approx(1.f - alpha);
    const float4 dp = ws.dpix[pix];
    float cdp = cd.x * dp.x + cd.y * dp.y + cd.z * dp.z;   // <colour (and depth) of the entry, upstream gradient of the pixel>
    if (EXTRAS) cdp = fmaf(cd.w, dp.w, cdp);
    // pairs of this round that fall on the same pixel go one after the other, in list order (= lane order)
    const uint32_t peers = __match_any_sync(0xffffffffu, act ? pix : 32 + lane);
    const int rank = __popc(peers & lt);
    const int maxrank = __reduce_max_sync(0xffffffffu, act ? rank : 0);
    // The recurrence of a pixel over its pairs of this round is a chain of affine maps of (T, S):
    //   T' = m T,  S' = S + v T   with m = 1/(1-alpha), v = <c, dL/dpixel> alpha m;
    // chains compose as (m1,v1) then (m2,v2) = (m1 m2, v1 + m1 v2), so every pair gets the composite of its pixel's
    // earlier pairs by pointer jumping over the peer lanes: ceil(log2(longest chain)) shuffle steps instead of one
    // shared-memory turn per chain link.
    int prev = act ? 31 - __clz(peers & lt) : -1;     // the pixel's previous pair of this round (lane), -1: none
    float M = act ? rinv : 1.f;
    float Vv = act ? cdp * alpha * rinv : 0.f;
    const float2 ts0 = ws.ts[pix];
    for (int span = maxrank; span > 0; span >>= 1) {
      const float Mp = __shfl_sync(0xffffffffu, M, prev);
      const float Vp = __shfl_sync(0xffffffffu, Vv, prev);
      const int pp = __shfl_sync(0xffffffffu, prev, prev);
      if (prev >= 0) {
        Vv = fmaf(Mp, Vv, Vp);
        M *= Mp;
        prev = pp;
      }
    }
    const float Ti = ts0.x * M;                        // transmittance in front of the entry
    const float Sa = fmaf(Vv, ts0.x, ts0.y);           // S with the entry included ...
    const float Sb = fmaf(-cdp * alpha, Ti, Sa);       // ... and behind it
    __syncwarp();
    if (act && (peers >> lane) == 1u) ws.ts[pix] = make_float2(Ti, Sa);   // the pixel's last pair of this round
    __syncwarp();
    if (act) {
      const float4 g = s_xy[stage][j];
      const float dx = g.x - (float)(bx0 + (pix & 15)), dy = g.y - (float)(by0 + (pix >> 4));
      const float w = alpha * Ti;
      // dC/dalpha_i = c_i*Ti - B/(1-alpha_i)   (same quantity as backward.cu:515-525's (c - accum_rec)*T), contracted
      // with the upstream gradient; the -T_final/(1-alpha_i) * tail term (backward.cu:527-533) rides in S
      const float dL_dalpha = cdp * Ti - Sb * rinv;
      const float gz = EXTRAS ? w * dp.w : 0.f;
      // the 0.99 cap is straight-through in the reference (backward.cu:494-497 recomputes alpha with the min)
      const float dL_dG = con_o.w * dL_dalpha;
      const float gdx = G * dx;
      const float gdy = G * dy;
      const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
      const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
      float* row = grad_acc + (size_t)s_id[stage][j] * GRAD_ACC;
      red_add_v4(row + 0, dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, -0.5f * gdx * dx * dL_dG,
                 -0.5f * gdx * dy * dL_dG);
      red_add_v4(row + 4, -0.5f * gdy * dy * dL_dG, G * dL_dalpha, w * dp.x, w * dp.y);
      if (EXTRAS) red_add_v2(row + 8, w * dp.z, gz);
      else atomicAdd(row + 8, w * dp.z);
    }
    phead = (phead + n) & (PB_CAP - 1);
    npairs -= n;
  };

  for (int b = 0; b < rounds; ++b) {
    const int stage = b % BL_STAGES;
    const int batch_first_pos = seg_hi - 1 - b * BL_BATCH;  // list position of batch entry 0 (descending)
    mbar_wait(&s_full[stage], (b / BL_STAGES) & 1);
    if (batch_first_pos - (BL_BATCH - 1) < warp_last) {  // else: whole batch lies behind this block's last contributor
      // entries at list positions >= warp_last lie behind this block's last contributor: batch entry e sits at
      // position batch_first_pos - e, so only e > batch_first_pos - warp_last can matter
      // live columns: pixels whose last contributor lies at or before the batch's nearest entry
      const int batch_near_pos = batch_first_pos - min(BL_BATCH, count - b * BL_BATCH) + 1;
      const RowWindow rw = row_window(__ballot_sync(0xffffffffu, batch_near_pos < my_last), bx0);
      const int nhits = classify_hits(max(0, batch_first_pos - warp_last + 1), min(BL_BATCH, count - b * BL_BATCH),
                                      s_xy[stage], s_co[stage], ws.hits, bx0, by0, rw, lane);
      for (int h0 = 0; h0 < nhits; h0 += 32) {
        // lane i <- hit h0 + i; rectangles laid end to end: start = exclusive prefix of the pixel counts
        const uint32_t hw = (h0 + lane < nhits) ? ws.hits[h0 + lane] : 0u;
        const int n = (h0 + lane < nhits) ? hit_pixels(hw) : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int start = incl - n;
        const int cand_total = __shfl_sync(0xffffffffu, incl, 31);
        for (int base = 0; base < cand_total; base += 32) {
          int j, pix;
          bool ok = expand_candidate(hw, n, start, cand_total, base, lane, j, pix);
          pix &= 31;
          const int pos = batch_first_pos - j;   // 0-based list position of the entry
          ok = ok && pos < ws.last[pix];
          const float4 g = s_xy[stage][j];
          const float4 con_o = s_co[stage][j];
          const float dx = g.x - (float)(bx0 + (pix & 15)), dy = g.y - (float)(by0 + (pix >> 4));
          const float power = -0.5f * (con_o.x * dx * dx + con_o.z * dy * dy) - con_o.y * dx * dy;
          const float G = ex2_approx(power * 1.4426950408889634f);
          const float alpha = min(0.99f, con_o.w * G);
          ok = ok && (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
          const uint32_t bal = __ballot_sync(0xffffffffu, ok);
          if (ok) ws.pairs[(phead + npairs + __popc(bal & lt)) & (PB_CAP - 1)] = make_uint2((uint32_t)j | ((uint32_t)pix << 7) | ((uint32_t)stage << 12), __float_as_uint(G));
          npairs += __popc(bal);
          __syncwarp();
          if (npairs >= 32) commit();
        }
      }
    }
    // Pairs that do not fill a commit round are carried into the next batch.  Their records stay valid: a work unit is
    // at most SEG = BL_STAGES * BL_BATCH entries long, so no ring stage is ever refilled within a unit (the producer
    // never has to wait for empty[], which is why nothing arrives on it here).
  }
  while (npairs > 0) commit();
}

int launch_blend_bwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s) {
  if (int rc = launch_unit_build(rb, s)) return rc;
  // one CTA per work unit; the grid is sized for the largest capacity, surplus CTAs exit on the device-side count
  uint32_t ucap = 0;
  for (int v = 0; v < rb.V; ++v) ucap = std::max(ucap, rb.v[v].units_cap);
  if (ucap == 0) return 0;
  const dim3 grid(ucap, rb.V, 1);
  constexpr int dyn = 8 * (int)sizeof(BwdWarpState);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(blend_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaFuncSetAttribute(blend_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    attr_set = true;
  }
  if (extras) blend_bwd_kernel<true><<<grid, BL_THREADS, dyn, s>>>(rb);
  else blend_bwd_kernel<false><<<grid, BL_THREADS, dyn, s>>>(rb);
  count_launch();
  return check_launch("blend_bwd", debug, s);
}

}  // namespace tgr
