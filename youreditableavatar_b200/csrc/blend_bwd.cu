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
//   * PERSISTENT CTAs with the forward's producer warp / mbarrier ring: the units of all views of the batch are drawn
//     from one device-side ticket counter; the producer warp publishes a unit's description, gathers its list into
//     the ring and runs ahead into the NEXT unit while the consumers finish the current one (round 1 and most of
//     round 2 launched one CTA per unit: every CTA then starts with a chain of four dependent global loads before
//     its first batch lands — 18 % of the warp-stall samples);
//   * PAIR-CENTRIC consumers (pipeline.cuh): a warp owns two pixel rows of the tile, but its lanes stand for (entry,
//     pixel) candidates — the pixels of the entry's ellipse on the rows' live columns — and then for surviving pairs,
//     not for pixels.  The first version mapped pixels to lanes and walked the candidates in a loop: ncu showed its
//     gradient block running with 5.8 of 32 lanes active and 1.5 G warp instructions per 8-view batch (2.17 ms).
//     Here the alpha evaluation runs on 32 candidates per instruction and the gradient arithmetic + reductions on 32
//     surviving pairs per instruction;
//   * the per-pixel recurrence carries TWO scalars, the transmittance T and S = <dL/dpixel, blended behind> (+ the
//     final-transmittance term): dL/dalpha only needs that inner product, not the four blended channels.  A pair is
//     then an affine map of (T, S) and the pairs of a commit round that fall on the same pixel compose their maps by
//     pointer jumping over the peer lanes (log2 steps of three shuffles) instead of taking turns through shared
//     memory (0.92 G warp instructions, 1.40 ms; with the four-channel state and turns: 1.08 G, 1.69 ms);
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

// Shared-memory ring.  A work unit is at most UNIT_BATCHES batches long and takes that many CONSECUTIVE stages; the
// ring is deeper than a unit so that the producer warp runs ahead into the next unit while the consumers finish the
// current one.  Stages are released when their unit is finished (pairs carried over batches refer to their stage).
constexpr int UNIT_BATCHES = SEG / BL_BATCH;
#ifndef TGR_BWD_STAGES
#define TGR_BWD_STAGES 5   // measured on C3 x8: 5 stages 1.406 ms, 6 (3 CTAs/SM fit) 1.48, 8 (3 CTAs/SM) 1.406; CTA per unit 1.431
#endif
constexpr int BL_STAGES = TGR_BWD_STAGES;
static_assert(UNIT_BATCHES * BL_BATCH == SEG && BL_STAGES > UNIT_BATCHES && BL_STAGES <= 8, "ring vs work unit");

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
  uint2 pairs[PB_CAP];            // {entry | pixel << 7 | ring stage << 12, G bits}
  float2 ts[32];                  // per pixel: T (after the current entry), S = <dL/dpixel, blended behind> + T_final * tail
  float4 dpix[32];                // per pixel: dL/dcolour rgb, dL/ddepth
  int last[32];                   // per pixel: last contributor (1-based list position)
};

// What the producer warp tells the consumers about a work unit.
struct BwdUnit {
  int view;            // index into the batch; < 0: no more work
  uint32_t tile_id;
  int seg_hi;          // list positions [seg_hi - count, seg_hi) of the tile, replayed back to front
  int count;
  uint32_t ckpt_slot;  // checkpoint holding the recurrence state at seg_hi; 0xffffffff: the final state is the end state
  int g0;              // running batch number of the unit's first batch (stage = g % BL_STAGES, parity from g / BL_STAGES)
};

#ifndef TGR_BWD_MIN_CTAS
#define TGR_BWD_MIN_CTAS 4
#endif
template <bool EXTRAS>
__global__ void __launch_bounds__(BL_THREADS, TGR_BWD_MIN_CTAS) blend_bwd_kernel(const __grid_constant__ RenderBatch rb) {
  __shared__ uint32_t s_id[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_xy[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_co[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_cd[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(8) uint64_t s_full[BL_STAGES], s_empty[BL_STAGES], s_ufull[2], s_uempty[2];
  __shared__ BwdUnit s_unit[2];
  __shared__ uint32_t s_prefix[MAX_BATCH + 1];   // units of the views before view v (one ticket space for the batch)
  extern __shared__ __align__(16) unsigned char s_dyn[];                 // 8 x BwdWarpState (static + dynamic > 48 KB)
  BwdWarpState* s_warp = reinterpret_cast<BwdWarpState*>(s_dyn);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < BL_STAGES; ++s) {
      mbar_init(&s_full[s], 32);   // every producer lane's copies arrive (cp.async.mbarrier.arrive)
      mbar_init(&s_empty[s], 8);   // one arrival per consumer warp, when the unit that used the stage is finished
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_ufull[s], 1);
      mbar_init(&s_uempty[s], 8);
    }
    uint32_t acc = 0;
    for (int v = 0; v < rb.V; ++v) {
      s_prefix[v] = acc;
      acc += *rb.v[v].unit_count;
    }
    s_prefix[rb.V] = acc;
  }
  __syncthreads();

  if (warp == 8) {
    // ======================= PRODUCER =======================
    // Persistent CTAs: work units (tile, segment) of ALL views of the batch are numbered through; CTA b starts with
    // unit b and draws further ones from one ticket counter (unit_count[1] of view 0, zeroed by unit_build_kernel).  Per unit: publish its description, then gather its
    // list back to front — batch entry e <-> list position seg_hi-1-e — into the next stages of the ring.
    uint32_t* ticket = rb.v[0].unit_count + 1;
    const uint32_t n_units = s_prefix[rb.V];
    int g0 = 0;
    for (int k = 0;; ++k) {
      // The first unit of a CTA is its block index, later ones come from the ticket counter: a batch with fewer
      // units than CTAs (small scenes) then runs one unit per CTA, all in parallel — with tickets from the start the
      // CTAs that launch first ran ahead into a second unit before the last CTAs had drawn their first (C1: +15 %).
      uint32_t u = blockIdx.x;
      if (k > 0) {
        if (lane == 0) u = gridDim.x + atomicAdd(ticket, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
      }
      if (k >= 2) mbar_wait(&s_uempty[k & 1], ((k >> 1) - 1) & 1);
      if (u >= n_units) {
        if (lane == 0) {
          s_unit[k & 1].view = -1;
          mbar_arrive(&s_ufull[k & 1]);
        }
        break;
      }
      int v = 0;
      while (u >= s_prefix[v + 1]) ++v;
      const RenderView& rv = rb.v[v];
      const uint2 unit = rv.units[u - s_prefix[v]];
      const uint2 range = rv.ranges[unit.x];
      const int total = (int)min(rv.tile_last[unit.x], range.y - range.x);  // entries [0,total) can contribute
      const int seg_lo = (int)unit.y * SEG;
      const int seg_hi = min(seg_lo + SEG, total);                          // exclusive
      const int count = max(seg_hi - seg_lo, 0);
      const int rounds = (count + BL_BATCH - 1) / BL_BATCH;
      const uint32_t* __restrict__ list = rv.point_list + range.x + seg_lo;
      uint32_t ids[BL_CHUNKS];
      prod_load_ids(list, count, 0, true, lane, ids);
      if (lane == 0) {
        BwdUnit un;
        un.view = v;
        un.tile_id = unit.x;
        un.seg_hi = seg_hi;
        un.count = count;
        un.ckpt_slot = (seg_hi < total) ? rv.seg_base[unit.x] + (uint32_t)(seg_hi / SEG) : 0xffffffffu;
        un.g0 = g0;
        s_unit[k & 1] = un;
        mbar_arrive(&s_ufull[k & 1]);   // release: the record is visible to whoever observes the phase
      }
      for (int b = 0; b < rounds; ++b) {
        const int g = g0 + b;
        const int st = g % BL_STAGES;
        if (g >= BL_STAGES) mbar_wait(&s_empty[st], ((g / BL_STAGES) - 1) & 1);
        prod_issue(ids, list, count, b * BL_BATCH, true, rv.xy_ext, rv.conic_opacity, rv.rgb_depth, s_xy[st], s_co[st],
                   s_cd[st], s_id[st], lane);
        cp_async_mbar_arrive(&s_full[st]);
        prod_load_ids(list, count, (b + 1) * BL_BATCH, true, lane, ids);
      }
      g0 += rounds;
    }
    cp_async_wait<0>();  // do not leave with copies in flight
    return;
  }

  // ========================= CONSUMERS =========================
  BwdWarpState& ws = s_warp[warp];
  const uint32_t lt = (1u << lane) - 1u;
  int npairs = 0;      // pairs waiting in the ring (warp-uniform)
  int phead = 0;       // ring position of the oldest one

  for (int k = 0;; ++k) {
    mbar_wait(&s_ufull[k & 1], (k >> 1) & 1);
    const BwdUnit un = s_unit[k & 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_uempty[k & 1]);
    if (un.view < 0) break;
    const RenderView& rv = rb.v[un.view];
    const int W = rv.W, H = rv.H;
    float* __restrict__ grad_acc = rv.grad_acc;
    const uint32_t tiles_x = (W + TILE - 1) / TILE;
    const int bx0 = (int)((un.tile_id % tiles_x) * TILE);     // pixel origin of this warp's 16x2 rows
    const int by0 = (int)((un.tile_id / tiles_x) * TILE) + warp * 2;
    const float ddelx_dx = 0.5f * W;
    const float ddely_dy = 0.5f * H;
    const int seg_hi = un.seg_hi, count = un.count;
    const int rounds = (count + BL_BATCH - 1) / BL_BATCH;
    int warp_last, my_last;
    {
      // lane l owns pixel (l & 15, l >> 4) of the two rows for the set-up of the per-pixel state
      const int lx = lane & 15, ly = lane >> 4;
      const uint32_t px = (uint32_t)(bx0 + lx), py = (uint32_t)(by0 + ly);
      const bool inside = px < (uint32_t)W && py < (uint32_t)H;
      const uint32_t pix_id = (uint32_t)W * py + px;
      const int pix_in_tile = (warp * 2 + ly) * TILE + lx;     // checkpoints are stored in tile raster order
      const int last_contributor = inside ? (int)rv.n_contrib[pix_id] : 0;
      my_last = last_contributor;
      warp_last = __reduce_max_sync(0xffffffffu, last_contributor);

      // State of the forward recurrence at the END of this segment.  Behind-colour B = sum_{j >= seg_hi} c_j a_j T_j
      // = C_final - C(before seg_hi); T = transmittance before entry seg_hi.  If nothing contributes at or after
      // seg_hi (it is the tile's last needed segment) the end state is the final state.
      float T_final = 0.f, T = 0.f, B0 = 0.f, B1 = 0.f, B2 = 0.f, Bz = 0.f;
      float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f, ddep = 0.f, dalp = 0.f;
      if (inside) {
        const float4 fs = rv.final_state[pix_id];
        T_final = fs.x;
        T = fs.x;
        if (un.ckpt_slot != 0xffffffffu) {
          const size_t slot = (size_t)un.ckpt_slot * TILE_PIX + pix_in_tile;
          const float4 ck = rv.ckpt[slot];
          T = ck.x;
          B0 = fs.y - ck.y; B1 = fs.z - ck.z; B2 = fs.w - ck.w;
          if (EXTRAS) Bz = rv.final_z[pix_id] - rv.ckpt_z[slot];
        }
        const size_t HW = (size_t)H * W;
        dp0 = rv.dL_dpix[0 * HW + pix_id];
        dp1 = rv.dL_dpix[1 * HW + pix_id];
        dp2 = rv.dL_dpix[2 * HW + pix_id];
        if (EXTRAS) {
          if (rv.dL_ddepth) ddep = rv.dL_ddepth[pix_id];
          if (rv.dL_dalpha) dalp = rv.dL_dalpha[pix_id];
        }
      }
      // d(out)/dalpha_i through the final transmittance: colour gets +T_final*bg, the alpha image -T_final
      float tail = rv.bg[0] * dp0 + rv.bg[1] * dp1 + rv.bg[2] * dp2;
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

    // Commits the oldest min(32, npairs) pairs of the ring: update of the per-pixel recurrences, then the gradient
    // arithmetic and the reductions at full width.
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
      const int stage = (int)((pw >> 12) & 7u);   // the ring stage the pair's records sit in (pairs are carried over batches)
      const float4 con_o = s_co[stage][j];
      const float4 cd = s_cd[stage][j];
      const float alpha = min(0.99f, con_o.w * G);
      const float rinv = rcp_approx(1.f - alpha);
      const float4 dp = ws.dpix[pix];
      float cdp = cd.x * dp.x + cd.y * dp.y + cd.z * dp.z;   // <colour (and depth) of the entry, upstream gradient of the pixel>
      if (EXTRAS) cdp = fmaf(cd.w, dp.w, cdp);
      // pairs of this round that fall on the same pixel: in list order (= lane order)
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
      const int g = un.g0 + b;
      const int stage = g % BL_STAGES;
      const int batch_first_pos = seg_hi - 1 - b * BL_BATCH;  // list position of batch entry 0 (descending)
      mbar_wait(&s_full[stage], (g / BL_STAGES) & 1);
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
            const float4 gq = s_xy[stage][j];
            const float4 con_o = s_co[stage][j];
            const float dx = gq.x - (float)(bx0 + (pix & 15)), dy = gq.y - (float)(by0 + (pix >> 4));
            const float power = -0.5f * (con_o.x * dx * dx + con_o.z * dy * dy) - con_o.y * dx * dy;
            const float G = ex2_approx(power * 1.4426950408889634f);
            const float alpha = min(0.99f, con_o.w * G);
            ok = ok && (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
            const uint32_t bal = __ballot_sync(0xffffffffu, ok);
            if (ok)
              ws.pairs[(phead + npairs + __popc(bal & lt)) & (PB_CAP - 1)] =
                  make_uint2((uint32_t)j | ((uint32_t)pix << 7) | ((uint32_t)stage << 12), __float_as_uint(G));
            npairs += __popc(bal);
            __syncwarp();
            if (npairs >= 32) commit();
          }
        }
      }
      // Pairs that do not fill a commit round are carried into the next batch: their records stay in place until
      // the unit is finished.
    }
    while (npairs > 0) commit();
    __syncwarp();
    if (lane == 0)
      for (int b = 0; b < rounds; ++b) mbar_arrive(&s_empty[(un.g0 + b) % BL_STAGES]);
  }
}

int launch_blend_bwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s) {
  if (int rc = launch_unit_build(rb, s)) return rc;
  uint32_t ucap = 0;
  for (int v = 0; v < rb.V; ++v) ucap = std::max(ucap, rb.v[v].units_cap);
  if (ucap == 0) return 0;
  // persistent CTAs, all resident at once; they draw work units from a device-side ticket counter
  const dim3 grid(num_queues() * TGR_BWD_MIN_CTAS, 1, 1);
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
