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
//   * same producer/consumer ring and warp-block footprint culling as the forward: a warp only replays the
//     Gaussians that can reach alpha >= 1/255 inside its 8x4 pixel block, at or before its last contributor;
//   * two candidates are evaluated together (loads / power / exp / reciprocal are independent; only the
//     transmittance and behind-colour updates are serial);
//   * gradients are committed per contributing pixel with 16-byte VECTOR reductions (red.global.add.v4.f32)
//     into one packed accumulator row grad_acc[g][12] = {dmean2D.xy, dconic.xx/xy/yy, dopacity, drgb, dz, -, -}:
//     3 instructions per (pixel, Gaussian) where the reference issues 9 scalar atomics; warp-shuffle
//     pre-reductions were measured slower on B200 (see the comment at the reduction);
//   * extras: gradients of the depth / alpha images flow through the same recurrence.
#include <algorithm>
#include "common.cuh"
#include "pipeline.cuh"

namespace tgr {

constexpr int BL_STAGES = 4;  // shared-memory ring depth (how far consumers may drift apart)

constexpr int BG = 2;  // candidates evaluated together by a consumer warp

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool EXTRAS>
__global__ void __launch_bounds__(BL_THREADS) blend_bwd_kernel(const __grid_constant__ RenderBatch rb) {
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
  __shared__ __align__(16) float4 s_xy[BL_STAGES][BL_BATCH + 1];   // +1: the PAD_ENTRY dummy record
  __shared__ __align__(16) float4 s_co[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(16) float4 s_cd[BL_STAGES][BL_BATCH + 1];
  __shared__ __align__(4) uint8_t s_list[8][SUB_GROUPS * LIST_BYTES];  // per consumer warp and lane group: candidates of the current batch

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
  if (tid < BL_STAGES) {
    init_pad_record(s_xy[tid], s_co[tid], s_cd[tid]);
    s_id[tid][PAD_ENTRY] = 0;
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
  int lx, ly, group;
  lane_pixel(warp, lane, lx, ly, group);
  const uint32_t px = tile_bx * TILE + lx;
  const uint32_t py = tile_by * TILE + ly;
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float2 pixf = {(float)px, (float)py};
  const int pix_in_tile = warp * 32 + lane;
  const float bx0 = (float)(tile_bx * TILE + (warp & 1) * 8);   // origin of this warp's 8x4 pixel block
  const float by0 = (float)(tile_by * TILE + (warp >> 1) * 4);

  const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
  const int warp_last = __reduce_max_sync(0xffffffffu, last_contributor);

  // State of the forward recurrence at the END of this segment.  Behind-colour B = sum_{j >= seg_hi} c_j a_j T_j
  // = C_final - C(before seg_hi); T = transmittance before entry seg_hi.  If nothing contributes at or after
  // seg_hi (it is the tile's last needed segment) the end state is the final state.
  float T_final = 0.f, T = 0.f;
  float B[3] = {0.f, 0.f, 0.f};
  float Bz = 0.f;
  if (inside) {
    const float4 fs = final_state[pix_id];
    T_final = fs.x;
    T = fs.x;
    if (seg_hi < total) {
      const size_t slot = ((size_t)seg_base[tile_id] + (size_t)(seg_hi / SEG)) * TILE_PIX + pix_in_tile;
      const float4 ck = ckpt[slot];
      T = ck.x;
      B[0] = fs.y - ck.y; B[1] = fs.z - ck.z; B[2] = fs.w - ck.w;
      if (EXTRAS) Bz = final_depth[pix_id] - ckpt_z[slot];
    }
  }

  const size_t HW = (size_t)H * W;
  float dpix[3] = {0.f, 0.f, 0.f};
  float ddep = 0.f, dalp = 0.f;
  if (inside) {
    dpix[0] = dL_dpix[0 * HW + pix_id];
    dpix[1] = dL_dpix[1 * HW + pix_id];
    dpix[2] = dL_dpix[2 * HW + pix_id];
    if (EXTRAS) {
      if (dL_ddepth) ddep = dL_ddepth[pix_id];
      if (dL_dalpha_img) dalp = dL_dalpha_img[pix_id];
    }
  }
  // d(out)/dalpha_i through the final transmittance: colour gets +T_final*bg, the alpha image -T_final
  float tail = bg[0] * dpix[0] + bg[1] * dpix[1] + bg[2] * dpix[2];
  if (EXTRAS) tail -= dalp;
  const float ddelx_dx = 0.5f * W;
  const float ddely_dy = 0.5f * H;

  for (int b = 0; b < rounds; ++b) {
    const int stage = b % BL_STAGES;
    const int batch_first_pos = seg_hi - 1 - b * BL_BATCH;  // list position of batch entry 0 (descending)
    mbar_wait(&s_full[stage], (b / BL_STAGES) & 1);
    if (batch_first_pos - (BL_BATCH - 1) < warp_last) {  // else: whole batch lies behind this block's last contributor
      // entries at list positions >= warp_last lie behind this block's last contributor: batch entry e sits at
      // position batch_first_pos - e, so only e > batch_first_pos - warp_last can matter
      int longest;
      const int ncand = cons_classify(max(0, batch_first_pos - warp_last + 1), min(BL_BATCH, count - b * BL_BATCH),
                                      s_xy[stage], s_list[warp], bx0, by0, lane, longest);
      const uint16_t* cand = reinterpret_cast<const uint16_t*>(s_list[warp] + group * LIST_BYTES);
      // Candidates are taken BG at a time (one 16-bit load = two batch-local indices) so the loads / power /
      // exp of one overlap the serial transmittance + behind-colour recurrences of the other.
      {
#pragma unroll 1
        for (int i = 0; i < longest; i += BG) {
          const uint32_t packed = i < ncand ? (uint32_t)cand[i >> 1] : (PAD_WORD & 0xffffu);
          int j[BG];
          bool valid[BG];
          float G[BG], alpha[BG], rinv[BG];
          float2 d[BG];
          float4 con_o[BG], cd[BG];
#pragma unroll
          for (int k = 0; k < BG; ++k) j[k] = (int)((packed >> (8 * k)) & 0xffu);
#pragma unroll
          for (int k = 0; k < BG; ++k) {
            const int pos = batch_first_pos - j[k];  // 0-based list position
            const float4 g = s_xy[stage][j[k]];
            con_o[k] = s_co[stage][j[k]];
            cd[k] = s_cd[stage][j[k]];
            d[k] = {g.x - pixf.x, g.y - pixf.y};
            const float power =
                -0.5f * (con_o[k].x * d[k].x * d[k].x + con_o[k].z * d[k].y * d[k].y) - con_o[k].y * d[k].x * d[k].y;
            G[k] = expf(power);
            alpha[k] = min(0.99f, con_o[k].w * G[k]);
            rinv[k] = __frcp_rn(1.f - alpha[k]);
            valid[k] = (pos < last_contributor) && (power <= 0.0f) && (alpha[k] >= 1.0f / 255.0f);
          }
#pragma unroll
          for (int k = 0; k < BG; ++k) {
            if (!__any_sync(0xffffffffu, valid[k])) continue;
            if (valid[k]) {
              // T holds the transmittance AFTER this entry; Ti before it.  With B the (unnormalised) colour
              // blended behind this entry:  dC/dalpha_i = c_i*Ti - B/(1-alpha_i)
              // (same quantity as backward.cu:515-525's (c - accum_rec)*T, written without the running average)
              const float Ti = T * rinv[k];
              const float w = alpha[k] * Ti;
              float dL_dalpha = (cd[k].x * Ti - B[0] * rinv[k]) * dpix[0] + (cd[k].y * Ti - B[1] * rinv[k]) * dpix[1] +
                                (cd[k].z * Ti - B[2] * rinv[k]) * dpix[2];
              B[0] += cd[k].x * w; B[1] += cd[k].y * w; B[2] += cd[k].z * w;
              float gz = 0.f;
              if (EXTRAS) {
                dL_dalpha += (cd[k].w * Ti - Bz * rinv[k]) * ddep;
                gz = w * ddep;
                Bz += cd[k].w * w;
              }
              T = Ti;
              dL_dalpha += (-T_final * rinv[k]) * tail;

              const float dL_dG = con_o[k].w * dL_dalpha;
              const float gdx = G[k] * d[k].x;
              const float gdy = G[k] * d[k].y;
              const float dG_ddelx = -gdx * con_o[k].x - gdy * con_o[k].y;
              const float dG_ddely = -gdy * con_o[k].z - gdx * con_o[k].y;
              // Every contributing pixel commits its own partials to the packed accumulator row with three
              // 16-byte vector reductions (red.global.add.v4.f32, sm_90+).  Measured on B200: faster than any
              // warp-shuffle pre-reduction (16-shuffle butterfly: 585 us; direct: 429 us) — the L2 reduction
              // units absorb the same-address traffic, the SM issue slots were the bottleneck.
              float* row = grad_acc + (size_t)s_id[stage][j[k]] * GRAD_ACC;
              red_add_v4(row + 0, dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, -0.5f * gdx * d[k].x * dL_dG,
                         -0.5f * gdx * d[k].y * dL_dG);
              red_add_v4(row + 4, -0.5f * gdy * d[k].y * dL_dG, G[k] * dL_dalpha, w * dpix[0], w * dpix[1]);
              if (EXTRAS) red_add_v2(row + 8, w * dpix[2], gz);
              else atomicAdd(row + 8, w * dpix[2]);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[stage]);
  }
}

int launch_blend_bwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s) {
  if (int rc = launch_unit_build(rb, s)) return rc;
  // one CTA per work unit; the grid is sized for the largest capacity, surplus CTAs exit on the device-side count
  uint32_t ucap = 0;
  for (int v = 0; v < rb.V; ++v) ucap = std::max(ucap, rb.v[v].units_cap);
  if (ucap == 0) return 0;
  const dim3 grid(ucap, rb.V, 1);
  if (extras) blend_bwd_kernel<true><<<grid, BL_THREADS, 0, s>>>(rb);
  else blend_bwd_kernel<false><<<grid, BL_THREADS, 0, s>>>(rb);
  count_launch();
  return check_launch("blend_bwd", debug, s);
}

}  // namespace tgr
