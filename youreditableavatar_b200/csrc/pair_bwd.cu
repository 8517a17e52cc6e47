// pair_bwd.cu — gradient of the alpha blending as a FLAT loop over contribution records.
//
// Behavioural reference: renderCUDA (backward), diff-gaussian-rasterization/cuda_rasterizer/backward.cu:399-557.
// The reference (and blend_bwd.cu, kept as the fallback) replays every tile's list back to front with one thread
// per pixel; on avatar scenes a splat touches ~3-5 pixels of a warp's 32, so the replay evaluates ~10 candidates
// per contribution and runs its gradient code at 5/32 lane utilisation (ncu).  But everything a contribution needs
// is known to the FORWARD when it blends the pair: the transmittance T_i and the colour C_i accumulated in front of
// it.  blend_fwd.cu therefore appends a 32-byte record {Gaussian id, lane, T_i, C_i.rgb, D_i, G_i} per blended
// (pixel, Gaussian) into 32-record chunks, and the colour blended BEHIND the pair follows from the final state:
//     B_i = C_final - C_i - c_i * alpha_i * T_i          (backward.cu:515-525's accumulation, in closed form)
// so the backward needs no order at all: one warp per chunk, one lane per record, every lane busy, no exp
// (G_i is stored), and the same three 16-byte vector reductions per contribution as the replay kernel.
// All 32 records of a chunk belong to the 32 pixels of ONE forward warp: lane l first loads pixel l's upstream
// gradients / final state (coalesced), and each record then fetches its pixel's values with warp shuffles.
#include "common.cuh"
#include <algorithm>
#include "pipeline.cuh"

namespace tgr {

__device__ __forceinline__ void pred_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void pred_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int PB_THREADS = 256;

template <bool EXTRAS>
__global__ void __launch_bounds__(PB_THREADS) pair_bwd_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.y];
  if (!rv.use_pairs || rv.unit_count[2] != 0u) return;   // no records, or they overflowed: blend_bwd.cu replays
  const uint32_t nchunks = min(rv.unit_count[1], rv.pair_blocks_cap) * (PAIR_BLOCK / 32);
  const int lane = threadIdx.x & 31;
  const int W = rv.W, H = rv.H;
  const size_t HW = (size_t)H * W;
  const uint32_t tiles_x = (W + TILE - 1) / TILE;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  const float bg0 = rv.bg[0], bg1 = rv.bg[1], bg2 = rv.bg[2];
  const uint32_t warps_total = gridDim.x * (PB_THREADS / 32);
  for (uint32_t chunk = blockIdx.x * (PB_THREADS / 32) + (threadIdx.x >> 5); chunk < nchunks; chunk += warps_total) {
    // ---- this lane's PIXEL of the forward warp that wrote the chunk -------------------------------------
    const uint2 bm = rv.pair_meta[chunk / (PAIR_BLOCK / 32)];          // {tile * 8 + warp, records used} of the block
    const uint32_t first = (chunk % (PAIR_BLOCK / 32)) * 32u;          // first record of this chunk inside the block
    if (first >= bm.y) continue;
    const uint32_t tile_id = bm.x >> 3;
    const int fwarp = (int)(bm.x & 7u);
    int lx, ly, group;
    lane_pixel(fwarp, lane, lx, ly, group);
    const uint32_t px = (tile_id % tiles_x) * TILE + lx, py = (tile_id / tiles_x) * TILE + ly;
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float4 fs = make_float4(0.f, 0.f, 0.f, 0.f);
    float p_d0 = 0.f, p_d1 = 0.f, p_d2 = 0.f, p_dd = 0.f, p_da = 0.f, p_fz = 0.f;
    if (inside) {
      fs = rv.final_state[pix_id];
      p_d0 = rv.dL_dpix[0 * HW + pix_id];
      p_d1 = rv.dL_dpix[1 * HW + pix_id];
      p_d2 = rv.dL_dpix[2 * HW + pix_id];
      if (EXTRAS) {
        if (rv.dL_ddepth) p_dd = rv.dL_ddepth[pix_id];
        if (rv.dL_dalpha) p_da = rv.dL_dalpha[pix_id];
        p_fz = rv.final_z[pix_id];
      }
    }
    // ---- this lane's RECORD -----------------------------------------------------------------------------
    const bool live = first + (uint32_t)lane < bm.y;
    const float4* rec = rv.pairs + ((size_t)chunk * 32 + lane) * 2;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    if (live) { r0 = rec[0]; r1 = rec[1]; }
    const uint32_t id = __float_as_uint(r0.x);
    const int src = live ? (int)(__float_as_uint(r0.y) & 31u) : lane;
    // pixel-side values of the record's pixel
    const float T_final = __shfl_sync(0xffffffffu, fs.x, src);
    const float Cf0 = __shfl_sync(0xffffffffu, fs.y, src), Cf1 = __shfl_sync(0xffffffffu, fs.z, src),
                Cf2 = __shfl_sync(0xffffffffu, fs.w, src);
    const float d0 = __shfl_sync(0xffffffffu, p_d0, src), d1 = __shfl_sync(0xffffffffu, p_d1, src),
                d2 = __shfl_sync(0xffffffffu, p_d2, src);
    float ddep = 0.f, dalp = 0.f, Zf = 0.f;
    if (EXTRAS) {
      ddep = __shfl_sync(0xffffffffu, p_dd, src);
      dalp = __shfl_sync(0xffffffffu, p_da, src);
      Zf = __shfl_sync(0xffffffffu, p_fz, src);
    }
    const float fpx = __shfl_sync(0xffffffffu, (float)px, src), fpy = __shfl_sync(0xffffffffu, (float)py, src);
    if (!live) continue;

    const float4 g = rv.xy_ext[id];
    const float4 con_o = rv.conic_opacity[id];
    const float4 cd = rv.rgb_depth[id];
    const float Ti = r0.z;                       // transmittance in front of this contribution
    const float G = r1.w;
    const float dx = g.x - fpx, dy = g.y - fpy;
    const float alpha = min(0.99f, con_o.w * G);
    const float rinv = __frcp_rn(1.f - alpha);
    const float w = alpha * Ti;
    // colour blended behind the pair: final - in front - own
    const float B0 = Cf0 - r0.w - cd.x * w, B1 = Cf1 - r1.x - cd.y * w, B2 = Cf2 - r1.y - cd.z * w;
    float tail = bg0 * d0 + bg1 * d1 + bg2 * d2;
    if (EXTRAS) tail -= dalp;
    float dL_dalpha = (cd.x * Ti - B0 * rinv) * d0 + (cd.y * Ti - B1 * rinv) * d1 + (cd.z * Ti - B2 * rinv) * d2;
    float gz = 0.f;
    if (EXTRAS) {
      const float Bz = Zf - r1.z - cd.w * w;
      dL_dalpha += (cd.w * Ti - Bz * rinv) * ddep;
      gz = w * ddep;
    }
    dL_dalpha += (-T_final * rinv) * tail;

    const float dL_dG = con_o.w * dL_dalpha;
    const float gdx = G * dx, gdy = G * dy;
    const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
    const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
    float* row = rv.grad_acc + (size_t)id * GRAD_ACC;
    pred_add_v4(row + 0, dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, -0.5f * gdx * dx * dL_dG,
                -0.5f * gdx * dy * dL_dG);
    pred_add_v4(row + 4, -0.5f * gdy * dy * dL_dG, G * dL_dalpha, w * d0, w * d1);
    if (EXTRAS) pred_add_v2(row + 8, w * d2, gz);
    else atomicAdd(row + 8, w * d2);
  }
}

int launch_pair_bwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s) {
  uint32_t cap = 0;
  bool any = false;
  for (int v = 0; v < rb.V; ++v) {
    if (!rb.v[v].use_pairs) continue;
    any = true;
    cap = std::max(cap, rb.v[v].pair_blocks_cap * (uint32_t)(PAIR_BLOCK / 32));
  }
  if (!any || cap == 0) return 0;
  // a few resident CTAs per SM, grid-stride over the chunks actually used (device-side count)
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)cap + 7) / 8, (uint64_t)NUM_SM * 8);
  const dim3 grid(blocks, rb.V, 1);
  if (extras) pair_bwd_kernel<true><<<grid, PB_THREADS, 0, s>>>(rb);
  else pair_bwd_kernel<false><<<grid, PB_THREADS, 0, s>>>(rb);
  count_launch();
  return check_launch("pair_bwd", debug, s);
}

}  // namespace tgr
