// blend_fwd.cu — front-to-back alpha blending, one CTA per 16x16 tile.
//
// Behavioural reference: renderCUDA, diff-gaussian-rasterization/cuda_rasterizer/forward.cu:261-374.
// The per-pair arithmetic (power, alpha = min(0.99, o*exp(power)), the 1/255 and 1e-4 thresholds, colour
// accumulation order) is evaluated with the same fp32 expressions so the image is bit-identical.
// Extras (new): depth = sum z_i alpha_i T_i, alpha = 1 - T_final.
//
// How it differs from the reference kernel:
//   * 256 threads; warp w owns the 8x4 pixel block at (8*(w&1), 4*(w>>1)) of the tile.  While a batch of
//     256 list entries is staged, every staging thread also tests its Gaussian's conservative footprint
//     (half-extents of the alpha >= 1/255 ellipse, computed once per Gaussian by the preprocess kernel)
//     against the eight warp blocks; eight __ballot_sync per warp turn that into a 256-bit "overlaps my
//     block" mask per consumer warp.  A warp then only evaluates the Gaussians whose footprint reaches its
//     block — pairs it skips would have failed the reference's alpha < 1/255 test, so results are unchanged.
//   * staging is a gather (point_list -> per-Gaussian records); it uses 16-byte cp.async copies into a
//     double-buffered shared-memory batch so the gather of batch b+1 overlaps the blending of batch b.
//   * the tile's maximum contributing list position is written out for the backward (tile_last).
#include "common.cuh"

namespace tgr {

constexpr int FB = 256;  // batch size == threads per CTA

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 8-bit mask of the warp blocks (bit = 2*row4 + col8) a footprint [x-hx,x+hx]x[y-hy,y+hy] can reach.
__device__ __forceinline__ uint32_t block_mask(float x, float y, float hx, float hy, float tile_x0, float tile_y0) {
  if (!(hx >= 0.f)) return 0u;  // never reaches alpha >= 1/255 (or NaN extents)
  // pixel centres are integers: pixels p with x-hx <= p <= x+hx
  const float fx0 = ceilf(x - hx) - tile_x0, fx1 = floorf(x + hx) - tile_x0;
  const float fy0 = ceilf(y - hy) - tile_y0, fy1 = floorf(y + hy) - tile_y0;
  if (fx1 < 0.f || fy1 < 0.f || fx0 > 15.f || fy0 > 15.f || fx0 > fx1 || fy0 > fy1) return 0u;
  const int x0 = (int)fmaxf(fx0, 0.f), x1 = (int)fminf(fx1, 15.f);
  const int y0 = (int)fmaxf(fy0, 0.f), y1 = (int)fminf(fy1, 15.f);
  const uint32_t colm = ((x0 < 8) ? 1u : 0u) | ((x1 >= 8) ? 2u : 0u);      // which 8-wide columns
  const int r0 = y0 >> 2, r1 = y1 >> 2;                                       // which 4-high rows
  uint32_t m = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r >= r0 && r <= r1) m |= colm << (2 * r);
  return m;
}

template <bool EXTRAS>
__global__ void __launch_bounds__(FB) blend_fwd_kernel(const uint2* __restrict__ ranges,
                                                       const uint32_t* __restrict__ point_list, int W, int H,
                                                       const float4* __restrict__ xy_ext,
                                                       const float4* __restrict__ conic_opacity,
                                                       const float4* __restrict__ rgb_depth,
                                                       const float* __restrict__ bg, float* __restrict__ final_T,
                                                       uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_last,
                                                       float* __restrict__ out_color, float* __restrict__ out_depth,
                                                       float* __restrict__ out_alpha) {
  __shared__ __align__(16) float4 s_xy[2][FB];   // x, y, hx, hy
  __shared__ __align__(16) float4 s_co[2][FB];   // conic xx, xy, yy, opacity
  __shared__ __align__(16) float4 s_cd[2][FB];   // r, g, b, depth
  __shared__ uint32_t s_ball[2][8][FB / 32];     // [buf][consumer block][producer warp]
  __shared__ uint32_t s_last[FB / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tiles_x = (W + TILE - 1) / TILE;
  const uint32_t tile_id = blockIdx.y * tiles_x + blockIdx.x;
  const uint32_t px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * TILE + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float2 pixf = {(float)px, (float)py};
  const float tile_x0 = (float)(blockIdx.x * TILE), tile_y0 = (float)(blockIdx.y * TILE);

  const uint2 range = ranges[tile_id];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + FB - 1) / FB;

  bool done = !inside;
  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float Dz = 0.f;

  auto issue = [&](int round, int buf) {  // gather batch `round` into buffer `buf` (asynchronously)
    const int progress = round * FB + tid;
    if (progress < total) {
      const uint32_t id = point_list[range.x + progress];
      cp_async16(&s_xy[buf][tid], &xy_ext[id]);
      cp_async16(&s_co[buf][tid], &conic_opacity[id]);
      cp_async16(&s_cd[buf][tid], &rgb_depth[id]);
    }
    cp_async_commit();
  };
  if (rounds > 0) issue(0, 0);

  for (int i = 0; i < rounds; ++i) {
    const int buf = i & 1;
    cp_async_wait<0>();
    // own record has landed: classify it against the eight warp blocks
    uint32_t mymask = 0;
    if (i * FB + tid < total) {
      const float4 g = s_xy[buf][tid];
      mymask = block_mask(g.x, g.y, g.z, g.w, tile_x0, tile_y0);
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (mymask >> b) & 1u);
      if (lane == 0) s_ball[buf][b][warp] = bal;
    }
    // all records + masks of batch i visible; nobody still reads buffer buf^1 (batch i-1)
    const int num_done = __syncthreads_count(done);
    if (num_done == FB) break;
    if (i + 1 < rounds) issue(i + 1, buf ^ 1);

    if (!__all_sync(0xffffffffu, done)) {
      const uint32_t base_pos = (uint32_t)(i * FB);
#pragma unroll 1
      for (int w8 = 0; w8 < FB / 32; ++w8) {
        uint32_t m = s_ball[buf][warp][w8];
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int j = w8 * 32 + bit;
          if (!done) {
            const float4 g = s_xy[buf][j];
            const float2 d = {g.x - pixf.x, g.y - pixf.y};
            const float4 con_o = s_co[buf][j];
            const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
            if (power <= 0.0f) {
              const float alpha = min(0.99f, con_o.w * expf(power));
              if (alpha >= 1.0f / 255.0f) {
                const float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) {
                  done = true;
                } else {
                  const float4 cd = s_cd[buf][j];
                  C[0] += cd.x * alpha * T;
                  C[1] += cd.y * alpha * T;
                  C[2] += cd.z * alpha * T;
                  if (EXTRAS) Dz += cd.w * alpha * T;
                  T = test_T;
                  last_contributor = base_pos + j + 1;
                }
              }
            }
          }
        }
        if (__all_sync(0xffffffffu, done)) break;
      }
    }
  }
  cp_async_wait<0>();

  if (inside) {
    final_T[pix_id] = T;
    n_contrib[pix_id] = last_contributor;
    const size_t HW = (size_t)H * W;
    out_color[0 * HW + pix_id] = C[0] + T * bg[0];
    out_color[1 * HW + pix_id] = C[1] + T * bg[1];
    out_color[2 * HW + pix_id] = C[2] + T * bg[2];
    if (EXTRAS) {
      out_depth[pix_id] = Dz;
      out_alpha[pix_id] = 1.0f - T;
    }
  }
  // tile-wide maximum of last_contributor: lets the backward start at the last useful list entry
  uint32_t wl = __reduce_max_sync(0xffffffffu, inside ? last_contributor : 0u);
  if (lane == 0) s_last[warp] = wl;
  __syncthreads();
  if (tid == 0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < FB / 32; ++w) m = max(m, s_last[w]);
    tile_last[tile_id] = m;
  }
}

int launch_blend_fwd(const tgr_params& p, const GeomView& g, const uint32_t* point_list, const ImageView& im,
                     cudaStream_t s) {
  dim3 grid((p.W + TILE - 1) / TILE, (p.H + TILE - 1) / TILE, 1);
  if (p.extras && p.out_depth && p.out_alpha)
    blend_fwd_kernel<true><<<grid, FB, 0, s>>>(im.ranges, point_list, p.W, p.H, g.xy_ext, g.conic_opacity, g.rgb_depth,
                                               p.background, im.final_T, im.n_contrib, im.tile_last, p.out_color,
                                               p.out_depth, p.out_alpha);
  else
    blend_fwd_kernel<false><<<grid, FB, 0, s>>>(im.ranges, point_list, p.W, p.H, g.xy_ext, g.conic_opacity, g.rgb_depth,
                                                p.background, im.final_T, im.n_contrib, im.tile_last, p.out_color,
                                                nullptr, nullptr);
  return check_launch("blend_fwd", p.debug != 0, s);
}

}  // namespace tgr
