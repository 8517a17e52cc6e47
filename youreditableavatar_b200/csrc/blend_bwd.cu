// blend_bwd.cu — gradient of the alpha blending, one CTA per 16x16 tile, back-to-front replay.
//
// Behavioural reference: renderCUDA (backward), diff-gaussian-rasterization/cuda_rasterizer/backward.cu:399-557.
// Differences in *how* (results agree to fp32 rounding / atomics order):
//   * the replay starts at the tile's last contributing list entry (tile_last, written by the forward)
//     instead of the end of the tile list, so the never-blended tail is not even staged;
//   * per-Gaussian partial gradients are reduced across the warp with shuffles and committed with one
//     lane's atomics per warp (the reference issues 9 global float atomics per (pixel, Gaussian) pair);
//   * all per-Gaussian 2D gradients land in one packed accumulator row grad_acc[g][12]
//     {dmean2D.x, dmean2D.y, dconic.xx, dconic.xy, dconic.yy, dopacity, dr, dg, db, dz, -, -};
//   * extras: gradients of the depth / alpha images flow through the same recurrence.
#include "common.cuh"

namespace tgr {

constexpr int BB = 256;

template <bool EXTRAS>
__global__ void __launch_bounds__(BB) blend_bwd_kernel(const uint2* __restrict__ ranges,
                                                       const uint32_t* __restrict__ point_list, int W, int H,
                                                       const float* __restrict__ bg, const float2* __restrict__ xy,
                                                       const float4* __restrict__ conic_opacity,
                                                       const float4* __restrict__ rgb_depth,
                                                       const float* __restrict__ final_T,
                                                       const uint32_t* __restrict__ n_contrib,
                                                       const uint32_t* __restrict__ tile_last,
                                                       const float* __restrict__ dL_dpix,
                                                       const float* __restrict__ dL_ddepth,
                                                       const float* __restrict__ dL_dalpha_img,
                                                       float* __restrict__ grad_acc) {
  __shared__ uint32_t s_id[BB];
  __shared__ float2 s_xy[BB];
  __shared__ float4 s_co[BB];
  __shared__ float4 s_cd[BB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tiles_x = (W + TILE - 1) / TILE;
  const uint32_t tile_id = blockIdx.y * tiles_x + blockIdx.x;
  const uint32_t px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * TILE + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float2 pixf = {(float)px, (float)py};

  const uint2 range = ranges[tile_id];
  const int total = (int)min(tile_last[tile_id], range.y - range.x);  // entries [0,total) can contribute
  if (total == 0) return;
  const int rounds = (total + BB - 1) / BB;

  const float T_final = inside ? final_T[pix_id] : 0.f;
  float T = T_final;
  const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

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
  // d(out)/dalpha_i through the final transmittance: colour gets +T_final*bg, alpha image gets -T_final
  const float bg_dot_dpixel = bg[0] * dpix[0] + bg[1] * dpix[1] + bg[2] * dpix[2];

  float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
  float accum_z = 0.f, last_z = 0.f;
  float last_alpha = 0.f;
  const float ddelx_dx = 0.5f * W;
  const float ddely_dy = 0.5f * H;

  int idx = total;  // list position (0-based) of the entry about to be visited is idx-1
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    const int progress = r * BB + tid;
    if (progress < total) {
      const uint32_t id = point_list[range.x + (total - 1 - progress)];
      s_id[tid] = id;
      s_xy[tid] = xy[id];
      s_co[tid] = conic_opacity[id];
      s_cd[tid] = rgb_depth[id];
    }
    __syncthreads();
    const int nb = min(BB, total - r * BB);
    for (int j = 0; j < nb; ++j) {
      --idx;
      bool valid = inside && idx < last_contributor;
      float G = 0.f, alpha = 0.f;
      float2 d = {0.f, 0.f};
      float4 con_o = s_co[j];
      if (valid) {
        const float2 m = s_xy[j];
        d = {m.x - pixf.x, m.y - pixf.y};
        const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
        if (power > 0.0f) valid = false;
        else {
          G = expf(power);
          alpha = min(0.99f, con_o.w * G);
          if (alpha < 1.0f / 255.0f) valid = false;
        }
      }
      if (!__any_sync(0xffffffffu, valid)) continue;

      float g_mx = 0.f, g_my = 0.f, g_cx = 0.f, g_cy = 0.f, g_cw = 0.f, g_op = 0.f;
      float g_c0 = 0.f, g_c1 = 0.f, g_c2 = 0.f, g_z = 0.f;
      if (valid) {
        T = T / (1.f - alpha);
        const float dchannel_dcolor = alpha * T;
        const float4 cd = s_cd[j];
        const float c[3] = {cd.x, cd.y, cd.z};
        float dL_dalpha = 0.0f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
          last_color[ch] = c[ch];
          dL_dalpha += (c[ch] - accum_rec[ch]) * dpix[ch];
        }
        g_c0 = dchannel_dcolor * dpix[0];
        g_c1 = dchannel_dcolor * dpix[1];
        g_c2 = dchannel_dcolor * dpix[2];
        if (EXTRAS) {
          accum_z = last_alpha * last_z + (1.f - last_alpha) * accum_z;
          last_z = cd.w;
          dL_dalpha += (cd.w - accum_z) * ddep;
          g_z = dchannel_dcolor * ddep;
        }
        dL_dalpha *= T;
        last_alpha = alpha;
        float tail = bg_dot_dpixel;
        if (EXTRAS) tail -= dalp;
        dL_dalpha += (-T_final / (1.f - alpha)) * tail;

        const float dL_dG = con_o.w * dL_dalpha;
        const float gdx = G * d.x;
        const float gdy = G * d.y;
        const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
        const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
        g_mx = dL_dG * dG_ddelx * ddelx_dx;
        g_my = dL_dG * dG_ddely * ddely_dy;
        g_cx = -0.5f * gdx * d.x * dL_dG;
        g_cy = -0.5f * gdx * d.y * dL_dG;
        g_cw = -0.5f * gdy * d.y * dL_dG;
        g_op = G * dL_dalpha;
      }
      // warp reduction, then one lane commits
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        g_mx += __shfl_xor_sync(0xffffffffu, g_mx, o);
        g_my += __shfl_xor_sync(0xffffffffu, g_my, o);
        g_cx += __shfl_xor_sync(0xffffffffu, g_cx, o);
        g_cy += __shfl_xor_sync(0xffffffffu, g_cy, o);
        g_cw += __shfl_xor_sync(0xffffffffu, g_cw, o);
        g_op += __shfl_xor_sync(0xffffffffu, g_op, o);
        g_c0 += __shfl_xor_sync(0xffffffffu, g_c0, o);
        g_c1 += __shfl_xor_sync(0xffffffffu, g_c1, o);
        g_c2 += __shfl_xor_sync(0xffffffffu, g_c2, o);
        if (EXTRAS) g_z += __shfl_xor_sync(0xffffffffu, g_z, o);
      }
      float* acc = grad_acc + (size_t)s_id[j] * GRAD_ACC;
      // spread the 9-10 atomics over lanes so they issue in one instruction
      float v = 0.f;
      switch (lane) {
        case 0: v = g_mx; break; case 1: v = g_my; break; case 2: v = g_cx; break; case 3: v = g_cy; break;
        case 4: v = g_cw; break; case 5: v = g_op; break; case 6: v = g_c0; break; case 7: v = g_c1; break;
        case 8: v = g_c2; break; case 9: v = g_z; break; default: break;
      }
      if (lane < (EXTRAS ? 10 : 9)) atomicAdd(acc + lane, v);
    }
  }
}

int launch_blend_bwd(const tgr_params& p, const GeomView& g, const uint32_t* point_list, const ImageView& im,
                     float* grad_acc, cudaStream_t s) {
  dim3 grid((p.W + TILE - 1) / TILE, (p.H + TILE - 1) / TILE, 1);
  const bool ex = p.extras && (p.dL_dout_depth || p.dL_dout_alpha);
  if (ex)
    blend_bwd_kernel<true><<<grid, BB, 0, s>>>(im.ranges, point_list, p.W, p.H, p.background, g.xy, g.conic_opacity,
                                               g.rgb_depth, im.final_T, im.n_contrib, im.tile_last, p.dL_dout_color,
                                               p.dL_dout_depth, p.dL_dout_alpha, grad_acc);
  else
    blend_bwd_kernel<false><<<grid, BB, 0, s>>>(im.ranges, point_list, p.W, p.H, p.background, g.xy, g.conic_opacity,
                                                g.rgb_depth, im.final_T, im.n_contrib, im.tile_last, p.dL_dout_color,
                                                nullptr, nullptr, grad_acc);
  return check_launch("blend_bwd", p.debug != 0, s);
}

}  // namespace tgr
