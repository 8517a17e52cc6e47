// preprocess.cu — per-Gaussian forward stage: (optional mesh binding + activations) -> cull -> project ->
// 3D/2D covariance -> conic, radius, tile rectangle -> SH->RGB.  One thread per Gaussian, SoA outputs.
//
// Behavioural reference (fp32 expression order reproduced so keys/radii are bit-identical):
//   diff-gaussian-rasterization/cuda_rasterizer/forward.cu:20-71   SH -> RGB with clamp flags
//   forward.cu:74-113  EWA 2D covariance (+0.3 low-pass)    forward.cu:118-152  3D covariance from scale/quat
//   forward.cu:155-256 preprocess                           auxiliary.h:41-56,139-164  ndc2Pix/getRect/in_frustum
// Binding (opt-in, fused): Edit_core/tetgs_scene/tetgs_model.py:252-286 (points = ori + normals*delta,
//   exp / sigmoid / normalize activations), barycentric attributes tetgs_model.py:335-377.
#include "common.cuh"

namespace tgr {

struct Cov3 { float c[6]; };

// Sigma = (S R)^T (S R) with glm's column-major conventions (forward.cu:118-152)
__device__ __forceinline__ Cov3 cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot) {
  M3 S = m3(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
  S.m[0][0] = mod * scale.x;
  S.m[1][1] = mod * scale.y;
  S.m[2][2] = mod * scale.z;
  float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
  M3 R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
            2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
            2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
  M3 M = m3_mul(S, R);
  M3 Sigma = m3_mul(m3_T(M), M);
  Cov3 o;
  o.c[0] = Sigma.m[0][0]; o.c[1] = Sigma.m[0][1]; o.c[2] = Sigma.m[0][2];
  o.c[3] = Sigma.m[1][1]; o.c[4] = Sigma.m[1][2]; o.c[5] = Sigma.m[2][2];
  return o;
}

// EWA projection of the 3D covariance (forward.cu:74-113)
__device__ __forceinline__ float3 cov2d(const float3& mean, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                        const float* cov3D, const float* __restrict__ view) {
  float3 t = xform4x3(mean, view);
  const float limx = 1.3f * tan_fovx;
  const float limy = 1.3f * tan_fovy;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = min(limx, max(-limx, txtz)) * t.z;
  t.y = min(limy, max(-limy, tytz)) * t.z;

  M3 J = m3(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
            0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
            0.0f, 0.0f, 0.0f);
  M3 Wm = m3(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
  M3 T = m3_mul(Wm, J);
  M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
  M3 cov = m3_mul(m3_mul(m3_T(T), m3_T(Vrk)), T);
  cov.m[0][0] += 0.3f;
  cov.m[1][1] += 0.3f;
  return {cov.m[0][0], cov.m[0][1], cov.m[1][1]};
}

// SH -> RGB (forward.cu:20-71).  `sh` holds (deg+1)^2 coefficient triples.
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* sh, float3 pos, float3 campos, uint8_t& clamped) {
  float3 dir = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
  float tx = dir.x * dir.x, ty = dir.y * dir.y, tz = dir.z * dir.z;
  float len = sqrtf(tx + ty + tz);
  dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
  float res[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) res[c] = SH_C0 * sh[c];
  if (deg > 0) {
    float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      res[c] = res[c] - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        res[c] = res[c] + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] +
                 SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + SH_C2[3] * xz * sh[21 + c] +
                 SH_C2[4] * (xx - yy) * sh[24 + c];
      if (deg > 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          res[c] = res[c] + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
                   SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                   SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                   SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
                   SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
      }
    }
  }
  clamped = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    res[c] += 0.5f;
    if (res[c] < 0) clamped |= (uint8_t)(1u << c);
    res[c] = fmaxf(res[c], 0.0f);
  }
  return {res[0], res[1], res[2]};
}

// Loads the first ncoef SH triples of Gaussian idx. 16-byte vector loads when the row is 16-byte aligned.
__device__ __forceinline__ void load_sh(const float* __restrict__ shs, size_t idx, int M, int ncoef, float* sh) {
  const float* base = shs + idx * (size_t)M * 3;
  const int nf = ncoef * 3;
  const bool vec_ok = (((M * 3) & 3) == 0) && ((reinterpret_cast<uintptr_t>(shs) & 15) == 0) && (((nf + 3) & ~3) <= M * 3);
  if (vec_ok) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    const int nv = (nf + 3) >> 2;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      if (k < nv) {
        float4 v = __ldg(b4 + k);
        sh[4 * k + 0] = v.x; sh[4 * k + 1] = v.y; sh[4 * k + 2] = v.z; sh[4 * k + 3] = v.w;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 48; ++k)
      if (k < nf) sh[k] = __ldg(base + k);
  }
}

constexpr int SH_ROW = 48;        // floats per SH row at M = 16
constexpr int SH_ROW_PAD = 49;    // padded shared-memory row stride (bank-conflict-free per-thread rows)

// STAGED: the block's 256 SH rows (48 KB, contiguous in HBM) are moved global -> shared with fully coalesced
// 16-byte loads; each thread then reads its own padded row.  Used when M == 16 and degree >= 2.
// The kernel serves a BATCH of views (ViewBatch, common.cuh): the Gaussian's parameters, its SH row and its 3D
// covariance are fetched / computed once and every view of the batch is projected from registers.
template <bool BOUND, bool STAGED>
__global__ void __launch_bounds__(256) preprocess_kernel(const tgr_params p, const tgr_binding bind,
                                                         const __grid_constant__ ViewBatch vb) {
  extern __shared__ float s_rows[];
  __shared__ float s_cam[MAX_BATCH][CAM_FLOATS];
  __shared__ uint32_t s_cnt[MAX_BATCH][4];   // instances, visible Gaussians, OR / AND of the visible depth keys
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = p.P;
  const int V = vb.V;
  for (int e = threadIdx.x; e < V * CAM_FLOATS; e += blockDim.x) {
    const int v = e / CAM_FLOATS, k = e % CAM_FLOATS;
    float x = 0.f;
    if (k < 16) x = vb.v[v].viewmatrix[k];
    else if (k < 32) x = vb.v[v].projmatrix[k - 16];
    else if (k < 35) x = vb.v[v].campos[k - 32];
    s_cam[v][k] = x;
  }
  if (threadIdx.x < 4 * MAX_BATCH) s_cnt[threadIdx.x >> 2][threadIdx.x & 3] = (threadIdx.x & 3) == 3 ? 0xffffffffu : 0u;
  // on the side: clear the temp area of the depth sort that follows (histograms, tickets, look-back flags); block b
  // takes slice b — this replaces one memset node per view
  for (int v = 0; v < V; ++v) {
    const uint32_t words = vb.v[v].sort_zero_words;
    const uint32_t per = ((words + gridDim.x - 1) / gridDim.x + 3u) & ~3u;
    const uint32_t lo = blockIdx.x * per, hi = min(lo + per, words);
    uint4* t4 = reinterpret_cast<uint4*>(vb.v[v].sort_temp);
    for (uint32_t i = lo / 4 + threadIdx.x; i * 4 < hi; i += blockDim.x) t4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (STAGED) {
    const int block_first = blockIdx.x * blockDim.x;
    const int nrows = min((int)blockDim.x, P - block_first);
    const float4* src = reinterpret_cast<const float4*>(p.shs + (size_t)block_first * SH_ROW);
    for (int v = threadIdx.x; v < nrows * (SH_ROW / 4); v += blockDim.x) {
      const float4 q = __ldg(src + v);
      const int e = v * 4;
      float* d = s_rows + (e / SH_ROW) * SH_ROW_PAD + (e % SH_ROW);
      d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    }
  }
  __syncthreads();

  const bool live = idx < P;
  float3 p_orig = {0.f, 0.f, 0.f};
  float3 scale = {0.f, 0.f, 0.f};
  float4 rot = {1.f, 0.f, 0.f, 0.f};
  float opacity = 0.f;
  Cov3 c3;
#pragma unroll
  for (int k = 0; k < 6; ++k) c3.c[k] = 0.f;
  float sh_local[STAGED ? 1 : 48];
  float hx_tau = 0.f;
  if (live) {
    if (BOUND) {
      const float d = bind.delta ? bind.delta[idx] : 0.f;
      float o[3], n[3];
      if (bind.origins != nullptr) {
        // direct form: mean = origin + normal * delta with per-Gaussian constants (tetgs_model.py:252-258 on the stored
        // ori_points / normals; tetgs_edit_3d.py:272-280)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[c] = bind.origins[3 * idx + c];
          n[c] = bind.normals ? bind.normals[3 * idx + c] : 0.f;
        }
      } else {
        // mean = sum_k w_k V[f_k] + (sum_k w_k N[f_k]) * delta      (tetgs_model.py:252-258, 335-377)
        const int f = bind.face_index[idx];
        const int i0 = bind.faces[3 * f + 0], i1 = bind.faces[3 * f + 1], i2 = bind.faces[3 * f + 2];
        const float w0 = bind.bary[3 * idx + 0], w1 = bind.bary[3 * idx + 1], w2 = bind.bary[3 * idx + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[c] = w0 * bind.verts[3 * i0 + c] + w1 * bind.verts[3 * i1 + c] + w2 * bind.verts[3 * i2 + c];
          n[c] = w0 * bind.vert_normals[3 * i0 + c] + w1 * bind.vert_normals[3 * i1 + c] + w2 * bind.vert_normals[3 * i2 + c];
        }
      }
      p_orig = {o[0] + n[0] * d, o[1] + n[1] * d, o[2] + n[2] * d};
      scale = {expf(bind.log_scales[3 * idx + 0]), expf(bind.log_scales[3 * idx + 1]), expf(bind.log_scales[3 * idx + 2])};
      float4 q = reinterpret_cast<const float4*>(bind.raw_quats)[idx];
      float qn = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);  // F.normalize eps
      rot = {q.x / qn, q.y / qn, q.z / qn, q.w / qn};
      opacity = 1.0f / (1.0f + expf(-bind.opacity_logits[idx]));
      bind.out_means3D[3 * idx + 0] = p_orig.x; bind.out_means3D[3 * idx + 1] = p_orig.y; bind.out_means3D[3 * idx + 2] = p_orig.z;
      bind.out_scales[3 * idx + 0] = scale.x; bind.out_scales[3 * idx + 1] = scale.y; bind.out_scales[3 * idx + 2] = scale.z;
      reinterpret_cast<float4*>(bind.out_rotations)[idx] = rot;
      bind.out_opacities[idx] = opacity;
    } else {
      p_orig = {p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
      opacity = p.opacities[idx];
    }
    if (p.cov3D_precomp != nullptr && !BOUND) {
#pragma unroll
      for (int k = 0; k < 6; ++k) c3.c[k] = p.cov3D_precomp[6 * (size_t)idx + k];
    } else {
      if (!BOUND) {
        scale = {p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]};
        rot = reinterpret_cast<const float4*>(p.rotations)[idx];
      }
      c3 = cov3d_from_scale_rot(scale, p.scale_modifier, rot);
    }
    if (!STAGED && p.colors_precomp == nullptr) {
      const int ncoef = (p.D + 1) * (p.D + 1);
      load_sh(p.shs, (size_t)idx, p.M, ncoef, sh_local);
    }
    // view-independent part of the alpha >= 1/255 footprint (see below): tau = ln(255 o) with margins
    hx_tau = (opacity < 1.0f / 255.0f) ? -1.f : (logf(255.0f * opacity) * 1.01f + 0.01f);
  }

  for (int v = 0; v < V; ++v) {
    const ViewDesc& vd = vb.v[v];
    const float* view = s_cam[v];
    const float* proj = s_cam[v] + 16;
    uint32_t my_tiles = 0, my_vis = 0, my_key = 0;
    if (live) {
      int radius_out = 0;
      ushort4 rect_out = make_ushort4(0, 0, 0, 0);
      uint32_t key_out = 0x7fffffffu;

      // near culling (auxiliary.h:139-164): dropped iff view-space z <= 0.2 (so a NaN depth passes, as it does there)
      const float3 p_view = xform4x3(p_orig, view);
      const bool culled = p_view.z <= 0.2f;
      if (culled && vd.prefiltered) vd.header->prefilter_violation = 1u;   // auxiliary.h:154-160
      if (!culled) {
        const float4 p_hom = xform4x4(p_orig, proj);
        const float p_w = 1.0f / (p_hom.w + 0.0000001f);
        const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

        const float focal_y = vd.H / (2.0f * vd.tan_fovy);
        const float focal_x = vd.W / (2.0f * vd.tan_fovx);
        const float3 cov = cov2d(p_orig, focal_x, focal_y, vd.tan_fovx, vd.tan_fovy, c3.c, view);

        const float det = (cov.x * cov.z - cov.y * cov.y);
        if (det != 0.0f) {
          const float det_inv = 1.f / det;
          const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};
          const float mid = 0.5f * (cov.x + cov.z);
          const float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
          const float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
          const float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
          const float2 point_image = {ndc2pix(p_proj.x, vd.W), ndc2pix(p_proj.y, vd.H)};
          const uint32_t gx = (vd.W + TILE - 1) / TILE, gy = (vd.H + TILE - 1) / TILE;
          uint2 rmin, rmax;
          tile_rect(point_image, (int)my_radius, gx, gy, rmin, rmax);
          const uint32_t ntiles = (rmax.x - rmin.x) * (rmax.y - rmin.y);
          if (ntiles != 0) {
            float3 rgb;
            uint8_t clamped = 0;
            if (p.colors_precomp != nullptr) {
              rgb = {p.colors_precomp[3 * idx], p.colors_precomp[3 * idx + 1], p.colors_precomp[3 * idx + 2]};
            } else {
              const float3 cam = {view[32], view[33], view[34]};
              rgb = sh_to_rgb(p.D, STAGED ? (s_rows + threadIdx.x * SH_ROW_PAD) : sh_local, p_orig, cam, clamped);
            }
            // Conservative footprint of {alpha >= 1/255}: |dx| <= sqrt(2 tau Sxx), |dy| <= sqrt(2 tau Syy) with
            // tau = ln(255 o) and S the inverse of the conic actually used by the blend kernels.  Margins cover
            // fp32 rounding of power/exp; ill-conditioned conics disable culling (huge extents).  Pairs outside
            // this box fail the reference's alpha < 1/255 test (forward.cu:343-345), so skipping them is exact.
            float hx = -1.f, hy = -1.f;
            if (!(opacity < 1.0f / 255.0f)) {
              const float tau = hx_tau;
              const float ac = conic.x * conic.z, bb = conic.y * conic.y;
              const float det_lo = (ac - bb) - 1e-6f * (fabsf(ac) + bb);
              if (det_lo > 0.f && ac <= 1000.f * det_lo && tau < 1e30f) {
                hx = sqrtf(2.f * tau * conic.z / det_lo) * 1.001f + 0.01f;
                hy = sqrtf(2.f * tau * conic.x / det_lo) * 1.001f + 0.01f;
              } else {
                hx = hy = 1e30f;
              }
            }
            vd.xy_ext[idx] = make_float4(point_image.x, point_image.y, hx, hy);
            vd.conic_opacity[idx] = {conic.x, conic.y, conic.z, opacity};
            vd.rgb_depth[idx] = {rgb.x, rgb.y, rgb.z, p_view.z};
            vd.clamped[idx] = clamped;
            radius_out = (int)my_radius;
            rect_out = make_ushort4((unsigned short)rmin.x, (unsigned short)rmin.y, (unsigned short)rmax.x,
                                    (unsigned short)rmax.y);
            key_out = __float_as_uint(p_view.z);
            my_tiles = ntiles;
            my_vis = 1;
            my_key = key_out;
          }
        }
      }
      vd.radii[idx] = radius_out;
      vd.rect[idx] = rect_out;
      vd.depth_key[idx] = key_out;
    }
    // instance count of this view: warp reduce -> shared atomics -> one global atomic per block (below)
    const uint32_t wt = __reduce_add_sync(0xffffffffu, my_tiles);
    const uint32_t wv = __reduce_add_sync(0xffffffffu, my_vis);
    const uint32_t wo = __reduce_or_sync(0xffffffffu, my_key);
    const uint32_t wa = __reduce_and_sync(0xffffffffu, my_vis ? my_key : 0xffffffffu);
    if ((threadIdx.x & 31) == 0) {
      if (wt) atomicAdd(&s_cnt[v][0], wt);
      if (wv) { atomicAdd(&s_cnt[v][1], wv); atomicOr(&s_cnt[v][2], wo); atomicAnd(&s_cnt[v][3], wa); }
    }
  }
  __syncthreads();
  if (threadIdx.x < V) {
    const uint32_t t = s_cnt[threadIdx.x][0], n = s_cnt[threadIdx.x][1];
    if (t) atomicAdd(&vb.v[threadIdx.x].header->num_rendered, t);
    if (n) {
      atomicAdd(&vb.v[threadIdx.x].header->num_visible, n);
      atomicOr(&vb.v[threadIdx.x].header->key_or, s_cnt[threadIdx.x][2]);
      atomicAnd(&vb.v[threadIdx.x].header->key_and, s_cnt[threadIdx.x][3]);
    }
  }
}

ViewDesc make_view_desc(const tgr_params& p, const GeomView& g, const float* grad_acc) {
  ViewDesc d{};
  d.viewmatrix = p.viewmatrix; d.projmatrix = p.projmatrix; d.campos = p.campos;
  d.tan_fovx = p.tan_fovx; d.tan_fovy = p.tan_fovy; d.W = p.W; d.H = p.H; d.prefiltered = p.prefiltered;
  d.sort_temp = g.sort_temp;
  d.sort_zero_words = (uint32_t)sort_zero_words((uint64_t)(p.P > 0 ? p.P : 0), RS_MAX_PASSES);
  d.radii = p.radii; d.header = g.header; d.depth_key = g.depth_key; d.rect = g.rect;
  d.xy_ext = g.xy_ext; d.conic_opacity = g.conic_opacity; d.rgb_depth = g.rgb_depth; d.clamped = g.clamped;
  d.grad_acc = grad_acc;
  return d;
}

// `p` carries the Gaussians (shared by every view of the batch); cameras and per-view outputs come from `vb`.
int launch_preprocess(const tgr_params& p, const tgr_binding* bind, const ViewBatch& vb, cudaStream_t s) {
  const int blocks = (p.P + 255) / 256;
  if (blocks == 0 || vb.V <= 0) return 0;
  const bool staged = p.colors_precomp == nullptr && p.shs != nullptr && p.M == 16 && p.D >= 2 &&
                      (reinterpret_cast<uintptr_t>(p.shs) & 15) == 0;
  const size_t smem = staged ? (size_t)256 * SH_ROW_PAD * sizeof(float) : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(preprocess_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * SH_ROW_PAD * 4);
    cudaFuncSetAttribute(preprocess_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * SH_ROW_PAD * 4);
    attr_set = true;
  }
  tgr_binding none{};
  const tgr_binding& b = bind ? *bind : none;
  if (bind) {
    if (staged) preprocess_kernel<true, true><<<blocks, 256, smem, s>>>(p, b, vb);
    else preprocess_kernel<true, false><<<blocks, 256, 0, s>>>(p, b, vb);
  } else {
    if (staged) preprocess_kernel<false, true><<<blocks, 256, smem, s>>>(p, b, vb);
    else preprocess_kernel<false, false><<<blocks, 256, 0, s>>>(p, b, vb);
  }
  count_launch();
  return check_launch("preprocess", p.debug != 0, s);
}

// visibility mask (checkFrustum, rasterizer_impl.cu:54-66 -> in_frustum, auxiliary.h:139-164): !(view-space z <= 0.2)
__global__ void mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float3 po = {means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]};
  const float3 pv = xform4x3(po, view);
  present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

int launch_mark_visible(int32_t P, const float* means3D, const float* view, const float* /*proj*/, uint8_t* present,
                        cudaStream_t s) {
  if (P <= 0) return 0;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
  count_launch();
  return check_launch("mark_visible", false, s);
}

}  // namespace tgr
