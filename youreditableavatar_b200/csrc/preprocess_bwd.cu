// preprocess_bwd.cu — per-Gaussian backward stage, ONE kernel for what the reference does in two
// (computeCov2DCUDA backward.cu:144-274 + preprocessCUDA backward.cu:346-396, SH backward :20-139,
// cov3D backward :278-341), plus the chain rule through the optional mesh binding / activations
// (tetgs_model.py:252-286).  Reads the packed 2D-gradient accumulator filled by blend_bwd and WRITES every
// element of every output gradient (zeros for culled Gaussians), so the caller never pre-zeroes them —
// the reference's glue issues nine torch::zeros fills per backward (rasterize_points.cu:151-159).
#include "common.cuh"

namespace tgr {

// acc = true: add into the destination (multi-view gradient accumulation), else overwrite
__device__ __forceinline__ void put(float* p, float v, bool acc) { *p = acc ? (*p + v) : v; }
__device__ __forceinline__ void store3(float* p, size_t i, float a, float b, float c, bool acc) {
  put(p + 3 * i + 0, a, acc); put(p + 3 * i + 1, b, acc); put(p + 3 * i + 2, c, acc);
}

// Coalesced flush of a per-block output tile (NC floats per Gaussian) staged in shared memory: 16-byte
// read-modify-write (acc) or store.  gdst + block_first*NC is 16-byte aligned because block_first % 256 == 0.
template <int NC>
__device__ __forceinline__ void flush_tile(float* __restrict__ gdst, const float* s_src, int block_first, int nrows,
                                           bool acc) {
  if (gdst == nullptr) return;
  const int total = nrows * NC;
  const int n4 = total >> 2;
  float4* d4 = reinterpret_cast<float4*>(gdst + (size_t)block_first * NC);
  const float4* s4 = reinterpret_cast<const float4*>(s_src);
  for (int v = threadIdx.x; v < n4; v += blockDim.x) {
    float4 o = s4[v];
    if (acc) { const float4 old = d4[v]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
    d4[v] = o;
  }
  float* dt = gdst + (size_t)block_first * NC;
  for (int e = (n4 << 2) + threadIdx.x; e < total; e += blockDim.x) dt[e] = acc ? (dt[e] + s_src[e]) : s_src[e];
}

constexpr int SH_ROW = 48;        // floats per SH row at M = 16
constexpr int SH_ROW_PAD = 49;    // padded shared-memory row stride: 49*t mod 32 is a bijection over a warp

// STAGED: the block's 256 SH rows (48 KB, contiguous in HBM) are moved global -> shared with fully coalesced
// 16-byte loads, each thread then walks its own padded row bank-conflict-free; dL/dsh goes back the same way.
// (One thread per Gaussian reading 48 floats at a 192-byte stride straight from global costs 32 sectors per
// request — ncu showed the unstaged kernel at ~30% of HBM bandwidth.)
// MULTI: the kernel serves a batch of views (ViewBatch, common.cuh).  Parameters are read once, every view's
// packed 2-D gradient row is chained back to 3-D and summed in registers / a second shared-memory SH tile, and
// the parameter gradients are written ONCE — per-view calls would read-modify-write 236 B per Gaussian and view.
// With MULTI = false (one view) the SH row is overwritten in place by its gradient, as before.
template <bool BOUND, bool STAGED, bool MULTI>
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(const tgr_params p, const tgr_binding bind,
                                                             const __grid_constant__ ViewBatch vb) {
  extern __shared__ float s_rows[];
  __shared__ float s_cam[MAX_BATCH][CAM_FLOATS];
  // the launch covers Gaussians [vb.first, vb.end): a caller may split a batch's backward into ranges so that the
  // all-reduce of one range's gradients overlaps the next range's kernel (parallel.py)
  const int block_first = vb.first + blockIdx.x * blockDim.x;
  const int idx = block_first + threadIdx.x;
  const int nrows = min((int)blockDim.x, vb.end - block_first);
  // frozen Gaussians (tgr_binding.n_frozen, the keep part of the edit models) take no part: zero gradient rows, no work
  const bool frozen = BOUND && idx < bind.n_frozen;
  const bool block_frozen = BOUND && block_first + (int)blockDim.x <= bind.n_frozen;
  const int V = frozen ? 0 : (MULTI ? vb.V : 1);
  for (int e = threadIdx.x; e < (MULTI ? vb.V : 1) * CAM_FLOATS; e += blockDim.x) {
    const int v = e / CAM_FLOATS, k = e % CAM_FLOATS;
    float x = 0.f;
    if (k < 16) x = vb.v[v].viewmatrix[k];
    else if (k < 32) x = vb.v[v].projmatrix[k - 16];
    else if (k < 35) x = vb.v[v].campos[k - 32];
    s_cam[v][k] = x;
  }
  if (STAGED && !block_frozen) {
    const float4* src = reinterpret_cast<const float4*>(p.shs + (size_t)block_first * SH_ROW);
    for (int v = threadIdx.x; v < nrows * (SH_ROW / 4); v += blockDim.x) {
      const float4 q = __ldg(src + v);
      const int e = v * 4;
      float* d = s_rows + (e / SH_ROW) * SH_ROW_PAD + (e % SH_ROW);
      d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    }
  }
  float* const row = s_rows + threadIdx.x * SH_ROW_PAD;
  // gradient tile of the SH rows: a second tile when several views are summed, else the SH row itself
  float* const s_out = (STAGED && MULTI) ? (s_rows + 256 * SH_ROW_PAD) : s_rows;
  float* const orow = s_out + threadIdx.x * SH_ROW_PAD;
  if (STAGED && MULTI) {
#pragma unroll
    for (int k = 0; k < SH_ROW; ++k) orow[k] = 0.f;
  }
  __syncthreads();

  float o_m2[2] = {0.f, 0.f}, o_col[3] = {0.f, 0.f, 0.f}, o_op = 0.f, o_mean[3] = {0.f, 0.f, 0.f};
  float o_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o_sc[3] = {0.f, 0.f, 0.f}, o_rot[4] = {0.f, 0.f, 0.f, 0.f};
  if (idx < vb.end) {
  const size_t i = (size_t)idx;
  const int M = p.M;
  const bool has_sh = (p.shs != nullptr && p.colors_precomp == nullptr && M > 0);
  const bool has_sr = BOUND || (p.scales != nullptr && p.rotations != nullptr && p.cov3D_precomp == nullptr);

  float dmean[3] = {0.f, 0.f, 0.f};
  float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dscale[3] = {0.f, 0.f, 0.f};
  float drot[4] = {0.f, 0.f, 0.f, 0.f};
  float dm2[2] = {0.f, 0.f}, dcol[3] = {0.f, 0.f, 0.f}, dop = 0.f;

  float3 mean = {0.f, 0.f, 0.f};
  float3 scale = {0.f, 0.f, 0.f};
  float4 rot = {1.f, 0.f, 0.f, 0.f};
  if (BOUND) {
    mean = {bind.out_means3D[3 * i], bind.out_means3D[3 * i + 1], bind.out_means3D[3 * i + 2]};
    scale = {bind.out_scales[3 * i], bind.out_scales[3 * i + 1], bind.out_scales[3 * i + 2]};
    rot = reinterpret_cast<const float4*>(bind.out_rotations)[i];
  } else {
    mean = {p.means3D[3 * i], p.means3D[3 * i + 1], p.means3D[3 * i + 2]};
    if (has_sr) {
      scale = {p.scales[3 * i], p.scales[3 * i + 1], p.scales[3 * i + 2]};
      rot = reinterpret_cast<const float4*>(p.rotations)[i];
    }
  }

  // ---- 3D covariance (recomputed rather than stored by the forward; view-independent) ------------
  float c3[6];
  const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
  M3 R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
            2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
            2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
  const float3 s = {p.scale_modifier * scale.x, p.scale_modifier * scale.y, p.scale_modifier * scale.z};
  M3 Mm;
  if (has_sr) {
    M3 S = m3(s.x, 0.f, 0.f, 0.f, s.y, 0.f, 0.f, 0.f, s.z);
    Mm = m3_mul(S, R);
    M3 Sg = m3_mul(m3_T(Mm), Mm);
    c3[0] = Sg.m[0][0]; c3[1] = Sg.m[0][1]; c3[2] = Sg.m[0][2]; c3[3] = Sg.m[1][1]; c3[4] = Sg.m[1][2]; c3[5] = Sg.m[2][2];
  } else {
    Mm = m3(0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 6; ++k) c3[k] = p.cov3D_precomp[6 * i + k];
  }
  float sh_local[STAGED ? 1 : 48];
  if (!STAGED && has_sh) {
    const int ncoef = (p.D + 1) * (p.D + 1);
    const float* base = p.shs + i * (size_t)M * 3;
#pragma unroll
    for (int k = 0; k < 48; ++k)
      if (k < ncoef * 3) sh_local[k] = __ldg(base + k);
  }
  bool any_visible = false;
  if (STAGED && !MULTI && frozen) {   // single view: the SH row doubles as its gradient row — a frozen Gaussian's is zero
#pragma unroll
    for (int k = 0; k < SH_ROW; ++k) orow[k] = 0.f;
  }

  for (int vi = 0; vi < V; ++vi) {
  const ViewDesc& vd = vb.v[vi];
  const int radius = vd.radii[idx];
  // non-staged SH path writes global memory directly: views after the first always add
  const bool acc_sh = p.accumulate != 0 || vi > 0;
  if (radius > 0) {
    any_visible = true;
    const float4* acc4 = reinterpret_cast<const float4*>(vd.grad_acc + i * GRAD_ACC);
    const float4 a0 = acc4[0], a1 = acc4[1], a2 = acc4[2];
    const float vm2[2] = {a0.x, a0.y};
    dm2[0] += a0.x; dm2[1] += a0.y;
    const float dconx = a0.z, dcony = a0.w, dconz = a1.x;
    dop += a1.y;
    const float vcol[3] = {a1.z, a1.w, a2.x};
    dcol[0] += a1.z; dcol[1] += a1.w; dcol[2] += a2.x;
    const float dz = a2.y;
    const float* view = s_cam[vi];
    const float* proj = s_cam[vi] + 16;

    // ---- conic -> cov2D -> cov3D and the covariance part of dL/dmean (backward.cu:144-274) --------
    {
      const float h_x = vd.W / (2.0f * vd.tan_fovx), h_y = vd.H / (2.0f * vd.tan_fovy);
      float3 t = xform4x3(mean, view);
      const float limx = 1.3f * vd.tan_fovx, limy = 1.3f * vd.tan_fovy;
      const float txtz = t.x / t.z, tytz = t.y / t.z;
      t.x = min(limx, max(-limx, txtz)) * t.z;
      t.y = min(limy, max(-limy, tytz)) * t.z;
      const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
      const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;

      M3 J = m3(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z), 0.f, 0.f, 0.f);
      M3 Wm = m3(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
      M3 Vrk = m3(c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]);
      M3 T = m3_mul(Wm, J);
      M3 cov2D = m3_mul(m3_mul(m3_T(T), m3_T(Vrk)), T);
      const float a = cov2D.m[0][0] + 0.3f, b = cov2D.m[0][1], c = cov2D.m[1][1] + 0.3f;
      const float denom = a * c - b * b;
      float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
      const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
      if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dconx + 2 * b * c * dcony + (denom - a * c) * dconz);
        dL_dc = denom2inv * (-a * a * dconz + 2 * a * b * dcony + (denom - a * c) * dconx);
        dL_db = denom2inv * 2 * (b * c * dconx - (denom + 2 * b * b) * dcony + a * b * dconz);
        const float (*Tm)[3] = T.m;
        dcov[0] += (Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc);
        dcov[3] += (Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc);
        dcov[5] += (Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc);
        dcov[1] += 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
        dcov[2] += 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
        dcov[4] += 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
      }
      // dL/dT (upper 2x3), then dL/dJ, then dL/dt
      float dT0[3], dT1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float A = T.m[0][0] * Vrk.m[k][0] + T.m[0][1] * Vrk.m[k][1] + T.m[0][2] * Vrk.m[k][2];
        const float B = T.m[1][0] * Vrk.m[k][0] + T.m[1][1] * Vrk.m[k][1] + T.m[1][2] * Vrk.m[k][2];
        dT0[k] = 2 * A * dL_da + B * dL_db;
        dT1[k] = 2 * B * dL_dc + A * dL_db;
      }
      const float dJ00 = Wm.m[0][0] * dT0[0] + Wm.m[0][1] * dT0[1] + Wm.m[0][2] * dT0[2];
      const float dJ02 = Wm.m[2][0] * dT0[0] + Wm.m[2][1] * dT0[1] + Wm.m[2][2] * dT0[2];
      const float dJ11 = Wm.m[1][0] * dT1[0] + Wm.m[1][1] * dT1[1] + Wm.m[1][2] * dT1[2];
      const float dJ12 = Wm.m[2][0] * dT1[0] + Wm.m[2][1] * dT1[1] + Wm.m[2][2] * dT1[2];
      const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
      const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
      const float dty = y_grad_mul * -h_y * tz2 * dJ12;
      const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t.x) * tz3 * dJ02 + (2 * h_y * t.y) * tz3 * dJ12;
      dmean[0] += view[0] * dtx + view[1] * dty + view[2] * dtz;
      dmean[1] += view[4] * dtx + view[5] * dty + view[6] * dtz;
      dmean[2] += view[8] * dtx + view[9] * dty + view[10] * dtz;
    }

    // ---- screen-space mean -> 3D mean through the projective divide (backward.cu:370-387) -------
    {
      const float4 m_hom = xform4x4(mean, proj);
      const float m_w = 1.0f / (m_hom.w + 0.0000001f);
      const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
      const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
      dmean[0] += (proj[0] * m_w - proj[3] * mul1) * vm2[0] + (proj[1] * m_w - proj[3] * mul2) * vm2[1];
      dmean[1] += (proj[4] * m_w - proj[7] * mul1) * vm2[0] + (proj[5] * m_w - proj[7] * mul2) * vm2[1];
      dmean[2] += (proj[8] * m_w - proj[11] * mul1) * vm2[0] + (proj[9] * m_w - proj[11] * mul2) * vm2[1];
      // extras: depth image gradient reaches the mean through row 2 of the view matrix
      dmean[0] += view[2] * dz; dmean[1] += view[6] * dz; dmean[2] += view[10] * dz;
    }

    // ---- colour -> SH coefficients and view direction (backward.cu:20-139) ------------------------
    if (has_sh) {
      const int D = p.D;
      const int ncoef = (D + 1) * (D + 1);
      const float* sh = STAGED ? row : sh_local;
      const float3 cam = {view[32], view[33], view[34]};
      const float3 dir_orig = {mean.x - cam.x, mean.y - cam.y, mean.z - cam.z};
      const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
      const float dx = dir_orig.x / len, dy = dir_orig.y / len, dzv = dir_orig.z / len;
      const uint8_t cl = vd.clamped[i];
      float dRGB[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) dRGB[c] = ((cl >> c) & 1) ? 0.f : vcol[c];

      float w[16];  // d(colour)/d(sh_k) basis weights
#pragma unroll
      for (int k = 0; k < 16; ++k) w[k] = 0.f;
      float dRGBdx[3] = {0.f, 0.f, 0.f}, dRGBdy[3] = {0.f, 0.f, 0.f}, dRGBdz[3] = {0.f, 0.f, 0.f};
      const float xx = dx * dx, yy = dy * dy, zz = dzv * dzv, xy_ = dx * dy, yz = dy * dzv, xz = dx * dzv;
      w[0] = SH_C0;
      if (D > 0) {
        w[1] = -SH_C1 * dy; w[2] = SH_C1 * dzv; w[3] = -SH_C1 * dx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          dRGBdx[c] = -SH_C1 * sh[9 + c];
          dRGBdy[c] = -SH_C1 * sh[3 + c];
          dRGBdz[c] = SH_C1 * sh[6 + c];
        }
        if (D > 1) {
          w[4] = SH_C2[0] * xy_; w[5] = SH_C2[1] * yz; w[6] = SH_C2[2] * (2.f * zz - xx - yy);
          w[7] = SH_C2[3] * xz; w[8] = SH_C2[4] * (xx - yy);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            dRGBdx[c] += SH_C2[0] * dy * sh[12 + c] + SH_C2[2] * 2.f * -dx * sh[18 + c] + SH_C2[3] * dzv * sh[21 + c] + SH_C2[4] * 2.f * dx * sh[24 + c];
            dRGBdy[c] += SH_C2[0] * dx * sh[12 + c] + SH_C2[1] * dzv * sh[15 + c] + SH_C2[2] * 2.f * -dy * sh[18 + c] + SH_C2[4] * 2.f * -dy * sh[24 + c];
            dRGBdz[c] += SH_C2[1] * dy * sh[15 + c] + SH_C2[2] * 2.f * 2.f * dzv * sh[18 + c] + SH_C2[3] * dx * sh[21 + c];
          }
          if (D > 2) {
            w[9] = SH_C3[0] * dy * (3.f * xx - yy); w[10] = SH_C3[1] * xy_ * dzv;
            w[11] = SH_C3[2] * dy * (4.f * zz - xx - yy); w[12] = SH_C3[3] * dzv * (2.f * zz - 3.f * xx - 3.f * yy);
            w[13] = SH_C3[4] * dx * (4.f * zz - xx - yy); w[14] = SH_C3[5] * dzv * (xx - yy);
            w[15] = SH_C3[6] * dx * (xx - 3.f * yy);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              dRGBdx[c] += (SH_C3[0] * sh[27 + c] * 3.f * 2.f * xy_ + SH_C3[1] * sh[30 + c] * yz + SH_C3[2] * sh[33 + c] * -2.f * xy_ +
                            SH_C3[3] * sh[36 + c] * -3.f * 2.f * xz + SH_C3[4] * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) +
                            SH_C3[5] * sh[42 + c] * 2.f * xz + SH_C3[6] * sh[45 + c] * 3.f * (xx - yy));
              dRGBdy[c] += (SH_C3[0] * sh[27 + c] * 3.f * (xx - yy) + SH_C3[1] * sh[30 + c] * xz +
                            SH_C3[2] * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * sh[36 + c] * -3.f * 2.f * yz +
                            SH_C3[4] * sh[39 + c] * -2.f * xy_ + SH_C3[5] * sh[42 + c] * -2.f * yz + SH_C3[6] * sh[45 + c] * -3.f * 2.f * xy_);
              dRGBdz[c] += (SH_C3[1] * sh[30 + c] * xy_ + SH_C3[2] * sh[33 + c] * 4.f * 2.f * yz +
                            SH_C3[3] * sh[36 + c] * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * sh[39 + c] * 4.f * 2.f * xz +
                            SH_C3[5] * sh[42 + c] * (xx - yy));
            }
          }
        }
      }
      // view-direction term first (with one view the staged SH row is about to be overwritten by its gradient)
      const float ddir[3] = {dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2],
                             dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2],
                             dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2]};
      // dL/dsh rows; rows beyond the active degree are zero
      if (STAGED) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float wk = (k < ncoef) ? w[k] : 0.f;
          if (MULTI) {
            orow[3 * k + 0] += wk * dRGB[0]; orow[3 * k + 1] += wk * dRGB[1]; orow[3 * k + 2] += wk * dRGB[2];
          } else {
            orow[3 * k + 0] = wk * dRGB[0]; orow[3 * k + 1] = wk * dRGB[1]; orow[3 * k + 2] = wk * dRGB[2];
          }
        }
      } else {
        float* out = p.dL_dsh + i * (size_t)M * 3;
        for (int k = 0; k < M; ++k) {
          const float wk = (k < ncoef && k < 16) ? w[k] : 0.f;
          put(out + 3 * k + 0, wk * dRGB[0], acc_sh); put(out + 3 * k + 1, wk * dRGB[1], acc_sh); put(out + 3 * k + 2, wk * dRGB[2], acc_sh);
        }
      }
      // ... into dL/dmean through the normalisation Jacobian (auxiliary.h:107-117)
      const float3 v = dir_orig;
      const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
      const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmean[0] += ((+sum2 - v.x * v.x) * ddir[0] - v.y * v.x * ddir[1] - v.z * v.x * ddir[2]) * invsum32;
      dmean[1] += (-v.x * v.y * ddir[0] + (sum2 - v.y * v.y) * ddir[1] - v.z * v.y * ddir[2]) * invsum32;
      dmean[2] += (-v.x * v.z * ddir[0] - v.y * v.z * ddir[1] + (sum2 - v.z * v.z) * ddir[2]) * invsum32;
    }
  } else if (has_sh) {
    // Gaussian culled in this view: its SH gradient rows are zero
    if (STAGED) {
      if (!MULTI) {
#pragma unroll
        for (int k = 0; k < SH_ROW; ++k) orow[k] = 0.f;
      }
    } else {
      float* out = p.dL_dsh + i * (size_t)M * 3;
      if (!acc_sh) for (int k = 0; k < M * 3; ++k) out[k] = 0.f;
    }
  }
  }  // views

  // ---- 3D covariance -> scale / quaternion (backward.cu:278-341); linear in dcov, so once per batch ----
  if (has_sr && any_visible) {
    M3 dSigma = m3(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                   0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
    M3 M2;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) M2.m[c][rr] = 2.0f * Mm.m[c][rr];
    M3 dM = m3_mul(M2, dSigma);
    M3 Rt = m3_T(R);
    M3 dMt = m3_T(dM);
    dscale[0] = Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2];
    dscale[1] = Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2];
    dscale[2] = Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) { dMt.m[0][k] *= s.x; dMt.m[1][k] *= s.y; dMt.m[2][k] *= s.z; }
    const float (*D)[3] = dMt.m;
    drot[0] = 2 * z * (D[0][1] - D[1][0]) + 2 * y * (D[2][0] - D[0][2]) + 2 * x * (D[1][2] - D[2][1]);
    drot[1] = 2 * y * (D[1][0] + D[0][1]) + 2 * z * (D[2][0] + D[0][2]) + 2 * r * (D[1][2] - D[2][1]) - 4 * x * (D[2][2] + D[1][1]);
    drot[2] = 2 * x * (D[1][0] + D[0][1]) + 2 * r * (D[2][0] - D[0][2]) + 2 * z * (D[1][2] + D[2][1]) - 4 * y * (D[2][2] + D[0][0]);
    drot[3] = 2 * r * (D[0][1] - D[1][0]) + 2 * x * (D[2][0] + D[0][2]) + 2 * y * (D[1][2] + D[2][1]) - 4 * z * (D[1][1] + D[0][0]);
  }

  // ---- the P-sized outputs leave through shared memory (see the epilogue) ------------------------------
  o_m2[0] = dm2[0]; o_m2[1] = dm2[1];
  o_col[0] = dcol[0]; o_col[1] = dcol[1]; o_col[2] = dcol[2];
  o_op = dop;
  o_mean[0] = dmean[0]; o_mean[1] = dmean[1]; o_mean[2] = dmean[2];
#pragma unroll
  for (int k = 0; k < 6; ++k) o_cov[k] = dcov[k];
  o_sc[0] = dscale[0]; o_sc[1] = dscale[1]; o_sc[2] = dscale[2];
  o_rot[0] = drot[0]; o_rot[1] = drot[1]; o_rot[2] = drot[2]; o_rot[3] = drot[3];
  const bool acc = p.accumulate != 0;

  if (BOUND) {
    // chain rule through points = ori + n*delta, scale = exp(.), q = normalize(.), o = sigmoid(.)
    const bool direct = bind.origins != nullptr;
    int i0 = 0, i1 = 0, i2 = 0;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    float n[3];
    if (direct) {
#pragma unroll
      for (int c = 0; c < 3; ++c) n[c] = bind.normals ? bind.normals[3 * i + c] : 0.f;
    } else {
      const int f = bind.face_index[idx];
      i0 = bind.faces[3 * f + 0]; i1 = bind.faces[3 * f + 1]; i2 = bind.faces[3 * f + 2];
      w0 = bind.bary[3 * i + 0]; w1 = bind.bary[3 * i + 1]; w2 = bind.bary[3 * i + 2];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        n[c] = w0 * bind.vert_normals[3 * i0 + c] + w1 * bind.vert_normals[3 * i1 + c] + w2 * bind.vert_normals[3 * i2 + c];
    }
    if (bind.dL_ddelta) put(bind.dL_ddelta + i, n[0] * dmean[0] + n[1] * dmean[1] + n[2] * dmean[2], acc);
    if (bind.dL_dlog_scales) store3(bind.dL_dlog_scales, i, dscale[0] * scale.x, dscale[1] * scale.y, dscale[2] * scale.z, acc);
    if (bind.dL_draw_quats) {
      const float4 q = reinterpret_cast<const float4*>(bind.raw_quats)[i];
      const float qn = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
      const float dotp = rot.x * drot[0] + rot.y * drot[1] + rot.z * drot[2] + rot.w * drot[3];
      float4* o = reinterpret_cast<float4*>(bind.dL_draw_quats) + i;
      float4 v = make_float4((drot[0] - rot.x * dotp) / qn, (drot[1] - rot.y * dotp) / qn, (drot[2] - rot.z * dotp) / qn,
                             (drot[3] - rot.w * dotp) / qn);
      if (acc) { const float4 old = *o; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
      *o = v;
    }
    if (bind.dL_dopacity_logits) {
      const float o = bind.out_opacities[i];
      put(bind.dL_dopacity_logits + i, dop * o * (1.f - o), acc);
    }
    if (bind.dL_dverts && any_visible && !direct) {
      const int vi[3] = {i0, i1, i2};
      const float wv[3] = {w0, w1, w2};
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(&bind.dL_dverts[3 * vi[k] + c], wv[k] * dmean[c]);
    }
  }
  }  // idx < P
  if (STAGED) {
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(p.dL_dsh + (size_t)block_first * SH_ROW);
    for (int v = threadIdx.x; v < nrows * (SH_ROW / 4); v += blockDim.x) {
      const int e = v * 4;
      const float* d = s_out + (e / SH_ROW) * SH_ROW_PAD + (e % SH_ROW);
      float4 o = make_float4(d[0], d[1], d[2], d[3]);
      if (p.accumulate) { const float4 old = dst[v]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
      dst[v] = o;
    }
    __syncthreads();  // SH rows are out; the buffer is reused for the small outputs below
  }
  {
    // stage [means3D 3 | scales 3 | rot 4 | opacity 1 | means2D 3 | colors 3 | cov3D 6] x 256 and flush coalesced
    float* t_mean = s_rows;
    float* t_sc = t_mean + 256 * 3;
    float* t_rot = t_sc + 256 * 3;
    float* t_op = t_rot + 256 * 4;
    float* t_m2 = t_op + 256;
    float* t_col = t_m2 + 256 * 3;
    float* t_cov = t_col + 256 * 3;
    const int t = threadIdx.x;
    t_mean[3 * t] = o_mean[0]; t_mean[3 * t + 1] = o_mean[1]; t_mean[3 * t + 2] = o_mean[2];
    t_sc[3 * t] = o_sc[0]; t_sc[3 * t + 1] = o_sc[1]; t_sc[3 * t + 2] = o_sc[2];
    reinterpret_cast<float4*>(t_rot)[t] = make_float4(o_rot[0], o_rot[1], o_rot[2], o_rot[3]);
    t_op[t] = o_op;
    t_m2[3 * t] = o_m2[0]; t_m2[3 * t + 1] = o_m2[1]; t_m2[3 * t + 2] = 0.f;
    t_col[3 * t] = o_col[0]; t_col[3 * t + 1] = o_col[1]; t_col[3 * t + 2] = o_col[2];
#pragma unroll
    for (int k = 0; k < 6; ++k) t_cov[6 * t + k] = o_cov[k];
    __syncthreads();
    const bool accf = p.accumulate != 0;
    flush_tile<3>(p.dL_dmeans3D, t_mean, block_first, nrows, accf);
    flush_tile<3>(p.dL_dscales, t_sc, block_first, nrows, accf);
    flush_tile<4>(p.dL_drotations, t_rot, block_first, nrows, accf);
    flush_tile<1>(p.dL_dopacity, t_op, block_first, nrows, accf);
    flush_tile<3>(p.dL_dmeans2D, t_m2, block_first, nrows, accf);
    flush_tile<3>(p.dL_dcolors, t_col, block_first, nrows, accf);
    flush_tile<6>(p.dL_dcov3D, t_cov, block_first, nrows, accf);
  }
}

// `p` carries the Gaussians and the output gradient tensors (shared by the batch); cameras, radii, clamp flags
// and the packed 2-D gradient rows of every view come from `vb`.
int launch_preprocess_bwd(const tgr_params& p, const tgr_binding* bind, const ViewBatch& vb, cudaStream_t s) {
  const int blocks = (vb.end - vb.first + 255) / 256;
  if (blocks <= 0 || vb.V <= 0) return 0;
  const bool has_sh = (p.shs != nullptr && p.colors_precomp == nullptr && p.M > 0);
  const bool staged = has_sh && p.M == 16 && p.D >= 2 && p.dL_dsh != nullptr &&
                      (reinterpret_cast<uintptr_t>(p.shs) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.dL_dsh) & 15) == 0;
  const bool multi = vb.V > 1;
  const size_t tile = (size_t)256 * SH_ROW_PAD * sizeof(float);
  const size_t smem = staged ? (multi ? 2 * tile : tile) : (size_t)256 * 23 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(preprocess_bwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile);
    cudaFuncSetAttribute(preprocess_bwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile);
    cudaFuncSetAttribute(preprocess_bwd_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * tile));
    cudaFuncSetAttribute(preprocess_bwd_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * tile));
    attr_set = true;
  }
  tgr_binding none{};
  const tgr_binding& b = bind ? *bind : none;
#define TGR_PB_LAUNCH(B, S, M) preprocess_bwd_kernel<B, S, M><<<blocks, 256, smem, s>>>(p, b, vb)
  if (bind) {
    if (staged) { if (multi) TGR_PB_LAUNCH(true, true, true); else TGR_PB_LAUNCH(true, true, false); }
    else        { if (multi) TGR_PB_LAUNCH(true, false, true); else TGR_PB_LAUNCH(true, false, false); }
  } else {
    if (staged) { if (multi) TGR_PB_LAUNCH(false, true, true); else TGR_PB_LAUNCH(false, true, false); }
    else        { if (multi) TGR_PB_LAUNCH(false, false, true); else TGR_PB_LAUNCH(false, false, false); }
  }
#undef TGR_PB_LAUNCH
  count_launch();
  return check_launch("preprocess_bwd", p.debug != 0, s);
}

}  // namespace tgr
