// cameras.cu — device-side construction of the rasterizer's camera block for a batch of views.
//
// Reference: the preamble of render_image_gaussian_rasterizer, Edit_core/tetgs_scene/tetgs_model.py:479-503
// (cat, axis flip, torch.inverse, getWorld2View, getProjectionMatrix, bmm — ~25 tiny ATen launches, two .item()
// host syncs and an H2D copy per view) and Edit_core/utils/graphics_utils.py:39-49,68-86.  One thread per view,
// one launch per batch, nothing touches the host; the 160-byte records are what tgr_params' camera pointers read.
#include "common.cuh"

namespace tgr {

__global__ void build_cameras_kernel(int V, const float* __restrict__ c2w, const float* __restrict__ intr, float znear,
                                     float zfar, float* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* M = c2w + (size_t)v * 12;   // [3,4] row-major
  // tetgs_model.py:482-485: OpenGL/Blender (Y up, Z back) -> COLMAP (Y down, Z forward): columns 1, 2 negated
  float A[3][3], t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    A[r][0] = M[4 * r + 0];
    A[r][1] = -M[4 * r + 1];
    A[r][2] = -M[4 * r + 2];
    t[r] = M[4 * r + 3];
  }
  // tetgs_model.py:488 torch.inverse of the affine 4x4: [A t; 0 1]^-1 = [A^-1, -A^-1 t; 0 1] (cofactor inverse, so
  // a c2w with scale or skew is handled like the general inverse does)
  float Ci[3][3];
  Ci[0][0] = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  Ci[0][1] = A[0][2] * A[2][1] - A[0][1] * A[2][2];
  Ci[0][2] = A[0][1] * A[1][2] - A[0][2] * A[1][1];
  Ci[1][0] = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  Ci[1][1] = A[0][0] * A[2][2] - A[0][2] * A[2][0];
  Ci[1][2] = A[0][2] * A[1][0] - A[0][0] * A[1][2];
  Ci[2][0] = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  Ci[2][1] = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  Ci[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  const float det = A[0][0] * Ci[0][0] + A[0][1] * Ci[1][0] + A[0][2] * Ci[2][0];
  const float idet = 1.0f / det;
  float W2C[4][4];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) W2C[r][c] = Ci[r][c] * idet;
    W2C[r][3] = -(W2C[r][0] * t[0] + W2C[r][1] * t[1] + W2C[r][2] * t[2]);
  }
  W2C[3][0] = W2C[3][1] = W2C[3][2] = 0.f;
  W2C[3][3] = 1.f;

  // graphics_utils.py:68-86 in double like the Python floats it is written with, stored fp32 like torch.zeros(4,4)
  const double fovx = (double)intr[4 * v + 0], fovy = (double)intr[4 * v + 1];
  const double thx = tan(fovx * 0.5), thy = tan(fovy * 0.5);
  const double zn = (double)znear, zf = (double)zfar;
  const double top = thy * zn, right = thx * zn;
  float P[4][4] = {};
  P[0][0] = (float)(2.0 * zn / (right - (-right)));
  P[1][1] = (float)(2.0 * zn / (top - (-top)));
  P[0][2] = -intr[4 * v + 2];                 // tetgs_model.py:498-499 (proj_transform[2,0/1] of the transposed P)
  P[1][2] = -intr[4 * v + 3];
  P[3][2] = 1.0f;
  P[2][2] = (float)(zf / (zf - zn));
  P[2][3] = (float)(-(zf * zn) / (zf - zn));

  float* o = out + (size_t)v * TGR_CAMERA_FLOATS;
  // viewmatrix = W2C^T, flat row-major of the transpose == element (r,c) of W2C at [4c + r]
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) o[4 * c + r] = W2C[r][c];
  // projmatrix = W2C^T P^T (tetgs_model.py:501): element [i][j] = sum_k W2C[k][i] * P[j][k], summed in k order
  // like the bmm's fp32 dot product
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fmaf(W2C[k][i], P[j][k], s);
      o[16 + 4 * i + j] = s;
    }
  o[32] = t[0];   // camera centre = translation of c2w (tetgs_model.py:502 get_camera_center)
  o[33] = t[1];
  o[34] = t[2];
  o[35] = (float)thx;
  o[36] = (float)thy;
  o[37] = o[38] = o[39] = 0.f;
}

}  // namespace tgr

using namespace tgr;

extern "C" int tgr_build_cameras(int32_t n_views, const float* camera_to_worlds, const float* intrinsics, float znear,
                                 float zfar, float* out_cameras, void* stream) {
  if (n_views < 0) { set_error("build_cameras: n_views %d", n_views); return 1; }
  if (n_views == 0) return 0;
  if (!camera_to_worlds || !intrinsics || !out_cameras) { set_error("build_cameras: null pointer"); return 1; }
  if (!(zfar > znear) || !(znear > 0.f)) { set_error("build_cameras: need 0 < znear < zfar"); return 1; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  build_cameras_kernel<<<(n_views + 63) / 64, 64, 0, s>>>(n_views, camera_to_worlds, intrinsics, znear, zfar, out_cameras);
  count_launch();
  return check_launch("build_cameras", false, s);
}
