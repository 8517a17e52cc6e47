// loss.cu — image loss of the texture / refinement stages, forward + gradient.
//
// Reference: Edit_core/utils/loss_utils.py:17-63 (l1_loss, l2_loss, ssim/_ssim) and the closure
// `(1 - dssim_factor) * l1_loss + dssim_factor * (1 - ssim)` at Edit_core/tetgs_texture/refine.py:245-247.
// There the step right after the rasterizer is five grouped 11x11 conv2d launches plus ~20 elementwise ATen
// kernels on [1,3,H,W] images, and autograd replays as many in the backward.  Here it is two launches for a
// whole batch of views:
//
//   loss_fwd_kernel  one CTA per 32x32 tile of one channel of one view: pred / target tile + 5-pixel halo staged
//                    in shared memory (zero padding = out-of-image loads return 0, loss_utils.py:45), separable
//                    11-tap Gaussian for the five moments (mu1, mu2, E[x^2], E[y^2], E[xy]), ssim map, loss terms
//                    block-reduced to one partial per CTA (fixed order: deterministic), and the three maps the
//                    gradient needs  d ssim/d mu1, d ssim/d E[x^2], d ssim/d E[xy]  pre-scaled by d loss_v/d ssim.
//   loss_bwd_kernel  the window is symmetric and the padding is zero, so the adjoint of the convolution is the
//                    same convolution: d loss_v/dx = conv(A) + 2 x conv(B) + y conv(C) + the L1 / L2 terms,
//                    times the upstream dL/d loss_v of the view (a device array, so autograd needs no extra pass).
//   loss_finish_kernel  per-view sums of the CTA partials in double, total = sum_v w_v loss_v.
//
// Bound: HBM for the two streaming kernels — algorithmic bytes per pixel and channel: forward 4 (pred) + 4 | 1
// (target fp32 | u8) + 12 (maps), backward 12 + 4 + 4 | 1 + 4 (gradient) = 44 | 38 B.
#include <cmath>
#include "common.cuh"

namespace tgr {

constexpr int LT = 32;             // output tile side
constexpr int LR = 5;              // window radius: window_size 11 (loss_utils.py:32)
constexpr int LW = 2 * LR + 1;     // 11
constexpr int LH = LT + 2 * LR;    // 42: tile + halo
constexpr int LTHREADS = 256;
constexpr int LROWS = LT / (LTHREADS / LT);   // 4 output rows per thread in the vertical pass

__constant__ float c_win[LW];
__constant__ float2 c_win2[LW];    // {w, w}: the window for the packed FMAs

// Packed fp32 arithmetic (sm_100a FFMA2 / FMUL2): one issue slot for two IEEE-rounded fp32 FMAs.  The separable
// passes are chains of FMAs that share the tap weight between two planes (mu1 | mu2, E[x^2] | E[y^2], ...), and
// both loss kernels are issue-bound (ncu: 82 % / 70 % issue-active, FMA pipe 52 % / 32 %), so pairing planes cuts the
// instruction count without changing a single rounding.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 upk(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 win2(int t) { return pk(c_win2[t].x, c_win2[t].y); }

struct LossArgs {
  const float* pred;
  const void* target;
  const float* view_w;   // [V] or nullptr (1/V)
  float* maps;           // [3][V*3*H*W]
  float* partial;        // [V*3*tiles]
  float* grad;           // [V*3*H*W]
  int W, H, V;
  int vec;               // W % 4 == 0 and 16-byte aligned images: tiles are staged with 16-byte loads
  float w_l1, w_l2, w_ss;
};

// u / 255 correctly rounded (general_utils.py:8 divides the 8-bit image by 255.0 on the CPU) without the IEEE
// division sequence: q = RN(u r), r = RN(1/255), one Newton step on the exact residual.  Equal to (float)u / 255.0f
// for all 256 inputs (checked exhaustively; tests compare the u8 and fp32 target paths bit for bit).
__device__ __forceinline__ float u8_unit(uint32_t u) {
  const float x = (float)u, r = 1.0f / 255.0f;
  const float q = x * r;
  return fmaf(fmaf(-q, 255.0f, x), r, q);
}

template <bool U8>
__device__ __forceinline__ float load_target(const void* t, size_t i) {
  if (U8) return u8_unit(static_cast<const uint8_t*>(t)[i]);
  return static_cast<const float*>(t)[i];
}

// Stages rows [y0 - 5, y0 + 37) x columns [x0 - 5, x0 + 37) of up to three fp32 planes (or a u8 plane) into shared
// memory with 16-byte global loads: the 42 columns are covered by the 12 ALIGNED float4 starting at x0 - 8, so a
// CTA issues 504 vector loads per plane instead of 1764 scalar ones; W % 4 == 0 makes every vector wholly inside
// or wholly outside the image (outside = the zero padding of the convolution, loss_utils.py:45).
constexpr int LV4 = 12;
__device__ __forceinline__ void put4(float* row, int c, const float4 v) {
  if (c >= 0) row[c] = v.x;
  if (c + 1 >= 0) row[c + 1] = v.y;
  if (c + 2 < LH) row[c + 2] = v.z;
  if (c + 3 < LH) row[c + 3] = v.w;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < LTHREADS / 32; ++i) s += red[i];
  }
  return s;   // valid in thread 0
}

// Shared-memory geometry: the staged tile rows have pitch 43 and the row-filtered planes pitch 33 so that the
// register-blocked horizontal pass (a thread owns 8 consecutive outputs of one row: 18 loads instead of 88, the
// products p*p, g*g, p*g formed once per staged pixel instead of once per tap) is bank-conflict free both ways:
// lane t reads row t/4, columns 8(t%4)+k -> bank (11 r + 8 j + k) mod 32, all 32 distinct; it writes bank
// (r + 8 j + i) mod 32, all distinct.
constexpr int LPP = LH + 1;        // 43: pitch of the staged tiles
constexpr int LHP = LT + 1;        // 33: pitch of the horizontally filtered planes
constexpr int LXB = 8;             // outputs per thread in the horizontal pass
constexpr int LTASKS = LH * (LT / LXB);   // 168 (row, 8-column group) tasks per plane

template <bool U8, bool SSIM>
__global__ void __launch_bounds__(LTHREADS) loss_fwd_kernel(const LossArgs a) {
  __shared__ float sP[LH][LPP];
  __shared__ float sG[LH][LPP];
  __shared__ float2 sH01[LH][LHP];   // row-filtered {mu1, mu2}
  __shared__ float2 sH23[LH][LHP];   // row-filtered {E[x^2], E[y^2]}
  __shared__ float sH4[LH][LHP];     // row-filtered E[xy]
  __shared__ float red[LTHREADS / 32];

  const int plane = blockIdx.z;                 // v * 3 + c
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t base = (size_t)plane * a.H * a.W;
  const int tid = threadIdx.x;
  const int tx = tid & (LT - 1), ty = tid / LT;  // 32 x 8

  const float inv_n = 1.0f / (3.0f * (float)a.H * (float)a.W);

  if (SSIM) {
    if (a.vec) {
      for (int t = tid; t < LH * LV4; t += LTHREADS) {
        const int r = t / LV4, j = t - r * LV4;
        const int gy = y0 + r - LR, gx = x0 - 8 + 4 * j;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f), g = p;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const size_t o = base + (size_t)gy * a.W + gx;
          p = *reinterpret_cast<const float4*>(a.pred + o);
          if (U8) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.target) + o);
            g = make_float4(u8_unit(w & 0xffu), u8_unit((w >> 8) & 0xffu), u8_unit((w >> 16) & 0xffu), u8_unit(w >> 24));
          } else {
            g = *reinterpret_cast<const float4*>(static_cast<const float*>(a.target) + o);
          }
        }
        put4(sP[r], 4 * j - 3, p);
        put4(sG[r], 4 * j - 3, g);
      }
    } else {
      for (int r = ty; r < LH; r += LTHREADS / LT) {
        const int gy = y0 + r - LR;
        const bool row_ok = gy >= 0 && gy < a.H;
        const size_t ro = base + (size_t)(row_ok ? gy : 0) * a.W;
#pragma unroll
        for (int c = tx; c < LH; c += LT) {
          const int gx = x0 + c - LR;
          float p = 0.f, g = 0.f;
          if (row_ok && gx >= 0 && gx < a.W) {
            p = a.pred[ro + gx];
            g = load_target<U8>(a.target, ro + gx);
          }
          sP[r][c] = p;
          sG[r][c] = g;
        }
      }
    }
    __syncthreads();
    // horizontal pass: 42 rows x 4 groups of 8 columns
    if (tid < LTASKS) {
      const int r = tid >> 2, c0 = (tid & 3) * LXB;
      f32x2 m12[LXB], e1122[LXB];
      float e12[LXB];
#pragma unroll
      for (int i = 0; i < LXB; ++i) { m12[i] = 0ull; e1122[i] = 0ull; e12[i] = 0.f; }
#pragma unroll
      for (int k = 0; k < LXB + LW - 1; ++k) {
        const float p = sP[r][c0 + k], g = sG[r][c0 + k];
        const f32x2 x = pk(p, g);
        const f32x2 xx = mul2(x, x);
        const float pg = p * g;
#pragma unroll
        for (int i = 0; i < LXB; ++i) {
          const int t = k - i;                   // tap index of staged column k for output i
          if (t >= 0 && t < LW) {
            const f32x2 w = win2(t);
            m12[i] = fma2(w, x, m12[i]);
            e1122[i] = fma2(w, xx, e1122[i]);
            e12[i] = fmaf(c_win[t], pg, e12[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < LXB; ++i) {
        sH01[r][c0 + i] = upk(m12[i]);
        sH23[r][c0 + i] = upk(e1122[i]);
        sH4[r][c0 + i] = e12[i];
      }
    }
    __syncthreads();
  }

  // vertical pass + map: thread owns column tx, rows ty*4 .. ty*4+3
  float acc = 0.f;
  float q[5][LROWS];
  if (SSIM) {
    {
      f32x2 a01[LROWS], a23[LROWS];
#pragma unroll
      for (int j = 0; j < LROWS; ++j) { a01[j] = 0ull; a23[j] = 0ull; q[4][j] = 0.f; }
#pragma unroll
      for (int k = 0; k < LROWS + LW - 1; ++k) {
        const float2 v01 = sH01[ty * LROWS + k][tx], v23 = sH23[ty * LROWS + k][tx];
        const float v4 = sH4[ty * LROWS + k][tx];
        const f32x2 c01 = pk(v01.x, v01.y), c23 = pk(v23.x, v23.y);
#pragma unroll
        for (int j = 0; j < LROWS; ++j) {
          const int t = k - j;
          if (t >= 0 && t < LW) {
            const f32x2 w = win2(t);
            a01[j] = fma2(w, c01, a01[j]);
            a23[j] = fma2(w, c23, a23[j]);
            q[4][j] = fmaf(c_win[t], v4, q[4][j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < LROWS; ++j) {
        const float2 u01 = upk(a01[j]), u23 = upk(a23[j]);
        q[0][j] = u01.x; q[1][j] = u01.y; q[2][j] = u23.x; q[3][j] = u23.y;
      }
    }
  }
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;   // loss_utils.py:54-55
  const float gs = -a.w_ss * inv_n;                     // d loss_v / d ssim_map(pixel)
  const size_t plane_sz = (size_t)a.V * 3 * a.H * a.W;
#pragma unroll
  for (int j = 0; j < LROWS; ++j) {
    const int gy = y0 + ty * LROWS + j, gx = x0 + tx;
    if (gy < a.H && gx < a.W) {
      const size_t o = base + (size_t)gy * a.W + gx;
      float p, g;
      if (SSIM) {
        p = sP[ty * LROWS + j + LR][tx + LR];
        g = sG[ty * LROWS + j + LR][tx + LR];
      } else {
        p = a.pred[o];
        g = load_target<U8>(a.target, o);
      }
      const float d = p - g;
      float term = a.w_l1 * fabsf(d) + a.w_l2 * d * d;
      if (SSIM) {
        const float mu1 = q[0][j], mu2 = q[1][j];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = q[2][j] - mu1_sq, s2 = q[3][j] - mu2_sq, s12 = q[4][j] - mu12;   // loss_utils.py:50-52
        const float A = mu1_sq + mu2_sq + C1, B = s1 + s2 + C2;
        const float Cn = 2.f * mu12 + C1, D = 2.f * s12 + C2;
        // A >= C1 and B ~ C2 + variances > 0: hardware reciprocals (1 ulp) instead of four IEEE divisions
        const float iA = __fdividef(1.f, A), iB = __fdividef(1.f, B);
        const float iAB = iA * iB;
        const float CD = Cn * D;
        const float ssim = CD * iAB;                                                      // loss_utils.py:57
        term += a.w_ss * (1.f - ssim);
        // derivatives of the map wrt the three window moments that depend on pred (mu1, E[x^2], E[xy]);
        // sigma1_sq = E[x^2] - mu1^2 and sigma12 = E[xy] - mu1 mu2 folded into d/d mu1
        const float d_s1 = -ssim * iB;
        const float d_s12 = 2.f * Cn * iAB;
        const float d_mu1 = 2.f * mu2 * D * iAB - 2.f * mu1 * ssim * iA - 2.f * mu1 * d_s1 - mu2 * d_s12;
        a.maps[o] = gs * d_mu1;
        a.maps[plane_sz + o] = gs * d_s1;
        a.maps[2 * plane_sz + o] = gs * d_s12;
      }
      acc += term;
    }
  }
  const float s = block_sum(acc, red);
  if (tid == 0) a.partial[((size_t)plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s * inv_n;
}

template <bool U8, bool SSIM>
__global__ void __launch_bounds__(LTHREADS) loss_bwd_kernel(const LossArgs a) {
  __shared__ float sM[3][LH][LPP];
  __shared__ float2 sH01[LH][LHP];   // row-filtered maps 0 | 1
  __shared__ float sH2[LH][LHP];     // row-filtered map 2

  const int plane = blockIdx.z;
  const int v = plane / 3;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t base = (size_t)plane * a.H * a.W;
  const size_t plane_sz = (size_t)a.V * 3 * a.H * a.W;
  const int tid = threadIdx.x;
  const int tx = tid & (LT - 1), ty = tid / LT;
  const float wv = a.view_w ? a.view_w[v] : 1.0f / (float)a.V;   // here: dL / d loss_v
  const float inv_n = 1.0f / (3.0f * (float)a.H * (float)a.W);

  float q[3][LROWS];
  if (SSIM) {
    if (a.vec) {
      for (int t = tid; t < LH * LV4; t += LTHREADS) {
        const int r = t / LV4, j = t - r * LV4;
        const int gy = y0 + r - LR, gx = x0 - 8 + 4 * j;
        float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0, m2 = m0;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const size_t o = base + (size_t)gy * a.W + gx;
          m0 = *reinterpret_cast<const float4*>(a.maps + o);
          m1 = *reinterpret_cast<const float4*>(a.maps + plane_sz + o);
          m2 = *reinterpret_cast<const float4*>(a.maps + 2 * plane_sz + o);
        }
        put4(sM[0][r], 4 * j - 3, m0);
        put4(sM[1][r], 4 * j - 3, m1);
        put4(sM[2][r], 4 * j - 3, m2);
      }
    } else {
      for (int r = ty; r < LH; r += LTHREADS / LT) {
        const int gy = y0 + r - LR;
        const bool row_ok = gy >= 0 && gy < a.H;
        const size_t ro = base + (size_t)(row_ok ? gy : 0) * a.W;
#pragma unroll
        for (int c = tx; c < LH; c += LT) {
          const int gx = x0 + c - LR;
          float m0 = 0.f, m1 = 0.f, m2 = 0.f;
          if (row_ok && gx >= 0 && gx < a.W) {
            m0 = a.maps[ro + gx];
            m1 = a.maps[plane_sz + ro + gx];
            m2 = a.maps[2 * plane_sz + ro + gx];
          }
          sM[0][r][c] = m0;
          sM[1][r][c] = m1;
          sM[2][r][c] = m2;
        }
      }
    }
    __syncthreads();
    if (tid < LTASKS) {
      const int r = tid >> 2, c0 = (tid & 3) * LXB;
      f32x2 h01[LXB];
      float h2[LXB];
#pragma unroll
      for (int i = 0; i < LXB; ++i) { h01[i] = 0ull; h2[i] = 0.f; }
#pragma unroll
      for (int k = 0; k < LXB + LW - 1; ++k) {
        const f32x2 x01 = pk(sM[0][r][c0 + k], sM[1][r][c0 + k]);
        const float x2 = sM[2][r][c0 + k];
#pragma unroll
        for (int i = 0; i < LXB; ++i) {
          const int t = k - i;
          if (t >= 0 && t < LW) {
            h01[i] = fma2(win2(t), x01, h01[i]);
            h2[i] = fmaf(c_win[t], x2, h2[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < LXB; ++i) {
        sH01[r][c0 + i] = upk(h01[i]);
        sH2[r][c0 + i] = h2[i];
      }
    }
    __syncthreads();
    {
      f32x2 a01[LROWS];
#pragma unroll
      for (int j = 0; j < LROWS; ++j) { a01[j] = 0ull; q[2][j] = 0.f; }
#pragma unroll
      for (int k = 0; k < LROWS + LW - 1; ++k) {
        const float2 v01 = sH01[ty * LROWS + k][tx];
        const float v2 = sH2[ty * LROWS + k][tx];
        const f32x2 c01 = pk(v01.x, v01.y);
#pragma unroll
        for (int j = 0; j < LROWS; ++j) {
          const int t = k - j;
          if (t >= 0 && t < LW) {
            a01[j] = fma2(win2(t), c01, a01[j]);
            q[2][j] = fmaf(c_win[t], v2, q[2][j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < LROWS; ++j) {
        const float2 u = upk(a01[j]);
        q[0][j] = u.x; q[1][j] = u.y;
      }
    }
  }
  const float k1 = a.w_l1 * wv * inv_n, k2 = 2.f * a.w_l2 * wv * inv_n;
#pragma unroll
  for (int j = 0; j < LROWS; ++j) {
    const int gy = y0 + ty * LROWS + j, gx = x0 + tx;
    if (gy < a.H && gx < a.W) {
      const size_t o = base + (size_t)gy * a.W + gx;
      const float p = a.pred[o];
      const float g = load_target<U8>(a.target, o);
      const float d = p - g;
      float gr = k1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) + k2 * d;   // torch.abs backward: sign(0) = 0
      if (SSIM) gr += wv * (q[0][j] + 2.f * p * q[1][j] + g * q[2][j]);
      a.grad[o] = gr;
    }
  }
}

// loss_out[1 + v] = sum of view v's partials (double, fixed order: one warp per view, lane-strided then a shuffle
// tree); loss_out[0] = sum_v w_v loss_v in view order.  Deterministic; a few microseconds for 8 x 3072 partials.
__global__ void __launch_bounds__(1024) loss_finish_kernel(const float* __restrict__ partial, int per_view, int V,
                                                           const float* __restrict__ view_w, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int v = warp; v < V; v += nwarps) {
    const float* __restrict__ pv = partial + (size_t)v * per_view;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;     // four independent chains hide the load latency
    int i = lane;
    for (; i + 96 < per_view; i += 128) {
      s0 += (double)pv[i]; s1 += (double)pv[i + 32]; s2 += (double)pv[i + 64]; s3 += (double)pv[i + 96];
    }
    for (; i < per_view; i += 32) s0 += (double)pv[i];
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[1 + v] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int v = 0; v < V; ++v) total += (double)out[1 + v] * (double)(view_w ? view_w[v] : 1.0f / (float)V);
    out[0] = (float)total;
  }
}

static void upload_window() {
  // loss_utils.py:23-25: float32 tensor of exp(-(x - 5)^2 / (2 sigma^2)), divided by its float32 sum
  static thread_local int done_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_dev == dev) return;
  float g[LW], s = 0.f;
  for (int x = 0; x < LW; ++x) {
    g[x] = (float)std::exp(-(double)((x - LW / 2) * (x - LW / 2)) / (2.0 * 1.5 * 1.5));
    s += g[x];
  }
  for (int x = 0; x < LW; ++x) g[x] /= s;
  cudaMemcpyToSymbol(c_win, g, sizeof(g));
  float2 g2[LW];
  for (int x = 0; x < LW; ++x) g2[x] = make_float2(g[x], g[x]);
  cudaMemcpyToSymbol(c_win2, g2, sizeof(g2));
  done_dev = dev;
}

static inline uint64_t loss_tiles(int W, int H) { return (uint64_t)((W + LT - 1) / LT) * ((H + LT - 1) / LT); }

}  // namespace tgr

using namespace tgr;

extern "C" uint64_t tgr_image_loss_bytes(int32_t n_views, int32_t W, int32_t H) {
  if (n_views <= 0 || W <= 0 || H <= 0) return 256;
  const uint64_t px = (uint64_t)n_views * 3 * (uint64_t)W * H;
  return align_up(3 * px * 4, 256) + align_up((uint64_t)n_views * 3 * loss_tiles(W, H) * 4, 256) + 256;
}

static int loss_check(int32_t V, int32_t W, int32_t H, const void* pred, const void* target, const void* ws,
                      uint64_t ws_bytes) {
  if (V <= 0 || W <= 0 || H <= 0) { set_error("image_loss: bad sizes V=%d W=%d H=%d", V, W, H); return 1; }
  if (!pred || !target || !ws) { set_error("image_loss: null pointer"); return 1; }
  if (ws_bytes < tgr_image_loss_bytes(V, W, H)) { set_error("image_loss: workspace too small"); return 1; }
  if ((uint64_t)V * 3 > 65535) { set_error("image_loss: too many views per call (%d)", V); return 1; }
  return 0;
}

static LossArgs loss_args(int32_t V, int32_t W, int32_t H, const float* pred, const void* target, const float* view_w,
                          float l1, float l2, float ss, float* grad, void* workspace) {
  const uint64_t px = (uint64_t)V * 3 * (uint64_t)W * H;
  LossArgs a;
  a.pred = pred;
  a.target = target;
  a.view_w = view_w;
  a.maps = static_cast<float*>(workspace);
  a.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up(3 * px * 4, 256));
  a.grad = grad;
  a.W = W; a.H = H; a.V = V;
  // plane_sz * 4 and the maps base are multiples of 16 when W % 4 == 0 (workspace is 16-byte aligned by contract)
  a.vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(workspace)) % 16 == 0) &&
          (reinterpret_cast<uintptr_t>(target) % 16 == 0);
  a.w_l1 = l1; a.w_l2 = l2; a.w_ss = ss;
  return a;
}

#define TGR_LOSS_LAUNCH(K, u8, ssim, grid, s, a)                                                                   \
  do {                                                                                                             \
    if (u8) { if (ssim) K<true, true><<<grid, LTHREADS, 0, s>>>(a); else K<true, false><<<grid, LTHREADS, 0, s>>>(a); }   \
    else    { if (ssim) K<false, true><<<grid, LTHREADS, 0, s>>>(a); else K<false, false><<<grid, LTHREADS, 0, s>>>(a); } \
  } while (0)

extern "C" int tgr_image_loss_forward(int32_t V, int32_t W, int32_t H, const float* pred, const void* target,
                                      int32_t target_is_u8, const float* view_weights, float l1_weight,
                                      float l2_weight, float dssim_weight, float* loss_out, void* workspace,
                                      uint64_t workspace_bytes, void* stream) {
  if (int rc = loss_check(V, W, H, pred, target, workspace, workspace_bytes)) return rc;
  if (!loss_out) { set_error("image_loss: null loss_out"); return 1; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  upload_window();
  const LossArgs a = loss_args(V, W, H, pred, target, view_weights, l1_weight, l2_weight, dssim_weight, nullptr, workspace);
  const dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, V * 3);
  TGR_LOSS_LAUNCH(loss_fwd_kernel, target_is_u8 != 0, dssim_weight != 0.f, grid, s, a);
  loss_finish_kernel<<<1, 1024, 0, s>>>(a.partial, (int)(3 * loss_tiles(W, H)), V, view_weights, loss_out);
  count_launch(2);
  return check_launch("image_loss_forward", false, s);
}

extern "C" int tgr_image_loss_backward(int32_t V, int32_t W, int32_t H, const float* pred, const void* target,
                                       int32_t target_is_u8, const float* dL_dloss_view, float l1_weight,
                                       float l2_weight, float dssim_weight, float* dL_dpred, void* workspace,
                                       uint64_t workspace_bytes, void* stream) {
  if (int rc = loss_check(V, W, H, pred, target, workspace, workspace_bytes)) return rc;
  if (!dL_dpred) { set_error("image_loss: null dL_dpred"); return 1; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  upload_window();
  const LossArgs a = loss_args(V, W, H, pred, target, dL_dloss_view, l1_weight, l2_weight, dssim_weight, dL_dpred, workspace);
  const dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, V * 3);
  TGR_LOSS_LAUNCH(loss_bwd_kernel, target_is_u8 != 0, dssim_weight != 0.f, grid, s, a);
  count_launch();
  return check_launch("image_loss_backward", false, s);
}

extern "C" int tgr_image_loss(int32_t V, int32_t W, int32_t H, const float* pred, const void* target,
                              int32_t target_is_u8, const float* view_weights, float l1_weight, float l2_weight,
                              float dssim_weight, float* loss_out, float* dL_dpred, void* workspace,
                              uint64_t workspace_bytes, void* stream) {
  if (int rc = tgr_image_loss_forward(V, W, H, pred, target, target_is_u8, view_weights, l1_weight, l2_weight,
                                      dssim_weight, loss_out, workspace, workspace_bytes, stream)) return rc;
  if (!dL_dpred) return 0;
  return tgr_image_loss_backward(V, W, H, pred, target, target_is_u8, view_weights, l1_weight, l2_weight, dssim_weight,
                                 dL_dpred, workspace, workspace_bytes, stream);
}
