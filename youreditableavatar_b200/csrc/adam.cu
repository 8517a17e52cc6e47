// adam.cu — Adam over every parameter group of the Gaussian model in ONE launch.
//
// Reference: Edit_core/tetgs_scene/tetgs_optimizer.py:66-104 builds `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over
// up to six groups (points, sh dc, sh rest, densities, scales, quaternions) and calls .step() once per iteration
// (:106-108); torch runs that as a chain of multi-tensor kernels (lerp, mul, addcmul, sqrt, div, add, addcdiv),
// each streaming the whole state.  Here: one pass, 16 B read + 12 B written per parameter, straight from the flat
// all-reduced gradient buffer.  The update follows torch.optim.adam._single_tensor_adam (the pinned dependency's
// published algorithm, torch 2.x; no weight decay, no amsgrad, maximize=False):
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
//   step_size = lr / (1 - beta1^t); denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps
//   param.addcdiv_(exp_avg, denom, value=-step_size)
// with the bias corrections evaluated in double on the host like Python does.
#include <cmath>
#include "common.cuh"

namespace tgr {

constexpr int AD_THREADS = 256;
constexpr int AD_VEC = 4;                               // one float4 per thread and iteration
constexpr int AD_ITERS = 4;                             // float4s per thread
constexpr int AD_CHUNK = AD_THREADS * AD_VEC * AD_ITERS;   // 4096 parameters per CTA

struct AdamArgs {
  tgr_adam_group g[TGR_ADAM_MAX_GROUPS];
  uint32_t first_chunk[TGR_ADAM_MAX_GROUPS + 1];   // CTA index where each group starts
  float step_scale[TGR_ADAM_MAX_GROUPS][2];        // -lr / bias_correction1, per group and {lr, lr_alt}
  int n_groups;
  float one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const float ss, const AdamArgs& a) {
  g *= a.grad_scale;
  m = fmaf(g - m, a.one_minus_b1, m);                     // lerp_(grad, 1 - beta1)
  v = fmaf(a.one_minus_b2 * g, g, v * a.b2);              // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = fmaf(ss, m / denom, p);                             // addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(AD_THREADS) adam_kernel(const __grid_constant__ AdamArgs a) {
  int gi = 0;
#pragma unroll 1
  while (gi + 1 < a.n_groups && blockIdx.x >= a.first_chunk[gi + 1]) ++gi;
  const tgr_adam_group& G = a.g[gi];
  const uint64_t begin = (uint64_t)(blockIdx.x - a.first_chunk[gi]) * AD_CHUNK;
  const float ss0 = a.step_scale[gi][0], ss1 = a.step_scale[gi][1];
  const uint32_t period = G.period, split = G.split;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(G.param) | reinterpret_cast<uintptr_t>(G.grad) |
                        reinterpret_cast<uintptr_t>(G.exp_avg) | reinterpret_cast<uintptr_t>(G.exp_avg_sq)) & 15) == 0;
#pragma unroll
  for (int it = 0; it < AD_ITERS; ++it) {
    const uint64_t i = begin + ((uint64_t)it * AD_THREADS + threadIdx.x) * AD_VEC;
    if (i >= G.count) break;
    if (vec_ok && i + AD_VEC <= G.count) {
      float4 p = *reinterpret_cast<const float4*>(G.param + i);
      const float4 g = __ldcs(reinterpret_cast<const float4*>(G.grad + i));
      float4 m = *reinterpret_cast<const float4*>(G.exp_avg + i);
      float4 v = *reinterpret_cast<const float4*>(G.exp_avg_sq + i);
      float s[4] = {ss0, ss0, ss0, ss0};
      if (period) {
        const uint32_t r = (uint32_t)(i % period);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t rk = r + k;
          if (rk >= period) rk -= period;
          s[k] = rk < split ? ss0 : ss1;
        }
      }
      adam_one(p.x, g.x, m.x, v.x, s[0], a);
      adam_one(p.y, g.y, m.y, v.y, s[1], a);
      adam_one(p.z, g.z, m.z, v.z, s[2], a);
      adam_one(p.w, g.w, m.w, v.w, s[3], a);
      *reinterpret_cast<float4*>(G.param + i) = p;
      *reinterpret_cast<float4*>(G.exp_avg + i) = m;
      *reinterpret_cast<float4*>(G.exp_avg_sq + i) = v;
    } else {
      for (uint64_t j = i; j < i + AD_VEC && j < G.count; ++j) {
        float p = G.param[j], m = G.exp_avg[j], v = G.exp_avg_sq[j];
        const float ss = (period && (uint32_t)(j % period) >= split) ? ss1 : ss0;
        adam_one(p, G.grad[j], m, v, ss, a);
        G.param[j] = p;
        G.exp_avg[j] = m;
        G.exp_avg_sq[j] = v;
      }
    }
  }
}

}  // namespace tgr

using namespace tgr;

extern "C" int tgr_adam_step(const tgr_adam_group* groups, int32_t n_groups, int32_t step, double beta1, double beta2,
                             double eps, float grad_scale, void* stream) {
  if (n_groups < 0 || n_groups > TGR_ADAM_MAX_GROUPS) { set_error("adam: %d groups (max %d)", n_groups, TGR_ADAM_MAX_GROUPS); return 1; }
  if (step < 1) { set_error("adam: step must be >= 1 (got %d)", step); return 1; }
  if (n_groups == 0) return 0;
  if (!groups) { set_error("adam: null groups"); return 1; }
  AdamArgs a;
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2 = 1.0 - std::pow(beta2, (double)step);
  uint64_t chunks = 0;
  int n = 0;
  for (int i = 0; i < n_groups; ++i) {
    const tgr_adam_group& g = groups[i];
    if (g.count == 0) continue;
    if (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq) { set_error("adam: null pointer in group %d", i); return 1; }
    if (g.period && g.split > g.period) { set_error("adam: group %d split %u > period %u", i, g.split, g.period); return 1; }
    a.g[n] = g;
    a.first_chunk[n] = (uint32_t)chunks;
    a.step_scale[n][0] = (float)(-(g.lr / bc1));
    a.step_scale[n][1] = (float)(-(g.lr_alt / bc1));
    chunks += (g.count + AD_CHUNK - 1) / AD_CHUNK;
    if (chunks > 0x7fffffffull) { set_error("adam: too many parameters"); return 1; }
    ++n;
  }
  if (n == 0) return 0;
  a.first_chunk[n] = (uint32_t)chunks;
  a.n_groups = n;
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.b2 = (float)beta2;
  a.one_minus_b2 = (float)(1.0 - beta2);
  a.bc2_sqrt = (float)std::sqrt(bc2);
  a.eps = (float)eps;
  a.grad_scale = grad_scale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adam_kernel<<<(unsigned)chunks, AD_THREADS, 0, s>>>(a);
  count_launch();
  return check_launch("adam_step", false, s);
}
