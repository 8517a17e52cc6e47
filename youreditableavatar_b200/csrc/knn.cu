// knn.cu — distCUDA2: mean squared distance of every point to its 3 nearest other points.
//
// Behavioural reference: simple-knn/simple_knn.cu:29-221 (Morton order, boxes of 1024 points, pruned exact
// search) and spatial.cu:15-26.  The search is exact for P >= 4, so results agree with the reference up to
// fp32 rounding of the distances.  Differences in how: no cudaMalloc / thrust vectors / blocking D2H copies
// inside the call (bounds stay on the device), the hand-written radix sort of sort.cu orders the Morton
// codes, and points are gathered into Morton order once so the inner loops read contiguous float4s.
#include <cfloat>
#include "common.cuh"

namespace tgr {

constexpr int KBOX = 1024;

struct KnnView {
  float* bounds;        // [8] min xyz, max xyz (as ordered-int encodings while reducing)
  uint32_t* codes_a;    // [P]
  uint32_t* idx_a;      // [P]
  uint32_t* codes_b;    // [P]
  uint32_t* idx_b;      // [P]
  float4* sorted_pts;   // [P]
  float4* box_min;      // [nbox]
  float4* box_max;      // [nbox]
  uint32_t* sort_temp;
  uint64_t bytes;
};

__host__ __device__ inline KnnView carve_knn(void* base, int32_t P) {
  KnnView k;
  char* p = static_cast<char*>(base);
  uint64_t n = (uint64_t)(P > 0 ? P : 0);
  uint64_t nb = (n + KBOX - 1) / KBOX;
  k.bounds = carve<float>(p, 32);
  k.codes_a = carve<uint32_t>(p, n);
  k.idx_a = carve<uint32_t>(p, n);
  k.codes_b = carve<uint32_t>(p, n);
  k.idx_b = carve<uint32_t>(p, n);
  k.sorted_pts = carve<float4>(p, n);
  k.box_min = carve<float4>(p, nb);
  k.box_max = carve<float4>(p, nb);
  k.sort_temp = carve<uint32_t>(p, sort_temp_bytes(n) / 4);
  k.bytes = (uint64_t)(p - static_cast<char*>(base)) + 128;
  return k;
}

// order-preserving float <-> int so atomicMin/atomicMax on ints reduce floats
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void knn_bounds_init(int* b) {
  // the reference folds {0,0,0} into both reductions (simple_knn.cu:191): start from zero
  if (threadIdx.x < 6) b[threadIdx.x] = f2ord(0.f);
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, int* __restrict__ b) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = pts[3 * (size_t)i + c];
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicMin(&b[c], f2ord(mn[c]));
      atomicMax(&b[3 + c], f2ord(mx[c]));
    }
  }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts, const int* __restrict__ b,
                                                         uint32_t* __restrict__ codes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float mnx = ord2f(b[0]), mny = ord2f(b[1]), mnz = ord2f(b[2]);
  const float mxx = ord2f(b[3]), mxy = ord2f(b[4]), mxz = ord2f(b[5]);
  const float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
  const uint32_t x = prep_morton((uint32_t)(((px - mnx) / (mxx - mnx)) * ((1 << 10) - 1)));
  const uint32_t y = prep_morton((uint32_t)(((py - mny) / (mxy - mny)) * ((1 << 10) - 1)));
  const uint32_t z = prep_morton((uint32_t)(((pz - mnz) / (mxz - mnz)) * ((1 << 10) - 1)));
  codes[i] = x | (y << 1) | (z << 2);
}

__global__ void __launch_bounds__(256) knn_gather_kernel(int P, const float* __restrict__ pts,
                                                         const uint32_t* __restrict__ idx, float4* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const size_t j = idx[i];
  sorted[i] = make_float4(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2], 0.f);
}

__global__ void __launch_bounds__(256) knn_box_kernel(int P, const float4* __restrict__ sorted, float4* __restrict__ bmin,
                                                      float4* __restrict__ bmax) {
  __shared__ float s[6][8];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  const int base = blockIdx.x * KBOX;
  for (int k = threadIdx.x; k < KBOX; k += 256) {
    const int i = base + k;
    if (i < P) {
      const float4 v = sorted[i];
      mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
      mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s[c][warp] = mn[c]; s[3 + c][warp] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      for (int c = 0; c < 3; ++c) { s[c][0] = fminf(s[c][0], s[c][w]); s[3 + c][0] = fmaxf(s[3 + c][0], s[3 + c][w]); }
    bmin[blockIdx.x] = make_float4(s[0][0], s[1][0], s[2][0], 0.f);
    bmax[blockIdx.x] = make_float4(s[3][0], s[4][0], s[5][0], 0.f);
  }
}

__device__ __forceinline__ void update3(const float4& ref, const float4& pt, float* best) {
  const float dx = pt.x - ref.x, dy = pt.y - ref.y, dz = pt.z - ref.z;
  float dist = dx * dx + dy * dy + dz * dz;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (best[j] > dist) { float t = best[j]; best[j] = dist; dist = t; }
  }
}

__device__ __forceinline__ float dist_box(const float4& bmin, const float4& bmax, const float4& p) {
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (p.x < bmin.x || p.x > bmax.x) dx = fminf(fabsf(p.x - bmin.x), fabsf(p.x - bmax.x));
  if (p.y < bmin.y || p.y > bmax.y) dy = fminf(fabsf(p.y - bmin.y), fabsf(p.y - bmax.y));
  if (p.z < bmin.z || p.z > bmax.z) dz = fminf(fabsf(p.z - bmin.z), fabsf(p.z - bmax.z));
  return dx * dx + dy * dy + dz * dz;
}

// One thread per (Morton-ordered) point; boxes are visited outward from the point's own box so the
// rejection bound tightens early, each candidate box is scanned with contiguous float4 loads.
__global__ void __launch_bounds__(256) knn_search_kernel(int P, const float4* __restrict__ sorted,
                                                         const uint32_t* __restrict__ idx,
                                                         const float4* __restrict__ bmin, const float4* __restrict__ bmax,
                                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 pt = sorted[i];
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  // seed bound from the +-3 Morton neighbours (simple_knn.cu:156-163)
  for (int k = max(0, i - 3); k <= min(P - 1, i + 3); ++k) {
    if (k == i) continue;
    update3(pt, sorted[k], best);
  }
  const float reject = best[2];
  best[0] = best[1] = best[2] = FLT_MAX;
  const int nbox = (P + KBOX - 1) / KBOX;
  const int home = i / KBOX;
  auto visit = [&](int b) {
    const float db = dist_box(bmin[b], bmax[b], pt);
    if (db > reject || db > best[2]) return;
    const int e = min(P, (b + 1) * KBOX);
    for (int k = b * KBOX; k < e; ++k) {
      if (k == i) continue;
      update3(pt, sorted[k], best);
    }
  };
  visit(home);
  for (int off = 1; off < nbox; ++off) {
    if (home + off < nbox) visit(home + off);
    if (home - off >= 0) visit(home - off);
  }
  out[idx[i]] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace tgr

using namespace tgr;

extern "C" uint64_t tgr_knn_bytes(int32_t P) { return carve_knn(nullptr, P).bytes; }

extern "C" int tgr_dist2(int32_t P, const float* points, float* mean_dist2, void* workspace, uint64_t workspace_bytes,
                         void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (P <= 0) return 0;
  if (!workspace || workspace_bytes < tgr_knn_bytes(P)) { set_error("knn workspace too small"); return 1; }
  KnnView k = carve_knn(workspace, P);
  int* b = reinterpret_cast<int*>(k.bounds);
  knn_bounds_init<<<1, 32, 0, s>>>(b);
  count_launch();
  const int blocks = (P + 255) / 256;
  knn_bounds_kernel<<<min(blocks, NUM_SM * 8), 256, 0, s>>>(P, points, b);
  count_launch();
  knn_morton_kernel<<<blocks, 256, 0, s>>>(P, points, b, k.codes_a);
  count_launch();
  bool in_b = false;
  if (int rc = launch_sort_pairs((uint64_t)P, nullptr, k.codes_a, k.idx_a, k.codes_b, k.idx_b, true, 0, 32, k.sort_temp, s, &in_b))
    return rc;
  const uint32_t* idx = in_b ? k.idx_b : k.idx_a;
  knn_gather_kernel<<<blocks, 256, 0, s>>>(P, points, idx, k.sorted_pts);
  count_launch();
  const int nbox = (P + KBOX - 1) / KBOX;
  knn_box_kernel<<<nbox, 256, 0, s>>>(P, k.sorted_pts, k.box_min, k.box_max);
  count_launch();
  knn_search_kernel<<<blocks, 256, 0, s>>>(P, k.sorted_pts, idx, k.box_min, k.box_max, mean_dist2);
  count_launch();
  return check_launch("dist2", false, s);
}
