// sort_vs_cub.cu — head-to-head: this library's hand-written onesweep (csrc/sort.cu, through the C ABI) against
// cub::DeviceRadixSort::SortPairs of the CUDA 12.9 toolkit (CCCL: onesweep, Policy1000 on sm_100) — the call the reference
// makes at cuda_rasterizer/rasterizer_impl.cu:303-308.  Same n, same bit range, same random keys; CUDA events around the
// sort calls only (inputs are restored between iterations outside the timed region), median of ITER runs.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude tools/cuda/sort_vs_cub.cu \
//        -Lyoureditableavatar_b200 -ltetgs_rast -Xlinker -rpath -Xlinker '$ORIGIN/../../youreditableavatar_b200' -o tools/cuda/sort_vs_cub
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "tetgs_rast.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int ITER = 15, SEGS = 8;

static float median(std::vector<float> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

template <typename K>
static void fill(std::vector<K>& h, int bits, uint64_t seed) {
  std::mt19937_64 g(seed);
  const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
  for (auto& x : h) x = (K)(g() & mask);
}

struct Case { const char* name; size_t n; int bits; };

// CUB, one call per segment (it has no batched pair sort with per-segment bit ranges worth using here: segmented sort is a different algorithm)
template <typename K>
static float time_cub(size_t n, int bits, int segs) {
  std::vector<K> hk(n);
  fill(hk, bits, 1);
  std::vector<K*> ki(segs), ko(segs);
  std::vector<uint32_t*> vi(segs), vo(segs);
  K* src; CK(cudaMalloc(&src, n * sizeof(K)));
  CK(cudaMemcpy(src, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice));
  for (int s = 0; s < segs; ++s) {
    CK(cudaMalloc(&ki[s], n * sizeof(K))); CK(cudaMalloc(&ko[s], n * sizeof(K)));
    CK(cudaMalloc(&vi[s], n * 4)); CK(cudaMalloc(&vo[s], n * 4));
    CK(cudaMemset(vi[s], 0, n * 4));
  }
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, ki[0], ko[0], vi[0], vo[0], (int)n, 0, bits);
  std::vector<void*> tmp(segs);
  for (int s = 0; s < segs; ++s) CK(cudaMalloc(&tmp[s], tb));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int it = 0; it < ITER + 3; ++it) {
    for (int s = 0; s < segs; ++s) CK(cudaMemcpyAsync(ki[s], src, n * sizeof(K), cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int s = 0; s < segs; ++s) cub::DeviceRadixSort::SortPairs(tmp[s], tb, ki[s], ko[s], vi[s], vo[s], (int)n, 0, bits);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    if (it >= 3) ms.push_back(t);
  }
  for (int s = 0; s < segs; ++s) { cudaFree(ki[s]); cudaFree(ko[s]); cudaFree(vi[s]); cudaFree(vo[s]); cudaFree(tmp[s]); }
  cudaFree(src);
  return median(ms);
}

static float time_ours(size_t n, int bits, int segs, bool* ok) {
  std::vector<uint32_t> hk(n);
  fill(hk, bits, 1);
  uint32_t* src; CK(cudaMalloc(&src, n * 4));
  CK(cudaMemcpy(src, hk.data(), n * 4, cudaMemcpyHostToDevice));
  std::vector<uint32_t*> ka(segs), va(segs), kb(segs), vb(segs);
  std::vector<void*> tmp(segs);
  std::vector<uint64_t> ns(segs, n);
  const uint64_t tb = tgr_sort_temp_bytes(n);
  std::vector<uint32_t> iota(n);
  for (size_t i = 0; i < n; ++i) iota[i] = (uint32_t)i;
  for (int s = 0; s < segs; ++s) {
    CK(cudaMalloc(&ka[s], n * 4)); CK(cudaMalloc(&va[s], n * 4)); CK(cudaMalloc(&kb[s], n * 4)); CK(cudaMalloc(&vb[s], n * 4));
    CK(cudaMalloc(&tmp[s], tb));
  }
  uint32_t* isrc; CK(cudaMalloc(&isrc, n * 4));
  CK(cudaMemcpy(isrc, iota.data(), n * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  int32_t in_b = 0;
  for (int it = 0; it < ITER + 3; ++it) {
    for (int s = 0; s < segs; ++s) {
      CK(cudaMemcpyAsync(ka[s], src, n * 4, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpyAsync(va[s], isrc, n * 4, cudaMemcpyDeviceToDevice));
    }
    CK(cudaEventRecord(e0));
    if (tgr_sort_pairs_u32_batch(segs, ns.data(), ka.data(), va.data(), kb.data(), vb.data(), 0, bits, tmp.data(), &in_b, nullptr)) {
      printf("ours failed: %s\n", tgr_last_error()); exit(1);
    }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    if (it >= 3) ms.push_back(t);
  }
  // check the last run: sorted and stable (values = original positions)
  std::vector<uint32_t> rk(n), rv(n);
  CK(cudaMemcpy(rk.data(), in_b ? kb[0] : ka[0], n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rv.data(), in_b ? vb[0] : va[0], n * 4, cudaMemcpyDeviceToHost));
  bool good = true;
  for (size_t i = 1; i < n && good; ++i) good = rk[i - 1] < rk[i] || (rk[i - 1] == rk[i] && rv[i - 1] < rv[i]);
  for (size_t i = 0; i < n && good; i += 997) good = hk[rv[i]] == rk[i];
  *ok = good;
  for (int s = 0; s < segs; ++s) { cudaFree(ka[s]); cudaFree(va[s]); cudaFree(kb[s]); cudaFree(vb[s]); cudaFree(tmp[s]); }
  cudaFree(src); cudaFree(isrc);
  return median(ms);
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("# %s, CUB %d.%d.%d, median of %d runs, CUDA events around the sort calls\n", pr.name, CUB_MAJOR_VERSION, CUB_MINOR_VERSION,
         CUB_SUBMINOR_VERSION, ITER);
  printf("# bytes = n * (key + 4) * 2 per radix pass (read + write) + one key read for the histograms\n");
  const Case cases[] = {{"depth sort, 1M Gaussians, all 32 key bits", 1000000, 32},
                        {"depth sort, 1M Gaussians, 24 varying bits", 1000000, 24},
                        {"tile sort, 1.87M instances (C3 view), 12 bits", 1870000, 12},
                        {"tile sort, 7.5M instances (C5 view), 14 bits", 7500000, 14}};
  printf("%-50s %5s | %12s %12s %7s | %12s %12s %7s\n", "same n, same bits, u32 keys + u32 values", "segs", "CUB ms", "ours ms", "x", "CUB GB/s", "ours GB/s", "ok");
  for (const Case& c : cases)
    for (int segs : {1, SEGS}) {
      if (c.n > 4000000 && segs > 2) segs = 2;
      bool ok = false;
      const float tc = time_cub<uint32_t>(c.n, c.bits, segs);
      const float to = time_ours(c.n, c.bits, segs, &ok);
      const int pc = (c.bits + 7) / 8;   // both use 8-bit digits at most
      const double bytes = (double)segs * c.n * (8.0 * 2 * pc + 4);
      printf("%-50s %5d | %12.4f %12.4f %7.2f | %12.0f %12.0f %7s\n", c.name, segs, tc, to, tc / to, bytes / tc / 1e6, bytes / to / 1e6, ok ? "yes" : "NO");
    }
  // the reference's own sort: 64-bit tile|depth keys over all instances, 32 + 13 bits at 1024^2 (rasterizer_impl.cu:300-308)
  // against what replaces it here: depth sort of the Gaussians + tile sort of the instances
  printf("\n%-62s %5s | %12s\n", "the reference's sort vs its replacement (per view of C3)", "segs", "ms");
  for (int segs : {1, SEGS}) {
    bool ok1, ok2;
    const float tref = time_cub<uint64_t>(1870000, 45, segs);
    const float td = time_ours(1000000, 24, segs, &ok1), tt = time_ours(1870000, 12, segs, &ok2);
    printf("%-62s %5d | %12.4f\n", "CUB SortPairs<u64,u32>, 1.87M instances, 45 bits", segs, tref);
    printf("%-62s %5d | %12.4f  (%.4f + %.4f), %.2fx\n", "ours: depth sort 1M x 24 bits + tile sort 1.87M x 12 bits", segs, td + tt, td, tt, tref / (td + tt));
  }
  return 0;
}
