// Prints which SM each CTA of a (4096 x 288-thread, ~25 KB smem) launch lands on while all stay resident.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(288) k(unsigned* smid, long long spin) {
  __shared__ float pad[6300];
  pad[threadIdx.x] = 0;
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  if (threadIdx.x == 0) smid[blockIdx.x] = s;
  long long t0 = clock64();
  while (clock64() - t0 < spin) { }
  if (pad[threadIdx.x] == 1.f) smid[0] = 0;
}
int main() {
  const int G = 1200;
  unsigned* d; cudaMalloc(&d, G * 4);
  k<<<G, 288>>>(d, 2000000);
  cudaDeviceSynchronize();
  unsigned h[G]; cudaMemcpy(h, d, G * 4, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 320; ++i) printf("%u%c", h[i], (i % 37 == 36) ? '\n' : ' ');
  printf("\n");
  int cnt[200] = {0};
  for (int i = 0; i < 1036; ++i) cnt[h[i]]++;
  printf("blocks per SM among first 1036: ");
  for (int i = 0; i < 148; ++i) printf("%d ", cnt[i]);
  printf("\n");
  // do blocks i and i+148 share an SM?
  int same = 0; for (int i = 0; i + 148 < 1036; ++i) same += (h[i] == h[i + 148]);
  printf("h[i]==h[i+148]: %d of %d\n", same, 1036 - 148);
  int same2 = 0; for (int i = 0; i + 1 < 1036; ++i) same2 += (h[i] / 2 == h[i + 1] / 2);
  printf("consecutive blocks on same TPC: %d\n", same2);
  return 0;
}
