// Dev probe: how does the block scheduler spread G one-warp CTAs (G below the machine's CTA slots) over the SMs?
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cta_spread scripts/probes/cta_spread.cu && /tmp/cta_spread
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(32, 32) k_spin(uint32_t* smid_out, long long cycles)
{
  extern __shared__ char dyn[];
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (threadIdx.x == 0) smid_out[blockIdx.x] = smid;
}

static void run(int grid, int limit)
{
  uint32_t* d;
  cudaMalloc(&d, sizeof(uint32_t) * grid);
  size_t dyn = 0;
  if (limit > 0 && limit < 32)
  {
    dyn = 32768 / limit - 1024;
    cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    cudaFuncSetAttribute(k_spin, cudaFuncAttributePreferredSharedMemoryCarveout, 14);  // percent of 228 KB ~ 32 KB
  }
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spin, 32, dyn);
  k_spin<<<grid, 32, dyn>>>(d, 400000);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint32_t> h(grid);
  cudaMemcpy(h.data(), d, sizeof(uint32_t) * grid, cudaMemcpyDeviceToHost);
  std::vector<int> per(256, 0);
  for (auto s : h) per[s]++;
  int used = 0, mx = 0, mn = 1 << 30;
  for (int s = 0; s < 256; ++s)
    if (per[s]) { ++used; mx = std::max(mx, per[s]); mn = std::min(mn, per[s]); }
  std::vector<int> hist(40, 0);
  for (int s = 0; s < 256; ++s) if (per[s]) hist[std::min(per[s], 39)]++;
  printf("grid %d limit %d dyn %zu occ %d err %d: SMs used %d, CTAs per SM min %d max %d; histogram:", grid, limit, dyn, occ, (int)e, used, mn, mx);
  for (int k = 0; k < 40; ++k) if (hist[k]) printf(" %dx%d", hist[k], k);
  printf("\n");
  cudaFree(d);
}

int main()
{
  for (int g : {148, 500, 1024, 2048, 2500, 4096, 4500, 4736})
    run(g, 0);
  run(4096, 28);
  run(2500, 17);
  run(2048, 14);
  return 0;
}
