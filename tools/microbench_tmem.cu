// Micro-benchmark (development aid, not part of the product): TMEM -> register (tcgen05.ld) and register -> TMEM
// (tcgen05.st) throughput per SM as a function of the number of warps issuing, with and without concurrent MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_tmem microbench_tmem.cu && ./microbench_tmem
#include <cstdio>
#include <cuda_runtime.h>
#include "../videopainter_b200/csrc/common.cuh"
using namespace vp;

__global__ void __launch_bounds__(512, 1) k_ldtm(int iters, int mode, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t r[32];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {          // loads: 4 x (32 lanes x 32 cols x 4 B) = 16 KiB per warp per iteration
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_x32(base + ((it + c) & 3) * 32 + (warp >> 2) * 128 % 512, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += r[i];
      }
    } else {                  // stores
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st_x32(base + ((it + c) & 3) * 32 + (warp >> 2) * 128 % 512, r);
      tmem_wait_st();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 2048 * sizeof(long long));
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8, 16}) {
      k_ldtm<<<148, warps * 32>>>(iters, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = 0;
      for (int i = 0; i < 148; ++i) cyc += h[i];
      cyc /= 148;
      const double bytes = (double)iters * 4 * 4096 * warps;
      printf("%s warps=%2d: %.0f cycles, %.1f B/clk/SM (%s)\n", mode ? "STTM" : "LDTM", warps, cyc, bytes / cyc,
             cudaGetErrorString(e));
    }
  return 0;
}
