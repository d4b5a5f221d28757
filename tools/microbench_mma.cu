// Micro-benchmark (development aid): issue-to-completion cost of tcgen05.mma for the shapes the attention kernel uses.
#include <cstdio>
#include <cuda_runtime.h>
#include "../videopainter_b200/csrc/common.cuh"
using namespace vp;

// MODE 0: SS, B K-major; 1: TS (A from TMEM), B K-major; 2: TS, B MN-major.  NACC independent accumulators, round-robin.
template <int MODE, int NACC, int N>
__global__ void __launch_bounds__(128, 1) k_mma(int n_outer, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tm = slot;
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, MODE == 2 ? 1 : 0);
    const uint64_t adesc = make_desc_sw128(smem_u32(smem), 1024, 0);
    const uint64_t bdesc = make_desc_sw128(smem_u32(smem) + 16384, 1024, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < n_outer; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          const uint32_t dcol = tm + a * N;
          if (MODE == 0) mma_ss(dcol, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
          else mma_ts(dcol, tm + 448 + k * 8, MODE == 2 ? bdesc + 128 * k : bdesc + 2 * k, idesc, 1);
        }
      }
    }
    tc_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t2 - t0;
    out[blockIdx.x * 2 + 1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE, int NACC, int N>
void run(long long* d) {
  const int n_outer = 1024;
  cudaFuncSetAttribute(k_mma<MODE, NACC, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_mma<MODE, NACC, N><<<148, 128, 100 * 1024>>>(n_outer, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double tot = 0, iss = 0;
  for (int i = 0; i < 148; ++i) { tot += h[2 * i]; iss += h[2 * i + 1]; }
  const double n = n_outer * 4.0 * NACC;
  const char* names[] = {"SS B=K ", "TS B=K ", "TS B=MN"};
  printf("%s M=128 N=%3d nacc=%d: %.1f clk/MMA complete, %.1f clk/MMA issue, ideal %d (%s)\n", names[MODE], N, NACC,
         tot / 148 / n, iss / 148 / n, N / 2, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 4096 * sizeof(long long));
  run<0, 1, 64>(d);  run<0, 2, 64>(d);  run<0, 4, 64>(d);
  run<0, 1, 128>(d); run<0, 2, 128>(d); run<0, 3, 128>(d);
  run<0, 1, 256>(d);
  run<2, 1, 64>(d);  run<2, 2, 64>(d);  run<2, 4, 64>(d);
  run<1, 1, 128>(d); run<1, 2, 128>(d);
  run<1, 1, 256>(d);
  return 0;
}
