// Micro-benchmark (development aid): reciprocal throughput of the SASS instructions the attention softmax is made of,
// per SM sub-partition, with 1, 2 and 4 resident warps per sub-partition.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_pipes tools/microbench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 512

enum Op { FFMA, FFMA2, FADD2, FMUL2, FMNMX, FMNMX3, EX2, EX2_BF16X2, F2FP, IMAD, LOP3, SHLADD, MIX_EX2_FFMA2, MIX_EX2_FMNMX3,
          MIX_FFMA2_FMNMX3, MIX_EX2_F2FP, MIX_ALL };
static const char* names[] = {"FFMA", "FFMA2", "FADD2", "FMUL2", "FMNMX", "FMNMX3", "MUFU.EX2", "MUFU.EX2.bf16x2", "F2FP.pack",
                              "IMAD", "LOP3", "SHL+IADD(LEA)", "mix EX2+FFMA2", "mix EX2+FMNMX3", "mix FFMA2+FMNMX3",
                              "mix EX2+F2FP", "mix EX2+FFMA2+FMNMX3+F2FP"};

template <int OP>
__global__ void k(long long* out, float seed) {
  float f[CHAINS];
  uint64_t d[CHAINS];
  uint32_t u[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    f[i] = seed * (i + 1) + threadIdx.x * 1e-6f;
    u[i] = __float_as_uint(f[i]);
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d[i]) : "f"(f[i]), "f"(f[i] * 0.5f));
  }
  const float c0 = seed * 0.999f, c1 = seed * 1e-3f;
  uint64_t dc0, dc1;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(dc0) : "f"(c0), "f"(c0));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(dc1) : "f"(c1), "f"(c1));
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == FFMA || OP == MIX_ALL) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c0), "f"(c1));
      if (OP == FFMA2 || OP == MIX_EX2_FFMA2 || OP == MIX_FFMA2_FMNMX3 || OP == MIX_ALL)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(dc0), "l"(dc1));
      if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dc1));
      if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dc0));
      if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
      if (OP == FMNMX3 || OP == MIX_EX2_FMNMX3 || OP == MIX_FFMA2_FMNMX3 || OP == MIX_ALL)
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c0));
      if (OP == EX2 || OP == MIX_EX2_FFMA2 || OP == MIX_EX2_FMNMX3 || OP == MIX_EX2_F2FP || OP == MIX_ALL)
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i]));
      if (OP == EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == F2FP || OP == MIX_EX2_F2FP || OP == MIX_ALL)
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(c0));
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(__float_as_uint(c0)), "r"(__float_as_uint(c1)));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(__float_as_uint(c0)), "r"(__float_as_uint(c1)));
      if (OP == SHLADD) {
        asm volatile("{.reg .b32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1;}" : "+r"(u[i]) : "r"(__float_as_uint(c0)));
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float a, b;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d[i]));
    acc += f[i] + a + b + __uint_as_float(u[i]);
  }
  if (acc == 123.456f) out[4096] = 1;
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int OP>
void run(long long* d, int n_instr_per_iter) {
  printf("%-28s", names[OP]);
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 128 * wps;
    k<OP><<<148, threads>>>(d, 1.0001f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" error %s", cudaGetErrorString(e)); continue; }
    static long long h[148 * 16];
    const int nw = 148 * (threads / 32);
    cudaMemcpy(h, d, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < nw; ++i) mx += (double)h[i];
    mx /= nw;
    // cycles per warp-instruction per sub-partition = elapsed / (instructions one warp issued * warps per sub-partition)
    printf("  %dw/SMSP: %6.2f clk/inst", wps, mx / ((double)ITERS * CHAINS * n_instr_per_iter * wps));
  }
  printf("\n");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8192 * sizeof(long long));
  run<FFMA>(d, 1); run<FFMA2>(d, 1); run<FADD2>(d, 1); run<FMUL2>(d, 1); run<FMNMX>(d, 1); run<FMNMX3>(d, 1);
  run<EX2>(d, 1); run<EX2_BF16X2>(d, 1); run<F2FP>(d, 1); run<IMAD>(d, 1); run<LOP3>(d, 1); run<SHLADD>(d, 2);
  run<MIX_EX2_FFMA2>(d, 2); run<MIX_EX2_FMNMX3>(d, 2); run<MIX_FFMA2_FMNMX3>(d, 2); run<MIX_EX2_F2FP>(d, 2);
  run<MIX_ALL>(d, 5);
  return 0;
}
