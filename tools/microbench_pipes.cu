// Issue-rate micro-benchmark of the instructions the attention softmax is made of (sm_100a).
// For each instruction: one CTA per SM, W warps (1, 2, 4 per scheduler), every thread runs 8 independent
// dependency chains; reports cycles per warp-instruction per scheduler (SMSP).  Build and run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench_pipes tools/microbench_pipes.cu
// Used to decide which pipe (FMA / ALU / XU) each softmax step should run on; results in profiles/.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CHAINS 8

enum Op { EX2, CVT_BF16X2, FMA2, FMNMX3, IMAD, SHL_ADD, PRMT, IADD, FFMA, CVT_BF16, EX2_PLUS_CVT, EX2_PLUS_FMA2, CVT_PLUS_FMA2, NOPS };
static const char* kNames[] = {"ex2.approx.ftz.f32", "cvt.rn.bf16x2.f32", "fma.rn.f32x2", "max.f32 (3-input)", "mad.lo.u32",
                               "shl+add (LEA?)", "prmt.b32", "add.u32", "fma.rn.f32", "cvt.rn.bf16.f32",
                               "ex2 + cvt.bf16x2 (1:1)", "ex2 + fma.f32x2 (1:1)", "cvt.bf16x2 + fma.f32x2 (1:1)"};

template <int OP>
__global__ void bench(int iters, long long* cycles, float* sink) {
  float a[CHAINS];
  uint32_t u[CHAINS];
  uint64_t d[CHAINS];
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = -0.001f * (threadIdx.x + i + 1);
    u[i] = threadIdx.x * 2654435761u + i;
    d[i] = (uint64_t(__float_as_uint(a[i])) << 32) | __float_as_uint(0.5f + i);
  }
  const float c1 = 0.9999f, c2 = 1e-6f;
  uint64_t dc1 = (uint64_t(__float_as_uint(c1)) << 32) | __float_as_uint(c1);
  uint64_t dc2 = (uint64_t(__float_as_uint(c2)) << 32) | __float_as_uint(c2);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == EX2 || OP == EX2_PLUS_CVT || OP == EX2_PLUS_FMA2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == CVT_BF16X2 || OP == EX2_PLUS_CVT || OP == CVT_PLUS_FMA2)
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(a[(i + 1) % CHAINS]));
      if (OP == FMA2 || OP == EX2_PLUS_FMA2 || OP == CVT_PLUS_FMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(dc1), "l"(dc2));
      if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % CHAINS]), "f"(c2));
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(0x800000u));
      if (OP == SHL_ADD) asm volatile("{.reg .u32 t; shl.b32 t, %1, 23; add.u32 %0, %0, t;}" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
      if (OP == CVT_BF16) {
        unsigned short h;
        asm volatile("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(a[i]));
        a[i] = __uint_as_float(uint32_t(h) << 16);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < CHAINS; ++i) s += a[i] + __uint_as_float(u[i]) + __uint_as_float(uint32_t(d[i])) + __uint_as_float(uint32_t(d[i] >> 32));
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
void run(long long* dc, float* ds) {
  const int iters = 4096;
  printf("%-30s", kNames[OP]);
  for (int warps : {4, 8, 16}) {
    bench<OP><<<148, warps * 32>>>(iters, dc, ds);
    cudaDeviceSynchronize();
    bench<OP><<<148, warps * 32>>>(iters, dc, ds);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    const int per_it = (OP >= EX2_PLUS_CVT) ? 2 * CHAINS : CHAINS;
    printf("  %2d warps/SMSP: %6.2f clk/inst", warps / 4, double(c) / (double(iters) * per_it * (warps / 4)));
  }
  printf("\n");
}

int main() {
  long long* dc;
  float* ds;
  cudaMalloc(&dc, 8);
  cudaMalloc(&ds, 4);
  run<EX2>(dc, ds);
  run<CVT_BF16X2>(dc, ds);
  run<CVT_BF16>(dc, ds);
  run<FMA2>(dc, ds);
  run<FFMA>(dc, ds);
  run<FMNMX3>(dc, ds);
  run<IMAD>(dc, ds);
  run<SHL_ADD>(dc, ds);
  run<PRMT>(dc, ds);
  run<IADD>(dc, ds);
  run<EX2_PLUS_CVT>(dc, ds);
  run<EX2_PLUS_FMA2>(dc, ds);
  run<CVT_PLUS_FMA2>(dc, ds);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
