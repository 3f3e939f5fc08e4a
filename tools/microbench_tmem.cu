// TMEM -> register bandwidth micro-benchmark (sm_100a): W warps of one CTA per SM each issue tcgen05.ld
// 32x32b.x32 (4 KB per warp instruction) back to back; reports cycles per load per scheduler and the implied
// bytes/clk/SM. Decides whether reading S (fp32) for the softmax is bounded by the TMEM read port.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_ab/microbench_tmem.so tools/microbench_tmem.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int PER_WAIT>
__global__ void bench(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t r[PER_WAIT][32];
#pragma unroll
    for (int k = 0; k < PER_WAIT; ++k) ld32(base + (k & 3) * 32, r[k]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < PER_WAIT; ++k) acc ^= r[k][0] ^ r[k][13] ^ r[k][31];
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int PER_WAIT>
void run(long long* dc, uint32_t* ds) {
  const int iters = 2048;
  printf("tcgen05.ld 32x32b.x32, %d loads per wait:", PER_WAIT);
  for (int warps : {4, 8, 16}) {
    bench<PER_WAIT><<<148, warps * 32>>>(iters, dc, ds);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_load = double(c) / (double(iters) * PER_WAIT);  // one warp's view
    printf("  %2d warps: %6.1f clk/load/warp = %6.1f B/clk/SM", warps, per_load, 4096.0 * warps / per_load);
  }
  printf("\n");
}

int main() {
  long long* dc;
  uint32_t* ds;
  cudaMalloc(&dc, 8);
  cudaMalloc(&ds, 4);
  run<1>(dc, ds);
  run<2>(dc, ds);
  run<4>(dc, ds);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
