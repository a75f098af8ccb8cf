// Microbenchmark: tcgen05.ld throughput per SM as a function of the number of warps (B200, sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <int INFLIGHT>
__global__ void k(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[INFLIGHT][32];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) tmem_ld32(base + (uint32_t)(((i * INFLIGHT + j) * 32) & 511), r[j]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j)
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[j][q];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; }
  if (acc == 0x12345678u) out[blockIdx.x * 2 + 1] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
  long long h[2];
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int inflight : {1, 2, 4}) {
      if (inflight == 1) k<1><<<148, warps * 32>>>(d, iters);
      if (inflight == 2) k<2><<<148, warps * 32>>>(d, iters);
      if (inflight == 4) k<4><<<148, warps * 32>>>(d, iters);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)warps * iters * inflight * 32 * 32 * 4;
      printf("warps %2d inflight %d: %lld clk, %.1f B/clk/SM, %.0f clk per x32 load per warp (%s)\n", warps, inflight, h[0],
             bytes / h[0], (double)h[0] / (iters * inflight), cudaGetErrorString(e));
    }
  }
  return 0;
}
