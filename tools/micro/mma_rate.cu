// Microbenchmark: issue rate of tcgen05.mma kind::f16 (bf16 x bf16 -> fp32, M = 128 per CTA, K = 16) on B200 as a function of
//   N (64..256), the number of independent accumulators the stream rotates over, the A source (shared memory descriptor
//   with conv-like tap offsets, or TMEM) and cta_group (1, or 2 = CTA pair with M = 256 and B split over the pair).
// It answers: what bounds the 3x3 conv kernel's N = 64 stream (measured 59-72 clk per MMA against a 32 clk tensor floor)?
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}

// operand fill: constant (low toggle rate) or pseudo-random bf16 pairs in +-[0.5, 2) (what real activations / weights toggle like)
__device__ int g_random;
__device__ __forceinline__ uint32_t fill_word(int i, int random) {
  if (!random) return 0x3c003c00u + (uint32_t)(i & 7);
  uint32_t x = (uint32_t)i * 2654435761u + (uint32_t)blockIdx.x * 40503u;
  x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
  const uint32_t lo = 0x3F00u | (x & 0x80FFu), hi = 0x3F00u | ((x >> 16) & 0x80FFu);
  return lo | (hi << 16);
}
constexpr int ROW_BYTES = 10 * 128;   // halo row of the conv kernel: 8 + 2 pixels x 128 B
constexpr int A_BYTES = 24 * 1024;
constexpr int B_BYTES = 96 * 1024;

// MODE 0: A from shared memory, tap offsets as in the conv kernel; MODE 1: A from shared memory, one fixed tile; MODE 2: A from TMEM
template <int CG, int NACC, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = fill_word(i, g_random);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t idesc = idesc_bf16(128 * CG, N);
    const uint64_t a_base = desc_sw128(smem_u32(smem), MODE == 0 ? ROW_BYTES : 1024);
    const uint64_t b_base = desc_sw128(smem_u32(smem + A_BYTES), 1024);
    const uint32_t tap16 = (uint32_t)((N / CG) * 128) >> 4;
    const uint32_t a_tmem = tmem + 480;
    if (rank == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
          for (int i = 0; i < 36; ++i) {
            const int s = i / 9, t = i % 9, ky = t / 3, kx = t % 3;
            const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
            const uint64_t b = b_base + (uint32_t)(t % 3) * tap16 + (uint32_t)(s * 2);
            if (MODE == 2) mma_ts<CG>(d, a_tmem + (uint32_t)(s * 8), b, idesc);
            else mma_ss<CG>(d, a_base + (uint32_t)(MODE == 0 ? ((kx * 128 + ky * ROW_BYTES) >> 4) : 0) + (uint32_t)(s * 2), b, idesc);
          }
        }
        __syncwarp();
      }
      if ((threadIdx.x & 31) == 0) {
        if (CG == 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      }
    }
    if ((threadIdx.x & 31) == 0) mbar_wait(smem_u32(&bar), 0);
    __syncwarp();
    t1 = clock64();
    if (rank == 0 && threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG, int NACC, int MODE>
void run(long long* d, int N, int grid) {
  if (NACC * N > (MODE == 2 ? 448 : 512)) return;
  const int iters = 200;
  auto kern = k<CG, NACC, MODE>;
  const size_t smem = A_BYTES + B_BYTES + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaMemset(d, 0, 1024 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, kern, d, N, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mn = 1LL << 60, mx = 0;
  for (int i = 0; i < grid; i += CG) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; }
  const double per = (double)mn / (iters * 36.0), floor = 128.0 * N / 256.0 / 1.0;   // tensor floor per CTA: M=128 x N x K=16 at 4096 MAC/clk
  printf("cg%d N=%3d acc=%d %s grid=%3d: %6.1f clk/MMA (max CTA %6.1f)  tensor floor %5.1f  -> %4.0f%% of peak  (%s)\n", CG, N, NACC,
         MODE == 0 ? "A=smem(taps)" : (MODE == 1 ? "A=smem(fixed)" : "A=tmem      "), grid, per, (double)mx / (iters * 36.0), floor,
         100.0 * floor / per, cudaGetErrorString(e));
  fflush(stdout);
}


// ---- strip pattern of conv_lin.cu: 12 MMAs N=192 per strip (A start offsets ky * 4096, dense atoms), NC tcgen05.commit per strip
// to distinct mbarriers (nobody waits on them until the end), accumulators alternate between two TMEM slots.
template <int NC>
__global__ void __launch_bounds__(128, 1) kstrip(long long* out, int N, int strips, int polls, int between) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[6];
  __shared__ uint32_t slot;
  __shared__ volatile uint32_t stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = fill_word(i, g_random);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = idesc_bf16(128, N);
    const uint64_t a_base = desc_sw128(smem_u32(smem), 1024);
    const uint64_t b_base = desc_sw128(smem_u32(smem + A_BYTES), 1024);
    long long t0 = clock64();
    if ((threadIdx.x & 31) == 0) {   // a barrier whose phase 0 is already complete: waiting on it never blocks
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[4])) : "memory");
    }
    __syncwarp();
    for (int g = 0; g < strips; ++g) {
      // what the conv kernel's MMA warp does between strips (the barriers are already satisfied there too)
      if (between & 1) { if ((threadIdx.x & 31) == 0) mbar_wait(smem_u32(&bars[4]), 0); __syncwarp(); }
      if (between & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (between & 4) { if ((threadIdx.x & 31) == 0) out[512 + (g & 63)] = clock64(); }
      if (between & 8) { if ((threadIdx.x & 31) == 0) mbar_wait(smem_u32(&bars[2 + (g & 1)]), (uint32_t)((((g >> 1) & 1) ^ 1))); __syncwarp(); }   // the commit barrier of strip g - 2 (2-slot pipeline)
      if ((threadIdx.x & 31) == 0) {
        const uint32_t d = tmem + (uint32_t)((g & 1) * N);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int s = 0; s < 4; ++s)
            mma_ss<1>(d, a_base + (uint32_t)((ky * 4096 + s * 32) >> 4), b_base + (uint32_t)(((ky % 3) * N * 128 + s * 32) >> 4), idesc);
        if (NC >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[g & 1])) : "memory");
        if (NC >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2 + (g & 1)])) : "memory");
      }
      __syncwarp();
    }
    if ((threadIdx.x & 31) == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
      // all earlier phases of bars[0] have completed in order; wait for the final one
      const uint32_t fin = (uint32_t)(((NC >= 1 ? (strips + 1) / 2 : 0)) & 1);
      mbar_wait(smem_u32(&bars[0]), fin);
    }
    __syncwarp();
    long long t1 = clock64();
    if (threadIdx.x == 32) { out[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp >= 2 && polls) {
    // other warps polling shared memory the way waiting epilogue / store / producer warps do
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(v) : "r"(smem_u32(&bars[5])), "r"(1u) : "memory");
      acc += v;
      if (polls > 1) break;
    }
    if (acc == 0xffffffffu) out[1000] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int NC>
void run_strip(long long* d, int N, int polls, int between = 0) {
  const int strips = 400;
  auto kern = kstrip<NC>;
  const size_t smem = A_BYTES + B_BYTES + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(d, 0, 1024 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) kern<<<148, 128, smem>>>(d, N, strips, polls, between);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mn = 1LL << 60;
  for (int i = 0; i < 148; ++i) if (h[i] < mn) mn = h[i];
  printf("strip pattern N=%3d, %d commit(s) per 12 MMAs, pollers %d, between-strip ops %2d: %6.1f clk/MMA (%s)\n", N, NC, polls, between, (double)mn / (strips * 12.0), cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  { int r = getenv("MMA_RANDOM") ? 1 : 0; cudaMemcpyToSymbol(g_random, &r, sizeof(int)); printf("operands: %s\n", r ? "random" : "constant"); }
  if (getenv("MMA_STRIP")) {
    long long* d0; cudaMalloc(&d0, 1024 * sizeof(long long));
    for (int N : {192, 64, 128}) { run_strip<0>(d0, N, 0); run_strip<1>(d0, N, 0); run_strip<2>(d0, N, 0); run_strip<2>(d0, N, 1); }
    for (int bt : {0, 1, 2, 4, 7, 8, 15}) run_strip<2>(d0, 192, 0, bt);
    for (int bt : {0, 7, 8}) run_strip<2>(d0, 64, 0, bt);
    return 0;
  }
  long long* d;
  cudaMalloc(&d, 1024 * sizeof(long long));
  for (int grid : {2, 148}) {
    for (int N : {64, 128, 192, 256}) {
      run<1, 1, 0>(d, N, grid); run<1, 2, 0>(d, N, grid); run<1, 4, 0>(d, N, grid); run<1, 8, 0>(d, N, grid);
      run<1, 4, 1>(d, N, grid);
      run<1, 1, 2>(d, N, grid); run<1, 2, 2>(d, N, grid); run<1, 4, 2>(d, N, grid);
      run<2, 1, 0>(d, N, grid); run<2, 2, 0>(d, N, grid); run<2, 4, 0>(d, N, grid); run<2, 8, 0>(d, N, grid);
      run<2, 2, 2>(d, N, grid);
    }
  }
  return 0;
}
