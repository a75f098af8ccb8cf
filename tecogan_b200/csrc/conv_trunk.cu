// Fused generator trunk on tcgen05: input conv + N residual blocks (2N+1 layers of 3x3 64->64) in ONE launch.
//
// Replaces the per-layer launches of generator_F's input stage and residual blocks (reference lib/frvsr.py:50-70) when
// the whole frame is a single wave of tiles (tiles <= SMs, e.g. 128x128 LR = 128 tiles).  Why: with one launch per
// layer a 16x8 tile's critical path is ~3.8 us, but every layer additionally paid ~2.4 us of grid-completion latency,
// prologue (barrier init, TMEM alloc, cluster sync) and weight fetch that nothing could hide (one 128 KB CTA per SM,
// profiles/conv_tc_r01_notes.md).  Here each CTA keeps its tile for the whole chain:
//   * layer l+1 starts as soon as the CTA's own tile and its <= 8 neighbours have published layer l
//     (per-tile counters in global memory, st.release / ld.acquire at gpu scope + async-proxy fences for the TMA reads);
//   * weights are double-buffered (2 x 72 KB): layer l+2 streams in while layers l, l+1 compute;
//   * TMEM, mbarriers and the bias table live across layers.
// Layer schedule (buffers A, B are NHWC bf16 [N,H,W,64]):  l = 0: X -> A (ReLU);  odd l: A -> B (ReLU);
// even l >= 2: B -> A with "+ A" (the residual).  Operand layouts / MMA issue / epilogue are those of conv_tc_sw.cu
// (one SWIZZLE_128B halo box with the horizontal taps as descriptor start offsets, 3 K-split accumulator chains,
// 8 epilogue warps writing a swizzled staging tile that one TMA store sends out; the residual stays in registers).
//
// All CTAs must be co-resident (they wait on each other): grid <= SM count and ~175 KB smem => one CTA per SM; the
// host refuses other shapes.
#include <type_traits>
#include "teco_common.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr int TILE_ROWS = 16, HALO_ROWS = 18, CH = 64;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr uint32_t ROW_BYTES = 10 * 128;                         // one halo row: 8 + 2 pixels x 128 B
constexpr uint32_t HALO_BYTES = HALO_ROWS * ROW_BYTES;            // ONE 18x10 box; horizontal taps = 128-byte start offsets
constexpr uint32_t HALO_REGION = (HALO_BYTES + 1023u) & ~1023u;   // 23 KB; doubles as the output staging tile (16 KB)
constexpr uint32_t W_LAYER_BYTES = 9 * CH * 128;                  // 72 KB
constexpr uint32_t SLAB_BYTES = 3 * CH * 128;                     // 3 taps
constexpr int MAX_LAYERS = 48;

struct TrunkParams {
  int N, H, W, L;
  int tiles_x, tiles_y, num_tiles;
  const uint8_t* wpk;        // [L][9][64][64] bf16, SW128 image (teco_pack_conv3x3_bf16)
  const float* bias;         // [L][64]
  __nv_bfloat16* a;
  __nv_bfloat16* b;
  unsigned int* flags;       // [num_tiles] layers published, zeroed by the host before launch
  long long* dbg;            // optional [cta][8 layers][8] clock64 stamps (teco_debug_timing)
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// tm_x / tm_a / tm_b: halo boxes (64 ch, 10 px, 18 rows) for loads; tm_ao / tm_bo: output tiles (64 ch, 8 px, 16 rows) of A / B
__global__ void __launch_bounds__(NUM_THREADS, 1)
trunk64_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_a,
                  const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_ao,
                  const __grid_constant__ CUtensorMap tm_bo, const TrunkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* halo = smem;                                  // one 18x10 halo box; output staging tile during the epilogue
  uint8_t* wbuf = smem + HALO_REGION;                    // 2 layers of weights
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + 2 * W_LAYER_BYTES);
  uint64_t* halo_full = bars;        // [1]
  uint64_t* halo_empty = bars + 1;   // [1]
  uint64_t* w_full = bars + 2;       // [2]
  uint64_t* w_empty = bars + 4;      // [2]
  uint64_t* acc_full = bars + 6;     // [1]
  uint64_t* acc_empty = bars + 7;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_bias = reinterpret_cast<float*>(bars + 9);    // [L][64]

  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int n = tile / p.tiles_y;
  const int x0 = tx * 8, y0 = ty * TILE_ROWS;
  long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;
#define TSTAMP(l, i) do { if (dbg && (l) < 8) dbg[(l) * 8 + (i)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(halo_full), 1);
    mbar_init(smem_u32(halo_empty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&w_full[i]), 1);
      mbar_init(smem_u32(&w_empty[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(acc_empty), NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ao) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.L * CH; i += NUM_THREADS) s_bias[i] = p.bias[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: weights (2 layers ahead), neighbour hand-shake, halo TMA =====================
    // one layer of weights = nine 8 KB bulk copies issued by nine lanes (a single 72 KB copy is processed serially)
    auto load_layer_weights = [&](int layer, int buf) {
      if (lane == 0) mbar_expect_tx(smem_u32(&w_full[buf]), W_LAYER_BYTES);
      __syncwarp();
      if (lane < 9)
        bulk_load_1d(smem_u32(wbuf + (size_t)buf * W_LAYER_BYTES + (size_t)lane * (CH * 128)),
                     p.wpk + (size_t)layer * W_LAYER_BYTES + (size_t)lane * (CH * 128), CH * 128, smem_u32(&w_full[buf]));
    };
    for (int l = 0; l < 2 && l < p.L; ++l) load_layer_weights(l, l);   // constants: fetch before the dependency wait
    pdl_wait();
    // the (up to) nine tiles whose layer-l output this tile's layer-(l+1) halo reads
    const unsigned int* nb = nullptr;
    if (lane < 9) {
      const int dy = lane / 3 - 1, dx = lane % 3 - 1;
      const int nty = ty + dy, ntx = tx + dx;
      if (nty >= 0 && nty < p.tiles_y && ntx >= 0 && ntx < p.tiles_x) nb = p.flags + ((size_t)n * p.tiles_y + nty) * p.tiles_x + ntx;
    }
    for (int l = 0; l < p.L; ++l) {
      if (l > 0) {
        if (nb) {
          while (ld_acquire(nb) < (unsigned int)l) { }   // pure spin: __nanosleep's quantum (~us) was most of a 5.6k-cycle stall
        }
        __syncwarp();
        if (lane == 0) TSTAMP(l, 0);
        fence_proxy_async_global();      // the neighbours' generic-proxy stores are ordered before our TMA reads
        if (lane == 0) TSTAMP(l, 1);
      }
      if (lane == 0) {
        mbar_wait(smem_u32(halo_empty), (uint32_t)((l & 1) ^ 1));
        mbar_expect_tx(smem_u32(halo_full), HALO_BYTES);
        const CUtensorMap* tm = (l == 0) ? &tm_x : ((l & 1) ? &tm_a : &tm_b);
        tma_load_4d(smem_u32(halo), tm, smem_u32(halo_full), 0, x0 - 1, y0 - 1, n);
      }
      __syncwarp();
      if (l + 2 < p.L) {    // weights of layer l+2 go into this layer's buffer once its MMAs have retired
        const int buf = l & 1;
        if (lane == 0) mbar_wait(smem_u32(&w_empty[buf]), (uint32_t)((l >> 1) & 1));
        __syncwarp();
        load_layer_weights(l + 2, buf);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc(CH);
    const uint32_t halo_addr = smem_u32(halo);
    const uint64_t a_base = umma_desc_sw128(halo_addr, ROW_BYTES);
    for (int l = 0; l < p.L; ++l) {
      const int buf = l & 1;
      mbar_wait_warp(smem_u32(acc_empty), (uint32_t)((l & 1) ^ 1));   // epilogue of layer l-1 has drained TMEM
      mbar_wait_warp(smem_u32(&w_full[buf]), (uint32_t)((l >> 1) & 1));
      mbar_wait_warp(smem_u32(halo_full), (uint32_t)(l & 1));
      tcgen05_fence_after();
      if (lane == 0) TSTAMP(l, 2);
      const uint64_t b_base = umma_desc_sw128(smem_u32(wbuf + (size_t)buf * W_LAYER_BYTES), 1024u);
      if (elect_one()) {
        // slab g = tap row ky; order k-step outer, kx inner: consecutive MMAs hit the three K-split chains in turn
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
          for (int s = 0; s < CH / 16; ++s) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t a_off16 = ((uint32_t)kx * 128u + (uint32_t)g * ROW_BYTES + (uint32_t)s * 32u) >> 4;
              const uint32_t b_off16 = ((uint32_t)(g * 3 + kx) * (CH * 128) + (uint32_t)s * 32u) >> 4;
              umma_bf16(tmem_base + (uint32_t)kx * CH, a_base + a_off16, b_base + b_off16, idesc, (g | s) ? 1u : 0u);
            }
          }
        }
        tcgen05_commit(smem_u32(halo_empty));
        tcgen05_commit(smem_u32(&w_empty[buf]));
        tcgen05_commit(smem_u32(acc_full));
      }
      if (lane == 0) TSTAMP(l, 3);
      __syncwarp();
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    pdl_wait();
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int m = 32 * q + lane;
    const int c0 = chalf * 32;
    const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
    unsigned int* my_flag = p.flags + blockIdx.x;
    uint32_t resp[16];   // residual carried across layers (packed bf16 pairs)
#pragma unroll
    for (int i = 0; i < 16; ++i) resp[i] = 0u;
    for (int l = 0; l < p.L; ++l) {
      mbar_wait_warp(smem_u32(acc_full), (uint32_t)(l & 1));
      tcgen05_fence_after();
      if (threadIdx.x == 64) TSTAMP(l, 4);
      uint32_t r[32], r2[32], r3[32];
      tmem_ld32(tcol, r);
      tmem_ld32(tcol + CH, r2);
      tmem_ld32(tcol + 2 * CH, r3);
      tmem_wait_ld();
      tcgen05_fence_before();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty)) : "memory");
      const bool relu = (l == 0) || (l & 1);
      const bool has_res = (l >= 2) && !(l & 1);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float a = __uint_as_float(r[i]) + __uint_as_float(r2[i]) + __uint_as_float(r3[i]) + s_bias[l * CH + c0 + i];
        v[i] = relu ? fmaxf(a, 0.f) : a;
      }
      if (has_res) {   // "+ A": this thread's own bf16 output of the last even layer, kept in registers
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&resp[i]));
          v[2 * i] += f.x;
          v[2 * i + 1] += f.y;
        }
      }
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        o[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      if (!(l & 1)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) resp[i] = o[i];
      }
      // staging tile = SWIZZLE_128B image of the output box: pixel m owns row m, its 64 B are chunks 4*chalf..+3 ^ (m & 7).
      // The halo buffer is free: this layer's MMAs have completed (acc_full).
      uint8_t* os = halo + (uint32_t)m * 128u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<uint4*>(os + ((((uint32_t)(4 * chalf + k)) ^ ((uint32_t)m & 7u)) << 4)) =
            make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (threadIdx.x == 64) TSTAMP(l, 5);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
      if (threadIdx.x == 64) {
        TSTAMP(l, 6);
        // one TMA store of the tile (clipped at the image border), complete before the layer counter is released:
        // the neighbours' TMA halo reads and our own next halo load (which overwrites the staging tile) wait on it
        tma_store_4d((l & 1) ? &tm_bo : &tm_ao, smem_u32(halo), 0, x0, y0, n);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        fence_proxy_async_global();
        st_release(my_flag, (unsigned int)(l + 1));
        TSTAMP(l, 7);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int make_map(PFN_encodeTiled enc, CUtensorMap* tm, const void* base, int N, int H, int W, int box_w = 10, int box_h = HALO_ROWS) {
  const cuuint64_t gdim[4] = {(cuuint64_t)CH, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)CH * 2, (cuuint64_t)W * CH * 2, (cuuint64_t)H * W * CH * 2};
  const cuuint32_t box[4] = {(cuuint32_t)CH, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return cr == CUDA_SUCCESS ? 0 : (int)cr;
}

}  // namespace

extern long long* teco_g_dbg_timing;   // conv_tc.cu

// 1 if teco_trunk64_tc can run this shape on the current device (single wave of co-resident CTAs), else 0.
extern "C" int teco_trunk64_supported(int32_t N, int32_t H, int32_t W, int32_t num_layers) {
  if (N <= 0 || H <= 0 || W <= 0 || num_layers < 1 || num_layers > MAX_LAYERS || (num_layers & 1) == 0) return 0;
  long long tiles = (long long)N * teco_ceil_div(H, TILE_ROWS) * teco_ceil_div(W, 8);
  return tiles <= teco_sm_count() ? 1 : 0;
}

extern "C" int teco_trunk64_tc(int32_t N, int32_t H, int32_t W, int32_t num_layers, const void* x_in, void* buf_a, void* buf_b,
                               const void* wpk_all, const float* bias_all, void* flags, void* stream) {
  TECO_CHECK_ARG(x_in && buf_a && buf_b && wpk_all && bias_all && flags, "teco_trunk64_tc: NULL argument");
  TECO_CHECK_ARG(teco_trunk64_supported(N, H, W, num_layers),
                 "teco_trunk64_tc: unsupported shape N=%d H=%d W=%d layers=%d (needs an odd layer count <= %d and "
                 "tiles <= SM count: all CTAs wait on each other and must be co-resident)", N, H, W, num_layers, MAX_LAYERS);
  TrunkParams p;
  p.N = N; p.H = H; p.W = W; p.L = num_layers;
  p.tiles_x = teco_ceil_div(W, 8);
  p.tiles_y = teco_ceil_div(H, TILE_ROWS);
  p.num_tiles = N * p.tiles_x * p.tiles_y;
  p.wpk = (const uint8_t*)wpk_all; p.bias = bias_all;
  p.a = (__nv_bfloat16*)buf_a; p.b = (__nv_bfloat16*)buf_b; p.flags = (unsigned int*)flags;
  p.dbg = teco_g_dbg_timing;
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    teco_set_error("teco_trunk64_tc: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return TECO_E_CUDA;
  }
  CUtensorMap tx, ta, tb, tao, tbo;
  int e1 = make_map(enc, &tx, x_in, N, H, W), e2 = make_map(enc, &ta, buf_a, N, H, W), e3 = make_map(enc, &tb, buf_b, N, H, W);
  e2 |= make_map(enc, &tao, buf_a, N, H, W, 8, TILE_ROWS);
  e3 |= make_map(enc, &tbo, buf_b, N, H, W, 8, TILE_ROWS);
  if (e1 || e2 || e3) {
    teco_set_error("teco_trunk64_tc: cuTensorMapEncodeTiled failed (%d %d %d)", e1, e2, e3);
    return TECO_E_CUDA;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TECO_CUDA_CALL(cudaMemsetAsync(flags, 0, sizeof(unsigned int) * (size_t)p.num_tiles, s));
  const size_t smem_bytes = 1024 + HALO_REGION + 2 * W_LAYER_BYTES + 16 * 8 + (size_t)num_layers * CH * sizeof(float);
  static bool attr = false;
  if (!attr) {
    TECO_CUDA_CALL(cudaFuncSetAttribute(trunk64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    attr = true;
  }
  TECO_CHECK_ARG(smem_bytes <= 226 * 1024, "teco_trunk64_tc: too many layers for shared memory");
  int max_blocks = 0;
  TECO_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, trunk64_tc_kernel, NUM_THREADS, smem_bytes));
  TECO_CHECK_ARG(max_blocks >= 1 && p.num_tiles <= max_blocks * teco_sm_count(), "teco_trunk64_tc: CTAs cannot all be resident");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.num_tiles);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, trunk64_tc_kernel, tx, ta, tb, tao, tbo, p);
  if (le != cudaSuccess) {
    teco_set_error("teco_trunk64_tc: launch failed: %s", cudaGetErrorString(le));
    return TECO_E_CUDA;
  }
  return TECO_OK;
}
