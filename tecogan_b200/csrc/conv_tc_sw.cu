// ONE-TILE-PER-CTA variant of the tcgen05 convolution (see conv_tc.cu for the design notes and the persistent variant).
// Used for single-wave launches (tiles <= SMs, e.g. the 128x128 generator trunk) and for the transposed convolutions:
// same-box A/B (profiles/conv_tc_r01_notes.md) showed this variant 1.3 us faster per single-wave layer and its
// two-CTAs-per-SM ring faster on the epilogue-heavy transposed conv, while the persistent kernel wins every multi-wave conv.
// bf16 3x3 convolution / 3x3-stride-2 transposed convolution on the 5th-gen tensor cores (sm_100a):
// TMA halo tiles -> shared memory -> tcgen05.mma (kind::f16, fp32 accumulators in TMEM) -> tcgen05.ld
// epilogue (bias, activation, residual, bf16 store; or fp32 "+bicubic, *2-1" output stage).
//
// Replaces, layer by layer, the cuDNN convolutions behind conv2()/conv2_tran() of the reference
// (lib/ops.py:35-56) as used by generator_F (lib/frvsr.py:44-88) and fnet (lib/frvsr.py:4-41).
//
// Implicit GEMM, no im2col buffer:
//   CTA tile  = 16 image rows x (8*J) pixels; every 16x8 sub-tile is one UMMA accumulator with M = 128.
//   A operand = NHWC bf16 activations, 64 channels = one 128-byte row per pixel.  Per 64-channel block the halo is
//               staged by THREE 4-D TMA boxes (64 ch, 8J px, 18 rows, 1 image), one per horizontal tap offset kx,
//               with SWIZZLE_128B.  Each box is the UMMA canonical K-major SW128 layout as it lands
//               (8 consecutive pixels = one 1024-byte swizzle atom, SBO = one box row).  Vertical taps ky are
//               descriptor start-address offsets of whole box rows (atom aligned), so 3 loads serve 9 taps.
//               TMA out-of-bounds zero fill implements TF 'SAME' padding.
//   B operand = weights pre-packed on the device as [cin/64][tap][cout][64 cin] bf16 in the same SW128 image,
//               streamed by 1-D bulk copies.  When a whole layer fits (64->64: 72 KB) the slabs are fetched once,
//               BEFORE the programmatic-dependent-launch wait, and multicast across a 4-CTA cluster.
//   D         = fp32 in TMEM, column block (sub-tile, phase) * Cout.
// Transposed conv (stride 2, TF 'SAME', y[i] = sum_j x[j] w[i-2j]) is the same loop with the nine taps routed to
// four sub-pixel phase accumulators and a 2x interleaving epilogue.
//
// Why SW128 and not the no-swizzle layout (round-1 measurement, profiles/conv_tc_r01_notes.md): with 16-byte core
// matrix rows every tcgen05.mma took ~250 cycles instead of ~32-48, and the 16-byte TMA rows ran at ~10 B/clk/SM.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue
// (one epilogue warp per scheduler is latency-bound: ~1000 clk per 32 channels; two per scheduler halve it).
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include "teco_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_ROWS = 16;
constexpr int HALO_ROWS = TILE_ROWS + 2;
constexpr int CB = 64;                 // channels per K block = one 128-byte swizzled row
constexpr int MAX_WST = 12;
constexpr int NUM_EPI_WARPS = 8;       // two warps per TMEM lane quarter, each taking half of the output channels
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;

struct TcParams {
  int N, H, W, Cin, Cout;             // Cout = channel pitch of y / res / bias (all output channels)
  int Ncta, nsplit, tiles_pad;        // output channels per CTA (UMMA N), Cout splits, tile count padded to the cluster size
  int tiles_x, tiles_y, J;
  int mode, act, out_f32_c;
  float post_scale, post_shift;
  int nblk, WST, TPS, KS;              // Cin/64, weight ring stages, taps per weight slab (1 or 3), K-split chains
  int HST, CS, mcast, num_tiles;       // halo stages, cluster size, resident+multicast weights, real tile count
  int late_trigger;                    // trigger the dependent launch after the MMAs instead of after the prologue
  uint32_t stage2_bytes;               // second staging region after the weights: residual tiles (conv) / ping-pong tile (tconv)
  int tma_out, tma_res;                // bf16 output / residual tiles travel through swizzled smem staging + TMA (Ncta == 64, conv)
  uint32_t copy_bytes, halo_stage_bytes, w_slab_bytes, tmem_cols;
  const uint8_t* wpk;
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  const float* res_f32;
  float* out_f32;
  long long* dbg;                      // optional [gridDim][32] clock64 stamps (teco_debug_timing)
};

using namespace tcptx;   // mbarrier / TMA / tcgen05 wrappers and the UMMA descriptors: tc_ptx.cuh

// ------------------------------------------------------------------ the kernel
// MODE 0 conv / 1 transposed conv; TPS taps per weight slab; J sub-tiles per CTA; KS K-split accumulator chains.
// They are compile-time so that the MMA issue loop is a fully unrolled stream of UTCHMMA whose descriptors differ
// from per-stage bases by immediates (uniform-datapath adds, no per-instruction R2UR).
// H1 = 1: ONE halo box (8J+2 pixels wide) per block instead of one box per horizontal tap; the horizontal taps are
// descriptor start offsets of whole 128-byte pixel rows (not swizzle-atom aligned: the SW128 XOR is a function of the
// absolute shared-memory address bits, which is also how TMA wrote the box; the descriptor's base-offset field stays 0 --
// setting it to (addr >> 7) & 7 was tested on the B200 and gives wrong results).  2.4x less L2->smem traffic, 23 KB per stage.
template <int MODE, int TPS, int J, int KS, int H1>
__global__ void __launch_bounds__(NUM_THREADS, 2)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_y,
                  const __grid_constant__ CUtensorMap tmap_r, const TcParams p) {
  // SWIZZLE_128B needs 1024-byte aligned tiles; the dynamic window starts on such a boundary (no static shared memory in
  // this kernel) -- relied upon instead of a 1 KB slack so that two trunk-layer CTAs (112.3 KB each) share an SM.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* halo_base = smem;                                             // HST stages x 3 kx-copies
  uint8_t* w_base = smem + (size_t)p.HST * p.halo_stage_bytes;           // WST weight slabs
  // residual staging tiles (J x 16 KB, TMA box image) follow the weights; the OUTPUT staging tiles reuse the halo
  // stages, which are dead once the accumulators are complete
  uint8_t* stage_res = w_base + (size_t)p.WST * p.w_slab_bytes;
  uint8_t* stage_out = halo_base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_res + p.stage2_bytes);
  uint64_t* halo_full = bars;
  uint64_t* halo_empty = bars + 2;
  uint64_t* w_full = bars + 4;
  uint64_t* w_empty = bars + 4 + MAX_WST;
  uint64_t* acc_full = bars + 4 + 2 * MAX_WST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * MAX_WST + 1);
  float* s_bias = reinterpret_cast<float*>(bars + 4 + 2 * MAX_WST + 2);   // [Cout]
  uint64_t* res_full = bars + 4 + 2 * MAX_WST + 2 + 128;                  // after 256 floats of bias

  // tile coordinates
  int tile = blockIdx.x % p.tiles_pad;      // grid = nsplit x (tiles padded to a multiple of the cluster size)
  const int n0 = (blockIdx.x / p.tiles_pad) * p.Ncta;   // first output channel of this CTA (all CTAs of a cluster share it)
  const bool active = tile < p.num_tiles;
  if (!active) tile = 0;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int n = tile / p.tiles_y;
  const int x0 = tx * 8 * J, y0 = ty * TILE_ROWS;
  long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 32 : nullptr;
#define STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
#define GSTAMP(i) do { if (dbg) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[i] = (long long)t_; } } while (0)
  if (threadIdx.x == 0) { STAMP(0); GSTAMP(26); }

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.HST; ++i) {
      mbar_init(smem_u32(&halo_full[i]), 1);
      mbar_init(smem_u32(&halo_empty[i]), 1);
    }
    for (int i = 0; i < p.WST; ++i) {
      mbar_init(smem_u32(&w_full[i]), 1);
      mbar_init(smem_u32(&w_empty[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(res_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    if (p.tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_y) : "memory");
    if (p.tma_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_r) : "memory");
  }
  if (warp == 1) {  // TMEM allocation (one full warp), result lands in smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  if (p.CS > 1) cluster_sync_all();   // every CTA's mbarriers are initialised before any multicast may signal them
  else __syncthreads();
  tcgen05_fence_after();
  // Programmatic dependent launch is triggered only once this CTA's own dependency wait has returned (producer warp):
  // the previous layer has retired by then, so the next layer's CTAs share an SM with exactly one CTA of this layer, run
  // their prologue + weight fetch under our halo load / MMAs / epilogue, and never pile up three layers deep (triggering
  // in the prologue let 2 CTAs of later layers land on the 20 idle SMs: 7.45 vs 6.5 us/layer).
  if (!active || !p.late_trigger) pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) STAMP(1);

  constexpr int nacc = MODE == 1 ? 4 : 1;
  constexpr int ncopies = H1 ? 1 : (MODE == 1 ? 2 : 3);   // horizontal tap offsets that occur (tconv only reads x-1, x)
  constexpr int slabs_per_blk = 9 / TPS;
  constexpr int row_bytes = (H1 ? 8 * J + 2 : 8 * J) * 128;   // one box row (pixels x 128 B)
  constexpr uint32_t copy_bytes = (uint32_t)(HALO_ROWS * row_bytes);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Issue cost of one bulk/tensor copy is a few hundred cycles, so independent copies are issued by different lanes.
    if (p.mcast) {
      // Weights do not depend on the previous layer: fetch them before the dependency wait.  Each CTA of the
      // cluster fetches 1/CS of every slab and multicasts it to all CS CTAs (one L2 read per cluster).
      const int sidx = lane;
      if (sidx < slabs_per_blk * p.nblk) {
        const uint32_t crank = p.CS > 1 ? cluster_ctarank() : 0;
        const uint32_t part = p.w_slab_bytes / (uint32_t)p.CS;
        const uint16_t mask = (uint16_t)((1u << p.CS) - 1u);
        const int b = sidx / slabs_per_blk, g = sidx - slabs_per_blk * b;
        mbar_expect_tx(smem_u32(&w_full[sidx]), p.w_slab_bytes);
        // global layout [blk][tap][cout][64]: a slab = TPS consecutive taps of one block (TPS == 3 only when nsplit == 1)
        const uint8_t* src = p.wpk + ((size_t)(b * 9 + g * TPS) * p.Cout + n0) * 128 + (size_t)crank * part;
        const uint32_t dst = smem_u32(w_base + (size_t)sidx * p.w_slab_bytes) + crank * part;
        if (p.CS > 1) bulk_load_1d_mcast(dst, src, part, smem_u32(&w_full[sidx]), mask);
        else bulk_load_1d(dst, src, part, smem_u32(&w_full[sidx]));
      }
      __syncwarp();
    }
    // Ring mode: the first WST weight slabs do not depend on the previous layer either -> issue them before the wait.
    const int total_slabs = slabs_per_blk * p.nblk;
    int next_slab = 0;
    if (!p.mcast && lane == 0) {
      for (; next_slab < total_slabs && next_slab < p.WST; ++next_slab) {
        const int b = next_slab / slabs_per_blk, g = next_slab - slabs_per_blk * b;
        mbar_expect_tx(smem_u32(&w_full[next_slab]), p.w_slab_bytes);
        const uint8_t* src = p.wpk + ((size_t)(b * 9 + g * TPS) * p.Cout + n0) * 128;
        bulk_load_1d(smem_u32(w_base + (size_t)next_slab * p.w_slab_bytes), src, p.w_slab_bytes, smem_u32(&w_full[next_slab]));
      }
    }
    if (lane == 0) STAMP(9);
    pdl_wait();   // the previous kernel's output (our input x) is complete and visible from here on
    if (p.late_trigger) pdl_launch_dependents();   // the layer before us has retired: at most two layers share an SM
    if (lane == 0) { STAMP(10); GSTAMP(27); }
    if (active) {
      int hs = 0;
      uint32_t hph = 0;
      for (int b = 0; b < p.nblk; ++b) {
        if (lane == 0) {
          mbar_wait(smem_u32(&halo_empty[hs]), hph ^ 1);
          mbar_expect_tx(smem_u32(&halo_full[hs]), copy_bytes * ncopies);
        }
        __syncwarp();
        if (lane < ncopies)
          tma_load_4d(smem_u32(halo_base + (size_t)hs * p.halo_stage_bytes + (size_t)lane * copy_bytes), &tmap,
                      smem_u32(&halo_full[hs]), b * CB, x0 - 1 + lane, y0 - 1, n);
        if (!p.mcast && lane == 0) {
          for (; next_slab < (b + 1) * slabs_per_blk; ++next_slab) {   // slabs of this block not yet in flight
            const int st = next_slab % p.WST, use = next_slab / p.WST;
            const int g = next_slab - slabs_per_blk * b;
            mbar_wait(smem_u32(&w_empty[st]), (uint32_t)((use - 1) & 1));   // the MMAs of the previous use have retired
            mbar_expect_tx(smem_u32(&w_full[st]), p.w_slab_bytes);
            const uint8_t* src = p.wpk + ((size_t)(b * 9 + g * TPS) * p.Cout + n0) * 128;
            bulk_load_1d(smem_u32(w_base + (size_t)st * p.w_slab_bytes), src, p.w_slab_bytes, smem_u32(&w_full[st]));
          }
        }
        __syncwarp();
        if (++hs == p.HST) { hs = 0; hph ^= 1; }
      }
      if (p.tma_res) {   // the residual tile (same box as the output tile), after the halo: the MMAs are waiting for that one
        if (lane == 0) mbar_expect_tx(smem_u32(res_full), (uint32_t)(J * 16384));
        __syncwarp();
        if (lane < J) tma_load_4d(smem_u32(stage_res + (size_t)lane * 16384), &tmap_r, smem_u32(res_full), n0, x0 + 8 * lane, y0, n);
      }
    } else if (p.mcast && lane == 0) {
      // padding CTA of a cluster: it only relays its share of the weights; stay until they have landed here too
      for (int sidx = 0; sidx < slabs_per_blk * p.nblk; ++sidx) mbar_wait(smem_u32(&w_full[sidx]), 0);
    }
    __syncwarp();
  } else if (warp == 1 && active) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc(p.Ncta);
    constexpr uint32_t a_sbo = (uint32_t)row_bytes;   // next 8-pixel group of the M=128 sub-tile = next image row
    constexpr uint32_t b_sbo = 1024u;                 // next 8 output channels
    int hs = 0, ws = 0;
    uint32_t hph = 0, wph = 0;
    uint32_t started = 0;  // bit (j*nacc+phase): accumulator already written once
    for (int b = 0; b < p.nblk; ++b) {
      mbar_wait_warp(smem_u32(&halo_full[hs]), hph);
      tcgen05_fence_after();
      if (lane == 0 && b == 0) { STAMP(2); GSTAMP(28); }
      const uint32_t halo_addr = smem_u32(halo_base + (size_t)hs * p.halo_stage_bytes);
      for (int g = 0; g < slabs_per_blk; ++g) {
        mbar_wait_warp(smem_u32(&w_full[ws]), wph);
        tcgen05_fence_after();
        if (lane == 0 && b == 0) STAMP(16 + g);
        {
          // Whole (converged) warp computes the warp-uniform bases; one elected lane issues the unrolled MMA stream.
          const uint32_t slab_addr = smem_u32(w_base + (size_t)ws * p.w_slab_bytes);
          const uint64_t a_base = umma_desc_sw128(halo_addr, a_sbo);
          const uint64_t b_base = umma_desc_sw128(slab_addr, b_sbo);
          const uint32_t tap_stride16 = (uint32_t)(p.Ncta * 128) >> 4;   // weight bytes per tap, in descriptor units
          // per-tap row/copy/phase (TPS == 3: ky = g, kx = tt; TPS == 1: tap = g)
          uint32_t a_off16[TPS], acc_idx[TPS];
#pragma unroll
          for (int tt = 0; tt < TPS; ++tt) {
            const int t = g * TPS + tt;
            const int ky = (TPS == 3) ? g : t / 3, kx = (TPS == 3) ? tt : t - 3 * (t / 3);
            int ry, rx, phase;
            if (MODE == 1) {  // transposed conv: tap -> (input offset, output phase)
              ry = (ky == 2) ? 0 : 1;
              rx = (kx == 2) ? 0 : 1;
              phase = ((ky == 1) ? 2 : 0) + ((kx == 1) ? 1 : 0);
            } else {
              ry = ky; rx = kx; phase = 0;
            }
            a_off16[tt] = H1 ? ((uint32_t)(rx * 128 + ry * row_bytes)) >> 4
                             : ((uint32_t)rx * copy_bytes + (uint32_t)(ry * row_bytes)) >> 4;
            acc_idx[tt] = (uint32_t)(phase * KS + (KS > 1 ? tt : 0));
          }
          const uint32_t started_now = started;
          if (elect_one()) {
            // order: k-step, tap, sub-tile -> consecutive MMAs hit different accumulators (J sub-tiles x KS chains)
#pragma unroll
            for (int s = 0; s < CB / 16; ++s) {
#pragma unroll
              for (int tt = 0; tt < TPS; ++tt) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                  const uint32_t acc = (uint32_t)(j * nacc * KS) + acc_idx[tt];
                  uint32_t accum = 1u;
                  if (s == 0) {   // first k-step of this slab: overwrite only if nobody has written this accumulator yet
                    accum = (started_now >> acc) & 1u;
#pragma unroll
                    for (int t2 = 0; t2 < tt; ++t2) accum |= (acc_idx[t2] == acc_idx[tt]) ? 1u : 0u;
                  }
                  umma_bf16(tmem_base + acc * (uint32_t)p.Ncta,
                            a_base + a_off16[tt] + (uint32_t)((j * 1024 + s * 32) >> 4),
                            b_base + tt * tap_stride16 + (uint32_t)((s * 32) >> 4), idesc, accum);
                }
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int tt = 0; tt < TPS; ++tt) started |= 1u << ((uint32_t)(j * nacc * KS) + acc_idx[tt]);
          if (elect_one()) tcgen05_commit(smem_u32(&w_empty[ws]));  // frees the weight slab when these MMAs retire
        }
        __syncwarp();
        if (++ws == p.WST) { ws = 0; wph ^= 1; }
      }
      if (elect_one()) tcgen05_commit(smem_u32(&halo_empty[hs]));
      __syncwarp();
      if (++hs == p.HST) { hs = 0; hph ^= 1; }
    }
    if (elect_one()) tcgen05_commit(smem_u32(acc_full));
    if (lane == 0) { STAMP(5); GSTAMP(29); }
    __syncwarp();
  } else if (warp >= 2 && active) {
    // ===================== epilogue (warps 2..5) =====================
    for (int c = (int)threadIdx.x - 64; c < p.Ncta; c += 32 * NUM_EPI_WARPS) s_bias[c] = p.bias ? p.bias[n0 + c] : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");   // the epilogue warps only
    pdl_wait();                                        // res / y belong to the dependency chain
    if (threadIdx.x == 64) STAMP(25);
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2; // which half of the channel steps this warp takes (0 or 1)
    const int m = 32 * q + lane;       // accumulator row = pixel within the 16x8 sub-tile
    const int ry = m >> 3, rx = m & 7;
    const float act_slope = p.act == TECO_ACT_RELU ? 0.f : (p.act == TECO_ACT_LRELU02 ? 0.2f : 1.f);
    mbar_wait_warp(smem_u32(acc_full), 0);
    tcgen05_fence_after();
    if (threadIdx.x == 64) STAMP(6);
    const int oy_in = y0 + ry;
    // EW output channels per step: 32 (two steps for 64 channels) or 16 (the 16-channel fp32 output stage)
    auto run = [&](auto ew_tag) {
      constexpr int EW = decltype(ew_tag)::value;
      for (int j = 0; j < J; ++j) {
        const int ox_in = x0 + 8 * j + rx;
        const bool in_img = (oy_in < p.H) && (ox_in < p.W);
        for (int ph = 0; ph < nacc; ++ph) {
          int oy, ox, OH, OW;
          if (MODE == 1) {
            oy = 2 * oy_in + (ph >> 1); ox = 2 * ox_in + (ph & 1); OH = 2 * p.H; OW = 2 * p.W;
          } else {
            oy = oy_in; ox = ox_in; OH = p.H; OW = p.W;
          }
          const size_t pix = ((size_t)n * OH + oy) * OW + ox;
          const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((j * nacc + ph) * KS * p.Ncta);
          for (int c0 = chalf * EW; c0 < p.Ncta; c0 += 2 * EW) {
            uint32_t r[EW];
            __syncwarp();
            if (EW == 32) tmem_ld32(tcol + (uint32_t)c0, r); else tmem_ld16(tcol + (uint32_t)c0, r);
            if (KS == 3) {   // K-split chains: issue all three TMEM loads, wait once, add
              uint32_t r2[EW], r3[EW];
              if (EW == 32) { tmem_ld32(tcol + (uint32_t)(p.Ncta + c0), r2); tmem_ld32(tcol + (uint32_t)(2 * p.Ncta + c0), r3); }
              else { tmem_ld16(tcol + (uint32_t)(p.Ncta + c0), r2); tmem_ld16(tcol + (uint32_t)(2 * p.Ncta + c0), r3); }
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < EW; ++i)
                r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]) + __uint_as_float(r3[i]));
            } else {
              tmem_wait_ld();
            }
            if (threadIdx.x == 64 && j == 0 && ph == 0 && c0 == 0) STAMP(11);
            float v[EW];
#pragma unroll
            for (int i = 0; i < EW; ++i) {
              const float a = __uint_as_float(r[i]) + s_bias[c0 + i];
              v[i] = fmaxf(a, a * act_slope);      // none: slope 1, relu: 0, lrelu: 0.2 -- no per-element branch
            }
            if (p.act >= TECO_ACT_TANH24) {        // uniform, outside the element loop
#pragma unroll
              for (int i = 0; i < EW; ++i) v[i] = teco_act(v[i], p.act);
            }
            if (MODE == 0 && EW == 32 && p.tma_out) {
              // c0 == 32 * chalf here (Ncta == 64).  Pixel m owns row m of the staging tile; its 64 B are chunks 4*chalf..+3,
              // XOR-swizzled with (m & 7) exactly like the TMA box -> conflict-free 16-byte accesses.
              const uint32_t rowoff = (uint32_t)m * 128u;
              if (p.tma_res) {
                mbar_wait_warp(smem_u32(res_full), 0);
                const uint8_t* rs = stage_res + (size_t)j * 16384 + rowoff;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(rs + ((((uint32_t)(4 * chalf + k)) ^ ((uint32_t)m & 7u)) << 4));
                  const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[i]));
                    v[8 * k + 2 * i] += f.x;
                    v[8 * k + 2 * i + 1] += f.y;
                  }
                }
              }
              uint8_t* os = stage_out + (size_t)j * 16384 + rowoff;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                  o[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(os + ((((uint32_t)(4 * chalf + k)) ^ ((uint32_t)m & 7u)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              }
              continue;
            }
            if (MODE == 1 && EW == 32 && p.tma_out) {
              // transposed conv: one staging tile per (sub-tile, output phase), two tiles in ping-pong (the dead halo stage and
              // the region after the weights); the phase's 16x8 pixels go out through a 5-D map over the 2x interleaved output
              const int t = j * 4 + ph;
              uint8_t* stg = (t & 1) ? stage_res : stage_out;
              if (t >= 2) {   // the store issued two steps ago has finished reading this tile
                if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
              }
              uint8_t* os = stg + (uint32_t)m * 128u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                  o[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(os + ((((uint32_t)(4 * chalf + k)) ^ ((uint32_t)m & 7u)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              }
              fence_async_smem();
              asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
              if (threadIdx.x == 64) {
                tma_store_5d(&tmap_y, smem_u32(stg), n0, ph & 1, x0 + 8 * j, ph >> 1, n * p.H + y0);
                bulk_commit();
              }
              continue;
            }
            if (!in_img) continue;
            if (p.out_f32) {
              for (int i = 0; i < EW; ++i) {
                int c = n0 + c0 + i;
                if (c < p.out_f32_c) {
                  float a = v[i] + (p.res_f32 ? p.res_f32[pix * p.out_f32_c + c] : 0.f);
                  p.out_f32[pix * p.out_f32_c + c] = a * p.post_scale + p.post_shift;
                }
              }
            }
            if (p.y) {
              if (p.res) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.Cout + n0 + c0);
#pragma unroll
                for (int k = 0; k < EW / 8; ++k) {
                  const uint4 rr = rp[k];
                  const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[i]));
                    v[8 * k + 2 * i] += f.x;
                    v[8 * k + 2 * i + 1] += f.y;
                  }
                }
              }
              uint4* yp = reinterpret_cast<uint4*>(p.y + pix * p.Cout + n0 + c0);
#pragma unroll
              for (int k = 0; k < EW / 8; ++k) {
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                  o[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                yp[k] = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
            if (threadIdx.x == 64 && j == 0 && ph == 0) STAMP(12 + ((c0 / EW) & 3));
          }
        }
      }
    };
    if (p.Ncta % 32 == 0) run(std::integral_constant<int, 32>{});
    else run(std::integral_constant<int, 16>{});
    if (MODE == 1 && p.tma_out && threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (MODE == 0 && p.tma_out) {
      // generic-proxy smem writes -> visible to the async proxy, all epilogue warps done, then one thread stores the tiles
      fence_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
      if (threadIdx.x == 64) {
        for (int j = 0; j < J; ++j) tma_store_4d(&tmap_y, smem_u32(stage_out + (size_t)j * 16384), n0, x0 + 8 * j, y0, n);
        bulk_commit();
        // shared memory must outlive the store's reads; global visibility to the next layer comes with grid completion
        // (griddepcontrol.wait / stream order), as in any epilogue that ends with a TMA store
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
  }

  if (threadIdx.x == 64) { STAMP(7); GSTAMP(30); }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) STAMP(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
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

}  // namespace

extern long long* teco_g_dbg_timing;   // set by teco_debug_timing (conv_tc.cu)

int teco_conv3x3_tc_one_tile(const teco_tc_desc* d, const void* x, const void* wpk, const float* bias, const void* res,
                               void* y, const float* res_f32, float* out_f32, void* stream) {
  TECO_CHECK_ARG(d && x && wpk, "teco_conv3x3_tc(one-tile): NULL argument");
  TECO_CHECK_ARG(y || out_f32, "teco_conv3x3_tc(one-tile): no output buffer");
  TECO_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "teco_conv3x3_tc(one-tile): bad shape N=%d H=%d W=%d", d->N, d->H, d->W);
  TECO_CHECK_ARG(d->Cin >= 64 && d->Cin % 64 == 0 && d->Cin <= 512, "teco_conv3x3_tc(one-tile): Cin must be a multiple of 64 in [64,512] (got %d)", d->Cin);
  TECO_CHECK_ARG(d->Cout >= 16 && d->Cout % 16 == 0 && d->Cout <= 512, "teco_conv3x3_tc(one-tile): Cout must be a multiple of 16 in [16,512] (got %d)", d->Cout);
  TECO_CHECK_ARG(d->mode == 0 || d->mode == 1, "teco_conv3x3_tc(one-tile): unknown mode %d", d->mode);
  TECO_CHECK_ARG(d->act >= 0 && d->act <= TECO_ACT_SIGMOID, "teco_conv3x3_tc(one-tile): unknown activation %d", d->act);
  TECO_CHECK_ARG(!out_f32 || (d->out_f32_c > 0 && d->out_f32_c <= d->Cout), "teco_conv3x3_tc(one-tile): bad out_f32_c");
  TECO_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)wpk) & 15) == 0 && (((uintptr_t)y) & 15) == 0 &&
                     (((uintptr_t)res) & 15) == 0,
                 "teco_conv3x3_tc(one-tile): tensors must be 16-byte aligned");

  TcParams p;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.mode = d->mode; p.act = d->act; p.out_f32_c = d->out_f32_c;
  p.post_scale = d->post_scale; p.post_shift = d->post_shift;
  p.wpk = (const uint8_t*)wpk; p.bias = bias; p.res = (const __nv_bfloat16*)res; p.y = (__nv_bfloat16*)y;
  p.res_f32 = res_f32; p.out_f32 = out_f32; p.dbg = teco_g_dbg_timing;
  if (p.dbg) {   // consecutive launches stamp consecutive [256][32] slices (tools/bench_conv.py chain)
    static long long* last_base = nullptr;
    static int launch_idx = 0;
    if (last_base != teco_g_dbg_timing) { last_base = teco_g_dbg_timing; launch_idx = 0; }
    p.dbg = teco_g_dbg_timing + (size_t)(launch_idx % 8) * 256 * 32;
    ++launch_idx;
  }
  p.nblk = d->Cin / CB;
  p.HST = p.nblk > 1 ? 2 : 1;
  constexpr int H1 = 1;   // single halo box per block, horizontal taps as 128-byte descriptor offsets (the three-box variant lost the A/B and is gone)
  // few spatial tiles but many output channels (FNet's 16x16 / 32x32 layers): split Cout over CTAs, 64 channels each
  {
    long long t1 = (long long)d->N * teco_ceil_div(d->H, TILE_ROWS) * teco_ceil_div(d->W, 8);
    p.nsplit = (d->Cout >= 128 && d->Cout % 64 == 0 && t1 * (d->Cout / 64) <= 2LL * teco_sm_count()) ? d->Cout / 64 : 1;
    if (d->Cout / p.nsplit > 256) p.nsplit = d->Cout / 256;
    p.Ncta = d->Cout / p.nsplit;
  }
  const int nacc = d->mode == 1 ? 4 : 1;
  const size_t budget = 208 * 1024;
  const int sms = teco_sm_count();
  // sub-tiles per CTA: prefer the wider tile when it still yields >= 2 waves of CTAs and fits TMEM / smem
  int J = 1;
  for (int j = 2; j >= 1; --j) {
    if (j * nacc * p.Ncta > 512) continue;
    size_t a_bytes = H1 ? (size_t)p.HST * (((size_t)HALO_ROWS * (8 * j + 2) * 128 + 1023) & ~(size_t)1023)
                        : (size_t)p.HST * 3 * HALO_ROWS * 8 * j * 128;
    if (a_bytes + 2 * (size_t)p.Ncta * 128 > budget) continue;
    long long tiles = (long long)d->N * teco_ceil_div(d->H, TILE_ROWS) * teco_ceil_div(d->W, 8 * j);
    static const int env_j = [] { const char* e = getenv("TECO_TC_J"); return e ? atoi(e) : 0; }();
    if (tiles >= 2LL * sms || j == 1 || j == env_j) { J = j; break; }
  }
  TECO_CHECK_ARG(J * nacc * p.Ncta <= 512, "teco_conv3x3_tc(one-tile): Cout=%d too large for mode %d (TMEM has 512 columns)", d->Cout, d->mode);
  p.J = J;
  p.tiles_x = teco_ceil_div(d->W, 8 * J);
  p.tiles_y = teco_ceil_div(d->H, TILE_ROWS);
  p.copy_bytes = (uint32_t)(HALO_ROWS * (H1 ? 8 * J + 2 : 8 * J) * 128);
  p.halo_stage_bytes = H1 ? ((p.copy_bytes + 1023u) & ~1023u) : 3 * p.copy_bytes;
  const size_t a_total = (size_t)p.HST * p.halo_stage_bytes;
  const size_t tap_bytes = (size_t)p.Ncta * 128;
  // whole layer resident?  then 3 taps per slab (3 barriers per block), fetched once, multicast over a 4-CTA cluster
  p.num_tiles = (int)((long long)d->N * p.tiles_x * p.tiles_y);
  // Weight staging.  Preferred: a 2-deep ring of 3-tap slabs -- with the halo copies that is ~105 KB, so TWO CTAs fit per SM
  // and programmatic dependent launch really overlaps the next layer's prologue + weight prefetch with this layer's
  // MMA/epilogue.  (A whole resident layer, 128 KB, multicast over a cluster, serialised the layers: round-1 notes.)
  const size_t half_sm = 112 * 1024;
  p.mcast = 0;
  const bool single_wave = (long long)p.num_tiles * p.nsplit <= (long long)sms;
  if (single_wave && p.nsplit == 1 && a_total + 9 * tap_bytes * p.nblk <= budget && 3 * p.nblk <= MAX_WST) {
    // one CTA per SM anyway (e.g. the 128x128 trunk: 128 tiles): whole layer resident, fetched before the dependency
    // wait and multicast over a 4-CTA cluster -- measured 6.2 us vs 6.9 us for the ring on the 64->64 layer
    p.mcast = 1; p.TPS = 3; p.WST = 3 * p.nblk;
  } else if (p.nsplit == 1 && 1024 + a_total + 2 * 3 * tap_bytes + 1024 <= half_sm) {
    // multi-wave grids: 2-deep ring of 3-tap slabs (~105 KB) so TWO CTAs share an SM (256x256: 14.7 us vs 24.0 us)
    p.TPS = 3; p.WST = 2;
  } else if (p.nsplit == 1 && a_total + 2 * 3 * tap_bytes <= budget) {
    p.TPS = 3; p.WST = (int)((budget - a_total) / (3 * tap_bytes));
    if (p.WST > 3 * p.nblk) p.WST = 3 * p.nblk;
    if (p.WST > MAX_WST) p.WST = MAX_WST;
  } else {
    p.TPS = 1;
    int wst = (int)((budget - a_total) / tap_bytes);
    if (wst > 9 * p.nblk) wst = 9 * p.nblk;
    if (wst > MAX_WST) wst = MAX_WST;
    TECO_CHECK_ARG(wst >= 2, "teco_conv3x3_tc(one-tile): shared memory budget too small (Cin=%d Cout=%d)", d->Cin, d->Cout);
    p.WST = wst;
  }
  p.w_slab_bytes = (uint32_t)(p.TPS * tap_bytes);
  p.KS = (p.TPS == 3 && d->mode == 0 && J * 3 * p.Ncta <= 512) ? 3 : 1;
  p.CS = (p.mcast && p.num_tiles >= 8) ? 4 : 1;
  uint32_t cols = (uint32_t)(J * nacc * p.KS * p.Ncta), tc = 32;
  while (tc < cols) tc <<= 1;
  p.tmem_cols = tc;
  constexpr int env_tma = 1;   // staged (TMA-store) epilogue wherever the layer qualifies
  p.tma_out = (env_tma && y && !out_f32 && p.Ncta == 64 &&
               (d->mode == 0 || (!res && p.nsplit == 1 && d->H % TILE_ROWS == 0))) ? 1 : 0;   // tconv: rows of (n, y) are one map dimension
  p.tma_res = (p.tma_out && res) ? 1 : 0;
  p.stage2_bytes = p.tma_res ? (uint32_t)(J * 16384) : ((p.tma_out && d->mode == 1) ? 16384u : 0u);
  TECO_CHECK_ARG(!p.tma_out || (size_t)J * 16384 <= a_total, "teco_conv3x3_tc(one-tile): output staging does not fit the halo stages");
  const size_t smem_bytes = a_total + (size_t)p.WST * p.w_slab_bytes + p.stage2_bytes +
                            (4 + 2 * MAX_WST + 2) * 8 + 256 * sizeof(float) + 16;
  static const int env_late = [] { const char* e = getenv("TECO_TC_LATE_TRIGGER"); return e ? atoi(e) : 1; }();
  p.late_trigger = (env_late && 2 * (smem_bytes + 1024) <= 228 * 1024) ? 1 : 0;   // only useful when two CTAs fit an SM
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    teco_set_error("teco_conv3x3_tc(one-tile): cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return TECO_E_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
  const cuuint64_t gstr[3] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->W * d->Cin * 2, (cuuint64_t)d->H * d->W * d->Cin * 2};
  const cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)(H1 ? 8 * J + 2 : 8 * J), (cuuint32_t)HALO_ROWS, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    teco_set_error("teco_conv3x3_tc(one-tile): cuTensorMapEncodeTiled failed with CUresult %d (N=%d H=%d W=%d Cin=%d)", (int)cr, d->N,
                   d->H, d->W, d->Cin);
    return TECO_E_CUDA;
  }
  CUtensorMap tmap_y = tmap, tmap_r = tmap;   // placeholders when the staged epilogue is off
  if (p.tma_out && d->mode == 1) {
    // output [N, 2H, 2W, C] as {C, px, x, py, n*H + y}: one box = the 16x8 pixels of one sub-pixel phase
    const cuuint64_t odim[5] = {(cuuint64_t)d->Cout, 2, (cuuint64_t)d->W, 2, (cuuint64_t)d->N * d->H};
    const cuuint64_t ostr[4] = {(cuuint64_t)d->Cout * 2, (cuuint64_t)d->Cout * 4, (cuuint64_t)d->W * d->Cout * 4,
                                (cuuint64_t)d->W * d->Cout * 8};
    const cuuint32_t obox[5] = {64, 1, 8, 1, (cuuint32_t)TILE_ROWS};
    const cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
    cr = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, y, odim, ostr, obox, estr5, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      teco_set_error("teco_conv3x3_tc(one-tile): cuTensorMapEncodeTiled (transposed-conv output) failed with CUresult %d", (int)cr);
      return TECO_E_CUDA;
    }
  } else if (p.tma_out) {
    const cuuint64_t odim[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    const cuuint64_t ostr[3] = {(cuuint64_t)d->Cout * 2, (cuuint64_t)d->W * d->Cout * 2, (cuuint64_t)d->H * d->W * d->Cout * 2};
    const cuuint32_t obox[4] = {64, 8, (cuuint32_t)TILE_ROWS, 1};
    cr = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, odim, ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS && p.tma_res)
      cr = enc(&tmap_r, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(res), odim, ostr, obox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      teco_set_error("teco_conv3x3_tc(one-tile): cuTensorMapEncodeTiled (output tile) failed with CUresult %d", (int)cr);
      return TECO_E_CUDA;
    }
  }
  using KernelT = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
  KernelT kern = nullptr;
#define TECO_PICK(M, T, JJ, K)                                              \
  if (d->mode == M && p.TPS == T && J == JJ && p.KS == K)                   \
    kern = conv3x3_tc_kernel<M, T, JJ, K, 1>;
  TECO_PICK(0, 3, 1, 3) TECO_PICK(0, 3, 2, 3) TECO_PICK(0, 3, 1, 1) TECO_PICK(0, 3, 2, 1) TECO_PICK(0, 1, 1, 1) TECO_PICK(0, 1, 2, 1)
  TECO_PICK(1, 3, 1, 1) TECO_PICK(1, 3, 2, 1) TECO_PICK(1, 1, 1, 1) TECO_PICK(1, 1, 2, 1)
#undef TECO_PICK
  if (!kern) {
    teco_set_error("teco_conv3x3_tc(one-tile): no kernel instantiation for mode=%d TPS=%d J=%d KS=%d", d->mode, p.TPS, J, p.KS);
    return TECO_E_UNSUPPORTED;
  }
  TECO_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024)));
  p.tiles_pad = (p.num_tiles + p.CS - 1) / p.CS * p.CS;
  const unsigned ctas = (unsigned)(p.tiles_pad * p.nsplit);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
  attrs[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (p.CS > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = (unsigned)p.CS;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmap, tmap_y, tmap_r, p);
  if (le != cudaSuccess) {
    teco_set_error("teco_conv3x3_tc(one-tile): launch failed: %s (grid %u, cluster %d, smem %zu)", cudaGetErrorString(le), ctas, p.CS, smem_bytes);
    return TECO_E_CUDA;
  }
  return TECO_OK;
}
