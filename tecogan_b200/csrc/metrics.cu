// metrics.cu -- PSNR and SSIM on the Y channel of uint8 RGB frames, on the device.
// Reference: metrics.py:37-55 (_rgb2ycbcr, maxVal 255), :57-61 (to_uint8 -- the identity on decoded PNGs), :63-70 (psnr),
// :72-75 (ssim = skimage.measure.compare_ssim(Y_true, Y_pred, data_range = Y_pred.max() - Y_pred.min()); skimage <= 0.15
// defaults: 7x7 uniform window, sample covariance, K1 0.01, K2 0.03, mean of S over the image cropped by 3 pixels), after
// the crop of metrics.py:77-92 (crop_8x8), which the caller passes as a window so that no cropped copy is made.
// The reference does this arithmetic in float64 (numpy); so do these kernels -- the frames are a few MB and the work is
// bounded by reading them (6 bytes per pixel), not by the FP64 pipe.
#include "teco_common.cuh"

namespace {
constexpr int TPB = 256;
constexpr int WIN = 7;                            // compare_ssim default win_size; S is averaged over the image cropped by 3
constexpr int TILE = 32, HALO = TILE + WIN - 1;  // 32 x 32 SSIM values per CTA from a 38 x 38 window of Y pairs

__device__ __forceinline__ double y_of_rgb(const uint8_t* __restrict__ p) {
  // row 0 of T plus offset 16 (metrics.py:39-44); summed in the order numpy's dot uses (r, g, b)
  return ((0.256788235294118 * (double)p[0] + 0.504129411764706 * (double)p[1]) + 0.097905882352941 * (double)p[2]) + 16.0;
}

// Positive finite doubles order like their bit patterns: max via atomicMax on the bits, min via atomicMax on the
// complemented bits, so that an all-zero accumulator is the neutral element of all four slots.
__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }

struct Frames {
  const uint8_t* tgt;
  const uint8_t* out;
  int tH, tW, oH, oW;      // the two frame sizes may differ (metrics.py:134-135: the output is cut to the target)
  int y0, x0, h, w;        // crop window, same offsets in both frames
};

// acc[n] = { sum (Y_true - Y_pred)^2, ~bits(min Y_pred), bits(max Y_pred), sum of S } as four 8-byte slots
__global__ void __launch_bounds__(TPB) psnr_minmax_kernel(Frames f, double* __restrict__ acc) {
  const int n = blockIdx.y;
  const uint8_t* t = f.tgt + (size_t)n * f.tH * f.tW * 3;
  const uint8_t* o = f.out + (size_t)n * f.oH * f.oW * 3;
  const long long total = (long long)f.h * f.w;
  double ss = 0.0, mn = 1e300, mx = 0.0;
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < total; i += (long long)gridDim.x * TPB) {
    const int r = (int)(i / f.w), c = (int)(i - (long long)r * f.w);
    const double yt = y_of_rgb(t + ((size_t)(f.y0 + r) * f.tW + f.x0 + c) * 3);
    const double yp = y_of_rgb(o + ((size_t)(f.y0 + r) * f.oW + f.x0 + c) * 3);
    const double d = yt - yp;
    ss += d * d;
    mn = fmin(mn, yp);
    mx = fmax(mx, yp);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, s);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  }
  __shared__ double red[3][TPB / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = ss; red[1][wid] = mn; red[2][wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < TPB / 32; ++k) { ss += red[0][k]; mn = fmin(mn, red[1][k]); mx = fmax(mx, red[2][k]); }
    atomicAdd(&acc[n * 4 + 0], ss);
    unsigned long long* u = reinterpret_cast<unsigned long long*>(acc + n * 4);
    if (mn < 1e299) atomicMax(u + 1, ~dbits(mn));
    atomicMax(u + 2, dbits(mx));
  }
}

__global__ void __launch_bounds__(TPB) ssim_kernel(Frames f, double* __restrict__ acc) {
  __shared__ double sx[HALO * HALO], sy[HALO * HALO];
  __shared__ double red[TPB / 32];
  const int n = blockIdx.z;
  const uint8_t* t = f.tgt + (size_t)n * f.tH * f.tW * 3;
  const uint8_t* o = f.out + (size_t)n * f.oH * f.oW * 3;
  const int oy0 = blockIdx.y * TILE, ox0 = blockIdx.x * TILE;   // first SSIM position of this tile (crop coordinates)
  const int vh = f.h - (WIN - 1), vw = f.w - (WIN - 1);         // positions whose whole 7x7 window lies inside the crop
  for (int i = threadIdx.x; i < HALO * HALO; i += TPB) {
    const int r = min(oy0 + i / HALO, f.h - 1), c = min(ox0 + i % HALO, f.w - 1);   // clamped reads feed masked outputs only
    sx[i] = y_of_rgb(t + ((size_t)(f.y0 + r) * f.tW + f.x0 + c) * 3);
    sy[i] = y_of_rgb(o + ((size_t)(f.y0 + r) * f.oW + f.x0 + c) * 3);
  }
  // data_range of compare_ssim: Y_pred.max() - Y_pred.min() over the crop (first pass)
  const unsigned long long* u = reinterpret_cast<const unsigned long long*>(acc + n * 4);
  const double R = __longlong_as_double((long long)u[2]) - __longlong_as_double((long long)~u[1]);
  const double C1 = (0.01 * R) * (0.01 * R), C2 = (0.03 * R) * (0.03 * R);
  const double inv_np = 1.0 / (double)(WIN * WIN), cov_norm = (double)(WIN * WIN) / (double)(WIN * WIN - 1);
  __syncthreads();
  const int tx = threadIdx.x & (TILE - 1), ty0 = threadIdx.x >> 5;
  double part = 0.0;
  for (int ty = ty0; ty < TILE; ty += TPB / TILE) {
    if (oy0 + ty >= vh || ox0 + tx >= vw) continue;
    double ax = 0, ay = 0, axx = 0, ayy = 0, axy = 0;
#pragma unroll
    for (int dy = 0; dy < WIN; ++dy) {
#pragma unroll
      for (int dx = 0; dx < WIN; ++dx) {
        const double x = sx[(ty + dy) * HALO + tx + dx], y = sy[(ty + dy) * HALO + tx + dx];
        ax += x; ay += y; axx += x * x; ayy += y * y; axy += x * y;
      }
    }
    const double ux = ax * inv_np, uy = ay * inv_np;
    const double vx = cov_norm * (axx * inv_np - ux * ux), vy = cov_norm * (ayy * inv_np - uy * uy);
    const double vxy = cov_norm * (axy * inv_np - ux * uy);
    const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
    part += (A1 * A2) / (B1 * B2);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < TPB / 32; ++k) part += red[k];
    atomicAdd(&acc[n * 4 + 3], part);
  }
}

int check_frames(const char* who, const Frames& f, int N) {
  TECO_CHECK_ARG(f.tgt && f.out && N > 0 && N <= 65535, "%s: NULL frame pointer or bad frame count %d", who, N);
  TECO_CHECK_ARG(f.tH > 0 && f.tW > 0 && f.oH > 0 && f.oW > 0, "%s: bad frame size %dx%d / %dx%d", who, f.tH, f.tW, f.oH, f.oW);
  TECO_CHECK_ARG(f.y0 >= 0 && f.x0 >= 0 && f.h > 0 && f.w > 0 && f.y0 + f.h <= min(f.tH, f.oH) && f.x0 + f.w <= min(f.tW, f.oW),
                 "%s: crop window %d+%d x %d+%d does not fit frames %dx%d / %dx%d", who, f.y0, f.h, f.x0, f.w, f.tH, f.tW, f.oH, f.oW);
  return TECO_OK;
}
}  // namespace

extern "C" {

int teco_metrics_psnr_y_u8(const uint8_t* tgt, int32_t tH, int32_t tW, const uint8_t* out, int32_t oH, int32_t oW, int32_t N,
                           int32_t y0, int32_t x0, int32_t h, int32_t w, double* acc, void* stream) {
  Frames f{tgt, out, tH, tW, oH, oW, y0, x0, h, w};
  int rc = check_frames("teco_metrics_psnr_y_u8", f, N);
  if (rc != TECO_OK) return rc;
  TECO_CHECK_ARG(acc != nullptr, "teco_metrics_psnr_y_u8: NULL accumulator");
  cudaStream_t st = (cudaStream_t)stream;
  TECO_CUDA_CALL(cudaMemsetAsync(acc, 0, sizeof(double) * 4 * (size_t)N, st));
  const long long total = (long long)h * w;
  const int bx = (int)max(1LL, min((total + TPB * 4 - 1) / (TPB * 4), (long long)(4 * teco_sm_count())));
  psnr_minmax_kernel<<<dim3(bx, N), TPB, 0, st>>>(f, acc);
  TECO_CUDA_LAUNCH_CHECK("teco_metrics_psnr_y_u8");
  return TECO_OK;
}

int teco_metrics_ssim_y_u8(const uint8_t* tgt, int32_t tH, int32_t tW, const uint8_t* out, int32_t oH, int32_t oW, int32_t N,
                           int32_t y0, int32_t x0, int32_t h, int32_t w, double* acc, void* stream) {
  Frames f{tgt, out, tH, tW, oH, oW, y0, x0, h, w};
  int rc = check_frames("teco_metrics_ssim_y_u8", f, N);
  if (rc != TECO_OK) return rc;
  TECO_CHECK_ARG(acc != nullptr, "teco_metrics_ssim_y_u8: NULL accumulator");
  // compare_ssim raises ValueError when win_size exceeds the image extent
  TECO_CHECK_ARG(h >= WIN && w >= WIN, "teco_metrics_ssim_y_u8: win_size 7 exceeds the %dx%d window", h, w);
  const int vh = h - (WIN - 1), vw = w - (WIN - 1);
  dim3 grid((vw + TILE - 1) / TILE, (vh + TILE - 1) / TILE, N);
  TECO_CHECK_ARG(grid.y <= 65535, "teco_metrics_ssim_y_u8: frame too tall");
  ssim_kernel<<<grid, TPB, 0, (cudaStream_t)stream>>>(f, acc);
  TECO_CUDA_LAUNCH_CHECK("teco_metrics_ssim_y_u8");
  return TECO_OK;
}

}  // extern "C"
