// Structural similarity of x_{k+1} against x_true on the device (per-iteration SSIM trace).
//
// Reference: utils/utils_eval.py:9-12 — skimage.metrics.structural_similarity(im1=x_true, im2=x,
// data_range = x.max() - x.min(), channel_axis=0) with scikit-image's defaults: uniform 7-tap window,
// K1 = 0.01, K2 = 0.03, sample covariance (N/(N-1)), mean over the positions where the window fits.
// channel_axis=0 quirk: a gray (H,W) image is treated as H channels of 1-D signals
// (7-tap windows along x only); a colour (3,H,W) image as three 2-D channels (7x7 windows).
// PARITY UNPINNED: scikit-image is not available to record vectors from; this restates its documented
// algorithm and is checked against the oracle's restatement only.
//
// Two small kernels per evaluation: per-item min/max of x (data_range), then window sums in double
// over a shared-memory tile, S accumulated per item with one atomicAdd(double) per block.
#include "kernels.cuh"

namespace pds {
namespace {

constexpr int kThreads = 256, kWin = 7, kTX = 32, kTY = 8;

__device__ __forceinline__ unsigned enc(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // monotone float -> unsigned
}
__device__ __forceinline__ float dec(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(kThreads) minmax_kernel(Dims d, const float* __restrict__ x, unsigned* __restrict__ mm /*[B][2]*/) {
  __shared__ unsigned smin[kThreads / 32], smax[kThreads / 32];
  const size_t base = (size_t)blockIdx.y * d.n;
  unsigned lo = 0xffffffffu, hi = 0u;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < d.n; i += gridDim.x * kThreads) {
    const unsigned k = enc(__ldg(x + base + i));
    lo = min(lo, k);
    hi = max(hi, k);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) { lo = min(lo, smin[w]); hi = max(hi, smax[w]); }
    atomicMin(&mm[blockIdx.y * 2], lo);
    atomicMax(&mm[blockIdx.y * 2 + 1], hi);
  }
}

// GRAY1D: 1-D windows along x (rows are "channels"); otherwise 2-D windows per (item, channel) plane.
template <bool GRAY1D>
__global__ void __launch_bounds__(kThreads) ssim_kernel(Dims d, const float* __restrict__ xt, const float* __restrict__ x,
                                                        const unsigned* __restrict__ mm, double* __restrict__ sums_cur) {
  constexpr int HY = GRAY1D ? kTY : kTY + kWin - 1, HX = kTX + kWin - 1;
  __shared__ float sa[HY][HX + 1], sb[HY][HX + 1];
  __shared__ double red[kThreads / 32];
  const int plane = blockIdx.y, b = plane / d.C;
  const int nx = d.W - (kWin - 1), ny = GRAY1D ? d.H : d.H - (kWin - 1);     // window positions
  const int tiles_x = (nx + kTX - 1) / kTX;
  const int x0 = (blockIdx.x % tiles_x) * kTX, y0 = (blockIdx.x / tiles_x) * kTY;
  const size_t pbase = (size_t)plane * d.hw;
  for (int i = threadIdx.x; i < HY * HX; i += kThreads) {
    const int hy = i / HX, hx = i - hy * HX;
    const int gy = y0 + hy, gx = x0 + hx;
    float va = 0.f, vb = 0.f;
    if (gy < d.H && gx < d.W) {
      va = __ldg(xt + pbase + (size_t)gy * d.W + gx);
      vb = __ldg(x + pbase + (size_t)gy * d.W + gx);
    }
    sa[hy][hx] = va;
    sb[hy][hx] = vb;
  }
  __syncthreads();
  const float range = dec(mm[b * 2 + 1]) - dec(mm[b * 2]);
  const double c1 = (0.01 * (double)range) * (0.01 * (double)range), c2 = (0.03 * (double)range) * (0.03 * (double)range);
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  double s = 0.0;
  if (x0 + tx < nx && y0 + ty < ny) {
    double ua = 0, ub = 0, uaa = 0, ubb = 0, uab = 0;
    constexpr int WY = GRAY1D ? 1 : kWin;
#pragma unroll
    for (int wy = 0; wy < WY; ++wy)
#pragma unroll
      for (int wx = 0; wx < kWin; ++wx) {
        const double p = sa[ty + wy][tx + wx], q = sb[ty + wy][tx + wx];
        ua += p; ub += q; uaa += p * p; ubb += q * q; uab += p * q;
      }
    constexpr double np_ = GRAY1D ? (double)kWin : (double)(kWin * kWin);
    constexpr double cov = np_ / (np_ - 1.0);
    ua /= np_; ub /= np_; uaa /= np_; ubb /= np_; uab /= np_;
    const double va = cov * (uaa - ua * ua), vb = cov * (ubb - ub * ub), vab = cov * (uab - ua * ub);
    s = ((2 * ua * ub + c1) * (2 * vab + c2)) / ((ua * ua + ub * ub + c1) * (va + vb + c2));
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(&sums_cur[(size_t)b * NSUM + SUM_SSIM], t);
  }
}

}  // namespace

// mm: device scratch [B][2] unsigned.  Adds sum of S over all window positions of item b to sums_cur[b][SUM_SSIM].
cudaError_t launch_ssim(const Dims& d, const float* xtrue, const float* x, unsigned* mm, double* sums_cur, cudaStream_t st) {
  if (d.W < kWin || (d.C != 1 && d.H < kWin)) return cudaSuccess;            // window does not fit: SSIM undefined
  cudaError_t e = cudaMemsetAsync(mm, 0, (size_t)d.B * 2 * sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  // min slots must start at 0xffffffff: set byte pattern 0xff on the even entries via a 2-D memset
  e = cudaMemset2DAsync(mm, 2 * sizeof(unsigned), 0xff, sizeof(unsigned), (size_t)d.B, st);
  if (e != cudaSuccess) return e;
  int gx = (d.n + kThreads * 8 - 1) / (kThreads * 8);
  gx = gx < 1 ? 1 : (gx > 148 * 4 ? 148 * 4 : gx);
  minmax_kernel<<<dim3(gx, d.B), kThreads, 0, st>>>(d, x, mm);
  const bool gray = d.C == 1;
  const int nx = d.W - (kWin - 1), ny = gray ? d.H : d.H - (kWin - 1);
  const int tiles = ((nx + kTX - 1) / kTX) * ((ny + kTY - 1) / kTY);
  if (gray) ssim_kernel<true><<<dim3(tiles, d.B * d.C), kThreads, 0, st>>>(d, xtrue, x, mm, sums_cur);
  else ssim_kernel<false><<<dim3(tiles, d.B * d.C), kThreads, 0, st>>>(d, xtrue, x, mm, sums_cur);
  return cudaGetLastError();
}

}  // namespace pds
