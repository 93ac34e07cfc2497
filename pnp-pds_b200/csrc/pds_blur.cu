// Degradation operator Phi / Phi^T for deg_op='blur' as a periodic stencil over a shared-memory
// halo tile, stand-alone and fused with the primal / dual updates of the PnP-PDS iteration.
//
// Reference: operators.py:7-22 (Phi: wrap-pad + FFT multiply + crop == periodic convolution with
// the kernel centred at l//2), operators.py:24-38 (Phi^T: periodic correlation), fused with
// iteration.py:50-52 / 55-58 / 61-63.
//
//   out[i,j] = sum_taps w * in[(i+dy) mod H, (j+dx) mod W]
//     Phi   : (dy,dx) = (c-a, c-b)     Phi^T : (dy,dx) = (a-c, b-c)     for every h[a,b] != 0
//
// Only the non-zero taps are applied (109 of 361 for blur_1.mat); the tile halo is the bounding
// box of the tap offsets.  One CTA = one 64x16 output tile of one (item, channel) plane; the
// halo tile is staged in shared memory once (periodic wrap resolved at load), each thread
// accumulates 4 outputs over the tap list (tap weights / offsets broadcast from shared memory).
#include "kernels.cuh"

namespace pds {
namespace {

constexpr int TW = 64, TH = 16, kThreads = 256, kRows = TH / (kThreads / TW);  // 4 outputs per thread

enum BlurMode { kApply = 0, kPrimal = 1, kDual = 2 };

struct BlurArgs {
  StepArgs s;            // used by kPrimal / kDual
  const float* in;       // kApply
  float* out;            // kApply
  const float* tap_w;
  const short2* tap_off;
  int ntaps, ry, rx;
  int tiles_x;
};

__device__ __forceinline__ int wrap(int v, int n) {
  v %= n;
  return v < 0 ? v + n : v;
}

template <int MODE, int METHOD>
__global__ void __launch_bounds__(kThreads) blur_kernel(BlurArgs a) {
  extern __shared__ float smem[];
  const Dims d = a.s.d;
  const int pitch = TW + 2 * a.rx;
  const int hrows = TH + 2 * a.ry;
  float* tile = smem;
  float* tw = smem + hrows * pitch;
  int* toff = reinterpret_cast<int*>(tw + a.ntaps);
  __shared__ double red[NSUM * (kThreads / 32)];

  const int plane = blockIdx.y;            // b*C + c
  const int b = plane / d.C;
  const int tyi = blockIdx.x / a.tiles_x, txi = blockIdx.x % a.tiles_x;
  const int x0 = txi * TW, y0 = tyi * TH;
  const size_t pbase = (size_t)plane * d.hw;

  for (int k = threadIdx.x; k < a.ntaps; k += kThreads) {
    tw[k] = a.tap_w[k];
    short2 o = a.tap_off[k];
    toff[k] = (int)o.x * pitch + (int)o.y;   // .x = dy, .y = dx
  }
  // stage the halo tile (periodic wrap)
  for (int idx = threadIdx.x; idx < hrows * pitch; idx += kThreads) {
    const int hy = idx / pitch, hx = idx - hy * pitch;
    const int gy = wrap(y0 - a.ry + hy, d.H), gx = wrap(x0 - a.rx + hx, d.W);
    const size_t g = pbase + (size_t)gy * d.W + gx;
    float v;
    if constexpr (MODE == kApply) v = __ldg(a.in + g);
    else if constexpr (MODE == kPrimal) v = __ldg(a.s.t + g);
    else v = 2.f * __ldg(a.s.xn + g) - __ldg(a.s.x + g);
    tile[idx] = v;
  }
  __syncthreads();

  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  float acc[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
  const int centre = (a.ry + ty) * pitch + a.rx + tx;
  const int rstep = (kThreads / TW) * pitch;
#pragma unroll 4
  for (int k = 0; k < a.ntaps; ++k) {
    const float w = tw[k];
    const float* p = tile + centre + toff[k];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = fmaf(w, p[r * rstep], acc[r]);
  }

  float acc_t = 0.f, acc_dx = 0.f, acc_x = 0.f, acc_e = 0.f;
  ItemParams p;
  float sg = 1.f, la = 0.f, lg4 = 0.f;
  if constexpr (MODE != kApply) {
    p = a.s.prm[b];
    sg = item_sigma(METHOD, a.s.sums_prev, b, p);
    la = p.lam * p.alpha;
    lg4 = 4.f * p.lam * p.g2;
  }
  const int gx = x0 + tx;
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int gy = y0 + ty + r * (kThreads / TW);
    if (gx >= d.W || gy >= d.H) continue;
    const size_t g = pbase + (size_t)gy * d.W + gx;
    if constexpr (MODE == kApply) {
      a.out[g] = acc[r];
    } else if constexpr (MODE == kPrimal) {
      a.s.u[g] = fmaf(-p.g1 * sg, acc[r], __ldg(a.s.x + g));
    } else {
      float v = acc[r];
      const float xn = __ldg(a.s.xn + g), x = __ldg(a.s.x + g);
      if constexpr (METHOD == PDS_METHOD_B) v += 2.f * __ldg(a.s.s_new + g) - __ldg(a.s.s_old + g);
      const float w = fmaf(p.g2, v, sg * a.s.t[g]);
      float tn;
      if constexpr (METHOD == PDS_METHOD_C) {
        tn = gkl_dual(w, __ldg(a.s.obs + g), la, lg4);
      } else {
        tn = fmaf(-p.g2, __ldg(a.s.obs + g), w);
        acc_t = fmaf(tn, tn, acc_t);
      }
      a.s.t[g] = tn;
      const float dx = xn - x;
      acc_dx = fmaf(dx, dx, acc_dx);
      acc_x = fmaf(x, x, acc_x);
      if (a.s.xtrue) {
        const float e = xn - __ldg(a.s.xtrue + g);
        acc_e = fmaf(e, e, acc_e);
      }
    }
  }
  if constexpr (MODE == kDual) {
    double v[NSUM] = {(double)acc_t, (double)acc_dx, (double)acc_x, (double)acc_e};
    block_accumulate<NSUM>(v, a.s.sums_cur + (size_t)b * NSUM, red);
  }
}

size_t smem_bytes(const BlurTaps& t) {
  return (size_t)((TH + 2 * t.ry) * (TW + 2 * t.rx)) * sizeof(float) + (size_t)t.ntaps * (sizeof(float) + sizeof(int));
}

template <int MODE, int METHOD>
cudaError_t launch(BlurArgs& a, const Dims& d, const BlurTaps& t, int which, cudaStream_t st) {
  a.tap_w = t.w[which];
  a.tap_off = t.off[which];
  a.ntaps = t.ntaps;
  a.ry = t.ry;
  a.rx = t.rx;
  a.tiles_x = (d.W + TW - 1) / TW;
  const int tiles_y = (d.H + TH - 1) / TH;
  dim3 grid(a.tiles_x * tiles_y, d.B * d.C);
  blur_kernel<MODE, METHOD><<<grid, kThreads, smem_bytes(t), st>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_blur_apply(const Dims& d, const BlurTaps& taps, int adjoint, const float* in, float* out, cudaStream_t st) {
  BlurArgs a{};
  a.s.d = d;
  a.in = in;
  a.out = out;
  return launch<kApply, 0>(a, d, taps, adjoint ? 1 : 0, st);
}

cudaError_t launch_primal_blur(const StepArgs& s, const BlurTaps& taps, cudaStream_t st) {
  BlurArgs a{};
  a.s = s;
  switch (s.method) {
    case PDS_METHOD_C: return launch<kPrimal, PDS_METHOD_C>(a, s.d, taps, 1, st);
    default: return launch<kPrimal, PDS_METHOD_A>(a, s.d, taps, 1, st);
  }
}

cudaError_t launch_dual_blur(const StepArgs& s, const BlurTaps& taps, cudaStream_t st) {
  BlurArgs a{};
  a.s = s;
  switch (s.method) {
    case PDS_METHOD_A: return launch<kDual, PDS_METHOD_A>(a, s.d, taps, 0, st);
    case PDS_METHOD_B: return launch<kDual, PDS_METHOD_B>(a, s.d, taps, 0, st);
    default: return launch<kDual, PDS_METHOD_C>(a, s.d, taps, 0, st);
  }
}

}  // namespace pds
