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
  __shared__ double red[NACC * (kThreads / 32)];

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
    double v[NACC] = {(double)acc_t, (double)acc_dx, (double)acc_x, (double)acc_e};
    block_accumulate<NACC>(v, a.s.sums_cur + (size_t)b * NSUM, red);
  }
}

// ------------------------------------------------------------------------------------------------
// Register-tiled variant (used when the tap bounding box fits RY x RX): one thread = 8 consecutive
// output columns of one row; lane <-> row, warp <-> column octet.  Per halo row the thread loads its
// (8 + 2 RX)-wide window with LDS.128 (pitch = 4 mod 32 floats: conflict-free) and the weight row by
// broadcast, and applies up to (2 RX + 1) x 8 FMAs; all-zero taps are skipped by a warp-uniform branch.
// ~7 FMAs per shared-memory load instead of 1, so the stencil runs at the FP32 pipe instead of the LSU.
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&r)[8]) {
  if constexpr (V == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  }
}

template <int MODE, int METHOD, int RY, int RX>
__global__ void __launch_bounds__(256) blur_rt_kernel(BlurArgs a) {
  constexpr int TWR = 64, THR = 32;
  constexpr int HC = TWR + 2 * RX, HR = THR + 2 * RY;
  constexpr int PITCH = ((HC + 27) / 32) * 32 + 4;
  constexpr int WROW = ((2 * RX + 1 + 3) / 4) * 4;
  constexpr int NV = 8 + 2 * RX, NV4 = (NV + 3) / 4;
  static_assert(8 * 7 + NV4 * 4 <= PITCH, "window read stays inside the padded row");
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                       // [HR][PITCH]
  float* wbox = smem + HR * PITCH;          // [2RY+1][WROW]
  __shared__ double red[NACC * 8];
  const Dims d = a.s.d;
  const int plane = blockIdx.y;
  const int b = plane / d.C;
  const int tyi = blockIdx.x / a.tiles_x, txi = blockIdx.x % a.tiles_x;
  const int x0 = txi * TWR, y0 = tyi * THR;
  const size_t pbase = (size_t)plane * d.hw;

  for (int k = threadIdx.x; k < (2 * RY + 1) * WROW; k += 256) wbox[k] = 0.f;
  for (int idx = threadIdx.x; idx < HR * HC; idx += 256) {
    const int hy = idx / HC, hx = idx - hy * HC;
    const int gy = wrap(y0 - RY + hy, d.H), gx = wrap(x0 - RX + hx, d.W);
    const size_t g = pbase + (size_t)gy * d.W + gx;
    float v;
    if constexpr (MODE == kApply) v = __ldg(a.in + g);
    else if constexpr (MODE == kPrimal) v = __ldg(a.s.t + g);
    else v = 2.f * __ldg(a.s.xn + g) - __ldg(a.s.x + g);
    tile[hy * PITCH + hx] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < a.ntaps; k += 256) {
    const short2 o = a.tap_off[k];
    wbox[((int)o.x + RY) * WROW + (int)o.y + RX] = a.tap_w[k];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
  for (int dyi = 0; dyi < 2 * RY + 1; ++dyi) {
    const float4* rp = reinterpret_cast<const float4*>(tile + (lane + dyi) * PITCH + 8 * warp);
    float v[NV4 * 4];
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
      const float4 q = rp[j];
      v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
    const float* wr = wbox + dyi * WROW;
#pragma unroll
    for (int dxi = 0; dxi < 2 * RX + 1; ++dxi) {
      const float w = wr[dxi];
      if (w != 0.f) {                       // warp-uniform: the weight does not depend on the thread
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(w, v[c + dxi], acc[c]);
      }
    }
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
  const int gy = y0 + lane, gx0 = x0 + 8 * warp;
  if (gy < d.H && gx0 < d.W) {
    const size_t g0 = pbase + (size_t)gy * d.W + gx0;
    const bool vec = ((d.W & 3) == 0);       // rows start 16-byte aligned and a float4 is never split by the edge
    const int nvalid = (d.W - gx0) < 8 ? (d.W - gx0) : 8;
    float res[8], xv[8], xnv[8], tv[8], ob[8], sn[8], so[8], xt[8];
    auto load8 = [&](const float* base, float (&r)[8]) {
      if (vec) {
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(base + g0));
        r[0] = q0.x; r[1] = q0.y; r[2] = q0.z; r[3] = q0.w;
        if (nvalid > 4) {
          const float4 q1 = __ldg(reinterpret_cast<const float4*>(base + g0) + 1);
          r[4] = q1.x; r[5] = q1.y; r[6] = q1.z; r[7] = q1.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c < nvalid) r[c] = __ldg(base + g0 + c);
      }
    };
    auto store8 = [&](float* base, const float (&r)[8]) {
      if (vec) {
        *reinterpret_cast<float4*>(base + g0) = make_float4(r[0], r[1], r[2], r[3]);
        if (nvalid > 4) *(reinterpret_cast<float4*>(base + g0) + 1) = make_float4(r[4], r[5], r[6], r[7]);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) if (c < nvalid) base[g0 + c] = r[c];
      }
    };
    if constexpr (MODE == kApply) {
      store8(a.out, acc);
    } else if constexpr (MODE == kPrimal) {
      load8(a.s.x, xv);
#pragma unroll
      for (int c = 0; c < 8; ++c) res[c] = fmaf(-p.g1 * sg, acc[c], xv[c]);
      store8(a.s.u, res);
    } else {
      load8(a.s.xn, xnv);
      load8(a.s.x, xv);
      load8(a.s.t, tv);                      // same thread reads then writes its own elements
      load8(a.s.obs, ob);
      if constexpr (METHOD == PDS_METHOD_B) {
        load8(a.s.s_new, sn);
        load8(a.s.s_old, so);
      }
      const bool have_true = a.s.xtrue != nullptr;
      if (have_true) load8(a.s.xtrue, xt);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c >= nvalid) continue;
        float v = acc[c];
        if constexpr (METHOD == PDS_METHOD_B) v += 2.f * sn[c] - so[c];
        const float w = fmaf(p.g2, v, sg * tv[c]);
        if constexpr (METHOD == PDS_METHOD_C) {
          res[c] = gkl_dual(w, ob[c], la, lg4);
        } else {
          res[c] = fmaf(-p.g2, ob[c], w);
          acc_t = fmaf(res[c], res[c], acc_t);
        }
        const float dx = xnv[c] - xv[c];
        acc_dx = fmaf(dx, dx, acc_dx);
        acc_x = fmaf(xv[c], xv[c], acc_x);
        if (have_true) {
          const float e = xnv[c] - xt[c];
          acc_e = fmaf(e, e, acc_e);
        }
      }
      store8(a.s.t, res);
    }
  }
  if constexpr (MODE == kDual) {
    double v[NACC] = {(double)acc_t, (double)acc_dx, (double)acc_x, (double)acc_e};
    block_accumulate<NACC>(v, a.s.sums_cur + (size_t)b * NSUM, red);
  }
}

template <int RY, int RX>
constexpr size_t rt_smem_bytes() {
  constexpr int HC = 64 + 2 * RX, HR = 32 + 2 * RY, PITCH = ((HC + 27) / 32) * 32 + 4, WROW = ((2 * RX + 1 + 3) / 4) * 4;
  return (size_t)(HR * PITCH + (2 * RY + 1) * WROW) * sizeof(float);
}

template <int MODE, int METHOD, int RY, int RX>
cudaError_t launch_rt(BlurArgs& a, const Dims& d, const BlurTaps& t, int which, cudaStream_t st) {
  a.tap_w = t.w[which];
  a.tap_off = t.off[which];
  a.ntaps = t.ntaps;
  a.ry = t.ry;
  a.rx = t.rx;
  a.tiles_x = (d.W + 63) / 64;
  const int tiles_y = (d.H + 31) / 32;
  dim3 grid(a.tiles_x * tiles_y, d.B * d.C);
  blur_rt_kernel<MODE, METHOD, RY, RX><<<grid, 256, rt_smem_bytes<RY, RX>(), st>>>(a);
  return cudaGetLastError();
}

size_t smem_bytes(const BlurTaps& t) {
  return (size_t)((TH + 2 * t.ry) * (TW + 2 * t.rx)) * sizeof(float) + (size_t)t.ntaps * (sizeof(float) + sizeof(int));
}

template <int MODE, int METHOD>
cudaError_t launch(BlurArgs& a, const Dims& d, const BlurTaps& t, int which, cudaStream_t st) {
  if (t.ry <= 8 && t.rx <= 4) return launch_rt<MODE, METHOD, 8, 4>(a, d, t, which, st);   // blur_1.mat: 17 x 9 box
  if (t.ry <= 9 && t.rx <= 9) return launch_rt<MODE, METHOD, 9, 9>(a, d, t, which, st);   // any 19 x 19 kernel
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
