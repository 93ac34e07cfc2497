// Total-variation pieces of the TV baselines A-PDS-TV / A-FBS-TV / comparisonB-3 (colour images, C = 3).
//
// Reference: operators.py:117-137 (D, D_T: forward differences and the reference's own "transpose"), operators.py:110-112
// (prox_l12), iteration.py:88-99,133-140 (the loops).  The dual variable y1 is (B, 6, H, W): channels 0..2 vertical
// differences of the three colours, 3..5 horizontal differences.
//
//   tv_primal :  x+ = u - gamma1 * D_T(y1)            u = x - gamma1 * Phi^T(...) comes from the primal / FBS kernels
//   tv_dual   :  w = y1 + gamma2 * D(2 x+ - x);  y1+ = w - gamma2 * prox_l12(w / gamma2, 1 / gamma2) = w * min(1, 1 / ||w||_pixel)
//                (the Moreau step of iteration.py:91 in closed form: the projection of the six differences of a pixel onto
//                the unit l2 ball; ||w|| = 0 leaves w = 0, as the reference's inf-arithmetic does)
//
// D_T is restated as the reference writes it: the last row / column returns +y[last] where the exact adjoint has
// +y[last-1] (operators.py:134-135).  Both kernels are pointwise with nearest-neighbour reads: HBM-bound, 40 B / 76 B per
// pixel-channel, neighbours served by L1/L2.
#include "kernels.cuh"

namespace pds {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) tv_primal_kernel(Dims d, const float* __restrict__ u, const float* __restrict__ y1,
                                                             const ItemParams* __restrict__ prm, float* __restrict__ xn) {
  const int b = blockIdx.y;
  const float g1 = prm[b].g1;
  const int H = d.H, W = d.W;
  for (int e = blockIdx.x * kThreads + threadIdx.x; e < d.n; e += gridDim.x * kThreads) {
    const int c = e / d.hw, p = e - c * d.hw;
    const int i = p / W, j = p - i * W;
    const float* v = y1 + ((size_t)b * 6 + c) * d.hw;
    const float* h = y1 + ((size_t)b * 6 + 3 + c) * d.hw;
    float xv, xh;
    if (i == 0) xv = -v[p];
    else if (i < H - 1) xv = v[p - W] - v[p];
    else xv = v[p];
    if (j == 0) xh = -h[p];
    else if (j < W - 1) xh = h[p - 1] - h[p];
    else xh = h[p];
    const size_t g = (size_t)b * d.n + e;
    xn[g] = fmaf(-g1, xv + xh, u[g]);
  }
}

__global__ void __launch_bounds__(kThreads) tv_dual_kernel(Dims d, const float* __restrict__ xn, const float* __restrict__ x,
                                                           const ItemParams* __restrict__ prm, float* __restrict__ y1) {
  const int b = blockIdx.y;
  const float g2 = prm[b].g2;
  const int H = d.H, W = d.W;
  for (int p = blockIdx.x * kThreads + threadIdx.x; p < d.hw; p += gridDim.x * kThreads) {
    const int i = p / W, j = p - i * W;
    float w[6];
    float n2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t g = ((size_t)b * 3 + c) * d.hw + p;
      const float z = 2.f * xn[g] - x[g];
      const float dv = i < H - 1 ? (2.f * xn[g + W] - x[g + W]) - z : 0.f;
      const float dh = j < W - 1 ? (2.f * xn[g + 1] - x[g + 1]) - z : 0.f;
      w[c] = fmaf(g2, dv, y1[((size_t)b * 6 + c) * d.hw + p]);
      w[3 + c] = fmaf(g2, dh, y1[((size_t)b * 6 + 3 + c) * d.hw + p]);
      n2 = fmaf(w[c], w[c], n2);
      n2 = fmaf(w[3 + c], w[3 + c], n2);
    }
    const float sc = n2 > 1.f ? 1.f / sqrtf(n2) : 1.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) y1[((size_t)b * 6 + c) * d.hw + p] = w[c] * sc;
  }
}

int grid_x(int n) {
  const int g = (n + kThreads - 1) / kThreads;
  return g < 148 * 8 ? (g > 0 ? g : 1) : 148 * 8;
}

}  // namespace

cudaError_t launch_tv_primal(const Dims& d, const float* u, const float* y1, const ItemParams* prm, float* xn, cudaStream_t st) {
  tv_primal_kernel<<<dim3(grid_x(d.n), d.B), kThreads, 0, st>>>(d, u, y1, prm, xn);
  return cudaGetLastError();
}

cudaError_t launch_tv_dual(const Dims& d, const float* xn, const float* x, const ItemParams* prm, float* y1, cudaStream_t st) {
  tv_dual_kernel<<<dim3(grid_x(d.hw), d.B), kThreads, 0, st>>>(d, xn, x, prm, y1);
  return cudaGetLastError();
}

}  // namespace pds
