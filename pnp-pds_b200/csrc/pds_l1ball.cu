// Projection onto the l1 ball (sparse-noise term of ours-B) without a sort.
//
// Reference: operators.py:94-100
//     tau = max(0, max_k (cumsum(sort_desc|z|)_k - eta)/k),  out = sign(z) max(|z| - tau, 0)
// with z = s - gamma1*y (iteration.py:56).  The same tau is the root of
//     f(tau) = sum_i max(|z_i| - tau, 0) - eta
// (piecewise linear, convex, decreasing).  Michelot's fixed point
//     tau <- (sum_{|z_i| > tau} |z_i| - eta) / #{|z_i| > tau}
// started from the full set is Newton's method on f from the left: tau increases monotonically and
// terminates, exactly, as soon as the active set stops shrinking.  Sums are accumulated in double,
// so the result agrees with the sort-based reference to fp32 round-off of the inputs.
//
// One thread-block CLUSTER (8 CTAs x 1024 threads) owns one item: every pass is a strided sweep
// over the item's elements, a block reduction, and one hardware cluster barrier; the 8 partial
// (sum, count) pairs are read through distributed shared memory, so there is no global atomics
// traffic and no host round trip for the data-dependent pass count.
//   * items of at most 8 x 1024 x 32 = 262 144 elements (a 512 x 512 gray image: BASELINE cfg2) are read ONCE: every thread
//     keeps its 32 values of z in registers, the passes are register sweeps + the cluster reduction;
//   * larger items re-read z every pass.  A single large item stays in the 126 MB L2; a batch does NOT (ncu, 64 items of
//     1024^2: L2 hit rate ~0, 8 B per element and pass from HBM) - the sweep then runs at HBM speed, which at the 2-4 passes
//     of the PDS steady state is 0.2 % of the iteration.
// The pass count is bounded by the element count (the active set shrinks strictly until it stops), so the loop has no
// artificial cap that could leave tau short of the root.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace pds {
namespace {

constexpr int kCluster = 8, kThreads = 1024, kRegs = 32;

struct L1Args {
  Dims d;
  const float* s_in;
  const float* t;            // may be null: z = s_in
  const ItemParams* prm;
  const double* sums_prev;
  float eta_override;        // < 0: use prm[b].eta
  float* s_out;
  float* tau_out;
};

template <bool CACHE>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads) l1ball_kernel(L1Args a) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double part[2][2];          // [slot][sum, count] of this CTA
  __shared__ double wred[2][kThreads / 32];
  const int b = blockIdx.y;
  const unsigned rank = cluster.block_rank();
  const size_t base = (size_t)b * a.d.n;
  const int n = a.d.n;
  float gs = 0.f, eta = a.eta_override;
  if (a.t != nullptr || eta < 0.f) {
    const ItemParams p = a.prm[b];
    if (a.t != nullptr) gs = p.g1 * item_sigma(PDS_METHOD_B, a.sums_prev, b, p);
    if (eta < 0.f) eta = p.eta;
  }
  const float* __restrict__ s = a.s_in + base;
  const float* __restrict__ t = a.t ? a.t + base : nullptr;
  const int stride = kCluster * kThreads;
  const int first = rank * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  float zr[CACHE ? kRegs : 1];
  int nk = 0;                 // valid entries of zr
  if constexpr (CACHE) {
    nk = first < n ? (n - first + stride - 1) / stride : 0;
#pragma unroll
    for (int k = 0; k < kRegs; ++k) {
      const int i = first + k * stride;
      float z = 0.f;
      if (k < nk) {
        z = __ldg(s + i);
        if (t) z = fmaf(-gs, __ldg(t + i), z);
      }
      zr[k] = z;
    }
  }

  double tau = -1.0;          // pass 0 selects everything
  double prev_cnt = -1.0;
  bool inside = false;        // ||z||_1 <= eta  ->  identity
  for (int pass = 0; pass <= n + 1; ++pass) {     // terminates by itself: the active set shrinks strictly until it stops
    const float tf = (float)tau;
    double sum = 0.0;
    int cnt = 0;
    if constexpr (CACHE) {
#pragma unroll
      for (int k = 0; k < kRegs; ++k) {
        const float az = fabsf(zr[k]);
        if (k < nk && az > tf) {
          sum += (double)az;
          ++cnt;
        }
      }
    } else {
      for (int i = first; i < n; i += stride) {
        float z = __ldg(s + i);
        if (t) z = fmaf(-gs, __ldg(t + i), z);
        const float az = fabsf(z);
        if (az > tf) {
          sum += (double)az;
          ++cnt;
        }
      }
    }
    double c = (double)cnt;
    sum = warp_sum(sum);
    c = warp_sum(c);
    if (lane == 0) {
      wred[0][warp] = sum;
      wred[1][warp] = c;
    }
    __syncthreads();
    if (warp == 0) {
      double s2 = wred[0][lane], c2 = wred[1][lane];   // kThreads/32 == 32 warps
      s2 = warp_sum(s2);
      c2 = warp_sum(c2);
      if (lane == 0) {
        part[pass & 1][0] = s2;
        part[pass & 1][1] = c2;
      }
    }
    cluster.sync();
    double S = 0.0, N = 0.0;
#pragma unroll
    for (int r = 0; r < kCluster; ++r) {
      const double* rp = cluster.map_shared_rank(&part[pass & 1][0], r);
      S += rp[0];
      N += rp[1];
    }
    if (pass == 0 && S <= (double)eta) {
      inside = true;
      break;
    }
    if (N == prev_cnt || N <= 0.0) break;     // active set unchanged: tau is exact
    prev_cnt = N;
    double nt = (S - (double)eta) / N;
    // tau must be compared in the precision the data has; stop when it no longer moves in fp32
    if (pass > 0 && (float)nt == tf) { tau = nt; break; }
    tau = nt;
  }
  // a CTA may not exit (or reuse `part`) while a peer can still read its shared memory
  cluster.sync();
  const float thr = inside ? 0.f : fmaxf((float)tau, 0.f);
  if (a.tau_out && rank == 0 && threadIdx.x == 0) a.tau_out[b] = thr;
  float* __restrict__ out = a.s_out + base;
  if constexpr (CACHE) {
#pragma unroll
    for (int k = 0; k < kRegs; ++k)
      if (k < nk) out[first + k * stride] = copysignf(fmaxf(fabsf(zr[k]) - thr, 0.f), zr[k]);
  } else {
    for (int i = first; i < n; i += stride) {
      float z = __ldg(s + i);
      if (t) z = fmaf(-gs, __ldg(t + i), z);
      const float m = fmaxf(fabsf(z) - thr, 0.f);
      out[i] = copysignf(m, z);
    }
  }
}

}  // namespace

cudaError_t launch_l1ball(const Dims& d, const float* s_in, const float* t, const ItemParams* prm, const double* sums_prev,
                          float eta_override, float* s_out, float* tau_out, cudaStream_t st) {
  L1Args a{d, s_in, t, prm, sums_prev, eta_override, s_out, tau_out};
  dim3 grid(kCluster, d.B);
  if (d.n <= kCluster * kThreads * kRegs) l1ball_kernel<true><<<grid, kThreads, 0, st>>>(a);
  else l1ball_kernel<false><<<grid, kThreads, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace pds
