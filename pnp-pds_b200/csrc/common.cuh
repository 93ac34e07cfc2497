// Shared device/host helpers for libpnp_pds (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/pnp_pds.h"

namespace pds {

// ---------------------------------------------------------------- errors
void set_error(const std::string& msg);
#define PDS_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::pds::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +  \
                       __FILE__ + ":" + std::to_string(__LINE__) + ")");                   \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)
#define PDS_REQUIRE(cond, msg)                                                             \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      ::pds::set_error(std::string(msg));                                                  \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a process-wide "done" flag would leave the kernel without
// its opt-in on a second GPU.  One bit per device ordinal, set atomically.
struct PerDeviceOnce {
  unsigned long long mask = 0ull;
  // true exactly once per device (and always for ordinals >= 64, where the attribute is simply set again)
  bool first_use() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const unsigned long long bit = 1ull << dev;
    return (__atomic_fetch_or(&mask, bit, __ATOMIC_ACQ_REL) & bit) == 0ull;
  }
  void retract() {   // the attribute call failed: try again on the next launch
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) __atomic_fetch_and(&mask, ~(1ull << dev), __ATOMIC_ACQ_REL);
  }
};

// ---------------------------------------------------------------- device-side layout
// Per-item hyper-parameters as the kernels read them (same layout as pds_item_params_t).
struct ItemParams {
  float g1, g2, eps, eta, lam, alpha;
};

enum : int { SUM_T2 = 0, SUM_DX2 = 1, SUM_X2 = 2, SUM_ERR2 = 3, SUM_SSIM = 4, NSUM = PDS_TRACE_WIDTH, NACC = 4 };  // NACC: sums the dual kernels produce

constexpr int kMid = 64;  // channel width of the DnCNN body (simple_CNN n_ch, basic_models.py:9)

#ifdef __CUDACC__
// sigma of the lazy l2-ball form:  y = sigma * t,
// sigma = max(0, 1 - gamma2*eps/||t||).  Reference: iteration.py:52 + operators.py:102-108.
__device__ __forceinline__ float sigma_from_norm2(double t2, float g2, float eps) {
  float nt = (float)sqrt(t2);
  float ge = g2 * eps;
  return (nt > ge) ? (1.0f - ge / nt) : 0.0f;
}

// sigma for item b given the previous iteration's sums (null on the first iteration: t == 0).
__device__ __forceinline__ float item_sigma(int method, const double* __restrict__ sums_prev, int b, const ItemParams& p) {
  if (method == PDS_METHOD_C || sums_prev == nullptr) return 1.0f;
  return sigma_from_norm2(sums_prev[(size_t)b * NSUM + SUM_T2], p.g2, p.eps);
}

// The generalised-KL dual update of ours-C (iteration.py:63 with operators.py:114-115):
//   y+ = w - g2*prox_GKL(w/g2, lam/g2, alpha, b) = 0.5*(w + lam*alpha - sqrt((w - lam*alpha)^2 + 4 lam g2 b))
// evaluated in the cancellation-free form.  la = lam*alpha, lg4 = 4*lam*g2.
__device__ __forceinline__ float gkl_dual(float w, float b, float la, float lg4) {
  float q = w - la;
  float c = lg4 * b;
  float d = sqrtf(fmaf(q, q, c));
  return (q > 0.f) ? (la - 0.5f * c / (q + d)) : (la + 0.5f * (q - d));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-reduce NV doubles held per thread and add them to global accumulators.
// smem must hold NV * (blockDim.x/32) doubles.
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double* __restrict__ gacc, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) smem[k * nwarp + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = (lane < nwarp) ? smem[k * nwarp + lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) atomicAdd(&gacc[k], s);
    }
  }
}

// fp32 -> (hi, lo) fp16 pair with hi + lo == v to ~2^-22 relative.
__device__ __forceinline__ void split_hi_lo(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__device__ __forceinline__ float leaky(float v, float slope) { return v >= 0.f ? v : v * slope; }
#endif  // __CUDACC__

}  // namespace pds
