// Fused elementwise primal / dual kernels of the PnP-PDS iteration for pointwise degradation
// operators (deg_op = Id or random_sampling), plus the stand-alone prox operators.
//
// Reference: iteration.py:48-63 (three proposed branches), operators.py:40-58 (mask),
// operators.py:102-108 (l2 ball), operators.py:114-115 (GKL prox), iteration.py:187 and
// utils_eval.py:4-7 (per-iteration metrics).
//
// HBM-bound streaming kernels: every operand is read once with 128-bit loads, one pass,
// reductions are block-reduced in double and committed with one atomicAdd per block and
// quantity.  The l2-ball projection is applied lazily: the state holds t with y = sigma*t and
// sigma is recomputed by every consumer from the previous iteration's ||t||^2 (no grid sync).
#include "kernels.cuh"

namespace pds {

namespace {

constexpr int kThreads = 256;

template <int V>
struct Vec;
template <>
struct Vec<4> {
  using type = float4;
};
template <>
struct Vec<1> {
  using type = float;
};

template <int V>
__device__ __forceinline__ void load(const float* __restrict__ p, size_t i, float (&r)[V]) {
  if constexpr (V == 4) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else {
    r[0] = __ldg(p + i);
  }
}
// plain (coherent) load for buffers the same kernel also writes
template <int V>
__device__ __forceinline__ void load_rw(const float* p, size_t i, float (&r)[V]) {
  if constexpr (V == 4) {
    float4 v = *reinterpret_cast<const float4*>(p + i);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else {
    r[0] = p[i];
  }
}
template <int V>
__device__ __forceinline__ void store(float* __restrict__ p, size_t i, const float (&r)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p + i) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    p[i] = r[0];
  }
}
template <int V>
__device__ __forceinline__ void load_mask(const uint8_t* __restrict__ m, int i, float (&r)[V]) {
  if constexpr (V == 4) {
    uchar4 v = __ldg(reinterpret_cast<const uchar4*>(m + i));
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else {
    r[0] = __ldg(m + i);
  }
}

// u = x - gamma1 * Phi^T(sigma t),  Phi^T = I or mask.
template <int V, bool MASK>
__global__ void __launch_bounds__(kThreads) primal_pw_kernel(StepArgs a) {
  const int b = blockIdx.y;
  const ItemParams p = a.prm[b];
  const float gs = p.g1 * item_sigma(a.method, a.sums_prev, b, p);
  const size_t base = (size_t)b * a.d.n;
  const int nv = a.d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float x[V], t[V], m[V], u[V];
    load<V>(a.x, base + (size_t)i * V, x);
    load<V>(a.t, base + (size_t)i * V, t);
    if constexpr (MASK) load_mask<V>(a.mask, (i * V) % a.d.hw, m);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float y = gs * t[k];
      if constexpr (MASK) y *= m[k];
      u[k] = x[k] - y;
    }
    store<V>(a.u, base + (size_t)i * V, u);
  }
}

// w = sigma t + g2 (Phi(2 x+ - x) [+ 2 s+ - s]);  A,B: t+ = w - g2 b ;  C: y+ = gkl(w)
template <int V, bool MASK, int METHOD>
__global__ void __launch_bounds__(kThreads) dual_pw_kernel(StepArgs a) {
  __shared__ double red[NACC * (kThreads / 32)];
  const int b = blockIdx.y;
  const ItemParams p = a.prm[b];
  const float sg = item_sigma(METHOD, a.sums_prev, b, p);
  const float la = p.lam * p.alpha, lg4 = 4.f * p.lam * p.g2;
  const size_t base = (size_t)b * a.d.n;
  const int nv = a.d.n / V;
  const bool have_true = a.xtrue != nullptr;
  float acc_t = 0.f, acc_dx = 0.f, acc_x = 0.f, acc_e = 0.f;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    const size_t o = base + (size_t)i * V;
    float xn[V], x[V], t[V], ob[V], m[V], sn[V], so[V], xt[V], tn[V];
    load<V>(a.xn, o, xn);
    load<V>(a.x, o, x);
    load_rw<V>(a.t, o, t);
    load<V>(a.obs, o, ob);
    if constexpr (MASK) load_mask<V>(a.mask, (i * V) % a.d.hw, m);
    if constexpr (METHOD == PDS_METHOD_B) {
      load<V>(a.s_new, o, sn);
      load<V>(a.s_old, o, so);
    }
    if (have_true) load<V>(a.xtrue, o, xt);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float v = 2.f * xn[k] - x[k];
      if constexpr (MASK) v *= m[k];
      if constexpr (METHOD == PDS_METHOD_B) v += 2.f * sn[k] - so[k];
      float w = fmaf(p.g2, v, sg * t[k]);
      if constexpr (METHOD == PDS_METHOD_C) {
        tn[k] = gkl_dual(w, ob[k], la, lg4);
      } else {
        tn[k] = fmaf(-p.g2, ob[k], w);
        acc_t = fmaf(tn[k], tn[k], acc_t);
      }
      float dx = xn[k] - x[k];
      acc_dx = fmaf(dx, dx, acc_dx);
      acc_x = fmaf(x[k], x[k], acc_x);
      if (have_true) {
        float e = xn[k] - xt[k];
        acc_e = fmaf(e, e, acc_e);
      }
    }
    store<V>(a.t, o, tn);
  }
  double v[NACC] = {(double)acc_t, (double)acc_dx, (double)acc_x, (double)acc_e};
  block_accumulate<NACC>(v, a.sums_cur + (size_t)b * NSUM, red);
}

template <int V>
__global__ void __launch_bounds__(kThreads) mask_apply_kernel(Dims d, const float* __restrict__ in,
                                                              const uint8_t* __restrict__ mask, float* __restrict__ out) {
  const size_t base = (size_t)blockIdx.y * d.n;
  const int nv = d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float x[V], m[V];
    load<V>(in, base + (size_t)i * V, x);
    load_mask<V>(mask, (i * V) % d.hw, m);
#pragma unroll
    for (int k = 0; k < V; ++k) x[k] *= m[k];
    store<V>(out, base + (size_t)i * V, x);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) scale_sigma_kernel(Dims d, const float* __restrict__ t, const ItemParams* prm,
                                                               const double* sums, int method, float* __restrict__ y) {
  const int b = blockIdx.y;
  const float sg = item_sigma(method, sums, b, prm[b]);
  const size_t base = (size_t)b * d.n;
  const int nv = d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float v[V];
    load<V>(t, base + (size_t)i * V, v);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] *= sg;
    store<V>(y, base + (size_t)i * V, v);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) diff_norm2_kernel(Dims d, const float* __restrict__ x, const float* __restrict__ c,
                                                              double* __restrict__ acc) {
  __shared__ double red[kThreads / 32];
  const size_t base = (size_t)blockIdx.y * d.n;
  const int nv = d.n / V;
  float s = 0.f;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float a[V], q[V];
    load<V>(x, base + (size_t)i * V, a);
    load<V>(c, base + (size_t)i * V, q);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float e = a[k] - q[k];
      s = fmaf(e, e, s);
    }
  }
  double v[1] = {(double)s};
  block_accumulate<1>(v, acc + blockIdx.y, red);
}

// operators.py:102-108: out = c + eps (x-c)/||x-c|| if ||x-c|| > eps else x
template <int V>
__global__ void __launch_bounds__(kThreads) proj_l2_apply_kernel(Dims d, const float* __restrict__ x, const float* __restrict__ c,
                                                                 float eps, const double* __restrict__ acc, float* __restrict__ out) {
  const size_t base = (size_t)blockIdx.y * d.n;
  const float nrm = (float)sqrt(acc[blockIdx.y]);
  const bool outside = nrm > eps;
  const float sc = outside ? eps / nrm : 1.f;
  const int nv = d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float a[V], q[V];
    load<V>(x, base + (size_t)i * V, a);
    load<V>(c, base + (size_t)i * V, q);
    if (outside) {
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] = fmaf(sc, a[k] - q[k], q[k]);
    }
    store<V>(out, base + (size_t)i * V, a);
  }
}

// same with the radius taken per item from prm[b].eps (ADMM loops)
template <int V>
__global__ void __launch_bounds__(kThreads) proj_l2_items_kernel(Dims d, const float* x, const float* __restrict__ c, const ItemParams* prm,
                                                                 const double* __restrict__ acc, float* out) {
  const size_t base = (size_t)blockIdx.y * d.n;
  const float eps = prm[blockIdx.y].eps;
  const float nrm = (float)sqrt(acc[blockIdx.y]);
  const bool outside = nrm > eps;
  const float sc = outside ? eps / nrm : 1.f;
  const int nv = d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float a[V], q[V];
    load_rw<V>(x, base + (size_t)i * V, a);
    load<V>(c, base + (size_t)i * V, q);
    if (outside) {
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] = fmaf(sc, a[k] - q[k], q[k]);
    }
    store<V>(out, base + (size_t)i * V, a);
  }
}

// operators.py:114-115, stable for x - gamma*alpha << 0.
template <int V>
__global__ void __launch_bounds__(kThreads) prox_gkl_kernel(Dims d, const float* __restrict__ x, const float* __restrict__ x0,
                                                            float gamma, float alpha, float* __restrict__ out) {
  const size_t base = (size_t)blockIdx.y * d.n;
  const int nv = d.n / V;
  const float ga = gamma * alpha, g4 = 4.f * gamma;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float a[V], b[V];
    load<V>(x, base + (size_t)i * V, a);
    load<V>(x0, base + (size_t)i * V, b);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float q = a[k] - ga, c = g4 * b[k];
      float dd = sqrtf(fmaf(q, q, c));
      a[k] = (q >= 0.f) ? 0.5f * (q + dd) : 0.5f * c / (dd - q);
    }
    store<V>(out, base + (size_t)i * V, a);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) metrics_kernel(Dims d, const float* __restrict__ xn, const float* __restrict__ x,
                                                           const float* __restrict__ xtrue, double* __restrict__ sums_cur) {
  __shared__ double red[NACC * (kThreads / 32)];
  const size_t base = (size_t)blockIdx.y * d.n;
  const int nv = d.n / V;
  float acc_dx = 0.f, acc_x = 0.f, acc_e = 0.f;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float a[V], q[V], tr[V];
    load<V>(xn, base + (size_t)i * V, a);
    load<V>(x, base + (size_t)i * V, q);
    if (xtrue) load<V>(xtrue, base + (size_t)i * V, tr);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float e = a[k] - q[k];
      acc_dx = fmaf(e, e, acc_dx);
      acc_x = fmaf(q[k], q[k], acc_x);
      if (xtrue) {
        float g = a[k] - tr[k];
        acc_e = fmaf(g, g, acc_e);
      }
    }
  }
  double v[NACC] = {0.0, (double)acc_dx, (double)acc_x, (double)acc_e};
  block_accumulate<NACC>(v, sums_cur + (size_t)blockIdx.y * NSUM, red);
}

__global__ void __launch_bounds__(kThreads) axpbypcz_kernel(size_t n, float a, const float* __restrict__ p, float b,
                                                            const float* __restrict__ q, float c, const float* __restrict__ r,
                                                            float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    float v = a * p[i];
    if (q) v = fmaf(b, q[i], v);
    if (r) v = fmaf(c, r[i], v);
    out[i] = v;
  }
}

// out[i] = sum_k coef[b][k] * in_k[i]   (k < nterms <= 6); out may alias an input (same index only)
template <int V>
__global__ void __launch_bounds__(kThreads) lincomb_kernel(LinArgs a) {
  const int b = blockIdx.y;
  float c[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) c[k] = (k < a.nterms) ? a.coef[(size_t)b * 6 + k] : 0.f;
  const size_t base = (size_t)b * a.d.n;
  const int nv = a.d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (k < a.nterms) {
        float v[V];
        load_rw<V>(a.in[k], base + (size_t)i * V, v);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fmaf(c[k], v[j], acc[j]);
      }
    }
    store<V>(a.out, base + (size_t)i * V, acc);
  }
}

// out = num / (alpha_b * den)      (admm.py:12: y / (poisson_alpha * phi(x)))
template <int V>
__global__ void __launch_bounds__(kThreads) ratio_kernel(Dims d, const float* num, const float* den, const ItemParams* prm, float* out) {
  const int b = blockIdx.y;
  const float alpha = prm[b].alpha;
  const size_t base = (size_t)b * d.n;
  const int nv = d.n / V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
    float p[V], q[V];
    load_rw<V>(num, base + (size_t)i * V, p);
    load_rw<V>(den, base + (size_t)i * V, q);
#pragma unroll
    for (int j = 0; j < V; ++j) p[j] = p[j] / (alpha * q[j]);
    store<V>(out, base + (size_t)i * V, p);
  }
}

__global__ void __launch_bounds__(kThreads) fill_kernel(size_t n, float v, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) out[i] = v;
}

inline bool vec4_ok(const Dims& d) { return (d.n % 4 == 0) && (d.hw % 4 == 0); }

// grid.x blocks per item so that the whole launch is a few waves of 148 SMs x 8 resident CTAs.
inline dim3 grid_for(const Dims& d, int V) {
  int nv = d.n / V;
  int want = (nv + kThreads * 4 - 1) / (kThreads * 4);  // >= 4 vectors per thread
  int cap = (148 * 8 * 2 + d.B - 1) / d.B;
  int gx = want < 1 ? 1 : (want > cap ? cap : want);
  return dim3(gx, d.B, 1);
}

}  // namespace

#define PDS_DISPATCH_V(d, CALL4, CALL1) \
  do {                                  \
    if (vec4_ok(d)) {                   \
      CALL4;                            \
    } else {                            \
      CALL1;                            \
    }                                   \
  } while (0)

cudaError_t launch_primal_pointwise(const StepArgs& a, cudaStream_t st) {
  const bool masked = a.mask != nullptr;
  if (vec4_ok(a.d)) {
    dim3 g = grid_for(a.d, 4);
    if (masked) primal_pw_kernel<4, true><<<g, kThreads, 0, st>>>(a);
    else primal_pw_kernel<4, false><<<g, kThreads, 0, st>>>(a);
  } else {
    dim3 g = grid_for(a.d, 1);
    if (masked) primal_pw_kernel<1, true><<<g, kThreads, 0, st>>>(a);
    else primal_pw_kernel<1, false><<<g, kThreads, 0, st>>>(a);
  }
  return cudaGetLastError();
}

template <int V, bool MASK>
static void launch_dual_m(const StepArgs& a, cudaStream_t st) {
  dim3 g = grid_for(a.d, V);
  switch (a.method) {
    case PDS_METHOD_A: dual_pw_kernel<V, MASK, PDS_METHOD_A><<<g, kThreads, 0, st>>>(a); break;
    case PDS_METHOD_B: dual_pw_kernel<V, MASK, PDS_METHOD_B><<<g, kThreads, 0, st>>>(a); break;
    default: dual_pw_kernel<V, MASK, PDS_METHOD_C><<<g, kThreads, 0, st>>>(a); break;
  }
}

cudaError_t launch_dual_pointwise(const StepArgs& a, cudaStream_t st) {
  const bool masked = a.mask != nullptr;
  if (vec4_ok(a.d)) {
    if (masked) launch_dual_m<4, true>(a, st);
    else launch_dual_m<4, false>(a, st);
  } else {
    if (masked) launch_dual_m<1, true>(a, st);
    else launch_dual_m<1, false>(a, st);
  }
  return cudaGetLastError();
}

cudaError_t launch_mask_apply(const Dims& d, const float* in, const uint8_t* mask, float* out, cudaStream_t st) {
  PDS_DISPATCH_V(d, (mask_apply_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, in, mask, out)),
                 (mask_apply_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, in, mask, out)));
  return cudaGetLastError();
}

cudaError_t launch_scale_by_sigma(const Dims& d, const float* t, const ItemParams* prm, const double* sums, int method, float* y,
                                  cudaStream_t st) {
  PDS_DISPATCH_V(d, (scale_sigma_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, t, prm, sums, method, y)),
                 (scale_sigma_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, t, prm, sums, method, y)));
  return cudaGetLastError();
}

cudaError_t launch_diff_norm2(const Dims& d, const float* x, const float* c, double* acc, cudaStream_t st) {
  PDS_DISPATCH_V(d, (diff_norm2_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, x, c, acc)),
                 (diff_norm2_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, x, c, acc)));
  return cudaGetLastError();
}

cudaError_t launch_proj_l2_apply(const Dims& d, const float* x, const float* c, float eps, const double* acc, float* out,
                                 cudaStream_t st) {
  PDS_DISPATCH_V(d, (proj_l2_apply_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, x, c, eps, acc, out)),
                 (proj_l2_apply_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, x, c, eps, acc, out)));
  return cudaGetLastError();
}

cudaError_t launch_proj_l2_items(const Dims& d, const float* x, const float* c, const ItemParams* prm, const double* acc, float* out,
                                 cudaStream_t st) {
  PDS_DISPATCH_V(d, (proj_l2_items_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, x, c, prm, acc, out)),
                 (proj_l2_items_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, x, c, prm, acc, out)));
  return cudaGetLastError();
}

cudaError_t launch_prox_gkl(const Dims& d, const float* x, const float* x0, float gamma, float alpha, float* out, cudaStream_t st) {
  PDS_DISPATCH_V(d, (prox_gkl_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, x, x0, gamma, alpha, out)),
                 (prox_gkl_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, x, x0, gamma, alpha, out)));
  return cudaGetLastError();
}

cudaError_t launch_metrics(const Dims& d, const float* xn, const float* x, const float* xtrue, double* sums_cur, cudaStream_t st) {
  PDS_DISPATCH_V(d, (metrics_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, xn, x, xtrue, sums_cur)),
                 (metrics_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, xn, x, xtrue, sums_cur)));
  return cudaGetLastError();
}

cudaError_t launch_lincomb(const LinArgs& a, cudaStream_t st) {
  PDS_DISPATCH_V(a.d, (lincomb_kernel<4><<<grid_for(a.d, 4), kThreads, 0, st>>>(a)), (lincomb_kernel<1><<<grid_for(a.d, 1), kThreads, 0, st>>>(a)));
  return cudaGetLastError();
}

cudaError_t launch_ratio(const Dims& d, const float* num, const float* den, const ItemParams* prm, float* out, cudaStream_t st) {
  PDS_DISPATCH_V(d, (ratio_kernel<4><<<grid_for(d, 4), kThreads, 0, st>>>(d, num, den, prm, out)),
                 (ratio_kernel<1><<<grid_for(d, 1), kThreads, 0, st>>>(d, num, den, prm, out)));
  return cudaGetLastError();
}

cudaError_t launch_fill(size_t n, float v, float* out, cudaStream_t st) {
  size_t blocks = (n + (size_t)kThreads * 4 - 1) / ((size_t)kThreads * 4);
  blocks = blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks);
  fill_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(n, v, out);
  return cudaGetLastError();
}

cudaError_t launch_axpbypcz(size_t n, float a, const float* p, float b, const float* q, float c, const float* r, float* out,
                            cudaStream_t st) {
  size_t blocks = (n + (size_t)kThreads * 4 - 1) / ((size_t)kThreads * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  axpbypcz_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(n, a, p, b, q, c, r, out);
  return cudaGetLastError();
}

}  // namespace pds
