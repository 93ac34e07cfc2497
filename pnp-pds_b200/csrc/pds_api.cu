// C-ABI of libpnp_pds.so: handle management, weight/operator set-up, stand-alone operators and
// the resident PnP-PDS loop.  See include/pnp_pds.h for the contract and the reference
// interfaces each entry point replaces.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <cmath>
#include <cuda_fp8.h>

#include "kernels.cuh"

namespace pds {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace pds

using namespace pds;

struct pds_handle_s {
  pds_config_t cfg{};
  Dims d{};
  // state of iteration.test_iter (iteration.py:23-32)
  float* xbuf[2] = {nullptr, nullptr};
  float* u = nullptr;
  float* t = nullptr;
  float* sbuf[2] = {nullptr, nullptr};
  float* obs = nullptr;
  float* xtrue = nullptr;
  float* tmp[2] = {nullptr, nullptr};
  float* y1 = nullptr;       // TV baselines: dual variable of the difference operator, (B, 6, H, W)
  // ADMM cross-check loops (algorithm/admm.py): z, d, work buffers, adj_phi(ones), coefficient table
  float* zbuf = nullptr;
  float* dbuf = nullptr;
  float* wrk[3] = {nullptr, nullptr, nullptr};
  float* ones_adj = nullptr;
  float* coef = nullptr;          // device [8 combos][B][6]
  bool coef_ready = false;
  std::vector<ItemParams> prm_host;
  int m1 = 15, m2 = 15;
  float gamma_step1 = 0.1f;
  bool have_true = false, have_problem = false;
  int cur = 0, scur = 0, iter = 0;
  uint8_t* mask = nullptr;
  bool have_mask = false;
  ItemParams* prm = nullptr;
  bool have_params = false;
  double* sums = nullptr;      // [max_iter][B][NSUM]
  double* scratch = nullptr;   // [B]
  unsigned* mm = nullptr;      // [B][2] min/max keys for the SSIM data range
  int ssim_mode = 0;           // 0 off, 1 every iteration, 2 last iteration of each run
  bool ssim_now = false;
  int conv_engine = PDS_CONV_TCGEN05;   // pds_debug_set_conv_engine: the fp32 CUDA-core engine is a cross-check, not a product backend
  // blur
  float* blur_w_dev = nullptr;          // capacity kMaxBlurTaps, allocated in pds_create when deg_op == blur
  short2* blur_off_dev[2] = {nullptr, nullptr};
  BlurTaps taps{};
  std::vector<float> blur_w_host;
  std::vector<short2> blur_off_host[2];
  bool have_blur = false;
  // denoiser
  bool have_net = false;
  int depth = 0;
  float slope = 0.f, res_sign = 1.f;
  int clamp = 1;
  std::vector<DncnnLayerW> layers;
  std::vector<float> first_w_host, first_b_host;
  __half* act[2] = {nullptr, nullptr};
  int chunk = 1;
  TcPlan* tc = nullptr;
  int tc_variant = 0;
  // bookkeeping
  std::vector<void*> allocs;
  size_t bytes = 0;
  long long launches = 0;
  // optional per-kernel timing (bench.py roofline): event pairs around the launches of each category
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;
  // pds_restore_host: x_obsrv / x_true are uploaded on a side stream while the first primal step and denoiser pass run
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_enter = nullptr, ev_inputs = nullptr;
  // pds_restore_host, ours-A/B/C: x_0 arrives and x_{final} leaves one denoiser chunk at a time (kPipeEvents chunks at most)
  static constexpr int kPipeEvents = 64;
  cudaEvent_t ev_pipe[kPipeEvents] = {};       // x_0 slice of chunk c is on the device
  cudaEvent_t ev_out = nullptr, ev_out_done = nullptr;
  bool pipe_in = false;                       // this iteration's primal step runs per chunk, behind the chunk's upload
  float* pipe_out_host = nullptr;             // this iteration's denoiser output is copied to the host chunk by chunk
  bool inputs_pending = false;
  size_t ev_used = 0;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  double prof_ms[PDS_PROF_NCAT] = {};
  long long prof_n[PDS_PROF_NCAT] = {};
};

namespace {

template <typename T>
int dev_alloc(pds_handle_s* h, T** p, size_t count) {
  void* q = nullptr;
  size_t nb = count * sizeof(T);
  if (nb == 0) nb = sizeof(T);
  PDS_CUDA_OK(cudaMalloc(&q, nb));
  h->allocs.push_back(q);
  h->bytes += nb;
  *p = static_cast<T*>(q);
  return 0;
}

// Operand split of the tcgen05 engine (dncnn_tc.cu): w = w_hi + w_lo with w_hi fp16; the correction products run in e4m3:
//   A8 = [e4m3(a) | e4m3(a_lo 2^10)],  B8 = [e4m3(w_lo 2^S) | e4m3(w_hi 2^(S-10))],  correction = 2^-S * (A8 . B8).
// S is per layer: with max|w| < 2^e,  |w_hi| 2^(S-10) <= 256 and |w_lo| 2^S <= 2^(e-11+S) <= 256 for S = 18 - e.
struct TcSplit {
  float s_lo, s_hi;   // 2^S, 2^(S-10)
  float lo_scale;     // 2^-S
};
TcSplit tc_split_scales(const float* w, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) m = std::max(m, std::fabs(w[i]));
  int e = 0;
  if (m > 0.f) {
    std::frexp(m, &e);                 // m = f * 2^e, f in [0.5, 1)  ->  m < 2^e
  }
  e = std::max(-20, std::min(e, 14));
  const int S = 18 - e;
  return TcSplit{std::ldexp(1.f, S), std::ldexp(1.f, S - 10), std::ldexp(1.f, -S)};
}
// One weight w[row][ci] into a fp16 tile and an e4m3 tile (rows of 128 B, 16-byte chunk j of a row stored at j ^ (row & 7)).
void tc_put_weight(__half* tile16, __half* tile8, int row, int ci, float v, const TcSplit& sp) {
  const __half hi = __float2half_rn(v);
  const float lo = v - __half2float(hi);
  tile16[(size_t)row * 64 + (((ci >> 3) ^ (row & 7)) << 3) + (ci & 7)] = hi;
  uint8_t* r8 = reinterpret_cast<uint8_t*>(tile8) + (size_t)row * 128;
  const int k_lo = ci, k_hi = 64 + ci;
  r8[(((k_lo >> 4) ^ (row & 7)) << 4) + (k_lo & 15)] = (uint8_t)__nv_cvt_float_to_fp8(lo * sp.s_lo, __NV_SATFINITE, __NV_E4M3);
  r8[(((k_hi >> 4) ^ (row & 7)) << 4) + (k_hi & 15)] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(hi) * sp.s_hi, __NV_SATFINITE, __NV_E4M3);
}

constexpr int kMaxBlurTaps = 63 * 63;

int tc_num_sms_cached() {
  static int n = tc_num_sms();
  return n;
}

#define PDS_TRY(expr)         \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

#define PDS_LAUNCH(h, expr)   \
  do {                        \
    PDS_CUDA_OK(expr);        \
    (h)->launches++;          \
  } while (0)

cudaEvent_t prof_event(pds_handle_s* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}

// launch + (optionally) bracket with events on the launching stream
#define PDS_LAUNCH_P(h, cat, st, expr)                         \
  do {                                                         \
    cudaEvent_t _ea = nullptr, _eb = nullptr;                  \
    if ((h)->prof) {                                           \
      _ea = prof_event(h);                                     \
      _eb = prof_event(h);                                     \
      cudaEventRecord(_ea, st);                                \
    }                                                          \
    PDS_CUDA_OK(expr);                                         \
    (h)->launches++;                                           \
    if ((h)->prof) {                                           \
      cudaEventRecord(_eb, st);                                \
      (h)->prof_recs.push_back({cat, _ea, _eb});               \
    }                                                          \
  } while (0)

int check_handle(pds_handle_t h) {
  PDS_REQUIRE(h != nullptr, "null handle");
  PDS_CUDA_OK(cudaSetDevice(h->cfg.device));
  return 0;
}

size_t total_elems(const pds_handle_s* h) { return (size_t)h->d.B * h->d.n; }

int apply_phi(pds_handle_s* h, bool adjoint, const float* in, float* out, cudaStream_t st) {
  switch (h->cfg.deg_op) {
    case PDS_OP_ID:
      if (in != out) PDS_CUDA_OK(cudaMemcpyAsync(out, in, total_elems(h) * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return 0;
    case PDS_OP_BLUR:
      PDS_REQUIRE(h->have_blur, "blur kernel not set (pds_set_blur_kernel)");
      PDS_REQUIRE(in != out, "blur cannot run in place");
      PDS_LAUNCH(h, launch_blur_apply(h->d, h->taps, adjoint ? 1 : 0, in, out, st));
      return 0;
    case PDS_OP_RANDOM_SAMPLING:
      PDS_REQUIRE(h->have_mask, "sampling mask not set (pds_set_mask)");
      PDS_LAUNCH(h, launch_mask_apply(h->d, in, h->mask, out, st));
      return 0;
  }
  PDS_REQUIRE(false, "unknown deg_op");
}

// Which kernel serves the 64->64 body layers of a launch of nimg images (see run_dncnn for the switches in tc_variant).
struct BodyDispatch {
  int band;        // > 0: row-streaming kernel with this band height
  bool two_cta;    // tile kernels: CTA-pair form
  bool chain;      // all body layers in one persistent launch (dncnn_chain.cu)
};
BodyDispatch body_dispatch(const pds_handle_s* h, int nimg) {
  BodyDispatch b{};
  const Dims& d = h->d;
  b.band = (h->tc_variant & (128 | 32 | 16)) ? 0 : roll_band_rows(nimg, d.H, d.W, tc_num_sms_cached(), (h->tc_variant & 64) != 0);
  b.two_cta = !(h->tc_variant & 16);
  b.chain = h->conv_engine == PDS_CONV_TCGEN05 && b.band == 0 && b.two_cta && !(h->tc_variant & 512) && h->depth > 2 && h->tc &&
            tc_chain_available(h->tc, nimg);
  return b;
}

StepArgs step_args(pds_handle_s* h);
// the items [b0, b0 + nimg) of a step as a step of their own (per-item rows and state slices start at item b0)
StepArgs chunk_view(StepArgs a, int b0, int nimg) {
  const size_t off = (size_t)b0 * a.d.n;
  a.d.B = nimg;
  a.x += off;
  a.xn += off;
  a.u += off;
  a.t += off;
  if (a.s_old) a.s_old += off;
  if (a.s_new) a.s_new += off;
  a.obs += off;
  if (a.xtrue) a.xtrue += off;
  a.prm += b0;
  if (a.sums_prev) a.sums_prev += (size_t)b0 * NSUM;
  a.sums_cur += (size_t)b0 * NSUM;
  return a;
}

// Denoiser.denoise over all B items, `chunk` images per pass so that activations can stay in L2.
int run_dncnn(pds_handle_s* h, const float* in, float* out, cudaStream_t st) {
  PDS_REQUIRE(h->have_net, "denoiser weights not loaded (pds_load_dncnn)");
  const Dims& d = h->d;
  for (int b0 = 0; b0 < d.B; b0 += h->chunk) {
    const int nimg = (d.B - b0 < h->chunk) ? d.B - b0 : h->chunk;
    const float* cin = in + (size_t)b0 * d.n;
    float* cout = out + (size_t)b0 * d.n;
    if (h->pipe_in) {
      // pds_restore_host, first iteration: the primal step of this chunk's items runs here, behind their x_0 upload, so the
      // denoiser starts after one chunk's transfer instead of the whole batch's
      PDS_CUDA_OK(cudaStreamWaitEvent(st, h->ev_pipe[b0 / h->chunk], 0));
      const StepArgs ca = chunk_view(step_args(h), b0, nimg);
      if (h->cfg.deg_op == PDS_OP_BLUR) PDS_LAUNCH_P(h, PDS_PROF_PRIMAL, st, launch_primal_blur(ca, h->taps, st));
      else PDS_LAUNCH_P(h, PDS_PROF_PRIMAL, st, launch_primal_pointwise(ca, st));
    }
    // body layers: the row-streaming kernel (dncnn_roll.cu) when the width splits into 128-pixel strips and its cost model
    // (roll_band_rows) beats the tiles; else the 2-CTA tile kernel (measured faster than the 1-CTA one at every size: half
    // the weight prologue per SM, fewer operand bytes; cfg1 264 vs 285 us per iteration, cfg2 694 vs 792).  The 1-CTA tile
    // kernel stays as a cross-check.  tc_variant (pds_debug_set_tc_variant) bit 7 disables row streaming, bit 6 forces it wherever the width
    // allows, bit 4 / bit 5 force the 1-CTA / 2-CTA tile kernel, bit 8 makes the row-streaming kernel read e4m3(a) from
    // HBM instead of rebuilding it on chip (then every layer stores it).
    const BodyDispatch bd = body_dispatch(h, nimg);
    const int band = bd.band;
    const bool two_cta = bd.two_cta;
    const bool derive = band > 0 && !(h->tc_variant & 256);
    if (h->conv_engine == PDS_CONV_TCGEN05) {
      tc_plan_set_probe_bits(h->tc, (h->tc_variant >> 16) & 7);   // bits 16 - 18: timing probes (wrong results by design, include/pnp_pds.h)
      PDS_LAUNCH_P(h, PDS_PROF_CONV_FIRST, st, launch_conv_first_tc(h->tc, nimg, d.C, cin, h->layers[0], h->slope, h->clamp, derive ? 0 : 1, (h->tc_variant & 32768) ? 1 : 0, st));
    } else {
      PDS_LAUNCH_P(h, PDS_PROF_CONV_FIRST, st, launch_conv_first(nimg, d.C, d.H, d.W, cin, h->layers[0], h->slope, h->clamp, h->act[0], st));
    }
    int src = 0;
    // small launches (the tile kernels' territory): all body layers in ONE persistent launch with tile-level dataflow between
    // layers (dncnn_chain.cu).  tc_variant bit 9 disables it (per-layer tile kernels: the bit-exact cross-check).
    const bool chain = bd.chain;
    if (chain) {
      PDS_LAUNCH_P(h, PDS_PROF_CONV_MID, st, launch_conv_body_chain(h->tc, src, nimg, h->slope, (h->tc_variant & 4096) ? 2 : ((h->tc_variant & 2048) ? 1 : 0), st));
      src ^= (h->depth - 2) & 1;
    }
    for (int l = 1; l < h->depth - 1 && !chain; ++l) {
      if (h->conv_engine == PDS_CONV_TCGEN05 && band > 0) {
        // e4m3(fp16(a)) is stored only for consumers that read it from HBM (the last layer does not: it takes the fp16 plane and
        // the a_lo half of plane 1)
        const int write_a8 = derive ? 0 : 1;
        PDS_LAUNCH_P(h, PDS_PROF_CONV_MID, st, launch_conv_mid_roll(h->tc, src, nimg, band, h->layers[l], h->slope, derive ? 1 : 0, write_a8, st));
      } else if (h->conv_engine == PDS_CONV_TCGEN05 && two_cta) {
        PDS_LAUNCH_P(h, PDS_PROF_CONV_MID, st, launch_conv_mid_tc2(h->tc, src, nimg, h->layers[l], h->slope, st));
      } else if (h->conv_engine == PDS_CONV_TCGEN05) {
        PDS_LAUNCH_P(h, PDS_PROF_CONV_MID, st, launch_conv_mid_tc(h->tc, src, nimg, h->layers[l], h->slope, st));
      } else {
        PDS_LAUNCH_P(h, PDS_PROF_CONV_MID, st, launch_conv_mid_simt(nimg, d.H, d.W, h->act[src], h->layers[l], h->slope, h->act[src ^ 1], st));
      }
      src ^= 1;
    }
    if (h->conv_engine == PDS_CONV_TCGEN05) {
      PDS_LAUNCH_P(h, PDS_PROF_CONV_LAST, st, launch_conv_last_tc(h->tc, src, nimg, d.C, h->layers[h->depth - 1], cin, h->res_sign, h->clamp, cout, st));
    } else {
      PDS_LAUNCH_P(h, PDS_PROF_CONV_LAST, st, launch_conv_last(nimg, d.C, d.H, d.W, h->act[src], h->layers[h->depth - 1], cin, h->res_sign, h->clamp, cout, st));
    }
    if (h->pipe_out_host) {
      // pds_restore_host, last iteration: x_{k+1} of this chunk is final (the dual update only reads it) — on its way to the host
      // while the remaining chunks are denoised
      PDS_CUDA_OK(cudaEventRecord(h->ev_out, st));
      PDS_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_out, 0));
      PDS_CUDA_OK(cudaMemcpyAsync(h->pipe_out_host + (size_t)b0 * d.n, cout, (size_t)nimg * d.n * sizeof(float), cudaMemcpyDeviceToHost,
                                  h->copy_stream));
    }
  }
  return 0;
}

// eval_ssim(x_true, x_{k+1}) of iteration.py:189 when requested for this iteration
int post_iteration(pds_handle_s* h, const float* x_new, cudaStream_t st) {
  if (h->ssim_now && h->have_true) {
    const size_t row = (size_t)h->d.B * NSUM;
    PDS_CUDA_OK(launch_ssim(h->d, h->xtrue, x_new, h->mm, h->sums + (size_t)h->iter * row, st));
    h->launches += 2;
  }
  return 0;
}

// which of the A / B / C update rules the fused primal / dual kernels apply for this handle's method
int kernel_method(const pds_handle_s* h) {
  if (h->cfg.method == PDS_METHOD_TV_A) return PDS_METHOD_A;
  if (h->cfg.method == PDS_METHOD_TV_B3) return PDS_METHOD_B;
  return h->cfg.method;
}

StepArgs step_args(pds_handle_s* h) {
  StepArgs a{};
  a.d = h->d;
  a.method = kernel_method(h);
  a.x = h->xbuf[h->cur];
  a.xn = h->xbuf[h->cur ^ 1];
  a.u = h->u;
  a.t = h->t;
  a.s_old = h->sbuf[h->scur];
  a.s_new = h->sbuf[h->scur ^ 1];
  a.obs = h->obs;
  a.xtrue = h->have_true ? h->xtrue : nullptr;
  a.mask = (h->cfg.deg_op == PDS_OP_RANDOM_SAMPLING) ? h->mask : nullptr;
  a.prm = h->prm;
  const size_t row = (size_t)h->d.B * NSUM;
  a.sums_prev = h->iter > 0 ? h->sums + (size_t)(h->iter - 1) * row : nullptr;
  a.sums_cur = h->sums + (size_t)h->iter * row;
  return a;
}

// The first consumer of x_obsrv / x_true after pds_restore_host waits for their side-stream upload here.
int wait_inputs(pds_handle_s* h, cudaStream_t st) {
  if (h->inputs_pending) {
    PDS_CUDA_OK(cudaStreamWaitEvent(st, h->ev_inputs, 0));
    h->inputs_pending = false;
  }
  return 0;
}

// One iteration of A-/B-/C-Proposed (iteration.py:48-63).
int pds_iteration(pds_handle_s* h, cudaStream_t st) {
  StepArgs a = step_args(h);
  const bool blur = h->cfg.deg_op == PDS_OP_BLUR;
  // x_{k+1} = D(x_k - gamma1 Phi^T y_k)
  if (h->pipe_in) { /* per chunk inside run_dncnn */ }
  else if (blur) PDS_LAUNCH_P(h, PDS_PROF_PRIMAL, st, launch_primal_blur(a, h->taps, st));
  else PDS_LAUNCH_P(h, PDS_PROF_PRIMAL, st, launch_primal_pointwise(a, st));
  // s_{k+1} = P_l1(s_k - gamma1 y_k)
  if (h->cfg.method == PDS_METHOD_B)
    PDS_LAUNCH_P(h, PDS_PROF_L1BALL, st, launch_l1ball(h->d, a.s_old, h->t, h->prm, a.sums_prev, -1.f, h->sbuf[h->scur ^ 1], nullptr, st));
  PDS_TRY(run_dncnn(h, h->u, h->xbuf[h->cur ^ 1], st));
  PDS_TRY(wait_inputs(h, st));              // x_obsrv, x_true are first read by the dual update
  // y_{k+1}
  if (blur) PDS_LAUNCH_P(h, PDS_PROF_DUAL, st, launch_dual_blur(a, h->taps, st));
  else PDS_LAUNCH_P(h, PDS_PROF_DUAL, st, launch_dual_pointwise(a, st));
  PDS_TRY(post_iteration(h, a.xn, st));
  h->cur ^= 1;
  if (h->cfg.method == PDS_METHOD_B) h->scur ^= 1;
  h->iter++;
  return 0;
}

// u = x - coef * v  with coef = g1*lam per item (A-PnPFBS-DnCNN) — small helper kernel
__global__ void fbs_combine_kernel(Dims d, const float* __restrict__ x, const float* __restrict__ v, const ItemParams* prm, int mode,
                                   const float* __restrict__ dx, float* __restrict__ out) {
  const int b = blockIdx.y;
  const ItemParams p = prm[b];
  const size_t base = (size_t)b * d.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) {
    const size_t g = base + i;
    if (mode == 0) {
      // iteration.py:73: x - gamma1*myLambda*0.5*(2 Phi^T(Phi x - b))
      out[g] = fmaf(-p.g1 * p.lam, v[g], x[g]);
    } else if (mode == 2) {
      // iteration.py:97 (A-FBS-TV): x - gamma1 * Phi^T(Phi x - b); the D_T(y1) term is added by tv_primal
      out[g] = fmaf(-p.g1, v[g], x[g]);
    } else {
      // iteration.py:102-105: mu = 2/(1/g1^2 + lam); x - mu*((1/g1^2) Phi^T(Phi x - b) + lam (x - D(x)))
      const float ig = 1.f / (p.g1 * p.g1);
      const float mu = 2.f / (ig + p.lam);
      out[g] = x[g] - mu * (ig * v[g] + p.lam * (x[g] - dx[g]));
    }
  }
}

// A-PnPFBS-DnCNN (iteration.py:71-73) and A-RED-DnCNN (iteration.py:100-105)
int fbs_red_iteration(pds_handle_s* h, cudaStream_t st) {
  const Dims& d = h->d;
  float* x = h->xbuf[h->cur];
  float* xn = h->xbuf[h->cur ^ 1];
  const size_t n = total_elems(h);
  // tmp0 = Phi x - b ; tmp1 = Phi^T tmp0
  PDS_TRY(apply_phi(h, false, x, h->tmp[0], st));
  PDS_LAUNCH(h, launch_axpbypcz(n, 1.f, h->tmp[0], -1.f, h->obs, 0.f, nullptr, h->tmp[0], st));
  if (h->cfg.deg_op == PDS_OP_ID) {
    PDS_CUDA_OK(cudaMemcpyAsync(h->tmp[1], h->tmp[0], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    PDS_TRY(apply_phi(h, true, h->tmp[0], h->tmp[1], st));
  }
  dim3 grid((d.n + 1023) / 1024 > 148 * 4 ? 148 * 4 : (d.n + 1023) / 1024, d.B);
  if (h->cfg.method == PDS_METHOD_FBS) {
    fbs_combine_kernel<<<grid, 256, 0, st>>>(d, x, h->tmp[1], h->prm, 0, nullptr, h->u);
    PDS_LAUNCH(h, cudaGetLastError());
    PDS_TRY(run_dncnn(h, h->u, xn, st));
  } else {
    PDS_TRY(run_dncnn(h, x, h->u, st));   // u = D(x)
    fbs_combine_kernel<<<grid, 256, 0, st>>>(d, x, h->tmp[1], h->prm, 1, h->u, xn);
    PDS_LAUNCH(h, cudaGetLastError());
  }
  const size_t row = (size_t)d.B * NSUM;
  PDS_LAUNCH(h, launch_metrics(d, xn, x, h->have_true ? h->xtrue : nullptr, h->sums + (size_t)h->iter * row, st));
  PDS_TRY(post_iteration(h, xn, st));
  h->cur ^= 1;
  h->iter++;
  return 0;
}


// TV baselines: A-PDS-TV (iteration.py:88-94), comparisonB-3 (iteration.py:133-140), A-FBS-TV (iteration.py:95-99).
// The y2 / s updates are those of A-/B-Proposed (same fused kernels, lazy l2-ball scaling); the denoiser is replaced by
// x+ = u - gamma1 D_T(y1), followed by the y1 update.
int tv_iteration(pds_handle_s* h, cudaStream_t st) {
  const Dims& d = h->d;
  StepArgs a = step_args(h);
  float* x = h->xbuf[h->cur];
  float* xn = h->xbuf[h->cur ^ 1];
  const bool blur = h->cfg.deg_op == PDS_OP_BLUR;
  if (h->cfg.method == PDS_METHOD_TV_FBS) {
    // u = x - gamma1 Phi^T(Phi x - b)
    const size_t n = total_elems(h);
    PDS_TRY(apply_phi(h, false, x, h->tmp[0], st));
    PDS_LAUNCH(h, launch_axpbypcz(n, 1.f, h->tmp[0], -1.f, h->obs, 0.f, nullptr, h->tmp[0], st));
    if (h->cfg.deg_op == PDS_OP_ID) PDS_CUDA_OK(cudaMemcpyAsync(h->tmp[1], h->tmp[0], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else PDS_TRY(apply_phi(h, true, h->tmp[0], h->tmp[1], st));
    dim3 grid((d.n + 1023) / 1024 > 148 * 4 ? 148 * 4 : (d.n + 1023) / 1024, d.B);
    fbs_combine_kernel<<<grid, 256, 0, st>>>(d, x, h->tmp[1], h->prm, 2, nullptr, h->u);
    PDS_LAUNCH(h, cudaGetLastError());
  } else {
    // u = x - gamma1 Phi^T y2 ;  s+ = P_l1(s - gamma1 y2)
    if (blur) PDS_LAUNCH_P(h, PDS_PROF_PRIMAL, st, launch_primal_blur(a, h->taps, st));
    else PDS_LAUNCH_P(h, PDS_PROF_PRIMAL, st, launch_primal_pointwise(a, st));
    if (h->cfg.method == PDS_METHOD_TV_B3)
      PDS_LAUNCH_P(h, PDS_PROF_L1BALL, st, launch_l1ball(d, a.s_old, h->t, h->prm, a.sums_prev, -1.f, h->sbuf[h->scur ^ 1], nullptr, st));
  }
  PDS_LAUNCH(h, launch_tv_primal(d, h->u, h->y1, h->prm, xn, st));
  PDS_LAUNCH(h, launch_tv_dual(d, xn, x, h->prm, h->y1, st));
  if (h->cfg.method == PDS_METHOD_TV_FBS) {
    const size_t row = (size_t)d.B * NSUM;
    PDS_LAUNCH(h, launch_metrics(d, xn, x, h->have_true ? h->xtrue : nullptr, h->sums + (size_t)h->iter * row, st));
  } else {
    if (blur) PDS_LAUNCH_P(h, PDS_PROF_DUAL, st, launch_dual_blur(a, h->taps, st));
    else PDS_LAUNCH_P(h, PDS_PROF_DUAL, st, launch_dual_pointwise(a, st));
  }
  PDS_TRY(post_iteration(h, xn, st));
  h->cur ^= 1;
  if (h->cfg.method == PDS_METHOD_TV_B3) h->scur ^= 1;
  h->iter++;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// ADMM cross-check loops (algorithm/admm.py, iteration.py:127-132,161-172).
// Compositions of the operators above plus per-item linear combinations (launch_lincomb).
// ---------------------------------------------------------------------------------------------
enum { CB_R = 0, CB_XIN = 1, CB_S = 2, CB_V = 3, CB_Y = 4, CC_X = 0, CC_V = 1, CC_D = 2, CC_Z = 3 };

int build_coefs(pds_handle_s* h) {
  const int B = h->d.B;
  std::vector<float> c((size_t)8 * B * 6, 0.f);
  auto at = [&](int combo, int b) { return &c[((size_t)combo * B + b) * 6]; };
  for (int b = 0; b < B; ++b) {
    const ItemParams& p = h->prm_host[b];
    if (h->cfg.method == PDS_METHOD_ADMM_B2) {
      const float ig = 1.f / p.g1;
      float* q;
      q = at(CB_R, b);   q[0] = 1.f; q[1] = 1.f; q[2] = -1.f; q[3] = 1.f;          // Phi x + s - z + y          (admm.py:34,42)
      q = at(CB_XIN, b); q[0] = 1.f; q[1] = -ig;                                    // x - (1/g1) Phi^T r          (admm.py:34)
      q = at(CB_S, b);   q[0] = 1.f - ig; q[1] = -ig; q[2] = ig; q[3] = -ig;         // s - (1/g1)(Phi x + s - z + y) (admm.py:42)
      q = at(CB_V, b);   q[0] = 1.f; q[1] = 1.f; q[2] = 1.f;                        // Phi x + s + y               (iteration.py:131)
      q = at(CB_Y, b);   q[0] = 1.f; q[1] = 1.f; q[2] = 1.f; q[3] = -1.f;           // y + Phi x + s - z           (iteration.py:132)
    } else {
      const float g = h->gamma_step1, gl = g * p.lam, ga = g / p.alpha;
      float* q;
      q = at(CC_X, b); q[0] = 1.f - gl; q[1] = ga; q[2] = -ga; q[3] = gl; q[4] = -gl;  // x - g*grad                (admm.py:12-13)
      q = at(CC_V, b); q[0] = 1.f; q[1] = 1.f;                                         // x + d                     (iteration.py:163)
      q = at(CC_D, b); q[0] = 1.f; q[1] = 1.f; q[2] = -1.f;                            // d + x - z                 (iteration.py:164)
      q = at(CC_Z, b); q[0] = p.g1 / (p.lam + p.g1); q[1] = p.lam / (p.lam + p.g1);    // (g1 D(z) + lam z*)/(lam+g1) (admm.py:27)
    }
  }
  PDS_CUDA_OK(cudaMemcpy(h->coef, c.data(), c.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->coef_ready = true;
  return 0;
}

int lin(pds_handle_s* h, int combo, int nterms, float* out, const float* i0, const float* i1, const float* i2, const float* i3,
        const float* i4, cudaStream_t st) {
  LinArgs a{};
  a.d = h->d;
  a.in[0] = i0; a.in[1] = i1; a.in[2] = i2; a.in[3] = i3; a.in[4] = i4; a.in[5] = nullptr;
  a.out = out;
  a.coef = h->coef + (size_t)combo * h->d.B * 6;
  a.nterms = nterms;
  PDS_LAUNCH(h, launch_lincomb(a, st));
  return 0;
}

int phi_into(pds_handle_s* h, bool adjoint, const float* in, float* out, cudaStream_t st) { return apply_phi(h, adjoint, in, out, st); }

// comparisonB-2 (iteration.py:127-132 + admm.py:30-44).  State: x (xbuf), s (sbuf), z, y (= t, stored directly).
int admm_b2_iteration(pds_handle_s* h, cudaStream_t st) {
  const size_t n = total_elems(h);
  float* x_prev = h->xbuf[h->cur];
  float* x = h->xbuf[h->cur ^ 1];
  float* s = h->sbuf[h->scur];
  float* y = h->t;
  float* z = h->zbuf;
  float *w0 = h->wrk[0], *w1 = h->wrk[1], *w2 = h->wrk[2];
  // x-step: m1 trips from x = 1
  PDS_LAUNCH(h, launch_fill(n, 1.f, x, st));
  for (int i = 0; i < h->m1; ++i) {
    PDS_TRY(phi_into(h, false, x, w0, st));
    PDS_TRY(lin(h, CB_R, 4, w0, w0, s, z, y, nullptr, st));
    PDS_TRY(phi_into(h, true, w0, w1, st));
    PDS_TRY(lin(h, CB_XIN, 2, w2, x, w1, nullptr, nullptr, nullptr, st));
    PDS_TRY(run_dncnn(h, w2, x, st));
  }
  // s-step: m2 trips from s = 1 (Phi x fixed)
  PDS_TRY(phi_into(h, false, x, w0, st));
  float* sn = h->sbuf[h->scur ^ 1];
  PDS_LAUNCH(h, launch_fill(n, 1.f, w1, st));
  for (int i = 0; i < h->m2; ++i) {
    PDS_TRY(lin(h, CB_S, 4, w2, w1, w0, z, y, nullptr, st));
    PDS_LAUNCH(h, launch_l1ball(h->d, w2, nullptr, h->prm, nullptr, -1.f, w1, nullptr, st));
  }
  PDS_CUDA_OK(cudaMemcpyAsync(sn, w1, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // z = P_l2(Phi x + s + y) ; y += Phi x + s - z
  PDS_TRY(lin(h, CB_V, 3, w2, w0, sn, y, nullptr, nullptr, st));
  PDS_CUDA_OK(cudaMemsetAsync(h->scratch, 0, (size_t)h->d.B * sizeof(double), st));
  PDS_LAUNCH(h, launch_diff_norm2(h->d, w2, h->obs, h->scratch, st));
  PDS_LAUNCH(h, launch_proj_l2_items(h->d, w2, h->obs, h->prm, h->scratch, z, st));
  PDS_TRY(lin(h, CB_Y, 4, y, y, w0, sn, z, nullptr, st));
  const size_t row = (size_t)h->d.B * NSUM;
  PDS_LAUNCH(h, launch_metrics(h->d, x, x_prev, h->have_true ? h->xtrue : nullptr, h->sums + (size_t)h->iter * row, st));
  PDS_TRY(post_iteration(h, x, st));
  h->cur ^= 1;
  h->scur ^= 1;
  h->iter++;
  return 0;
}

// C-PnPADMM-DnCNN (iteration.py:161-165) and C-RED-DnCNN (iteration.py:166-172) + admm.py:4-28.
int admm_c_iteration(pds_handle_s* h, cudaStream_t st) {
  const size_t n = total_elems(h);
  float* x_prev = h->xbuf[h->cur];
  float* x = h->xbuf[h->cur ^ 1];
  float* z = h->zbuf;
  float* dd = h->dbuf;
  float *w0 = h->wrk[0], *w1 = h->wrk[1], *w2 = h->wrk[2];
  PDS_LAUNCH(h, launch_fill(n, 1.f, x, st));
  for (int i = 0; i < h->m1; ++i) {
    PDS_TRY(phi_into(h, false, x, w0, st));
    PDS_LAUNCH(h, launch_ratio(h->d, h->obs, w0, h->prm, w0, st));            // y / (alpha Phi x)
    PDS_TRY(phi_into(h, true, w0, w1, st));
    PDS_TRY(lin(h, CC_X, 5, x, x, w1, h->ones_adj, z, dd, st));
  }
  if (h->cfg.method == PDS_METHOD_ADMM_C) {
    PDS_TRY(lin(h, CC_V, 2, w2, x, dd, nullptr, nullptr, nullptr, st));
    PDS_TRY(run_dncnn(h, w2, z, st));
  } else {
    PDS_TRY(lin(h, CC_V, 2, w2, x, dd, nullptr, nullptr, nullptr, st));        // z* = x + d
    for (int i = 0; i < h->m2; ++i) {
      PDS_TRY(run_dncnn(h, z, w0, st));
      PDS_TRY(lin(h, CC_Z, 2, z, w0, w2, nullptr, nullptr, nullptr, st));
    }
  }
  PDS_TRY(lin(h, CC_D, 3, dd, dd, x, z, nullptr, nullptr, st));
  const size_t row = (size_t)h->d.B * NSUM;
  PDS_LAUNCH(h, launch_metrics(h->d, x, x_prev, h->have_true ? h->xtrue : nullptr, h->sums + (size_t)h->iter * row, st));
  PDS_TRY(post_iteration(h, x, st));
  h->cur ^= 1;
  h->iter++;
  return 0;
}

int admm_prepare(pds_handle_s* h, cudaStream_t st) {
  if (!h->coef_ready) PDS_TRY(build_coefs(h));
  if (h->iter == 0) {
    const size_t nb = total_elems(h) * sizeof(float);
    PDS_CUDA_OK(cudaMemsetAsync(h->zbuf, 0, nb, st));
    PDS_CUDA_OK(cudaMemsetAsync(h->dbuf, 0, nb, st));
    if (h->cfg.method != PDS_METHOD_ADMM_B2) {
      PDS_LAUNCH(h, launch_fill(total_elems(h), 1.f, h->wrk[2], st));
      PDS_TRY(phi_into(h, true, h->wrk[2], h->ones_adj, st));               // adj_phi(ones)  (admm.py:12)
    }
  }
  return 0;
}

}  // namespace

extern "C" {

const char* pds_last_error(void) { return g_err.c_str(); }
int pds_abi_version(void) { return PDS_ABI_VERSION; }

int pds_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pds_create(const pds_config_t* cfg, pds_handle_t* out) {
  PDS_REQUIRE(cfg && out, "null argument");
  PDS_REQUIRE(cfg->batch >= 1 && cfg->height >= 1 && cfg->width >= 1, "bad shape");
  PDS_REQUIRE(cfg->channels == 1 || cfg->channels == 3, "channels must be 1 or 3 (reference ch)");
  PDS_REQUIRE(cfg->method >= PDS_METHOD_A && cfg->method <= PDS_METHOD_TV_FBS, "unknown method");
  PDS_REQUIRE(cfg->method < PDS_METHOD_TV_A || (cfg->channels == 3 && cfg->height >= 2 && cfg->width >= 2),
              "the TV baselines are defined for colour images only (operators.py:122-123)");
  PDS_REQUIRE(cfg->deg_op >= PDS_OP_ID && cfg->deg_op <= PDS_OP_RANDOM_SAMPLING, "unknown deg_op");
  PDS_REQUIRE(cfg->max_iter >= 1, "max_iter must be >= 1");
  PDS_REQUIRE(cfg->reserved == 0, "pds_config_t.reserved must be 0");
  PDS_REQUIRE((long long)cfg->batch * cfg->channels <= 65535, "batch*channels exceeds the grid limit");
  PDS_REQUIRE((long long)cfg->channels * cfg->height * cfg->width <= 0x7fffffffLL, "C*H*W exceeds 2^31-1 elements per item");
  PDS_REQUIRE(pds_device_count() > cfg->device && cfg->device >= 0,
              "no usable CUDA device: this library has no CPU fallback");
  PDS_CUDA_OK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop{};
  PDS_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  PDS_REQUIRE(prop.major == 10, "libpnp_pds is built for sm_100a (B200) only");
  pds_handle_s* h = new (std::nothrow) pds_handle_s();
  PDS_REQUIRE(h, "out of host memory");
  h->cfg = *cfg;
  h->d = Dims{cfg->batch, cfg->channels, cfg->height, cfg->width, cfg->height * cfg->width,
              cfg->channels * cfg->height * cfg->width};
  const size_t n = total_elems(h);
  int rc = 0;
  auto A = [&](auto** p, size_t c) { if (!rc) rc = dev_alloc(h, p, c); };
  A(&h->xbuf[0], n); A(&h->xbuf[1], n); A(&h->u, n); A(&h->t, n);
  A(&h->obs, n); A(&h->xtrue, n);
  if (cfg->method == PDS_METHOD_B) { A(&h->sbuf[0], n); A(&h->sbuf[1], n); }
  if (cfg->method == PDS_METHOD_FBS || cfg->method == PDS_METHOD_RED) { A(&h->tmp[0], n); A(&h->tmp[1], n); }
  if (cfg->method >= PDS_METHOD_TV_A) {
    A(&h->y1, 2 * n);
    if (cfg->method == PDS_METHOD_TV_B3) { A(&h->sbuf[0], n); A(&h->sbuf[1], n); }
    if (cfg->method == PDS_METHOD_TV_FBS) { A(&h->tmp[0], n); A(&h->tmp[1], n); }
  } else if (cfg->method >= PDS_METHOD_ADMM_B2) {
    A(&h->zbuf, n); A(&h->dbuf, n); A(&h->wrk[0], n); A(&h->wrk[1], n); A(&h->wrk[2], n); A(&h->ones_adj, n);
    A(&h->coef, (size_t)8 * cfg->batch * 6);
    if (cfg->method == PDS_METHOD_ADMM_B2 && !h->sbuf[0]) { A(&h->sbuf[0], n); A(&h->sbuf[1], n); }
  }
  A(&h->mask, (size_t)h->d.hw);
  A(&h->prm, (size_t)cfg->batch);
  A(&h->sums, (size_t)cfg->max_iter * cfg->batch * NSUM);
  A(&h->scratch, (size_t)cfg->batch);
  A(&h->mm, (size_t)cfg->batch * 2);
  if (cfg->deg_op == PDS_OP_BLUR) {        // tap tables of pds_set_blur_kernel (l <= 63): nothing is allocated after create / load
    A(&h->blur_w_dev, (size_t)kMaxBlurTaps); A(&h->blur_off_dev[0], (size_t)kMaxBlurTaps); A(&h->blur_off_dev[1], (size_t)kMaxBlurTaps);
  }
  if (rc) { pds_destroy(h); return rc; }
  // side stream + events of pds_restore_host (x_obsrv / x_true upload overlapped with the first primal step)
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_enter, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_inputs, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_out_done, cudaEventDisableTiming) != cudaSuccess) {
    pds_destroy(h);
    PDS_REQUIRE(false, "could not create the copy stream / events");
  }
  for (int i = 0; i < pds_handle_s::kPipeEvents; ++i)
    if (cudaEventCreateWithFlags(&h->ev_pipe[i], cudaEventDisableTiming) != cudaSuccess) {
      pds_destroy(h);
      PDS_REQUIRE(false, "could not create the chunk events");
    }
  *out = h;
  return 0;
}

int pds_destroy(pds_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->tc) tc_plan_destroy(h->tc);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev_enter) cudaEventDestroy(h->ev_enter);
  if (h->ev_inputs) cudaEventDestroy(h->ev_inputs);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  if (h->ev_out_done) cudaEventDestroy(h->ev_out_done);
  for (cudaEvent_t e : h->ev_pipe)
    if (e) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return 0;
}

int pds_set_blur_kernel(pds_handle_t h, const double* k, int l) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(k && l >= 1 && l <= 63 && (l % 2 == 1), "blur kernel must be l x l with odd l <= 63");
  PDS_REQUIRE(h->cfg.deg_op == PDS_OP_BLUR, "the handle was not created with deg_op = blur");
  const int c = l / 2;
  std::vector<float> w;
  std::vector<short2> off[2];
  int ry = 0, rx = 0;
  for (int a = 0; a < l; ++a)
    for (int b = 0; b < l; ++b) {
      const double v = k[a * l + b];
      if (v == 0.0) continue;
      w.push_back((float)v);
      off[0].push_back(make_short2((short)(c - a), (short)(c - b)));  // Phi   (operators.py:7-22)
      off[1].push_back(make_short2((short)(a - c), (short)(b - c)));  // Phi^T (operators.py:24-38)
      ry = std::max(ry, std::abs(a - c));
      rx = std::max(rx, std::abs(b - c));
    }
  PDS_REQUIRE(!w.empty(), "blur kernel is all zeros");
  float* dw = h->blur_w_dev;               // capacity kMaxBlurTaps >= l*l (may be called again: the tables are overwritten)
  PDS_CUDA_OK(cudaMemcpy(dw, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  for (int q = 0; q < 2; ++q) {
    short2* dof = h->blur_off_dev[q];
    PDS_CUDA_OK(cudaMemcpy(dof, off[q].data(), off[q].size() * sizeof(short2), cudaMemcpyHostToDevice));
    h->taps.w[q] = dw;
    h->taps.off[q] = dof;
  }
  h->blur_w_host = w;
  h->blur_off_host[0] = off[0];
  h->blur_off_host[1] = off[1];
  h->taps.w_host = h->blur_w_host.data();
  h->taps.off_host[0] = h->blur_off_host[0].data();
  h->taps.off_host[1] = h->blur_off_host[1].data();
  h->taps.ntaps = (int)w.size();
  h->taps.ry = ry;
  h->taps.rx = rx;
  h->have_blur = true;
  return 0;
}

int pds_set_mask(pds_handle_t h, const uint8_t* mask_host) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(mask_host, "null mask");
  PDS_CUDA_OK(cudaMemcpy(h->mask, mask_host, (size_t)h->d.hw, cudaMemcpyHostToDevice));
  h->have_mask = true;
  return 0;
}

int pds_set_item_params(pds_handle_t h, const pds_item_params_t* p, int n) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(p && (n == 1 || n == h->d.B), "params: n must be 1 or batch");
  std::vector<ItemParams> v(h->d.B);
  for (int b = 0; b < h->d.B; ++b) {
    const pds_item_params_t& q = p[n == 1 ? 0 : b];
    v[b] = ItemParams{q.gamma1, q.gamma2, q.epsilon, q.eta, q.lambda, q.alpha};
  }
  PDS_CUDA_OK(cudaMemcpy(h->prm, v.data(), v.size() * sizeof(ItemParams), cudaMemcpyHostToDevice));
  h->prm_host = v;
  h->coef_ready = false;
  h->have_params = true;
  return 0;
}

int pds_load_dncnn(pds_handle_t h, const void* blob, size_t nbytes) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(!h->have_net, "denoiser already loaded");
  PDS_REQUIRE(blob && nbytes >= 48, "weight blob too small");
  const unsigned char* p = static_cast<const unsigned char*>(blob);
  PDS_REQUIRE(std::memcmp(p, "PDSW", 4) == 0, "not a PDSW weight blob");
  int32_t hdr[5];
  std::memcpy(hdr, p + 4, sizeof(hdr));
  float fl[2];
  std::memcpy(fl, p + 24, sizeof(fl));
  int32_t clampv;
  std::memcpy(&clampv, p + 32, 4);
  const int ver = hdr[0], depth = hdr[1], cin = hdr[2], nch = hdr[3], cout = hdr[4];
  PDS_REQUIRE(ver == 1, "unsupported PDSW version");
  PDS_REQUIRE(depth >= 3 && depth <= 64, "unsupported depth");
  PDS_REQUIRE(nch == kMid, "only 64-channel DnCNN bodies are supported (all reference checkpoints)");
  PDS_REQUIRE(cin == h->d.C && cout == h->d.C, "checkpoint channel count does not match the handle (ch)");
  size_t need = 48;
  for (int l = 0; l < depth; ++l) {
    const int ci = l == 0 ? cin : nch, co = l == depth - 1 ? cout : nch;
    need += (size_t)(co * ci * 9 + co) * 4;
  }
  PDS_REQUIRE(need == nbytes, "PDSW blob size mismatch");
  {
    // the fp16 / e4m3 operand split of the tcgen05 engine covers |w| < 2^14 (tc_split_scales); every shipped checkpoint has |w| < 4
    const float* wv = reinterpret_cast<const float*>(p + 48);
    const size_t nw = (nbytes - 48) / 4;
    bool ok = true;
    for (size_t i = 0; i < nw; ++i) ok = ok && std::isfinite(wv[i]) && std::fabs(wv[i]) < 16384.f;
    PDS_REQUIRE(ok, "checkpoint holds non-finite or huge (>= 2^14) parameters");
  }
  h->depth = depth;
  h->slope = fl[0];
  h->res_sign = fl[1];
  h->clamp = clampv;
  h->layers.assign(depth, DncnnLayerW{});
  const float* src = reinterpret_cast<const float*>(p + 48);
  std::vector<float> buf;
  for (int l = 0; l < depth; ++l) {
    const int ci = l == 0 ? cin : nch, co = l == depth - 1 ? cout : nch;
    const float* w = src;                // [co][ci][3][3]
    const float* b = src + (size_t)co * ci * 9;
    src = b + co;
    DncnnLayerW& L = h->layers[l];
    float* db = nullptr;
    PDS_TRY(dev_alloc(h, &db, (size_t)co));
    PDS_CUDA_OK(cudaMemcpy(db, b, (size_t)co * 4, cudaMemcpyHostToDevice));
    L.bias = db;
    if (l == 0) {
      buf.assign((size_t)9 * ci * 64, 0.f);
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int tp = 0; tp < 9; ++tp) buf[((size_t)tp * ci + c) * 64 + o] = w[((size_t)o * ci + c) * 9 + tp];
      {
        // tcgen05 first layer: rows 0-63 = w_hi[oc], rows 64-127 = w_lo[oc]; half k of a row = weight of (tap, ci) with
        // k = tap*Cin + ci (< 27), zero beyond; 16-byte chunk j stored at j ^ (row & 7)
        std::vector<__half> img((size_t)128 * 64, __float2half_rn(0.f));
        for (int o = 0; o < co; ++o)
          for (int c = 0; c < ci; ++c)
            for (int tp = 0; tp < 9; ++tp) {
              const int kk = tp * ci + c;
              const float v = w[((size_t)o * ci + c) * 9 + tp];
              const __half hi = __float2half_rn(v);
              const __half lo = __float2half_rn(v - __half2float(hi));
              const int chunk = (kk >> 3) ^ (o & 7);
              img[(size_t)o * 64 + chunk * 8 + (kk & 7)] = hi;
              img[(size_t)(64 + o) * 64 + chunk * 8 + (kk & 7)] = lo;
            }
        __half* dh = nullptr;
        PDS_TRY(dev_alloc(h, &dh, img.size()));
        PDS_CUDA_OK(cudaMemcpy(dh, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
        L.w_first_tc = dh;
      }
      {
        // tap-shifted first layer (dncnn_tc.cu first2): [tap][K chunk][oc][8 halves]; K-slot s of a B row:
        //   [0,ci) w_hi, [ci,2ci) w_hi, [2ci,3ci) w_lo, 3ci..3ci+2 = the bias as three fp16 terms (centre tap only), rest 0
        std::vector<__half> img((size_t)9 * 2 * 64 * 8, __float2half_rn(0.f));
        auto put = [&](int tp, int o, int s, float v) { img[(((size_t)tp * 2 + (s >> 3)) * 64 + o) * 8 + (s & 7)] = __float2half_rn(v); };
        for (int o = 0; o < co; ++o) {
          for (int c = 0; c < ci; ++c)
            for (int tp = 0; tp < 9; ++tp) {
              const float v = w[((size_t)o * ci + c) * 9 + tp];
              const float hi = __half2float(__float2half_rn(v));
              put(tp, o, c, hi);
              put(tp, o, ci + c, hi);
              put(tp, o, 2 * ci + c, v - hi);
            }
          float rem = b[o];
          for (int k = 0; k < 3; ++k) {
            const float t = __half2float(__float2half_rn(rem));
            put(4, o, 3 * ci + k, t);
            rem -= t;
          }
        }
        __half* dh = nullptr;
        PDS_TRY(dev_alloc(h, &dh, img.size()));
        PDS_CUDA_OK(cudaMemcpy(dh, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
        L.w_first_tc2 = dh;
      }
      h->first_w_host = buf;
      h->first_b_host.assign(b, b + co);
      L.w_first_host = h->first_w_host.data();
      L.bias_host = h->first_b_host.data();
    } else if (l == depth - 1) {
      buf.assign((size_t)co * 9 * 64, 0.f);
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int tp = 0; tp < 9; ++tp) buf[((size_t)o * 9 + tp) * 64 + c] = w[((size_t)o * ci + c) * 9 + tp];
      float* dw = nullptr;
      PDS_TRY(dev_alloc(h, &dw, buf.size()));
      PDS_CUDA_OK(cudaMemcpy(dw, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice));
      L.w_last = dw;
      // tcgen05 engine (dncnn_tc.cu namespace last): B row n = tap*Cout + c holds W[c][tap][ci 0..63], all nine taps side by side
      // along N, rows >= 9*Cout zero.  fp16 tile: 64 rows x 128 B, rows 0-31 w_hi[n], rows 32-63 (w_lo 2^S)[n], 16-byte chunk j of
      // row r at j ^ (r & 7) (SWIZZLE_128B); then the e4m3 tile: 32 rows x 64 B of e4m3(w_hi 2^(S-10))[n], chunk j at
      // j ^ ((r >> 1) & 3) (SWIZZLE_64B)
      const TcSplit sp = tc_split_scales(w, (size_t)co * ci * 9);
      L.lo_scale = sp.lo_scale;
      std::vector<__half> img((size_t)64 * 64 + 32 * 32, __float2half_rn(0.f));
      uint8_t* img8 = reinterpret_cast<uint8_t*>(img.data() + (size_t)64 * 64);
      for (int tp = 0; tp < 9; ++tp)
        for (int o = 0; o < co; ++o)
          for (int c = 0; c < 64; ++c) {
            const int n = tp * co + o;
            const float v = w[((size_t)o * ci + c) * 9 + tp];
            const __half hi = __float2half_rn(v);
            const float lo = v - __half2float(hi);
            img[(size_t)n * 64 + (((c >> 3) ^ (n & 7)) << 3) + (c & 7)] = hi;
            img[(size_t)(32 + n) * 64 + (((c >> 3) ^ ((32 + n) & 7)) << 3) + (c & 7)] = __float2half_rn(lo * sp.s_lo);
            img8[(size_t)n * 64 + (((c >> 4) ^ ((n >> 1) & 3)) << 4) + (c & 15)] =
                (uint8_t)__nv_cvt_float_to_fp8(__half2float(hi) * sp.s_hi, __NV_SATFINITE, __NV_E4M3);
          }
      __half* dh = nullptr;
      PDS_TRY(dev_alloc(h, &dh, img.size()));
      PDS_CUDA_OK(cudaMemcpy(dh, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
      L.w_last_tc = dh;
    } else {
      // SIMT engine: [ci][tap][oc] fp32
      buf.assign((size_t)64 * 9 * 64, 0.f);
      for (int o = 0; o < 64; ++o)
        for (int c = 0; c < 64; ++c)
          for (int tp = 0; tp < 9; ++tp) buf[((size_t)c * 9 + tp) * 64 + o] = w[((size_t)o * 64 + c) * 9 + tp];
      float* dw = nullptr;
      PDS_TRY(dev_alloc(h, &dw, buf.size()));
      PDS_CUDA_OK(cudaMemcpy(dw, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice));
      L.w_mid = dw;
      // tcgen05 engine: shared-memory image [tap][fp16 tile | e4m3 tile][oc][128 B]: K-major rows of 128 B,
      // 16-byte chunk j of row oc stored at chunk (j ^ (oc & 7))  (SWIZZLE_128B).
      //   fp16 tile row = w_hi[oc][ci 0..63];  e4m3 tile row = [e4m3(w_lo 2^S)[ci 0..63] | e4m3(w_hi 2^(S-10))[ci 0..63]]
      const TcSplit sp = tc_split_scales(w, (size_t)64 * 64 * 9);
      L.lo_scale = sp.lo_scale;
      std::vector<__half> img((size_t)2 * 9 * 64 * 64);
      for (int tp = 0; tp < 9; ++tp)
        for (int o = 0; o < 64; ++o)
          for (int c = 0; c < 64; ++c)
            tc_put_weight(img.data() + (size_t)(tp * 2) * 64 * 64, img.data() + (size_t)(tp * 2 + 1) * 64 * 64, o, c,
                          w[((size_t)o * 64 + c) * 9 + tp], sp);
      __half* dh = nullptr;
      PDS_TRY(dev_alloc(h, &dh, img.size()));
      PDS_CUDA_OK(cudaMemcpy(dh, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
      L.w_mid_tc = dh;
      // 2-CTA engine: CTA r keeps, per tap, the fp16 tile of oc [32r, 32r+32) followed by their e4m3 tile
      // (2 x 32 rows of 128 B, swizzled by local row)
      std::vector<__half> img2((size_t)2 * 9 * 64 * 64);
      for (int r = 0; r < 2; ++r)
        for (int tp = 0; tp < 9; ++tp)
          for (int lr = 0; lr < 32; ++lr)
            for (int c = 0; c < 64; ++c) {
              __half* t16 = img2.data() + ((size_t)r * 9 + tp) * 64 * 64;
              tc_put_weight(t16, t16 + 32 * 64, lr, c, w[((size_t)(32 * r + lr) * 64 + c) * 9 + tp], sp);
            }
      __half* dh2 = nullptr;
      PDS_TRY(dev_alloc(h, &dh2, img2.size()));
      PDS_CUDA_OK(cudaMemcpy(dh2, img2.data(), img2.size() * sizeof(__half), cudaMemcpyHostToDevice));
      L.w_mid_tc2 = dh2;
    }
  }
  // activation buffers
  const size_t px = (size_t)h->d.hw;
  int chunk = h->cfg.denoiser_chunk;
  if (chunk <= 0) {
    const size_t target_px = (size_t)8 << 20;   // 8 Mpx per pass = 2 GiB per activation buffer
    chunk = (int)std::max<size_t>(1, target_px / px);
  }
  chunk = std::min(chunk, h->d.B);
  h->chunk = chunk;
  const size_t act_elems = (size_t)chunk * 2 * px * 64;
  PDS_TRY(dev_alloc(h, &h->act[0], act_elems));
  PDS_TRY(dev_alloc(h, &h->act[1], act_elems));
  if (h->conv_engine == PDS_CONV_TCGEN05) {
    PDS_TRY(tc_plan_create(chunk, h->d.H, h->d.W, h->act[0], h->act[1], &h->tc));
    size_t chain_bytes = 0;
    PDS_TRY(tc_plan_chain(h->tc, h->layers.data(), depth, &chain_bytes));
    h->bytes += chain_bytes;
  }
  h->have_net = true;
  return 0;
}

int pds_phi(pds_handle_t h, const float* in, float* out, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  return apply_phi(h, false, in, out, (cudaStream_t)stream);
}

int pds_phi_adj(pds_handle_t h, const float* in, float* out, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  return apply_phi(h, true, in, out, (cudaStream_t)stream);
}

int pds_proj_l2_ball(pds_handle_t h, const float* x, const float* c, float eps, float* out, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  cudaStream_t st = (cudaStream_t)stream;
  PDS_CUDA_OK(cudaMemsetAsync(h->scratch, 0, (size_t)h->d.B * sizeof(double), st));
  PDS_LAUNCH(h, launch_diff_norm2(h->d, x, c, h->scratch, st));
  PDS_LAUNCH(h, launch_proj_l2_apply(h->d, x, c, eps, h->scratch, out, st));
  return 0;
}

int pds_proj_l1_ball(pds_handle_t h, const float* x, float eta, float* out, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(x != out, "proj_l1_ball cannot run in place");
  PDS_REQUIRE(eta >= 0.f, "eta must be >= 0");
  PDS_LAUNCH(h, launch_l1ball(h->d, x, nullptr, h->prm, nullptr, eta, out, nullptr, (cudaStream_t)stream));
  return 0;
}

int pds_prox_gkl(pds_handle_t h, const float* x, const float* x0, float gamma, float alpha, float* out, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_LAUNCH(h, launch_prox_gkl(h->d, x, x0, gamma, alpha, out, (cudaStream_t)stream));
  return 0;
}

int pds_dncnn_forward(pds_handle_t h, const float* in, float* out, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(in != out, "dncnn_forward cannot run in place");
  return run_dncnn(h, in, out, (cudaStream_t)stream);
}

int pds_set_problem(pds_handle_t h, const float* x0, const float* obs, const float* xtrue, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(x0 && obs, "x0 and obs are required");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = total_elems(h) * sizeof(float);
  h->cur = h->scur = h->iter = 0;
  PDS_CUDA_OK(cudaMemcpyAsync(h->xbuf[0], x0, nb, cudaMemcpyDeviceToDevice, st));
  PDS_CUDA_OK(cudaMemcpyAsync(h->obs, obs, nb, cudaMemcpyDeviceToDevice, st));
  h->have_true = xtrue != nullptr;
  if (xtrue) PDS_CUDA_OK(cudaMemcpyAsync(h->xtrue, xtrue, nb, cudaMemcpyDeviceToDevice, st));
  PDS_CUDA_OK(cudaMemsetAsync(h->t, 0, nb, st));
  if (h->y1) PDS_CUDA_OK(cudaMemsetAsync(h->y1, 0, 2 * nb, st));
  if (h->sbuf[0]) PDS_CUDA_OK(cudaMemsetAsync(h->sbuf[0], 0, nb, st));
  PDS_CUDA_OK(cudaMemsetAsync(h->sums, 0, (size_t)h->cfg.max_iter * h->d.B * NSUM * sizeof(double), st));
  h->have_problem = true;
  return 0;
}

int pds_run(pds_handle_t h, int n_iter, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(h->have_problem, "pds_set_problem has not been called");
  PDS_REQUIRE(h->have_params, "pds_set_item_params has not been called");
  PDS_REQUIRE(n_iter >= 0 && h->iter + n_iter <= h->cfg.max_iter, "n_iter exceeds the max_iter the handle was created with");
  cudaStream_t st = (cudaStream_t)stream;
  float* const host_out = h->pipe_out_host;         // set by pds_restore_host for this run only
  const bool pipe_first = h->pipe_in;
  for (int i = 0; i < n_iter; ++i) {
    h->pipe_in = pipe_first && i == 0;
    h->pipe_out_host = (i == n_iter - 1) ? host_out : nullptr;
    h->ssim_now = h->ssim_mode == 1 || (h->ssim_mode == 2 && i == n_iter - 1);
    if (h->cfg.method > PDS_METHOD_C) PDS_TRY(wait_inputs(h, st));    // every other loop reads x_obsrv at its first step
    if (h->cfg.method <= PDS_METHOD_C) PDS_TRY(pds_iteration(h, st));
    else if (h->cfg.method <= PDS_METHOD_RED) PDS_TRY(fbs_red_iteration(h, st));
    else if (h->cfg.method >= PDS_METHOD_TV_A) PDS_TRY(tv_iteration(h, st));
    else {
      PDS_TRY(admm_prepare(h, st));
      if (h->cfg.method == PDS_METHOD_ADMM_B2) PDS_TRY(admm_b2_iteration(h, st));
      else PDS_TRY(admm_c_iteration(h, st));
    }
  }
  h->pipe_in = false;
  h->pipe_out_host = nullptr;
  return 0;
}

int pds_set_ssim(pds_handle_t h, int mode) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(mode >= 0 && mode <= 2, "ssim mode must be 0 (off), 1 (every iteration) or 2 (last iteration of each run)");
  h->ssim_mode = mode;
  return 0;
}

int pds_set_admm(pds_handle_t h, int m1, int m2, float gamma_step1) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(m1 >= 0 && m2 >= 0, "m1, m2 must be >= 0");
  h->m1 = m1;
  h->m2 = m2;
  h->gamma_step1 = gamma_step1;
  h->coef_ready = false;
  return 0;
}

int pds_iterations_done(pds_handle_t h) { return h ? h->iter : -1; }

int pds_get_state(pds_handle_t h, float* x, float* s, float* y, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(h->have_problem, "no problem set");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = total_elems(h) * sizeof(float);
  if (x) PDS_CUDA_OK(cudaMemcpyAsync(x, h->xbuf[h->cur], nb, cudaMemcpyDeviceToDevice, st));
  if (s) {
    if (h->sbuf[0]) PDS_CUDA_OK(cudaMemcpyAsync(s, h->sbuf[h->scur], nb, cudaMemcpyDeviceToDevice, st));
    else PDS_CUDA_OK(cudaMemsetAsync(s, 0, nb, st));
  }
  if (y) {
    const size_t row = (size_t)h->d.B * NSUM;
    const double* sums = h->iter > 0 ? h->sums + (size_t)(h->iter - 1) * row : nullptr;
    if (kernel_method(h) <= PDS_METHOD_B) PDS_LAUNCH(h, launch_scale_by_sigma(h->d, h->t, h->prm, sums, kernel_method(h), y, st));
    else PDS_CUDA_OK(cudaMemcpyAsync(y, h->t, nb, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int pds_get_traces(pds_handle_t h, double* trace_host, size_t cap, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  const size_t n = (size_t)h->iter * h->d.B * NSUM;
  PDS_REQUIRE(trace_host && cap >= n, "trace buffer too small");
  cudaStream_t st = (cudaStream_t)stream;
  PDS_CUDA_OK(cudaMemcpyAsync(trace_host, h->sums, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  PDS_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int pds_restore_host(pds_handle_t h, const float* x0, const float* obs, const float* xtrue, int n_iter, float* x_out, float* s_out,
                     double* trace_host, size_t trace_cap, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(x0 && obs && x_out, "x0, obs and x_out are required");
  PDS_REQUIRE(n_iter >= 0 && n_iter <= h->cfg.max_iter, "n_iter exceeds max_iter");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = total_elems(h) * sizeof(float);
  h->cur = h->scur = h->iter = 0;
  // ours-A/B/C with a denoiser: x_0 goes up one denoiser chunk at a time on the side stream and the first iteration's primal step
  // and denoiser pass follow chunk by chunk (run_dncnn); in the last iteration every chunk of x_{k+1} leaves for the host as soon as
  // its last layer is done.  Otherwise x_0 goes up on the caller's stream.  x_obsrv and x_true always follow on the side stream,
  // ordered after whatever the caller's stream was still doing with those buffers, and are joined by the first kernel that reads
  // them (wait_inputs).
  const int nchunks = h->have_net && h->chunk > 0 ? (h->d.B + h->chunk - 1) / h->chunk : 0;
  const bool pipe = h->cfg.method <= PDS_METHOD_C && h->have_net && n_iter >= 1 && nchunks >= 2 && nchunks <= pds_handle_s::kPipeEvents;
  PDS_CUDA_OK(cudaEventRecord(h->ev_enter, st));
  PDS_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_enter, 0));
  if (pipe) {
    for (int c = 0; c < nchunks; ++c) {
      const size_t b0 = (size_t)c * h->chunk, nimg = std::min((size_t)h->chunk, (size_t)h->d.B - b0);
      PDS_CUDA_OK(cudaMemcpyAsync(h->xbuf[0] + b0 * h->d.n, x0 + b0 * h->d.n, nimg * h->d.n * sizeof(float), cudaMemcpyHostToDevice,
                                  h->copy_stream));
      PDS_CUDA_OK(cudaEventRecord(h->ev_pipe[c], h->copy_stream));
    }
  } else {
    PDS_CUDA_OK(cudaMemcpyAsync(h->xbuf[0], x0, nb, cudaMemcpyHostToDevice, st));
    PDS_CUDA_OK(cudaEventRecord(h->ev_enter, st));        // after x_0: it gets the whole link first
    PDS_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_enter, 0));
  }
  PDS_CUDA_OK(cudaMemcpyAsync(h->obs, obs, nb, cudaMemcpyHostToDevice, h->copy_stream));
  h->have_true = xtrue != nullptr;
  if (xtrue) PDS_CUDA_OK(cudaMemcpyAsync(h->xtrue, xtrue, nb, cudaMemcpyHostToDevice, h->copy_stream));
  PDS_CUDA_OK(cudaEventRecord(h->ev_inputs, h->copy_stream));
  h->inputs_pending = true;
  PDS_CUDA_OK(cudaMemsetAsync(h->t, 0, nb, st));
  if (h->y1) PDS_CUDA_OK(cudaMemsetAsync(h->y1, 0, 2 * nb, st));
  if (h->sbuf[0]) PDS_CUDA_OK(cudaMemsetAsync(h->sbuf[0], 0, nb, st));
  PDS_CUDA_OK(cudaMemsetAsync(h->sums, 0, (size_t)h->cfg.max_iter * h->d.B * NSUM * sizeof(double), st));
  h->have_problem = true;
  // the chunk-wise download needs a page-locked destination: a copy into pageable memory blocks the host until it is done, which
  // would stall the launches of the following chunks
  cudaPointerAttributes pa{};
  const bool out_pinned = cudaPointerGetAttributes(&pa, x_out) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  (void)cudaGetLastError();
  const bool pipe_out = pipe && out_pinned;
  h->pipe_in = pipe;
  h->pipe_out_host = pipe_out ? x_out : nullptr;
  const int rc_run = pds_run(h, n_iter, stream);
  h->pipe_in = false;
  h->pipe_out_host = nullptr;
  PDS_TRY(rc_run);
  PDS_TRY(wait_inputs(h, st));              // n_iter == 0: still join the side stream before returning
  if (pipe_out) {                           // the chunk copies of x_{final} on the side stream
    PDS_CUDA_OK(cudaEventRecord(h->ev_out_done, h->copy_stream));
    PDS_CUDA_OK(cudaStreamWaitEvent(st, h->ev_out_done, 0));
  } else {
    PDS_CUDA_OK(cudaMemcpyAsync(x_out, h->xbuf[h->cur], nb, cudaMemcpyDeviceToHost, st));
  }
  if (s_out) {
    if (h->sbuf[0]) PDS_CUDA_OK(cudaMemcpyAsync(s_out, h->sbuf[h->scur], nb, cudaMemcpyDeviceToHost, st));
    else std::memset(s_out, 0, nb);
  }
  if (trace_host) {
    const size_t n = (size_t)h->iter * h->d.B * NSUM;
    PDS_REQUIRE(trace_cap >= n, "trace buffer too small");
    PDS_CUDA_OK(cudaMemcpyAsync(trace_host, h->sums, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  PDS_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int pds_profile_enable(pds_handle_t h, int on) {
  PDS_TRY(check_handle(h));
  h->prof = on != 0;
  return 0;
}

int pds_profile_read(pds_handle_t h, double* ms_out, long long* count_out, int reset, pds_stream_t stream) {
  PDS_TRY(check_handle(h));
  PDS_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  for (const auto& r : h->prof_recs) {
    float ms = 0.f;
    PDS_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
    h->prof_ms[r.cat] += ms;
    h->prof_n[r.cat] += 1;
  }
  h->prof_recs.clear();
  h->ev_used = 0;
  for (int c = 0; c < PDS_PROF_NCAT; ++c) {
    if (ms_out) ms_out[c] = h->prof_ms[c];
    if (count_out) count_out[c] = h->prof_n[c];
    if (reset) { h->prof_ms[c] = 0; h->prof_n[c] = 0; }
  }
  return 0;
}

long long pds_kernel_launches(pds_handle_t h) { return h ? h->launches : -1; }
size_t pds_workspace_bytes(pds_handle_t h) { return h ? h->bytes : 0; }

int pds_debug_roll_band_rows(int nimg, int H, int W, int force) { return pds::roll_band_rows(nimg, H, W, pds::tc_num_sms(), force != 0); }

/* test hook: the fp32 CUDA-core convolution engine (on-device cross-check of the tcgen05 engine); before pds_load_dncnn */
int pds_debug_set_conv_engine(pds_handle_t h, int engine) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(engine == PDS_CONV_TCGEN05 || engine == PDS_CONV_SIMT, "unknown conv engine");
  PDS_REQUIRE(!h->have_net, "pds_debug_set_conv_engine must precede pds_load_dncnn");
  h->conv_engine = engine;
  return 0;
}

/* test hook: %globaltimer timeline of the chain kernel's pipeline events (tools/chain_timeline.py) */
int pds_debug_chain_trace(pds_handle_t h, unsigned long long* out_host) {
  PDS_TRY(check_handle(h));
  PDS_REQUIRE(h->tc != nullptr, "no tcgen05 plan (pds_load_dncnn first)");
  return tc_chain_trace(h->tc, out_host);
}

/* test hook: which kernel serves the body layers of a launch of nimg images: 0 fp32 CUDA-core engine, 1 row-streaming
 * (conv_roll_d_kernel / conv_roll_kernel), 2 CTA-pair tiles (conv_tc2_kernel), 3 1-CTA tiles (conv_tc_kernel), 4 chain kernel */
int pds_debug_body_kernel(pds_handle_t h, int nimg) {
  if (!h || !h->have_net) return -1;
  if (h->conv_engine != PDS_CONV_TCGEN05) return 0;
  if (nimg <= 0 || nimg > h->chunk) nimg = h->chunk;
  const BodyDispatch b = body_dispatch(h, nimg);
  if (b.band > 0) return 1;
  if (b.chain) return 4;
  return b.two_cta ? 2 : 3;
}

/* test hook: perf-experiment switches of the tcgen05 engine (see run_dncnn) */
int pds_debug_set_tc_variant(pds_handle_t h, int variant) {
  if (!h) return 1;
  h->tc_variant = variant;
  h->taps.debug_generic = (variant & 8192) ? 1 : 0;
  h->taps.debug_ox = (variant & 16384) ? 8 : 0;          // bit 14: 64 x 32 stencil tiles (8 outputs per thread) for large launches     // bit 13: generic blur stencils instead of blur_1.mat's compile-time tap list
  return 0;
}

}  // extern "C"
