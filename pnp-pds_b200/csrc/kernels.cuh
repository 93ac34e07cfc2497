// Internal launch interface between the C-ABI (pds_api.cu) and the kernel translation units.
#pragma once
#include "common.cuh"

namespace pds {

struct Dims {
  int B, C, H, W;
  int hw;  // H*W
  int n;   // C*H*W  (elements per item)
};

// Arguments shared by the fused primal / dual kernels.  Pointers are (B,C,H,W) fp32 unless noted.
struct StepArgs {
  Dims d;
  int method;                // PDS_METHOD_A/B/C
  const float* x;            // x_k
  const float* xn;           // x_{k+1} (denoiser output)              [dual]
  float* u;                  // denoiser input x_k - gamma1 Phi^T y_k   [primal]
  float* t;                  // dual state, y = sigma t (A,B) or y (C); read in both, written in dual
  const float* s_old;        // s_k          (B)
  const float* s_new;        // s_{k+1}      (B)   [dual]
  const float* obs;          // x_obsrv
  const float* xtrue;        // may be null
  const uint8_t* mask;       // (H*W) keep mask or null
  const ItemParams* prm;     // [B]
  const double* sums_prev;   // [B][NSUM] of the previous iteration (null on the first)
  double* sums_cur;          // [B][NSUM] of this iteration          [dual]
};

// ---- pds_elementwise.cu: Phi in {Id, mask} ---------------------------------
cudaError_t launch_primal_pointwise(const StepArgs& a, cudaStream_t st);
cudaError_t launch_dual_pointwise(const StepArgs& a, cudaStream_t st);
cudaError_t launch_mask_apply(const Dims& d, const float* in, const uint8_t* mask, float* out, cudaStream_t st);
cudaError_t launch_scale_by_sigma(const Dims& d, const float* t, const ItemParams* prm, const double* sums, int method,
                                  float* y, cudaStream_t st);
// stand-alone prox operators
cudaError_t launch_diff_norm2(const Dims& d, const float* x, const float* c, double* acc /*[B]*/, cudaStream_t st);
cudaError_t launch_proj_l2_apply(const Dims& d, const float* x, const float* c, float eps, const double* acc, float* out,
                                 cudaStream_t st);
cudaError_t launch_proj_l2_items(const Dims& d, const float* x, const float* c, const ItemParams* prm, const double* acc, float* out,
                                 cudaStream_t st);
cudaError_t launch_prox_gkl(const Dims& d, const float* x, const float* x0, float gamma, float alpha, float* out,
                            cudaStream_t st);
// metrics for methods without a dual kernel: fills SUM_DX2, SUM_X2, SUM_ERR2
cudaError_t launch_metrics(const Dims& d, const float* xn, const float* x, const float* xtrue, double* sums_cur,
                           cudaStream_t st);
// out[i] = sum_k coef[b][k] * in[k][i] per item b (ADMM cross-check loops)
struct LinArgs {
  Dims d;
  const float* in[6];
  float* out;
  const float* coef;   // device [B][6]
  int nterms;
};
cudaError_t launch_lincomb(const LinArgs& a, cudaStream_t st);
cudaError_t launch_ratio(const Dims& d, const float* num, const float* den, const ItemParams* prm, float* out, cudaStream_t st);
cudaError_t launch_fill(size_t n, float v, float* out, cudaStream_t st);
// out = a*p + b*q + c*r   (r, q may be null)
cudaError_t launch_axpbypcz(size_t n, float a, const float* p, float b, const float* q, float c, const float* r, float* out,
                            cudaStream_t st);

// ---- pds_blur.cu: Phi = periodic stencil -----------------------------------
struct BlurTaps {            // device arrays, built by pds_set_blur_kernel
  const float* w[2];         // [0] = Phi taps, [1] = Phi^T taps
  const short2* off[2];      // (dy, dx) per tap: out[i,j] += w * in[i+dy, j+dx] (periodic)
  int ntaps;
  int ry, rx;                // max |dy|, max |dx|
  int debug_generic;         // test hook: 1 = never use the compile-time tap list of blur_1.mat (the generic kernels are the cross-check)
  int debug_ox;              // test hook: outputs per thread of the register-tiled stencil for large launches (0 = default)
  const float* w_host;       // host copies (owned by the handle): the register-tiled stencil takes the weight box by value
  const short2* off_host[2];
};
cudaError_t launch_blur_apply(const Dims& d, const BlurTaps& taps, int adjoint, const float* in, float* out, cudaStream_t st);
cudaError_t launch_primal_blur(const StepArgs& a, const BlurTaps& taps, cudaStream_t st);
cudaError_t launch_dual_blur(const StepArgs& a, const BlurTaps& taps, cudaStream_t st);

// ---- pds_tv.cu: TV baselines (C = 3) --------------------------------------------
// xn = u - gamma1 * D_T(y1);   y1 <- unit-ball projection per pixel of y1 + gamma2 * D(2 xn - x)     (y1: (B,6,H,W))
cudaError_t launch_tv_primal(const Dims& d, const float* u, const float* y1, const ItemParams* prm, float* xn, cudaStream_t st);
cudaError_t launch_tv_dual(const Dims& d, const float* xn, const float* x, const ItemParams* prm, float* y1, cudaStream_t st);

// ---- pds_ssim.cu ---------------------------------------------------------------
cudaError_t launch_ssim(const Dims& d, const float* xtrue, const float* x, unsigned* mm, double* sums_cur, cudaStream_t st);

// ---- pds_l1ball.cu ----------------------------------------------------------
// s_out = P_{l1-ball(eta)}(z), z = s_in - gamma1*sigma*t (t != null) or z = s_in.
// eta per item from prm (eta_override < 0) or the scalar eta_override.
cudaError_t launch_l1ball(const Dims& d, const float* s_in, const float* t, const ItemParams* prm, const double* sums_prev,
                          float eta_override, float* s_out, float* tau_out /*[B] or null*/, cudaStream_t st);

// ---- dncnn_*.cu ---------------------------------------------------------------
// Activations between layers: [img][2 planes][H][W][128 B] (NHWC).  Plane 0 = fp16(v) x 64 channels in both engines.
// Plane 1: SIMT engine = fp16(v - fp16(v)) x 64;  tcgen05 engine = e4m3(v) x 64 | e4m3((v - fp16(v)) 2^10) x 64.
// The two engines never share a buffer.
struct DncnnLayerW {
  const float* w_first_host;  // HOST [9*Cin][64] (k = tap*Cin + ci): handed to the first-layer kernel by value (constant bank)
  const float* bias_host;     // HOST [64] bias of the first layer
  const __half* w_first_tc;   // tcgen05 first layer: [w_hi 64 rows ; w_lo 64 rows] x 128 B, k = tap*Cin+ci in the first 27 halves, swizzled
  const __half* w_first_tc2;  // tap-shifted first layer (dncnn_tc.cu first2): [tap 9][K chunk 2][oc 64][8 halves], K-slots
                              // [w_hi(c) | w_hi(c) | w_lo(c) | bias in three fp16 terms (centre tap) | 0]
  const float* w_mid;     // [64 ci][9][64 oc] fp32                      SIMT engine
  const __half* w_mid_tc; // smem image for the tcgen05 engine: [9 taps][fp16 tile | e4m3 tile][64 oc][128 B], 128B-swizzled rows
  const __half* w_mid_tc2; // 2-CTA engine: [cta 2][tap 9][fp16 tile 32 rows | e4m3 tile 32 rows][128 B], swizzled
  const float* w_last;    // [Cout][9][64 ci]                            last layer (SIMT engine)
  const __half* w_last_tc; // last layer (tcgen05 engine): fp16 tile [w_hi 32 rows ; w_lo 2^S 32 rows] x 128 B (SWIZZLE_128B), then e4m3 tile
                           // [w_hi 2^(S-10)] 32 rows x 64 B (SWIZZLE_64B); row n = tap*Cout + c, rows >= 9*Cout zero
  float lo_scale;         // 2^-S of the e4m3 correction accumulator (tcgen05 engine, pds_api.cu tc_split_scales)
  const float* bias;      // [Cout of this layer]
};
cudaError_t launch_conv_first(int nimg, int C, int H, int W, const float* in /*(nimg,C,H,W)*/, const DncnnLayerW& L, float slope,
                              int clamp_in, __half* act_out, cudaStream_t st);
cudaError_t launch_conv_mid_simt(int nimg, int H, int W, const __half* act_in, const DncnnLayerW& L, float slope, __half* act_out,
                                 cudaStream_t st);
cudaError_t launch_conv_last(int nimg, int C, int H, int W, const __half* act_in, const DncnnLayerW& L, const float* net_in,
                             float residual_sign, int clamp, float* out, cudaStream_t st);

// tcgen05 engine (dncnn_tc.cu)
struct TcPlan;  // opaque: tensor maps for the two activation buffers
int tc_plan_create(int nimg, int H, int W, __half* act0, __half* act1, TcPlan** out);  // returns 0 / error (message set)
void tc_plan_destroy(TcPlan* p);
void tc_plan_set_probe_bits(TcPlan* p, int bits);
// in_buf: 0 or 1 (which activation buffer is the input; the other is the output)
cudaError_t launch_conv_mid_tc(TcPlan* plan, int in_buf, int nimg, const DncnnLayerW& L, float slope, cudaStream_t st);
// write_a8: also store the e4m3(fp16(v)) half of plane 1 (0 when the next layer is the row-streaming kernel, which rebuilds it)
// im2col: 1 = the im2col kernel (cross-check), 0 = the tap-shifted kernel (default)
cudaError_t launch_conv_first_tc(TcPlan* plan, int nimg, int C, const float* in, const DncnnLayerW& L, float slope, int clamp_in,
                                 int write_a8, int im2col, cudaStream_t st);
cudaError_t launch_conv_mid_tc2(TcPlan* plan, int in_buf, int nimg, const DncnnLayerW& L, float slope, cudaStream_t st);
cudaError_t launch_conv_last_tc(TcPlan* plan, int in_buf, int nimg, int C, const DncnnLayerW& L, const float* net_in,
                                float residual_sign, int clamp, float* out, cudaStream_t st);
// all body layers in one persistent launch (dncnn_chain.cu); set up by tc_plan_chain after the layers are uploaded
int tc_plan_chain(TcPlan* plan, const DncnnLayerW* layers, int depth, size_t* bytes_out);
bool tc_chain_available(const TcPlan* plan, int nimg);
cudaError_t launch_conv_body_chain(TcPlan* plan, int in_buf, int nimg, float slope, int interleave, cudaStream_t st);
int tc_chain_trace(TcPlan* plan, unsigned long long* out_host);   // debug timeline: null = arm, else read back [4][64][8] and disarm
int tc_num_sms();
// row-streaming body layer (dncnn_roll.cu): band height for a launch of nimg images (0 = not applicable, use the tile kernels)
int roll_setup();
int roll_band_rows(int nimg, int H, int W, int num_sms, bool force);
// derive: 1 = rebuild e4m3(fp16(v)) from plane 0 on chip (input layers may skip that store), 0 = read it from plane 1
cudaError_t launch_conv_mid_roll(TcPlan* plan, int in_buf, int nimg, int band_rows, const DncnnLayerW& L, float slope, int derive,
                                 int write_a8, cudaStream_t st);

}  // namespace pds
