/*
 * pnp_pds.h — C ABI of the B200-native PnP-PDS hot path (libpnp_pds.so).
 *
 * The reference (yodai49/PnP-PDS) is function-level Python with no FFI boundary of its own
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds (ctypes stub in
 * INTEGRATION.md); every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pds_last_error() gives the message
 *     (thread-local).  No C++ types or exceptions cross the boundary.
 *   - images are fp32, planar, (B, C, H, W) contiguous — the reference layout (C,H,W) with a
 *     leading batch of independent restorations; gray images have C = 1.
 *   - `*_dev` pointers are CUDA device pointers on the handle's device, `*_host` are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - a handle owns its workspace: device memory, the side stream and the events are created in pds_create and
 *     pds_load_dncnn and nowhere else (the only later allocations are the CUDA events of the optional profiler,
 *     pds_profile_enable); it is not thread-safe: one handle per stream.
 *   - kernel selection never depends on the environment.
 *   - there is no CPU fallback: if no CUDA device is usable every call fails with an error.
 */
#ifndef PNP_PDS_H_
#define PNP_PDS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDS_ABI_VERSION 1

typedef struct pds_handle_s* pds_handle_t;
typedef void* pds_stream_t;

/* deg_op of operators.get_observation_operators (operators.py:60-79) */
enum { PDS_OP_ID = 0, PDS_OP_BLUR = 1, PDS_OP_RANDOM_SAMPLING = 2 };

/* method branches of iteration.test_iter (iteration.py:48-63, 71-73, 100-105) */
enum {
  PDS_METHOD_A = 0,   /* A-Proposed / ours-A : iteration.py:48-52 */
  PDS_METHOD_B = 1,   /* B-Proposed / ours-B : iteration.py:53-58 */
  PDS_METHOD_C = 2,   /* C-Proposed / ours-C : iteration.py:59-63 */
  PDS_METHOD_FBS = 3, /* A-PnPFBS-DnCNN / comparisonA-1 : iteration.py:71-73 */
  PDS_METHOD_RED = 4, /* A-RED-DnCNN / comparisonA-6 : iteration.py:100-105 */
  PDS_METHOD_ADMM_B2 = 5, /* comparisonB-2 : iteration.py:127-132 + algorithm/admm.py:30-44 */
  PDS_METHOD_ADMM_C = 6,  /* C-PnPADMM-DnCNN / comparisonC-2 : iteration.py:161-165 + admm.py:4-16 */
  PDS_METHOD_RED_C = 7,   /* C-RED-DnCNN / comparisonC-3 : iteration.py:166-172 + admm.py:4-28 */
  /* TV baselines (no denoiser; colour images only, as the reference's D / D_T, operators.py:117-137) */
  PDS_METHOD_TV_A = 8,    /* A-PDS-TV / comparisonA-4 : iteration.py:88-94 */
  PDS_METHOD_TV_B3 = 9,   /* comparisonB-3 : iteration.py:133-140 */
  PDS_METHOD_TV_FBS = 10  /* A-FBS-TV : iteration.py:95-99 */
};

typedef struct {
  int32_t batch;        /* B: independent restorations (images x grid points) */
  int32_t channels;     /* C: 1 or 3 (reference ch) */
  int32_t height;       /* H = shape[-2] */
  int32_t width;        /* W = shape[-1] */
  int32_t method;       /* PDS_METHOD_* */
  int32_t deg_op;       /* PDS_OP_* */
  int32_t max_iter;     /* capacity of the per-iteration traces */
  int32_t reserved;     /* must be 0 */
  int32_t device;       /* CUDA device ordinal */
  int32_t denoiser_chunk; /* images per denoiser pass (0 = library default) */
} pds_config_t;

/* per-item hyper-parameters (one restoration = one item), all precomputed on the host:
 *   epsilon = sqrt(n (1-sp_nl)) r alpha_n gaussian_nl      (operators.py:104)
 *   eta     = alpha_s n sp_nl r / 2                        (operators.py:96)
 * gamma1/gamma2/lambda/alpha as passed to iteration.test_iter (iteration.py:10). */
typedef struct {
  float gamma1;
  float gamma2;
  float epsilon;
  float eta;
  float lambda;   /* myLambda */
  float alpha;    /* poisson_alpha */
} pds_item_params_t;

/* Number of double-precision partial sums recorded per item per iteration. */
#define PDS_TRACE_WIDTH 5
/* trace[it][b][0] = ||t||^2  with y = sigma t the dual variable before the l2-ball scaling (methods A, B)
 * trace[it][b][1] = ||x_{k+1} - x_k||^2      numerator of c[i]      (iteration.py:187)
 * trace[it][b][2] = ||x_k||^2                denominator of c[i]
 * trace[it][b][3] = ||x_{k+1} - x_true||^2   n * mse of eval_psnr   (utils_eval.py:4-7)
 * trace[it][b][4] = sum of the SSIM map over all window positions (eval_ssim, utils_eval.py:9-12) on the
 *                   iterations selected by pds_set_ssim, 0 elsewhere; mean SSIM = value / positions with
 *                   positions = C (H-6)(W-6) for colour, H (W-6) for gray (channel_axis=0 quirk) */

const char* pds_last_error(void);
int pds_abi_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int pds_device_count(void);

/* lifetime — replaces the per-call state set-up of iteration.test_iter (iteration.py:23-41) */
int pds_create(const pds_config_t* cfg, pds_handle_t* out);
int pds_destroy(pds_handle_t h);

/* blur kernel h (l x l, odd l <= 63, row-major float64 as stored in blur_models/*.mat; operators.py:77-78); handles created
 * with deg_op = PDS_OP_BLUR only; may be called again to replace the kernel */
int pds_set_blur_kernel(pds_handle_t h, const double* kernel_host, int l);
/* keep-mask (H*W bytes, 1 = observed) of get_random_sampling_operator (operators.py:40-58);
 * generated on the host with numpy's legacy MT19937 so it is bit-exact */
int pds_set_mask(pds_handle_t h, const uint8_t* mask_host);
/* n == batch, or n == 1 to broadcast */
int pds_set_item_params(pds_handle_t h, const pds_item_params_t* params_host, int n);
/* per-iteration SSIM on the device: 0 = off, 1 = every iteration (as the reference, iteration.py:189),
 * 2 = only the last iteration of each pds_run / pds_restore_host call.  Needs x_true. */
int pds_set_ssim(pds_handle_t h, int mode);
/* inner trip counts and step size of the ADMM cross-check loops: m1, m2, gammaInADMMStep1 (iteration.py:10) */
int pds_set_admm(pds_handle_t h, int m1, int m2, float gamma_step1);
/* PDSW weight blob (models/weights.py) — replaces Denoiser.__init__/load_network (denoiser.py:9-32) */
int pds_load_dncnn(pds_handle_t h, const void* blob_host, size_t nbytes);

/* ---- stand-alone operators (device pointers, (B,C,H,W) fp32) ---- */
/* phi / adj_phi of get_observation_operators (operators.py:60-79) for the handle's deg_op */
int pds_phi(pds_handle_t h, const float* in_dev, float* out_dev, pds_stream_t stream);
int pds_phi_adj(pds_handle_t h, const float* in_dev, float* out_dev, pds_stream_t stream);
/* proj_l2_ball (operators.py:102-108) with the radius precomputed; per item, all channels jointly */
int pds_proj_l2_ball(pds_handle_t h, const float* x_dev, const float* center_dev, float epsilon,
                     float* out_dev, pds_stream_t stream);
/* proj_l1_ball (operators.py:94-100) with the radius precomputed; per item */
int pds_proj_l1_ball(pds_handle_t h, const float* x_dev, float eta, float* out_dev, pds_stream_t stream);
/* prox_GKL (operators.py:114-115) */
int pds_prox_gkl(pds_handle_t h, const float* x_dev, const float* x0_dev, float gamma, float alpha,
                 float* out_dev, pds_stream_t stream);
/* Denoiser.denoise (denoiser.py:14-16,34-46) / KAIR DnCNN.forward (network_dncnn.py:75-77) */
int pds_dncnn_forward(pds_handle_t h, const float* in_dev, float* out_dev, pds_stream_t stream);

/* ---- the resident loop: iteration.test_iter (iteration.py:44-189) ---- */
/* copy x_0, x_obsrv (and x_true, may be NULL) into the handle, zero the dual/sparse state and traces */
int pds_set_problem(pds_handle_t h, const float* x0_dev, const float* obs_dev, const float* xtrue_dev,
                    pds_stream_t stream);
/* run n_iter more iterations (asynchronous on `stream`) */
int pds_run(pds_handle_t h, int n_iter, pds_stream_t stream);
int pds_iterations_done(pds_handle_t h);
/* x_n, s_n (without the reference's +0.5, iteration.py:196), y_n = sigma t; any pointer may be NULL */
int pds_get_state(pds_handle_t h, float* x_dev, float* s_dev, float* y_dev, pds_stream_t stream);
/* synchronises `stream`; writes iterations_done*B*PDS_TRACE_WIDTH doubles */
int pds_get_traces(pds_handle_t h, double* trace_host, size_t capacity_doubles, pds_stream_t stream);

/* ---- host-buffer entry point: one whole restoration job, H2D and D2H inside ---- */
/* equivalent of one iteration.test_iter call per item (iteration.py:10-196):
 * inputs (B,C,H,W) fp32 on the host; outputs x (and s, may be NULL) and the traces.  The transfers overlap the loop (pinned host
 * buffers): x_obsrv / x_true go up on a side stream under the first iteration; for ours-A/B/C with a denoiser x_0 goes up one
 * denoiser chunk at a time with the first iteration's primal step and denoiser pass following chunk by chunk, and in the last
 * iteration every chunk of x leaves for the host as soon as its last layer is done (page-locked x_host only; a pageable
 * destination is filled by one copy at the end).  Returns after `stream` is synchronised. */
int pds_restore_host(pds_handle_t h, const float* x0_host, const float* obs_host, const float* xtrue_host,
                     int n_iter, float* x_out_host, float* s_out_host, double* trace_host,
                     size_t trace_capacity_doubles, pds_stream_t stream);

/* ---- introspection for tests / bench ---- */
/* per-kernel device timing: when enabled, every launch of the categories below is bracketed by CUDA
 * events on the launching stream; pds_profile_read synchronises the stream and returns accumulated
 * milliseconds and launch counts per category (arrays of PDS_PROF_NCAT). */
enum { PDS_PROF_PRIMAL = 0, PDS_PROF_DUAL = 1, PDS_PROF_L1BALL = 2, PDS_PROF_CONV_FIRST = 3, PDS_PROF_CONV_MID = 4,
       PDS_PROF_CONV_LAST = 5, PDS_PROF_NCAT = 6 };
int pds_profile_enable(pds_handle_t h, int on);
int pds_profile_read(pds_handle_t h, double* ms_out, long long* count_out, int reset, pds_stream_t stream);
/* kernels launched by this handle since creation */
long long pds_kernel_launches(pds_handle_t h);
/* bytes of device workspace owned by the handle */
size_t pds_workspace_bytes(pds_handle_t h);

/* ---- KAIR UNet forward ----
 * models/network_unet.py:13-66 (UNet: head, 3 x [nb convs + 2x2 stride-2 conv], nb+1 body convs, 3 x [2x2 stride-2 transposed
 * conv + nb convs], tail; skip additions and the input residual), blocks of models/basicblock.py:61-63, 413-419, 439-445.
 * The reference class cannot be constructed (load_state_dict before any layer exists, network_unet.py:17) and ships no weights;
 * this is the forward operator for that architecture in exact fp32, with its own handle.  Images: fp32 planar (B, C, H, W),
 * H and W divisible by 8. */
typedef struct pds_unet_s* pds_unet_t;
typedef struct {
  int32_t batch;
  int32_t in_nc;      /* UNet(in_nc=...) */
  int32_t out_nc;     /* == in_nc (input residual, network_unet.py:62) */
  int32_t nc[4];      /* channel widths of the four levels (reference default 64, 128, 256, 512) */
  int32_t nb;         /* convs per level (reference default 2) */
  int32_t height;
  int32_t width;
  int32_t device;
} pds_unet_config_t;
int pds_unet_create(const pds_unet_config_t* cfg, pds_unet_t* out);
int pds_unet_destroy(pds_unet_t h);
/* size of the PDSU weight blob for this configuration: 48-byte header ("PDSU", version 1, in_nc, out_nc, nc[4], nb) followed by,
 * in module order (m_head, m_down1..3, m_body, m_up3..1, m_tail), each layer's weight in its torch layout (Conv2d
 * [cout][cin][k][k], ConvTranspose2d [cin][cout][k][k]) and bias, fp32 */
size_t pds_unet_blob_bytes(const pds_unet_config_t* cfg);
int pds_unet_load(pds_unet_t h, const void* blob_host, size_t nbytes);
/* UNet.forward (network_unet.py:52-64); device pointers, not in place */
int pds_unet_forward(pds_unet_t h, const float* in_dev, float* out_dev, pds_stream_t stream);
long long pds_unet_kernel_launches(pds_unet_t h);
size_t pds_unet_workspace_bytes(pds_unet_t h);

/* ---- test hooks (hardware probes used by tests/test_gpu_tcgen05.py; not part of the drop-in surface) ---- */
/* The denoiser always runs on the tcgen05 engine (TMA-fed implicit GEMM: fp16 product + e4m3 operand corrections, fp32 TMEM
 * accumulators).  A fp32 CUDA-core direct convolution is kept as an on-device cross-check for the tests; it is selected per
 * handle with this hook (before pds_load_dncnn), never through the configuration. */
enum { PDS_CONV_TCGEN05 = 0, PDS_CONV_SIMT = 1 };
int pds_debug_set_conv_engine(pds_handle_t h, int engine);
/* kernel-selection switches of the tcgen05 engine: bit 4 / bit 5 force the 1-CTA / 2-CTA tile kernel for the body layers,
 * bit 6 forces the row-streaming body kernels (dncnn_roll.cu) wherever the image is at least 128 pixels wide, bit 7 disables
 * them, bit 8 makes them read the e4m3(a) operand from HBM instead of rebuilding it on chip (every layer then stores it),
 * bit 9 disables the chain kernel (all body layers of a small launch in one persistent launch) in favour of one tile-kernel
 * launch per layer, bits 11 / 12 are timing probes of the chain kernel's MMA issue order (12 gives wrong results by design),
 * bit 13 makes the blur stencils ignore the compile-time tap list of blur_models/blur_1.mat (generic kernels: the cross-check),
 * bit 14 runs large stencil launches on 64 x 32 tiles (8 outputs per thread), bit 15 runs the first layer on the im2col kernel
 * instead of the tap-shifted one (the cross-check), bits 16 - 18 are timing probes of the first layer and of the row-streaming
 * body kernel's epilogue (no activation stores / no epilogue arithmetic or consecutive-address stores / a third of the first
 * layer's MMAs: wrong results by design) */
int pds_debug_set_tc_variant(pds_handle_t h, int variant);
/* SM-cycle-counter stamps of the chain kernel's pipeline events (dncnn_chain.cu: flags polled, TMA issued, TMEM stage free, first
 * plane landed, MMAs issued, accumulator ready, stored, published) for the first 64 units of CTA 0 of the first 4 clusters:
 * out_host == NULL arms the trace for the following launches, otherwise [4][64][8] uint64 are copied back and it is disarmed */
int pds_debug_chain_trace(pds_handle_t h, unsigned long long* out_host);
/* which kernel serves the 64->64 body layers of a launch of nimg images (nimg <= 0: a full denoiser chunk) on this handle:
 * 0 fp32 CUDA-core engine, 1 row-streaming, 2 CTA-pair tile kernel per layer, 3 1-CTA tile kernel per layer, 4 chain kernel
 * (all body layers in one persistent launch); -1 before pds_load_dncnn */
int pds_debug_body_kernel(pds_handle_t h, int nimg);
/* rows per CTA pair when the row-streaming body kernels serve a launch of nimg images of H x W on the current device;
 * 0 = the tile kernels run instead (narrower than 128 pixels, or the cost model prefers tiles; force != 0 skips that comparison) */
int pds_debug_roll_band_rows(int nimg, int H, int W, int force);
/* one tcgen05.mma (M=128, N=16, K=16, B = identity) over a shared-memory region whose 16-byte chunk c
 * holds (c & 1023, c >> 10) repeated; out_host[128][16] therefore reveals which chunk fed every (row, k) */
int pds_debug_umma_probe(unsigned a_off, unsigned sbo, unsigned base_off, unsigned region_bytes, float* out_host);
/* one activation-tile TMA box load at (x, y, plane_index); out_host receives the 23040 shared-memory bytes (18 rows x 10 pixels x 128 B) */
int pds_debug_tma_probe(const void* act_dev, int nimg, int H, int W, int x, int y, int plane_index, void* out_host);

#ifdef __cplusplus
}
#endif
#endif /* PNP_PDS_H_ */
