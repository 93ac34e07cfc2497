// KAIR UNet forward as direct fp32 convolutions on the CUDA cores.
//
// Reference: models/network_unet.py:13-66 (UNet: head, 3 x [nb convs + 2x2 stride-2 conv], body of nb+1 convs,
// 3 x [2x2 stride-2 transposed conv + nb convs], tail, skip additions x+x4 / x+x3 / x+x2 / x+x1 and the input residual),
// built from models/basicblock.py:61-63 (`conv`: Conv2d / ConvTranspose2d + ReLU), :413-419 (`upsample_convtranspose`) and
// :439-445 (`downsample_strideconv`).  The reference class is dead code (it calls load_state_dict before it has any layer,
// network_unet.py:17, and ships no weights), so there is no trained network and no loop to match: this file provides the
// forward operator for the architecture, checked against outputs of the reference's own layers (tests/golden/unet.npz).
//
// Why CUDA cores: the denoiser on the measured path is the 64-channel DnCNN, whose tcgen05 kernels are specialised to
// 64 -> 64 (dncnn_roll.cu, dncnn_chain.cu).  The UNet's widths (64 / 128 / 256 / 512, any nc in general) would need a
// general-K implicit GEMM; until then it runs in exact fp32, which is also what its parity test wants.
//   conv3x3_kernel : 32 x 32 output tile per CTA, 2 x 2 pixels x 8 output channels per thread, input channels staged 8 at a
//                    time in shared memory (halo tile + weights), optional second input added on load (the skip additions),
//                    optional ReLU, optional residual.
//   down2x2_kernel : 2x2 stride-2 convolution (+ ReLU);  up2x2_kernel : 2x2 stride-2 transposed convolution (+ ReLU).
// Layout: planar (B, C, H, W) fp32, like the PDS state.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "kernels.cuh"

namespace pds {
namespace {

constexpr int kOcb = 8, kCib = 8;

struct ConvArgs {
  const float* in;      // (B, Cin, H, W)
  const float* in2;     // optional: added to `in` on load (skip connection), same shape
  const float* w;       // device, re-laid out: [Cin][taps][Cout_pad]  (Cout_pad = multiple of 8)
  const float* bias;    // [Cout_pad]
  const float* res;     // optional residual added to the output (B, Cout, Hout, Wout)
  float* out;           // (B, Cout, Hout, Wout)
  int Cin, Cout, Cout_pad, H, W;   // H, W: input size
  int relu;
};

// 3x3, stride 1, zero padding 1.
__global__ void __launch_bounds__(256) conv3x3_kernel(ConvArgs a) {
  __shared__ float tile[kCib][34][35];
  __shared__ __align__(16) float wsm[kCib][9][kOcb];
  const int tiles_x = (a.W + 31) / 32;
  const int ty0 = (blockIdx.x / tiles_x) * 32, tx0 = (blockIdx.x % tiles_x) * 32;
  const int oc0 = blockIdx.y * kOcb, b = blockIdx.z;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;      // thread -> 2 x 2 pixels at (2 ly, 2 lx)
  const size_t hw = (size_t)a.H * a.W;
  float acc[2][2][kOcb];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int o = 0; o < kOcb; ++o) acc[i][j][o] = 0.f;
  for (int c0 = 0; c0 < a.Cin; c0 += kCib) {
    const int nc = min(kCib, a.Cin - c0);
    for (int idx = threadIdx.x; idx < kCib * 34 * 34; idx += 256) {
      const int c = idx / (34 * 34), r = idx - c * 34 * 34, hy = r / 34, hx = r - hy * 34;
      const int gy = ty0 - 1 + hy, gx = tx0 - 1 + hx;
      float v = 0.f;
      if (c < nc && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
        const size_t g = ((size_t)b * a.Cin + c0 + c) * hw + (size_t)gy * a.W + gx;
        v = __ldg(a.in + g);
        if (a.in2) v += __ldg(a.in2 + g);
      }
      tile[c][hy][hx] = v;
    }
    for (int idx = threadIdx.x; idx < kCib * 9 * kOcb; idx += 256) {
      const int c = idx / (9 * kOcb), r = idx - c * 9 * kOcb, t = r / kOcb, o = r - t * kOcb;
      wsm[c][t][o] = c < nc ? __ldg(a.w + ((size_t)(c0 + c) * 9 + t) * a.Cout_pad + oc0 + o) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < kCib; ++c) {
      float v[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = tile[c][2 * ly + i][2 * lx + j];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3, dx = t - dy * 3;
        const float4 w0 = *reinterpret_cast<const float4*>(&wsm[c][t][0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&wsm[c][t][4]);
        const float wv[kOcb] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int o = 0; o < kOcb; ++o) acc[i][j][o] = fmaf(wv[o], v[i + dy][j + dx], acc[i][j][o]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 0; o < kOcb; ++o) {
    const int oc = oc0 + o;
    if (oc >= a.Cout) break;
    const float bv = __ldg(a.bias + oc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int gy = ty0 + 2 * ly + i, gx = tx0 + 2 * lx + j;
        if (gy >= a.H || gx >= a.W) continue;
        const size_t g = ((size_t)b * a.Cout + oc) * hw + (size_t)gy * a.W + gx;
        float r = acc[i][j][o] + bv;
        if (a.relu) r = fmaxf(r, 0.f);
        if (a.res) r += __ldg(a.res + g);
        a.out[g] = r;
      }
  }
}

// 2x2 convolution, stride 2, no padding: (H, W) -> (H/2, W/2).  One thread = one output pixel x 8 output channels.
__global__ void __launch_bounds__(256) down2x2_kernel(ConvArgs a) {
  __shared__ __align__(16) float wsm[kCib][4][kOcb];
  const int Ho = a.H / 2, Wo = a.W / 2;
  const int tiles_x = (Wo + 15) / 16;
  const int oy = (blockIdx.x / tiles_x) * 16 + (threadIdx.x >> 4), ox = (blockIdx.x % tiles_x) * 16 + (threadIdx.x & 15);
  const int oc0 = blockIdx.y * kOcb, b = blockIdx.z;
  const bool live = oy < Ho && ox < Wo;
  const size_t hw = (size_t)a.H * a.W;
  float acc[kOcb];
#pragma unroll
  for (int o = 0; o < kOcb; ++o) acc[o] = 0.f;
  for (int c0 = 0; c0 < a.Cin; c0 += kCib) {
    const int nc = min(kCib, a.Cin - c0);
    for (int idx = threadIdx.x; idx < kCib * 4 * kOcb; idx += 256) {
      const int c = idx / (4 * kOcb), r = idx - c * 4 * kOcb, t = r / kOcb, o = r - t * kOcb;
      wsm[c][t][o] = c < nc ? __ldg(a.w + ((size_t)(c0 + c) * 4 + t) * a.Cout_pad + oc0 + o) : 0.f;
    }
    __syncthreads();
    if (live) {
      for (int c = 0; c < nc; ++c) {
        const float* p = a.in + ((size_t)b * a.Cin + c0 + c) * hw + (size_t)(2 * oy) * a.W + 2 * ox;
        const float v[4] = {__ldg(p), __ldg(p + 1), __ldg(p + a.W), __ldg(p + a.W + 1)};
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int o = 0; o < kOcb; ++o) acc[o] = fmaf(wsm[c][t][o], v[t], acc[o]);
      }
    }
    __syncthreads();
  }
  if (!live) return;
#pragma unroll
  for (int o = 0; o < kOcb; ++o) {
    const int oc = oc0 + o;
    if (oc >= a.Cout) break;
    float r = acc[o] + __ldg(a.bias + oc);
    if (a.relu) r = fmaxf(r, 0.f);
    a.out[((size_t)b * a.Cout + oc) * ((size_t)Ho * Wo) + (size_t)oy * Wo + ox] = r;
  }
}

// 2x2 transposed convolution, stride 2: (H, W) -> (2H, 2W); out[co, 2y+dy, 2x+dx] = b[co] + sum_ci w[ci][co][dy][dx] (in + in2)[ci, y, x].
// One thread = one input pixel x 8 output channels x its 2 x 2 output block.
__global__ void __launch_bounds__(256) up2x2_kernel(ConvArgs a) {
  __shared__ __align__(16) float wsm[kCib][4][kOcb];
  const int tiles_x = (a.W + 15) / 16;
  const int iy = (blockIdx.x / tiles_x) * 16 + (threadIdx.x >> 4), ix = (blockIdx.x % tiles_x) * 16 + (threadIdx.x & 15);
  const int oc0 = blockIdx.y * kOcb, b = blockIdx.z;
  const bool live = iy < a.H && ix < a.W;
  const size_t hw = (size_t)a.H * a.W;
  float acc[4][kOcb];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int o = 0; o < kOcb; ++o) acc[t][o] = 0.f;
  for (int c0 = 0; c0 < a.Cin; c0 += kCib) {
    const int nc = min(kCib, a.Cin - c0);
    for (int idx = threadIdx.x; idx < kCib * 4 * kOcb; idx += 256) {
      const int c = idx / (4 * kOcb), r = idx - c * 4 * kOcb, t = r / kOcb, o = r - t * kOcb;
      wsm[c][t][o] = c < nc ? __ldg(a.w + ((size_t)(c0 + c) * 4 + t) * a.Cout_pad + oc0 + o) : 0.f;
    }
    __syncthreads();
    if (live) {
      for (int c = 0; c < nc; ++c) {
        const size_t g = ((size_t)b * a.Cin + c0 + c) * hw + (size_t)iy * a.W + ix;
        float v = __ldg(a.in + g);
        if (a.in2) v += __ldg(a.in2 + g);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int o = 0; o < kOcb; ++o) acc[t][o] = fmaf(wsm[c][t][o], v, acc[t][o]);
      }
    }
    __syncthreads();
  }
  if (!live) return;
  const int Wo = 2 * a.W;
  const size_t hwo = 4 * hw;
#pragma unroll
  for (int o = 0; o < kOcb; ++o) {
    const int oc = oc0 + o;
    if (oc >= a.Cout) break;
    const float bv = __ldg(a.bias + oc);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float r = acc[t][o] + bv;
      if (a.relu) r = fmaxf(r, 0.f);
      a.out[((size_t)b * a.Cout + oc) * hwo + (size_t)(2 * iy + (t >> 1)) * Wo + 2 * ix + (t & 1)] = r;
    }
  }
}

enum LayerKind { kConv3 = 0, kDown = 1, kUp = 2 };
struct Layer {
  int kind, cin, cout, relu;
  float* w = nullptr;      // device [cin][taps][cout_pad]
  float* bias = nullptr;   // device [cout_pad]
};

}  // namespace
}  // namespace pds

using namespace pds;

struct pds_unet_s {
  pds_unet_config_t cfg{};
  std::vector<Layer> layers;
  bool loaded = false;
  // activations: x1..x4 (skips) and two ping-pong buffers per resolution level
  float* skip[4] = {nullptr, nullptr, nullptr, nullptr};
  float* tmp[4][2] = {};
  std::vector<void*> allocs;
  size_t bytes = 0;
  long long launches = 0;
};

namespace {

int cout_pad(int c) { return (c + kOcb - 1) / kOcb * kOcb; }

// module order of network_unet.py:19-46 (= the order of the tensors in a PDSU blob)
std::vector<Layer> unet_layers(const pds_unet_config_t& c) {
  std::vector<Layer> L;
  auto conv = [&](int ci, int co, int relu) { Layer l; l.kind = kConv3; l.cin = ci; l.cout = co; l.relu = relu; L.push_back(l); };
  auto down = [&](int ci, int co) { Layer l; l.kind = kDown; l.cin = ci; l.cout = co; l.relu = 1; L.push_back(l); };
  auto up = [&](int ci, int co) { Layer l; l.kind = kUp; l.cin = ci; l.cout = co; l.relu = 1; L.push_back(l); };
  conv(c.in_nc, c.nc[0], 1);                                          // m_head
  for (int i = 0; i < 3; ++i) {                                       // m_down1..3
    for (int k = 0; k < c.nb; ++k) conv(c.nc[i], c.nc[i], 1);
    down(c.nc[i], c.nc[i + 1]);
  }
  for (int k = 0; k < c.nb + 1; ++k) conv(c.nc[3], c.nc[3], 1);       // m_body
  for (int i = 3; i >= 1; --i) {                                      // m_up3..1
    up(c.nc[i], c.nc[i - 1]);
    for (int k = 0; k < c.nb; ++k) conv(c.nc[i - 1], c.nc[i - 1], 1);
  }
  conv(c.nc[0], c.out_nc, 0);                                         // m_tail
  return L;
}

template <typename T>
int ualloc(pds_unet_s* h, T** p, size_t count) {
  void* q = nullptr;
  PDS_CUDA_OK(cudaMalloc(&q, count * sizeof(T)));
  h->allocs.push_back(q);
  h->bytes += count * sizeof(T);
  *p = static_cast<T*>(q);
  return 0;
}

int run_layer(pds_unet_s* h, const Layer& l, const float* in, const float* in2, const float* res, float* out, int H, int W, cudaStream_t st) {
  ConvArgs a{};
  a.in = in; a.in2 = in2; a.res = res; a.out = out;
  a.w = l.w; a.bias = l.bias;
  a.Cin = l.cin; a.Cout = l.cout; a.Cout_pad = cout_pad(l.cout); a.H = H; a.W = W; a.relu = l.relu;
  const int B = h->cfg.batch;
  if (l.kind == kConv3) {
    dim3 grid(((W + 31) / 32) * ((H + 31) / 32), a.Cout_pad / kOcb, B);
    conv3x3_kernel<<<grid, 256, 0, st>>>(a);
  } else if (l.kind == kDown) {
    dim3 grid(((W / 2 + 15) / 16) * ((H / 2 + 15) / 16), a.Cout_pad / kOcb, B);
    down2x2_kernel<<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(((W + 15) / 16) * ((H + 15) / 16), a.Cout_pad / kOcb, B);
    up2x2_kernel<<<grid, 256, 0, st>>>(a);
  }
  PDS_CUDA_OK(cudaGetLastError());
  h->launches++;
  return 0;
}

}  // namespace

extern "C" {

int pds_unet_create(const pds_unet_config_t* cfg, pds_unet_t* out) {
  PDS_REQUIRE(cfg && out, "null argument");
  PDS_REQUIRE(cfg->batch >= 1 && cfg->height >= 8 && cfg->width >= 8, "bad shape");
  PDS_REQUIRE(cfg->height % 8 == 0 && cfg->width % 8 == 0, "UNet needs H and W divisible by 8 (three stride-2 levels, network_unet.py:31-33)");
  PDS_REQUIRE(cfg->in_nc >= 1 && cfg->out_nc >= 1 && cfg->nb >= 1 && cfg->nb <= 16, "bad in_nc / out_nc / nb");
  PDS_REQUIRE(cfg->in_nc == cfg->out_nc, "the input residual (network_unet.py:62) needs out_nc == in_nc");
  for (int i = 0; i < 4; ++i) PDS_REQUIRE(cfg->nc[i] >= 1 && cfg->nc[i] <= 4096, "bad channel width");
  PDS_REQUIRE((long long)cfg->batch <= 65535, "batch exceeds the grid limit");
  int ndev = 0;
  PDS_REQUIRE(cudaGetDeviceCount(&ndev) == cudaSuccess && cfg->device >= 0 && cfg->device < ndev, "no usable CUDA device: this library has no CPU fallback");
  PDS_CUDA_OK(cudaSetDevice(cfg->device));
  pds_unet_s* h = new (std::nothrow) pds_unet_s();
  PDS_REQUIRE(h, "out of host memory");
  h->cfg = *cfg;
  h->layers = unet_layers(*cfg);
  int rc = 0;
  for (int lv = 0; lv < 4 && !rc; ++lv) {
    const size_t n = (size_t)cfg->batch * cfg->nc[lv] * (cfg->height >> lv) * (cfg->width >> lv);
    rc = ualloc(h, &h->skip[lv], n);
    if (!rc) rc = ualloc(h, &h->tmp[lv][0], n);
    if (!rc) rc = ualloc(h, &h->tmp[lv][1], n);
  }
  for (Layer& l : h->layers) {
    if (rc) break;
    const int taps = l.kind == kConv3 ? 9 : 4;
    rc = ualloc(h, &l.w, (size_t)l.cin * taps * cout_pad(l.cout));
    if (!rc) rc = ualloc(h, &l.bias, (size_t)cout_pad(l.cout));
  }
  if (rc) { pds_unet_destroy(h); return rc; }
  *out = h;
  return 0;
}

int pds_unet_destroy(pds_unet_t h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return 0;
}

size_t pds_unet_blob_bytes(const pds_unet_config_t* cfg) {
  if (!cfg) return 0;
  size_t n = 48;
  for (const Layer& l : unet_layers(*cfg)) n += ((size_t)l.cin * l.cout * (l.kind == kConv3 ? 9 : 4) + l.cout) * 4;
  return n;
}

int pds_unet_load(pds_unet_t h, const void* blob, size_t nbytes) {
  PDS_REQUIRE(h && blob, "null argument");
  PDS_CUDA_OK(cudaSetDevice(h->cfg.device));
  const unsigned char* p = static_cast<const unsigned char*>(blob);
  PDS_REQUIRE(nbytes >= 48 && std::memcmp(p, "PDSU", 4) == 0, "not a PDSU weight blob");
  int32_t hdr[8];
  std::memcpy(hdr, p + 4, sizeof(hdr));            // version, in_nc, out_nc, nc[4], nb
  PDS_REQUIRE(hdr[0] == 1, "unsupported PDSU version");
  PDS_REQUIRE(hdr[1] == h->cfg.in_nc && hdr[2] == h->cfg.out_nc && hdr[7] == h->cfg.nb, "PDSU blob does not match the handle (in_nc / out_nc / nb)");
  for (int i = 0; i < 4; ++i) PDS_REQUIRE(hdr[3 + i] == h->cfg.nc[i], "PDSU blob does not match the handle (nc)");
  PDS_REQUIRE(nbytes == pds_unet_blob_bytes(&h->cfg), "PDSU blob size mismatch");
  const float* src = reinterpret_cast<const float*>(p + 48);
  std::vector<float> buf;
  for (Layer& l : h->layers) {
    const int taps = l.kind == kConv3 ? 9 : 4, cp = cout_pad(l.cout);
    const size_t nw = (size_t)l.cin * l.cout * taps;
    for (size_t i = 0; i < nw + l.cout; ++i) PDS_REQUIRE(std::isfinite(src[i]), "PDSU blob holds non-finite parameters");
    buf.assign((size_t)l.cin * taps * cp, 0.f);
    // torch layouts: Conv2d weight [cout][cin][kh][kw]; ConvTranspose2d weight [cin][cout][kh][kw]  ->  [cin][tap][cout_pad]
    for (int co = 0; co < l.cout; ++co)
      for (int ci = 0; ci < l.cin; ++ci)
        for (int t = 0; t < taps; ++t) {
          const size_t s = l.kind == kUp ? ((size_t)ci * l.cout + co) * taps + t : ((size_t)co * l.cin + ci) * taps + t;
          buf[((size_t)ci * taps + t) * cp + co] = src[s];
        }
    PDS_CUDA_OK(cudaMemcpy(l.w, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> bb((size_t)cp, 0.f);
    std::memcpy(bb.data(), src + nw, (size_t)l.cout * 4);
    PDS_CUDA_OK(cudaMemcpy(l.bias, bb.data(), bb.size() * 4, cudaMemcpyHostToDevice));
    src += nw + l.cout;
  }
  h->loaded = true;
  return 0;
}

// network_unet.py:52-64
int pds_unet_forward(pds_unet_t h, const float* in_dev, float* out_dev, pds_stream_t stream) {
  PDS_REQUIRE(h && in_dev && out_dev, "null argument");
  PDS_REQUIRE(h->loaded, "UNet weights not loaded (pds_unet_load)");
  PDS_REQUIRE(in_dev != out_dev, "unet_forward cannot run in place");
  PDS_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int H = h->cfg.height, W = h->cfg.width, nb = h->cfg.nb;
  size_t li = 0;
  auto next = [&]() -> const Layer& { return h->layers[li++]; };
#define UNET_RUN(...)                         \
  do {                                        \
    int _r = run_layer(h, __VA_ARGS__);       \
    if (_r) return _r;                        \
  } while (0)
  // x1 = m_head(x0)
  UNET_RUN(next(), in_dev, nullptr, nullptr, h->skip[0], H, W, st);
  // x_{lv+2} = m_down_{lv+1}(x_{lv+1}): nb convs at level lv, then the stride-2 conv into level lv+1
  for (int lv = 0; lv < 3; ++lv) {
    const float* cur = h->skip[lv];
    for (int k = 0; k < nb; ++k) {
      float* o = h->tmp[lv][k & 1];
      UNET_RUN(next(), cur, nullptr, nullptr, o, H >> lv, W >> lv, st);
      cur = o;
    }
    UNET_RUN(next(), cur, nullptr, nullptr, h->skip[lv + 1], H >> lv, W >> lv, st);
  }
  // x = m_body(x4)
  const float* cur = h->skip[3];
  for (int k = 0; k < nb + 1; ++k) {
    float* o = h->tmp[3][k & 1];
    UNET_RUN(next(), cur, nullptr, nullptr, o, H >> 3, W >> 3, st);
    cur = o;
  }
  // x = m_up_lv(x + x_{lv+1}): transposed conv of the sum into level lv-1, then nb convs there
  for (int lv = 3; lv >= 1; --lv) {
    float* o = h->tmp[lv - 1][0];
    UNET_RUN(next(), cur, h->skip[lv], nullptr, o, H >> lv, W >> lv, st);
    cur = o;
    for (int k = 0; k < nb; ++k) {
      float* o2 = h->tmp[lv - 1][(k + 1) & 1];
      UNET_RUN(next(), cur, nullptr, nullptr, o2, H >> (lv - 1), W >> (lv - 1), st);
      cur = o2;
    }
  }
  // m_tail(x + x1) + x0
  UNET_RUN(next(), cur, h->skip[0], in_dev, out_dev, H, W, st);
#undef UNET_RUN
  return 0;
}

long long pds_unet_kernel_launches(pds_unet_t h) { return h ? h->launches : -1; }
size_t pds_unet_workspace_bytes(pds_unet_t h) { return h ? h->bytes : 0; }

}  // extern "C"
