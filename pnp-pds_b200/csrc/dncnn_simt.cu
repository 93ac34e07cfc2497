// DnCNN denoiser — CUDA-core layers.
//
// Reference: models/basic_models.py:25-38 (simple_CNN.forward: 3x3 conv, zero padding 1, bias,
// LeakyReLU(0.01), input skip), models/denoiser.py:34-46 (clamp in / clamp out),
// models/network_dncnn.py:42-77 (KAIR DnCNN: ReLU, x - model(x), no clamps).
//
//   conv_first : Cin (1|3) -> 64, fused input clamp, bias, LeakyReLU; writes NHWC hi/lo fp16 planes
//   conv_mid   : 64 -> 64 fp32 direct convolution (cross-check engine for the tcgen05 kernel)
//   conv_last  : 64 -> Cout (1|3), fused bias, residual with the (clamped) network input, output clamp
//
// Inter-layer activations are [img][2][H][W][64] fp16: plane 0 holds hi = fp16(v), plane 1 holds
// lo = fp16(v - hi), so hi + lo carries ~22 mantissa bits and both engines read the same data.
#include <cstring>

#include "kernels.cuh"

namespace pds {
namespace {

__device__ __forceinline__ void unpack8(const uint4& hi, const uint4& lo, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 a = __half22float2(h[k]), b = __half22float2(l[k]);
    v[2 * k] = a.x + b.x;
    v[2 * k + 1] = a.y + b.y;
  }
}

__device__ __forceinline__ void pack8(const float (&v)[8], uint4& hi, uint4& lo) {
  __half2* h = reinterpret_cast<__half2*>(&hi);
  __half2* l = reinterpret_cast<__half2*>(&lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __half h0, l0, h1, l1;
    split_hi_lo(v[2 * k], h0, l0);
    split_hi_lo(v[2 * k + 1], h1, l1);
    h[k] = __halves2half2(h0, h1);
    l[k] = __halves2half2(l0, l1);
  }
}

__device__ __forceinline__ void st_global_256(void* p, const uint4 (&v)[2]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w),
               "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w)
               : "memory");
}

// ------------------------------------------------------------------ first layer
// One thread = 4 horizontally adjacent pixels x 16 output channels (64 fp32 accumulators).  The 4 warps
// of a block work on the SAME 32 pixel quads and each owns one group of 16 output channels, so the
// weights a warp needs are warp-uniform: they travel as a __grid_constant__ kernel parameter and reach
// the FFMAs through the constant bank / uniform registers (no shared-memory weight traffic; the
// earlier LDS-broadcast version was LSU-bound).  Each lane then writes one full 32-byte sector per
// pixel and plane with a 256-bit store.
constexpr int kFirstThreads = 128;

template <int CIN>
struct FirstW {
  float w[9 * CIN][64];   // k = tap*CIN + ci
  float b[64];
};

template <int CIN>
__global__ void __launch_bounds__(kFirstThreads, 4) conv_first_kernel(int nimg, int H, int W, const float* __restrict__ in,
                                                                      const __grid_constant__ FirstW<CIN> wk, float slope, int clamp_in,
                                                                      __half* __restrict__ act) {
  const int ocg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) & 3;     // warp-uniform channel group
  const int lane = threadIdx.x & 31;
  const int qpr = (W + 3) >> 2;                        // pixel quads per row
  const long long nquads = (long long)nimg * H * qpr;
  const size_t hw = (size_t)H * W;
  for (long long q0 = (long long)blockIdx.x * 32; q0 < nquads; q0 += (long long)gridDim.x * 32) {
    const long long q = q0 + lane;
    if (q >= nquads) continue;
    const int img = (int)(q / ((long long)H * qpr));
    const int r = (int)(q - (long long)img * H * qpr);
    const int y = r / qpr, x0 = (r - y * qpr) * 4;
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[p][o] = wk.b[ocg * 16 + o];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float v[3][6];            // 3x6 input window of this channel
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        const float* row = in + ((size_t)(img * CIN + ci) * H + yy) * W;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int xx = x0 + j - 1;
          float t = 0.f;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            t = __ldg(row + xx);
            if (clamp_in) t = fminf(fmaxf(t, 0.f), 1.f);
          }
          v[dy][j] = t;
        }
      }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = &wk.w[(dy * 3 + dx) * CIN + ci][ocg * 16];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float a = v[dy][p + dx];
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[p][o] = fmaf(a, wp[o], acc[p][o]);
          }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (x0 + p >= W) continue;
      const size_t pix = (size_t)y * W + x0 + p;
      uint4 hi[2], lo[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float t8[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) t8[o] = leaky(acc[p][g * 8 + o], slope);
        pack8(t8, hi[g], lo[g]);
      }
      st_global_256(act + (((size_t)img * 2 + 0) * hw + pix) * 64 + ocg * 16, hi);
      st_global_256(act + (((size_t)img * 2 + 1) * hw + pix) * 64 + ocg * 16, lo);
    }
  }
}

template <int CIN>
cudaError_t launch_first_t(int nimg, int H, int W, const float* in, const DncnnLayerW& L, float slope, int clamp_in, __half* act_out,
                           cudaStream_t st) {
  FirstW<CIN> wk;
  std::memcpy(wk.w, L.w_first_host, sizeof(wk.w));
  std::memcpy(wk.b, L.bias_host, sizeof(wk.b));
  const long long nquads = (long long)nimg * H * ((W + 3) / 4);
  long long blocks = (nquads + 31) / 32;
  const int grid = (int)(blocks < 148 * 32 ? (blocks < 1 ? 1 : blocks) : 148 * 32);
  conv_first_kernel<CIN><<<grid, kFirstThreads, 0, st>>>(nimg, H, W, in, wk, slope, clamp_in, act_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ middle layers (fp32 SIMT)
constexpr int kMidThreads = 256, kTile = 16, kHalo = 18, kPitch = 20, kChunk = 16;
constexpr int kInS = kChunk * kHalo * kPitch;        // 5760 floats
constexpr int kWS = kChunk * 9 * 64;                 // 9216 floats
constexpr size_t kMidSmem = (size_t)(kInS + kWS) * sizeof(float);

__global__ void __launch_bounds__(kMidThreads, 2) conv_mid_simt_kernel(int nimg, int H, int W, const __half* __restrict__ act_in,
                                                                       const float* __restrict__ wk /*[64][9][64]*/,
                                                                       const float* __restrict__ bias, float slope,
                                                                       __half* __restrict__ act_out) {
  extern __shared__ __align__(16) float sm[];
  float* in_s = sm;
  float* w_s = sm + kInS;
  const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile;
  const int tile = blockIdx.x;
  const int img = tile / (tiles_x * tiles_y);
  const int trem = tile - img * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * kTile, tx0 = (trem % tiles_x) * kTile;
  const int hw = H * W;
  const __half* in_hi = act_in + ((size_t)img * 2 + 0) * hw * 64;
  const __half* in_lo = act_in + ((size_t)img * 2 + 1) * hw * 64;

  const int og = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, row = lane & 15, x0 = half * 8;

  float acc[8][8];
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[p][o] = 0.f;

  for (int chunk = 0; chunk < 64 / kChunk; ++chunk) {
    __syncthreads();
    for (int task = threadIdx.x; task < 2 * kHalo * kHalo; task += kMidThreads) {
      const int g = task / (kHalo * kHalo), hp = task - g * (kHalo * kHalo);
      const int hy = hp / kHalo, hx = hp - hy * kHalo;
      const int gy = ty0 - 1 + hy, gx = tx0 - 1 + hx;
      float v[8];
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const size_t o = ((size_t)gy * W + gx) * 64 + chunk * kChunk + g * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(in_hi + o));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(in_lo + o));
        unpack8(a, b, v);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) in_s[(g * 8 + k) * (kHalo * kPitch) + hy * kPitch + hx] = v[k];
    }
    {
      const float4* src = reinterpret_cast<const float4*>(wk + (size_t)chunk * kWS);
      float4* dst = reinterpret_cast<float4*>(w_s);
      for (int i = threadIdx.x; i < kWS / 4; i += kMidThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < kChunk; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* rp = in_s + ci * (kHalo * kPitch) + (row + dy) * kPitch + x0;
        float v[10];
        const float4 a0 = *reinterpret_cast<const float4*>(rp);
        const float4 a1 = *reinterpret_cast<const float4*>(rp + 4);
        const float2 a2 = *reinterpret_cast<const float2*>(rp + 8);
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
        v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        v[8] = a2.x; v[9] = a2.y;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wp = reinterpret_cast<const float4*>(w_s + (ci * 9 + dy * 3 + dx) * 64 + og * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[p][o] = fmaf(v[p + dx], w[o], acc[p][o]);
        }
      }
    }
  }
  float bv[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) bv[o] = __ldg(bias + og * 8 + o);
  const int gy = ty0 + row;
  if (gy < H) {
    __half* out_hi = act_out + ((size_t)img * 2 + 0) * hw * 64;
    __half* out_lo = act_out + ((size_t)img * 2 + 1) * hw * 64;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int gx = tx0 + x0 + p;
      if (gx >= W) continue;
      float r[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) r[o] = leaky(acc[p][o] + bv[o], slope);
      uint4 hi, lo;
      pack8(r, hi, lo);
      const size_t o = ((size_t)gy * W + gx) * 64 + og * 8;
      *reinterpret_cast<uint4*>(out_hi + o) = hi;
      *reinterpret_cast<uint4*>(out_lo + o) = lo;
    }
  }
}

// ------------------------------------------------------------------ last layer
constexpr int kLastThreads = 256, kLPix = 68;   // padded pixel stride in floats (bank-conflict-free LDS.128)

template <int COUT>
__global__ void __launch_bounds__(kLastThreads) conv_last_kernel(int nimg, int H, int W, const __half* __restrict__ act_in,
                                                                 const float* __restrict__ wk /*[COUT][9][64]*/,
                                                                 const float* __restrict__ bias, const float* __restrict__ net_in,
                                                                 float residual_sign, int clamp, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                                 // [18*18][68]
  float* ws = sm + kHalo * kHalo * kLPix;           // [COUT][9][64]
  const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile;
  const int t = blockIdx.x;
  const int img = t / (tiles_x * tiles_y);
  const int trem = t - img * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * kTile, tx0 = (trem % tiles_x) * kTile;
  const int hw = H * W;
  const __half* in_hi = act_in + ((size_t)img * 2 + 0) * hw * 64;
  const __half* in_lo = act_in + ((size_t)img * 2 + 1) * hw * 64;
  for (int i = threadIdx.x; i < COUT * 9 * 64; i += kLastThreads) ws[i] = __ldg(wk + i);
  for (int task = threadIdx.x; task < kHalo * kHalo * 8; task += kLastThreads) {
    const int hp = task >> 3, g = task & 7;
    const int hy = hp / kHalo, hx = hp - hy * kHalo;
    const int gy = ty0 - 1 + hy, gx = tx0 - 1 + hx;
    float v[8];
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const size_t o = ((size_t)gy * W + gx) * 64 + g * 8;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(in_hi + o));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(in_lo + o));
      unpack8(a, b, v);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(tile + hp * kLPix + g * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const float4* ap = reinterpret_cast<const float4*>(tile + ((ty + dy) * kHalo + tx + dx) * kLPix);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 a = ap[c4];
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float4 w = *reinterpret_cast<const float4*>(ws + (o * 9 + dy * 3 + dx) * 64 + c4 * 4);
          acc[o] = fmaf(a.x, w.x, acc[o]);
          acc[o] = fmaf(a.y, w.y, acc[o]);
          acc[o] = fmaf(a.z, w.z, acc[o]);
          acc[o] = fmaf(a.w, w.w, acc[o]);
        }
      }
    }
  const int gy = ty0 + ty, gx = tx0 + tx;
  if (gy < H && gx < W) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      const size_t g = ((size_t)(img * COUT + o) * H + gy) * W + gx;
      float xin = __ldg(net_in + g);
      if (clamp) xin = fminf(fmaxf(xin, 0.f), 1.f);
      const float n = acc[o] + __ldg(bias + o);
      float r = residual_sign > 0.f ? n + xin : xin - n;
      if (clamp) r = fminf(fmaxf(r, 0.f), 1.f);
      out[g] = r;
    }
  }
}

template <int COUT>
cudaError_t launch_last_t(int nimg, int H, int W, const __half* act_in, const DncnnLayerW& L, const float* net_in, float rs, int clamp,
                          float* out, cudaStream_t st) {
  const size_t smem = (size_t)(kHalo * kHalo * kLPix + COUT * 9 * 64) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(conv_last_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { once.retract(); return e; }
  }
  const int tiles = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * nimg;
  conv_last_kernel<COUT><<<tiles, kLastThreads, smem, st>>>(nimg, H, W, act_in, L.w_last, L.bias, net_in, rs, clamp, out);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv_first(int nimg, int C, int H, int W, const float* in, const DncnnLayerW& L, float slope, int clamp_in,
                              __half* act_out, cudaStream_t st) {
  if (C == 1) return launch_first_t<1>(nimg, H, W, in, L, slope, clamp_in, act_out, st);
  if (C == 3) return launch_first_t<3>(nimg, H, W, in, L, slope, clamp_in, act_out, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_conv_mid_simt(int nimg, int H, int W, const __half* act_in, const DncnnLayerW& L, float slope, __half* act_out,
                                 cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(conv_mid_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMidSmem);
    if (e != cudaSuccess) { once.retract(); return e; }
  }
  const int tiles = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * nimg;
  conv_mid_simt_kernel<<<tiles, kMidThreads, kMidSmem, st>>>(nimg, H, W, act_in, L.w_mid, L.bias, slope, act_out);
  return cudaGetLastError();
}

cudaError_t launch_conv_last(int nimg, int C, int H, int W, const __half* act_in, const DncnnLayerW& L, const float* net_in,
                             float residual_sign, int clamp, float* out, cudaStream_t st) {
  if (C == 1) return launch_last_t<1>(nimg, H, W, act_in, L, net_in, residual_sign, clamp, out, st);
  if (C == 3) return launch_last_t<3>(nimg, H, W, act_in, L, net_in, residual_sign, clamp, out, st);
  return cudaErrorInvalidValue;
}

}  // namespace pds
