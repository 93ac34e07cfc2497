// Stand-alone probe: issue rate of fp32 FMA forms on sm_100a — FFMA with a constant-bank operand (what the stencils use), FFMA with
// three register operands, FFMA2 (fma.rn.f32x2) — 16 independent accumulators per thread, 8 warps x 4 blocks per SM.
#include <cuda_runtime.h>
#include <cstdio>
struct W { float w[32]; };
template <int MODE>
__global__ void __launch_bounds__(256, 4) k(const __grid_constant__ W cw, const float* wg, float* out, int iters) {
  float acc[16], v[16], wr[8];
  for (int i = 0; i < 16; ++i) { acc[i] = 0.f; v[i] = threadIdx.x * 0.001f + i; }
  for (int i = 0; i < 8; ++i) wr[i] = wg[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(cw.w[t], v[(c + t) & 15], acc[c]);
      } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(wr[t], v[(c + t) & 15], acc[c]);
      } else if (MODE == 3) {          // operand pattern that cannot be paired: stays scalar FFMA
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(cw.w[t + (c & 1)], v[(c * 7 + t) & 15], acc[c]);
      } else {
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          unsigned long long a, b, w2, r;
          asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc[c]), "f"(acc[c + 1]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(v[(c + 2 * (t & 3)) & 15]), "f"(v[(c + 2 * (t & 3) + 1) & 15]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(w2) : "f"(wr[t]), "f"(wr[t]));
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(w2), "l"(b), "l"(a));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[c]), "=f"(acc[c + 1]) : "l"(r));
        }
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main() {
  W cw; for (int i = 0; i < 32; ++i) cw.w[i] = 1.0f / (i + 3);
  float *wg, *out; cudaMalloc(&wg, 64); cudaMemcpy(wg, cw.w, 32, cudaMemcpyHostToDevice); cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000, grid = 148 * 4;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const char* names[4] = {"FFMA const-bank operand", "FFMA three registers", "FFMA2 (f32x2)", "scalar FFMA (unpairable)"};
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (m == 0) k<0><<<grid, 256>>>(cw, wg, out, iters); else if (m == 1) k<1><<<grid, 256>>>(cw, wg, out, iters); else if (m == 2) k<2><<<grid, 256>>>(cw, wg, out, iters); else k<3><<<grid, 256>>>(cw, wg, out, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fma = (double)grid * 256 * iters * 8 * 16;
    printf("%-26s %.3f ms  %.2f TFMA/s  = %.1f FMA/clk/SM at 1.965 GHz\n", names[m], ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
