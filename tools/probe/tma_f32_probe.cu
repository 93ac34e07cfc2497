// Standalone probe: fp32 planar image -> TMA box (bw x 18 x C) into shared memory, rank 3 or 4, SWIZZLE_NONE.
// usage: tma_f32_probe rank bw C W H
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int nfloat, int x0, int y0, int img, int C) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u, bar = base + 16384;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nfloat * 4) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(base),
                   "l"(&tmap), "r"(bar), "r"(x0), "r"(y0), "r"(img * C) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(base),
                   "l"(&tmap), "r"(bar), "r"(x0), "r"(y0), "r"(0), "r"(img) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem + (base - smem_u32(smem)))[i];
}
int main(int argc, char** argv) {
  const int rank = atoi(argv[1]), bw = atoi(argv[2]), C = atoi(argv[3]), W = atoi(argv[4]), H = atoi(argv[5]);
  const int nimg = 2;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<float> h((size_t)nimg * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) + 1.f;
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nfloat = bw * 18 * C; cudaMalloc(&o, nfloat * 4);
  CUtensorMap m; CUresult r;
  if (rank == 3) {
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg * C}; cuuint64_t st[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, 18, (cuuint32_t)C}; cuuint32_t es[3] = {1, 1, 1};
    r = enc(&m, (argc > 8 && atoi(argv[8])) ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (argc > 9 && atoi(argv[9])) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, (argc > 10 && atoi(argv[10])) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)nimg}; cuuint64_t st[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, 18, (cuuint32_t)C, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (argc > 9 && atoi(argv[9])) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, (argc > 10 && atoi(argv[10])) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("rank %d bw %d C %d W %d H %d args:", rank, bw, C, W, H); for (int i = 6; i < argc; ++i) printf(" %s", argv[i]); printf(" encode rc=%d", (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  const int x0 = argc > 6 ? atoi(argv[6]) : 7, y0 = argc > 7 ? atoi(argv[7]) : 15, img = 1;
  if (rank == 3) probe<3><<<1, 128, 16384 + 64 + 1024>>>(m, o, nfloat, x0, y0, img, C); else probe<4><<<1, 128, 16384 + 64 + 1024>>>(m, o, nfloat, x0, y0, img, C);
  cudaError_t e = cudaDeviceSynchronize();
  printf(" kernel: %s", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> got(nfloat); cudaMemcpy(got.data(), o, nfloat * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < C; ++c) for (int y = 0; y < 18; ++y) for (int x = 0; x < bw; ++x) {
      const int gy = y0 + y, gx = x0 + x;
      const float exp = (gy < H && gx < W) ? h[((size_t)(img * C + c) * H + gy) * W + gx] : 0.f;
      if (got[(c * 18 + y) * bw + x] != exp) ++bad;
    }
    printf(" mismatches=%d", bad);
  }
  printf("\n");
  return 0;
}
