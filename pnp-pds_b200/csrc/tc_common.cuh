// Device-side building blocks shared by the tcgen05 kernels (dncnn_tc.cu, dncnn_roll.cu): mbarrier / TMA / tcgen05 PTX
// wrappers, the K-major SWIZZLE_128B descriptor helpers, the activation epilogue (fp16 plane + e4m3 plane) and the
// launch plan.  Everything is internal to the library (anonymous namespace: one copy per translation unit).
#pragma once
#include <cuda.h>
#include <cuda_fp8.h>

#include <vector>

#include "kernels.cuh"

namespace pds {

// Per-layer entry of the chain kernel's device table (dncnn_chain.cu).
struct ChainLayer {
  const __half* w;          // DncnnLayerW::w_mid_tc2 (per-CTA halves of the weight image)
  const float* bias;
  float lo_scale;
  float pad;
};
constexpr int kChainMaxLayers = 40;     // biases of the whole chain are staged in shared memory (reference networks: 15 / 18 body layers)
constexpr int kChainMaxPairs = 16384;   // launches of up to 2 Mpx can run as one chain launch (flags: layers x pairs x 4 B)

// Tensor maps and geometry of the two activation buffers of one engine handle.
struct TcPlan {
  CUtensorMap map[2];       // halo-tile boxes (64 ch x 10 px x 18 rows) for the 16x8-pixel tile kernels
  CUtensorMap map_row[2];   // row boxes (64 ch x 130 px x 1 row) for the row-streaming body kernel
  CUtensorMap map_lo[2];    // halo-tile boxes over the a_lo half of plane 1 (32 halves x 10 px x 18 rows, SWIZZLE_64B): last layer
  __half* act[2];
  int nimg, H, W;
  int num_sms;
  // chain kernel (all body layers in one launch, dncnn_chain.cu); null when the plan's launches are too large for it
  ChainLayer* chain_layers = nullptr;
  uint32_t* chain_flags = nullptr;      // [chain_nlayers][chain_stride] per-unit arrival counters, monotone across launches
  int chain_nlayers = 0, chain_stride = 0, chain_npairs_last = -1;
  uint32_t chain_epoch = 0;
  size_t chain_bytes = 0;
  unsigned long long* chain_trace = nullptr;   // debug timeline buffer (pds_debug_chain_trace), normally null
  // first layer (tap-shifted kernel): tensor maps over the fp32 network input, one per (pointer, planes) seen
  struct InMap { const float* ptr; int planes, C; CUtensorMap map; };
  std::vector<InMap> in_maps;
  int probe_bits = 0;                   // timing probes of the first layer and of the row-streaming epilogue (tc_variant bits 16 - 18, wrong results by design)
};
int chain_setup();
int tc_plan_set_chain(TcPlan* plan, const std::vector<ChainLayer>& layers);
bool chain_available(const TcPlan* plan, int nimg);
cudaError_t launch_conv_chain(TcPlan* plan, int in_buf, int nimg, float slope, int interleave, cudaStream_t st);

namespace {

constexpr int kTileRows = 16, kTileCols = 8;          // output tile (M = 128)
constexpr int kHaloRows = 18, kHaloPitch = 10;        // pixels
constexpr uint32_t kPlaneBytes = kHaloRows * kHaloPitch * 128;   // 23040 bytes landed by one TMA box
constexpr uint32_t kPlaneSlot = 23 * 1024;                        // slot stride: keeps every slot 1024-B aligned
constexpr int kThreads = 192;
constexpr uint32_t kIdescBase = (1u << 4) /*D=f32*/ | (0u << 7) /*A=f16*/ | (0u << 10) /*B=f16*/ | ((128u >> 4) << 24) /*M=128*/;

struct TcArgs {
  const __half* w_img;
  const float* bias;
  __half* out;            // body layers: next activation buffer
  float slope;
  float lo_scale;         // 2^-S of this layer's e4m3 correction accumulator
  int H, W, nimg, tiles_x, tiles_y, ntiles;
  // last layer only
  int write_a8;           // body / first layer: also store e4m3(fp16(v)) (0 when the consumer rebuilds it: row-streaming kernel)
  const float* net_in;    // (nimg, C, H, W) network input (residual)
  float* out_f32;         // (nimg, C, H, W)
  int C;
  float res_sign;
  int clamp;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// "TMEM stage drained" signal: the only accesses it has to follow are the warp's tcgen05.ld (ordered by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync), so no release fence is needed — the default .release arrive compiles to a MEMBAR that
// waits for every global load / store the thread still has in flight.
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.  The bound is wall time (%globaltimer,
// kWaitTrapNs), not a poll count: under compute-sanitizer, cuda-gdb or time-slicing a healthy kernel can need any number of polls.
constexpr unsigned long long kWaitTrapNs = 8000000000ull;   // 8 s
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && (spin & 1023u) == 1023u) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitTrapNs) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor").
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7u) << 49) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// Packed fp32 pairs (FMUL2 / FADD2, sm_100): one issue slot for two lanes of the epilogue's elementwise maths.  Same roundings as
// the scalar forms.
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}

constexpr float kActLoScale = 1024.f;     // 2^10: e4m3(a_lo * 2^10) stays finite for |a| < 448 (a_lo <= 2^-11 * 2^ceil(log2|a|))

__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}
// two fp16 values -> two e4m3 bytes (low half -> low byte), round to nearest even, saturating
__device__ __forceinline__ uint32_t e4m3x2_from_f16x2(uint32_t h2) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
  return r;
}

// 32 channels [c0, c0+32) of one pixel: v = d0 + lo_scale*d1 + bias -> LeakyReLU, written as full 32-byte sectors:
//   plane 0 (dst_p0, fp16 x 64):  fp16(v) at channels c0..c0+31                         (two 256-bit stores)
//   plane 1 (dst_p1, 128 bytes):  e4m3(fp16(v)) at byte c0.. (only if write_a8), e4m3((v - fp16(v)) * 2^10) at byte 64+c0..
// The first half of plane 1 is a function of plane 0 alone, so a consumer can rebuild it on chip: the row-streaming body
// kernel does (dncnn_roll.cu), and the layers feeding it skip the store (write_a8 = 0: 192 instead of 256 bytes per pixel).
__device__ __forceinline__ void store_half_row(__half* dst_p0, uint8_t* dst_p1, const uint32_t (&d0)[32], const uint32_t (&d1)[32],
                                               const float* bias_s, int c0, float slope, float lo_scale, int write_a8, uint32_t cs = 32) {
  // cs: byte distance between the 32-byte pieces of a row (32; other values only in the store-pattern timing probe of dncnn_roll.cu)
  uint32_t a8[8], l8[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t hi[8];
    float l[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = q * 16 + 2 * k;
      // packed fp32 pairs (FFMA2 / FADD2 / FMUL2): the same roundings as the scalar forms in two thirds of the issue slots
      const float2 b = *reinterpret_cast<const float2*>(bias_s + c0 + c);
      float2 v = add2(fma2(make_float2(__uint_as_float(d1[c]), __uint_as_float(d1[c + 1])), make_float2(lo_scale, lo_scale),
                           make_float2(__uint_as_float(d0[c]), __uint_as_float(d0[c + 1]))), b);
      const float2 vs = mul2(v, make_float2(slope, slope));
      v.x = fmaxf(v.x, vs.x);                // LeakyReLU for 0 <= slope <= 1 (0.01 simple_CNN, 0 KAIR ReLU)
      v.y = fmaxf(v.y, vs.y);
      const __half2 hh = __floats2half2_rn(v.x, v.y);
      const float2 lo = mul2(sub2(v, __half22float2(hh)), make_float2(kActLoScale, kActLoScale));
      hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
      l[2 * k] = lo.x;
      l[2 * k + 1] = lo.y;
    }
    st_global_256(reinterpret_cast<uint8_t*>(dst_p0) + (size_t)(c0 / 16 + q) * cs, hi);
    if (write_a8) {                          // warp-uniform
#pragma unroll
      for (int k = 0; k < 4; ++k) a8[q * 4 + k] = e4m3x2_from_f16x2(hi[2 * k]) | (e4m3x2_from_f16x2(hi[2 * k + 1]) << 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) l8[q * 4 + k] = pack_e4m3x4(l[4 * k], l[4 * k + 1], l[4 * k + 2], l[4 * k + 3]);
  }
  if (write_a8) st_global_256(dst_p1 + (size_t)(c0 / 32) * cs, a8);
  st_global_256(dst_p1 + (size_t)(2 + c0 / 32) * cs, l8);
}

// Programmatic dependent launch: a layer's CTAs may start (barrier init, TMEM alloc, weight loads) while the
// previous layer's grid is still draining; everything that touches the previous layer's output waits here.
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred;
}

__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// ---- cta_group::2 (CTA pair) variants --------------------------------------------------------
namespace two {
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// "TMEM stage drained" signal to the MMA issuer in CTA 0.  Relaxed on purpose: the only accesses it has to follow are this
// warp's tcgen05.ld (ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync).  A .release arrive at cluster scope
// compiles to MEMBAR.ALL.GPU, i.e. it waits for every global store the thread has in flight — the activations of the
// previous tile — which made the epilogue, not the tensor pipe, the pacing stage (ncu: stall_membar 2.0 of 8.3 cycles/inst).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// kind / collector usage as template parameters (COLL: 0 none, 1 collector::a::fill, 2 collector::a::use, 3 collector::a::lastuse).
// The A collector keeps the A tile of the previous MMA inside the tensor core: consecutive MMAs with the same A descriptor
// then read it from shared memory once (measured with ncu: l1tex tc wavefronts drop by exactly the A share).
template <bool F16, int COLL>
__device__ __forceinline__ void umma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
#define PDS_MMA2(KIND, C)                                                                                              \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                     \
               "tcgen05.mma.cta_group::2.kind::" KIND C " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), \
               "r"(idesc), "r"(accumulate)                                                                             \
               : "memory")
  if constexpr (F16) {
    if constexpr (COLL == 0) PDS_MMA2("f16", "");
    else if constexpr (COLL == 1) PDS_MMA2("f16", ".collector::a::fill");
    else if constexpr (COLL == 2) PDS_MMA2("f16", ".collector::a::use");
    else PDS_MMA2("f16", ".collector::a::lastuse");
  } else {
    if constexpr (COLL == 0) PDS_MMA2("f8f6f4", "");
    else if constexpr (COLL == 1) PDS_MMA2("f8f6f4", ".collector::a::fill");
    else if constexpr (COLL == 2) PDS_MMA2("f8f6f4", ".collector::a::use");
    else PDS_MMA2("f8f6f4", ".collector::a::lastuse");
  }
#undef PDS_MMA2
}

// ---- geometry and MMA issue of the CTA-pair tile kernels (conv_tc2_kernel in dncnn_tc.cu, conv_chain_kernel in dncnn_chain.cu)
constexpr uint32_t kWHalf = 9 * 64 * 128;             // 73728: per-CTA weight image
constexpr int kSlots2 = 5;
constexpr uint32_t kOffA2 = kWHalf, kOffBar2 = kOffA2 + kSlots2 * kPlaneSlot;
constexpr uint32_t kOffBias2 = kOffBar2 + 192, kSmemBytes2 = kOffBias2 + 256 + 1024;
constexpr uint32_t kAccCols2 = 128, kTmemCols2 = 512;
constexpr int kAccStages2 = 4;                        // 4 x 128 columns: the epilogue may lag the tensor pipe by three tiles
constexpr uint32_t kIdescBase2 = (1u << 4) | ((256u >> 4) << 24);     // D=f32, A=B=f16 (or e4m3: same code 0), M=256
constexpr uint32_t kIdescN64 = kIdescBase2 | ((64u >> 3) << 17);


template <bool P0>
__device__ __forceinline__ void issue_plane2(uint32_t d_tmem, uint32_t a_lo, uint32_t w_lo) {
  constexpr uint32_t kHiA = ((kHaloPitch * 128u) >> 4) | (1u << 14) | (2u << 29);
  constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ao = (uint32_t)((dy * kHaloPitch + dx) * 128 + k * 32) >> 4;
      const uint32_t bo = (uint32_t)(tap * 8192 + (P0 ? 0 : 4096) + k * 32) >> 4;
      const uint32_t acc = (tap == 0 && k == 0) ? 0u : 1u;
      if (P0) umma_f16_2sm(d_tmem, desc64(a_lo + ao, kHiA), desc64(w_lo + bo, kHiB), kIdescN64, acc);
      else umma_f8_2sm(d_tmem + 64u, desc64(a_lo + ao, kHiA), desc64(w_lo + bo, kHiB), kIdescN64, acc);
    }
  }
}
// Both planes of a unit, the two MMA kinds alternating: consecutive MMAs then accumulate into DIFFERENT TMEM accumulators
// (f16 -> columns [0,64), f8f6f4 -> [64,128)), so an MMA does not wait for its predecessor's accumulator update.  The order of
// the MMAs within each accumulator is that of issue_plane2 -> bit-identical results.
__device__ __forceinline__ void issue_unit2_interleaved(uint32_t d_tmem, uint32_t a0_lo, uint32_t a1_lo, uint32_t w_lo) {
  constexpr uint32_t kHiA = ((kHaloPitch * 128u) >> 4) | (1u << 14) | (2u << 29);
  constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ao = (uint32_t)((dy * kHaloPitch + dx) * 128 + k * 32) >> 4;
      const uint32_t b0 = (uint32_t)(tap * 8192 + k * 32) >> 4;
      const uint32_t b1 = (uint32_t)(tap * 8192 + 4096 + k * 32) >> 4;
      const uint32_t acc = (tap == 0 && k == 0) ? 0u : 1u;
      umma_f16_2sm(d_tmem, desc64(a0_lo + ao, kHiA), desc64(w_lo + b0, kHiB), kIdescN64, acc);
      umma_f8_2sm(d_tmem + 64u, desc64(a1_lo + ao, kHiA), desc64(w_lo + b1, kHiB), kIdescN64, acc);
    }
  }
}
// TIMING PROBE ONLY (wrong results): every MMA of the plane reads the SAME A tile and keeps it in the collector, so shared memory
// only delivers the B operand.  Separates "bound by shared-memory operand reads" from "bound by MMA issue" (tools/chain_timeline.py).
template <bool P0>
__device__ __forceinline__ void issue_plane2_probe_collector(uint32_t d_tmem, uint32_t a_lo, uint32_t w_lo) {
  constexpr uint32_t kHiA = ((kHaloPitch * 128u) >> 4) | (1u << 14) | (2u << 29);
  constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t bo = (uint32_t)(tap * 8192 + (P0 ? 0 : 4096) + k * 32) >> 4;
      const uint32_t acc = (tap == 0 && k == 0) ? 0u : 1u;
      const bool first = tap == 0 && k == 0, last = tap == 8 && k == 3;
      if (first) umma_2sm<P0, 1>(d_tmem + (P0 ? 0u : 64u), desc64(a_lo, kHiA), desc64(w_lo + bo, kHiB), kIdescN64, acc);
      else if (last) umma_2sm<P0, 3>(d_tmem + (P0 ? 0u : 64u), desc64(a_lo, kHiA), desc64(w_lo + bo, kHiB), kIdescN64, acc);
      else umma_2sm<P0, 2>(d_tmem + (P0 ? 0u : 64u), desc64(a_lo, kHiA), desc64(w_lo + bo, kHiB), kIdescN64, acc);
    }
  }
}
}  // namespace two

// ---- host: launch with programmatic stream serialization (PDL) -------------------------------
template <typename Kern, typename... Args>
static cudaError_t launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace
}  // namespace pds
