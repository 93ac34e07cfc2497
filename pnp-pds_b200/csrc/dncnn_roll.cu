// DnCNN 64->64 3x3 body layer, row-streaming variant for large launches (sm_100a, cta_group::2).
//
// Reference: models/basic_models.py:31-35 (conv_list[i] + LeakyReLU), network_dncnn.py:60-66 (KAIR body).
//
// Why a second body kernel: with N = 64 output channels an SS-mode tcgen05.mma reads 4 KB of A from shared memory for
// 32 tensor-pipe cycles, i.e. the 16x8-tile kernels of dncnn_tc.cu are bound by shared-memory operand reads (ncu: l1tex tc
// wavefronts 89 % of peak, tensor pipe 35 % active), because every tap re-reads its own shifted view of the activations.
// Here the A operand is read ONCE per input row and x-shift and used for the three vertical taps:
//   * a CTA owns a strip of 128 consecutive pixels of an image row (UMMA M = 128 per CTA, 256 per CTA pair) and walks
//     down a band of rows.  Input row r (130 pixels incl. the x halo, one TMA row box per plane, zero-filled outside
//     the image = the convolution's padding) contributes to the output rows r-1, r, r+1 through the taps dy = 2, 1, 0;
//   * for every x-shift dx and k-step the three MMAs (one per dy, each into the accumulator of "its" output row) are
//     issued back to back with collector::a::fill / use / lastuse, so the tensor core keeps the A tile in its collector
//     buffer and shared memory is read once instead of three times (operand traffic per 128 pixels: 168 KB vs 360 KB);
//   * four output rows are live in TMEM (4 x [64 fp32 columns kind::f16 | 64 columns kind::f8f6f4] = 512 columns):
//     row r+1 is being started, r and r-1 accumulate, r-2 is drained by the epilogue warps while the next input row runs.
// Operand split, weight image (per CTA: 9 taps x [fp16 tile of 32 output channels | e4m3 tile]) and epilogue are those of
// two::conv_tc2_kernel (dncnn_tc.cu); the activation layout in HBM is unchanged, so the kernels are interchangeable.
#include "tc_common.cuh"

namespace pds {

namespace {
namespace roll {

constexpr int kStripW = 128;                              // output pixels per CTA and row
constexpr int kRowPix = kStripW + 2;                      // pixels landed per row box (x halo of 1 on both sides)
constexpr uint32_t kRowBytes = kRowPix * 128;             // 16640
constexpr uint32_t kRowSlot = 17 * 1024;                  // slot stride, 1024-B aligned
constexpr int kSlotsR = 8;                                // plane ring: four input rows in flight
constexpr uint32_t kWHalf = 9 * 64 * 128;                 // 73728: per-CTA weight image (same as two::conv_tc2_kernel)
constexpr uint32_t kOffAR = kWHalf, kOffBarR = kOffAR + kSlotsR * kRowSlot;
constexpr uint32_t kOffBiasR = kOffBarR + 256, kSmemBytesR = kOffBiasR + 256 + 1024;
constexpr uint32_t kIdescN64 = (1u << 4) | ((256u >> 4) << 24) | ((64u >> 3) << 17);   // D=f32, A=B=f16|e4m3, M=256, N=64

struct RollArgs {
  const __half* w_img;
  const float* bias;
  const __half* in;       // input activations (conv_roll_d_kernel reads the a_lo half of plane 1 with plain loads)
  __half* out;
  int write_a8;           // see store_half_row (tc_common.cuh)
  float slope, lo_scale;
  int H, W, nimg;
  int npairs_x;           // 256-pixel strip pairs per image row
  int total_rows;         // nimg * npairs_x * H: rows of all strip-pair columns, column after column
  int rows_per_cluster;   // contiguous share of that sequence owned by one CTA pair
  int dbg;                // timing probes of conv_roll_d_kernel's epilogue (wrong results): 1 = no activation stores, 2 = every warp-level store covers 1 KB of consecutive addresses
};

// The work of one CTA pair: rows [cid * rows_per_cluster, ...) of the global row sequence, cut into bands at the column
// (image / strip pair) boundaries.  Every pair gets the same number of rows (+ two halo rows per band), so there is no wave
// quantisation: a fixed band height of 64 on 74 pairs costs 462 row steps per pair at the cfg4 shape, this split 447.
struct BandWalk {
  int g, end, H, npairs_x;
  __device__ __forceinline__ BandWalk(const RollArgs& a, int cid)
      : g(cid * a.rows_per_cluster), end(min(a.total_rows, (cid + 1) * a.rows_per_cluster)), H(a.H), npairs_x(a.npairs_x) {}
  __device__ __forceinline__ bool next(int& img, int& px, int& yb, int& rb) {
    if (g >= end) return false;
    const int col = g / H;
    yb = g - col * H;
    rb = min(H - yb, end - g);
    img = col / npairs_x;
    px = col - img * npairs_x;
    g += rb;
    return true;
  }
};

// All MMAs of one plane of one input row.  MASK bit 2/1/0: the output rows r-1 / r / r+1 (taps dy = 2 / 1 / 0) exist in
// this band; d2/d1/d0 are the TMEM addresses of their accumulator blocks.  The block of row r+1 is started here
// (accumulate = 0 on its first MMA), the others already hold the contributions of earlier input rows.
template <bool F16, int MASK>
__device__ __forceinline__ void issue_row(uint32_t d2, uint32_t d1, uint32_t d0, uint32_t a_lo, uint32_t w_lo) {
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 8 contiguous 128-byte rows | version 1 | SWIZZLE_128B
  constexpr int kCount = ((MASK >> 2) & 1) + ((MASK >> 1) & 1) + (MASK & 1);
  constexpr uint32_t kCol = F16 ? 0u : 64u;
  constexpr uint32_t kTile = F16 ? 0u : 4096u;
  // collector usage of the 1st / 2nd / 3rd MMA of a group of kCount MMAs sharing A
  constexpr int kC2 = kCount > 1 ? 1 : 0;                                           // dy=2 is always first when present
  constexpr int kC1 = (MASK & 4) ? ((MASK & 1) ? 2 : 3) : ((MASK & 1) ? 1 : 0);     // middle, last, first or alone
  constexpr int kC0 = kCount > 1 ? 3 : 0;                                           // dy=0 is always last when present
#pragma unroll
  for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = desc64(a_lo + ((uint32_t)(dx * 128 + k * 32) >> 4), kHi);
      if constexpr ((MASK & 4) != 0)
        two::umma_2sm<F16, kC2>(d2 + kCol, ad, desc64(w_lo + ((uint32_t)((6 + dx) * 8192 + kTile + k * 32) >> 4), kHi), kIdescN64, 1u);
      if constexpr ((MASK & 2) != 0)
        two::umma_2sm<F16, kC1>(d1 + kCol, ad, desc64(w_lo + ((uint32_t)((3 + dx) * 8192 + kTile + k * 32) >> 4), kHi), kIdescN64, 1u);
      if constexpr ((MASK & 1) != 0)
        two::umma_2sm<F16, kC0>(d0 + kCol, ad, desc64(w_lo + ((uint32_t)((0 + dx) * 8192 + kTile + k * 32) >> 4), kHi), kIdescN64,
                                (dx == 0 && k == 0) ? 0u : 1u);
    }
  }
}

template <bool F16>
__device__ __forceinline__ void issue_row_masked(int mask, uint32_t d2, uint32_t d1, uint32_t d0, uint32_t a_lo, uint32_t w_lo) {
  switch (mask) {       // warp-uniform
    case 7: issue_row<F16, 7>(d2, d1, d0, a_lo, w_lo); break;
    case 6: issue_row<F16, 6>(d2, d1, d0, a_lo, w_lo); break;
    case 3: issue_row<F16, 3>(d2, d1, d0, a_lo, w_lo); break;
    case 4: issue_row<F16, 4>(d2, d1, d0, a_lo, w_lo); break;
    case 2: issue_row<F16, 2>(d2, d1, d0, a_lo, w_lo); break;
    case 1: issue_row<F16, 1>(d2, d1, d0, a_lo, w_lo); break;
    default: break;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    conv_roll_kernel(const __grid_constant__ CUtensorMap tmap, RollArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sA = base + kOffAR, sBar = base + kOffBarR;
  // barriers: full[8] @0 (used in CTA 0), empty[8] @64, wfull @128, tfull[4] @136, tempty[4] @168 (CTA 0), tmem slot @200
  const uint32_t bFull = sBar, bEmpty = sBar + 64, bW = sBar + 128, bTFull = sBar + 136, bTEmpty = sBar + 168;
  const uint32_t sTmemSlot = sBar + 200;
  float* bias_s = reinterpret_cast<float*>(gbase + kOffBiasR);
  const uint32_t rank = two::cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlotsR; ++i) {
      mbar_init(bFull + 8 * i, 1);          // CTA 0: one arrive.expect_tx for both CTAs' boxes
      mbar_init(bEmpty + 8 * i, 1);         // multicast commit from CTA 0
    }
    mbar_init(bW, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 8);        // 4 epilogue warps x 2 CTAs arrive on CTA 0's barrier
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) bias_s[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                          // barriers initialised, TMEM slot written
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bW, kWHalf);           // this CTA's half of the weight image
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_img) + (size_t)rank * kWHalf;
      for (int i = 0; i < 9; ++i) bulk_load(sW + i * 8192u, src + (size_t)i * 8192u, 8192u, bW);
    }
    __syncwarp();
    mbar_wait(bW, 0);
  }
  __syncthreads();
  two::cluster_sync_all();                  // both halves of the weights landed; all barriers of both CTAs are initialised
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kOffBarR + 200);
  pdl_launch_dependents();

  const int cid = blockIdx.x >> 1;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; boxes signal CTA 0's full barrier)
    pdl_wait_prior_grid();
    const uint32_t full0 = two::map_to_cta(bFull, 0);
    uint32_t j = 0;
    BandWalk walk(a, cid);
    for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
      const int x0 = (px * 2 + (int)rank) * kStripW;
      for (int s = 0; s < rb + 2; ++s) {
#pragma unroll
        for (int p = 0; p < 2; ++p, ++j) {
          const uint32_t slot = j % kSlotsR, use = j / kSlotsR;
          mbar_wait(bEmpty + 8 * slot, (use & 1) ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(bFull + 8 * slot, 2 * kRowBytes);
            two::tma_load_4d_2sm(sA + slot * kRowSlot, &tmap, full0 + 8 * slot, 0, x0 - 1, yb - 1 + s, img * 2 + p);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CTA 0 only)
    if (rank == 0) {
      const uint32_t w_lo = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t j = 0;
      uint32_t n0 = 0;                      // output rows started so far by this cluster: row n lives in TMEM block n & 3
      BandWalk walk(a, cid);
      for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
        for (int s = 0; s < rb + 2; ++s) {
          // input row yb-1+s feeds output rows s-2 (dy=2), s-1 (dy=1), s (dy=0) of the band
          const int mask = (s >= 2 ? 4 : 0) | ((s >= 1 && s <= rb) ? 2 : 0) | (s < rb ? 1 : 0);
          const uint32_t nn = n0 + (uint32_t)s;
          if (mask & 1) mbar_wait(bTEmpty + 8 * (nn & 3), ((nn >> 2) & 1) ^ 1);      // block of the row started here is drained
          const uint32_t d0 = tmem_base + (nn & 3) * 128u, d1 = tmem_base + ((nn - 1) & 3) * 128u, d2 = tmem_base + ((nn - 2) & 3) * 128u;
#pragma unroll
          for (int p = 0; p < 2; ++p, ++j) {
            const uint32_t slot = j % kSlotsR, use = j / kSlotsR;
            mbar_wait(bFull + 8 * slot, use & 1);
            tc_fence_after();
            const uint32_t a_lo = (((sA + slot * kRowSlot) & 0x3FFFFu) >> 4) | (1u << 16);
            if (elect_one()) {
              if (p == 0) issue_row_masked<true>(mask, d2, d1, d0, a_lo, w_lo);
              else issue_row_masked<false>(mask, d2, d1, d0, a_lo, w_lo);
              two::umma_commit_2sm(bEmpty + 8 * slot);
              if (p == 1 && (mask & 4)) two::umma_commit_2sm(bTFull + 8 * ((nn - 2) & 3));   // output row s-2 is complete
            }
            __syncwarp();
          }
        }
        n0 += (uint32_t)rb;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (each CTA drains its own 128 TMEM lanes = pixels)
    const int q = warp & 3;
    const int t = q * 32 + lane;
    const size_t hw = (size_t)a.H * a.W;
    const uint32_t tempty0 = two::map_to_cta(bTEmpty, 0);
    uint32_t n = 0;
    BandWalk walk(a, cid);
    for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
      const int x = (px * 2 + (int)rank) * kStripW + t;
      for (int jr = 0; jr < rb; ++jr, ++n) {
        const uint32_t blk = n & 3;
        mbar_wait(bTFull + 8 * blk, (n >> 2) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + blk * 128u;
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32(taddr + 0, r0);
        tmem_ld32(taddr + 64, r2);
        tmem_ld32(taddr + 32, r1);
        tmem_ld32(taddr + 96, r3);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) two::mbar_arrive_cluster(tempty0 + 8 * blk);      // block released before any arithmetic or store
        if (x < a.W) {
          const size_t pix = (size_t)(yb + jr) * a.W + x;
          __half* o_p0 = a.out + (((size_t)img * 2 + 0) * hw + pix) * 64;
          uint8_t* o_p1 = reinterpret_cast<uint8_t*>(a.out + (((size_t)img * 2 + 1) * hw + pix) * 64);
          store_half_row(o_p0, o_p1, r0, r2, bias_s, 0, a.slope, a.lo_scale, a.write_a8);
          store_half_row(o_p0, o_p1, r1, r3, bias_s, 32, a.slope, a.lo_scale, a.write_a8);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  two::cluster_sync_all();                  // the peer may still be reading TMEM / signalling our barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Same layer, with the e4m3(a) operand rebuilt on chip ("derive" form, the default).
// Plane 1 of the activations is [e4m3(fp16(v)) x 64 | e4m3((v - fp16(v)) 2^10) x 64]: its first half is a function of plane 0,
// so storing and re-loading it costs 128 of the layer's 512 HBM bytes per pixel for nothing.  Here
//   * the TMA producer of each CTA lands only the fp16 row (plane 0) in the "F" slot of a ring stage (local full barrier);
//   * four converter warps per CTA build the "E" slot of the stage — the K = 128 e4m3 operand row in the SWIZZLE_128B layout the
//     MMA descriptors expect: bytes 0..63 = cvt.rn.satfinite.e4m3x2.f16x2 of the F row (read back from shared memory), bytes
//     64..127 = the a_lo half of plane 1, fetched with plain 16-byte loads (ld.global.L2::64B, so that DRAM does not deliver
//     the unused half of the 128-byte line) one row ahead, zero outside the image like TMA; then fence.proxy.async (every
//     writing lane) + one relaxed cluster-scope arrive per warp on the stage's "ready" barrier in CTA 0.  Relaxed on
//     purpose, as for the TMEM-drained signal: a release at cluster scope compiles to MEMBAR.ALL.GPU per warp and row, and
//     the writes it would order are already performed (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC of the proxy fence);
//   * two converter warps, not four: with 8 warps per CTA every scheduler holds two and the register cap stays at 255; with
//     10 warps it drops to 168, ptxas spills in the MMA-issue loop and the tensor pipe starves (measured: 1.43 vs 1.12 ms).
//     The two warps take alternate rows (a whole row each): splitting every row between them made the per-row fixed costs
//     (barrier waits, proxy fence, load latency) the pacing stage;
//   * the F ring (5 slots) is one deeper than the E ring (4), so the row loads run ahead of the converters;
//   * the MMA issuer waits for "ready" only (it implies that both F rows have landed), and frees F and E separately;
//   * the layers feeding this kernel skip the e4m3(a) store (write_a8 = 0): 384 instead of 512 HBM bytes per pixel and layer.
// MMA order, accumulators and epilogue are those of conv_roll_kernel, and e4m3(fp16(v)) is what every producer stores, so the
// result is bit-identical to the other body kernels.
// ---------------------------------------------------------------------------------------------
constexpr int kConvWarps = 2;                             // 8 warps per CTA = 2 per scheduler: the register cap stays at 255 (10 warps: 168, spills)
constexpr int kConvThreads = 32 * kConvWarps;
constexpr int kThreadsD = kThreads + kConvThreads;
constexpr int kFD = 5, kED = 4;                           // F ring (fp16 rows, TMA) one deeper than the E ring (e4m3 rows, converters)
constexpr uint32_t kOffED = kOffAR + kFD * kRowSlot, kOffBarD = kOffED + kED * kRowSlot;
constexpr uint32_t kOffBiasD = kOffBarD + 256, kSmemBytesD = kOffBiasD + 256 + 1024;
static_assert(kSmemBytesD <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.L2::64B.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsD, 1)
    conv_roll_d_kernel(const __grid_constant__ CUtensorMap tmap, RollArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sF = base + kOffAR, sE = base + kOffED, sBar = base + kOffBarD;
  // barriers: fullF[5] @0 (local), emptyF[5] @40 (local), emptyE[4] @80 (local), ready[4] @112 (CTA 0), wfull @144,
  //           tfull[4] @152, tempty[4] @184 (CTA 0), tmem slot @216
  const uint32_t bFullF = sBar, bEmptyF = sBar + 40, bEmptyE = sBar + 80, bReady = sBar + 112, bW = sBar + 144;
  const uint32_t bTFull = sBar + 152, bTEmpty = sBar + 184, sTmemSlot = sBar + 216;
  float* bias_s = reinterpret_cast<float*>(gbase + kOffBiasD);
  const uint32_t rank = two::cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFD; ++i) {
      mbar_init(bFullF + 8 * i, 1);         // this CTA's producer: arrive.expect_tx for its own row box
      mbar_init(bEmptyF + 8 * i, 1 + 1);    // multicast commit of the fp16 MMAs + the converter warp that took the row
    }
    for (int i = 0; i < kED; ++i) {
      mbar_init(bEmptyE + 8 * i, 1);        // multicast commit of the e4m3 MMAs
      mbar_init(bReady + 8 * i, 2);         // one converter warp per CTA and row (used in CTA 0)
    }
    mbar_init(bW, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 8);        // 4 epilogue warps x 2 CTAs arrive on CTA 0's barrier
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) bias_s[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                          // barriers initialised, TMEM slot written
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bW, kWHalf);           // this CTA's half of the weight image
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_img) + (size_t)rank * kWHalf;
      for (int i = 0; i < 9; ++i) bulk_load(sW + i * 8192u, src + (size_t)i * 8192u, 8192u, bW);
    }
    __syncwarp();
    mbar_wait(bW, 0);
  }
  __syncthreads();
  two::cluster_sync_all();                  // both halves of the weights landed; all barriers of both CTAs are initialised
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kOffBarD + 216);
  pdl_launch_dependents();

  const int cid = blockIdx.x >> 1;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: this CTA's fp16 rows
    pdl_wait_prior_grid();
    uint32_t j = 0;
    BandWalk walk(a, cid);
    for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
      const int x0 = (px * 2 + (int)rank) * kStripW;
      for (int s = 0; s < rb + 2; ++s, ++j) {
        const uint32_t fs = j % kFD, fuse = j / kFD;
        mbar_wait(bEmptyF + 8 * fs, (fuse & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bFullF + 8 * fs, kRowBytes);
          tma_load_4d(sF + fs * kRowSlot, &tmap, bFullF + 8 * fs, 0, x0 - 1, yb - 1 + s, img * 2);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CTA 0 only)
    if (rank == 0) {
      const uint32_t w_lo = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t j = 0;
      uint32_t n0 = 0;                      // output rows started so far by this cluster: row n lives in TMEM block n & 3
      BandWalk walk(a, cid);
      for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
        for (int s = 0; s < rb + 2; ++s, ++j) {
          // input row yb-1+s feeds output rows s-2 (dy=2), s-1 (dy=1), s (dy=0) of the band
          const int mask = (s >= 2 ? 4 : 0) | ((s >= 1 && s <= rb) ? 2 : 0) | (s < rb ? 1 : 0);
          const uint32_t nn = n0 + (uint32_t)s;
          if (mask & 1) mbar_wait(bTEmpty + 8 * (nn & 3), ((nn >> 2) & 1) ^ 1);      // block of the row started here is drained
          const uint32_t d0 = tmem_base + (nn & 3) * 128u, d1 = tmem_base + ((nn - 1) & 3) * 128u, d2 = tmem_base + ((nn - 2) & 3) * 128u;
          const uint32_t fs = j % kFD, es = j % kED, euse = j / kED;
          mbar_wait(bReady + 8 * es, euse & 1);                           // E rows of both CTAs written (=> both F rows landed)
          tc_fence_after();
          const uint32_t f_lo = (((sF + fs * kRowSlot) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t e_lo = (((sE + es * kRowSlot) & 0x3FFFFu) >> 4) | (1u << 16);
          if (elect_one()) {
            issue_row_masked<true>(mask, d2, d1, d0, f_lo, w_lo);
            two::umma_commit_2sm(bEmptyF + 8 * fs);
            issue_row_masked<false>(mask, d2, d1, d0, e_lo, w_lo);
            two::umma_commit_2sm(bEmptyE + 8 * es);
            if (mask & 4) two::umma_commit_2sm(bTFull + 8 * ((nn - 2) & 3));        // output row s-2 is complete
          }
          __syncwarp();
        }
        n0 += (uint32_t)rb;
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ epilogue (each CTA drains its own 128 TMEM lanes = pixels)
    const int q = warp & 3;
    const int t = q * 32 + lane;
    const size_t hw = (size_t)a.H * a.W;
    const uint32_t tempty0 = two::map_to_cta(bTEmpty, 0);
    uint32_t n = 0;
    BandWalk walk(a, cid);
    for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
      const int x = (px * 2 + (int)rank) * kStripW + t;
      for (int jr = 0; jr < rb; ++jr, ++n) {
        const uint32_t blk = n & 3;
        mbar_wait(bTFull + 8 * blk, (n >> 2) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + blk * 128u;
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32(taddr + 0, r0);
        tmem_ld32(taddr + 64, r2);
        tmem_ld32(taddr + 32, r1);
        tmem_ld32(taddr + 96, r3);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) two::mbar_arrive_cluster(tempty0 + 8 * blk);      // block released before any arithmetic or store
        if (x < a.W && !(a.dbg & 1)) {
          const size_t pix = (size_t)(yb + jr) * a.W + x;
          __half* o_p0 = a.out + (((size_t)img * 2 + 0) * hw + pix) * 64;
          uint8_t* o_p1 = reinterpret_cast<uint8_t*>(a.out + (((size_t)img * 2 + 1) * hw + pix) * 64);
          uint32_t cs = 32;
          if (a.dbg & 2) {            // same 4 KB piece per warp and plane, lane-linear inside it
            o_p0 = o_p0 - (size_t)lane * 64 + lane * 16;
            o_p1 = o_p1 - (size_t)lane * 128 + lane * 32;
            cs = 1024;
          }
          store_half_row(o_p0, o_p1, r0, r2, bias_s, 0, a.slope, a.lo_scale, a.write_a8, cs);
          store_half_row(o_p0, o_p1, r1, r3, bias_s, 32, a.slope, a.lo_scale, a.write_a8, cs);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ converters: build the e4m3 operand row of every stage
    // The two converter warps take alternate rows (a whole row each), so each has two row periods for its fixed costs (barrier
    // waits, proxy fence, load latency).  Task k < 16 of a lane: pixel group k (8 consecutive pixels of the row box),
    // 16-channel quarter lane & 3; a quarter-warp (one 128-byte shared-memory transaction of the 16-byte accesses) covers the
    // two pixels {i, i ^ 5} of the group: their swizzle phases differ in bit 0 (the two fp16 reads land in disjoint 16-byte
    // columns) and in bit 2 (so do the two e4m3 writes) — no bank conflicts.  Task 16 (lanes 0..7): the pixels 128, 129.
    const size_t hw = (size_t)a.H * a.W;
    const uint32_t ready0 = two::map_to_cta(bReady, 0);
    const uint32_t mine = (uint32_t)(warp - 6);                          // rows with (j & 1) == mine
    constexpr int kFull = kRowPix / 8;                                   // 16
    constexpr int kTasks = kFull + 1;
    const bool tail = lane < 4 * (kRowPix - 8 * kFull);                  // lanes 0..7
    const int pig = ((lane >> 2) & 1) ? ((lane >> 3) ^ 5) : (lane >> 3);
    const uint32_t qd = (uint32_t)lane & 3u;
    auto task_px = [&](int k) { return k < kFull ? k * 8 + pig : 8 * kFull + (lane >> 2); };
    pdl_wait_prior_grid();                                               // plain loads of the previous layer's output below
    uint32_t j = 0;
    BandWalk walk(a, cid);
    for (int img, px, yb, rb; walk.next(img, px, yb, rb);) {
      const int x0 = (px * 2 + (int)rank) * kStripW - 1;                 // image x of box pixel 0
      const uint8_t* p1 = reinterpret_cast<const uint8_t*>(a.in) + ((size_t)img * 2 + 1) * hw * 128 + 64 + qd * 16;
      uint4 lo[kTasks];
      auto fetch = [&](int s) {                                          // a_lo chunks of input row yb-1+s (zero outside the image)
        const int y = yb - 1 + s;
        const bool yok = y >= 0 && y < a.H;
#pragma unroll
        for (int k = 0; k < kTasks; ++k) {
          const int x = x0 + task_px(k);
          const bool ok = yok && (k < kFull || tail) && x >= 0 && x < a.W;
          lo[k] = ok ? ldg128(p1 + ((size_t)y * a.W + x) * 128) : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      int s = (int)((j ^ mine) & 1u);                                    // this warp's first row of the band
      if (s < rb + 2) fetch(s);
      for (; s < rb + 2; s += 2) {
        const uint32_t jj = j + (uint32_t)s;
        const uint32_t fs = jj % kFD, fuse = jj / kFD, es = jj % kED, euse = jj / kED;
        const uint32_t F = sF + fs * kRowSlot, E = sE + es * kRowSlot;
        mbar_wait(bFullF + 8 * fs, fuse & 1);                            // fp16 row landed
        mbar_wait(bEmptyE + 8 * es, (euse & 1) ^ 1);                     // e4m3 MMAs of the previous use are done
#pragma unroll
        for (int k0 = 0; k0 < kTasks; k0 += 9) {                         // two batches: 18 / 16 shared-memory reads in flight
          uint4 h0[9], h1[9];
#pragma unroll
          for (int k = k0; k < k0 + 9 && k < kTasks; ++k) {
            if (k < kFull || tail) {
              const uint32_t p = (uint32_t)task_px(k), sw = p & 7u;
              h0[k - k0] = lds128(F + p * 128u + (((2u * qd) ^ sw) << 4));
              h1[k - k0] = lds128(F + p * 128u + (((2u * qd + 1u) ^ sw) << 4));
            }
          }
#pragma unroll
          for (int k = k0; k < k0 + 9 && k < kTasks; ++k) {
            if (k < kFull || tail) {
              const uint32_t p = (uint32_t)task_px(k), sw = p & 7u;
              uint4 a8;
              a8.x = e4m3x2_from_f16x2(h0[k - k0].x) | (e4m3x2_from_f16x2(h0[k - k0].y) << 16);
              a8.y = e4m3x2_from_f16x2(h0[k - k0].z) | (e4m3x2_from_f16x2(h0[k - k0].w) << 16);
              a8.z = e4m3x2_from_f16x2(h1[k - k0].x) | (e4m3x2_from_f16x2(h1[k - k0].y) << 16);
              a8.w = e4m3x2_from_f16x2(h1[k - k0].z) | (e4m3x2_from_f16x2(h1[k - k0].w) << 16);
              sts128(E + p * 128u + ((qd ^ sw) << 4), a8);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < kTasks; ++k) {                               // a_lo half (loaded while this warp's previous row was handled)
          if (k < kFull || tail) {
            const uint32_t p = (uint32_t)task_px(k), sw = p & 7u;
            sts128(E + p * 128u + (((4u + qd) ^ sw) << 4), lo[k]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA's reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bEmptyF + 8 * fs);                                 // F row no longer read by the converters
          two::mbar_arrive_cluster(ready0 + 8 * es);
        }
        if (s + 2 < rb + 2) fetch(s + 2);                                // this warp's next row (after the fence: it would wait for them)
      }
      j += (uint32_t)(rb + 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  two::cluster_sync_all();                  // the peer may still be reading TMEM / signalling our barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace roll
}  // namespace

int roll_setup() {
  cudaError_t e = cudaFuncSetAttribute(roll::conv_roll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roll::kSmemBytesR);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(roll::conv_roll_d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roll::kSmemBytesD);
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(conv_roll_kernel): ") + cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// Dispatch decision: rows per CTA pair (> 0) when the row-streaming kernels are the better choice for this launch, 0 for the
// tile kernels.  Cost model in units of "one row step of a CTA pair" (2 x 128 pixels): the launch splits the rows of all
// (image, strip pair) columns evenly over the pairs (BandWalk), a pair pays its rows plus two halo rows per band; strips that
// hang over the right image edge (width not a multiple of 256) are paid in full.  The 2-CTA tile kernel needs ~1.2 steps per
// tile pair (it re-reads the A operand: 3246 vs 2912 cycles per 128 pixels in ncu, and it re-fetches halos; calibrated on
// single 512x512 and 1024x1024 images).  `force`: ignore the comparison with the tile kernels (tests).
int roll_rows_per_pair(int nimg, int H, int W, int num_sms) {
  const long long npx = (W + 2 * roll::kStripW - 1) / (2 * roll::kStripW);
  const long long pairs = num_sms / 2 > 0 ? num_sms / 2 : 1;
  const long long total = (long long)nimg * npx * H;
  const long long rpc = (total + pairs - 1) / pairs;
  return (int)(rpc < 8 ? 8 : rpc);                                      // tiny launches: fewer pairs rather than halo-dominated bands
}
int roll_band_rows(int nimg, int H, int W, int num_sms, bool force) {
  if (W < roll::kStripW || H < 8 || nimg < 1) return 0;
  const long long pairs = num_sms / 2 > 0 ? num_sms / 2 : 1;
  const int rpc = roll_rows_per_pair(nimg, H, W, num_sms);
  if (force) return rpc;
  const double bands = 1.0 + (double)rpc / (double)H;                   // columns a pair's share touches, on average
  const double roll_cost = (double)rpc + 2.0 * bands;
  const long long tile_pairs = ((long long)nimg * ((H + kTileRows - 1) / kTileRows) * ((W + kTileCols - 1) / kTileCols) + 1) / 2;
  const double tile_cost = 1.2 * (double)((tile_pairs + pairs - 1) / pairs);
  return roll_cost < tile_cost ? rpc : 0;
}

cudaError_t launch_conv_mid_roll(TcPlan* plan, int in_buf, int nimg, int band_rows, const DncnnLayerW& L, float slope, int derive,
                                 int write_a8, cudaStream_t st) {
  roll::RollArgs a{};
  a.w_img = L.w_mid_tc2;
  a.bias = L.bias;
  a.in = plan->act[in_buf];
  a.out = plan->act[in_buf ^ 1];
  a.write_a8 = write_a8;
  a.dbg = plan->probe_bits;
  a.slope = slope;
  a.lo_scale = L.lo_scale;
  a.H = plan->H;
  a.W = plan->W;
  a.nimg = nimg;
  a.npairs_x = (plan->W + 2 * roll::kStripW - 1) / (2 * roll::kStripW);
  a.total_rows = nimg * a.npairs_x * plan->H;
  a.rows_per_cluster = band_rows > 0 ? band_rows : roll_rows_per_pair(nimg, plan->H, plan->W, plan->num_sms);
  const int nclusters = (a.total_rows + a.rows_per_cluster - 1) / a.rows_per_cluster;
  if (derive) return launch_pdl(roll::conv_roll_d_kernel, 2 * nclusters, roll::kThreadsD, roll::kSmemBytesD, st, plan->map_row[in_buf], a);
  return launch_pdl(roll::conv_roll_kernel, 2 * nclusters, kThreads, roll::kSmemBytesR, st, plan->map_row[in_buf], a);
}

}  // namespace pds
