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
  __half* out;
  float slope, lo_scale;
  int H, W, nimg;
  int band_rows, nbands, npairs_x, nunits;
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

  const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; boxes signal CTA 0's full barrier)
    pdl_wait_prior_grid();
    const uint32_t full0 = two::map_to_cta(bFull, 0);
    uint32_t j = 0;
    for (int unit = cid; unit < a.nunits; unit += nclusters) {
      const int px = unit % a.npairs_x, t = unit / a.npairs_x;
      const int band = t % a.nbands, img = t / a.nbands;
      const int yb = band * a.band_rows;
      const int rb = min(a.band_rows, a.H - yb);
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
      for (int unit = cid; unit < a.nunits; unit += nclusters) {
        const int band = (unit / a.npairs_x) % a.nbands;
        const int rb = min(a.band_rows, a.H - band * a.band_rows);
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
    for (int unit = cid; unit < a.nunits; unit += nclusters) {
      const int px = unit % a.npairs_x, tt = unit / a.npairs_x;
      const int band = tt % a.nbands, img = tt / a.nbands;
      const int yb = band * a.band_rows;
      const int rb = min(a.band_rows, a.H - yb);
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
          store_half_row(o_p0, o_p1, r0, r2, bias_s, 0, a.slope, a.lo_scale);
          store_half_row(o_p0, o_p1, r1, r3, bias_s, 32, a.slope, a.lo_scale);
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

}  // namespace roll
}  // namespace

int roll_setup() {
  cudaError_t e = cudaFuncSetAttribute(roll::conv_roll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roll::kSmemBytesR);
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(conv_roll_kernel): ") + cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// Rows per band, or 0 when the tile kernels are the better choice for this launch.  Cost model in units of "one 128-pixel row
// step of a CTA pair": a band of rb rows costs rb + 2 steps (two halo rows), the launch costs ceil(units / pairs) bands per
// pair; the 2-CTA tile kernel needs ~1.2 steps per tile pair (it re-reads the A operand: 3246 vs 2912 cycles per 128
// pixels in ncu, and it re-fetches halos; calibrated on single 512x512 and 1024x1024 images).
// `force`: ignore the comparison with the tile kernels (tests).
int roll_band_rows(int nimg, int H, int W, int num_sms, bool force) {
  if (W % roll::kStripW != 0 || H < 8 || nimg < 1) return 0;
  const long long npx = (W + 2 * roll::kStripW - 1) / (2 * roll::kStripW);
  const long long pairs = num_sms / 2 > 0 ? num_sms / 2 : 1;
  long long best_cost = -1;
  int best_rb = 0;
  for (int rb = 8; rb <= 64; ++rb) {
    const long long units = (long long)nimg * ((H + rb - 1) / rb) * npx;
    const long long cost = ((units + pairs - 1) / pairs) * (rb + 2);
    if (best_cost < 0 || cost <= best_cost) {
      best_cost = cost;
      best_rb = rb;
    }
  }
  if (force) return best_rb;
  const long long tile_pairs = ((long long)nimg * ((H + kTileRows - 1) / kTileRows) * ((W + kTileCols - 1) / kTileCols) + 1) / 2;
  const double tile_cost = 1.2 * (double)((tile_pairs + pairs - 1) / pairs);
  return (double)best_cost < tile_cost ? best_rb : 0;
}

cudaError_t launch_conv_mid_roll(TcPlan* plan, int in_buf, int nimg, int band_rows, const DncnnLayerW& L, float slope, cudaStream_t st) {
  roll::RollArgs a{};
  a.w_img = L.w_mid_tc2;
  a.bias = L.bias;
  a.out = plan->act[in_buf ^ 1];
  a.slope = slope;
  a.lo_scale = L.lo_scale;
  a.H = plan->H;
  a.W = plan->W;
  a.nimg = nimg;
  a.band_rows = band_rows;
  a.nbands = (plan->H + band_rows - 1) / band_rows;
  a.npairs_x = (plan->W + 2 * roll::kStripW - 1) / (2 * roll::kStripW);
  a.nunits = nimg * a.nbands * a.npairs_x;
  const int nclusters = a.nunits < plan->num_sms / 2 ? a.nunits : plan->num_sms / 2;
  return launch_pdl(roll::conv_roll_kernel, 2 * nclusters, kThreads, roll::kSmemBytesR, st, plan->map_row[in_buf], a);
}

}  // namespace pds
