// DnCNN 64->64 3x3 layers as a TMA-fed tcgen05 implicit GEMM (sm_100a).
//
// Reference: models/basic_models.py:31-35 (conv_list[i] + LeakyReLU), the only dense contraction
// on the PnP-PDS path: per layer  D[pixel, oc] = sum_{tap, ci} A[pixel+tap, ci] W[oc, tap, ci].
//
// Mapping
//   * one CTA tile = 16 rows x 8 pixels = 128 output pixels = UMMA M; N = 64 output channels;
//     K = 64 input channels per tap, 9 taps -> 36 k-steps of 16.
//   * operands are split as a = a_hi + a_lo, w = w_hi + w_lo with a_hi, w_hi fp16.  The leading product
//     a_hi*w_hi runs as kind::f16 MMAs (K = 64 per tap); the two first-order correction products
//     a*w_lo + a_lo*w_hi only need ~4 significant bits each, so they run as ONE kind::f8f6f4 MMA per
//     k-step with K = 128 e4m3 bytes per tap:  A8 = [e4m3(a) | e4m3(a_lo 2^10)],  B8 = [e4m3(w_lo 2^S) | e4m3(w_hi 2^(S-10))]
//     into a second fp32 TMEM accumulator that the epilogue folds in with the factor 2^-S.
//     72 tcgen05.mma per tile (was 108 with three fp16 products) at ~2^-16 relative operand precision;
//     needed for the 1e-4 iterate gate (single-pass fp16 misses it; tools/emulate_split.py
//     measures 3e-6 for this scheme against 6e-5 for one fp16 pass on the same loop).
//   * the activation halo tile (18 rows x 10 pixels x 128 B, one per plane) is fetched by
//     ONE 4-D TMA box each with SWIZZLE_128B; out-of-image pixels are zero-filled by TMA, which is
//     exactly the convolution's zero padding.  All 9 taps read that single tile: the A descriptor
//     of tap (dy,dx) starts at pixel-row offset (dy*10+dx)*128 B with SBO = one tile row (1280 B).
//     (Probed on B200, tests/test_gpu_tcgen05_probe.py: the 128B-swizzle phase is taken from the
//     shared-memory ADDRESS bits, so any 128-B-aligned start and any SBO multiple of 128 B work with
//     base_offset = 0.)
//   * the layer's weights (9 taps x [64x64 fp16 | 64x128 e4m3], pre-swizzled on the host) stay resident in
//     shared memory for the whole persistent CTA (147 KB), loaded once with cp.async.bulk.
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane,
//     warp-uniform control flow so descriptors live in uniform registers), warps 2-5 = epilogue
//     (tcgen05.ld -> bias + LeakyReLU -> fp16 plane + e4m3 plane -> 32-byte stores).
//     The accumulators are double-buffered in TMEM (2 x 128 columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1; activation planes go through a 3-slot ring (p0_t, p1_t,
//     p0_t+1, ...) so the producer runs up to three planes ahead of the tensor pipe.
//   Activation layout in HBM: [img][2][H][W][128 B]; plane 0 = fp16(v) x 64 channels, plane 1 =
//     e4m3(v) x 64 followed by e4m3((v - fp16(v)) * 2^10) x 64.
#include "tc_common.cuh"

namespace pds {

namespace {


// Compile-time geometry of a body layer with NW = 64 output channels (rows of the B operand per tile).
template <int NW>
struct Geo {
  static constexpr uint32_t kWTile = NW * 128;                 // one (tap, split) NW x 64 fp16 tile
  static constexpr uint32_t kWBytes = 9 * 2 * kWTile;          // 147456 (NW=64) / 36864 (NW=16), 1024-multiples
  static constexpr int kSlots = 3;                             // activation-plane ring depth
  static constexpr uint32_t kOffA = kWBytes, kOffBar = kOffA + kSlots * kPlaneSlot;
  static constexpr uint32_t kOffBias = kOffBar + 192, kSmemUsed = kOffBias + 256;
  static constexpr uint32_t kSmemBytes = kSmemUsed + 1024;     // slack for manual 1024-B alignment
  static constexpr uint32_t kAccCols = 2 * NW;                 // cols [0,NW): a_hi*w_hi (f16 kind) ; [NW,2NW): e4m3 correction (f8f6f4 kind)
  static constexpr uint32_t kTmemCols = 4 * NW < 32 ? 32 : 4 * NW;   // 2 accumulator stages
  static constexpr uint32_t kIdesc = kIdescBase | ((uint32_t)(NW >> 3) << 17);   // same bits for both kinds: format 0 = F16 / E4M3
};



// All MMAs of one activation plane (36 k-steps).  The weight image keeps, per tap, the fp16 tile w_hi (NW rows x 128 B)
// directly followed by the e4m3 tile [w_lo 2^S | w_hi 2^(S-10)] (NW rows x 128 B).  Plane 0 (fp16) accumulates
// a_hi*w_hi into columns [0,NW) with kind::f16 (K = 16 per MMA); plane 1 (e4m3, 128 K-bytes per pixel and tap)
// accumulates the correction into columns [NW,2NW) with kind::f8f6f4 (K = 32 per MMA) — the same 32 operand bytes
// per row and k-step for both, so the descriptor arithmetic is identical.
// a_lo / w_lo are the low descriptor words of the plane / weight bases; every offset is an immediate.
template <int NW, bool P0>
__device__ __forceinline__ void issue_plane(uint32_t d_tmem, uint32_t a_lo, uint32_t w_lo) {
  using G = Geo<NW>;
  constexpr uint32_t kHiA = ((kHaloPitch * 128u) >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
  constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ao = (uint32_t)((dy * kHaloPitch + dx) * 128 + k * 32) >> 4;
      const uint32_t bo = (uint32_t)((tap * 2 + (P0 ? 0 : 1)) * (int)G::kWTile + k * 32) >> 4;
      const uint32_t acc = (tap == 0 && k == 0) ? 0u : 1u;
      if (P0) umma_f16(d_tmem, desc64(a_lo + ao, kHiA), desc64(w_lo + bo, kHiB), G::kIdesc, acc);
      else umma_f8(d_tmem + NW, desc64(a_lo + ao, kHiA), desc64(w_lo + bo, kHiB), G::kIdesc, acc);
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcArgs a) {
  using G = Geo<NW>;
  constexpr int kSlots = G::kSlots;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sA = base + G::kOffA, sBar = base + G::kOffBar;
  // barriers: full[6] @0, empty[6] @48, wfull @96, tfull[2] @104, tempty[2] @120, tmem slot @136
  const uint32_t bFull = sBar, bEmpty = sBar + 48, bW = sBar + 96, bTFull = sBar + 104, bTEmpty = sBar + 120;
  const uint32_t sTmemSlot = sBar + 136;
  float* bias_s = reinterpret_cast<float*>(gbase + G::kOffBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(bFull + 8 * i, 1);
      mbar_init(bEmpty + 8 * i, 1);
    }
    mbar_init(bW, 1);
    mbar_init(bTFull, 1); mbar_init(bTFull + 8, 1);
    mbar_init(bTEmpty, 4); mbar_init(bTEmpty + 8, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) {
    const int c = threadIdx.x - 64;
    bias_s[c] = a.bias[c];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(G::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + G::kOffBar + 136);
  pdl_launch_dependents();

  const int per_img = a.tiles_x * a.tiles_y;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    if (elect_one()) {
      mbar_expect_tx(bW, G::kWBytes);
      for (int i = 0; i < 18; ++i)
        bulk_load(sW + i * G::kWTile, reinterpret_cast<const uint8_t*>(a.w_img) + (size_t)i * G::kWTile, G::kWTile, bW);
    }
    __syncwarp();
    pdl_wait_prior_grid();                                  // activations of the previous layer are complete and visible
    uint32_t j = 0;     // plane sequence number: 2*it + p
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y0 = (rem / a.tiles_x) * kTileRows, x0 = (rem % a.tiles_x) * kTileCols;
#pragma unroll
      for (int p = 0; p < 2; ++p, ++j) {
        const uint32_t slot = j % kSlots, use = j / kSlots;
        mbar_wait(bEmpty + 8 * slot, (use & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bFull + 8 * slot, kPlaneBytes);
          tma_load_4d(sA + slot * kPlaneSlot, &tmap, bFull + 8 * slot, 0, x0 - 1, y0 - 1, img * 2 + p);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    mbar_wait(bW, 0);
    const uint32_t w_lo = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
    uint32_t j = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      mbar_wait(bTEmpty + 8 * acc, (uint32_t)(((it >> 1) & 1) ^ 1));
      const uint32_t d_tmem = tmem_base + acc * G::kAccCols;
#pragma unroll
      for (int p = 0; p < 2; ++p, ++j) {
        const uint32_t slot = j % kSlots, use = j / kSlots;
        mbar_wait(bFull + 8 * slot, use & 1);
        tc_fence_after();
        const uint32_t a_lo = (((sA + slot * kPlaneSlot) & 0x3FFFFu) >> 4) | (1u << 16);
        if (elect_one()) {
          if (p == 0) issue_plane<NW, true>(d_tmem, a_lo, w_lo);
          else issue_plane<NW, false>(d_tmem, a_lo, w_lo);
          umma_commit(bEmpty + 8 * slot);              // slot may be overwritten once these MMAs retire
          if (p == 1) umma_commit(bTFull + 8 * acc);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const size_t hw = (size_t)a.H * a.W;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y = (rem / a.tiles_x) * kTileRows + ty, x = (rem % a.tiles_x) * kTileCols + tx;
      const uint32_t acc = it & 1;
      mbar_wait(bTFull + 8 * acc, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * G::kAccCols;
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 64, r2);
      tmem_ld32(taddr + 32, r1);
      tmem_ld32(taddr + 96, r3);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bTEmpty + 8 * acc);
      if (y < a.H && x < a.W) {
        const size_t pix = (size_t)y * a.W + x;
        __half* o_p0 = a.out + (((size_t)img * 2 + 0) * hw + pix) * 64;
        uint8_t* o_p1 = reinterpret_cast<uint8_t*>(a.out + (((size_t)img * 2 + 1) * hw + pix) * 64);
        store_half_row(o_p0, o_p1, r0, r2, bias_s, 0, a.slope, a.lo_scale, a.write_a8);
        store_half_row(o_p0, o_p1, r1, r3, bias_s, 32, a.slope, a.lo_scale, a.write_a8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// First layer (Cin = 1|3 -> 64) on the tensor pipe.  K = 9*Cin <= 27 is padded to 32 (two k-steps); the
// A operand is an explicit im2col tile that two groups of four producer warps build in shared memory: thread m
// owns pixel m of the 16x8 tile, reads its 3x3xCin window (zero padding, input clamp of denoiser.py:40), splits
// to fp16 hi/lo and writes the 64 meaningful bytes of its 128-byte row in SWIZZLE_128B order.  With only two
// k-steps the tensor work is negligible, so this layer keeps three exact fp16 products (N=128 [w_hi;w_lo] against
// the hi tile, N=64 against the lo tile); the epilogue is the body layers' (fp16 plane + e4m3 plane), on two groups
// of four warps.  The kernel is bound by issue slots and by writing the 256 B/pixel of activations.
// ---------------------------------------------------------------------------------------------
namespace first {
constexpr int kStages = 4, kThreadsF = 544;                 // 2 x 4 producer warps, 1 MMA warp, 2 x 4 epilogue warps
constexpr int kMmaWarpF = 8;
constexpr uint32_t kWBytesF = 128 * 128;                    // [w_hi 64 rows ; w_lo 64 rows] x 128 B
constexpr uint32_t kATile = 128 * 128;                      // one plane of one stage
constexpr uint32_t kOffAF = kWBytesF, kOffBarF = kOffAF + kStages * 2 * kATile;
constexpr uint32_t kOffBiasF = kOffBarF + 192, kOffStgF = kOffBiasF + 256;
constexpr uint32_t kSmemBytesF = kOffStgF + 4 * 18 * 10 * 3 * 4 + 1024;       // + double-buffered input window per producer group (Cin <= 3)
constexpr uint32_t kIdescN64 = kIdescBase | ((64u >> 3) << 17), kIdescN128 = kIdescBase | ((128u >> 3) << 17);

struct FirstArgs {
  const float* in;        // (nimg, CIN, H, W)
  const __half* w_img;    // swizzled [128][64] fp16 image
  const float* bias;
  __half* out;
  float slope;
  int clamp_in;
  int write_a8;           // see store_half_row (tc_common.cuh)
  int H, W, nimg, tiles_x, tiles_y, ntiles;
  int dbg;                // timing probes of the tap-shifted kernel (wrong results): 1 = no activation stores, 2 = no epilogue arithmetic, 4 = a third of the MMAs
};

template <int CIN>
__global__ void __launch_bounds__(kThreadsF, 1) conv_first_tc_kernel(FirstArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sA = base + kOffAF, sBar = base + kOffBarF;
  // barriers: full[4] @0, empty[4] @32, wfull @64, tfull[2] @72, tempty[2] @88, tmem slot @104
  const uint32_t bFull = sBar, bEmpty = sBar + 32, bW = sBar + 64, bTFull = sBar + 72, bTEmpty = sBar + 88, sTmemSlot = sBar + 104;
  float* bias_s = reinterpret_cast<float*>(gbase + kOffBiasF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bFull + 8 * i, 128);        // every producer thread arrives after its row is written
      mbar_init(bEmpty + 8 * i, 1);
    }
    mbar_init(bW, 1);
    mbar_init(bTFull, 1); mbar_init(bTFull + 8, 1);
    mbar_init(bTEmpty, 4); mbar_init(bTEmpty + 8, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 288 && threadIdx.x < 352) bias_s[threadIdx.x - 288] = a.bias[threadIdx.x - 288];
  if (warp == kMmaWarpF) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the A tiles once: chunks 4..7 of every row are never written again and never read by the two k-steps,
  // but the padded part of chunk 3 (k = 27..31) must be zero
  for (uint32_t i = threadIdx.x; i < kStages * 2 * kATile / 16; i += kThreadsF)
    *reinterpret_cast<uint4*>(gbase + kOffAF + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kOffBarF + 104);
  const int per_img = a.tiles_x * a.tiles_y;
  const size_t hw = (size_t)a.H * a.W;

  if (warp < kMmaWarpF) {
    // ------------------------------------------------------------ im2col producers (thread m = pixel m of the tile).  Two groups of
    // four warps; group g builds the tiles with it % 2 == g (its own staging window, named barrier and A stages g, g+2).
    if (threadIdx.x == 0) {
      mbar_expect_tx(bW, kWBytesF);
      bulk_load(sW, a.w_img, kWBytesF, bW);
    }
    const int grp = warp >> 2;
    const int m = threadIdx.x & 127, ty = m >> 3, tx = m & 7;
    // The (18 x 10 x Cin) input window of a tile is fetched cooperatively (each of the 128 producer threads loads
    // PER elements, coalesced along x), kDepth tiles ahead of its use so DRAM latency is off the critical path,
    // staged through a double-buffered shared-memory window, and then every thread gathers its own 3x3xCin patch.
    constexpr int kWinPix = kHaloRows * kHaloPitch, kWin = kWinPix * CIN, PER = (kWin + 127) / 128, kDepth = 3;
    float* stg = reinterpret_cast<float*>(gbase + kOffStgF);           // [2 groups][2][kWin]
    auto fetch = [&](int tile, float (&r)[PER]) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y0 = (rem / a.tiles_x) * kTileRows - 1, x0 = (rem % a.tiles_x) * kTileCols - 1;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int idx = m + 128 * j;
        float t = 0.f;
        if (idx < kWin) {
          const int c = idx / kWinPix, p = idx - c * kWinPix;
          const int gy = y0 + p / kHaloPitch, gx = x0 + p % kHaloPitch;
          if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) t = __ldg(a.in + ((size_t)(img * CIN + c) * a.H + gy) * a.W + gx);
        }
        r[j] = t;
      }
    };
    float pre[kDepth][PER];
    const int tstep = 2 * (int)gridDim.x;                         // tile stride of this group's sequence
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      const int t0 = blockIdx.x + grp * (int)gridDim.x + d * tstep;
      if (t0 < a.ntiles) fetch(t0, pre[d]);
    }
    int it = grp;
    for (int tile0 = blockIdx.x + grp * (int)gridDim.x; tile0 < a.ntiles; tile0 += kDepth * tstep) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d, it += 2) {
        const int tile = tile0 + d * tstep;
        if (tile >= a.ntiles) break;
        float* sw = stg + (grp * 2 + ((it >> 1) & 1)) * kWin;
        // the input clamp (denoiser.py:40) is applied here, kDepth tiles after the load was issued, so that nothing
        // touches a prefetched register while its load is still in flight
#pragma unroll
        for (int j = 0; j < PER; ++j)
          if (m + 128 * j < kWin) sw[m + 128 * j] = a.clamp_in ? fminf(fmaxf(pre[d][j], 0.f), 1.f) : pre[d][j];
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // window of this tile complete (the group's 4 warps)
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) v[(dy * 3 + dx) * CIN + ci] = sw[ci * kWinPix + (ty + dy) * kHaloPitch + tx + dx];
      const uint32_t stage = it % kStages, use = it / kStages;
      mbar_wait(bEmpty + 8 * stage, (use & 1) ^ 1);
      const uint32_t row_hi = sA + stage * 2 * kATile + (uint32_t)m * 128u, row_lo = row_hi + kATile;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = v[8 * j + 2 * e], v1 = v[8 * j + 2 * e + 1];
          const __half2 hh = __floats2half2_rn(v0, v1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const uint32_t off = (uint32_t)((j ^ (m & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
      }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(bFull + 8 * stage);
        // prefetch for the tile kDepth rounds ahead — issued AFTER the releasing arrive above, whose MEMBAR would otherwise wait
        // for these loads to return and expose the full DRAM latency on every tile
        if (tile + kDepth * tstep < a.ntiles) fetch(tile + kDepth * tstep, pre[d]);
      }
    }
  } else if (warp == kMmaWarpF) {
    // ------------------------------------------------------------ MMA issuer
    mbar_wait(bW, 0);
    const uint32_t w_lo = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, stage = it % kStages, use = it / kStages;
      mbar_wait(bTEmpty + 8 * acc, (uint32_t)(((it >> 1) & 1) ^ 1));
      mbar_wait(bFull + 8 * stage, use & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 128u;
      const uint32_t a_hi = (((sA + stage * 2 * kATile) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_lo = (((sA + stage * 2 * kATile + kATile) & 0x3FFFFu) >> 4) | (1u << 16);
      if (elect_one()) {
        umma_f16(d_tmem, desc64(a_hi, kHi), desc64(w_lo, kHi), kIdescN128, 0u);
        umma_f16(d_tmem, desc64(a_hi + 2, kHi), desc64(w_lo + 2, kHi), kIdescN128, 1u);
        umma_f16(d_tmem, desc64(a_lo, kHi), desc64(w_lo, kHi), kIdescN64, 1u);
        umma_f16(d_tmem, desc64(a_lo + 2, kHi), desc64(w_lo + 2, kHi), kIdescN64, 1u);
        umma_commit(bEmpty + 8 * stage);
        umma_commit(bTFull + 8 * acc);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: two groups of four warps (the epilogue of a tile is a
    // ~700-instruction dependent chain per thread and this layer has almost no tensor work to hide it behind); group g drains
    // the tiles with it % 2 == g, which are exactly the tiles of accumulator stage g
    const int q = warp & 3;                               // TMEM lane quadrant this warp may read
    const int grp = (warp - (kMmaWarpF + 1)) >> 2;
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    int it = grp;
    for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < a.ntiles; tile += 2 * gridDim.x, it += 2) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y = (rem / a.tiles_x) * kTileRows + ty, x = (rem % a.tiles_x) * kTileCols + tx;
      const uint32_t acc = (uint32_t)grp;
      mbar_wait(bTFull + 8 * acc, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128u;
      const bool st = y < a.H && x < a.W;
      const size_t pix = (size_t)y * a.W + x;
      __half* o_p0 = a.out + (((size_t)img * 2 + 0) * hw + pix) * 64;
      uint8_t* o_p1 = reinterpret_cast<uint8_t*>(a.out + (((size_t)img * 2 + 1) * hw + pix) * 64);
      {
        uint32_t r0[32], r2[32];
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 64, r2);
        tmem_ld_wait();
        if (st) store_half_row(o_p0, o_p1, r0, r2, bias_s, 0, a.slope, 1.f, a.write_a8);      // columns [64,128) hold a_hi*w_lo at scale 1
      }
      {
        uint32_t r1[32], r3[32];
        tmem_ld32(taddr + 32, r1);
        tmem_ld32(taddr + 96, r3);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(bTEmpty + 8 * acc);
        if (st) store_half_row(o_p0, o_p1, r1, r3, bias_s, 32, a.slope, 1.f, a.write_a8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpF) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}
}  // namespace first

// ---------------------------------------------------------------------------------------------
// First layer, tap-shifted form (the default; conv_first_tc_kernel above is kept as the cross-check, tc_variant bit 15, and serves
// widths with W % 4 != 0).  The im2col kernel above spends half of its issue slots building the 128 x 27 patch matrix (ncu:
// producers 47 % of the stall samples, epilogue 48 %).  Here nothing is gathered: every pixel of the input window (4 rows x 130
// pixels for a pair of 128-pixel tiles, see below) is ONE 32-byte record of 16 fp16 K-slots
//     [ a_hi(c) | a_lo(c) | a_hi(c) | 1 1 1 | 0 .. ]        c = 0..Cin-1,  a = a_hi + a_lo
// stored as two un-swizzled K-major planes (slots 0-7 / 8-15, 16 bytes per pixel each), and tap (dy,dx) is the same window read
// through a descriptor whose start address is shifted by (dy*130 + dx) * 16 bytes (8-row groups = 8 consecutive pixels, 128 bytes
// apart) — the trick of the body layers at K = 16 instead of 64.  The B rows of tap t hold
//     [ w_hi(c) | w_hi(c) | w_lo(c) | bias as three fp16 terms (centre tap only) | 0 .. ]
// so one M=128, N=64, K=16 MMA per tap accumulates a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (+ bias): 9 MMAs per tile into ONE fp32
// accumulator, no correction accumulator, no bias add in the epilogue.  The fp32 window is landed by TMA (one 136 x 4 x Cin box
// per tile pair); three builder warps turn it into the 520 records of a pair; sixteen epilogue warps in four groups drain four
// TMEM stages and store whole 128-byte lines through per-warp staging rows.
// ---------------------------------------------------------------------------------------------
namespace first2 {
constexpr int kStagesA = 3, kAcc = 4;
// one builder warp per record stage and per window slot: a warp always returns to the same stage / slot, so it can never be two
// barrier phases ahead of their consumers (a parity wait cannot tell phase n from phase n + 2)
constexpr int kBuilders = 3, kGrpThreads = 32, kMmaWarp2 = 6, kEpiWarp0 = 8, kThreadsF2 = (kEpiWarp0 + 4 * kAcc) * 32;   // 768
// Tile = 128 consecutive pixels of ONE image row (M = 128): its activations are 16 KB of consecutive addresses in plane 0 (DRAM
// pages, full lines).  Work unit = a PAIR of vertically adjacent tiles: one window of 4 rows x 130 pixels (4 x Cin TMA rows of 544
// bytes), ONE record block, and the two tiles' MMAs read it at row offsets 0 and 1 into two TMEM stages — every hand-off of the
// pipeline (window -> records -> MMAs -> TMEM stage -> epilogue) carries two tiles, and a third of the records are shared.
// (With 16 x 8-pixel tiles the stores were 1 KB runs 128 KB apart and the window was 54 TMA rows of 64 bytes.)
constexpr int kTW = 128, kPairRows = 2, kWinRows = kPairRows + 2, kWinPitch = kTW + 2;
constexpr int kWinPix = kWinRows * kWinPitch;                       // 520 records per tile pair
constexpr int kRecPerThread = (kWinPix + 31) / 32;                  // 17
constexpr uint32_t kChunkPlane = kWinPix * 16;                      // 8320 bytes: K-slots 0-7 (plane 0) / 8-15 (plane 1) of every record
constexpr uint32_t kStageBytes = 2 * kChunkPlane;
constexpr uint32_t kWTap = 2 * 64 * 16, kWBytesF2 = 9 * kWTap;      // per tap: [chunk][oc][8 halves]
constexpr int kStg = 3, kBoxW = kTW + 8, kBoxX = 3;                  // input-window ring.  The box starts 4 pixels left of the tile and is 136 wide:
                                                                     // TMA wants the innermost start coordinate and extent in 16-byte multiples
                                                                     // (measured: tools/probe/tma_f32_probe.cu — x0 = 7 is an illegal instruction, 8 is fine)
template <int CIN> struct Stg {
  static constexpr uint32_t kBytes = (uint32_t)(CIN * kWinRows * kBoxW * 4);
  static constexpr uint32_t kSlot = (kBytes + 127u) & ~127u;
};
constexpr uint32_t kOffA2 = kWBytesF2, kOffStg2 = kOffA2 + kStagesA * kStageBytes, kOffBar2 = kOffStg2 + kStg * Stg<3>::kSlot;
constexpr uint32_t kOffWst2 = kOffBar2 + 256;                        // per epilogue warp: 32 rows x 128 B (plane 0) + 32 x 128 B (plane 1)
constexpr uint32_t kSmemBytesF2 = kOffWst2 + 4 * kAcc * 8192 + 128;
static_assert(kOffWst2 % 128 == 0, "staging rows are 128-byte aligned");
static_assert(kOffStg2 % 128 == 0, "TMA destination alignment");
constexpr uint32_t kIdescF2 = kIdescBase | ((64u >> 3) << 17);
static_assert(kOffA2 % 16 == 0 && kStageBytes % 16 == 0 && kOffBar2 % 8 == 0, "alignment");
static_assert(kBuilders == kStagesA && kBuilders == kStg, "one builder warp per stage and slot (barrier phases)");

// K-major, no swizzle: 8 rows x 16 bytes per core matrix (rows 16 bytes apart), lbo = distance between the two K chunks of an MMA,
// sbo = distance between 8-row groups
__device__ __forceinline__ uint64_t desc_k_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

// 32 channels [c0, c0+32) of one pixel from ONE accumulator that already holds the bias: LeakyReLU, then the three encodings of
// store_half_row (tc_common.cuh) — fp16(v) | e4m3(fp16(v)) (if write_a8) | e4m3((v - fp16(v)) 2^10) — written to the warp's
// staging rows in shared memory (row = lane = pixel, 16-byte chunk c of a 128-byte row at c ^ (lane & 7): conflict-free both ways).
// The warp then copies the rows out with every store instruction covering whole 128-byte lines (flush_rows below): with each
// thread storing its own pixel's 32-byte pieces directly, a warp-level store touches 32 different lines, and the L1 store-request
// rate — 768 requests per tile — was what bound this kernel (probe with consecutive addresses: 0.70 -> 0.45 ms per launch).
__device__ __forceinline__ void stage_half_row1(uint32_t row_p0, uint32_t row_p1, int lane, const uint32_t (&d)[32], int half, float slope,
                                                int write_a8) {
  uint32_t a8[8], l8[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t hi[8];
    float l[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = q * 16 + 2 * k;
      float2 v = make_float2(__uint_as_float(d[c]), __uint_as_float(d[c + 1]));
      const float2 vs = mul2(v, make_float2(slope, slope));
      v.x = fmaxf(v.x, vs.x);
      v.y = fmaxf(v.y, vs.y);
      const __half2 hh = __floats2half2_rn(v.x, v.y);
      const float2 lo = mul2(sub2(v, __half22float2(hh)), make_float2(kActLoScale, kActLoScale));
      hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
      l[2 * k] = lo.x;
      l[2 * k + 1] = lo.y;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t chunk = (uint32_t)((half * 4 + q * 2 + k) ^ (lane & 7));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_p0 + chunk * 16u), "r"(hi[4 * k]), "r"(hi[4 * k + 1]), "r"(hi[4 * k + 2]),
                   "r"(hi[4 * k + 3]) : "memory");
    }
    if (write_a8) {
#pragma unroll
      for (int k = 0; k < 4; ++k) a8[q * 4 + k] = e4m3x2_from_f16x2(hi[2 * k]) | (e4m3x2_from_f16x2(hi[2 * k + 1]) << 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) l8[q * 4 + k] = pack_e4m3x4(l[4 * k], l[4 * k + 1], l[4 * k + 2], l[4 * k + 3]);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (write_a8) {
      const uint32_t ca = (uint32_t)((half * 2 + k) ^ (lane & 7));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_p1 + ca * 16u), "r"(a8[4 * k]), "r"(a8[4 * k + 1]), "r"(a8[4 * k + 2]),
                   "r"(a8[4 * k + 3]) : "memory");
    }
    const uint32_t cl = (uint32_t)((4 + half * 2 + k) ^ (lane & 7));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_p1 + cl * 16u), "r"(l8[4 * k]), "r"(l8[4 * k + 1]), "r"(l8[4 * k + 2]),
                 "r"(l8[4 * k + 3]) : "memory");
  }
}

// Copies chunks [C0, 8) of the warp's 32 staged pixel rows (32 consecutive pixels of an image row) to global memory:
// lane -> (pixel, 16-byte chunk), 32 / (8 - C0) pixels per instruction, every instruction writes whole lines of consecutive
// addresses (512 bytes for C0 = 0).  g0 = address of this lane's chunk in the first instruction; left = pixels inside the image.
template <int C0>
__device__ __forceinline__ void flush_rows(uint32_t stage, uint8_t* g0, int lane, int left) {
  constexpr int NC = 8 - C0, PPI = 32 / NC, NI = 32 / PPI;
  const int px0 = lane / NC, c = C0 + lane % NC;
  uint32_t v[NI][4];
#pragma unroll
  for (int j = 0; j < NI; ++j) {                                // all shared-memory reads first: independent, latencies overlap
    const int r = j * PPI + px0;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3])
                 : "r"(stage + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) * 16)));
  }
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    if (j * PPI + px0 < left)
      asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(g0 + (size_t)(j * PPI) * 128), "r"(v[j][0]), "r"(v[j][1]), "r"(v[j][2]),
                   "r"(v[j][3])
                   : "memory");
  }
}

// n / d for a divisor fixed per launch (round-up multiplier, Granlund-Montgomery): the tile -> (image, row, column) split is done by
// every warp for every tile, and an integer division is ~25 instructions
struct FastDiv {
  uint32_t mul, sh1, sh2, d;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.mul, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int CIN>
__global__ void __launch_bounds__(kThreadsF2, 1) conv_first2_kernel(const __grid_constant__ CUtensorMap tmap_in, first::FirstArgs a, FastDiv div_img,
                                                                    FastDiv div_tx) {
  static_assert(3 * CIN + 3 <= 16, "K-slots of one record");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sW = base, sA = base + kOffA2, sBar = base + kOffBar2;
  // barriers: fullA[4] @0, emptyA[4] @32, wfull @64, tfull[4] @72, tempty[4] @104, tmem slot @136, stgfull[4] @144, stgempty[4] @176
  const uint32_t bFull = sBar, bEmpty = sBar + 32, bW = sBar + 64, bTFull = sBar + 72, bTEmpty = sBar + 104, sTmemSlot = sBar + 136;
  const uint32_t bSFull = sBar + 144, bSEmpty = sBar + 176, sStg = base + kOffStg2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesA; ++i) {
      mbar_init(bFull + 8 * i, kGrpThreads);       // every thread of the builder group arrives after its records are written
      mbar_init(bEmpty + 8 * i, 1);
    }
    mbar_init(bW, 1);
    for (int i = 0; i < kStg; ++i) {
      mbar_init(bSFull + 8 * i, 1);
      mbar_init(bSEmpty + 8 * i, kGrpThreads / 32);
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_in) : "memory");
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + kOffBar2 + 136);
  const int per_img = a.tiles_x * a.tiles_y;
  const size_t hw = (size_t)a.H * a.W;

  if (warp < kBuilders) {
    // ------------------------------------------------------------ record builders.  The window itself (Cin x 3 x 136 floats,
    // zero-filled outside the image = the convolution's padding) is landed by TMA, so no thread of this kernel has a global load
    // in flight when it reaches the proxy fence below (fence.proxy.async = MEMBAR + FENCE.VIEW.ASYNC: with register-prefetched
    // loads every tile waited out a DRAM round trip there).  Three builder warps, warp g builds the tile pairs with it % 3 == g (17
    // records per lane): a tile's build is one serial chain (wait for the window, build, wait for a free stage, store, proxy
    // fence, arrive) whose hand-offs cost more than its arithmetic — measured with the epilogue switched off: two groups of
    // three warps delivered a tile every 1 120 cycles.
    if (threadIdx.x == 0) {
      mbar_expect_tx(bW, kWBytesF2);
      bulk_load(sW, a.w_img, kWBytesF2, bW);
    }
    const int grp = warp;
    const int t = lane;
    float* const stg_base = reinterpret_cast<float*>(smem_raw + kOffStg2);
    int off[kRecPerThread];
    bool active[kRecPerThread];
#pragma unroll
    for (int j = 0; j < kRecPerThread; ++j) {
      const int p = t + j * kGrpThreads;
      active[j] = p < kWinPix;
      const int hy = p / kWinPitch, hx = p - hy * kWinPitch;
      off[j] = hy * kBoxW + hx + kBoxX;
    }
    for (int it = grp, tile = blockIdx.x + grp * (int)gridDim.x; tile < a.ntiles; tile += kBuilders * (int)gridDim.x, it += kBuilders) {
      const uint32_t slot = it % kStg;
      mbar_wait(bSFull + 8 * slot, (uint32_t)((it / kStg) & 1));
      float v[kRecPerThread][CIN];
#pragma unroll
      for (int j = 0; j < kRecPerThread; ++j)
#pragma unroll
        for (int c = 0; c < CIN; ++c) v[j][c] = active[j] ? stg_base[slot * (Stg<CIN>::kSlot / 4) + c * (kWinRows * kBoxW) + off[j]] : 0.f;
      // fp16 hi / lo split of the (clamped, denoiser.py:40) input.  The "1" slots are set in every record: they only meet
      // non-zero weights (the bias) in the centre tap, whose pixel is the output pixel itself.
      uint32_t rec[kRecPerThread][8];
#pragma unroll
      for (int j = 0; j < kRecPerThread; ++j) {
        __half h16[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) h16[k2] = __float2half_rn(0.f);
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float x = a.clamp_in ? fminf(fmaxf(v[j][c], 0.f), 1.f) : v[j][c];
          const __half hi = __float2half_rn(x);
          const __half lo = __float2half_rn(x - __half2float(hi));
          h16[c] = hi;
          h16[CIN + c] = lo;
          h16[2 * CIN + c] = hi;
        }
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) h16[3 * CIN + k2] = __float2half_rn(1.f);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2)
          rec[j][k2] = (uint32_t)__half_as_ushort(h16[2 * k2]) | ((uint32_t)__half_as_ushort(h16[2 * k2 + 1]) << 16);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bSEmpty + 8 * slot);          // the window slot may be refilled
      const uint32_t stage = it % kStagesA, use = it / kStagesA;
      mbar_wait(bEmpty + 8 * stage, (use & 1) ^ 1);
#pragma unroll
      for (int j = 0; j < kRecPerThread; ++j)
        if (active[j]) {
          const uint32_t dst = sA + stage * kStageBytes + (uint32_t)(t + j * kGrpThreads) * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(rec[j][0]), "r"(rec[j][1]), "r"(rec[j][2]), "r"(rec[j][3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kChunkPlane), "r"(rec[j][4]), "r"(rec[j][5]), "r"(rec[j][6]), "r"(rec[j][7]) : "memory");
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(bFull + 8 * stage);
    }
  } else if (warp == kMmaWarp2 + 1) {
    // ------------------------------------------------------------ TMA producer: one (136 x 4 x Cin) box per tile pair
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const int img = (int)fdiv((uint32_t)tile, div_img), rem = tile - img * per_img;
        const int trow = (int)fdiv((uint32_t)rem, div_tx);
        const int y0 = trow * kPairRows - 1, x0 = (rem - trow * a.tiles_x) * kTW - 1 - kBoxX;
        const uint32_t slot = it % kStg;
        mbar_wait(bSEmpty + 8 * slot, (uint32_t)(((it / kStg) & 1) ^ 1));
        mbar_expect_tx(bSFull + 8 * slot, Stg<CIN>::kBytes);
        tma_load_4d(sStg + slot * Stg<CIN>::kSlot, &tmap_in, bSFull + 8 * slot, x0, y0, 0, img);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp2) {
    // ------------------------------------------------------------ MMA issuer: nine K=16 MMAs per tile, one per tap; the two tiles
    // of a pair read the same records one window row apart and fill the TMEM stages 2 * (it % 2) + {0, 1}
    mbar_wait(bW, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const uint32_t stage = it % kStagesA, use = it / kStagesA;
      mbar_wait(bFull + 8 * stage, use & 1);
      const uint32_t a0 = sA + stage * kStageBytes;
#pragma unroll
      for (int r = 0; r < kPairRows; ++r) {
        const uint32_t acc = (uint32_t)(2 * (it & 1) + r);
        mbar_wait(bTEmpty + 8 * acc, (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64u;
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if ((a.dbg & 4) && t >= 3) break;              // timing probe: a third of the MMAs
            const uint32_t shift = (uint32_t)((t / 3 + r) * kWinPitch + (t % 3)) * 16u;
            umma_f16(d_tmem, desc_k_none(a0 + shift, kChunkPlane, 128), desc_k_none(sW + t * kWTap, 64 * 16, 128), kIdescF2,
                     t ? 1u : 0u);
          }
          if (r == kPairRows - 1) umma_commit(bEmpty + 8 * stage);
          umma_commit(bTFull + 8 * acc);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue: four groups of four warps, group g drains TMEM stage g
    // = row g & 1 of the tile pairs with it % 2 == g >> 1
    const int q = warp & 3;                               // TMEM lane quadrant this warp may read
    const int grp = (warp - kEpiWarp0) >> 2;
    const int prow = grp & 1;
    int it = grp >> 1;
    for (int tile = blockIdx.x + (grp >> 1) * (int)gridDim.x; tile < a.ntiles; tile += 2 * gridDim.x, it += 2) {
      const uint32_t acc = (uint32_t)grp;
      mbar_wait(bTFull + 8 * acc, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u;
      const uint32_t wst = base + kOffWst2 + (uint32_t)(warp - kEpiWarp0) * 8192u;
      const uint32_t row_p0 = wst + (uint32_t)lane * 128u, row_p1 = row_p0 + 4096u;
      {
        uint32_t r0[32];
        tmem_ld32(taddr, r0);
        tmem_ld_wait();
        if (!(a.dbg & 2)) stage_half_row1(row_p0, row_p1, lane, r0, 0, a.slope, a.write_a8);
      }
      {
        uint32_t r1[32];
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(bTEmpty + 8 * acc);
        if (!(a.dbg & 2)) stage_half_row1(row_p0, row_p1, lane, r1, 1, a.slope, a.write_a8);
      }
      __syncwarp();
      if (!(a.dbg & 1)) {
        const int img = (int)fdiv((uint32_t)tile, div_img), rem = tile - img * per_img;
        const int trow = (int)fdiv((uint32_t)rem, div_tx);
        const int tx0 = (rem - trow * a.tiles_x) * kTW + q * 32;          // this warp's 32 pixels of image row y
        const int y = trow * kPairRows + prow;
        const int left = y < a.H ? a.W - tx0 : 0;                         // pixels of the piece inside the image (may be <= 0)
        uint8_t* p0 = reinterpret_cast<uint8_t*>(a.out) + ((size_t)img * 2 * hw + (size_t)y * a.W + tx0) * 128;
        uint8_t* p1 = p0 + hw * 128;
        flush_rows<0>(wst, p0 + (lane >> 3) * 128 + (lane & 7) * 16, lane, left);
        if (a.write_a8) flush_rows<0>(wst + 4096u, p1 + (lane >> 3) * 128 + (lane & 7) * 16, lane, left);
        else flush_rows<4>(wst + 4096u, p1 + (lane >> 2) * 128 + (4 + (lane & 3)) * 16, lane, left);
      }
      __syncwarp();                                     // the staging rows are rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}
}  // namespace first2

// ---------------------------------------------------------------------------------------------
// Last layer (64 -> Cout = 1|3).  With N this small, one MMA group per tap is bound by re-reading the same A tile from
// shared memory nine times.  Instead the contraction over the 64 input channels is done ONCE per halo pixel for all
// nine taps at the same time:
//     P[q][tap*C + c] = sum_ci act[q][ci] * W[c][tap][ci]          q over the 18x10 halo tile,  N = 9*C <= 27 -> 32 columns
// as two M=128 MMA blocks over the contiguous halo tile (halo pixels 0..127 and 56..183: both 1024-byte aligned; the rows past
// pixel 179 are never gathered), and the epilogue gathers
//     out[y][x][c] = sum_tap P[(y+dy)*10 + (x+dx)][tap*C + c]  + bias, residual, clamp
// through a shared-memory copy of P.  The kernel is bound by streaming the activations, so it reads as few of them as the
// operand split allows: the fp16 plane (128 B per pixel) and the a_lo half of plane 1 (64 B per pixel, its own SWIZZLE_64B
// tensor map and K-major descriptors) — not e4m3(fp16(a)), which the body layers use for their a*w_lo correction.  With almost no
// tensor work here, that correction runs in fp16 on the plane that is loaded anyway:
//     kind::f16,    K = 64:  a_hi x [ w_hi ; w_lo 2^S ]          N = 64: columns 0-31 the product, 32-63 the first correction
//     kind::f8f6f4, K = 64:  e4m3(a_lo 2^10) x e4m3(w_hi 2^(S-10))   N = 32: the second correction, same scale 2^S
// So no body layer stores e4m3(fp16(a)) for this kernel's sake (192 instead of 256 bytes per pixel read here).
// ---------------------------------------------------------------------------------------------
namespace last {
constexpr int kSlotsL = 4;                                // tile slots: fp16 plane box + a_lo half box
constexpr int kThreadsL = 64 + 2 * 128;                   // TMA warp, MMA warp, two epilogue groups of four warps
constexpr int kNL = 32;                                   // B rows (tap*C + c), zero beyond 9*C
constexpr uint32_t kW16 = 2 * kNL * 128, kW8 = kNL * 64, kWBytesL = kW16 + kW8;   // fp16 tile [w_hi ; w_lo 2^S] (SW128) + e4m3 tile (SW64)
constexpr int kHaloPix = kHaloRows * kHaloPitch;          // 180
constexpr int kBlk1 = 56;                                 // second MMA block starts at halo pixel 56
constexpr uint32_t kLoBytes = kHaloPix * 64;              // 11520 bytes landed by the a_lo box
constexpr uint32_t kLoSlot = 12 * 1024;
constexpr uint32_t kTileSlot = kPlaneSlot + kLoSlot;      // 35 KB
constexpr int kPRows = kBlk1 + 128;                       // P rows written (184; rows >= 180 are dead)
constexpr int kPStride = 29;                              // floats per halo pixel in the P copy (odd: conflict-free rows)
constexpr uint32_t kOffAL = 11 * 1024, kOffBarL = kOffAL + kSlotsL * kTileSlot;
constexpr uint32_t kOffBiasL = kOffBarL + 192, kOffPL = kOffBiasL + 64;
constexpr uint32_t kSmemBytesL = kOffPL + 2 * kPRows * kPStride * 4 + 1024;
constexpr uint32_t kIdescL16 = kIdescBase | ((uint32_t)(2 * kNL >> 3) << 17), kIdescL8 = kIdescBase | ((uint32_t)(kNL >> 3) << 17);
constexpr uint32_t kAccCols = 192;                        // per TMEM stage: [block 0: 64 | block 1: 64 | f8 block 0: 32 | f8 block 1: 32]
static_assert(kWBytesL <= kOffAL && kPlaneBytes + (kPRows - kHaloPix) * 128 <= kPlaneSlot && kPRows * 64 <= kLoSlot, "slot slack");

template <int C>
__global__ void __launch_bounds__(kThreadsL, 1) conv_last_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_lo,
                                                                    TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sA = base + kOffAL, sBar = base + kOffBarL;
  // barriers: full[4] @0, empty[4] @32, wfull @64, tfull[2] @72, tempty[2] @88, tmem slot @104
  const uint32_t bFull = sBar, bEmpty = sBar + 32, bW = sBar + 64, bTFull = sBar + 72, bTEmpty = sBar + 88;
  const uint32_t sTmemSlot = sBar + 104;
  float* bias_s = reinterpret_cast<float*>(gbase + kOffBiasL);
  float* P = reinterpret_cast<float*>(gbase + kOffPL);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlotsL; ++i) {
      mbar_init(bFull + 8 * i, 1);
      mbar_init(bEmpty + 8 * i, 1);
    }
    mbar_init(bW, 1);
    mbar_init(bTFull, 1); mbar_init(bTFull + 8, 1);
    mbar_init(bTEmpty, 4); mbar_init(bTEmpty + 8, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_lo) : "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C) bias_s[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kOffBarL + 104);
  pdl_launch_dependents();

  const int per_img = a.tiles_x * a.tiles_y;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: two boxes per tile on one barrier
    if (elect_one()) {
      mbar_expect_tx(bW, kWBytesL);
      bulk_load(sW, a.w_img, kWBytesL, bW);
    }
    __syncwarp();
    pdl_wait_prior_grid();
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++j) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y0 = (rem / a.tiles_x) * kTileRows, x0 = (rem % a.tiles_x) * kTileCols;
      const uint32_t slot = j % kSlotsL, use = j / kSlotsL;
      mbar_wait(bEmpty + 8 * slot, (use & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(bFull + 8 * slot, kPlaneBytes + kLoBytes);
        tma_load_4d(sA + slot * kTileSlot, &tmap, bFull + 8 * slot, 0, x0 - 1, y0 - 1, img * 2);
        tma_load_4d(sA + slot * kTileSlot + kPlaneSlot, &tmap_lo, bFull + 8 * slot, 32, x0 - 1, y0 - 1, img * 2 + 1);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: 2 row blocks x (4 fp16 + 2 e4m3 k-steps)
    mbar_wait(bW, 0);
    const uint32_t w16 = ((sW & 0x3FFFFu) >> 4) | (1u << 16), w8 = (((sW + kW16) & 0x3FFFFu) >> 4) | (1u << 16);
    constexpr uint32_t kHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO = 8 contiguous 128-byte rows, SWIZZLE_128B
    constexpr uint32_t kHi64 = (512u >> 4) | (1u << 14) | (4u << 29);       // SBO = 8 contiguous 64-byte rows, SWIZZLE_64B
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++j) {
      const uint32_t acc = j & 1, slot = j % kSlotsL, use = j / kSlotsL;
      mbar_wait(bTEmpty + 8 * acc, (uint32_t)(((j >> 1) & 1) ^ 1));
      mbar_wait(bFull + 8 * slot, use & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      const uint32_t a16 = (((sA + slot * kTileSlot) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a8 = (((sA + slot * kTileSlot + kPlaneSlot) & 0x3FFFFu) >> 4) | (1u << 16);
      if (elect_one()) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_tmem + (uint32_t)(b * 64), desc64(a16 + ((uint32_t)(b * kBlk1 * 128 + k * 32) >> 4), kHi128), desc64(w16 + ((uint32_t)(k * 32) >> 4), kHi128),
                     kIdescL16, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f8(d_tmem + (uint32_t)(128 + b * 32), desc64(a8 + ((uint32_t)(b * kBlk1 * 64 + k * 32) >> 4), kHi64), desc64(w8 + ((uint32_t)(k * 32) >> 4), kHi64),
                    kIdescL8, k ? 1u : 0u);
        }
        umma_commit(bEmpty + 8 * slot);
        umma_commit(bTFull + 8 * acc);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> P (shared) -> 3x3 gather -> planar fp32.
    // Two groups of four warps: a tile's epilogue is one dependent chain per thread (TMEM read, P write, barrier, gather, store)
    // and this layer has only 12 MMAs per tile to hide it behind; group g drains the tiles with it % 2 == g, i.e. accumulator
    // stage g, through its own P buffer and named barrier.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int t = q * 32 + lane;
    const int ty = t >> 3, tx = t & 7;
    const size_t hw = (size_t)a.H * a.W;
    int it = grp;
    for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < a.ntiles; tile += 2 * gridDim.x, it += 2) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y = (rem / a.tiles_x) * kTileRows + ty, x = (rem % a.tiles_x) * kTileCols + tx;
      const uint32_t acc = (uint32_t)grp;
      // residual input of this thread's pixel: issued before the wait so its latency hides behind the MMAs
      const bool live = y < a.H && x < a.W;
      const size_t pix = (size_t)y * a.W + x;
      float xin[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xin[c] = live ? __ldg(a.net_in + ((size_t)img * C + c) * hw + pix) : 0.f;
      mbar_wait(bTFull + 8 * acc, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols;
      float* Pb = P + grp * (kPRows * kPStride);
      float p0[9 * C], p1[9 * C];
      {
        uint32_t r0[32], r1[32], r2[32];
        tmem_ld32(taddr, r0);                    // block 0: a_hi w_hi
        tmem_ld32(taddr + 32, r1);               //          a_hi (w_lo 2^S)
        tmem_ld32(taddr + 128, r2);              //          (a_lo 2^10) (w_hi 2^(S-10))
        tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 9 * C; ++n) p0[n] = fmaf(__uint_as_float(r1[n]) + __uint_as_float(r2[n]), a.lo_scale, __uint_as_float(r0[n]));
      }
      {
        uint32_t r0[32], r1[32], r2[32];
        tmem_ld32(taddr + 64, r0);               // block 1
        tmem_ld32(taddr + 96, r1);
        tmem_ld32(taddr + 160, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(bTEmpty + 8 * acc);
#pragma unroll
        for (int n = 0; n < 9 * C; ++n) p1[n] = fmaf(__uint_as_float(r1[n]) + __uint_as_float(r2[n]), a.lo_scale, __uint_as_float(r0[n]));
      }
      // the group's previous tile has been gathered out of Pb by all four warps
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      {
        float* row0 = Pb + t * kPStride;                   // halo pixel t            (block 0)
#pragma unroll
        for (int n = 0; n < 9 * C; ++n) row0[n] = p0[n];
        if (t >= 128 - kBlk1 && kBlk1 + t < kHaloPix) {    // halo pixel 56 + t >= 128 (block 1; the rest duplicates block 0 or is dead)
          float* row1 = Pb + (kBlk1 + t) * kPStride;
#pragma unroll
          for (int n = 0; n < 9 * C; ++n) row1[n] = p1[n];
        }
      }
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // P of this tile complete
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (live) {
        // out = clamp(sign * (conv + bias) + clamp(net_in))   (basic_models.py:36, denoiser.py:40-42, network_dncnn.py:77)
        float sum[C];
#pragma unroll
        for (int c = 0; c < C; ++c) sum[c] = bias_s[c];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float* pr = Pb + ((ty + dy) * kHaloPitch + tx + dx) * kPStride + (dy * 3 + dx) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) sum[c] += pr[c];
          }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float xi = xin[c];
          if (a.clamp) xi = fminf(fmaxf(xi, 0.f), 1.f);
          float o = a.res_sign > 0.f ? sum[c] + xi : xi - sum[c];
          if (a.clamp) o = fminf(fmaxf(o, 0.f), 1.f);
          a.out_f32[((size_t)img * C + c) * hw + pix] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}
}  // namespace last

// ---------------------------------------------------------------------------------------------
// 2-CTA variant of the body layer (cta_group::2): a cluster of two CTAs works on two pixel tiles at
// once as ONE M=256 UMMA issued by CTA 0.  The B operand (weights) is split between the two CTAs'
// shared memories, so every SM reads only half of B per MMA and keeps only half of the weight image
// (72 KB instead of 144 KB): the SS-mode MMA at N=64 is bound by shared-memory operand reads (4 KB of A
// per 32 tensor cycles), so halving the B bytes matters, and the freed room holds a 5-slot plane ring.
//   per-CTA weight image, per tap: [ fp16 w_hi[32r:32r+32] (4 KB) | e4m3 [w_lo 2^S | w_hi 2^(S-10)][32r:32r+32] (4 KB) ]
//   B rows (N = 64, either plane):  [ oc 0..31 | oc 32..63 ]   (CTA 0 | CTA 1)
//   accumulator columns per stage: [0,64) a_hi*w_hi (kind::f16), [64,128) e4m3 correction * 2^S (kind::f8f6f4)
// ---------------------------------------------------------------------------------------------
namespace two {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sA = base + kOffA2, sBar = base + kOffBar2;
  // barriers: full[5] @0 (used in CTA 0), empty[5] @40, wfull @80, tfull[4] @88, tempty[4] @120 (CTA 0), tmem slot @152
  const uint32_t bFull = sBar, bEmpty = sBar + 40, bW = sBar + 80, bTFull = sBar + 88, bTEmpty = sBar + 120;
  const uint32_t sTmemSlot = sBar + 152;
  float* bias_s = reinterpret_cast<float*>(gbase + kOffBias2);
  const uint32_t rank = cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots2; ++i) {
      mbar_init(bFull + 8 * i, 1);          // CTA 0: one arrive.expect_tx for both CTAs' boxes
      mbar_init(bEmpty + 8 * i, 1);         // multicast commit from CTA 0
    }
    mbar_init(bW, 1);
    for (int i = 0; i < kAccStages2; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 8);        // 4 epilogue warps x 2 CTAs arrive on CTA 0's barrier
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) bias_s[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(kTmemCols2) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                          // barriers initialised, TMEM slot written
  if (warp == 0) {
    if (elect_one()) {
      // this CTA's half of the weight image
      mbar_expect_tx(bW, kWHalf);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_img) + (size_t)rank * kWHalf;
      for (int i = 0; i < 9; ++i) bulk_load(sW + i * 8192u, src + (size_t)i * 8192u, 8192u, bW);
    }
    __syncwarp();
    mbar_wait(bW, 0);                       // own half landed ...
  }
  __syncthreads();
  cluster_sync_all();                       // ... and the peer's too; all barriers of both CTAs are initialised
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kOffBar2 + 152);
  pdl_launch_dependents();

  const int per_img = a.tiles_x * a.tiles_y;
  const int npairs = (a.ntiles + 1) >> 1;
  const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; boxes signal CTA 0's full barrier)
    pdl_wait_prior_grid();
    uint32_t j = 0;
    for (int pair = cid; pair < npairs; pair += nclusters) {
      int tile = 2 * pair + (int)rank;
      if (tile >= a.ntiles) tile = a.ntiles - 1;            // odd tail: load a valid tile, its result is not stored
      const int img = tile / per_img, rem = tile - img * per_img;
      const int y0 = (rem / a.tiles_x) * kTileRows, x0 = (rem % a.tiles_x) * kTileCols;
#pragma unroll
      for (int p = 0; p < 2; ++p, ++j) {
        const uint32_t slot = j % kSlots2, use = j / kSlots2;
        mbar_wait(bEmpty + 8 * slot, (use & 1) ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(bFull + 8 * slot, 2 * kPlaneBytes);
          tma_load_4d_2sm(sA + slot * kPlaneSlot, &tmap, map_to_cta(bFull + 8 * slot, 0), 0, x0 - 1, y0 - 1, img * 2 + p);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CTA 0 only)
    if (rank == 0) {
      const uint32_t w_lo = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t j = 0;
      int it = 0;
      for (int pair = cid; pair < npairs; pair += nclusters, ++it) {
        const uint32_t acc = it % kAccStages2;
        mbar_wait(bTEmpty + 8 * acc, (uint32_t)(((it / kAccStages2) & 1) ^ 1));
        const uint32_t d_tmem = tmem_base + acc * kAccCols2;
#pragma unroll
        for (int p = 0; p < 2; ++p, ++j) {
          const uint32_t slot = j % kSlots2, use = j / kSlots2;
          mbar_wait(bFull + 8 * slot, use & 1);
          tc_fence_after();
          const uint32_t a_lo = (((sA + slot * kPlaneSlot) & 0x3FFFFu) >> 4) | (1u << 16);
          if (elect_one()) {
            if (p == 0) issue_plane2<true>(d_tmem, a_lo, w_lo);
            else issue_plane2<false>(d_tmem, a_lo, w_lo);
            umma_commit_2sm(bEmpty + 8 * slot);
            if (p == 1) umma_commit_2sm(bTFull + 8 * acc);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (each CTA drains its own 128 TMEM lanes)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const size_t hw = (size_t)a.H * a.W;
    const uint32_t tempty0 = map_to_cta(bTEmpty, 0);
    int it = 0;
    for (int pair = cid; pair < npairs; pair += nclusters, ++it) {
      const int tile = 2 * pair + (int)rank;
      const bool live = tile < a.ntiles;
      const int tl = live ? tile : a.ntiles - 1;
      const int img = tl / per_img, rem = tl - img * per_img;
      const int y = (rem / a.tiles_x) * kTileRows + ty, x = (rem % a.tiles_x) * kTileCols + tx;
      const uint32_t acc = it % kAccStages2;
      mbar_wait(bTFull + 8 * acc, (uint32_t)((it / kAccStages2) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols2;
      const bool st = live && y < a.H && x < a.W;
      const size_t pix = (size_t)y * a.W + x;
      __half* o_p0 = a.out + (((size_t)img * 2 + 0) * hw + pix) * 64;
      uint8_t* o_p1 = reinterpret_cast<uint8_t*>(a.out + (((size_t)img * 2 + 1) * hw + pix) * 64);
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(taddr + 0, r0);
      tmem_ld32(taddr + 64, r2);
      tmem_ld32(taddr + 32, r1);
      tmem_ld32(taddr + 96, r3);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);      // stage released before any arithmetic or store
      if (st) {
        store_half_row(o_p0, o_p1, r0, r2, bias_s, 0, a.slope, a.lo_scale, a.write_a8);
        store_half_row(o_p0, o_p1, r1, r3, bias_s, 32, a.slope, a.lo_scale, a.write_a8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still be reading TMEM / signalling our barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
  }
}
}  // namespace two

// ---------------------------------------------------------------------------------------------
// Probes (test hooks): establish, on hardware, how a tcgen05 shared-memory descriptor addresses
// memory, and what a TMA box load leaves in shared memory.
// ---------------------------------------------------------------------------------------------
struct ProbeArgs {
  uint32_t a_off, sbo, base_off, region_bytes;
  float* out;   // [128][16]
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  // region [0, region_bytes): A pattern; then 2 KiB B tile; then barrier + tmem slot
  __half* A = reinterpret_cast<__half*>(g);
  const uint32_t nchunk = a.region_bytes / 16;
  for (uint32_t c = threadIdx.x; c < nchunk; c += blockDim.x)
    for (int e = 0; e < 8; ++e) A[c * 8 + e] = __float2half_rn((e & 1) ? (float)(c >> 10) : (float)(c & 1023u));
  __half* B = reinterpret_cast<__half*>(g + a.region_bytes);
  for (uint32_t i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const uint32_t n = i / 64, kk = i % 64;               // logical B[n][kk]
    const uint32_t chunk = (kk >> 3) ^ (n & 7);
    B[n * 64 + chunk * 8 + (kk & 7)] = __float2half_rn(kk == n ? 1.f : 0.f);
  }
  const uint32_t sBar = base + a.region_bytes + 2048, sSlot = sBar + 16;
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sSlot), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(g + a.region_bytes + 2048 + 16);
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad = make_desc(base + a.a_off, a.sbo, a.base_off);
    const uint64_t bd = make_desc(base + a.region_bytes, 1024, 0);
    umma_f16(tmem_base, ad, bd, idesc, 0);
    umma_commit(sBar);
  }
  mbar_wait(sBar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[16];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) a.out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
  }
}

__global__ void __launch_bounds__(128, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c1, int c2, int c3,
                                                           uint4* out /* kPlaneBytes/16 */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  const uint32_t sBar = base + kPlaneSlot;
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(sBar, kPlaneBytes);
    tma_load_4d(base, &tmap, sBar, 0, c1, c2, c3);
  }
  mbar_wait(sBar, 0);
  const uint4* s = reinterpret_cast<const uint4*>(g);
  for (uint32_t i = threadIdx.x; i < kPlaneBytes / 16; i += blockDim.x) out[i] = s[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// box_c = 64: whole 128-byte pixel rows (SWIZZLE_128B); box_c = 32: half rows (the last layer's a_lo box, SWIZZLE_64B)
int make_act_map(CUtensorMap* map, __half* act, int nimg, int H, int W, int box_w, int box_h, int box_c = 64) {
  EncodeTiledFn enc = get_encode();
  PDS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg * 2};
  const cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, act, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   box_c == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B,     // half rows: no promotion to whole lines
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return 0;
}

first2::FastDiv make_fastdiv(uint32_t d) {
  first2::FastDiv f{};
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                       // ceil(log2 d)
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l < 1 ? 0 : l - 1;
  return f;
}

// fp32 planar network input (nimg, C, H, W) as a 4-D tensor map with a (136 x 3 x C x 1) box, cached per (pointer, planes, C)
const CUtensorMap* plan_input_map(TcPlan* plan, const float* in, int planes, int C) {
  for (auto& e : plan->in_maps)
    if (e.ptr == in && e.planes == planes && e.C == C) return &e.map;
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return nullptr;
  TcPlan::InMap e{};
  e.ptr = in;
  e.planes = planes;
  e.C = C;
  const cuuint64_t dims[4] = {(cuuint64_t)plan->W, (cuuint64_t)plan->H, (cuuint64_t)C, (cuuint64_t)(planes / C)};
  const cuuint64_t strides[3] = {(cuuint64_t)plan->W * 4, (cuuint64_t)plan->H * plan->W * 4, (cuuint64_t)C * plan->H * plan->W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)first2::kBoxW, (cuuint32_t)first2::kWinRows, (cuuint32_t)C, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return nullptr;
  if (plan->in_maps.size() >= 32) plan->in_maps.erase(plan->in_maps.begin());
  plan->in_maps.push_back(e);
  return &plan->in_maps.back().map;
}

}  // namespace


int tc_num_sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

int tc_plan_create(int nimg, int H, int W, __half* act0, __half* act1, TcPlan** out) {
  TcPlan* p = new TcPlan();
  p->act[0] = act0;
  p->act[1] = act1;
  p->nimg = nimg;
  p->H = H;
  p->W = W;
  p->num_sms = tc_num_sms();
  int rc = make_act_map(&p->map[0], act0, nimg, H, W, kHaloPitch, kHaloRows);
  if (!rc) rc = make_act_map(&p->map[1], act1, nimg, H, W, kHaloPitch, kHaloRows);
  if (!rc) rc = make_act_map(&p->map_row[0], act0, nimg, H, W, 130, 1);      // dncnn_roll.cu: 128-pixel strip + x halo
  if (!rc) rc = make_act_map(&p->map_row[1], act1, nimg, H, W, 130, 1);
  if (!rc) rc = make_act_map(&p->map_lo[0], act0, nimg, H, W, kHaloPitch, kHaloRows, 32);    // last layer: a_lo half of plane 1
  if (!rc) rc = make_act_map(&p->map_lo[1], act1, nimg, H, W, kHaloPitch, kHaloRows, 32);
  if (!rc) rc = roll_setup();
  if (!rc) rc = chain_setup();
  if (!rc) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<64>::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(last::conv_last_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)last::kSmemBytesL);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(last::conv_last_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)last::kSmemBytesL);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(two::conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)two::kSmemBytes2);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(first::conv_first_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first::kSmemBytesF);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(first::conv_first_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first::kSmemBytesF);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(first2::conv_first2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first2::kSmemBytesF2);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(first2::conv_first2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first2::kSmemBytesF2);
    if (e != cudaSuccess) {
      set_error(std::string("cudaFuncSetAttribute(conv_tc_kernel): ") + cudaGetErrorString(e));
      rc = 1;
    }
  }
  if (rc) {
    delete p;
    return rc;
  }
  *out = p;
  return 0;
}

void tc_plan_set_probe_bits(TcPlan* p, int bits) { p->probe_bits = bits; }

void tc_plan_destroy(TcPlan* p) {
  if (!p) return;
  if (p->chain_layers) cudaFree(p->chain_layers);
  if (p->chain_flags) cudaFree(p->chain_flags);
  if (p->chain_trace) cudaFree(p->chain_trace);
  delete p;
}

static void fill_common(TcArgs& a, TcPlan* plan, int nimg) {
  a.H = plan->H;
  a.W = plan->W;
  a.nimg = nimg;
  a.tiles_x = (plan->W + kTileCols - 1) / kTileCols;
  a.tiles_y = (plan->H + kTileRows - 1) / kTileRows;
  a.ntiles = a.tiles_x * a.tiles_y * nimg;
}

cudaError_t launch_conv_mid_tc(TcPlan* plan, int in_buf, int nimg, const DncnnLayerW& L, float slope, cudaStream_t st) {
  TcArgs a{};
  a.w_img = L.w_mid_tc;
  a.bias = L.bias;
  a.out = plan->act[in_buf ^ 1];
  a.slope = slope;
  a.lo_scale = L.lo_scale;
  a.write_a8 = 1;
  a.C = 64;
  fill_common(a, plan, nimg);
  const int grid = a.ntiles < plan->num_sms ? a.ntiles : plan->num_sms;
  return launch_pdl(conv_tc_kernel<64>, grid, kThreads, Geo<64>::kSmemBytes, st, plan->map[in_buf], a);
}

cudaError_t launch_conv_mid_tc2(TcPlan* plan, int in_buf, int nimg, const DncnnLayerW& L, float slope, cudaStream_t st) {
  TcArgs a{};
  a.w_img = L.w_mid_tc2;
  a.bias = L.bias;
  a.out = plan->act[in_buf ^ 1];
  a.slope = slope;
  a.lo_scale = L.lo_scale;
  a.write_a8 = 1;
  a.C = 64;
  fill_common(a, plan, nimg);
  const int npairs = (a.ntiles + 1) / 2;
  const int nclusters = npairs < plan->num_sms / 2 ? npairs : plan->num_sms / 2;
  return launch_pdl(two::conv_tc2_kernel, 2 * nclusters, kThreads, two::kSmemBytes2, st, plan->map[in_buf], a);
}

cudaError_t launch_conv_first_tc(TcPlan* plan, int nimg, int C, const float* in, const DncnnLayerW& L, float slope, int clamp_in,
                                 int write_a8, int im2col, cudaStream_t st) {
  first::FirstArgs a{};
  a.in = in;
  a.w_img = im2col ? L.w_first_tc : L.w_first_tc2;
  a.bias = L.bias;
  a.out = plan->act[0];
  a.slope = slope;
  a.clamp_in = clamp_in;
  a.write_a8 = write_a8;
  a.H = plan->H;
  a.W = plan->W;
  a.nimg = nimg;
  a.tiles_x = (plan->W + kTileCols - 1) / kTileCols;
  a.tiles_y = (plan->H + kTileRows - 1) / kTileRows;
  a.ntiles = a.tiles_x * a.tiles_y * nimg;
  const int grid = a.ntiles < plan->num_sms ? a.ntiles : plan->num_sms;
  // the tap-shifted kernel lands its input windows by TMA: rows must be 16-byte multiples (W % 4 == 0) from a 16-byte aligned base;
  // other shapes take the im2col kernel
  if (!im2col && (plan->W & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (C == 1 || C == 3)) {
    const CUtensorMap* m = plan_input_map(plan, in, nimg * C, C);
    a.dbg = plan->probe_bits;
    a.tiles_x = (plan->W + first2::kTW - 1) / first2::kTW;      // strips of 128 pixels, two image rows per work unit
    a.tiles_y = (plan->H + first2::kPairRows - 1) / first2::kPairRows;
    a.ntiles = a.tiles_x * a.tiles_y * nimg;
    const int grid2 = a.ntiles < plan->num_sms ? a.ntiles : plan->num_sms;
    if (m == nullptr) return cudaErrorInvalidValue;
    const first2::FastDiv di = make_fastdiv((uint32_t)(a.tiles_x * a.tiles_y)), dx = make_fastdiv((uint32_t)a.tiles_x);
    if (C == 1) first2::conv_first2_kernel<1><<<grid2, first2::kThreadsF2, first2::kSmemBytesF2, st>>>(*m, a, di, dx);
    else first2::conv_first2_kernel<3><<<grid2, first2::kThreadsF2, first2::kSmemBytesF2, st>>>(*m, a, di, dx);
    return cudaGetLastError();
  }
  a.w_img = L.w_first_tc;
  if (C == 1) first::conv_first_tc_kernel<1><<<grid, first::kThreadsF, first::kSmemBytesF, st>>>(a);
  else if (C == 3) first::conv_first_tc_kernel<3><<<grid, first::kThreadsF, first::kSmemBytesF, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_conv_last_tc(TcPlan* plan, int in_buf, int nimg, int C, const DncnnLayerW& L, const float* net_in, float residual_sign,
                                int clamp, float* out, cudaStream_t st) {
  TcArgs a{};
  a.w_img = L.w_last_tc;
  a.bias = L.bias;
  a.net_in = net_in;
  a.out_f32 = out;
  a.C = C;
  a.res_sign = residual_sign;
  a.clamp = clamp;
  a.lo_scale = L.lo_scale;
  fill_common(a, plan, nimg);
  const int grid = a.ntiles < plan->num_sms ? a.ntiles : plan->num_sms;
  if (C == 1) return launch_pdl(last::conv_last_tc_kernel<1>, grid, last::kThreadsL, last::kSmemBytesL, st, plan->map[in_buf], plan->map_lo[in_buf], a);
  if (C == 3) return launch_pdl(last::conv_last_tc_kernel<3>, grid, last::kThreadsL, last::kSmemBytesL, st, plan->map[in_buf], plan->map_lo[in_buf], a);
  return cudaErrorInvalidValue;
}

}  // namespace pds

// ---- debug exports (declared in include/pnp_pds.h under "test hooks") ----
extern "C" int pds_debug_umma_probe(unsigned a_off, unsigned sbo, unsigned base_off, unsigned region_bytes, float* out_host /*128*16*/) {
  using namespace pds;
  PDS_REQUIRE(out_host && region_bytes % 1024 == 0 && region_bytes <= 200 * 1024, "bad probe arguments");
  float* d = nullptr;
  PDS_CUDA_OK(cudaMalloc(&d, 128 * 16 * sizeof(float)));
  const size_t smem = (size_t)region_bytes + 2048 + 64 + 1024;
  PDS_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProbeArgs a{a_off, sbo, base_off, region_bytes, d};
  umma_probe_kernel<<<1, 128, smem>>>(a);
  PDS_CUDA_OK(cudaGetLastError());
  PDS_CUDA_OK(cudaDeviceSynchronize());
  PDS_CUDA_OK(cudaMemcpy(out_host, d, 128 * 16 * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

extern "C" int pds_debug_tma_probe(const void* act_dev, int nimg, int H, int W, int x, int y, int plane_index, void* out_host) {
  using namespace pds;
  CUtensorMap map;
  int rc = make_act_map(&map, (__half*)act_dev, nimg, H, W, kHaloPitch, kHaloRows);
  if (rc) return rc;
  uint4* d = nullptr;
  PDS_CUDA_OK(cudaMalloc(&d, kPlaneBytes));
  const size_t smem = (size_t)kPlaneSlot + 64 + 1024;
  PDS_CUDA_OK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tma_probe_kernel<<<1, 128, smem>>>(map, x, y, plane_index, d);
  PDS_CUDA_OK(cudaGetLastError());
  PDS_CUDA_OK(cudaDeviceSynchronize());
  PDS_CUDA_OK(cudaMemcpy(out_host, d, kPlaneBytes, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}
