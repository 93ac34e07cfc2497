// All 64->64 body layers of one DnCNN pass in ONE persistent launch (small launches: single images, a few small images).
//
// Reference: models/basic_models.py:31-35 — `for i in range(depth-2): out = nl_list[i+1](conv_list[i](out))`; the
// reference (and the per-layer kernels of dncnn_tc.cu) serialise the 18 layers.  For a 256 x 256 image a layer is only
// 256 CTA-pair tiles on 74 CTA pairs: 3.46 rounds that a per-layer launch rounds up to 4, plus a prologue per layer
// (barrier init, TMEM allocation, 72 KB of weights per CTA) that cannot overlap the previous layer's tail because one CTA
// fills an SM's shared memory.  Here the (layer, tile pair) units of the whole chain form one sequence,
//     unit u = layer * npairs + pair,        cluster c works on u = c, c + nclusters, c + 2 nclusters, ...
// so every cluster gets the same number of units (+-1) over the chain, and a unit starts as soon as the 3 x 3 neighbourhood
// of tile pairs it reads has been written by the previous layer — tile-level dataflow through per-unit counters in global
// memory instead of a grid-wide barrier between layers:
//   * epilogue warps: after the unit's activation stores, __syncwarp + one CTA-scope release arrive on a shared-memory
//     mbarrier; a dedicated publisher warp per CTA waits on it and does the GPU-scope `red.release.gpu` (+4) on
//     flag[layer][pair], so the MEMBAR.ALL.GPU that waits for the stores' acknowledgements is off the epilogue's critical
//     path (first version, release in the epilogue warps: ncu stall_membar = 20 % of all samples, tensor pipe 33 % busy);
//   * TMA producer warp: before the unit's loads, up to nine lanes poll (`ld.acquire.gpu`) the flags of the neighbouring
//     pairs of the previous layer, then `fence.proxy.async` (the tiles were written through the generic proxy by other SMs
//     and are read through the async proxy) and the two plane loads.
//   Deadlock freedom: units are processed in increasing u by every cluster and a unit only depends on smaller u, so the
//   smallest unfinished unit can always run (all CTAs are co-resident: grid <= #SMs, one CTA per SM).  Write-after-read on
//   the two ping-pong activation buffers is covered by the same flags: (l+1, p) overwrites what the neighbours of p read in
//   layer l, and it waits for exactly those units.  Every wait is bounded by wall time and traps instead of hanging.
//   * flags count up across launches (target = 8 * epoch), so nothing is reset between denoiser calls.
// Weights: the layer changes every ~3.5 units, so the per-CTA half of the weight image is double buffered (2 x 72 KB) and
// the next layer's image is fetched while the current layer runs (3-slot activation-plane ring instead of 5).  MMA order per
// output element is that of conv_tc2_kernel -> bit-identical results (tests/test_gpu_dncnn.py).
#include "tc_common.cuh"

namespace pds {

namespace {
namespace chain {
using namespace two;

constexpr int kSlotsC = 3;
constexpr int kEpiWarpsC = 8;                                         // two warps per TMEM lane quarter, 32 of the 64 channels each
constexpr int kPubWarpC = 2 + kEpiWarpsC;
constexpr int kThreadsC = 32 * (kPubWarpC + 1);                       // producer, MMA / weight forwarder, 8 epilogue warps, publisher
constexpr int kMaxLayersC = kChainMaxLayers;
constexpr uint32_t kOffAC = 2 * kWHalf;                               // after the two weight buffers
constexpr uint32_t kOffBarC = kOffAC + kSlotsC * kPlaneSlot;
constexpr uint32_t kOffBiasC = kOffBarC + 256;                        // bias[nlayers][64] + lo_scale[nlayers] of the whole chain
constexpr uint32_t kSmemBytesC = kOffBiasC + kMaxLayersC * (256 + 4) + 1024;   // + slack for the manual 1024-B alignment
constexpr int kAccStagesC = 4;
constexpr int kDoneRing = 8;
constexpr int kTraceClusters = 4, kTraceUnits = 64;
// event slots of the debug timeline
enum { TR_POLLED = 0, TR_TMA = 1, TR_TEMPTY = 2, TR_FULL0 = 3, TR_ISSUED = 4, TR_TFULL = 5, TR_STORED = 6, TR_PUBLISHED = 7 };
static_assert(kSmemBytesC <= 227 * 1024, "chain kernel shared memory");

struct ChainArgs {
  const ChainLayer* layers;   // device table [nlayers]
  __half* act0;
  __half* act1;
  uint32_t* flags;            // [nlayers][flag_stride]
  uint32_t target;            // 8 * epoch of this launch
  int flag_stride;
  int interleave;             // MMA issue order: 0 = plane 0 then plane 1, 1 = the two kinds alternating (both planes first)
  unsigned long long* trace;  // debug (pds_debug_chain_trace): [kTraceClusters][kTraceUnits][8] globaltimer stamps of CTA 0 of the first clusters
  int in_buf;                 // buffer read by chain layer 0
  int nlayers;
  float slope;
  int H, W, nimg, tiles_x, tiles_y, ntiles, npairs;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// global state space only (FENCE.VIEW.ASYNC.G).  The unqualified form is MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC.S: it also waits for
// the TMA writes into shared memory that the producer has in flight — measured at 1.1 us per unit on the producer's critical path.
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void trace_stamp(const ChainArgs& a, uint32_t rank, int cid, int k, int ev) {
  if (a.trace != nullptr && rank == 0 && cid < kTraceClusters && k < kTraceUnits)
    a.trace[((size_t)cid * kTraceUnits + k) * 8 + ev] = (unsigned long long)clock64();   // SM cycle counter: all stamps come from CTA 0's SM
}

// first layer > l among the units u = cid + k * nclusters of this cluster (-1: none)
__device__ __forceinline__ int next_layer_of_cluster(int l, int cid, int nclusters, int npairs, int nunits) {
  const int base = (l + 1) * npairs;
  int k = (base - cid + nclusters - 1) / nclusters;
  if (k < 0) k = 0;
  const int u = cid + k * nclusters;
  return u < nunits ? u / npairs : -1;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsC, 1)
    conv_chain_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1, ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sW = base, sA = base + kOffAC, sBar = base + kOffBarC;
  // barriers: full[3] @0 (CTA 0), empty[3] @24, tfull[4] @48, tempty[4] @80 (CTA 0), wfull[2] @112, wpeer[2] @128 (CTA 0),
  // wempty[2] @144, tmem slot @160, done[8] @168 (epilogue -> publisher), publisher progress counter @232
  const uint32_t bFull = sBar, bEmpty = sBar + 24, bTFull = sBar + 48, bTEmpty = sBar + 80;
  const uint32_t bWFull = sBar + 112, bWPeer = sBar + 128, bWEmpty = sBar + 144, sTmemSlot = sBar + 160;
  const uint32_t bDone = sBar + 168, sPubCount = sBar + 232;
  float* bias_all = reinterpret_cast<float*>(gbase + kOffBiasC);             // [nlayers][64]
  float* lo_all = bias_all + a.nlayers * 64;                                 // [nlayers]
  const uint32_t rank = cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlotsC; ++i) {
      mbar_init(bFull + 8 * i, 1);          // CTA 0: one arrive.expect_tx for both CTAs' boxes
      mbar_init(bEmpty + 8 * i, 1);         // multicast commit from CTA 0
    }
    for (int i = 0; i < kAccStagesC; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 2 * kEpiWarpsC);   // the epilogue warps of both CTAs arrive on CTA 0's barrier
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bWFull + 8 * i, 1);         // own half of a layer's weight image landed
      mbar_init(bWPeer + 8 * i, 1);         // CTA 0: the peer's half landed too
      mbar_init(bWEmpty + 8 * i, 1);        // every MMA that read this weight buffer has completed (multicast commit)
    }
    for (int i = 0; i < kDoneRing; ++i) mbar_init(bDone + 8 * i, kEpiWarpsC);   // this CTA's epilogue warps stored their part of a unit
    sts_volatile(sPubCount, 0u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap1) : "memory");
  }
  // biases / correction scales of every layer of the chain (constants: no dependence on the previous grid)
  for (int i = threadIdx.x; i < a.nlayers * 64; i += kThreadsC) bias_all[i] = __ldg(a.layers[i >> 6].bias + (i & 63));
  for (int i = threadIdx.x; i < a.nlayers; i += kThreadsC) lo_all[i] = a.layers[i].lo_scale;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemSlot), "r"(kTmemCols2) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // all barriers of both CTAs are initialised, TMEM slot written
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kOffBarC + 160);
  pdl_launch_dependents();

  const int per_img = a.tiles_x * a.tiles_y;
  const int npairs = a.npairs;
  const int nunits = a.nlayers * npairs;
  const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ producer: dependencies, activation planes, weights
    pdl_wait_prior_grid();
    uint32_t j = 0;
    int gen = -1, issued = -1, prev_l = -1, kk = 0;
    for (int u = cid; u < nunits; u += nclusters, ++kk) {
      const int l = u / npairs, p = u - l * npairs;
      const bool first_of_gen = l != prev_l;
      if (first_of_gen) ++gen;
      prev_l = l;
      int tile = 2 * p + (int)rank;
      if (tile >= a.ntiles) tile = a.ntiles - 1;            // odd tail: load a valid tile, its result is not stored
      const int img = tile / per_img, rem = tile - img * per_img;
      const int tyi = rem / a.tiles_x, txi = rem - tyi * a.tiles_x;
      // 1. the 3 x 3 neighbourhood of this tile in the previous layer's output
      if (l > 0) {
        if (lane < 9) {
          const int ny = tyi + lane / 3 - 1, nx = txi + lane % 3 - 1;
          if (ny >= 0 && ny < a.tiles_y && nx >= 0 && nx < a.tiles_x) {
            const int nt = img * per_img + ny * a.tiles_x + nx;
            const uint32_t* f = a.flags + (size_t)(l - 1) * a.flag_stride + (nt >> 1);
            unsigned long long t0 = 0;
            for (uint32_t spin = 0; (int32_t)(ld_acquire_gpu(f) - a.target) < 0; ++spin) {
              if ((spin & 255u) == 255u) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kWaitTrapNs) __trap();
              }
            }
          }
        }
        __syncwarp();
        fence_proxy_async_global();         // generic-proxy writes of other SMs (acquired above) -> async-proxy TMA reads below
      }
      if (lane == 0) trace_stamp(a, rank, cid, kk, TR_POLLED);
      // 2. first plane slot of the unit (all MMAs of unit k-2 have completed once this passes)
      mbar_wait(bEmpty + 8 * (j % kSlotsC), ((j / kSlotsC) & 1) ^ 1);
      // 3. weights: this generation must be on its way; from the second unit of a generation on, fetch the next layer's image
      //    into the other buffer (its previous user, generation gen-1, has completed by then: the wait below is immediate)
      {
        const int next_l = next_layer_of_cluster(l, cid, nclusters, npairs, nunits);
        const int want = (!first_of_gen && next_l >= 0) ? gen + 1 : gen;
        while (issued < want) {
          const int G = issued + 1, b = G & 1;
          const int lw = (G == gen) ? l : next_l;
          if (G >= 2) mbar_wait(bWEmpty + 8 * b, (uint32_t)(((G >> 1) & 1) ^ 1));
          if (elect_one()) {
            mbar_expect_tx(bWFull + 8 * b, kWHalf);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(a.layers[lw].w) + (size_t)rank * kWHalf;
            for (int i = 0; i < 9; ++i) bulk_load(sW + b * kWHalf + i * 8192u, src + (size_t)i * 8192u, 8192u, bWFull + 8 * b);
          }
          __syncwarp();
          issued = G;
        }
      }
      // 4. the two activation planes of the unit
      const CUtensorMap* tm = ((a.in_buf + l) & 1) ? &tmap1 : &tmap0;
      const int y0 = tyi * kTileRows, x0 = txi * kTileCols;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl, ++j) {
        const uint32_t slot = j % kSlotsC, use = j / kSlotsC;
        if (pl == 1) mbar_wait(bEmpty + 8 * slot, (use & 1) ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(bFull + 8 * slot, 2 * kPlaneBytes);
          tma_load_4d_2sm(sA + slot * kPlaneSlot, tm, map_to_cta(bFull + 8 * slot, 0), 0, x0 - 1, y0 - 1, img * 2 + pl);
        }
        __syncwarp();
      }
      if (lane == 0) trace_stamp(a, rank, cid, kk, TR_TMA);
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ------------------------------------------------------------ MMA issuer (CTA 0 only)
      uint32_t j = 0;
      int it = 0, gen = -1, prev_l = -1;
      uint32_t wbuf = 0;
      for (int u = cid; u < nunits; u += nclusters, ++it) {
        const int l = u / npairs;
        if (l != prev_l) {
          ++gen;
          wbuf = (uint32_t)(gen & 1);
          const uint32_t par = (uint32_t)((gen >> 1) & 1);
          mbar_wait(bWFull + 8 * wbuf, par);                 // own half ...
          mbar_wait(bWPeer + 8 * wbuf, par);                 // ... and the peer's
          prev_l = l;
        }
        const int un = u + nclusters;
        const bool last_of_gen = un >= nunits || un / npairs != l;
        const uint32_t w_lo = (((sW + wbuf * kWHalf) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t acc = it % kAccStagesC;
        mbar_wait(bTEmpty + 8 * acc, (uint32_t)(((it / kAccStagesC) & 1) ^ 1));
        if (lane == 0) trace_stamp(a, rank, cid, it, TR_TEMPTY);
        const uint32_t d_tmem = tmem_base + acc * kAccCols2;
        if (a.interleave == 1) {
          // both planes first, then the two MMA kinds alternating (independent accumulators)
          const uint32_t s0 = j % kSlotsC, u0 = j / kSlotsC, s1 = (j + 1) % kSlotsC, u1 = (j + 1) / kSlotsC;
          mbar_wait(bFull + 8 * s0, u0 & 1);
          if (lane == 0) trace_stamp(a, rank, cid, it, TR_FULL0);
          mbar_wait(bFull + 8 * s1, u1 & 1);
          tc_fence_after();
          const uint32_t a0_lo = (((sA + s0 * kPlaneSlot) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t a1_lo = (((sA + s1 * kPlaneSlot) & 0x3FFFFu) >> 4) | (1u << 16);
          if (elect_one()) {
            issue_unit2_interleaved(d_tmem, a0_lo, a1_lo, w_lo);
            umma_commit_2sm(bEmpty + 8 * s0);
            umma_commit_2sm(bEmpty + 8 * s1);
            umma_commit_2sm(bTFull + 8 * acc);
            if (last_of_gen) umma_commit_2sm(bWEmpty + 8 * wbuf);
          }
          __syncwarp();
          j += 2;
        } else {
#pragma unroll
          for (int pl = 0; pl < 2; ++pl, ++j) {
            const uint32_t slot = j % kSlotsC, use = j / kSlotsC;
            mbar_wait(bFull + 8 * slot, use & 1);
            tc_fence_after();
            if (pl == 0 && lane == 0) trace_stamp(a, rank, cid, it, TR_FULL0);
            const uint32_t a_lo = (((sA + slot * kPlaneSlot) & 0x3FFFFu) >> 4) | (1u << 16);
            if (elect_one()) {
              if (a.interleave == 2) {            // timing probe (wrong results): A from the collector, see tc_common.cuh
                if (pl == 0) issue_plane2_probe_collector<true>(d_tmem, a_lo, w_lo);
                else issue_plane2_probe_collector<false>(d_tmem, a_lo, w_lo);
              } else if (pl == 0) issue_plane2<true>(d_tmem, a_lo, w_lo);
              else issue_plane2<false>(d_tmem, a_lo, w_lo);
              umma_commit_2sm(bEmpty + 8 * slot);
              if (pl == 1) {
                umma_commit_2sm(bTFull + 8 * acc);
                if (last_of_gen) umma_commit_2sm(bWEmpty + 8 * wbuf);
              }
            }
            __syncwarp();
          }
        }
        if (lane == 0) trace_stamp(a, rank, cid, it, TR_ISSUED);
      }
    } else {
      // ------------------------------------------------------------ CTA 1: tell the MMA issuer when this half of a weight image landed
      int gen = -1, prev_l = -1;
      for (int u = cid; u < nunits; u += nclusters) {
        const int l = u / npairs;
        if (l == prev_l) continue;
        prev_l = l;
        ++gen;
        const uint32_t wbuf = (uint32_t)(gen & 1);
        mbar_wait(bWFull + 8 * wbuf, (uint32_t)((gen >> 1) & 1));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_cta(bWPeer + 8 * wbuf, 0));
      }
    }
  } else if (warp == kPubWarpC) {
    // ------------------------------------------------------------ publisher: GPU-scope release of finished units
    int it = 0;
    for (int u = cid; u < nunits; u += nclusters, ++it) {
      const int l = u / npairs, p = u - l * npairs;
      mbar_wait(bDone + 8 * (it & (kDoneRing - 1)), (uint32_t)((it / kDoneRing) & 1));   // acquire.cta: the epilogue warps' stores
      if (lane == 0) {
        red_release_gpu_add(a.flags + (size_t)l * a.flag_stride + p, 4u);                // cumulative: release.gpu over what was acquired
        sts_volatile(sPubCount, (uint32_t)(it + 1));
        trace_stamp(a, rank, cid, it, TR_PUBLISHED);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (each CTA drains its own 128 TMEM lanes)
    // The steady state of the chain has no per-layer tail to hide a slow epilogue in (a per-layer launch runs its <= 4 units
    // into the 4 TMEM stages and drains them while the next launch sets up): with four warps the epilogue took 2.2 us per unit
    // against 1.6 us of MMAs (ncu: epilogue warps 80 % busy, tensor pipe 43 %).  Eight warps: each drains 32 of the 64 channels.
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;         // channels [32 half, 32 half + 32)
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const size_t hw = (size_t)a.H * a.W;
    const uint32_t tempty0 = map_to_cta(bTEmpty, 0);
    int it = 0;
    for (int u = cid; u < nunits; u += nclusters, ++it) {
      const int l = u / npairs, p = u - l * npairs;
      const int tile = 2 * p + (int)rank;
      const bool live = tile < a.ntiles;
      const int tl = live ? tile : a.ntiles - 1;
      const int img = tl / per_img, rem = tl - img * per_img;
      const int y = (rem / a.tiles_x) * kTileRows + ty, x = (rem % a.tiles_x) * kTileCols + tx;
      const float* bias = bias_all + l * 64;
      const float lo_scale = lo_all[l];
      __half* outb = ((a.in_buf + l + 1) & 1) ? a.act1 : a.act0;
      const uint32_t acc = it % kAccStagesC;
      mbar_wait(bTFull + 8 * acc, (uint32_t)((it / kAccStagesC) & 1));
      tc_fence_after();
      if (warp == 2 && lane == 0) trace_stamp(a, rank, cid, it, TR_TFULL);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols2;
      const bool st = live && y < a.H && x < a.W;
      const size_t pix = (size_t)y * a.W + x;
      __half* o_p0 = outb + (((size_t)img * 2 + 0) * hw + pix) * 64;
      uint8_t* o_p1 = reinterpret_cast<uint8_t*>(outb + (((size_t)img * 2 + 1) * hw + pix) * 64);
      uint32_t r0[32], r2[32];
      tmem_ld32(taddr + 32 * half, r0);       // a_hi * w_hi accumulator
      tmem_ld32(taddr + 64 + 32 * half, r2);  // e4m3 correction accumulator
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);      // stage released before any arithmetic or store
      if (st) store_half_row(o_p0, o_p1, r0, r2, bias, 32 * half, a.slope, lo_scale, 1);
      __syncwarp();
      if (warp == 2 && lane == 0) trace_stamp(a, rank, cid, it, TR_STORED);
      if (lane == 0) {
        // ring-slot reuse: the publisher has consumed this slot's previous phase (it never lags 8 units in practice)
        while (it >= kDoneRing && (int)lds_volatile(sPubCount) < it - (kDoneRing - 1)) {}
        mbar_arrive(bDone + 8 * (it & (kDoneRing - 1)));     // release.cta, cumulative over the warp's stores (after __syncwarp)
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

}  // namespace chain
}  // namespace

int chain_setup() {
  cudaError_t e = cudaFuncSetAttribute(chain::conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain::kSmemBytesC);
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(conv_chain_kernel): ") + cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// Device tables of the chain kernel: one ChainLayer per body layer, and the per-unit flags (zeroed once).
int tc_plan_set_chain(TcPlan* plan, const std::vector<ChainLayer>& layers) {
  const int tiles = ((plan->W + kTileCols - 1) / kTileCols) * ((plan->H + kTileRows - 1) / kTileRows) * plan->nimg;
  const int npairs = (tiles + 1) / 2;
  if (layers.empty() || npairs > kChainMaxPairs || (int)layers.size() > kChainMaxLayers) return 0;   // -> per-layer kernels
  PDS_CUDA_OK(cudaMalloc(&plan->chain_layers, layers.size() * sizeof(ChainLayer)));
  PDS_CUDA_OK(cudaMemcpy(plan->chain_layers, layers.data(), layers.size() * sizeof(ChainLayer), cudaMemcpyHostToDevice));
  const size_t nflags = layers.size() * (size_t)npairs;
  PDS_CUDA_OK(cudaMalloc(&plan->chain_flags, nflags * sizeof(uint32_t)));
  PDS_CUDA_OK(cudaMemset(plan->chain_flags, 0, nflags * sizeof(uint32_t)));
  plan->chain_nlayers = (int)layers.size();
  plan->chain_stride = npairs;
  plan->chain_epoch = 0;
  plan->chain_npairs_last = -1;
  plan->chain_bytes = layers.size() * sizeof(ChainLayer) + nflags * sizeof(uint32_t);
  return 0;
}

bool chain_available(const TcPlan* plan, int nimg) {
  if (!plan || !plan->chain_flags) return false;
  const int tiles = ((plan->W + kTileCols - 1) / kTileCols) * ((plan->H + kTileRows - 1) / kTileRows) * nimg;
  return (tiles + 1) / 2 <= plan->chain_stride;
}

cudaError_t launch_conv_chain(TcPlan* plan, int in_buf, int nimg, float slope, int interleave, cudaStream_t st) {
  chain::ChainArgs a{};
  a.interleave = interleave;
  a.layers = plan->chain_layers;
  a.act0 = plan->act[0];
  a.act1 = plan->act[1];
  a.flags = plan->chain_flags;
  a.flag_stride = plan->chain_stride;
  a.trace = plan->chain_trace;
  a.in_buf = in_buf;
  a.nlayers = plan->chain_nlayers;
  a.slope = slope;
  a.H = plan->H;
  a.W = plan->W;
  a.nimg = nimg;
  a.tiles_x = (plan->W + kTileCols - 1) / kTileCols;
  a.tiles_y = (plan->H + kTileRows - 1) / kTileRows;
  a.ntiles = a.tiles_x * a.tiles_y * nimg;
  a.npairs = (a.ntiles + 1) / 2;
  // the flags of the units this launch runs must all stand at 8 * epoch: a launch over a different number of pairs than the
  // previous one (last partial chunk of a batch), or an epoch counter close to wrapping, starts from zeroed flags
  if (a.npairs != plan->chain_npairs_last || plan->chain_epoch >= (1u << 27)) {
    cudaError_t e = cudaMemsetAsync(plan->chain_flags, 0, (size_t)plan->chain_nlayers * plan->chain_stride * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    plan->chain_epoch = 0;
    plan->chain_npairs_last = a.npairs;
  }
  a.target = 8u * ++plan->chain_epoch;
  const int nunits = a.nlayers * a.npairs;
  const int half = plan->num_sms / 2;
  const int nclusters = nunits < half ? nunits : half;
  return launch_pdl(chain::conv_chain_kernel, 2 * nclusters, chain::kThreadsC, chain::kSmemBytesC, st, plan->map[0], plan->map[1], a);
}

// Debug timeline (pds_debug_chain_trace): arm -> the next chain launches stamp the SM cycle counter at the pipeline events of the first
// units of the first clusters; read -> copies the stamps back and disarms.
int tc_chain_trace(TcPlan* plan, unsigned long long* out_host) {
  const size_t n = (size_t)chain::kTraceClusters * chain::kTraceUnits * 8;
  if (out_host == nullptr) {
    if (!plan->chain_trace) PDS_CUDA_OK(cudaMalloc(&plan->chain_trace, n * sizeof(unsigned long long)));
    PDS_CUDA_OK(cudaMemset(plan->chain_trace, 0, n * sizeof(unsigned long long)));
    return 0;
  }
  PDS_REQUIRE(plan->chain_trace != nullptr, "chain trace is not armed");
  PDS_CUDA_OK(cudaDeviceSynchronize());
  PDS_CUDA_OK(cudaMemcpy(out_host, plan->chain_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(plan->chain_trace);
  plan->chain_trace = nullptr;
  return 0;
}

// ---- interface used by pds_api.cu (kernels.cuh) ----
int tc_plan_chain(TcPlan* plan, const DncnnLayerW* layers, int depth, size_t* bytes_out) {
  std::vector<ChainLayer> v;
  for (int l = 1; l < depth - 1; ++l) v.push_back(ChainLayer{layers[l].w_mid_tc2, layers[l].bias, layers[l].lo_scale, 0.f});
  const int rc = tc_plan_set_chain(plan, v);
  if (bytes_out) *bytes_out = plan->chain_bytes;
  return rc;
}
bool tc_chain_available(const TcPlan* plan, int nimg) { return chain_available(plan, nimg); }
cudaError_t launch_conv_body_chain(TcPlan* plan, int in_buf, int nimg, float slope, int interleave, cudaStream_t st) {
  return launch_conv_chain(plan, in_buf, nimg, slope, interleave, st);
}

}  // namespace pds
