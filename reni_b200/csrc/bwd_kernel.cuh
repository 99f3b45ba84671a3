// Fused RENI decoder backward (delta chain) for sm_100a.
//
//   per 128-direction tile:
//     g_y   = dLoss/dy                      fused loss (sine-weighted MSE + cosine term, loss_functions.py:6-32,60-71)
//                                           or an external grad_out, times tanh' = 1 - o^2; scaled by S into fp16 range
//     d_L   = (g_y  W_out'') * cos(a_L)     tcgen05.mma K = 16
//     d_l-1 = (d_l  W_l'')   * cos(a_l-1)   tcgen05.mma 128 x 256 x 256, l = L..1; cos rebuilt (MUFU) from the forward's
//                                           phase stash (phase.cuh); omega of the consuming layer folded into W''
//     dM_b, dc_b += [f | 1]^T d_0           tcgen05.mma N = 16 with d_0 read as an MN-major operand (contraction over the
//                                           tile's rows): the per-map layer-0 reduction costs 16 tiny MMAs per tile
//   d_l tiles stay in shared memory for the next GEMM; when weight gradients are wanted d_1..d_L are also stashed as
//   fp16 tile images for the weight-gradient GEMM.  d_0 never leaves the SM.
//
// CTA pairs (cluster of 2, tcgen05.mma.cta_group::2, M = 256) like the forward kernel: the leader issues every MMA for
// sub-tile g of BOTH CTAs, each CTA streams only its 128-column half of every backward weight image (TMA tile loads
// that complete on the leader's mbarrier) and holds its own accumulators.  The per-tile reductions ride along: the
// K = 16 head GEMM splits W_out'' by columns, and the layer-0 reduction (MN-major operands) gives each CTA 8 of the
// 16 accumulator columns for its own [f | 1] features.
// Per CTA: producer warp, MMA warp (leader only), two ping-ponging epilogue groups of 8 warps.
// Each epilogue thread owns (row, 128 columns): the 16 x 16-B stash loads of its whole layer slice are issued before it
// waits for the GEMM, so their HBM/L2 latency hides under the tensor pipe; the producer warp additionally prefetches the
// next layer's stash tiles into L2 (cp.async.bulk.prefetch.L2).
#pragma once
#include <cuda.h>  // CUtensorMap

#include "layout.cuh"
#include "phase.cuh"
#include "ptx.cuh"

namespace reni {

constexpr int kBwdThreads = 576;
constexpr int kBwdStages = 4;
#ifndef RENI_LAYER_RESIDENT
#define RENI_LAYER_RESIDENT 0  // see fwd_kernel.cuh: measured slower, kept as a build switch
#endif
constexpr bool kBwdLayerResident = RENI_LAYER_RESIDENT != 0;

struct BwdParams {
  const float* out;       // (B, P, 3) forward output (after tanh)
  const float* grad_out;  // (B, P, 3) external gradient, or null for the fused loss
  const float* aout;      // sine output layer (RENI.py:164-171): (B, P, 3) pre-activations a_out, else null
  const float* target;    // fused loss
  const float* sw;
  int64_t sw_bstride;
  const float* map_loss;  // (B, 32): [16..18] = coefA_c, [19..21] = coefB_c (cosine term, pre-scaled by S)
  const float* scalars;   // [0] = S (gradient scale)
  const __half* wb;       // L backward weight images [j/8][k][8] (unpaired mode)
  const __half* wb2;      // ... split for CTA pairs: [l][k half][j/8][128][8]
  const __half* w6b;      // [2][256][8]
  const uint16_t* stash_u;  // phases of a_l (12 or 16 bits each, phase.cuh), per tile (L+1) tile-layer images
  __half* stash_d;        // delta stash: per tile nslots images (nslots = L+1 or 1)
  __half* stash_gy;       // per tile [2 halves][2][64][8]
  const float* D;         // directions (for the layer-0 feature columns f)
  int64_t d_bstride;
  float* dmc;             // (B, 5, 256): dM_b rows 0..3, dc_b row 4; accumulated with atomics (caller zeroes)
  const float* film;      // kFilm: (B, L, 2, 256) per-map (freq_l, phase_l) of the hidden layers
  int B, P, tiles_per_map, ntiles, L;
  int out_tanh, d_slots, so2;
  int use_cos;            // fused loss: 0 = no cosine term (map_loss is not read, it may still be in flight)
  int grid_w, dir_grid, sw_grid;  // analytic grid (RENI_FLAG_GRID_DIRECTIONS / RENI_FLAG_GRID_SINEWEIGHT), see fwd_kernel.cuh
  const uint32_t* mask_bits;
  int w_map_rows;         // per-map backward images (FiLM on per-map images): rows of 256 B per map in wmap, 0 = shared
  uint32_t* ready;        // overlap mode (weight-gradient kernel co-resident on the other SMs): per-tile counter, +1 per
                          // epilogue warp each time a stashed delta_l of the tile is complete in global memory
  unsigned long long* trace;  // debug builds (-DRENI_BWD_TRACE=1): clock64 timeline of CTA 0, see tools/trace_bwd.py
  alignas(64) CUtensorMap wmap;  // wb2 as rows of 256 B, box = one 16 KB half chunk
};

#ifndef RENI_BWD_TRACE
#define RENI_BWD_TRACE 0
#endif
DEVINL void btrace_ev(const BwdParams& p, int role, uint32_t& n, uint32_t code) {
  if (RENI_BWD_TRACE && p.trace != nullptr && blockIdx.x == 0 && n < 4096) {
    p.trace[role * 4096 + n] = ((unsigned long long)code << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    ++n;
  }
}

struct BwdSmem {
  static constexpr int kA = 0;
  static constexpr int kRing = kA + 2 * kTileImageBytes;
  static constexpr int kW6 = kRing + kBwdStages * kWChunkBytes;
  static constexpr int kF = kW6 + kW6ImageBytes;               // 2 x [2][128][8] fp16 feature tiles [f0..f3, 1, 0..]
  static constexpr int kBars = kF + 2 * kGyImageBytes;
  static constexpr int kNumBars = 2 * kBwdStages + 6;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemPtr + 16;
};
static_assert(BwdSmem::kTotal <= 232448, "backward kernel shared memory over budget");

#ifndef RENI_BWD_L2_HINTS
#define RENI_BWD_L2_HINTS 0  // bit 0: evict_first on the delta-stash stores, bit 1: evict_last on the phase prefetch
#endif
// L2 cache policies: the delta stash is a pure stream for this kernel (read back only by the next kernel), the phase
// tiles pulled in by the prefetch must survive until the epilogue reads them
DEVINL uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
DEVINL uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
DEVINL void bulk_prefetch_l2(const void* gmem, uint32_t bytes) {
#if RENI_BWD_L2_HINTS & 2
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(gmem), "r"(bytes),
               "l"(l2_policy_evict_last())
               : "memory");
#else
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
#endif
}

// bulk async copy shared -> global (delta-stash store), tracked by the issuing thread's bulk async-group
DEVINL void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
#if RENI_BWD_L2_HINTS & 1
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(l2_policy_evict_first())
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
#endif
}
DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DEVINL void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// gpu-scope release of everything this warp has written (call after __syncwarp, one lane): the co-resident
// weight-gradient CTAs acquire the counter before they pull the tile's stash blocks
DEVINL void ready_signal(uint32_t* ctr) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}

#ifndef RENI_BWD_PHASE_HINT
#define RENI_BWD_PHASE_HINT 1  // phase-stash loads: 0 ld.global.nc, 1 ld.global.cs (streaming), 2 ld.global.lu (293 -> 289..291 us)
#endif
#ifndef RENI_BWD_PF_DIST
#define RENI_BWD_PF_DIST 1
#endif
#ifndef RENI_BWD_PF_DIST_LATENT
#define RENI_BWD_PF_DIST_LATENT 1  // the same for the latent-only chain (no delta-stash writes competing for bandwidth)
#endif
#ifndef RENI_BWD_PF_NEXT_UNIT
#define RENI_BWD_PF_NEXT_UNIT 1  // L2-prefetch the next unit's top-layer phase tiles during the current unit's last pass:
                                 // 0 off, 1 latent-only chain (225 -> 213 us at cfg 2), 2 also with weight gradients
                                 // (293 -> 307 us: that kernel is short of bandwidth, early requests hurt it); also
                                 // prefetching the unit's out / target rows changed nothing
#endif
#ifndef RENI_BWD_SIGNAL_FIRST
#define RENI_BWD_SIGNAL_FIRST 1  // sub-tile handed to the MMA issuer before (1) / after (0) its stash copies are queued.
                                 // 1 removes a ~3900 clk wait from the issuer's timeline.  With the 16-bit phase stash
                                 // the kernel got SLOWER for it (295 -> 304 us: the layer step was bound by its 256 KB of
                                 // HBM traffic per SM, tools/ubench/sm_traffic.cu); with 12-bit phases it is no longer
                                 // purely HBM-bound and gains (264 -> 260 us)
#endif
#ifndef RENI_BWD_BULK_STASH
#define RENI_BWD_BULK_STASH 1  // 1: delta stash written by per-warp bulk copies of the finished smem pieces; 0: st.global
#endif

// kFilm (FiLM conditioning, RENI.py:515-524): a_l = freq_l[b] * u_l + phase_l[b] with u_l = W_l h_{l-1} + b_l, so
// dL/du_l = delta_l * freq_l[b].  The epilogue of hidden layer l keeps TWO values per element: the unscaled
// delta_l = dL/da_l goes to the stash (the weight-gradient kernel turns it into dW_l, dfreq_l, dphase_l per map) and
// delta_l * freq_l[b] goes to the shared-memory tile that feeds the next GEMM.  The stash is therefore written with
// st.global (the bulk copies of the non-FiLM kernel move the shared-memory tile as it is).
template <bool kNeedDW, bool kPair, bool kFilm = false>
__global__ void __launch_bounds__(kBwdThreads, 1) reni_bwd_kernel(const __grid_constant__ BwdParams p) {
  static_assert(!kFilm || kNeedDW, "the FiLM backward always stashes delta (dfreq / dphase come from the dW kernel)");
  constexpr bool kBulk = RENI_BWD_BULK_STASH && !kFilm;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBars);
  uint64_t* w_full = bars;                        // [kBwdStages] leader: both halves of a chunk have landed
  uint64_t* w_empty = bars + kBwdStages;          // [kBwdStages] multicast commit: the slot is free in both CTAs
  uint64_t* a_ready = bars + 2 * kBwdStages;      // [2] this CTA's epilogue group g -> (leader's) MMA issuer
  uint64_t* acc_full = a_ready + 2;               // [2] multicast commit -> epilogue group g of both CTAs
  uint64_t* a_ready_peer = acc_full + 2;          // [2] leader only: the peer's group g (one arrival per warp)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + BwdSmem::kTmemPtr);

  const int L = p.L;
  // Work units are taken in DESCENDING tile order (the forward kernel's last tiles are the freshest in L2).
  // kPair (cluster of 2): cluster k owns tile quads; leader: tiles {4q, 4q+1}, peer: {4q+2, 4q+3}.
  // Unpaired (compile-time fallback): CTA c owns tile pairs.
  const uint32_t crank = kPair ? cluster_ctarank() : 0;
  constexpr int kUnitTiles = kPair ? 4 : 2;
  const int nunits = (p.ntiles + kUnitTiles - 1) / kUnitTiles;
  const int nworkers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int iters = (nunits - worker + nworkers - 1) / nworkers;
  auto clamp02 = [](int x) { return x < 0 ? 0 : (x > 2 ? 2 : x); };
  auto unit_base = [&](int it) { return (nunits - 1 - (worker + it * nworkers)) * kUnitTiles; };  // first tile of the unit
  // per layer pass: paired 4 x [8 j-groups][128 k][8] (K 64 of this CTA's N half), unpaired 8 x [4 j-groups][256 k][8]
  constexpr int kChunks = kPair ? 4 : 8;
  constexpr int kSteps = kPair ? 4 : 2;          // K = 16 steps per chunk
  constexpr uint32_t kBRows = kPair ? 128 : 256;  // rows of B in this CTA's chunk

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(&a_ready[0], 256);
    mbar_init(&a_ready[1], 256);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&a_ready_peer[0], 8);
    mbar_init(&a_ready_peer[1], 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc2<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  {  // this CTA's 128 of the 256 columns of W_out'': [c/8 2][128 k][8]
    const uint4* src = reinterpret_cast<const uint4*>(p.w6b);
    uint4* dst = reinterpret_cast<uint4*>(smem + BwdSmem::kW6);
    if (kPair) {
      for (int i = threadIdx.x; i < kW6ImageBytes / 32; i += kBwdThreads)
        dst[i] = src[(i >> 7) * kH + crank * 128 + (i & 127)];
    } else {
      for (int i = threadIdx.x; i < kW6ImageBytes / 16; i += kBwdThreads) dst[i] = src[i];
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t ra_peer = kPair ? mapa_u32(smem_u32(a_ready_peer), 0) : 0;
  auto signal_ready = [&](int g) {  // "sub-tile g of this CTA is in shared memory"
    if (kPair && crank == 1) {
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(ra_peer + g * 8);
    } else {
      mbar_arrive(&a_ready[g]);
    }
  };

  if (warp == 0) {
    // ============================================================ weight-chunk producer (layers L..1, this CTA's half)
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int it = 0; it < iters; ++it) {
        const int ubase = unit_base(it);
        const int tbase = ubase + 2 * (int)crank;
        const int nsub = clamp02(p.ntiles - tbase);          // this CTA's live sub-tiles
        const int nstream = clamp02(p.ntiles - ubase);       // passes the leader makes over each layer
        const int32_t wrow0 = (ubase / p.tiles_per_map) * p.w_map_rows;  // (per-map images: one map per unit)
        // phase-stash tiles are pulled towards L2 kPfDist layers before the epilogue that multiplies by their cosine
        constexpr int kPfDist = kNeedDW ? RENI_BWD_PF_DIST : RENI_BWD_PF_DIST_LATENT;
        // (the first unit's top layer here; with kPfNext every later unit's top layer is requested during the previous
        // unit's last layer pass, so that its first epilogue does not start with an HBM miss)
        constexpr bool kPfNext = RENI_BWD_PF_NEXT_UNIT == 2 || (RENI_BWD_PF_NEXT_UNIT == 1 && !kNeedDW);
        if (it == 0 || !kPfNext)
          for (int g = 0; g < nsub; ++g)
            for (int d = 0; d < kPfDist && L - d >= 0; ++d)
              bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.stash_u) +
                                   ((size_t)(tbase + g) * (L + 1) + (L - d)) * kPhaseTileBytes,
                               kPhaseTileBytes);
        auto prefetch_next_unit = [&]() {
          if (!kPfNext || it + 1 >= iters) return;
          const int tnext = unit_base(it + 1) + 2 * (int)crank;
          const int nnext = clamp02(p.ntiles - tnext);
          for (int g = 0; g < nnext; ++g) {
            for (int d = 0; d < kPfDist && L - d >= 0; ++d)
              bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.stash_u) +
                                   ((size_t)(tnext + g) * (L + 1) + (L - d)) * kPhaseTileBytes,
                               kPhaseTileBytes);
          }
        };
        if (kPair && kBwdLayerResident) {
          // layer-resident weights (see the forward kernel): slot c = chunk c of the current layer for both passes
          for (int l = L; l >= 1; --l) {
            for (int g = 0; g < nsub; ++g)
              if (kPfDist > 0 && l - kPfDist >= 0)
                bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.stash_u) +
                                     ((size_t)(tbase + g) * (L + 1) + (l - kPfDist)) * kPhaseTileBytes,
                                 kPhaseTileBytes);
            if (l == 1) prefetch_next_unit();
            for (int c = 0; c < kChunks; ++c) {
              mbar_wait(&w_empty[c], ph ^ 1);
              if (crank == 0) mbar_arrive_expect_tx(&w_full[c], 2 * kWChunkBytes);
              const int32_t row =
                  (int32_t)(((size_t)((l - 1) * 2 + crank) * (kWImageBytes / 2) + (size_t)c * kWChunkBytes) / 256);
              tma2_load_2d(smem + BwdSmem::kRing + c * kWChunkBytes, &p.wmap, 0, wrow0 + row, mapa_u32(smem_u32(&w_full[c]), 0));
            }
            ph ^= 1;
          }
        } else {
        for (int l = L; l >= 1; --l) {
          for (int g = 0; g < nstream; ++g) {
            if (kPfDist > 0 && g < nsub && l - kPfDist >= 0)
              bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.stash_u) +
                                   ((size_t)(tbase + g) * (L + 1) + (l - kPfDist)) * kPhaseTileBytes,
                               kPhaseTileBytes);
            if (l == 1 && g == 0) prefetch_next_unit();
            for (int c = 0; c < kChunks; ++c) {
              mbar_wait(&w_empty[st], ph ^ 1);
              if (kPair) {
                if (crank == 0) mbar_arrive_expect_tx(&w_full[st], 2 * kWChunkBytes);
                const int32_t row =
                    (int32_t)(((size_t)((l - 1) * 2 + crank) * (kWImageBytes / 2) + (size_t)c * kWChunkBytes) / 256);
                tma2_load_2d(smem + BwdSmem::kRing + st * kWChunkBytes, &p.wmap, 0, wrow0 + row,
                             mapa_u32(smem_u32(&w_full[st]), 0));
              } else {
                mbar_arrive_expect_tx(&w_full[st], kWChunkBytes);
                bulk_g2s(smem + BwdSmem::kRing + st * kWChunkBytes,
                         reinterpret_cast<const uint8_t*>(p.wb) + (size_t)(l - 1) * kWImageBytes + (size_t)c * kWChunkBytes,
                         kWChunkBytes, &w_full[st]);
              }
              if (++st == kBwdStages) { st = 0; ph ^= 1; }
            }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (leader CTA: issues for the pair)
    if (lane == 0 && crank == 0) {
      constexpr uint32_t kM = kPair ? 256 : 128;
      constexpr uint32_t idesc_h = umma_idesc_f16(kM, 256, 0, 0);
      constexpr uint32_t idesc_r = umma_idesc_f16(kM, kW6N, 1, 1);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if (kPair) umma2_f16_ss(d, da, db, idesc, acc);
        else umma_f16_ss(d, da, db, idesc, acc);
      };
      auto commit = [&](uint64_t* bar, uint16_t mask) {
        if (kPair) umma2_commit_multicast(bar, mask);
        else umma_commit(bar);
      };
      const uint32_t a_base = smem_u32(smem + BwdSmem::kA);
      const uint32_t ring_base = smem_u32(smem + BwdSmem::kRing);
      const uint32_t w6_base = smem_u32(smem + BwdSmem::kW6);
      uint32_t st = 0, ph = 0;
      uint32_t a_ph = 0, ap_ph = 0;  // bit g: parity of a_ready[g] / a_ready_peer[g]
      uint32_t tn = 0;
      for (int it = 0; it < iters; ++it) {
        const int ubase = unit_base(it);
        const int nsub = clamp02(p.ntiles - ubase);           // live sub-tiles of the leader
        const int nsub_peer = kPair ? clamp02(p.ntiles - ubase - 2) : 0;  // ... of the peer
        for (int l = L + 1; l >= 0; --l) {  // l = L+1: output layer (K = 16); L..1: hidden layer l; 0: layer-0 reduction
          for (int g = 0; g < nsub; ++g) {
            mbar_wait(&a_ready[g], (a_ph >> g) & 1);
            a_ph ^= 1u << g;
            btrace_ev(p, 0, tn, 0x700 | ((uint32_t)l << 4) | g);  // own epilogue group seen
            if (g < nsub_peer) {
              mbar_wait(&a_ready_peer[g], (ap_ph >> g) & 1);
              ap_ph ^= 1u << g;
            }
            tc_fence_after();
            btrace_ev(p, 0, tn, 0x100 | ((uint32_t)l << 4) | g);  // operand ready seen
            const uint32_t a_tile = a_base + g * kTileImageBytes;
            const uint32_t d_tmem = tmem_base + g * 256;
            if (l == L + 1) {
              // A: g_y [128 x 16] of each CTA; B: each CTA's [2][128 k][8] half of W_out''
              const uint64_t da = umma_smem_desc(a_tile, 2048, 128);
              const uint64_t db = umma_smem_desc(w6_base, kBRows * 16, 128);
              mma(d_tmem, da, db, idesc_h, 0);
            } else if (l == 0) {
              // D_r[j, i] = sum_rows delta0_r[row, j] * F[row, i]: MN-major views of [k/8][128][8] images; CTA r
              // contributes columns 8r..8r+7 of B (its own [f | 1] features), so it reads back columns 8r.. of D
              const uint32_t f_tile = smem_u32(smem + BwdSmem::kF) + g * kGyImageBytes;
#pragma unroll
              for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
                for (int ks = 0; ks < kTileRows / 16; ++ks) {
                  const uint64_t da = umma_smem_desc(a_tile + mh * 16 * 2048 + ks * 256, 128, 2048);
                  const uint64_t db = umma_smem_desc(f_tile + ks * 256, 128, 2048);
                  mma(d_tmem + mh * kW6N, da, db, idesc_r, ks != 0);
                }
              }
            } else if (kPair && kBwdLayerResident) {
              const bool first = (g == 0), last = (g == nsub - 1);
              for (int c = 0; c < kChunks; ++c) {
                if (first) {
                  mbar_wait(&w_full[c], ph);
                  tc_fence_after();
                }
                const uint32_t b_tile = ring_base + c * kWChunkBytes;
#pragma unroll
                for (int ks = 0; ks < kSteps; ++ks) {
                  const uint64_t da = umma_smem_desc(a_tile + (c * kSteps + ks) * 4096, 2048, 128);
                  const uint64_t db = umma_smem_desc(b_tile + ks * (kBRows * 32), kBRows * 16, 128);
                  mma(d_tmem, da, db, idesc_h, (c | ks) != 0);
                }
                if (last) commit(&w_empty[c], 0x3);
              }
              if (last) ph ^= 1;
            } else {
              for (int c = 0; c < kChunks; ++c) {
                mbar_wait(&w_full[st], ph);
                tc_fence_after();
                const uint32_t b_tile = ring_base + st * kWChunkBytes;
#pragma unroll
                for (int ks = 0; ks < kSteps; ++ks) {
                  // A: [j/8][128][8] -> 2048 B per 8-column group; B: [j/8][kBRows k][8] -> kBRows * 16 B per group
                  const uint64_t da = umma_smem_desc(a_tile + (c * kSteps + ks) * 4096, 2048, 128);
                  const uint64_t db = umma_smem_desc(b_tile + ks * (kBRows * 32), kBRows * 16, 128);
                  mma(d_tmem, da, db, idesc_h, (c | ks) != 0);
                }
                commit(&w_empty[st], 0x3);
                if (++st == kBwdStages) { st = 0; ph ^= 1; }
              }
            }
            commit(&acc_full[g], (g < nsub_peer) ? 0x3 : 0x1);
            btrace_ev(p, 0, tn, 0x200 | ((uint32_t)l << 4) | g);  // pass issued
          }
        }
      }
    }
  } else {
    // ============================================================ epilogue groups
    const int g = (warp - 2) >> 3;
    const uint32_t e = warp - 2 - 8 * g;
    const uint32_t q = warp & 3;
    const uint32_t chalf = e >> 2;  // 128-column half handled by this warp
    const uint32_t row = q * 32 + lane;
    uint8_t* a_tile = smem + BwdSmem::kA + g * kTileImageBytes;
    const uint32_t t_acc = tmem_base + ((q * 32) << 16) + g * 256 + chalf * 128;
    uint32_t acc_ph = 0;
    uint32_t tn = 0;
    const bool tracer = e == 0 && lane == 0;
    const float S = __ldg(p.scalars);

    for (int it = 0; it < iters; ++it) {
      const int tile = unit_base(it) + 2 * (int)crank + g;
      if (tile >= p.ntiles) continue;  // (the ragged last quad is visited first)
      const int b = tile / p.tiles_per_map;
      const int pix = (tile - b * p.tiles_per_map) * kTileRows + row;
      const bool rvalid = pix < p.P;
      const uint8_t* st_u = reinterpret_cast<const uint8_t*>(p.stash_u) + (size_t)tile * (L + 1) * kPhaseTileBytes;
      uint8_t* st_d = reinterpret_cast<uint8_t*>(p.stash_d) + (size_t)tile * p.d_slots * kTileImageBytes;

      // ---- g_y (scaled by S) -> fp16 [128 x 16] operand at the head of the tile image
      float gy[3] = {0.f, 0.f, 0.f};
      if (rvalid) {
        const size_t e = ((size_t)b * p.P + pix) * 3;
        if (p.grad_out != nullptr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float o = __ldg(p.out + e + c);
            float gg = __ldg(p.grad_out + e + c) * S;
            if (p.out_tanh) gg *= (1.f - o * o);
            if (p.aout != nullptr) gg *= cosf(__ldg(p.aout + e + c));  // d sin(a_out) / d a_out
            gy[c] = gg;
          }
        } else {
          const float* wp = p.sw_grid ? nullptr : p.sw + (size_t)b * p.sw_bstride + (size_t)pix * 3;
          const float wg = p.sw_grid ? grid_sineweight(pix, p.grid_w, p.mask_bits) : 0.f;
          const float* ml = p.map_loss + (size_t)b * 32;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float o = __ldg(p.out + e + c);
            const float t = __ldg(p.target + e + c);
            // S*g_o = (o-t)*sw + coefA*t + coefB*o   with S = 3P/2
            float gg = (o - t) * (p.sw_grid ? wg : __ldg(wp + c));
            if (p.use_cos) gg += __ldg(ml + 16 + c) * t + __ldg(ml + 19 + c) * o;
            if (p.out_tanh) gg *= (1.f - o * o);
            if (p.aout != nullptr) gg *= cosf(__ldg(p.aout + e + c));
            gy[c] = gg;
          }
        }
      }
      if (chalf == 0) {
        // feature row [f0, f1, f2, f3, 1, 0, ...] for the layer-0 reduction GEMM (zero for rows beyond P)
        float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, one = 0.f;
        if (rvalid) {
          float dx, dy, dz;
          if (p.dir_grid) {
            float sp_;
            grid_point(pix, p.grid_w, dx, dy, dz, sp_);
          } else {
            const float* d = p.D + (size_t)b * p.d_bstride + (size_t)pix * 3;
            dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
          }
          if (p.so2) { f0 = dx; f1 = dz; f2 = sqrtf(dx * dx + dz * dz); f3 = dy; }
          else       { f0 = dx; f1 = dy; f2 = dz; }
          one = 1.f;
        }
        uint8_t* f_tile = smem + BwdSmem::kF + g * kGyImageBytes;
        *reinterpret_cast<uint4*>(f_tile + tile_image_off(kTileRows, row, 0)) =
            make_uint4(pack_half2(f0, f1), pack_half2(f2, f3), pack_half2(one, 0.f), 0u);
        *reinterpret_cast<uint4*>(f_tile + tile_image_off(kTileRows, row, 1)) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (chalf == 0) {
        uint4 v0, v1;
        v0.x = pack_half2(gy[0], gy[1]);
        v0.y = pack_half2(gy[2], 0.f);
        v0.z = 0u;
        v0.w = 0u;
        v1 = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, 0)) = v0;
        *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, 1)) = v1;
        if (kNeedDW) {
          uint8_t* sg = reinterpret_cast<uint8_t*>(p.stash_gy) + (size_t)tile * kGyImageBytes;
          *reinterpret_cast<uint4*>(sg + stash_off(row, 0, kW6N)) = v0;
          *reinterpret_cast<uint4*>(sg + stash_off(row, 1, kW6N)) = v1;
        }
      }
      fence_proxy_async_smem();
      signal_ready(g);

      // ---- delta_l = acc * cos(a_l), l = L..0 ; this thread: (row, columns chalf*128 .. +127)
      for (int l = L; l >= 0; --l) {
        const uint8_t* hl = st_u + (size_t)l * kPhaseTileBytes;
        uint8_t* dl = nullptr;
        if (kNeedDW && l > 0) dl = st_d + (size_t)l * kTileImageBytes;
        // the whole layer slice of the phase stash (16 records of 8 columns) is requested before waiting for the GEMM
        PhaseRec hh[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) phase_fetch4<RENI_BWD_PHASE_HINT>(hl, row, chalf * 4 + k, hh[k]);
        if (tracer) btrace_ev(p, 1 + g, tn, 0x300 | ((uint32_t)l << 4) | g);  // phase loads issued, waiting for the accumulator
        mbar_wait(&acc_full[g], acc_ph);
        acc_ph ^= 1;
        tc_fence_after();
        if (tracer) btrace_ev(p, 1 + g, tn, 0x400 | ((uint32_t)l << 4) | g);  // accumulator seen
        const float* fl = nullptr;  // this map's freq_l (hidden layers only; layer 0 is modulated by the caller)
        if (kFilm && l > 0) fl = p.film + ((size_t)b * L + (l - 1)) * 2 * kH;
        if (kNeedDW && p.ready != nullptr && l < L) {
          // overlap mode: delta_{l+1} (and, behind the first signal, g_y) of this warp's block must be complete in global
          // memory -- not just read out of the tile image -- before the tile's counter moves
          if (kBulk && lane < 16) {
            bulk_wait0();
            fence_proxy_async_all();
          }
          __syncwarp();
          if (lane == 0) ready_signal(p.ready + tile);
        } else if (kBulk && kNeedDW) {  // this warp's previous pieces have been read out of the tile image
          if (lane < 16) bulk_wait_read0();
          __syncwarp();
        }
        if (tracer) btrace_ev(p, 1 + g, tn, 0x600 | ((uint32_t)l << 4) | g);  // previous stash copies have left the tile
        auto process16 = [&](const uint32_t (&v)[16], int it) {
#pragma unroll
          for (int q8 = 0; q8 < 2; ++q8) {
            const int kl = it * 2 + q8;
            const int kg = chalf * 16 + kl;
            const PhaseRec& hw = hh[kl >> 2][kl & 3];
            float d[8];  // delta = acc * cos(a), cos rebuilt from the stashed phases
            d[0] = __uint_as_float(v[q8 * 8 + 0]) * abl_cos(phase_angle_of<0>(hw));
            d[1] = __uint_as_float(v[q8 * 8 + 1]) * abl_cos(phase_angle_of<1>(hw));
            d[2] = __uint_as_float(v[q8 * 8 + 2]) * abl_cos(phase_angle_of<2>(hw));
            d[3] = __uint_as_float(v[q8 * 8 + 3]) * abl_cos(phase_angle_of<3>(hw));
            d[4] = __uint_as_float(v[q8 * 8 + 4]) * abl_cos(phase_angle_of<4>(hw));
            d[5] = __uint_as_float(v[q8 * 8 + 5]) * abl_cos(phase_angle_of<5>(hw));
            d[6] = __uint_as_float(v[q8 * 8 + 6]) * abl_cos(phase_angle_of<6>(hw));
            d[7] = __uint_as_float(v[q8 * 8 + 7]) * abl_cos(phase_angle_of<7>(hw));
            uint4 dv;
            dv.x = pack_half2(d[0], d[1]);
            dv.y = pack_half2(d[2], d[3]);
            dv.z = pack_half2(d[4], d[5]);
            dv.w = pack_half2(d[6], d[7]);
            if (kFilm && fl != nullptr) {
              // stash the unscaled delta (streaming store: read back only by the weight-gradient kernel), hand
              // delta * freq to the next GEMM
              __stcs(reinterpret_cast<uint4*>(dl + stash_off(row, kg, kH)), dv);
              const float4 f0 = __ldg(reinterpret_cast<const float4*>(fl + kg * 8));
              const float4 f1 = __ldg(reinterpret_cast<const float4*>(fl + kg * 8 + 4));
              dv.x = pack_half2(d[0] * f0.x, d[1] * f0.y);
              dv.y = pack_half2(d[2] * f0.z, d[3] * f0.w);
              dv.z = pack_half2(d[4] * f1.x, d[5] * f1.y);
              dv.w = pack_half2(d[6] * f1.z, d[7] * f1.w);
              *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, kg)) = dv;
              continue;
            }
            *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, kg)) = dv;
            if (!kBulk && dl != nullptr && !(RENI_ABL & 1))
              *reinterpret_cast<uint4*>(dl + stash_off(row, kg, kH)) = dv;
          }
        };
        {
          uint32_t va[16], vb[16];
          tmem_ld16(t_acc, va);
#pragma unroll
          for (int it = 0; it < 8; it += 2) {
            tmem_ld_wait();
            tmem_ld16(t_acc + (it + 1) * 16, vb);
            process16(va, it);
            tmem_ld_wait();
            if (it + 2 < 8) tmem_ld16(t_acc + (it + 2) * 16, va);
            process16(vb, it + 1);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        if (tracer) btrace_ev(p, 1 + g, tn, 0x500 | ((uint32_t)l << 4) | g);  // this warp's epilogue done
        else if (lane == 0) btrace_ev(p, 1 + warp, tn, 0x500 | ((uint32_t)l << 4) | g);  // (roles 3..18: the other warps)
#if RENI_BWD_SIGNAL_FIRST
        // the next GEMM only READS the tile, like the stash copies below, so the issuer may be told before they are queued
        // (the copy engine drains ~28 B/clk and the issue of a copy blocks behind its queue, profiles/r2_ubench_s2g.txt)
        signal_ready(g);
#endif
        if (kBulk && dl != nullptr && !(RENI_ABL & 1)) {
          // this warp's block of the finished tile image (32 rows x 16 column groups) goes to the stash as 16 bulk
          // copies of 512 B, one per lane: no st.global in the epilogue, the copy engine reads while the tensor core does
          __syncwarp();
          if (lane < 16) {
            const uint32_t kg = chalf * 16 + lane;
            bulk_s2g(dl + (q >> 1) * kHalfImageBytes + kg * (kHalfRows * 16) + (q & 1) * 512,
                     a_tile + kg * (kTileRows * 16) + q * 512, 512);
            bulk_commit();
          }
        }
#if !RENI_BWD_SIGNAL_FIRST
        signal_ready(g);
#endif
      }
      // ---- layer-0 reduction result: D[j, 0..4] for j = row (columns 0..15) and j = 128 + row (columns 16..31)
      mbar_wait(&acc_full[g], acc_ph);
      acc_ph ^= 1;
      tc_fence_after();
      if (chalf == 0) {
        uint32_t v[32];
        tmem_ld32(t_acc, v);
        tmem_ld_wait();
        const float inv_s = __ldg(p.scalars + 1);
        float* dst = p.dmc + (size_t)b * 5 * kH;
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int i = 0; i < 5; ++i)
            atomicAdd(dst + i * kH + mh * 128 + row, __uint_as_float(v[mh * kW6N + crank * 8 + i]) * inv_s);
      }
      tc_fence_before();
    }
  }

  if (kBulk && kNeedDW && warp >= 2 && lane < 16) bulk_wait0();  // stash writes have landed
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // the pair's MMAs, multicast commits and remote arrivals are all behind us
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace reni
