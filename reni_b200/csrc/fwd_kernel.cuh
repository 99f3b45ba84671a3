// Fused RENI decoder forward for sm_100a.
//
//   per 128-direction tile:  f = [dx, dz, |d_xz|, dy]           (SO(2) invariants, registers only)
//                            h0 = sin(f . M_b + c_b)            (layer 0 hoisted to a per-map 4x256 matrix)
//                            h_l = sin(h_{l-1} W_l'^T + b_l')   (tcgen05.mma, fp16 operands, fp32 TMEM accumulators,
//                                                                bias + sin fused in the TMEM->register epilogue)
//                            o = tanh(h_L W_out^T + b_out)      (tcgen05.mma N=16) + optional fused loss partial sums
//
// One persistent CTA per SM, 576 threads:
//   warp 0        : weight-chunk producer (cp.async.bulk global->smem ring, mbarrier complete_tx)
//   warp 1        : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..17   : sixteen epilogue warps over two 128-row sub-tiles (TMEM columns 0..255 / 256..511) that ping-pong:
//                   while the sin epilogue of layer l runs for one sub-tile the tensor pipe runs layer l of the other.
//     grouped mode (training): warps 2..9 own sub-tile 0, warps 10..17 own sub-tile 1; two warps share a TMEM lane
//                   quarter and split the 256 columns in halves.  A store-stalled group only delays its own sub-tile.
//     all-hands mode (inference): all sixteen warps work on one sub-tile at a time (four per lane quarter, 64 columns
//                   each), so no warp idles through a GEMM.  Measured on B200 (tools/variant_time.py): all-hands wins
//                   without stash stores (63.5 % vs 61 % of the bf16 peak at cfg 5), grouped wins with them
//                   (234 vs 262 us at cfg 2) -- hence one mode per use.
// Activations never leave the SM (smem tile image, overwritten in place).
// With kTrain the phase of every pre-activation (12 bits by default, phase.cuh) is additionally stashed in the
// tile-image geometry: the backward kernels rebuild cos(a_l) (delta chain) and h_l = sin(a_l) (weight-gradient GEMM
// operand) from it to ~5e-5, so one 2-byte stash replaces an h stash plus a cos stash and costs no second MUFU here.
// (Rebuilding cos as +-sqrt(1 - h^2) from an fp16 h was tried and measured: gradient error 0.6-1.6e-2, rejected.)
//
// Reference semantics: src/models/RENI.py:31-53 (encoding), :63-87 (SineLayer), :132-178 (net).
#pragma once
#include <cuda.h>  // CUtensorMap

#include "layout.cuh"
#include "phase.cuh"
#include "ptx.cuh"

namespace reni {

constexpr int kFwdThreads = 576;
#ifndef RENI_FWD_STAGES
#define RENI_FWD_STAGES 4
#endif
constexpr int kFwdStages = RENI_FWD_STAGES;
#ifndef RENI_LAYER_RESIDENT
#define RENI_LAYER_RESIDENT 0  // paired kernels: each layer's weight half enters the SM once for both sub-tile passes
                               // (measured: inference unchanged at 67 %, training step 795 -> 856 us -- the refill can only
                               // start behind the last pass, so the ring no longer runs ahead of the tensor pipe)
#endif
constexpr bool kLayerResident = RENI_LAYER_RESIDENT != 0 && RENI_FWD_STAGES == 4;
constexpr int kGroupThreads = 256;  // epilogue threads per sub-tile (grouped mode)
constexpr int kEpiThreads = 512;    // all epilogue threads (all-hands mode)

struct FwdParams {
  const float* D;        // (B or 1, P, 3) unit directions
  int64_t d_bstride;     // elements between maps (0: one grid shared by all maps)
  const float* mc;       // (B, 5, 256): rows 0..3 = omega0*M_b, row 4 = omega0*c_b
  const __half* wf;      // L weight images [k/8][n][8] of omega_l * W_l
  const __half* wf2;     // the same, split for CTA pairs: [l][n half][k/8][128][8]
  const __half* w6f;     // [k/8 32][n 16][8] final-layer image
  const float* bias;     // L*256 (omega_l * b_l) then 16 (final bias, zero padded)
  float* out;            // (B, P, 3)
  uint16_t* stash_u;     // kTrain: per tile (L+1) tile-layer images of the phases of a_l (phase.cuh)
  const float* target;   // fused loss partial sums (optional, may be null)
  const float* sw;       // (B or 1, P, 3)
  int64_t sw_bstride;
  float* loss_part;      // (ntiles, 4 warps, 10)
  float* aout;           // kTrain with a sine output layer: (B, P, 3) pre-activations a_out (the backward needs cos a_out)
  const float* film;     // kFilm: (B, L, 2, 256) per-map FiLM modulation of the hidden layers: a_l = freq_l * acc + phase_l
  int B, P, tiles_per_map, ntiles, L;
  int out_tanh, last_sine, so2;
  int w_map_rows, b_map_rows;  // per-map weight / bias images (FiLM on per-map images): rows of 256 B per map in wmap / bmap,
                               // 0 = one image shared by all maps
  unsigned long long* trace;  // debug: per-role clock64 timeline of CTA 0 (reni_debug_set_trace), else null
  int grid_w;                 // side length W of the analytic equirectangular grid (P = W * W / 2), used by:
  int dir_grid;               // 1: D is not read, directions come from grid_point(pix, grid_w) (RENI_FLAG_GRID_DIRECTIONS)
  int sw_grid;                // 1: sw is not read, sine weights come from grid_sineweight (RENI_FLAG_GRID_SINEWEIGHT)
  const uint32_t* mask_bits;  // sw_grid: one bit per pixel (1 = kept), or null
  int split;                  // paired mode: 1 = two-term weights, every K chunk is followed by its fp16 residual chunk
                              // (W' = W_hi + W_lo, acc = h W_hi^T + h W_lo^T): removes the weight-rounding half of the
                              // fp16 operand error of the hidden layers at twice the tensor-core work
  alignas(64) CUtensorMap wmap;  // paired mode: wf2 as rows of 256 B, box = one 16 KB half chunk (TMA tile loads)
  alignas(64) CUtensorMap wmap_lo;  // ... the residual images wf2lo, same geometry
  alignas(64) CUtensorMap bmap;  // paired mode: wbias2 as rows of 256 B, box = one 4 KB bias block
};

struct FwdSmem {
  static constexpr int kA = 0;                                        // 2 x 64 KB activation tile images
  static constexpr int kRing = kA + 2 * kTileImageBytes;              // weight chunk ring
  static constexpr int kMaxStages = kFwdStages;
  static constexpr int kRingBytes = kFwdStages * kWChunkBytes;
  static constexpr int kW6 = kRing + kRingBytes;                      // final-layer image
  static constexpr int kBias = kW6 + kW6ImageBytes;                   // (kMaxHiddenLayers*256 + 16) floats
  static constexpr int kMc = kBias + (kMaxHiddenLayers * kH + 16) * 4;  // 2 x 5 x 256 floats
  static constexpr int kOnes = kMc + 2 * 5 * kH * 4;                  // [2 k-groups][128][8] fp16: columns (1, 1, 0, ...)
  static constexpr int kBiasSlot = (kOnes + kBiasBlockBytes + 127) / 128 * 128;  // paired: the current layer's bias block (TMA dst)
  static constexpr int kBars = kBiasSlot + kBiasBlockBytes;           // mbarriers
  static constexpr int kNumBars = 2 * (kMaxStages + 1) + 6;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemPtr + 16;
};
static_assert(FwdSmem::kTotal <= 232448, "forward kernel shared memory over budget");

// (debug timeline) wall-clock stamp: the same record with the low 48 bits of %globaltimer (ns) instead of clock64 --
// two such pairs give the SM clock the kernel actually ran at
DEVINL void trace_wall(const FwdParams& p, int role, uint32_t& n, uint32_t code) {
  if (p.trace != nullptr && blockIdx.x == 0 && n < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[role * 4096 + n] = ((unsigned long long)code << 48) | (t & 0xFFFFFFFFFFFFull);
    ++n;
  }
}
// debug timeline: role r (0 MMA issuer, 1/2 epilogue group 0/1) appends (code << 48 | clock) to its 4096-entry lane
DEVINL void trace_ev(const FwdParams& p, int role, uint32_t& n, uint32_t code) {
  if (p.trace != nullptr && blockIdx.x == 0 && n < 4096) {
    p.trace[role * 4096 + n] = ((unsigned long long)code << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    ++n;
  }
}

#ifndef RENI_FWD_PIPE_L0
#define RENI_FWD_PIPE_L0 2  // all-hands epilogue: 1 = layer 0 of the next unit ahead of the current unit's output epilogue,
                           // 2 = and the output epilogue finished behind the next unit's first hidden epilogue
#endif
#ifndef RENI_FWD_STASH_HINT
#define RENI_FWD_STASH_HINT 1  // phase-stash stores: 0 plain st.global, 1 st.global.cs (streaming: fwd 214 -> 207 us), 2 st.global.wt (no change)
#endif
// sin of 8 pre-activations -> packed fp16 and, if kPhase, their packed phases
template <bool kPhase>
DEVINL void sin8(const float (&a)[8], uint4& hv, PhaseRec& uv) {
  hv.x = pack_half2(epi_sin<0>(a[0]), epi_sin<1>(a[1]));
  hv.y = pack_half2(epi_sin<2>(a[2]), epi_sin<3>(a[3]));
  hv.z = pack_half2(epi_sin<4>(a[4]), epi_sin<5>(a[5]));
  hv.w = pack_half2(epi_sin<6>(a[6]), epi_sin<7>(a[7]));
  if (kPhase) uv = phase_encode8(a);
}

// kFilm (FiLM conditioning, RENI.py:515-524,666-678): the hidden layers compute sin(freq_l[b] * (W_l h + b_l) + phase_l[b])
// with per-map vectors freq_l, phase_l; W_l and b_l enter unscaled (omega = 1), the bias rides in the accumulator, and
// the epilogue applies one FFMA per element with the map's (freq, phase) read through L1 (the same 2 KB for a whole
// sub-tile).  Layer 0 arrives hoisted and already modulated in p.mc.
template <bool kTrain, bool kAllHands, bool kPair, bool kFilm = false>
__global__ void __launch_bounds__(kFwdThreads, 1) reni_fwd_kernel(const __grid_constant__ FwdParams p) {
  static_assert(!kFilm || (kAllHands && kPair), "FiLM epilogue exists for the paired all-hands kernel only");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBars);
  constexpr int kStages = kFwdStages;
  uint64_t* w_full = bars;                                 // [kStages + 1]
  uint64_t* w_empty = bars + FwdSmem::kMaxStages + 1;      // [kStages + 1]
  uint64_t* a_ready = bars + 2 * (FwdSmem::kMaxStages + 1);  // [2]  epilogue -> MMA (one arrival per thread)
  uint64_t* acc_full = a_ready + 2;             // [2]  MMA -> epilogue group (tcgen05.commit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + FwdSmem::kTmemPtr);
  float* s_bias = reinterpret_cast<float*>(smem + FwdSmem::kBias);

  const int L = p.L;
  // Work distribution.  Unpaired: CTA c owns tiles {2u, 2u+1}, u = c, c + grid, ...  Paired (cluster of 2,
  // tcgen05.mma.cta_group::2): cluster k owns tile quads q = k, k + nclusters, ...; the leader (rank 0) owns tiles
  // {4q, 4q+1}, its peer {4q+2, 4q+3}.  One M = 256 MMA covers sub-tile g of both CTAs, each CTA holds (and streams
  // from L2) only its 128-column half of every weight matrix: half the weight bytes enter each SM.
  const uint32_t crank = kPair ? cluster_ctarank() : 0;
  const int unit_tiles = kPair ? 4 : 2;
  const int nunits = (p.ntiles + unit_tiles - 1) / unit_tiles;
  const int nworkers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int iters = (nunits - worker + nworkers - 1) / nworkers;  // same for both CTAs of a pair
  auto clamp02 = [](int x) { return x < 0 ? 0 : (x > 2 ? 2 : x); };
  constexpr int kChunkBytes = kWChunkBytes;            // paired: [8 k-groups][128 n][8] = K 64 of this CTA's N half
  constexpr int kChunks = kPair ? 4 : kChunksPerLayer;
  const int nsplit = (kPair && !kLayerResident && p.split) ? 2 : 1;
  uint64_t* a_ready_peer = acc_full + 2;  // [2] leader only: the peer's sub-tile g is ready (one arrival per warp)
  constexpr uint32_t kPeerWarps = kAllHands ? 16 : 8;

  // ---------------------------------------------------------------- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages + 1; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(&a_ready[0], kAllHands ? kEpiThreads : kGroupThreads);
    mbar_init(&a_ready[1], kAllHands ? kEpiThreads : kGroupThreads);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&a_ready_peer[0], kPeerWarps);
    mbar_init(&a_ready_peer[1], kPeerWarps);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc2<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  {  // resident small operands: final-layer weight image + all biases
    const uint4* src = reinterpret_cast<const uint4*>(p.w6f);
    uint4* dst = reinterpret_cast<uint4*>(smem + FwdSmem::kW6);
    if (kPair) {  // this CTA's 8 of the 16 padded output rows: [k/8][8][8]
      for (int i = threadIdx.x; i < kW6ImageBytes / 32; i += kFwdThreads)
        dst[i] = src[(i >> 3) * kW6N + crank * 8 + (i & 7)];
    } else {
      for (int i = threadIdx.x; i < kW6ImageBytes / 16; i += kFwdThreads) dst[i] = src[i];
    }
    const int nb = L * kH + 16;
    for (int i = threadIdx.x; i < nb; i += kFwdThreads) s_bias[i] = p.bias[i];
    // A operand of the bias K-step: every row = (1, 1, 0, ..., 0)
    uint4* ones = reinterpret_cast<uint4*>(smem + FwdSmem::kOnes);
    for (int i = threadIdx.x; i < kBiasBlockBytes / 16; i += kFwdThreads)
      ones[i] = (i < kTileRows) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // "sub-tile g of this CTA is in shared memory": the leader's (or an unpaired CTA's) epilogue threads arrive on the
  // local barrier; the peer's warps arrive on the leader's a_ready_peer, one remote arrival per warp
  const uint32_t ra_peer = kPair ? mapa_u32(smem_u32(a_ready_peer), 0) : 0;
  auto signal_ready = [&](int g) {
    if (kPair && crank == 1) {
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(ra_peer + g * 8);
    } else {
      mbar_arrive(&a_ready[g]);
    }
  };

  if (warp == 0) {
    // ============================================================ weight-chunk producer
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(kPair ? p.wf2 : p.wf);
      for (int it = 0; it < iters; ++it) {
        const int ubase = (worker + it * nworkers) * unit_tiles;
        const int nstream = clamp02(p.ntiles - ubase);  // passes the (leader's) MMA issuer makes over each layer
        const int umap = ubase / p.tiles_per_map;       // (per-map images: the unit's four tiles lie in one map)
        const int32_t wrow0 = umap * p.w_map_rows, brow0 = umap * p.b_map_rows;
        if (kPair && kLayerResident) {
          // Layer-resident weights: ring slot c holds chunk c of the current layer for BOTH sub-tile passes of the pair
          // (the ring is exactly one layer's 128-column half), so the half enters the SM once per tile quad and layer
          // instead of once per pass; slot kStages holds the layer's bias block.  A slot is released by the commit
          // behind the LAST pass's MMAs on it, and refilled with the next layer's chunk while the epilogues run.
          for (int l = 0; l < L; ++l) {
            mbar_wait(&w_empty[kStages], ph ^ 1);
            if (crank == 0) mbar_arrive_expect_tx(&w_full[kStages], 2 * kBiasBlockBytes);
            tma2_load_2d(smem + FwdSmem::kBiasSlot, &p.bmap, 0,
                         brow0 + (int32_t)((l * 2 + crank) * (kBiasBlockBytes / 256)), mapa_u32(smem_u32(&w_full[kStages]), 0));
            for (int c = 0; c < kChunks; ++c) {
              mbar_wait(&w_empty[c], ph ^ 1);
              if (crank == 0) mbar_arrive_expect_tx(&w_full[c], 2 * kChunkBytes);
              const int32_t row = (int32_t)(((size_t)(l * 2 + crank) * (kWImageBytes / 2) + (size_t)c * kChunkBytes) / 256);
              tma2_load_2d(smem + FwdSmem::kRing + c * kChunkBytes, &p.wmap, 0, wrow0 + row, mapa_u32(smem_u32(&w_full[c]), 0));
            }
            ph ^= 1;
          }
        } else {
        for (int l = 0; l < L; ++l) {
          for (int g = 0; g < nstream; ++g) {
            if (kPair) {  // the layer's bias block rides the ring as a small extra chunk ahead of the weights
              mbar_wait(&w_empty[st], ph ^ 1);
              if (crank == 0) mbar_arrive_expect_tx(&w_full[st], 2 * kBiasBlockBytes);
              tma2_load_2d(smem + FwdSmem::kRing + st * kChunkBytes, &p.bmap, 0,
                           brow0 + (int32_t)((l * 2 + crank) * (kBiasBlockBytes / 256)), mapa_u32(smem_u32(&w_full[st]), 0));
              if (++st == kStages) { st = 0; ph ^= 1; }
            }
            for (int c = 0; c < kChunks; ++c) {
              for (int sp = 0; sp < nsplit; ++sp) {  // the chunk of W_hi, then (two-term weights) the chunk of W_lo
              mbar_wait(&w_empty[st], ph ^ 1);
              if (kPair) {
                // both halves complete on the LEADER's barrier (TMA tile load with .cta_group::2): no relay, one wait
                if (crank == 0) mbar_arrive_expect_tx(&w_full[st], 2 * kChunkBytes);
                const int32_t row = (int32_t)(((size_t)(l * 2 + crank) * (kWImageBytes / 2) + (size_t)c * kChunkBytes) / 256);
                tma2_load_2d(smem + FwdSmem::kRing + st * kChunkBytes, sp ? &p.wmap_lo : &p.wmap, 0, wrow0 + row,
                             mapa_u32(smem_u32(&w_full[st]), 0));
              } else {
                mbar_arrive_expect_tx(&w_full[st], kChunkBytes);
                bulk_g2s(smem + FwdSmem::kRing + st * kChunkBytes, wsrc + (size_t)l * kWImageBytes + (size_t)c * kChunkBytes,
                         kChunkBytes, &w_full[st]);
              }
              if (++st == kStages) { st = 0; ph ^= 1; }
              }
            }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ============================================================ MMA issuer (paired: the leader issues for both)
      constexpr uint32_t idesc_h = umma_idesc_f16(kPair ? 256 : 128, 256, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(kPair ? 256 : 128, kW6N, 0, 0);
      constexpr uint32_t kBRows = kPair ? 128 : 256;   // rows of B in this CTA's chunk
      constexpr uint32_t kW6Rows = kPair ? 8 : kW6N;
      const uint32_t a_base = smem_u32(smem + FwdSmem::kA);
      const uint32_t ring_base = smem_u32(smem + FwdSmem::kRing);
      const uint32_t w6_base = smem_u32(smem + FwdSmem::kW6);
      uint32_t st = 0, ph = 0;
      uint32_t a_ph = 0, ap_ph = 0;  // bit g: parity of a_ready[g] / a_ready_peer[g]
      uint32_t tn = 0;
      trace_ev(p, 0, tn, 0x900);    // issuer starts (clock64) ...
      trace_wall(p, 0, tn, 0xA00);  // ... and the wall clock at the same moment
      for (int it = 0; it < iters; ++it) {
        const int ubase = (worker + it * nworkers) * unit_tiles;
        const int nsub = clamp02(p.ntiles - ubase);                      // live sub-tiles of this (the leader) CTA
        const int nsub_peer = kPair ? clamp02(p.ntiles - ubase - 2) : 0;  // ... of the peer
        for (int l = 1; l <= L + 1; ++l) {
          for (int g = 0; g < nsub; ++g) {
            mbar_wait(&a_ready[g], (a_ph >> g) & 1);
            a_ph ^= 1u << g;
            if (g < nsub_peer) {
              mbar_wait(&a_ready_peer[g], (ap_ph >> g) & 1);
              ap_ph ^= 1u << g;
            }
            tc_fence_after();
            trace_ev(p, 0, tn, 0x100 | (l << 4) | g);  // operand ready seen
            const uint32_t a_tile = a_base + g * kTileImageBytes;
            const uint32_t d_tmem = tmem_base + g * 256;
            const uint16_t live_mask = (g < nsub_peer) ? 0x3 : 0x1;
            if (kPair && kLayerResident && l <= L) {
              // layer-resident weights: wait for each slot's fill in the first pass only, release it in the last
              const bool first = (g == 0), last = (g == nsub - 1);
              if (first) mbar_wait(&w_full[kStages], ph);
              tc_fence_after();
              {
                const uint64_t da = umma_smem_desc(smem_u32(smem + FwdSmem::kOnes), 2048, 128);
                const uint64_t db = umma_smem_desc(smem_u32(smem + FwdSmem::kBiasSlot), 2048, 128);
                umma2_f16_ss(d_tmem, da, db, idesc_h, 0);
                if (last) umma2_commit_multicast(&w_empty[kStages], 0x3);
              }
              for (int c = 0; c < kChunks; ++c) {
                if (first) {
                  mbar_wait(&w_full[c], ph);
                  tc_fence_after();
                }
                const uint32_t b_tile = ring_base + c * kChunkBytes;
                constexpr int kSteps = kChunkBytes / (kBRows * 32);  // K = 16 steps per chunk
#pragma unroll
                for (int ks = 0; ks < kSteps; ++ks) {
                  const uint64_t da = umma_smem_desc(a_tile + (c * kSteps + ks) * 4096, 2048, 128);
                  const uint64_t db = umma_smem_desc(b_tile + ks * (kBRows * 32), kBRows * 16, 128);
                  umma2_f16_ss(d_tmem, da, db, idesc_h, 1);
                }
                if (last) umma2_commit_multicast(&w_empty[c], 0x3);
              }
              if (last) ph ^= 1;
            } else if (l <= L) {
              if (kPair) {  // acc = [1 1 0 ..] x [b_hi b_lo 0 ..]^T = bias, then the weight chunks accumulate on top
                mbar_wait(&w_full[st], ph);
                tc_fence_after();
                const uint64_t da = umma_smem_desc(smem_u32(smem + FwdSmem::kOnes), 2048, 128);
                const uint64_t db = umma_smem_desc(ring_base + st * kChunkBytes, 2048, 128);
                umma2_f16_ss(d_tmem, da, db, idesc_h, 0);
                umma2_commit_multicast(&w_empty[st], 0x3);
                if (++st == kStages) { st = 0; ph ^= 1; }
              }
              for (int c = 0; c < kChunks; ++c) {
                for (int sp = 0; sp < nsplit; ++sp) {  // (two-term weights: the same A columns against W_hi, then W_lo)
                mbar_wait(&w_full[st], ph);
                tc_fence_after();
                const uint32_t b_tile = ring_base + st * kChunkBytes;
                constexpr int kSteps = kChunkBytes / (kBRows * 32);  // K = 16 steps per chunk
#pragma unroll
                for (int ks = 0; ks < kSteps; ++ks) {
                  // A: [k/8][128][8] -> 2048 B per 8-column group; B: [k/8][kBRows][8] -> kBRows*16 B per group
                  const uint64_t da = umma_smem_desc(a_tile + (c * kSteps + ks) * 4096, 2048, 128);
                  const uint64_t db = umma_smem_desc(b_tile + ks * (kBRows * 32), kBRows * 16, 128);
                  if (kPair) umma2_f16_ss(d_tmem, da, db, idesc_h, 1);
                  else umma_f16_ss(d_tmem, da, db, idesc_h, (c | ks) != 0);
                }
                if (kPair) umma2_commit_multicast(&w_empty[st], 0x3);
                else umma_commit(&w_empty[st]);
                if (++st == kStages) { st = 0; ph ^= 1; }
                }
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < kH / 16; ++ks) {
                const uint64_t da = umma_smem_desc(a_tile + ks * 4096, 2048, 128);
                const uint64_t db = umma_smem_desc(w6_base + ks * (kW6Rows * 32), kW6Rows * 16, 128);
                if (kPair) umma2_f16_ss(d_tmem, da, db, idesc_o, ks != 0);
                else umma_f16_ss(d_tmem, da, db, idesc_o, ks != 0);
              }
            }
            if (kPair) umma2_commit_multicast(&acc_full[g], live_mask);
            else umma_commit(&acc_full[g]);
            trace_ev(p, 0, tn, 0x200 | (l << 4) | g);  // all MMAs of this pass issued
          }
        }
      }
      trace_ev(p, 0, tn, 0x900);
      trace_wall(p, 0, tn, 0xA00);
    }
  } else if (kAllHands) {
    // ============================================================ epilogue warps
    const uint32_t e = warp - 2;                    // 0..15
    const uint32_t q = warp & 3;                    // TMEM lane quarter this warp may access
    const uint32_t cq = e >> 2;                     // which 64-column quarter this warp handles
    const uint32_t row = q * 32 + lane;             // row inside a sub-tile == TMEM lane
    const uint32_t etid = threadIdx.x - 64;         // 0..511
    float* s_mc_all = reinterpret_cast<float*>(smem + FwdSmem::kMc);
    uint32_t acc_ph = 0;  // bit g = phase parity of acc_full[g]
    auto unit_tbase = [&](int it) { return (worker + it * nworkers) * unit_tiles + 2 * (int)crank; };  // this CTA's first tile
    auto tile_of = [&](int tb, int g) { return g ? min(tb + 1, p.ntiles - 1) : tb; };

    // per-map layer-0 operands of a unit's sub-tiles -> smem
    auto load_mc = [&](int tb) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4* src = reinterpret_cast<const float4*>(p.mc + (size_t)(tile_of(tb, g) / p.tiles_per_map) * 5 * kH);
        float4* dst = reinterpret_cast<float4*>(s_mc_all + g * 5 * kH);
        for (int i = etid; i < 5 * kH / 4; i += kEpiThreads) dst[i] = __ldg(src + i);
      }
    };
    // layer 0 of sub-tile g of the unit at tb: h0 = sin(f . M' + c')  (omega folded into M', c'); this warp's 64 columns
    auto layer0 = [&](int tb, int g) {
      const int tile = tile_of(tb, g);
      const int b = tile / p.tiles_per_map;
      const int pix = (tile - b * p.tiles_per_map) * kTileRows + row;
      // invariant direction features, registers only (RENI.py:37-49)
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
      if (pix < p.P) {
        float dx, dy, dz;
        if (p.dir_grid) {
          float sp_;
          grid_point(pix, p.grid_w, dx, dy, dz, sp_);
        } else {
          const float* d = p.D + (size_t)b * p.d_bstride + (size_t)pix * 3;
          dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
        }
        if (p.so2) {
          f0 = dx;
          f1 = dz;
          f2 = sqrtf(dx * dx + dz * dz);
          f3 = dy;
        } else {  // SO3 / None: the three inner-product columns (RENI.py:25,57)
          f0 = dx;
          f1 = dy;
          f2 = dz;
        }
      }
      uint8_t* a_tile = smem + FwdSmem::kA + g * kTileImageBytes;
      const float* s_mc = s_mc_all + g * 5 * kH;
      uint8_t* st_u = nullptr;
      if (kTrain) st_u = reinterpret_cast<uint8_t*>(p.stash_u) + (size_t)(tb + g) * (L + 1) * kPhaseTileBytes;
      PhaseRec ubuf[4];
#if RENI_PHASE_BITS == 16
#pragma unroll 2
#else
#pragma unroll 4
#endif
      for (int k8 = 0; k8 < 8; ++k8) {
        const int kg = cq * 8 + k8;
        float a[8];
        {
          const float4* m = reinterpret_cast<const float4*>(s_mc + kg * 8);
          const float4 c0 = m[4 * (kH / 4)], c1 = m[4 * (kH / 4) + 1];
          a[0] = c0.x; a[1] = c0.y; a[2] = c0.z; a[3] = c0.w;
          a[4] = c1.x; a[5] = c1.y; a[6] = c1.z; a[7] = c1.w;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float fi = (i == 0) ? f0 : (i == 1) ? f1 : (i == 2) ? f2 : f3;
            const float4 m0 = m[i * (kH / 4)], m1 = m[i * (kH / 4) + 1];
            a[0] = fmaf(fi, m0.x, a[0]); a[1] = fmaf(fi, m0.y, a[1]);
            a[2] = fmaf(fi, m0.z, a[2]); a[3] = fmaf(fi, m0.w, a[3]);
            a[4] = fmaf(fi, m1.x, a[4]); a[5] = fmaf(fi, m1.y, a[5]);
            a[6] = fmaf(fi, m1.z, a[6]); a[7] = fmaf(fi, m1.w, a[7]);
          }
        }
        uint4 hv;
        PhaseRec uv;
        sin8<kTrain>(a, hv, uv);
        *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, kg)) = hv;
        if (kTrain && !(RENI_ABL & 1)) phase_put<RENI_FWD_STASH_HINT>(ubuf, k8 & 3, st_u, row, kg, uv);
      }
      fence_proxy_async_smem();
      signal_ready(g);
    };
    // output layer (N = 16 padded), first half: wait for the accumulator and take the three pre-activations out of TMEM
    auto read_out = [&](int g, float (&y)[3]) {
      mbar_wait(&acc_full[g], (acc_ph >> g) & 1);
      acc_ph ^= 1u << g;
      tc_fence_after();
      if (cq == 0) {
        const uint32_t t_acc = tmem_base + ((q * 32) << 16) + g * 256;
        uint32_t v[16];
        tmem_ld16(t_acc, v);
        tmem_ld_wait();
        const float* bo = s_bias + L * kH;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          y[c] = (__uint_as_float(v[c]) + __uint_as_float(v[c + 3])) + bo[c];  // (W_out hi + lo columns)
      }
      tc_fence_before();
    };
    // ... second half: optional sin, optional tanh, store, fused loss partials
    auto finish_out = [&](int tb, int g, const float (&y)[3]) {
      if (cq != 0) return;
      const int tile = tb + g;
      const int b = tile_of(tb, g) / p.tiles_per_map;
      const int pix = (tile_of(tb, g) - b * p.tiles_per_map) * kTileRows + row;
      const bool rvalid = pix < p.P;
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float yy = y[c];
        if (p.last_sine) {
          if (kTrain && rvalid) p.aout[((size_t)b * p.P + pix) * 3 + c] = yy;
          yy = sinf(yy);
        }
        if (p.out_tanh) yy = tanhf(yy);
        o[c] = yy;
      }
      if (rvalid) {
        float* op = p.out + ((size_t)b * p.P + pix) * 3;
        op[0] = o[0];
        op[1] = o[1];
        op[2] = o[2];
      }
      if (p.loss_part != nullptr) {
        float part[kLossPartials];
#pragma unroll
        for (int i = 0; i < kLossPartials; ++i) part[i] = 0.f;
        if (rvalid) {
          const float* tp = p.target + ((size_t)b * p.P + pix) * 3;
          const float* wp = p.sw_grid ? nullptr : p.sw + (size_t)b * p.sw_bstride + (size_t)pix * 3;
          const float wg = p.sw_grid ? grid_sineweight(pix, p.grid_w, p.mask_bits) : 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float t = __ldg(tp + c), w = p.sw_grid ? wg : __ldg(wp + c);
            const float er = o[c] - t;
            part[0] = fmaf(er * er, w, part[0]);
            part[1 + c] = o[c] * t;
            part[4 + c] = o[c] * o[c];
            part[7 + c] = t * t;
          }
        }
#pragma unroll
        for (int i = 0; i < kLossPartials; ++i) {
          float x = part[i];
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
          part[i] = x;
        }
        if (lane == 0) {
          float* lp = p.loss_part + ((size_t)tile * 4 + q) * kLossPartials;
#pragma unroll
          for (int i = 0; i < kLossPartials; ++i) lp[i] = part[i];
        }
      }
    };

    // Layer 0 of unit it+1 is computed at the END of unit it, between reading the output accumulator of a sub-tile and
    // finishing its outputs (RENI_FWD_PIPE_L0): the activation tile of sub-tile g is free as soon as its output GEMM
    // has completed, so the first hidden GEMM of the next unit starts ~7 K clk earlier than when the whole output
    // epilogue of both sub-tiles came first (timeline: the tensor pipe idled ~10 K of 69 K clk at every unit boundary).
    if (iters > 0 && clamp02(p.ntiles - unit_tbase(0)) > 0) {
      const int tb = unit_tbase(0), ns = clamp02(p.ntiles - tb);
      load_mc(tb);
      named_bar_sync(1, kEpiThreads);
      for (int g = 0; g < ns; ++g) layer0(tb, g);
    }
    // outputs of the previous unit that are still to be finished (behind the first hidden epilogue of this unit, where
    // the epilogue warps have slack: the tensor pipe is busy with the other sub-tile's first GEMM)
    float y0[3] = {0.f, 0.f, 0.f}, y1[3] = {0.f, 0.f, 0.f};
    int pend_tb = 0, pend_n = 0;
    for (int it = 0; it < iters; ++it) {
      const int tbase = unit_tbase(it);
      const int nsub = clamp02(p.ntiles - tbase);
      if (nsub == 0) break;
      const int bmap0 = tile_of(tbase, 0) / p.tiles_per_map, bmap1 = tile_of(tbase, 1) / p.tiles_per_map;

      // ---- hidden layers: bias + sin epilogue, TMEM -> registers -> smem tile image (in place); the sub-tiles alternate
      for (int l = 1; l <= L; ++l) {
        const float* bl = s_bias + (l - 1) * kH + cq * 64;
        for (int g = 0; g < nsub; ++g) {
          uint8_t* a_tile = smem + FwdSmem::kA + g * kTileImageBytes;
          uint8_t* su = nullptr;
          if (kTrain)
            su = reinterpret_cast<uint8_t*>(p.stash_u) + ((size_t)(tbase + g) * (L + 1) + l) * kPhaseTileBytes;
          const uint32_t t_acc = tmem_base + ((q * 32) << 16) + g * 256 + cq * 64;
          const float* fl = nullptr;  // this map's (freq_l, phase_l)
          if (kFilm) fl = p.film + ((size_t)(g ? bmap1 : bmap0) * L + (l - 1)) * 2 * kH;
          mbar_wait(&acc_full[g], (acc_ph >> g) & 1);
          acc_ph ^= 1u << g;
          tc_fence_after();
          // TMEM -> registers in 16-column slices, double buffered: the next tcgen05.ld is in flight while this
          // slice goes through bias + sin + pack + store
          PhaseRec ubuf[4];  // (12-bit phases leave as groups of 32 columns: phase.cuh)
          auto process16 = [&](const uint32_t (&v)[16], int it) {
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
              const int kl = it * 2 + q8;          // 8-column group inside this warp's quarter
              const int kg = cq * 8 + kl;          // ... inside the tile
              float a[8];
              if (kFilm) {  // a = freq * (acc incl. bias) + phase, per-map vectors (uniform over the warp's rows)
                const float4 f0 = __ldg(reinterpret_cast<const float4*>(fl + kg * 8));
                const float4 f1 = __ldg(reinterpret_cast<const float4*>(fl + kg * 8 + 4));
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(fl + kH + kg * 8));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(fl + kH + kg * 8 + 4));
                a[0] = fmaf(__uint_as_float(v[q8 * 8 + 0]), f0.x, s0.x);
                a[1] = fmaf(__uint_as_float(v[q8 * 8 + 1]), f0.y, s0.y);
                a[2] = fmaf(__uint_as_float(v[q8 * 8 + 2]), f0.z, s0.z);
                a[3] = fmaf(__uint_as_float(v[q8 * 8 + 3]), f0.w, s0.w);
                a[4] = fmaf(__uint_as_float(v[q8 * 8 + 4]), f1.x, s1.x);
                a[5] = fmaf(__uint_as_float(v[q8 * 8 + 5]), f1.y, s1.y);
                a[6] = fmaf(__uint_as_float(v[q8 * 8 + 6]), f1.z, s1.z);
                a[7] = fmaf(__uint_as_float(v[q8 * 8 + 7]), f1.w, s1.w);
              } else if (kPair) {  // the bias is already in the accumulator
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __uint_as_float(v[q8 * 8 + i]);
              } else {
                const float4 b0 = *reinterpret_cast<const float4*>(bl + kl * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(bl + kl * 8 + 4);
                a[0] = __uint_as_float(v[q8 * 8 + 0]) + b0.x;
                a[1] = __uint_as_float(v[q8 * 8 + 1]) + b0.y;
                a[2] = __uint_as_float(v[q8 * 8 + 2]) + b0.z;
                a[3] = __uint_as_float(v[q8 * 8 + 3]) + b0.w;
                a[4] = __uint_as_float(v[q8 * 8 + 4]) + b1.x;
                a[5] = __uint_as_float(v[q8 * 8 + 5]) + b1.y;
                a[6] = __uint_as_float(v[q8 * 8 + 6]) + b1.z;
                a[7] = __uint_as_float(v[q8 * 8 + 7]) + b1.w;
              }
              uint4 hv;
              PhaseRec uv;
              sin8<kTrain>(a, hv, uv);
              *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, kg)) = hv;
              if (kTrain && !(RENI_ABL & 1)) phase_put<RENI_FWD_STASH_HINT>(ubuf, kl & 3, su, row, kg, uv);
            }
          };
          {
            uint32_t va[16], vb[16];
            tmem_ld16(t_acc, va);
#pragma unroll
            for (int it = 0; it < 4; it += 2) {
              tmem_ld_wait();
              tmem_ld16(t_acc + (it + 1) * 16, vb);
              process16(va, it);
              tmem_ld_wait();
              if (it + 2 < 4) tmem_ld16(t_acc + (it + 2) * 16, va);
              process16(vb, it + 1);
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          signal_ready(g);
          if (pend_n > 0) {
            finish_out(pend_tb, 0, y0);
            if (pend_n > 1) finish_out(pend_tb, 1, y1);
            pend_n = 0;
          }
        }
      }

      // ---- unit boundary: output layer of this unit, layer 0 of the next one
      const int tb_next = unit_tbase(it + 1);
      const int ns_next = (it + 1 < iters) ? clamp02(p.ntiles - tb_next) : 0;
      if (ns_next > 0) {
        // (every warp has seen a hidden-layer accumulator of each live sub-tile, i.e. all of them are past this unit's
        // layer 0: its operands can be replaced.  L = 0 has no such guarantee: the barrier in front covers it.)
        if (L == 0 || nsub < 2) named_bar_sync(1, kEpiThreads);
        load_mc(tb_next);
        named_bar_sync(1, kEpiThreads);
      }
      if (pend_n > 0) {  // (L = 0: nothing ran in between)
        finish_out(pend_tb, 0, y0);
        if (pend_n > 1) finish_out(pend_tb, 1, y1);
        pend_n = 0;
      }
      read_out(0, y0);
      if (RENI_FWD_PIPE_L0 && ns_next > 0) layer0(tb_next, 0);
      if (nsub > 1) read_out(1, y1);
      if (RENI_FWD_PIPE_L0 && ns_next > 1) layer0(tb_next, 1);
      if (RENI_FWD_PIPE_L0 == 2 && ns_next > 0) {
        pend_tb = tbase;
        pend_n = nsub;
      } else {
        finish_out(tbase, 0, y0);
        if (nsub > 1) finish_out(tbase, 1, y1);
      }
      if (!RENI_FWD_PIPE_L0)
        for (int g = 0; g < ns_next; ++g) layer0(tb_next, g);
    }
  } else {
    // ============================================================ epilogue groups
    const int g = (warp - 2) >> 3;                  // sub-tile handled by this group
    const uint32_t e = warp - 2 - 8 * g;            // warp inside the group, 0..7
    const uint32_t q = warp & 3;                    // TMEM lane quarter this warp may access
    const uint32_t chalf = e >> 2;                  // which 128-column half this warp handles
    const uint32_t row = q * 32 + lane;             // row inside the tile == TMEM lane
    const uint32_t gtid = e * 32 + lane;            // 0..255 inside the group
    uint8_t* a_tile = smem + FwdSmem::kA + g * kTileImageBytes;
    float* s_mc = reinterpret_cast<float*>(smem + FwdSmem::kMc) + g * 5 * kH;
    const uint32_t t_acc = tmem_base + ((q * 32) << 16) + g * 256 + chalf * 128;
    uint32_t acc_ph = 0;
    uint32_t tn = 0;

    for (int it = 0; it < iters; ++it) {
      const int tile = (worker + it * nworkers) * unit_tiles + 2 * (int)crank + g;
      if (tile >= p.ntiles) break;
      const int b = tile / p.tiles_per_map;
      const int pix = (tile - b * p.tiles_per_map) * kTileRows + row;
      const bool rvalid = pix < p.P;
      uint8_t* st_u = nullptr;
      if (kTrain) st_u = reinterpret_cast<uint8_t*>(p.stash_u) + (size_t)tile * (L + 1) * kPhaseTileBytes;

      // ---- per-map layer-0 operands -> smem (group-private)
      named_bar_sync(1 + g, kGroupThreads);
      {
        const float4* src = reinterpret_cast<const float4*>(p.mc + (size_t)b * 5 * kH);
        float4* dst = reinterpret_cast<float4*>(s_mc);
        for (int i = gtid; i < 5 * kH / 4; i += kGroupThreads) dst[i] = __ldg(src + i);
      }
      named_bar_sync(1 + g, kGroupThreads);

      // ---- invariant direction features, registers only (RENI.py:37-49)
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
      if (rvalid) {
        float dx, dy, dz;
        if (p.dir_grid) {
          float sp_;
          grid_point(pix, p.grid_w, dx, dy, dz, sp_);
        } else {
          const float* d = p.D + (size_t)b * p.d_bstride + (size_t)pix * 3;
          dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
        }
        if (p.so2) {
          f0 = dx;
          f1 = dz;
          f2 = sqrtf(dx * dx + dz * dz);
          f3 = dy;
        } else {  // SO3 / None: the three inner-product columns (RENI.py:25,57)
          f0 = dx;
          f1 = dy;
          f2 = dz;
        }
      }

      // ---- layer 0: h0 = sin(f . M' + c')  (omega folded into M', c'); this warp's 128 columns
      PhaseRec ubuf0[4];
#if RENI_PHASE_BITS == 16
#pragma unroll 2
#else
#pragma unroll 4
#endif
      for (int k8 = 0; k8 < 16; ++k8) {
        const int kg = chalf * 16 + k8;
        float a[8];
        {
          const float4* m = reinterpret_cast<const float4*>(s_mc + kg * 8);
          const float4 c0 = m[4 * (kH / 4)], c1 = m[4 * (kH / 4) + 1];
          a[0] = c0.x; a[1] = c0.y; a[2] = c0.z; a[3] = c0.w;
          a[4] = c1.x; a[5] = c1.y; a[6] = c1.z; a[7] = c1.w;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float fi = (i == 0) ? f0 : (i == 1) ? f1 : (i == 2) ? f2 : f3;
            const float4 m0 = m[i * (kH / 4)], m1 = m[i * (kH / 4) + 1];
            a[0] = fmaf(fi, m0.x, a[0]); a[1] = fmaf(fi, m0.y, a[1]);
            a[2] = fmaf(fi, m0.z, a[2]); a[3] = fmaf(fi, m0.w, a[3]);
            a[4] = fmaf(fi, m1.x, a[4]); a[5] = fmaf(fi, m1.y, a[5]);
            a[6] = fmaf(fi, m1.z, a[6]); a[7] = fmaf(fi, m1.w, a[7]);
          }
        }
        uint4 hv;
        PhaseRec uv;
        sin8<kTrain>(a, hv, uv);
        *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, kg)) = hv;
        if (kTrain && !(RENI_ABL & 1)) phase_put<RENI_FWD_STASH_HINT>(ubuf0, k8 & 3, st_u, row, kg, uv);
      }
      fence_proxy_async_smem();
      signal_ready(g);

      // ---- hidden layers: bias + sin epilogue, TMEM -> registers -> smem tile image (in place)
      for (int l = 1; l <= L; ++l) {
        const float* bl = s_bias + (l - 1) * kH + chalf * 128;
        uint8_t* su = kTrain ? st_u + (size_t)l * kPhaseTileBytes : nullptr;
        if (e == 0 && lane == 0) trace_ev(p, 1 + g, tn, 0x300 | (l << 4) | g);  // waiting for the accumulator
        mbar_wait(&acc_full[g], acc_ph);
        acc_ph ^= 1;
        tc_fence_after();
        if (e == 0 && lane == 0) trace_ev(p, 1 + g, tn, 0x400 | (l << 4) | g);  // accumulator seen
        // TMEM -> registers in 16-column slices, double buffered: the next tcgen05.ld is in flight while this
        // slice goes through bias + sin + pack + store
        PhaseRec ubuf[4];
        auto process16 = [&](const uint32_t (&v)[16], int it) {
#pragma unroll
          for (int q8 = 0; q8 < 2; ++q8) {
            const int kl = it * 2 + q8;          // 8-column group inside this warp's half
            const int kg = chalf * 16 + kl;      // ... inside the tile
            float a[8];
            if (kPair) {  // the bias is already in the accumulator
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = __uint_as_float(v[q8 * 8 + i]);
            } else {
              const float4 b0 = *reinterpret_cast<const float4*>(bl + kl * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bl + kl * 8 + 4);
              a[0] = __uint_as_float(v[q8 * 8 + 0]) + b0.x;
              a[1] = __uint_as_float(v[q8 * 8 + 1]) + b0.y;
              a[2] = __uint_as_float(v[q8 * 8 + 2]) + b0.z;
              a[3] = __uint_as_float(v[q8 * 8 + 3]) + b0.w;
              a[4] = __uint_as_float(v[q8 * 8 + 4]) + b1.x;
              a[5] = __uint_as_float(v[q8 * 8 + 5]) + b1.y;
              a[6] = __uint_as_float(v[q8 * 8 + 6]) + b1.z;
              a[7] = __uint_as_float(v[q8 * 8 + 7]) + b1.w;
            }
            uint4 hv;
            PhaseRec uv;
            sin8<kTrain>(a, hv, uv);
            *reinterpret_cast<uint4*>(a_tile + tile_image_off(kTileRows, row, kg)) = hv;
            if (kTrain && !(RENI_ABL & 1)) phase_put<RENI_FWD_STASH_HINT>(ubuf, kl & 3, su, row, kg, uv);
          }
        };
        {
          uint32_t va[16], vb[16];
          tmem_ld16(t_acc, va);
#pragma unroll 1
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
        signal_ready(g);
        if (e == 0 && lane == 0) trace_ev(p, 1 + g, tn, 0x500 | (l << 4) | g);  // this warp's epilogue done
      }

      // ---- output layer (N = 16 padded): bias, optional sin, optional tanh, store, fused loss partials
      mbar_wait(&acc_full[g], acc_ph);
      acc_ph ^= 1;
      tc_fence_after();
      if (chalf == 0) {
        float o[3];
        {
          uint32_t v[16];
          tmem_ld16(t_acc, v);
          tmem_ld_wait();
          const float* bo = s_bias + L * kH;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float y = (__uint_as_float(v[c]) + __uint_as_float(v[c + 3])) + bo[c];  // (W_out hi + lo columns)
            if (p.last_sine) {
              if (kTrain && rvalid) p.aout[((size_t)b * p.P + pix) * 3 + c] = y;
              y = sinf(y);
            }
            if (p.out_tanh) y = tanhf(y);
            o[c] = y;
          }
        }
        if (rvalid) {
          float* op = p.out + ((size_t)b * p.P + pix) * 3;
          op[0] = o[0];
          op[1] = o[1];
          op[2] = o[2];
        }
        if (p.loss_part != nullptr) {
          float part[kLossPartials];
#pragma unroll
          for (int i = 0; i < kLossPartials; ++i) part[i] = 0.f;
          if (rvalid) {
            const float* tp = p.target + ((size_t)b * p.P + pix) * 3;
            const float* wp = p.sw_grid ? nullptr : p.sw + (size_t)b * p.sw_bstride + (size_t)pix * 3;
            const float wg = p.sw_grid ? grid_sineweight(pix, p.grid_w, p.mask_bits) : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float t = __ldg(tp + c), w = p.sw_grid ? wg : __ldg(wp + c);
              const float er = o[c] - t;
              part[0] = fmaf(er * er, w, part[0]);
              part[1 + c] = o[c] * t;
              part[4 + c] = o[c] * o[c];
              part[7 + c] = t * t;
            }
          }
#pragma unroll
          for (int i = 0; i < kLossPartials; ++i) {
            float x = part[i];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
            part[i] = x;
          }
          if (lane == 0) {
            float* lp = p.loss_part + ((size_t)tile * 4 + q) * kLossPartials;
#pragma unroll
            for (int i = 0; i < kLossPartials; ++i) lp[i] = part[i];
          }
        }
      }
      tc_fence_before();
    }
  }

  // ---------------------------------------------------------------- teardown
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
