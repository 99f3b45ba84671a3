// Layer-major backward of the RENI decoder for sm_100a: delta chain AND weight gradients in one pass over the stash.
//
// The tile-major backward (bwd_kernel.cuh + dw_kernel.cuh) writes every delta_l to HBM in the chain kernel and then
// re-reads delta_l AND the phase stash in the weight-gradient kernel: 256 KB of HBM traffic per (tile, layer).  dW_l is
// a 256 x 256 fp32 accumulator -- all 512 TMEM columns of an SM -- so a tile-major kernel cannot keep the five of them
// next to its chain accumulators.  Here the loop order is turned around: ONE launch per hidden layer l = L..1, every
// CTA keeps HALF of that layer's dW (256 TMEM columns) for the whole launch, and each tile passes through
//
//     delta_l tile (HBM/L2 -> smem, 64 KB bulk copy)                       A operand of both GEMMs below
//     acc      = delta_l  W_l''[:, k half]        tcgen05.mma 128 x 128 x 256 (K-major A, resident weight half)
//     delta_l-1[:, k half] = acc * cos(a_l-1)     epilogue; phases a_l-1 from the forward's 16-bit stash (one read)
//     h_l-1[:, k half]     = sin(a_l-1)           same phases, same epilogue -> smem operand image
//     dW_l[:, k half]     += delta_l^T h_l-1      tcgen05.mma 2 x (128 x 128 x 128), both operands MN-major views
//     db_l[j half]        += column sums of the delta_l tile (from shared memory)
//
// so a (tile, layer) costs 192 KB of HBM traffic (delta_l in, phase in, delta_l-1 out) and one MUFU pair per element.
// Two CTAs (k halves r = 0, 1; blockIdx = 2 c + r) walk the same tile list, the second read of a delta tile hits L2.
// Launch l walks the tiles in the opposite direction to launch l+1, so the delta tiles written last are read first
// (still in L2).  The last launch (l = 1, kFirst) keeps delta_0 on chip and reduces it to the per-map layer-0
// gradients dM_b, dc_b += [f | 1]^T delta_0 with an N = 16 MMA, as the tile-major chain does.
//
// reni_lbwd_head_kernel opens the chain: g_y from the fused loss (or an external gradient), delta_L = (g_y W_out'')
// cos(a_L) on the CUDA cores (K = 3), dW_out / db_out += g_y^T h_L on the tensor core (N = 16).
//
// Reference semantics: autograd of src/models/RENI.py:63-87,132-178 and src/utils/loss_functions.py:6-32.
#pragma once
#include "dw_kernel.cuh"
#include "layout.cuh"
#include "phase.cuh"
#include "ptx.cuh"

namespace reni {

#ifndef RENI_LBWD_PF_DIST
#define RENI_LBWD_PF_DIST 0  // L2 prefetch distance, in tiles ahead of the tile whose shared-memory load is being issued
#endif
#ifndef RENI_LBWD_STAGGER
#define RENI_LBWD_STAGGER 0  // start CTA pair c after (c % 8) * RENI_LBWD_STAGGER clocks: de-synchronises the load / compute
                             // phases of the SMs (all of them otherwise hit HBM at the same moments)
#endif
#ifndef RENI_LBWD_STORE_HINT
#define RENI_LBWD_STORE_HINT 0  // delta_{l-1} stores: 0 plain st.global (write-back: the next launch reads the newest tiles from L2), 1 st.cs
#endif
constexpr int kLbwdThreads = 576;  // warp 0 producer, warp 1 MMA issuer, warps 2..17 epilogue (row x 32 columns each)
constexpr int kLbwdHalfCols = 128;
constexpr int kLbwdHImageBytes = kTileRows * kLbwdHalfCols * 2;  // 32 KB: [16 k-groups][128 rows][8]

struct LbwdParams {
  const uint16_t* stash_u;  // forward phase stash: per tile (L+1) images of two 64-row halves (stash_off)
  __half* stash_d;          // delta stash of this path: per tile (L+1) FULL 128-row tile images, slot l = delta_l
  const __half* wb2;        // backward weight images [l][k half][j/8][128 k][8] = omega_{l-1} W_l[j][k]
  float* dW;                // this layer's weight gradient (256 x 256), accumulated with reductions
  float* db;                // this layer's bias gradient (256)
  const float* scalars;     // [1] = 1 / S
  const float* D;           // kFirst: directions for the layer-0 feature columns
  int64_t d_bstride;
  float* dmc;               // kFirst: (B, 5, 256) dM_b rows 0..3, dc_b row 4 (atomics; caller zeroes)
  int P, tiles_per_map, ntiles, L, l, rev, so2;
  int grid_w, dir_grid;     // analytic directions (RENI_FLAG_GRID_DIRECTIONS): grid_point(pix, grid_w) instead of D
  unsigned long long* trace;  // debug: clock64 timeline of CTA 0 (reni_debug_set_trace), else null
};

// debug timeline: role 0 MMA issuer, 1 first epilogue warp, 2 producer; entry = code << 48 | clock
DEVINL void lbwd_trace(const LbwdParams& p, int role, uint32_t& n, uint32_t code) {
  if (p.trace != nullptr && blockIdx.x == 0 && n < 4096) {
    p.trace[role * 4096 + n] = ((unsigned long long)code << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    ++n;
  }
}

struct LbwdSmem {
  static constexpr int kX = 0;                                  // 2 x 64 KB delta_l tiles
  static constexpr int kW = 2 * kTileImageBytes;                // resident 64 KB weight half
  static constexpr int kHimg = kW + kWImageBytes / 2;           // 32 KB h_{l-1} half image (B operand of the dW GEMM);
                                                                // kFirst: alternately the delta_0 half image
  static constexpr int kBars = kHimg + kLbwdHImageBytes;
  static constexpr int kNumBars = 16;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kF = kBars + 256;                        // kFirst: [128 rows][8] fp16 rows [f0..f3, 1, 0, 0, 0]
  static constexpr int kFBytes = 2048 + 512;                    //   (the N = 16 operand's second column group aliases
                                                                //   rows 32.. of the first: its 8 result columns are unused)
  static constexpr int kTotal = kF + kFBytes;
};
static_assert(LbwdSmem::kTotal <= 232448, "layer-major backward shared memory over budget");

// One hidden layer l of the backward for all tiles.  Schedule per tile i (X = delta_l tile buffer i % 2):
//   tensor pipe : chain(i) = X W -> acc[i % 2] ; wgrad(i) = X^T H(i) -> dW   (H does not depend on chain(i): the two GEMMs
//                 are issued back to back and the tile buffer is released behind them)
//   epilogue    : H(i) = sin(phase(i)) as soon as wgrad(i-1) has read H (chain(i) runs underneath); bias sums of X(i);
//                 delta_{l-1}(i) = acc * cos(phase(i)) once chain(i) is complete (wgrad(i) runs underneath)
// The phase slice of tile i+1 is requested (registers) at the top of iteration i.
// kFirst (l = 1): one chain accumulator (16 TMEM columns go to the layer-0 reduction); delta_0(i) takes the h buffer's
// place behind wgrad(i) and is reduced against the feature rows before H(i+1) is written.
//
// Measured at cfg 2 (tools/step_phases.py, tools/trace_lbwd.py; profiles/r2_lbwd_*): this schedule 448 us for the five
// launches (+ 68 us head) against 541 us for the tile-major chain + weight-gradient GEMM.  Variants that were built,
// verified against the tile-major path and measured SLOWER, in this order: L2 prefetch of the next 1 / 2 / 3 tiles
// (448 / 468 / 491 us); one chain accumulator + bias gradient on the tensor core + phases two tiles ahead (493 us);
// sin pass one tile ahead of the cos pass (585 us); the same with delta_{l-1} leaving through the copy engine from the h
// buffer instead of st.global (695 us).  The timeline shows why none of them helps: a 64 KB tile load takes 5000..6000
// clk when all SMs stream (only one of the two buffers is ever in flight), and the 32 KB of delta stores drain at about
// 13 B/clk/SM whichever engine issues them, blocking the epilogue's next shared-memory instruction for ~2500 clk.
template <bool kFirst>
__global__ void __launch_bounds__(kLbwdThreads, 1) reni_lbwd_layer_kernel(const LbwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  constexpr int kNA = kFirst ? 1 : 2;  // chain accumulators

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LbwdSmem::kBars);
  uint64_t* w_full = bars;          // weight half landed
  uint64_t* full = bars + 1;        // [2] delta_l tile landed in X[b]
  uint64_t* empty = bars + 3;       // [2] commit: both GEMMs have read X[b]
  uint64_t* acc_full = bars + 5;    // [2] commit: chain accumulator complete
  uint64_t* acc_free = bars + 7;    // [2] epilogue warps: accumulator has been read out (16 arrivals)
  uint64_t* h_ready = bars + 9;     // epilogue warps: h image written, bias sums of the tile taken (16 arrivals)
  uint64_t* h_free = bars + 10;     // commit: the dW GEMM has read the h image
  uint64_t* d_ready = bars + 11;    // kFirst: delta_0 image + feature rows written (16 arrivals)
  uint64_t* hd_free = bars + 12;    // kFirst: commit: the layer-0 reduction has read them
  uint64_t* done = bars + 13;       // commit: everything issued has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + LbwdSmem::kTmemPtr);

  const int L = p.L, l = p.l;
  const uint32_t r = blockIdx.x & 1;             // k half of this CTA
  const int c = (int)(blockIdx.x >> 1);          // tile-list index shared by the two halves
  const int nC = (int)(gridDim.x >> 1);
  const int ntl = c < p.ntiles ? (p.ntiles - c + nC - 1) / nC : 0;
  auto tile_of = [&](int i) {
    const int idx = c + i * nC;
    return p.rev ? p.ntiles - 1 - idx : idx;
  };
  const uint8_t* stash_u = reinterpret_cast<const uint8_t*>(p.stash_u);
  uint8_t* stash_d = reinterpret_cast<uint8_t*>(p.stash_d);
  auto slot_off = [&](int tile, int layer) { return ((size_t)tile * (L + 1) + layer) * kTileImageBytes; };
  auto pslot_off = [&](int tile, int layer) { return ((size_t)tile * (L + 1) + layer) * kPhaseTileBytes; };

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 16);
    }
    mbar_init(h_ready, 16);
    mbar_init(h_free, 1);
    mbar_init(d_ready, 16);
    mbar_init(hd_free, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: chain accumulators [0,128) [128,256) (kFirst: one, and the layer-0 reduction at [128,144));
  // dW accumulator rows j < 128 at [256,384), rows j >= 128 at [384,512)
  constexpr uint32_t kColRed = 128, kColDw = 256;
  if (RENI_LBWD_STAGGER > 0) {
    const long long t0 = clock64(), wait = (long long)(c % 8) * RENI_LBWD_STAGGER;
    while (clock64() - t0 < wait) {
    }
  }

  if (warp == 0) {
    // ============================================================ producer: weight half once, one delta tile per step
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wb2) + ((size_t)(l - 1) * 2 + r) * (kWImageBytes / 2);
      mbar_arrive_expect_tx(w_full, kWImageBytes / 2);
      bulk_g2s(smem + LbwdSmem::kW, wsrc, kWImageBytes / 4, w_full);
      bulk_g2s(smem + LbwdSmem::kW + kWImageBytes / 4, wsrc + kWImageBytes / 4, kWImageBytes / 4, w_full);
      auto prefetch = [&](int i) {  // delta tile + this CTA's half of the phase tile towards L2 (build switch, off)
        if (i >= ntl) return;
        const int t = tile_of(i);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(stash_d + slot_off(t, l)), "r"(kTileImageBytes)
                     : "memory");
        const uint8_t* ph = stash_u + pslot_off(t, l - 1) + (size_t)r * 16 * (kHalfRows * kPhaseRecBytes);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ph), "r"(16 * kHalfRows * kPhaseRecBytes) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ph + kPhaseHalfBytes),
                     "r"(16 * kHalfRows * kPhaseRecBytes)
                     : "memory");
      };
      for (int d = 0; d < RENI_LBWD_PF_DIST; ++d) prefetch(d);
      uint32_t tn = 0;
      for (int i = 0; i < ntl; ++i) {
        const int b = i & 1;
        if (RENI_LBWD_PF_DIST > 0) prefetch(i + RENI_LBWD_PF_DIST);
        if (i >= 2) mbar_wait(&empty[b], (uint32_t)(i / 2 - 1) & 1u);
        lbwd_trace(p, 2, tn, 0x100 | (i & 0xff));  // load issued
        mbar_arrive_expect_tx(&full[b], kTileImageBytes);
        const uint8_t* src = stash_d + slot_off(tile_of(i), l);
        uint8_t* dst = smem + LbwdSmem::kX + b * kTileImageBytes;
        bulk_g2s(dst, src, kTileImageBytes / 2, &full[b]);
        bulk_g2s(dst + kTileImageBytes / 2, src + kTileImageBytes / 2, kTileImageBytes / 2, &full[b]);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_c = umma_idesc_f16(128, kLbwdHalfCols, 0, 0);  // chain: K-major A and B
      constexpr uint32_t idesc_w = umma_idesc_f16(128, kLbwdHalfCols, 1, 1);  // dW: MN-major A and B
      constexpr uint32_t idesc_r = umma_idesc_f16(128, kW6N, 1, 1);           // layer-0 reduction
      const uint32_t x_base = smem_u32(smem + LbwdSmem::kX);
      const uint32_t w_base = smem_u32(smem + LbwdSmem::kW);
      const uint32_t h_base = smem_u32(smem + LbwdSmem::kHimg);
      mbar_wait(w_full, 0);
      uint32_t tn = 0;
      auto chain = [&](int i) {
        const int b = i & 1, a = i % kNA;
        mbar_wait(&full[b], (uint32_t)(i / 2) & 1u);
        lbwd_trace(p, 0, tn, 0x100 | (i & 0xff));  // tile landed
        if (i >= kNA) mbar_wait(&acc_free[a], (uint32_t)(i / kNA - 1) & 1u);
        lbwd_trace(p, 0, tn, 0x200 | (i & 0xff));  // accumulator free: chain issued
        tc_fence_after();
        const uint32_t a_tile = x_base + b * kTileImageBytes;
#pragma unroll
        for (int ks = 0; ks < kH / 16; ++ks) {
          // A: [j/8][128 rows][8] -> 2048 B per 8-column group; B: [j/8][128 k][8] -> 2048 B per group
          const uint64_t da = umma_smem_desc(a_tile + ks * 4096, 2048, 128);
          const uint64_t db = umma_smem_desc(w_base + ks * 4096, 2048, 128);
          umma_f16_ss(tmem_base + a * kLbwdHalfCols, da, db, idesc_c, ks != 0);
        }
        umma_commit(&acc_full[a]);
      };
      auto wgrad = [&](int j) {
        const int b = j & 1;
        mbar_wait(h_ready, (uint32_t)j & 1u);
        lbwd_trace(p, 0, tn, 0x300 | (j & 0xff));  // h image ready: dW GEMM issued
        tc_fence_after();
        const uint32_t a_tile = x_base + b * kTileImageBytes;
#pragma unroll
        for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
          for (int ks = 0; ks < kTileRows / 16; ++ks) {
            // MN-major views of [x/8][128 rows][8] images: 8-element groups 2048 B apart (SBO), 8-row groups 128 B (LBO)
            const uint64_t da = umma_smem_desc(a_tile + mh * 16 * 2048 + ks * 256, 128, 2048);
            const uint64_t db = umma_smem_desc(h_base + ks * 256, 128, 2048);
            umma_f16_ss(tmem_base + kColDw + mh * kLbwdHalfCols, da, db, idesc_w, (j != 0) || (ks != 0));
          }
        }
        umma_commit(&empty[b]);
        umma_commit(h_free);
      };
      auto reduce0 = [&](int j) {  // D[k, i] = sum_rows delta_0[row, k] F[row, i]   (kFirst)
        mbar_wait(d_ready, (uint32_t)j & 1u);
        tc_fence_after();
        const uint32_t f_tile = smem_u32(smem + LbwdSmem::kF);
#pragma unroll
        for (int ks = 0; ks < kTileRows / 16; ++ks) {
          const uint64_t da = umma_smem_desc(h_base + ks * 256, 128, 2048);
          const uint64_t db = umma_smem_desc(f_tile + ks * 256, 128, 512);
          umma_f16_ss(tmem_base + kColRed, da, db, idesc_r, ks != 0);
        }
        umma_commit(hd_free);
      };
      for (int i = 0; i < ntl; ++i) {
        chain(i);
        if (kFirst && i >= 1) reduce0(i - 1);
        wgrad(i);
      }
      if (kFirst && ntl > 0) reduce0(ntl - 1);
      umma_commit(done);
    }
  } else {
    // ============================================================ epilogue warps: thread = (row, 32 columns)
    const uint32_t e = warp - 2;         // 0..15
    const uint32_t q = warp & 3;         // TMEM lane quarter this warp may access
    const uint32_t cq = e >> 2;          // 32-column quarter of this CTA's 128 columns
    const uint32_t row = q * 32 + lane;
    const float inv_s = __ldg(p.scalars + 1);
    float bsum[8];                       // column sums of delta_l: this warp owns j-group r*16 + e, lanes own row sets
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;

    // layer-0 reduction result of a tile -> dmc (kFirst; warps with cq == 0 hold the 128 accumulator rows)
    auto reduce_out = [&](int tile) {
      if (!kFirst || cq != 0) return;
      uint32_t v[16];
      tmem_ld16(tmem_base + ((q * 32) << 16) + kColRed, v);
      tmem_ld_wait();
      const int b = tile / p.tiles_per_map;
      float* dst = p.dmc + (size_t)b * 5 * kH + r * kLbwdHalfCols + row;
#pragma unroll
      for (int i = 0; i < 5; ++i) atomicAdd(dst + i * kH, __uint_as_float(v[i]) * inv_s);
      tc_fence_before();
    };
    auto load_phases = [&](int i, PhaseRec (&ph)[4]) {
      if (i >= ntl) return;
      const uint8_t* ph_tile = stash_u + pslot_off(tile_of(i), l - 1);
      phase_fetch4<1>(ph_tile, row, r * 4 + cq, ph);
    };

    uint32_t tn = 0;
    const bool tr = (e == 0 && lane == 0);
    // one tile: `ph` holds its phase slice, `nxt` (free now) takes the next tile's
    auto body = [&](int i, const PhaseRec (&ph)[4], PhaseRec (&nxt)[4]) {
      const int b = i & 1, a = i % kNA;
      const int tile = tile_of(i);
      if (tr) lbwd_trace(p, 1, tn, 0x100 | (i & 0xff));  // iteration start
      load_phases(i + 1, nxt);
      // (A) h_{l-1} = sin(a_{l-1}) -> operand image, once the previous tile's GEMMs have read the buffer
      if (i >= 1) {
        if (kFirst) {
          mbar_wait(hd_free, (uint32_t)(i - 1) & 1u);
          tc_fence_after();
          reduce_out(tile_of(i - 1));
        } else {
          mbar_wait(h_free, (uint32_t)(i - 1) & 1u);
        }
      }
      if (tr) lbwd_trace(p, 1, tn, 0x200 | (i & 0xff));  // h buffer free seen
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const PhaseRec& hw = ph[g];
        uint4 hv;
        hv.x = pack_half2(abl_sin(phase_angle_of<0>(hw)), abl_sin(phase_angle_of<1>(hw)));
        hv.y = pack_half2(abl_sin(phase_angle_of<2>(hw)), abl_sin(phase_angle_of<3>(hw)));
        hv.z = pack_half2(abl_sin(phase_angle_of<4>(hw)), abl_sin(phase_angle_of<5>(hw)));
        hv.w = pack_half2(abl_sin(phase_angle_of<6>(hw)), abl_sin(phase_angle_of<7>(hw)));
        *reinterpret_cast<uint4*>(smem + LbwdSmem::kHimg + tile_image_off(kTileRows, row, cq * 4 + g)) = hv;
      }
      fence_proxy_async_smem();
      // (B) bias gradient: column sums of the delta_l tile, straight from shared memory -- BEFORE this warp signals
      // h_ready: the dW GEMM behind that signal releases the tile buffer to the producer
      mbar_wait(&full[b], (uint32_t)(i / 2) & 1u);
      {
        const uint8_t* xg = smem + LbwdSmem::kX + b * kTileImageBytes + (size_t)(r * 16 + e) * (kTileRows * 16);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint4 v = *reinterpret_cast<const uint4*>(xg + (lane + 32 * s) * 16);
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
          const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
          const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
          bsum[0] += f0.x; bsum[1] += f0.y; bsum[2] += f1.x; bsum[3] += f1.y;
          bsum[4] += f2.x; bsum[5] += f2.y; bsum[6] += f3.x; bsum[7] += f3.y;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
      if (tr) lbwd_trace(p, 1, tn, 0x300 | (i & 0xff));  // h image written
      // (C) delta_{l-1} = acc * cos(a_{l-1})
      mbar_wait(&acc_full[a], (uint32_t)(i / kNA) & 1u);
      if (tr) lbwd_trace(p, 1, tn, 0x400 | (i & 0xff));  // accumulator seen
      tc_fence_after();
      uint4 dv[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((q * 32) << 16) + a * kLbwdHalfCols + cq * 32 + hf * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const int g = hf * 2 + g2;
          const PhaseRec& hw = ph[g];
          dv[g].x = pack_half2(__uint_as_float(v[g2 * 8 + 0]) * abl_cos(phase_angle_of<0>(hw)),
                               __uint_as_float(v[g2 * 8 + 1]) * abl_cos(phase_angle_of<1>(hw)));
          dv[g].y = pack_half2(__uint_as_float(v[g2 * 8 + 2]) * abl_cos(phase_angle_of<2>(hw)),
                               __uint_as_float(v[g2 * 8 + 3]) * abl_cos(phase_angle_of<3>(hw)));
          dv[g].z = pack_half2(__uint_as_float(v[g2 * 8 + 4]) * abl_cos(phase_angle_of<4>(hw)),
                               __uint_as_float(v[g2 * 8 + 5]) * abl_cos(phase_angle_of<5>(hw)));
          dv[g].w = pack_half2(__uint_as_float(v[g2 * 8 + 6]) * abl_cos(phase_angle_of<6>(hw)),
                               __uint_as_float(v[g2 * 8 + 7]) * abl_cos(phase_angle_of<7>(hw)));
          if (!kFirst) {  // delta_{l-1} leaves for the next launch: a warp writes 512 contiguous bytes per group
            uint4* dptr = reinterpret_cast<uint4*>(stash_d + slot_off(tile, l - 1) +
                                                   tile_image_off(kTileRows, row, r * 16 + cq * 4 + g));
            if (RENI_LBWD_STORE_HINT == 1) __stcs(dptr, dv[g]);
            else *dptr = dv[g];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[a]);
      if (tr) lbwd_trace(p, 1, tn, 0x500 | (i & 0xff));  // cos pass done
      if (kFirst) {
        // (D) delta_0 never leaves the SM: its image takes the h buffer's place once the dW GEMM has read h, with the
        // feature rows [f0, f1, f2, f3, 1, 0, 0, 0] (zero for rows beyond P) beside it
        mbar_wait(h_free, (uint32_t)i & 1u);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(smem + LbwdSmem::kHimg + tile_image_off(kTileRows, row, cq * 4 + g)) = dv[g];
        if (cq == 0) {
          const int bm = tile / p.tiles_per_map;
          const int pix = (tile - bm * p.tiles_per_map) * kTileRows + (int)row;
          float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, one = 0.f;
          if (pix < p.P) {
            float dx, dy, dz;
            if (p.dir_grid) {
              float sp_;
              grid_point(pix, p.grid_w, dx, dy, dz, sp_);
            } else {
              const float* d = p.D + (size_t)bm * p.d_bstride + (size_t)pix * 3;
              dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
            }
            if (p.so2) { f0 = dx; f1 = dz; f2 = sqrtf(dx * dx + dz * dz); f3 = dy; }
            else       { f0 = dx; f1 = dy; f2 = dz; }
            one = 1.f;
          }
          *reinterpret_cast<uint4*>(smem + LbwdSmem::kF + row * 16) =
              make_uint4(pack_half2(f0, f1), pack_half2(f2, f3), pack_half2(one, 0.f), 0u);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_ready);
      }
    };

    {
      PhaseRec pa[4], pb[4];
      load_phases(0, pa);
      for (int i = 0; i < ntl; i += 2) {
        body(i, pa, pb);
        if (i + 1 < ntl) body(i + 1, pb, pa);
      }
    }

    // ---- flush: every GEMM of this CTA is complete
    mbar_wait(done, 0);
    tc_fence_after();
    if (ntl > 0) {
      reduce_out(tile_of(ntl - 1));
      // bias gradient: add up the lanes' row sets, one atomic per column
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = bsum[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        bsum[i] = x;
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(p.db + (r * 16 + e) * 8 + i, bsum[i] * inv_s);
      }
      // weight gradient: warp (q, cq) flushes rows mh*128 + q*32.., columns (cq>>1)*64.. of this CTA's k half
      const uint32_t mh = cq & 1, chh = cq >> 1;
      const uint32_t j = mh * 128 + row;
      float* dst = p.dW + (size_t)j * kH + r * kLbwdHalfCols + chh * 64;
      const uint32_t t_acc = tmem_base + ((q * 32) << 16) + kColDw + mh * kLbwdHalfCols + chh * 64;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t w[32];
        tmem_ld32(t_acc + ch * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          red_add_v4(dst + ch * 32 + i, __uint_as_float(w[i]) * inv_s, __uint_as_float(w[i + 1]) * inv_s,
                     __uint_as_float(w[i + 2]) * inv_s, __uint_as_float(w[i + 3]) * inv_s);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Head of the layer-major backward: per tile
//   g_y     = S * dLoss/dy                           (fused loss or external gradient; tanh' and the sine output layer's
//                                                     cosine folded in, exactly as reni_bwd_kernel does)
//   delta_L = (g_y W_out'') * cos(a_L)               K = 3 on the CUDA cores, fp32 -> fp16 tile image in the stash
//   dW_out += g_y^T h_L, db_out += sum g_y           h_L = sin(a_L) image in shared memory, N = 16 tcgen05.mma
// warp 0: TMEM + MMA issuer; warps 1..16: workers, thread = (row, 64 columns), phase loads software-pipelined in halves.
// ------------------------------------------------------------------------------------------------
constexpr int kLbwdHeadThreads = 544;  // warp 0: TMEM + MMA issuer; warps 1..16: workers, thread = (row, 64 columns)

struct LbwdHeadParams {
  const float* out;       // (B, P, 3) forward output
  const float* grad_out;  // external gradient or null (fused loss)
  const float* aout;      // sine output layer: pre-activations, else null
  const float* target;
  const float* sw;
  int64_t sw_bstride;
  const float* map_loss;  // (B, 32): cosine-term coefficients
  const float* scalars;   // [0] = S, [1] = 1/S
  const __half* w6b;      // [c/8][256 k][8] = omega_L * s * W_out[c][k]
  const uint16_t* stash_u;
  __half* stash_d;
  float* dW_out;          // (out_features, 256)
  float* db_out;
  float out_scale;
  int P, tiles_per_map, ntiles, L, out_tanh, use_cos, out_features;
  int grid_w, sw_grid;      // analytic sine weights (RENI_FLAG_GRID_SINEWEIGHT)
  const uint32_t* mask_bits;
};

struct LbwdHeadSmem {
  static constexpr int kHimg = 0;                               // 2 x 64 KB h_L images
  static constexpr int kG = 2 * kTileImageBytes;                // 2 x 4 KB g_y tiles [2][128][8]
  static constexpr int kW6 = kG + 2 * kGyImageBytes;            // 3 x 256 floats
  static constexpr int kBars = kW6 + 3 * kH * 4;
  static constexpr int kTmemPtr = kBars + 8 * 8;
  static constexpr int kTotal = kTmemPtr + 16;
};

__global__ void __launch_bounds__(kLbwdHeadThreads, 1) reni_lbwd_head_kernel(const LbwdHeadParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LbwdHeadSmem::kBars);
  uint64_t* h_ready = bars;      // [2] workers -> MMA (16 arrivals)
  uint64_t* h_free = bars + 2;   // [2] commit: the GEMM has read buffer b
  uint64_t* done = bars + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + LbwdHeadSmem::kTmemPtr);
  float* s_w6 = reinterpret_cast<float*>(smem + LbwdHeadSmem::kW6);
  const int L = p.L;
  const int ntl = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
  const uint8_t* stash_u = reinterpret_cast<const uint8_t*>(p.stash_u);
  uint8_t* stash_d = reinterpret_cast<uint8_t*>(p.stash_d);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_ready[i], 16);
      mbar_init(&h_free[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<32>(tmem_ptr);
  for (int i = threadIdx.x; i < 3 * kH; i += kLbwdHeadThreads) {
    const int c = i / kH, k = i % kH;
    s_w6[i] = __half2float(p.w6b[(size_t)k * 8 + c]);  // (group 0 of the [c/8][k][8] image)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, kW6N, 1, 1);
      for (int i = 0; i < ntl; ++i) {
        const int b = i & 1;
        if (i + 2 < ntl)  // a later tile's phases towards L2
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                           stash_u + ((size_t)tile_of(i + 2) * (L + 1) + L) * kPhaseTileBytes),
                       "r"(kPhaseTileBytes)
                       : "memory");
        mbar_wait(&h_ready[b], (uint32_t)(i >> 1) & 1u);
        tc_fence_after();
        const uint32_t h_tile = smem_u32(smem + LbwdHeadSmem::kHimg) + b * kTileImageBytes;
        const uint32_t g_tile = smem_u32(smem + LbwdHeadSmem::kG) + b * kGyImageBytes;
#pragma unroll
        for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
          for (int ks = 0; ks < kTileRows / 16; ++ks) {
            const uint64_t da = umma_smem_desc(h_tile + mh * 16 * 2048 + ks * 256, 128, 2048);
            const uint64_t db = umma_smem_desc(g_tile + ks * 256, 128, 2048);
            umma_f16_ss(tmem_base + mh * kW6N, da, db, idesc, (i != 0) || (ks != 0));
          }
        }
        umma_commit(&h_free[b]);
      }
      umma_commit(done);
    }
  } else {
    const uint32_t e = warp - 1;        // 0..15
    const uint32_t rq = e & 3;          // row quarter
    const uint32_t cq = e >> 2;         // 64-column quarter
    const uint32_t row = rq * 32 + lane;
    const float S = __ldg(p.scalars);
    float dbo[3] = {0.f, 0.f, 0.f};
    auto load_phases = [&](int tile, int half, PhaseRec (&ph)[4]) {
      const uint8_t* ph_tile = stash_u + ((size_t)tile * (L + 1) + L) * kPhaseTileBytes;
      phase_fetch4<1>(ph_tile, row, cq * 2 + half, ph);
    };
    // per-row loss inputs of a tile -> S * dLoss/dy (zero for rows beyond P)
    auto row_gy = [&](int tile, float (&gy)[3]) {
      const int b = tile / p.tiles_per_map;
      const int pix = (tile - b * p.tiles_per_map) * kTileRows + (int)row;
      gy[0] = gy[1] = gy[2] = 0.f;
      if (pix >= p.P) return;
      const size_t eo = ((size_t)b * p.P + pix) * 3;
      if (p.grad_out != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float o = __ldg(p.out + eo + c);
          float gg = __ldg(p.grad_out + eo + c) * S;
          if (p.out_tanh) gg *= (1.f - o * o);
          if (p.aout != nullptr) gg *= cosf(__ldg(p.aout + eo + c));
          gy[c] = gg;
        }
      } else {
        const float* wp = p.sw_grid ? nullptr : p.sw + (size_t)b * p.sw_bstride + (size_t)pix * 3;
        const float wg = p.sw_grid ? grid_sineweight(pix, p.grid_w, p.mask_bits) : 0.f;
        const float* ml = p.map_loss + (size_t)b * 32;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float o = __ldg(p.out + eo + c);
          const float t = __ldg(p.target + eo + c);
          float gg = (o - t) * (p.sw_grid ? wg : __ldg(wp + c));  // S * g_o = (o - t) * sw + coefA * t + coefB * o  with S = 3P/2
          if (p.use_cos) gg += __ldg(ml + 16 + c) * t + __ldg(ml + 19 + c) * o;
          if (p.out_tanh) gg *= (1.f - o * o);
          if (p.aout != nullptr) gg *= cosf(__ldg(p.aout + eo + c));
          gy[c] = gg;
        }
      }
      // the tensor core sees g_y as fp16 (dW_out operand); the chain uses the same rounded values so that this path
      // and the tile-major chain (g_y W_out'' as an fp16 GEMM) agree to summation order
#pragma unroll
      for (int c = 0; c < 3; ++c) gy[c] = __half2float(__float2half_rn(gy[c]));
    };

    PhaseRec pa[4], pb[4];
    float gy[3], gy_next[3];
    if (ntl > 0) {
      load_phases(tile_of(0), 0, pa);
      row_gy(tile_of(0), gy_next);
    }
    for (int i = 0; i < ntl; ++i) {
      const int buf = i & 1;
      const int tile = tile_of(i);
      gy[0] = gy_next[0]; gy[1] = gy_next[1]; gy[2] = gy_next[2];
      load_phases(tile, 1, pb);
      if (cq == 0) {
        dbo[0] += gy[0];
        dbo[1] += gy[1];
        dbo[2] += gy[2];
      }
      if (i >= 2) mbar_wait(&h_free[buf], (uint32_t)((i >> 1) - 1) & 1u);
      uint8_t* h_tile = smem + LbwdHeadSmem::kHimg + buf * kTileImageBytes;
      uint8_t* d_tile = stash_d + ((size_t)tile * (L + 1) + L) * kTileImageBytes;
      if (cq == 0) {
        uint8_t* g_tile = smem + LbwdHeadSmem::kG + buf * kGyImageBytes;
        *reinterpret_cast<uint4*>(g_tile + tile_image_off(kTileRows, row, 0)) =
            make_uint4(pack_half2(gy[0], gy[1]), pack_half2(gy[2], 0.f), 0u, 0u);
        *reinterpret_cast<uint4*>(g_tile + tile_image_off(kTileRows, row, 1)) = make_uint4(0u, 0u, 0u, 0u);
      }
      auto half_pass = [&](const PhaseRec (&ph)[4], int half) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int kg = cq * 8 + half * 4 + g;
          float a[8];
          phase_decode8(ph[g], a);
          float d[8], h[8];
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const int k = kg * 8 + x;
            const float m = fmaf(gy[0], s_w6[k], fmaf(gy[1], s_w6[kH + k], gy[2] * s_w6[2 * kH + k]));
            d[x] = m * abl_cos(a[x]);
            h[x] = abl_sin(a[x]);
          }
          *reinterpret_cast<uint4*>(d_tile + tile_image_off(kTileRows, row, kg)) =
              make_uint4(pack_half2(d[0], d[1]), pack_half2(d[2], d[3]), pack_half2(d[4], d[5]), pack_half2(d[6], d[7]));
          *reinterpret_cast<uint4*>(h_tile + tile_image_off(kTileRows, row, kg)) =
              make_uint4(pack_half2(h[0], h[1]), pack_half2(h[2], h[3]), pack_half2(h[4], h[5]), pack_half2(h[6], h[7]));
        }
      };
      half_pass(pa, 0);
      if (i + 1 < ntl) {  // the next tile's first half and loss inputs are in flight through the second half
        load_phases(tile_of(i + 1), 0, pa);
        row_gy(tile_of(i + 1), gy_next);
      }
      half_pass(pb, 1);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[buf]);
    }

    mbar_wait(done, 0);
    tc_fence_after();
    if (ntl > 0) {
      const float inv_s = __ldg(p.scalars + 1) * p.out_scale;
      if (cq == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float x = dbo[c];
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
          if (lane == 0 && c < p.out_features) atomicAdd(p.db_out + c, x * inv_s);
        }
      }
      // accumulator rows k = mh*128 + lane-quarter rows; warp w may read TMEM lanes 32 * (w % 4) ..
      if (cq < 2) {
        const uint32_t q = warp & 3, mh = cq;
        uint32_t v[16];
        tmem_ld16(tmem_base + ((q * 32) << 16) + mh * kW6N, v);
        tmem_ld_wait();
        const uint32_t k = mh * 128 + q * 32 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < p.out_features) atomicAdd(p.dW_out + (size_t)c * kH + k, __uint_as_float(v[c]) * inv_s);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

}  // namespace reni
