// Data layout shared by all kernels of the RENI decoder path (sm_100a).
//
// "Tile image" = the one layout every 16-bit matrix on this path uses, in shared memory AND in HBM:
//     element (r, k) of a [R rows x K cols] fp16 block lives at byte  ((k/8) * R + r) * 16 + (k%8) * 2
// i.e. [K/8][R][8].  It is the tcgen05 SWIZZLE_NONE canonical layout twice over:
//   * as a K-major operand (contraction over k): core matrix = 8 rows x 16 B, SBO = 128 B, LBO = R*16 B
//   * as an MN-major operand (contraction over r): core matrix = 8 k-rows... of the transposed view,
//     SBO = R*16 B (next 8 columns), LBO = 128 B (next 8 rows)
// so the activations written by an epilogue (thread = row, 16 B = 8 consecutive columns, a warp writes
// 512 contiguous bytes -> conflict-free in smem, fully coalesced in HBM) feed the next layer's GEMM
// (K-major) and the weight-gradient GEMM (MN-major) without any transpose or swizzle pass.
#pragma once
#include <stdint.h>

namespace reni {

constexpr int kH = 256;            // hidden features (compile-time for the tcgen05 path)
constexpr int kTileRows = 128;     // directions per tile == UMMA M == TMEM lanes
constexpr int kMaxHiddenLayers = 6;
constexpr int kTileImageBytes = kTileRows * kH * 2;  // 65536: one [128 x 256] fp16 tile image
constexpr int kHalfRows = 64;                        // stash sub-block rows (weight-gradient GEMM K step)
constexpr int kHalfImageBytes = kHalfRows * kH * 2;  // 32768
constexpr int kWImageBytes = kH * kH * 2;            // 131072: one [256 x 256] fp16 weight image
constexpr int kWChunkK = 32;                         // K columns per streamed weight chunk
constexpr int kWChunkBytes = kH * kWChunkK * 2;      // 16384
constexpr int kChunksPerLayer = kH / kWChunkK;       // 8
constexpr int kW6N = 16;                             // final layer padded to N = 16
constexpr int kW6ImageBytes = kH * kW6N * 2;         // 8192
constexpr int kGyImageBytes = kTileRows * kW6N * 2;  // 4096 : g_y tile, [16/8][128][8] per 64-row half
constexpr int kBiasBlockBytes = 2 * 128 * 16;         // 4096: [2 k-groups][128 n][8] bias block of one N half
constexpr int kLossPartials = 10;                    // se, dot[3], oo[3], tt[3]

// byte offset of the 16-byte group (row r, columns 8*kg .. 8*kg+7) inside a [R x *] tile image
__host__ __device__ inline uint32_t tile_image_off(uint32_t R, uint32_t r, uint32_t kg) { return (kg * R + r) * 16u; }

// stash tiles are stored as two 64-row half images so the weight-gradient GEMM can stream K = 64 rows per stage
__host__ __device__ inline uint32_t stash_off(uint32_t r, uint32_t kg, uint32_t ncols) {
  return (r / kHalfRows) * (kHalfRows * ncols * 2u) + (kg * kHalfRows + (r % kHalfRows)) * 16u;
}

struct WorkspaceLayout {
  // all offsets in bytes from the workspace base; 0-size regions are absent
  int64_t wf, wb, wf2, wf2lo, wb2, wbias2, w6f, w6b, bias, mc;       // weight images + biases + per-map layer-0 (M_b, c_b)
  int64_t stash_h, stash_c, stash_d, stash_gy;   // per-tile activation / cos / delta stashes
  int64_t aout;                                  // output pre-activations (sine output layer only)
  int64_t loss_part, map_loss, dmc, scalars;     // loss partials, per-map loss coefficients, per-map dM/dc
  int64_t xc, dxc, dip;                          // per-map constant encoding columns and their gradients
  int64_t film_S, film_cs;                       // FiLM backward: per-map delta_l^T h_{l-1} and column sums of delta_l
  int64_t wf2m, wf2m_lo, wb2m, wbias2m;                   // FiLM per-map weight / bias images (RENI_FLAG_FILM_PERMAP)
  int64_t ready;                                 // overlap mode: per-tile "delta_l is out" counters (+ 1 word: stuck flag)
  int64_t total;
};

}  // namespace reni
