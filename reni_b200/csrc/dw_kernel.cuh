// Weight-gradient GEMMs of the RENI decoder for sm_100a (split-K over directions).
//
//   dW_l[j,k] = sum_r delta_l[r,j] * h_{l-1}[r,k]     l = 1..L      (256 x 256, K = all directions)
//   db_l[j]   = sum_r delta_l[r,j]
//   dW_out[c,k] = sum_r g_y[r,c] * h_L[r,k] ,  db_out[c] = sum_r g_y[r,c]
//
// delta_l / g_y are the fp16 tile images the delta-chain kernel stashed ([k/8][64 rows][8]); h_l is rebuilt here from
// the forward's phase stash (phase.cuh) (sin via MUFU on the otherwise idle CUDA cores, written straight into the operand
// image).  Read as MN-major UMMA operands the contraction runs over the rows, so no transpose is ever materialised.
// CTA i works on layer job (i mod (L+1)) and a contiguous slice of the 64-row stash blocks; the 256 x 256 fp32
// accumulator lives in TMEM (2 x 256 columns) for the whole slice and is flushed once with vector reductions.
//   warp 0 : bulk-copy producer (3-stage ring: delta block + phase block)     warp 1 : tcgen05.mma issuer
//   warps 2..9 : phase -> h operand image (in place), column sums for the bias gradient, then the TMEM -> HBM flush
#pragma once
#include "layout.cuh"
#include "phase.cuh"
#include "ptx.cuh"

namespace reni {

constexpr int kDwThreads = 320;
constexpr int kDwStages = 3;
constexpr int kPhaseLand = kHalfImageBytes - kPhaseHalfBytes;  // where a phase block lands inside its operand region
constexpr int kDwStageBytes = 2 * kHalfImageBytes;  // 64 KB: A operand + B operand; the phase block lands in the
                                                    // operand slot of h and is converted in place

#ifndef RENI_DW_INTERLEAVE
#define RENI_DW_INTERLEAVE 1  // 1: a CTA walks blocks slice, slice+nslices, ... (ascending) instead of one contiguous range: the
                              // delta chain walks the units downwards, so every CTA starts on the tiles it wrote last (L2)
#endif
#ifndef RENI_DW_LOAD_HINT
#define RENI_DW_LOAD_HINT 1  // stash blocks loaded with an L2 evict-first policy (258 -> 252 us at cfg 2)
#endif

struct DwParams {
  const uint16_t* stash_u;
  const __half* stash_d;
  const __half* stash_gy;
  float* dW[kMaxHiddenLayers + 2];  // [1..L] hidden (256x256), [L+1] output (out_features x 256); [0] unused
  float* db[kMaxHiddenLayers + 2];
  const float* scalars;  // [1] = 1 / S
  float out_scale;       // omega of a sine output layer (its pre-activation is omega * (W h + b)), else 1
  int ntiles, L, out_features;
  // FiLM mode (film_S != null): grid.y = maps; the CTAs of row b reduce over the tiles of map b only and the hidden
  // jobs accumulate into per-map buffers -- S[b][l-1] = delta_l^T h_{l-1} (256 x 256) and cs[b][l-1] = column sums of
  // delta_l -- from which reni_film_reduce_kernel forms dW_l, db_l, dfreq_l[b], dphase_l[b].  The output-layer job
  // (not modulated) still accumulates straight into dW[L+1] / db[L+1]; njobs = L drops it (frozen decoder).
  float* film_S;
  float* film_cs;
  int tiles_per_map, njobs;
  int out_ctas;  // > 0: CTAs given to the output-layer job (non-FiLM grid), 0: plain round robin
  // Overlap mode (ready != null): this grid runs BESIDE the delta-chain kernel on the SMs that kernel leaves free.  The
  // chain bumps ready[tile] by one per epilogue warp (8 per tile) each time a delta_l of the tile is complete in global
  // memory (l = L..1; g_y is behind the first bump), so hidden job l may pull the tile's blocks once
  // ready[tile] >= 8 * (L - l + 1) and the output job once it is >= 8.  Blocks are then taken in DESCENDING order,
  // interleaved over the job's slices -- the order the chain finishes them -- and arrive from L2 instead of HBM.
  // FiLM, persistent mode (film_persistent = 1, grid = (SMs, 1)): the B * njobs * (2 tiles_per_map) stash blocks of all
  // (map, job) items form one list that the CTAs share out evenly in contiguous ranges; a CTA flushes its accumulator
  // whenever its range crosses into the next item (about two flushes per CTA instead of one per (map, job, slice) CTA,
  // and no partially filled last round).
  int film_persistent, B;
  const uint32_t* ready;
  uint32_t* stuck;  // set to 1 if a wait ran into its poll limit (the result is then wrong; the host checks in tests)
};

struct DwSmem {
  static constexpr int kRing = 0;
  static constexpr int kBars = kRing + kDwStages * kDwStageBytes;
  static constexpr int kNumBars = 3 * kDwStages + 2;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemPtr + 16;
};
static_assert(DwSmem::kTotal <= 232448, "dW kernel shared memory over budget");

DEVINL void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kDwThreads, 1) reni_dw_kernel(const DwParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DwSmem::kBars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDwStages;
  uint64_t* conv = bars + 2 * kDwStages;  // converters -> MMA: operand images ready
  uint64_t* done = bars + 3 * kDwStages;
  uint64_t* flushed = done + 1;  // converters -> MMA: the accumulator has been read out (segment boundary)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + DwSmem::kTmemPtr);

  const int L = p.L;
  const int njobs = p.njobs;
  const bool film = p.film_S != nullptr;
  // CTA -> (job, slice).  Round robin over the jobs, except that the output-layer job (whose stash blocks are half
  // the bytes and whose MMAs are N = 16) only gets p.out_ctas CTAs when that is set: the hidden jobs, which carry the
  // HBM traffic, then share the remaining CTAs evenly.
  int job, slice, nslices;
  if (p.out_ctas > 0 && njobs == L + 1) {
    if ((int)blockIdx.x < p.out_ctas) {
      job = L; slice = blockIdx.x; nslices = p.out_ctas;
    } else {
      const int i = (int)blockIdx.x - p.out_ctas, nh = (int)gridDim.x - p.out_ctas;
      job = i % L; slice = i / L; nslices = (nh - 1 - job) / L + 1;
    }
  } else {
    job = blockIdx.x % njobs;   // 0..L-1 -> hidden layer job+1 ; L -> output layer
    slice = blockIdx.x / njobs;
    nslices = ((int)gridDim.x - 1 - job) / njobs + 1;
  }
  const int total = (film ? p.tiles_per_map : p.ntiles) * 2;  // 64-row stash blocks this grid row reduces over
  const int s_base = film ? (int)blockIdx.y * p.tiles_per_map * 2 : 0;
  const bool overlap = p.ready != nullptr;
  const int s_begin = s_base + (int)((int64_t)slice * total / nslices);
  const int s_end = s_base + (int)((int64_t)(slice + 1) * total / nslices);
  const bool inter = RENI_DW_INTERLEAVE && !overlap && !p.film_persistent;
  const int nst = (overlap || inter) ? (total > slice ? (total - slice + nslices - 1) / nslices : 0) : s_end - s_begin;
  // i-th stash block of this CTA
  auto blk = [&](int i) { return overlap ? total - 1 - (slice + i * nslices) : s_begin + i; };
  // Work of this CTA as a sequence of segments (job, map, first block, blocks): ONE for the grids above, one per
  // (map, job) item touched by the CTA's range in the persistent FiLM mode.
  struct Seg { int job, map, s0, n; };
  const int nblk = p.tiles_per_map * 2;
  const int64_t g_total = (int64_t)p.B * njobs * nblk;
  const int64_t g_begin = p.film_persistent ? (int64_t)blockIdx.x * g_total / gridDim.x : 0;
  const int64_t g_end = p.film_persistent ? (int64_t)(blockIdx.x + 1) * g_total / gridDim.x : 1;
  auto next_seg = [&](int64_t& g, Seg& sg) -> bool {
    if (g >= g_end) return false;
    if (!p.film_persistent) {
      sg = Seg{job, film ? (int)blockIdx.y : 0, s_begin, nst};
      g = g_end;
      return true;
    }
    const int64_t item = g / nblk;
    const int64_t stop = (item + 1) * nblk < g_end ? (item + 1) * nblk : g_end;
    sg.map = (int)(item / njobs);
    sg.job = (int)(item % njobs);
    sg.s0 = sg.map * nblk + (int)(g - item * nblk);
    sg.n = (int)(stop - g);
    g = stop;
    return true;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kDwStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);  // tcgen05.commit
      mbar_init(&conv[i], 8);   // one arrival per converter warp
    }
    mbar_init(done, 1);
    mbar_init(flushed, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // sources for stash block s (tile = s/2, half = s%2): the fp16 image that is used as it is, and the phase block
  // from which the h operand is rebuilt
  auto img_src = [&](int s, int job_) -> const uint8_t* {
    const int tile = s >> 1, half = s & 1;
    const bool is_out = job_ == L;
    const int layer = job_ + 1;
    if (is_out)
      return reinterpret_cast<const uint8_t*>(p.stash_gy) + (size_t)tile * kGyImageBytes +
             (size_t)half * (kHalfRows * kW6N * 2);
    return reinterpret_cast<const uint8_t*>(p.stash_d) + ((size_t)tile * (L + 1) + layer) * kTileImageBytes +
           (size_t)half * kHalfImageBytes;
  };
  auto phase_src = [&](int s, int job_) -> const uint8_t* {
    const int tile = s >> 1, half = s & 1;
    const int lh = job_ == L ? L : job_;  // h_L for the output layer, h_{l-1} for hidden layer l = job + 1
    return reinterpret_cast<const uint8_t*>(p.stash_u) + ((size_t)tile * (L + 1) + lh) * kPhaseTileBytes +
           (size_t)half * kPhaseHalfBytes;
  };
  // hidden job: A = delta_l (copied), B = h_{l-1} (converted).  output job: A = h_L (converted), B = g_y (copied)
  auto img_off_of = [&](int job_) -> uint32_t { return job_ == L ? kHalfImageBytes : 0; };
  auto cvt_off_of = [&](int job_) -> uint32_t { return job_ == L ? 0 : kHalfImageBytes; };
  auto img_bytes_of = [&](int job_) -> uint32_t { return job_ == L ? (kHalfRows * kW6N * 2) : kHalfImageBytes; };

  // block i of a segment (overlap mode has a single segment and walks its interleaved, descending order)
  auto seg_blk = [&](const Seg& sg, int i) {
    return overlap ? blk(i) : (inter ? s_base + slice + i * nslices : sg.s0 + i);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      int tile_ok = -1;
      int64_t g = g_begin;
      Seg sg;
      while (next_seg(g, sg)) {
        const bool is_out = sg.job == L;
        const uint32_t need = is_out ? 8u : 8u * (uint32_t)(L - sg.job);
        const uint32_t img_off = img_off_of(sg.job), cvt_off = cvt_off_of(sg.job), img_bytes = img_bytes_of(sg.job);
        for (int i = 0; i < sg.n; ++i) {
          const int s = seg_blk(sg, i);
          if (overlap && (s >> 1) != tile_ok) {
            const uint32_t* ctr = p.ready + (s >> 1);
            uint32_t v, spins = 0;
            for (;;) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
              if (v >= need) break;
              if (++spins > (1u << 22)) {  // ~ seconds: never hang the device on a scheduling surprise
                *p.stuck = 1u;
                break;
              }
              __nanosleep(200);
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies below read what was just acquired
            tile_ok = s >> 1;
          }
          mbar_wait(&empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&full[st], img_bytes + kPhaseHalfBytes);
          uint8_t* dst = smem + DwSmem::kRing + st * kDwStageBytes;
#if RENI_DW_LOAD_HINT  // both stash blocks are read exactly once in this kernel: do not keep them in L2
          bulk_g2s_stream(dst + img_off, img_src(s, sg.job), img_bytes, &full[st]);
          if (kPhaseLand == 0) {
            bulk_g2s_stream(dst + cvt_off, phase_src(s, sg.job), kPhaseHalfBytes, &full[st]);
          } else {
            // 12-bit phases: the quarter of the block that converter column set c turns into its 8 KB of the operand
            // image lands at the END of those 8 KB, so a converter thread only ever overwrites its own input
            for (int c = 0; c < 4; ++c)
              bulk_g2s_stream(dst + cvt_off + c * (kHalfImageBytes / 4) + kPhaseLand / 4,
                              phase_src(s, sg.job) + c * (kPhaseHalfBytes / 4), kPhaseHalfBytes / 4, &full[st]);
          }
#else
          bulk_g2s(dst + img_off, img_src(s, sg.job), img_bytes, &full[st]);
          if (kPhaseLand == 0) {
            bulk_g2s(dst + cvt_off, phase_src(s, sg.job), kPhaseHalfBytes, &full[st]);
          } else {
            for (int c = 0; c < 4; ++c)
              bulk_g2s(dst + cvt_off + c * (kHalfImageBytes / 4) + kPhaseLand / 4,
                       phase_src(s, sg.job) + c * (kPhaseHalfBytes / 4), kPhaseHalfBytes / 4, &full[st]);
          }
#endif
          if (++st == kDwStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t ring_base = smem_u32(smem + DwSmem::kRing);
      uint32_t st = 0, ph = 0;
      int64_t g = g_begin;
      Seg sg;
      int seg_idx = 0;
      while (next_seg(g, sg)) {
        const uint32_t idesc = sg.job == L ? umma_idesc_f16(128, kW6N, 1, 1) : umma_idesc_f16(128, 256, 1, 1);
        if (seg_idx > 0) {  // the previous segment's accumulator has been read out
          mbar_wait(flushed, (uint32_t)(seg_idx - 1) & 1u);
          tc_fence_after();
        }
        for (int i = 0; i < sg.n; ++i) {
          mbar_wait(&conv[st], ph);
          tc_fence_after();
          const uint32_t sa = ring_base + st * kDwStageBytes;
          const uint32_t sb = sa + kHalfImageBytes;
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
            for (int ks = 0; ks < kHalfRows / 16; ++ks) {
              // MN-major operands: 8-column groups are 64 rows x 16 B = 1024 B apart (SBO), 8-row groups 128 B (LBO)
              const uint64_t da = umma_smem_desc(sa + mh * 16 * 1024 + ks * 256, 128, 1024);
              const uint64_t db = umma_smem_desc(sb + ks * 256, 128, 1024);
              umma_f16_ss(tmem_base + mh * 256, da, db, idesc, (i != 0) || (ks != 0));
            }
          }
          umma_commit(&empty[st]);
          if (++st == kDwStages) { st = 0; ph ^= 1; }
        }
        umma_commit(done);
        ++seg_idx;
      }
    }
  } else {
    // ---- converters; bias-gradient column sums straight from the smem operand: thread t owns the 8 columns of group
    // jg = t >> 3 and every 8th row (rset = t & 7), so a quarter-warp reads 128 contiguous bytes (conflict-free) and
    // the running sums take 8 registers (a per-row layout took 64 and kept this kernel at 157 registers, which left no
    // room on the SM for the small kernels that are forked beside it)
    const uint32_t t = threadIdx.x - 64;  // 0..255
    const uint32_t r = t & 63;            // conversion: row inside the 64-row block
    const uint32_t kset = t >> 6;         // conversion: 8-column groups kset*8 .. kset*8+7
    const uint32_t jg = t >> 3, rset = t & 7;
    uint32_t st = 0, ph = 0;
    int64_t g = g_begin;
    Seg sg;
    int seg_idx = 0;
    while (next_seg(g, sg)) {
      const bool is_out = sg.job == L;
      const int layer = sg.job + 1;
      const uint32_t img_off = img_off_of(sg.job), cvt_off = cvt_off_of(sg.job);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int i = 0; i < sg.n; ++i) {
        mbar_wait(&full[st], ph);
        uint8_t* stage = smem + DwSmem::kRing + st * kDwStageBytes;
        // (1) phase block -> h = sin(angle) operand image [k/8][64][8].  16-bit phases: in place, group by group.
        // 12-bit phases: this thread's two 32-column groups sit at the end of its column set's 8 KB of the operand region
        // (see the producer); it takes them into registers, then writes its eight 16-byte pieces over that region.
        {
          uint8_t* hd = stage + cvt_off;
          PhaseRec u[2][4];
          if (kPhaseLand != 0) {
            const uint8_t* us = hd + kset * (kHalfImageBytes / 4) + kPhaseLand / 4;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) phase_fetch4_smem(us, r, kk, u[kk]);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t off = ((kset * 8 + kk) * kHalfRows + r) * 16;
            if (kPhaseLand == 0 && (kk & 3) == 0) phase_fetch4_smem(hd, r, kset * 2 + (kk >> 2), u[kk >> 2]);
            const PhaseRec& uw = u[kk >> 2][kk & 3];
            uint4 h;
            h.x = pack_half2(abl_sin(phase_angle_of<0>(uw)), abl_sin(phase_angle_of<1>(uw)));
            h.y = pack_half2(abl_sin(phase_angle_of<2>(uw)), abl_sin(phase_angle_of<3>(uw)));
            h.z = pack_half2(abl_sin(phase_angle_of<4>(uw)), abl_sin(phase_angle_of<5>(uw)));
            h.w = pack_half2(abl_sin(phase_angle_of<6>(uw)), abl_sin(phase_angle_of<7>(uw)));
            *reinterpret_cast<uint4*>(hd + off) = h;
          }
        }
        // (2) bias-gradient column sums from the delta / g_y image
        const uint8_t* base = stage + img_off;
        if (!is_out || jg == 0) {  // (g_y has 16 padded columns: only group 0 carries data)
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + (jg * kHalfRows + rset + 8 * q8) * 16);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
            const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
            acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
            acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[st]);
        if (++st == kDwStages) { st = 0; ph ^= 1; }
      }

      // ---- flush: all MMAs of the segment done -> TMEM holds its dW
      mbar_wait(done, (uint32_t)seg_idx & 1u);
      tc_fence_after();
      const float inv_s = __ldg(p.scalars + 1) * (is_out ? p.out_scale : 1.f);
      float* dW_dst = p.dW[layer];
      float* db_dst = p.db[layer];
      if (film && !is_out) {
        dW_dst = p.film_S + ((size_t)sg.map * L + sg.job) * kH * kH;
        db_dst = p.film_cs + ((size_t)sg.map * L + sg.job) * kH;
      }
      if (sg.n > 0) {
        // column sums: add up the 8 row subsets (lanes that differ in their low 3 bits), one atomic per column
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = acc[i];
          x += __shfl_xor_sync(0xffffffffu, x, 1);
          x += __shfl_xor_sync(0xffffffffu, x, 2);
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          acc[i] = x;
        }
        if (rset == 0) {
          const int ncol = is_out ? p.out_features : kH;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if ((int)(jg * 8 + i) < ncol && (!is_out || jg == 0)) atomicAdd(db_dst + jg * 8 + i, acc[i] * inv_s);
        }
        const uint32_t q = warp & 3;
        const uint32_t mh = (warp - 2) >> 2;
        const uint32_t j = mh * 128 + q * 32 + lane;  // accumulator row
        const uint32_t t_acc = tmem_base + ((q * 32) << 16) + mh * 256;
        if (!is_out) {
          float* dst = dW_dst + (size_t)j * kH;
#pragma unroll 1
          for (int ch = 0; ch < kH / 32; ++ch) {
            uint32_t v[32];
            tmem_ld32(t_acc + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              red_add_v4(dst + ch * 32 + i, __uint_as_float(v[i]) * inv_s, __uint_as_float(v[i + 1]) * inv_s,
                         __uint_as_float(v[i + 2]) * inv_s, __uint_as_float(v[i + 3]) * inv_s);
          }
        } else {
          uint32_t v[16];
          tmem_ld16(t_acc, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < p.out_features) atomicAdd(dW_dst + (size_t)c * kH + j, __uint_as_float(v[c]) * inv_s);
        }
      }
      // the accumulator may be overwritten by the next segment's first MMA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(flushed);
      ++seg_idx;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace reni
