// FiLM per-map stage as ONE cooperative launch per direction (RENI.py:481-512 mapping network, :515-524 first FiLM layer).
//
// EXPERIMENT, off by default (RENI_FILM_MAP_FUSED in abi.cu).  The staged version (small_kernels.cuh: 2 + n_linears
// launches forward, 3 + 2 n_linears backward) takes 47 us forward and ~100 us backward at 32 maps inside a replayed step
// graph for ~0.1 GFLOP (tools/film_map_stage_time.py); this one measured 60 / ~130 us: per stage (globaltimer stamps,
// RENI_MAP_FUSED_STAMPS) input 2.1, linears 18.4 / 5.1 / 4.7 / 16.4, finish 4.4 us -- a k-step of 16 loads per lane costs
// ~0.8 us of L2 latency with 8 warps per SM, exactly as in the staged kernels, and a barrier costs what a launch did.
// Here every stage is a grid-stride loop over warp-sized work units and the stages are
// separated by a grid-wide barrier (one counter in the caller's scratch, zeroed by the launcher; the launch is
// cooperative, so all CTAs are resident).  Activations cross the barrier through L2 (ld.global.cg / plain stores +
// __threadfence), weights and inputs are read-only for the whole launch (__ldg).
#pragma once
#include "small_kernels.cuh"

namespace reni {

#ifndef RENI_MAP_FUSED_STAMPS
#define RENI_MAP_FUSED_STAMPS 0  // debug: globaltimer of block 0 behind every barrier, in the sync words (tools/film_map_stage_time.py)
#endif
constexpr int kMapFusedThreads = 256;
constexpr int kMapFusedMaxCtas = 128;

// All CTAs of the (cooperative) grid: arrive, then wait until `target` arrivals have been counted.  Bounded spin: a
// launch that is not cooperative after all must not hang the device -- it flags *stuck and traps.
DEVINL void map_grid_sync(unsigned int* ctr, unsigned int target, unsigned int* stuck) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 24)) {  // (seconds) flag it and fail the launch rather than return garbage
        *stuck = 1u;
        __threadfence_system();
        asm volatile("trap;");
      }
    }
  }
  __syncthreads();
#if RENI_MAP_FUSED_STAMPS
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(ctr)[2 + target / gridDim.x] = t;
  }
#endif
}

struct FilmMapFusedParams {
  const float* Z;    // (B, N, 3)
  const float* W0;   // first FiLM layer (256, in0)
  const float* b0;
  const float* W[kFilmMapMaxLinears];
  const float* bias[kFilmMapMaxLinears];
  float* act[kFilmMapMaxLinears + 1];  // act[0] mapping input ... act[n] raw output
  int dims[kFilmMapMaxLinears + 1];
  int n_linears;
  float* mc;
  float* film;
  int B, N, so2, Lf;
  unsigned int* sync;  // [0] barrier counter (zero at launch), [1] stuck flag
};

// y[b, o] = act(bias[o] + sum_k W[o, k] x[b, k]) for a 4-row x 4-map unit per warp: lanes stride k (coalesced weight
// rows, every weight read once for four maps), 16 partial sums per lane, butterfly reduction
DEVINL void map_linear_unit(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                            float* __restrict__ y, int B, int in, int out, int leaky, int o0, int b0, int lane) {
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[r][m] = 0.f;
  const float* w[4];
  const float* xs[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = W + (size_t)min(o0 + r, out - 1) * in;
#pragma unroll
  for (int m = 0; m < 4; ++m) xs[m] = x + (size_t)min(b0 + m, B - 1) * in;
  int k = lane;
  for (; k + 32 < in; k += 64) {
    float a[4], c[4], u[4], v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { a[r] = __ldg(w[r] + k); c[r] = __ldg(w[r] + k + 32); }
#pragma unroll
    for (int m = 0; m < 4; ++m) { u[m] = __ldcg(xs[m] + k); v[m] = __ldcg(xs[m] + k + 32); }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][m] = fmaf(c[r], v[m], fmaf(a[r], u[m], acc[r][m]));
  }
  for (; k < in; k += 32) {
    float a[4], u[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = __ldg(w[r] + k);
#pragma unroll
    for (int m = 0; m < 4; ++m) u[m] = __ldcg(xs[m] + k);
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][m] = fmaf(a[r], u[m], acc[r][m]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float s = acc[r][m];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
      acc[r][m] = s;
    }
  if (lane < 16) {  // lane -> (row, map) of the unit
    const int r = lane >> 2, m = lane & 3;
    float s = 0.f;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int mm = 0; mm < 4; ++mm)
        if (rr == r && mm == m) s = acc[rr][mm];
    if (o0 + r < out && b0 + m < B) {
      s += __ldg(bias + o0 + r);
      y[(size_t)(b0 + m) * out + o0 + r] = (leaky && s < 0.f) ? 0.2f * s : s;
    }
  }
}

__global__ void __launch_bounds__(kMapFusedThreads) reni_film_map_fused_fwd_kernel(const FilmMapFusedParams p) {
  const int lane = threadIdx.x & 31;
  const int gwarp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x), nthreads = (int)(gridDim.x * blockDim.x);
  const int N = p.N, B = p.B;
  unsigned int epoch = 0;
#if RENI_MAP_FUSED_STAMPS
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.sync)[2] = t;
  }
#endif

  // ---- mapping input per map (RENI.py:424-435 SO2: [vec(Z_xz Z_xz^T), Z_y]; :407-415 SO3: vec(Z Z^T))
  {
    const int mn_in = p.dims[0];
    for (int idx = gtid; idx < B * mn_in; idx += nthreads) {
      const int b = idx / mn_in, i = idx - b * mn_in;
      const float* z = p.Z + (size_t)b * 3 * N;
      float v;
      if (i < N * N) {
        const int n = i / N, m = i - n * N;
        v = __ldg(z + n * 3) * __ldg(z + m * 3) + __ldg(z + n * 3 + 2) * __ldg(z + m * 3 + 2);
        if (!p.so2) v = fmaf(__ldg(z + n * 3 + 1), __ldg(z + m * 3 + 1), v);
      } else {
        v = __ldg(z + (i - N * N) * 3 + 1);
      }
      p.act[0][idx] = v;
    }
  }
  map_grid_sync(p.sync, ++epoch * gridDim.x, p.sync + 1);

  // ---- the mapping network, LeakyReLU(0.2) between the linears
  for (int i = 0; i < p.n_linears; ++i) {
    const int in = p.dims[i], out = p.dims[i + 1];
    const int nrb = (out + 3) >> 2, nmb = (B + 3) >> 2;
    for (int u = gwarp; u < nrb * nmb; u += nwarps) {  // consecutive warps: same maps, neighbouring rows
      const int mb = u / nrb, rb = u - mb * nrb;
      map_linear_unit(p.act[i], p.W[i], p.bias[i], p.act[i + 1], B, in, out, i + 1 < p.n_linears, rb * 4, mb * 4, lane);
    }
    map_grid_sync(p.sync, ++epoch * gridDim.x, p.sync + 1);
  }

  // ---- freq = 15 raw + 30 (RENI.py:667), film for the hidden layers, hoisted + modulated first layer
  {
    const int half = p.Lf * kH;
    const int in0 = p.so2 ? N + 2 : N;
    for (int idx = gtid; idx < B * kH; idx += nthreads) {
      const int b = idx / kH, j = idx - b * kH;
      const float* raw = p.act[p.n_linears] + (size_t)b * 2 * half;
      const float* z = p.Z + (size_t)b * 3 * N;
      for (int l = 1; l < p.Lf; ++l) {
        float* o = p.film + ((size_t)b * (p.Lf - 1) + (l - 1)) * 2 * kH;
        o[j] = fmaf(__ldcg(raw + l * kH + j), 15.f, 30.f);
        o[kH + j] = __ldcg(raw + half + l * kH + j);
      }
      const float f0 = fmaf(__ldcg(raw + j), 15.f, 30.f), ph0 = __ldcg(raw + half + j);
      const float* w = p.W0 + (size_t)j * in0;
      float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
      if (p.so2) {
        for (int n = 0; n < N; ++n) {
          const float wv = __ldg(w + 2 + n);
          m0 = fmaf(wv, __ldg(z + n * 3), m0);      // d_x
          m1 = fmaf(wv, __ldg(z + n * 3 + 2), m1);  // d_z
        }
        m2 = __ldg(w);      // |d_xz|
        m3 = __ldg(w + 1);  // d_y
      } else {
        for (int n = 0; n < N; ++n) {
          const float wv = __ldg(w + n);
          m0 = fmaf(wv, __ldg(z + n * 3), m0);
          m1 = fmaf(wv, __ldg(z + n * 3 + 1), m1);
          m2 = fmaf(wv, __ldg(z + n * 3 + 2), m2);
        }
      }
      float* o = p.mc + (size_t)b * 5 * kH + j;
      o[0] = f0 * m0;
      o[kH] = f0 * m1;
      o[2 * kH] = f0 * m2;
      o[3 * kH] = f0 * m3;
      o[4 * kH] = fmaf(f0, __ldg(p.b0 + j), ph0);
    }
  }
#if RENI_MAP_FUSED_STAMPS
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.sync)[2 + epoch + 1] = t;
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// Backward of the per-map stage in one cooperative launch (the staged version and the derivation: small_kernels.cuh,
// "FiLM per-map stage, BACKWARD").  Stages, a grid barrier between them:
//   head | linear n-1: dX units + dW units + the first layer's dW_0 units | ... | linear 0: dX + dW units | dz
// dX and dW of one linear only read dpre = dY * act'(y), so they share a stage; every unit is a CTA-sized piece of the
// staged kernels' grids (dX: 256 input columns x 16 output rows x 32 maps, slices added with atomics into the zeroed
// dX; dW: 256 columns x 8 rows over all maps, each element owned by one unit).
// ------------------------------------------------------------------------------------------------
struct FilmMapFusedBwdParams {
  const float* Z;
  const float* W0;
  const float* b0;
  const float* W[kFilmMapMaxLinears];
  const float* act[kFilmMapMaxLinears + 1];  // saved by the forward
  float* dact[kFilmMapMaxLinears + 1];       // gradients w.r.t. act[i]; [0..n-1] zeroed by the launcher
  int dims[kFilmMapMaxLinears + 1];
  int n_linears;
  const float* d_mc;
  const float* d_film;
  float* dM;   // (B, 4, 256) scratch
  float* dZ;   // written
  float* dW0;  // accumulated (null: a frozen decoder only wants dZ)
  float* db0;
  float* dW[kFilmMapMaxLinears];
  float* db[kFilmMapMaxLinears];
  int B, N, so2, Lf;
  unsigned int* sync;
};

constexpr int kMapFusedDxRows = 16;

__global__ void __launch_bounds__(kMapFusedThreads) reni_film_map_fused_bwd_kernel(const FilmMapFusedBwdParams p) {
  __shared__ __align__(16) float s_g[32 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x), nthreads = (int)(gridDim.x * blockDim.x);
  const int gwarp = gtid >> 5, nwarps = nthreads >> 5;
  const int N = p.N, B = p.B, n = p.n_linears;
  const int half = p.Lf * kH;
  const int in0 = p.so2 ? N + 2 : N;
  const bool want_dw = p.dW0 != nullptr;
  unsigned int epoch = 0;

  // ---- head: d_raw = [15 d_freq | d_phase], dM[r] = d_mc[r] freq_0
  for (int idx = gtid; idx < B * kH; idx += nthreads) {
    const int b = idx / kH, j = idx - b * kH;
    const float* raw = p.act[n] + (size_t)b * 2 * half;
    const float* z = p.Z + (size_t)b * 3 * N;
    float* d_raw = p.dact[n] + (size_t)b * 2 * half;
    for (int l = 1; l < p.Lf; ++l) {
      const float* g = p.d_film + ((size_t)b * (p.Lf - 1) + (l - 1)) * 2 * kH;
      d_raw[l * kH + j] = 15.f * __ldg(g + j);
      d_raw[half + l * kH + j] = __ldg(g + kH + j);
    }
    const float f0 = fmaf(__ldg(raw + j), 15.f, 30.f);
    const float* w = p.W0 + (size_t)j * in0;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    if (p.so2) {
      for (int q = 0; q < N; ++q) {
        const float wv = __ldg(w + 2 + q);
        m0 = fmaf(wv, __ldg(z + q * 3), m0);
        m1 = fmaf(wv, __ldg(z + q * 3 + 2), m1);
      }
      m2 = __ldg(w);
      m3 = __ldg(w + 1);
    } else {
      for (int q = 0; q < N; ++q) {
        const float wv = __ldg(w + q);
        m0 = fmaf(wv, __ldg(z + q * 3), m0);
        m1 = fmaf(wv, __ldg(z + q * 3 + 1), m1);
        m2 = fmaf(wv, __ldg(z + q * 3 + 2), m2);
      }
    }
    const float* g = p.d_mc + (size_t)b * 5 * kH + j;
    const float g0 = __ldg(g), g1 = __ldg(g + kH), g2 = __ldg(g + 2 * kH), g3 = __ldg(g + 3 * kH), g4 = __ldg(g + 4 * kH);
    d_raw[j] = 15.f * fmaf(g0, m0, fmaf(g1, m1, fmaf(g2, m2, fmaf(g3, m3, g4 * __ldg(p.b0 + j)))));
    d_raw[half + j] = g4;
    float* dM = p.dM + (size_t)b * 4 * kH + j;
    dM[0] = g0 * f0;
    dM[kH] = g1 * f0;
    dM[2 * kH] = g2 * f0;
    dM[3 * kH] = g3 * f0;
  }
  map_grid_sync(p.sync, ++epoch * gridDim.x, p.sync + 1);

  // ---- the linears, last to first
  for (int i = n - 1; i >= 0; --i) {
    const int in = p.dims[i], out = p.dims[i + 1];
    const int leaky = i + 1 < n ? 1 : 0;
    const float* dY = p.dact[i + 1];
    const float* y = p.act[i + 1];
    const int nkb = (in + kMapFusedThreads - 1) / kMapFusedThreads;
    const int nos = (out + kMapFusedDxRows - 1) / kMapFusedDxRows, nmb = (B + 31) / 32;
    const int n_dx = nkb * nos * nmb;
    const int nrb = (out + kMapBwdRows - 1) / kMapBwdRows;
    const int n_dw = want_dw ? nkb * nrb : 0;
    const int n_w0 = (want_dw && i == n - 1) ? kH / 8 : 0;
    for (int u = blockIdx.x; u < n_dx + n_dw + n_w0; u += gridDim.x) {
      __syncthreads();  // (the previous unit is done with s_g)
      if (u < n_dx) {
        // dX[b, k] += sum_{o in slice} dpre[b, o] W[o, k]: thread = input column k, 32 maps in registers
        const int kb = u % nkb, os = (u / nkb) % nos, mb = u / (nkb * nos);
        const int o0 = os * kMapFusedDxRows, b0 = mb * 32, k = kb * kMapFusedThreads + (int)threadIdx.x;
        for (int t = threadIdx.x; t < kMapFusedDxRows * 32; t += kMapFusedThreads) {
          const int bb = t / kMapFusedDxRows, o = t % kMapFusedDxRows;  // consecutive threads: consecutive o of one map
          float g = 0.f;
          if (b0 + bb < B && o0 + o < out) {
            g = __ldcg(dY + (size_t)(b0 + bb) * out + o0 + o);
            if (leaky && __ldg(y + (size_t)(b0 + bb) * out + o0 + o) < 0.f) g *= 0.2f;
          }
          s_g[o * 32 + bb] = g;
        }
        __syncthreads();
        if (k < in) {
          float acc[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) acc[t] = 0.f;
          const int no = min(kMapFusedDxRows, out - o0);
          for (int oc = 0; oc < no; oc += 8) {
            float w[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w[t] = (oc + t < no) ? __ldg(p.W[i] + (size_t)(o0 + oc + t) * in + k) : 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float4* g4 = reinterpret_cast<const float4*>(s_g + (oc + t) * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 g = g4[q];
                acc[q * 4 + 0] = fmaf(g.x, w[t], acc[q * 4 + 0]);
                acc[q * 4 + 1] = fmaf(g.y, w[t], acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(g.z, w[t], acc[q * 4 + 2]);
                acc[q * 4 + 3] = fmaf(g.w, w[t], acc[q * 4 + 3]);
              }
            }
          }
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (b0 + t < B) atomicAdd(p.dact[i] + (size_t)(b0 + t) * in + k, acc[t]);
        }
      } else if (u < n_dx + n_dw) {
        // dW[o, k] += sum_b dpre[b, o] x[b, k], db[o] += sum_b dpre[b, o]: thread = input column k, 8 output rows
        const int v = u - n_dx;
        const int kb = v % nkb, rb = v / nkb;
        const int o0 = rb * kMapBwdRows, k = kb * kMapFusedThreads + (int)threadIdx.x;
        const float* x = p.act[i];
        float acc[kMapBwdRows], accb = 0.f;
#pragma unroll
        for (int r = 0; r < kMapBwdRows; ++r) acc[r] = 0.f;
        for (int b0 = 0; b0 < B; b0 += 32) {
          __syncthreads();
          {
            const int bb = threadIdx.x / kMapBwdRows, r = threadIdx.x % kMapBwdRows;  // 32 x 8 = 256 entries
            float g = 0.f;
            if (b0 + bb < B && o0 + r < out) {
              g = __ldcg(dY + (size_t)(b0 + bb) * out + o0 + r);
              if (leaky && __ldg(y + (size_t)(b0 + bb) * out + o0 + r) < 0.f) g *= 0.2f;
            }
            s_g[bb * kMapBwdRows + r] = g;
          }
          float xr[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) xr[t] = (k < in && b0 + t < B) ? __ldg(x + (size_t)(b0 + t) * in + k) : 0.f;
          __syncthreads();
#pragma unroll
          for (int t = 0; t < 32; ++t) {
#pragma unroll
            for (int r = 0; r < kMapBwdRows; ++r) acc[r] = fmaf(s_g[t * kMapBwdRows + r], xr[t], acc[r]);
          }
          if (kb == 0 && threadIdx.x < kMapBwdRows) {
            for (int t = 0; t < 32; ++t) accb += s_g[t * kMapBwdRows + threadIdx.x];
          }
        }
        if (k < in) {
#pragma unroll
          for (int r = 0; r < kMapBwdRows; ++r)
            if (o0 + r < out) p.dW[i][(size_t)(o0 + r) * in + k] += acc[r];
        }
        if (kb == 0 && threadIdx.x < kMapBwdRows && o0 + (int)threadIdx.x < out) p.db[i][o0 + threadIdx.x] += accb;
      } else {
        // dW_0, db_0 of the hoisted first FiLM layer: warp = feature j, lanes stride the input columns
        const int j = (u - n_dx - n_dw) * 8 + warp;
        for (int col = lane; col < in0; col += 32) {
          float acc = 0.f;
          for (int b0 = 0; b0 < B; b0 += 8) {
            float t[8];
#pragma unroll
            for (int uu = 0; uu < 8; ++uu) {
              const int b = b0 + uu;
              t[uu] = 0.f;
              if (b < B) {
                const float* dM = p.dM + (size_t)b * 4 * kH + j;
                const float* z = p.Z + (size_t)b * 3 * N;
                if (p.so2) {
                  if (col == 0) t[uu] = __ldcg(dM + 2 * kH);
                  else if (col == 1) t[uu] = __ldcg(dM + 3 * kH);
                  else t[uu] = fmaf(__ldcg(dM), __ldg(z + (col - 2) * 3), __ldcg(dM + kH) * __ldg(z + (col - 2) * 3 + 2));
                } else {
                  t[uu] = fmaf(__ldcg(dM), __ldg(z + col * 3),
                               fmaf(__ldcg(dM + kH), __ldg(z + col * 3 + 1), __ldcg(dM + 2 * kH) * __ldg(z + col * 3 + 2)));
                }
              }
            }
#pragma unroll
            for (int uu = 0; uu < 8; ++uu) acc += t[uu];
          }
          p.dW0[(size_t)j * in0 + col] += acc;
        }
        {  // db_0[j] += sum_b d_mc[b][4][j] * freq_0[b][j]: lanes stride the maps
          float acc = 0.f;
          for (int b = lane; b < B; b += 32)
            acc = fmaf(__ldg(p.d_mc + (size_t)b * 5 * kH + 4 * kH + j),
                       fmaf(__ldg(p.act[n] + (size_t)b * 2 * half + j), 15.f, 30.f), acc);
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
          if (lane == 0) p.db0[j] += acc;
        }
      }
    }
    map_grid_sync(p.sync, ++epoch * gridDim.x, p.sync + 1);
  }

  // ---- dZ: hoisted-layer part (dM W_ip) + mapping-input part ((dG + dG^T) Z, d Z_y); warp = (map, latent row, component)
  {
    const int mn_in = p.dims[0];
    for (int u = gwarp; u < B * 3 * N; u += nwarps) {
      const int b = u / (3 * N), ic = u - b * 3 * N;
      const int q = ic / 3, c = ic - q * 3;
      const float* z = p.Z + (size_t)b * 3 * N;
      const float* dx0 = p.dact[0] + (size_t)b * mn_in;
      const float* dM = p.dM + (size_t)b * 4 * kH;
      float acc = 0.f;
      if (p.so2 && c == 1) {
        if (lane == 0) acc = __ldcg(dx0 + N * N + q);  // Z_y enters the mapping input as it is
      } else {
        const float* dm = dM + (p.so2 ? (c == 0 ? 0 : 1) : c) * kH;
        const float* w = p.W0 + (p.so2 ? 2 : 0) + q;
#pragma unroll
        for (int jj = 0; jj < kH / 32; ++jj) {
          const int j = jj * 32 + lane;
          acc = fmaf(__ldcg(dm + j), __ldg(w + (size_t)j * in0), acc);
        }
        for (int m = lane; m < N; m += 32)
          acc = fmaf(__ldcg(dx0 + q * N + m) + __ldcg(dx0 + m * N + q), __ldg(z + m * 3 + c), acc);
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
      if (lane == 0) p.dZ[(size_t)b * 3 * N + ic] = acc;
    }
  }
}

}  // namespace reni
