// HBM-bound helper kernels of the RENI decoder path: weight image preparation, the per-map
// layer-0 hoisting prologue (M_b, c_b), loss reduction, and a tcgen05 descriptor self-test.
#pragma once
#include "layout.cuh"
#include "ptx.cuh"

namespace reni {

// ------------------------------------------------------------------------------------------------
// Weight preparation: fp32 nn.Linear parameters -> fp16 tile images with omega folded in.
//   wf[l]  (l = 1..L)  [k/8][n][8]  = omega_l     * W_l[n][k]          forward  B operand (N = out, K = in)
//   wb[l]              [j/8][k][8]  = omega_{l-1} * W_l[j][k]          backward B operand (N = in,  K = out)
//   w6f                [k/8][16][8] = s * W_out[n][k]   (n < out_features, else 0);  s = omega if sine-last
//   w6b                [c/8][256][8]= omega_L * s * W_out[c][k]
//   bias               L*256: omega_l * b_l ; then 16: s * b_out
// Reference: SineLayer.forward sin(omega*(xW^T+b)) (RENI.py:86-87), final Linear (RENI.py:153-162).
// ------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* w[kMaxHiddenLayers + 2];  // w[0] = first layer (unused here), w[1..L] hidden, w[L+1] = output
  const float* b[kMaxHiddenLayers + 2];
  __half* wf;
  __half* wb;
  __half* w6f;
  __half* w6b;
  float* bias;
  int L, out_features, last_sine;
  float first_omega, hidden_omega;
};

__global__ void reni_prep_weights_kernel(const PrepParams p) {
  const int l = blockIdx.y;  // 0..L-1 -> hidden layer l+1 ; L -> output layer
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  if (l < p.L) {
    const float* W = p.w[l + 1];
    const float om_f = p.hidden_omega;                            // omega of layer l+1 (hidden)
    const float om_b = (l == 0) ? p.first_omega : p.hidden_omega;  // omega of the layer feeding it
    __half* wf = p.wf + (size_t)l * kH * kH;
    __half* wb = p.wb + (size_t)l * kH * kH;
    for (int i = tid; i < kH * kH; i += nthreads) {
      const int n = i / kH, k = i % kH;  // W[n][k], coalesced read
      const float w = W[i];
      wf[((k >> 3) * kH + n) * 8 + (k & 7)] = __float2half_rn(om_f * w);
      wb[((n >> 3) * kH + k) * 8 + (n & 7)] = __float2half_rn(om_b * w);
    }
    for (int i = tid; i < kH; i += nthreads) p.bias[l * kH + i] = om_f * p.b[l + 1][i];
  } else {
    const float* W = p.w[p.L + 1];
    const float s = p.last_sine ? p.hidden_omega : 1.0f;
    const float om_b = (p.L == 0) ? p.first_omega : p.hidden_omega;
    for (int i = tid; i < kW6N * kH; i += nthreads) {
      const int n = i / kH, k = i % kH;
      const float w = (n < p.out_features) ? s * W[n * kH + k] : 0.f;
      p.w6f[((k >> 3) * kW6N + n) * 8 + (k & 7)] = __float2half_rn(w);
      p.w6b[((n >> 3) * kH + k) * 8 + (n & 7)] = __float2half_rn(om_b * w);
    }
    for (int i = tid; i < kW6N; i += nthreads)
      p.bias[p.L * kH + i] = (i < p.out_features) ? s * p.b[p.L + 1][i] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Prologue: per-map hoisting of layer 0.  With the invariant encodings (RENI.py:23-60) every column of
// the first-layer input is either constant per map or linear in <= 4 direction features f, so
//     omega0 * (x W0^T + b0) = f . M_b' + c_b'        (M_b' 4x256, c_b' 256, omega0 folded)
//   SO2 : f = [dx, dz, |d_xz|, dy]   columns [ip N | vec(G) N^2 | |d_xz| | Z_y N | dy]   (RENI.py:51)
//   SO3 : f = [dx, dy, dz, 0]        columns [ip N | vec(Z Z^T) N^2]                     (RENI.py:27)
//   None: f = [dx, dy, dz, 0]        columns [ip N | vec(Z) 3N]                          (RENI.py:59)
// Grid (256/8, B): one block per (map, 8 output features); a warp per output feature reads its W0 row
// coalesced and dots it with the per-map constant vector staged in shared memory.
// ------------------------------------------------------------------------------------------------
struct PrologueParams {
  const float* Z;   // (B, N, 3)
  const float* W0;  // (256, in_features)
  const float* b0;  // (256)
  float* mc;        // (B, 5, 256)
  float* xc;        // (B, in_features) per-map constant columns (optional, for the dW0 kernel)
  int B, N, in_features, equivariance;  // 0 None, 1 SO2, 2 SO3
  float omega0;
};

__global__ void __launch_bounds__(256) reni_prologue_kernel(const PrologueParams p) {
  extern __shared__ float s_x[];  // in_features constants (0 where the column depends on the direction) + 3N latents
  const int b = blockIdx.y;
  const int N = p.N;
  float* s_z = s_x + p.in_features;
  const float* Zb = p.Z + (size_t)b * N * 3;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_z[i] = Zb[i];
  __syncthreads();
  for (int i = threadIdx.x; i < p.in_features; i += blockDim.x) {
    float v = 0.f;
    if (p.equivariance == 1) {
      if (i >= N && i < N + N * N) {
        const int n = (i - N) / N, m = (i - N) % N;
        v = s_z[n * 3] * s_z[m * 3] + s_z[n * 3 + 2] * s_z[m * 3 + 2];  // G = Z_xz Z_xz^T (RENI.py:40)
      } else if (i > N + N * N && i < 2 * N + N * N + 1) {
        v = s_z[(i - N - N * N - 1) * 3 + 1];  // Z_y (RENI.py:47)
      }
    } else if (p.equivariance == 2) {
      if (i >= N) {
        const int n = (i - N) / N, m = (i - N) % N;
        v = s_z[n * 3] * s_z[m * 3] + s_z[n * 3 + 1] * s_z[m * 3 + 1] + s_z[n * 3 + 2] * s_z[m * 3 + 2];
      }
    } else {
      if (i >= N) v = s_z[i - N];
    }
    s_x[i] = v;
    if (p.xc != nullptr && blockIdx.x == 0) p.xc[(size_t)b * p.in_features + i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const float* w = p.W0 + (size_t)j * p.in_features;
  float c = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
  for (int i = lane; i < p.in_features; i += 32) {
    const float wv = __ldg(w + i);
    c = fmaf(wv, s_x[i], c);
    if (i < N) {
      if (p.equivariance == 1) {
        m0 = fmaf(wv, s_z[i * 3], m0);
        m1 = fmaf(wv, s_z[i * 3 + 2], m1);
      } else {
        m0 = fmaf(wv, s_z[i * 3], m0);
        m1 = fmaf(wv, s_z[i * 3 + 1], m1);
        m2 = fmaf(wv, s_z[i * 3 + 2], m2);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, s);
    m0 += __shfl_xor_sync(0xffffffffu, m0, s);
    m1 += __shfl_xor_sync(0xffffffffu, m1, s);
    m2 += __shfl_xor_sync(0xffffffffu, m2, s);
  }
  if (lane == 0) {
    float* o = p.mc + (size_t)b * 5 * kH;
    float r2, r3;
    if (p.equivariance == 1) {
      r2 = w[N + N * N];          // |d_xz| column
      r3 = w[2 * N + N * N + 1];  // d_y column
    } else {
      r2 = m2;
      r3 = 0.f;
    }
    o[0 * kH + j] = p.omega0 * m0;
    o[1 * kH + j] = p.omega0 * m1;
    o[2 * kH + j] = p.omega0 * r2;
    o[3 * kH + j] = p.omega0 * r3;
    o[4 * kH + j] = p.omega0 * (c + p.b0[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 descriptor self-test: D[128 x N] = A * B^T with both operands given as ready-made shared
// memory images and every descriptor field supplied at run time.  Used by tests to pin the operand
// layouts (K-major and MN-major, SWIZZLE_NONE) independently of the pipelined kernels.
// ------------------------------------------------------------------------------------------------
struct SelfTestParams {
  const uint8_t* a_img;
  const uint8_t* b_img;
  float* d_out;  // [128][N] row-major
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;  // start-address advance per K = 16 step
  uint32_t a_mn, b_mn, N, ksteps;
};

__global__ void __launch_bounds__(128, 1) reni_selftest_umma_kernel(const SelfTestParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) & ~1023u);
  for (uint32_t i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
  for (uint32_t i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(p.b_img + i);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc<256>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, p.N, p.a_mn, p.b_mn);
    for (uint32_t k = 0; k < p.ksteps; ++k) {
      const uint64_t da = umma_smem_desc(smem_u32(sa) + k * p.a_kstep, p.a_lbo, p.a_sbo);
      const uint64_t db = umma_smem_desc(smem_u32(sb) + k * p.b_kstep, p.b_lbo, p.b_sbo);
      umma_f16_ss(tmem_base, da, db, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row = warp * 32 + lane;
  for (uint32_t c = 0; c < p.N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) p.d_out[row * p.N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<256>(tmem_base);
}

}  // namespace reni

namespace reni {

// ------------------------------------------------------------------------------------------------
// Loss finish (one block per map): reduce the forward kernel's per-tile partial sums into
//   mse_b = (1/(3P)) sum (o-t)^2 sw                          (loss_functions.py:6-13)
//   cosine_b = 1 - mean_c( cos_sim_c(o,t over pixels) * sw[b,0,c] )   (loss_functions.py:25-32)
// accumulate loss_out = [loss, mse, prior, cosine] and leave the per-map coefficients the backward
// kernel needs for d(beta*cosine)/do, pre-multiplied by the gradient scale S = 3P/2:
//   S*g_o = (o-t)*sw + coefA_c*t + coefB_c*o
// ------------------------------------------------------------------------------------------------
struct LossFinishParams {
  const float* loss_part;  // (ntiles, 4, 10)
  const float* sw;         // (B or 1, P, 3)
  int64_t sw_bstride;
  const float* Z;          // (B, N, 3) for the prior term (may be null)
  float* map_loss;         // (B, 32)
  float* loss_out;         // [4] accumulated with atomics (caller zeroes)
  float* scalars;          // [0] = S, [1] = 1/S
  int B, P, tiles_per_map, nz;  // nz = N*3
  float alpha, beta;
  int use_cos;
};

__global__ void __launch_bounds__(128) reni_loss_finish_kernel(const LossFinishParams p) {
  const int b = blockIdx.x;
  __shared__ float s_part[4][kLossPartials];
  float part[kLossPartials];
#pragma unroll
  for (int i = 0; i < kLossPartials; ++i) part[i] = 0.f;
  const float* lp = p.loss_part + (size_t)b * p.tiles_per_map * 4 * kLossPartials;
  for (int i = threadIdx.x; i < p.tiles_per_map * 4; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < kLossPartials; ++k) part[k] += lp[(size_t)i * kLossPartials + k];
  }
  float zz = 0.f;
  if (p.Z != nullptr)
    for (int i = threadIdx.x; i < p.nz; i += blockDim.x) {
      const float z = p.Z[(size_t)b * p.nz + i];
      zz = fmaf(z, z, zz);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kLossPartials; ++k) {
    float x = part[k];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
    if (lane == 0) s_part[warp][k] = x;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) zz += __shfl_xor_sync(0xffffffffu, zz, s);
  __shared__ float s_zz[4];
  if (lane == 0) s_zz[warp] = zz;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[kLossPartials];
    for (int k = 0; k < kLossPartials; ++k) t[k] = s_part[0][k] + s_part[1][k] + s_part[2][k] + s_part[3][k];
    const float S = 1.5f * (float)p.P;
    const float mse = t[0] / (3.f * (float)p.P);
    float cosl = 0.f;
    float* ml = p.map_loss + (size_t)b * 32;
    for (int c = 0; c < 3; ++c) {
      float cA = 0.f, cB = 0.f;
      if (p.use_cos) {
        const float no = fmaxf(sqrtf(t[4 + c]), 1e-20f), nt = fmaxf(sqrtf(t[7 + c]), 1e-20f);
        const float w0 = p.sw[(size_t)b * p.sw_bstride + c];  // sine weight (x mask) of pixel 0
        const float cs = t[1 + c] / (no * nt);
        cosl += cs * w0;
        const float k = S * p.beta * (-w0 / 3.f);
        cA = k / (no * nt);
        cB = -k * t[1 + c] / (no * no * no * nt);
      }
      ml[16 + c] = cA;
      ml[19 + c] = cB;
    }
    cosl = p.use_cos ? p.beta * (1.f - cosl / 3.f) : 0.f;
    const float prior = p.alpha * (s_zz[0] + s_zz[1] + s_zz[2] + s_zz[3]);
    ml[0] = mse;
    ml[1] = cosl;
    atomicAdd(p.loss_out + 0, mse + prior + cosl);
    atomicAdd(p.loss_out + 1, mse);
    atomicAdd(p.loss_out + 2, prior);
    atomicAdd(p.loss_out + 3, cosl);
    if (b == 0) {
      p.scalars[0] = S;
      p.scalars[1] = 1.f / S;
    }
  }
}

// gradient scale for an external grad_out: S = 1 / max|g| (keeps the fp16 deltas in range)
__global__ void reni_absmax_kernel(const float* g, int64_t n, unsigned int* slot) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(g[i]));
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}
__global__ void reni_scale_from_absmax_kernel(const unsigned int* slot, float* scalars) {
  const float m = __uint_as_float(*slot);
  const float S = (m > 0.f && isfinite(m)) ? 1.f / m : 1.f;
  scalars[0] = S;
  scalars[1] = 1.f / S;
}

// ------------------------------------------------------------------------------------------------
// Layer-0 reduction (one block per tile): dM_b[i][j] += sum_p f_i delta0[p,j], dc_b[j] += sum_p delta0[p,j]
// from the fp16 delta_0 stash; f recomputed from the directions (never stored).
// ------------------------------------------------------------------------------------------------
struct L0ReduceParams {
  const __half* stash_d;
  const float* D;
  int64_t d_bstride;
  const float* scalars;
  float* dmc;  // (B, 5, 256), caller zeroes
  int P, tiles_per_map, d_slots, so2;
};

__global__ void __launch_bounds__(256) reni_layer0_reduce_kernel(const L0ReduceParams p) {
  __shared__ float s_f[kTileRows][4];
  __shared__ float s_red[8][5][kH];  // 40 KB
  const int tile = blockIdx.x;
  const int b = tile / p.tiles_per_map;
  const int p0 = (tile - b * p.tiles_per_map) * kTileRows;
  if (threadIdx.x < kTileRows) {
    const int pix = p0 + threadIdx.x;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
    if (pix < p.P) {
      const float* d = p.D + (size_t)b * p.d_bstride + (size_t)pix * 3;
      const float dx = d[0], dy = d[1], dz = d[2];
      if (p.so2) { f0 = dx; f1 = dz; f2 = sqrtf(dx * dx + dz * dz); f3 = dy; }
      else       { f0 = dx; f1 = dy; f2 = dz; }
    }
    s_f[threadIdx.x][0] = f0; s_f[threadIdx.x][1] = f1; s_f[threadIdx.x][2] = f2; s_f[threadIdx.x][3] = f3;
  }
  __syncthreads();
  const uint8_t* src = reinterpret_cast<const uint8_t*>(p.stash_d) + (size_t)tile * p.d_slots * kTileImageBytes;
  // thread -> (8-column group kg, group rg of 16 rows); rows are read in pairs (32 contiguous bytes = one sector)
  const int kg = threadIdx.x & 31;
  const int rg = threadIdx.x >> 5;
  float acc[5][8];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
  for (int rr = 0; rr < 16; rr += 2) {
    const int r = rg * 16 + rr;
    const uint4* ptr = reinterpret_cast<const uint4*>(src + stash_off(r, kg, kH));
    const uint4 vv[2] = {__ldg(ptr), __ldg(ptr + 1)};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 v = vv[h];
      float d[8];
      float2 t2;
      t2 = __half22float2(*reinterpret_cast<const __half2*>(&v.x)); d[0] = t2.x; d[1] = t2.y;
      t2 = __half22float2(*reinterpret_cast<const __half2*>(&v.y)); d[2] = t2.x; d[3] = t2.y;
      t2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z)); d[4] = t2.x; d[5] = t2.y;
      t2 = __half22float2(*reinterpret_cast<const __half2*>(&v.w)); d[6] = t2.x; d[7] = t2.y;
      const float f0 = s_f[r + h][0], f1 = s_f[r + h][1], f2 = s_f[r + h][2], f3 = s_f[r + h][3];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[0][k] = fmaf(f0, d[k], acc[0][k]);
        acc[1][k] = fmaf(f1, d[k], acc[1][k]);
        acc[2][k] = fmaf(f2, d[k], acc[2][k]);
        acc[3][k] = fmaf(f3, d[k], acc[3][k]);
        acc[4][k] += d[k];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) s_red[rg][i][kg * 8 + k] = acc[i][k];
  __syncthreads();
  const float inv_s = p.scalars[1];
  float* dst = p.dmc + (size_t)b * 5 * kH;
  for (int i = threadIdx.x; i < 5 * kH; i += blockDim.x) {
    const int a = i / kH, j = i % kH;
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += s_red[g][a][j];
    atomicAdd(dst + i, s * inv_s);
  }
}

// ------------------------------------------------------------------------------------------------
// Map-level backward of the hoisted first layer.
//   xc (B, in)   : per-map constant input columns (written by the prologue)
//   dmc (B,5,256): rows 0..3 = dM_b, row 4 = dc_b
//   (1) dW0[j,i] += sum_b dc_b[j] xc_b[i] + [i<N] sum_b dM_b[.][j] Z_b[i,.] + direction columns ; db0 += sum_b dc_b
//   (2) dxc[b,i]  = sum_j dc_b[j] W0[j,i]      ;  dip[b,c,n] = sum_j dM_b[c][j] W0[j,n]
//   (3) dZ from dxc / dip (chain rule through G = Z Z^T etc.) + 2 alpha Z
// ------------------------------------------------------------------------------------------------
struct MapBwdParams {
  const float* Z;
  const float* W0;
  const float* xc;
  const float* dmc;
  float* dW0;   // (256, in) accumulated
  float* db0;   // (256) accumulated
  float* dxc;   // (B, in) scratch
  float* dip;   // (B, 3, N) scratch
  float* dZ;    // (B, N, 3) written (+= if accumulate)
  int B, N, in_features, equivariance;
  float alpha2;  // 2*alpha (prior gradient), 0 if none
  int accumulate;
};

__global__ void __launch_bounds__(256) reni_dw0_kernel(const MapBwdParams p) {
  // one thread per dW0 element, i fastest (coalesced); loops over the maps
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int in = p.in_features, N = p.N;
  if (idx >= (int64_t)kH * in) return;
  const int j = (int)(idx / in), i = (int)(idx % in);
  float s = 0.f, sb = 0.f;
  const int col_dn = N + N * N, col_dy = 2 * N + N * N + 1;
  for (int b = 0; b < p.B; ++b) {
    const float* m = p.dmc + (size_t)b * 5 * kH;
    const float dc = m[4 * kH + j];
    s = fmaf(dc, p.xc[(size_t)b * in + i], s);
    if (i < N) {
      const float* z = p.Z + ((size_t)b * N + i) * 3;
      if (p.equivariance == 1) s += m[j] * z[0] + m[kH + j] * z[2];
      else s += m[j] * z[0] + m[kH + j] * z[1] + m[2 * kH + j] * z[2];
    } else if (p.equivariance == 1) {
      if (i == col_dn) s += m[2 * kH + j];
      if (i == col_dy) s += m[3 * kH + j];
    }
    if (i == 0) sb += dc;
  }
  p.dW0[idx] += s;
  if (i == 0) p.db0[j] += sb;
}

__global__ void __launch_bounds__(256) reni_dxc_kernel(const MapBwdParams p) {
  // grid (ceil(in/256), B): dxc[b,i] = sum_j dc_b[j] W0[j,i]; dip[b,c,n] = sum_j dM_b[c][j] W0[j,n]
  __shared__ float s_m[5 * kH];
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = threadIdx.x; k < 5 * kH; k += blockDim.x) s_m[k] = p.dmc[(size_t)b * 5 * kH + k];
  __syncthreads();
  if (i >= p.in_features) return;
  float s = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
  const bool ip = i < p.N;
  for (int j = 0; j < kH; ++j) {
    const float w = __ldg(p.W0 + (size_t)j * p.in_features + i);
    s = fmaf(s_m[4 * kH + j], w, s);
    if (ip) {
      d0 = fmaf(s_m[j], w, d0);
      d1 = fmaf(s_m[kH + j], w, d1);
      d2 = fmaf(s_m[2 * kH + j], w, d2);
    }
  }
  p.dxc[(size_t)b * p.in_features + i] = s;
  if (ip) {
    float* o = p.dip + (size_t)b * 3 * p.N;
    o[i] = d0;
    o[p.N + i] = d1;
    o[2 * p.N + i] = d2;
  }
}

__global__ void __launch_bounds__(128) reni_dz_kernel(const MapBwdParams p) {
  // grid (B): one thread per latent row n
  const int b = blockIdx.x, N = p.N;
  const float* Z = p.Z + (size_t)b * N * 3;
  const float* dxc = p.dxc + (size_t)b * p.in_features;
  const float* dip = p.dip + (size_t)b * 3 * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (p.equivariance == 1) {
      const float* dG = dxc + N;
      for (int m = 0; m < N; ++m) {
        const float s = dG[n * N + m] + dG[m * N + n];
        g0 = fmaf(s, Z[m * 3], g0);
        g2 = fmaf(s, Z[m * 3 + 2], g2);
      }
      g0 += dip[n];
      g2 += dip[N + n];
      g1 = dxc[N + N * N + 1 + n];
    } else if (p.equivariance == 2) {
      const float* dG = dxc + N;
      for (int m = 0; m < N; ++m) {
        const float s = dG[n * N + m] + dG[m * N + n];
        g0 = fmaf(s, Z[m * 3], g0);
        g1 = fmaf(s, Z[m * 3 + 1], g1);
        g2 = fmaf(s, Z[m * 3 + 2], g2);
      }
      g0 += dip[n];
      g1 += dip[N + n];
      g2 += dip[2 * N + n];
    } else {
      g0 = dxc[N + n * 3] + dip[n];
      g1 = dxc[N + n * 3 + 1] + dip[N + n];
      g2 = dxc[N + n * 3 + 2] + dip[2 * N + n];
    }
    g0 = fmaf(p.alpha2, Z[n * 3], g0);
    g1 = fmaf(p.alpha2, Z[n * 3 + 1], g1);
    g2 = fmaf(p.alpha2, Z[n * 3 + 2], g2);
    float* o = p.dZ + ((size_t)b * N + n) * 3;
    if (p.accumulate) { o[0] += g0; o[1] += g1; o[2] += g2; }
    else              { o[0] = g0;  o[1] = g1;  o[2] = g2; }
  }
}

}  // namespace reni
