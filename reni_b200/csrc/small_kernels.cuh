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
