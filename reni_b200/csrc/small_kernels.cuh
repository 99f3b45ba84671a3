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
//   wf2[l] [n/128][k/8][n%128][8]   = omega_l     * W_l[n][k]          forward B operand split in two N halves for
//                                                                     CTA pairs (tcgen05.mma.cta_group::2)
//   wb2[l] [k/128][j/8][k%128][8]   = omega_{l-1} * W_l[j][k]          backward B operand, same split
//   w6f                [k/8][16][8] = s * W_out[n][k]   (n < out_features, else 0);  s = omega if sine-last
//   w6b                [c/8][256][8]= omega_L * s * W_out[c][k]
//   bias               L*256: omega_l * b_l ; then 16: s * b_out
//   wbias2[l][n/128]   [2][128][8]  = (hi, lo, 0, ...) fp16 split of omega_l * b_l: B operand of the bias K-step
// Reference: SineLayer.forward sin(omega*(xW^T+b)) (RENI.py:86-87), final Linear (RENI.py:153-162).
// ------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* w[kMaxHiddenLayers + 2];  // w[0] = first layer (unused here), w[1..L] hidden, w[L+1] = output
  const float* b[kMaxHiddenLayers + 2];
  __half* wf;
  __half* wb;
  __half* wf2;
  __half* wf2lo;   // residual of wf2: fp16(omega_l W_l - wf2), same geometry (two-term forward weights)
  __half* wb2;
  __half* wbias2;  // [l][n/128][2 k-groups][128][8]: column k=0 = hi, k=1 = lo fp16 halves of omega_l*b_l, rest 0
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
    __half* wf2 = p.wf2 + (size_t)l * kH * kH;
    __half* wf2lo = p.wf2lo + (size_t)l * kH * kH;
    __half* wb2 = p.wb2 + (size_t)l * kH * kH;
    for (int i = tid; i < kH * kH; i += nthreads) {
      const int n = i / kH, k = i % kH;  // W[n][k], coalesced read
      const float w = W[i];
      wf[((k >> 3) * kH + n) * 8 + (k & 7)] = __float2half_rn(om_f * w);
      wb[((n >> 3) * kH + k) * 8 + (n & 7)] = __float2half_rn(om_b * w);
      const __half whi = __float2half_rn(om_f * w);
      wf2[(((n >> 7) * (kH / 8) + (k >> 3)) * 128 + (n & 127)) * 8 + (k & 7)] = whi;
      wf2lo[(((n >> 7) * (kH / 8) + (k >> 3)) * 128 + (n & 127)) * 8 + (k & 7)] =
          __float2half_rn(om_f * w - __half2float(whi));
      wb2[(((k >> 7) * (kH / 8) + (n >> 3)) * 128 + (k & 127)) * 8 + (n & 7)] = __float2half_rn(om_b * w);
    }
    for (int i = tid; i < kH; i += nthreads) {
      const float bv = om_f * p.b[l + 1][i];
      p.bias[l * kH + i] = bv;
      // the paired forward adds the bias on the tensor core: one extra K = 16 step [1 1 0 ..] x [hi lo 0 ..]^T
      const __half hi = __float2half_rn(bv);
      const __half lo = __float2half_rn(bv - __half2float(hi));
      __half* blk = p.wbias2 + ((size_t)l * 2 + (i >> 7)) * (2 * 128 * 8);
      uint4* row0 = reinterpret_cast<uint4*>(blk + (i & 127) * 8);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      __half2 h2 = __halves2half2(hi, lo);
      v.x = *reinterpret_cast<uint32_t*>(&h2);
      row0[0] = v;
      reinterpret_cast<uint4*>(blk + (128 + (i & 127)) * 8)[0] = make_uint4(0u, 0u, 0u, 0u);
    }
  } else {
    const float* W = p.w[p.L + 1];
    const float s = p.last_sine ? p.hidden_omega : 1.0f;
    const float om_b = (p.L == 0) ? p.first_omega : p.hidden_omega;
    for (int i = tid; i < kW6N * kH; i += nthreads) {
      const int n = i / kH, k = i % kH;
      const float w = (n < p.out_features) ? s * W[n * kH + k] : 0.f;
      // forward image: rows 0..2 = fp16(W_out), rows 3..5 = the fp16 residual (the epilogue adds columns c and c + 3)
      float wfwd = w;
      if (n >= 3 && n < 3 + p.out_features) {
        const float w0 = s * W[(n - 3) * kH + k];
        wfwd = w0 - __half2float(__float2half_rn(w0));
      }
      p.w6f[((k >> 3) * kW6N + n) * 8 + (k & 7)] = __float2half_rn(wfwd);
      p.w6b[((n >> 3) * kH + k) * 8 + (n & 7)] = __float2half_rn(om_b * w);
    }
    for (int i = tid; i < kW6N; i += nthreads)
      p.bias[p.L * kH + i] = (i < p.out_features) ? s * p.b[p.L + 1][i] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// FiLM per-map operand images (RENI_FLAG_FILM_PERMAP).  FiLMLayer (RENI.py:515-524):
//     sin(freq_l[b] * (W_l h + b_l) + phase_l[b]) = sin((diag(freq_l[b]) W_l) h + (freq_l[b] * b_l + phase_l[b]))
// so with V[n][k] = freq_l[b][n] * W_l[n][k] (one fp16 rounding of the fp32 product) as map b's forward image, the same
// values in the backward layout (delta_{l-1} = cos(a_{l-1}) * sum_j delta_l[j] freq_l[b][j] W_l[j][k]) and
// freq * b + phase as its bias block, the FiLM layers run on the kernels of the Cond-by-Concat decoder.
//   wf2m[b][l][n/128][k/8][n%128][k%8]   wb2m[b][l][k/128][n/8][k%128][n%8]   wbias2m[b][l][n/128] [2][128][8]
// Grid (8, L, B) x 256 threads; every thread writes whole 16-byte groups, a warp 512 contiguous bytes.
// ------------------------------------------------------------------------------------------------
struct FilmPrepParams {
  const float* w[kMaxHiddenLayers + 2];  // [1..L] are read
  const float* b[kMaxHiddenLayers + 2];
  const float* film;                     // (B, L, 2, 256)
  __half* wf2m;
  __half* wf2m_lo;  // residual images of wf2m (two-term forward weights)
  __half* wb2m;
  __half* wbias2m;
  int L;
};

__global__ void __launch_bounds__(256) reni_film_prep_maps_kernel(const FilmPrepParams p) {
  const int l = blockIdx.y, b = blockIdx.z;
  const float* W = p.w[l + 1];
  const float* freq = p.film + ((size_t)b * p.L + l) * 2 * kH;
  const float* phase = freq + kH;
  __shared__ float s_freq[kH];
  s_freq[threadIdx.x] = freq[threadIdx.x];
  __syncthreads();
  uint4* wf = reinterpret_cast<uint4*>(p.wf2m + ((size_t)b * p.L + l) * kH * kH);
  uint4* wflo = reinterpret_cast<uint4*>(p.wf2m_lo + ((size_t)b * p.L + l) * kH * kH);
  uint4* wb = reinterpret_cast<uint4*>(p.wb2m + ((size_t)b * p.L + l) * kH * kH);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
#pragma unroll 4  // (2048 threads per (map, layer): four independent iterations, their loads issued together)
  for (int i = tid; i < kH * (kH / 8); i += nthreads) {
    {  // forward image: thread = (n, 8 consecutive k)
      const int n = i & (kH - 1), kg = i >> 8;
      const float f = s_freq[n];
      const float4 a = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * kH + kg * 8));
      const float4 c = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * kH + kg * 8 + 4));
      const float x[8] = {f * a.x, f * a.y, f * a.z, f * a.w, f * c.x, f * c.y, f * c.z, f * c.w};
      uint4 v;
      v.x = pack_half2(x[0], x[1]);
      v.y = pack_half2(x[2], x[3]);
      v.z = pack_half2(x[4], x[5]);
      v.w = pack_half2(x[6], x[7]);
      wf[((n >> 7) * (kH / 8) + kg) * 128 + (n & 127)] = v;
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
      uint4 lo;  // residuals of the four rounded pairs
      lo.x = pack_half2(x[0] - __low2float(hv[0]), x[1] - __high2float(hv[0]));
      lo.y = pack_half2(x[2] - __low2float(hv[1]), x[3] - __high2float(hv[1]));
      lo.z = pack_half2(x[4] - __low2float(hv[2]), x[5] - __high2float(hv[2]));
      lo.w = pack_half2(x[6] - __low2float(hv[3]), x[7] - __high2float(hv[3]));
      wflo[((n >> 7) * (kH / 8) + kg) * 128 + (n & 127)] = lo;
    }
    {  // backward image: thread = (8 consecutive n, k)
      const int k = i & (kH - 1), ng = i >> 8;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = s_freq[ng * 8 + j] * __ldg(W + (size_t)(ng * 8 + j) * kH + k);
      uint4 v;
      v.x = pack_half2(x[0], x[1]);
      v.y = pack_half2(x[2], x[3]);
      v.z = pack_half2(x[4], x[5]);
      v.w = pack_half2(x[6], x[7]);
      wb[((k >> 7) * (kH / 8) + ng) * 128 + (k & 127)] = v;
    }
  }
  if (blockIdx.x == 0) {  // bias block: (hi, lo) fp16 split of freq * b + phase in k-group 0, zeros in k-group 1
    const int i = threadIdx.x;
    const float bv = s_freq[i] * __ldg(p.b[l + 1] + i) + __ldg(phase + i);
    const __half hi = __float2half_rn(bv);
    const __half lo = __float2half_rn(bv - __half2float(hi));
    __half2 h2 = __halves2half2(hi, lo);
    uint4* blk = reinterpret_cast<uint4*>(p.wbias2m + (((size_t)b * p.L + l) * 2 + (i >> 7)) * (2 * 128 * 8));
    blk[i & 127] = make_uint4(*reinterpret_cast<uint32_t*>(&h2), 0u, 0u, 0u);
    blk[128 + (i & 127)] = make_uint4(0u, 0u, 0u, 0u);
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
  float* xfull;     // (B, 5, in_features): per-map input columns paired with [dM0..dM3, dc] for the dW0 GEMM (optional)
  int B, N, in_features, equivariance;  // 0 None, 1 SO2, 2 SO3
  float omega0;
  float* scalars;  // fused loss: [0] = S, [1] = 1/S written here so the backward does not wait for the loss reduction
  float fused_S;
};

__global__ void __launch_bounds__(256) reni_prologue_kernel(const PrologueParams p) {
  extern __shared__ float s_x[];  // in_features constants (0 where the column depends on the direction) + 3N latents
  const int b = blockIdx.y;
  const int N = p.N;
  if (p.scalars != nullptr && blockIdx.x == 0 && b == 0 && threadIdx.x == 0) {
    p.scalars[0] = p.fused_S;
    p.scalars[1] = 1.f / p.fused_S;
  }
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
    if (p.xfull != nullptr && blockIdx.x == 0) {
      // rows 0..3 multiply dM_b[0..3] (direction features), row 4 multiplies dc_b (constant columns)
      float* xf = p.xfull + (size_t)b * 5 * p.in_features + i;
      const int in = p.in_features;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
      if (p.equivariance == 1) {
        if (i < N) { r0 = s_z[i * 3]; r1 = s_z[i * 3 + 2]; }
        if (i == N + N * N) r2 = 1.f;
        if (i == 2 * N + N * N + 1) r3 = 1.f;
      } else if (i < N) {
        r0 = s_z[i * 3]; r1 = s_z[i * 3 + 1]; r2 = s_z[i * 3 + 2];
      }
      xf[0] = r0; xf[in] = r1; xf[2 * in] = r2; xf[3 * in] = r3; xf[4 * in] = v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const float* w = p.W0 + (size_t)j * p.in_features;
  float c = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
  // the first N columns (inner products with the direction) also feed M_b; handled apart so the long loop is branch-free
  for (int i = lane; i < N; i += 32) {
    const float wv = __ldg(w + i);
    c = fmaf(wv, s_x[i], c);
    if (p.equivariance == 1) {
      m0 = fmaf(wv, s_z[i * 3], m0);
      m1 = fmaf(wv, s_z[i * 3 + 2], m1);
    } else {
      m0 = fmaf(wv, s_z[i * 3], m0);
      m1 = fmaf(wv, s_z[i * 3 + 1], m1);
      m2 = fmaf(wv, s_z[i * 3 + 2], m2);
    }
  }
  {  // remaining columns: 4 independent loads in flight per lane (the row is read once, latency-bound otherwise)
    float c1 = 0.f, c2 = 0.f, c3 = 0.f;
    int i = N + lane;
    for (; i + 96 < p.in_features; i += 128) {
      const float w0 = __ldg(w + i), w1 = __ldg(w + i + 32), w2 = __ldg(w + i + 64), w3 = __ldg(w + i + 96);
      c = fmaf(w0, s_x[i], c);
      c1 = fmaf(w1, s_x[i + 32], c1);
      c2 = fmaf(w2, s_x[i + 64], c2);
      c3 = fmaf(w3, s_x[i + 96], c3);
    }
    for (; i < p.in_features; i += 32) c = fmaf(__ldg(w + i), s_x[i], c);
    c += c1 + c2 + c3;
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

// ------------------------------------------------------------------------------------------------
// Probe: can a (non-tensor) bulk copy into THIS CTA's shared memory complete its bytes on an mbarrier that lives in
// the PEER CTA?  (cp.async.bulk.shared::cluster.global with a mapa'd mbarrier address.)  Both CTAs of a cluster copy
// `bytes` from global into their own smem and signal the leader's barrier, which expects 2 x bytes.  The leader polls
// a bounded number of times and reports: result[0] = 1 completed / 0 timed out, result[1 + rank] = byte sum seen by
// each CTA in its own buffer.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) reni_probe_remote_tx_kernel(const uint8_t* src, uint32_t bytes,
                                                                      uint32_t* result,
                                                                      const __grid_constant__ CUtensorMap tmap,
                                                                      int use_tma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(&bar, 2 * bytes);
    const uint32_t leader_bar = mapa_u32(smem_u32(&bar), 0);
    if (use_tma) {  // tensor map: rows of 256 B; this CTA's box starts at row rank * bytes / 256
      tma2_load_2d(smem, &tmap, 0, (int32_t)(rank * bytes / 256), leader_bar);
    } else {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem)),
                   "l"(src + (size_t)rank * bytes), "r"(bytes), "r"(leader_bar)
                   : "memory");
    }
    if (rank == 0) {
      uint32_t ok = 0;
      for (int i = 0; i < 2000000 && !ok; ++i) ok = mbar_try_wait(&bar, 0) ? 1u : 0u;
      result[0] = ok;
    }
  }
  __syncthreads();
  cluster_sync_all();  // (the leader's poll loop bounds how long the peer waits here)
  if (threadIdx.x == 0) {
    uint32_t sum = 0;
    for (uint32_t i = 0; i < bytes; ++i) sum += smem[i];
    result[1 + rank] = sum;
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the self-test: one cluster of two CTAs, D[256 x N] = [A0; A1] * [B0; B1]^T with
// tcgen05.mma.cta_group::2 -- CTA r supplies A_r (its 128 rows) and B_r (its N/2 rows of B), the leader (rank 0)
// issues, the multicast commit releases both CTAs, each reads its own 128 accumulator rows.  Also exercises the
// cross-CTA hand-shake the pipelined kernels use (remote mbarrier arrive + cluster-scope acquire).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) reni_selftest_umma2_kernel(const SelfTestParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_peer;
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) & ~1023u);
  const uint8_t* ga = p.a_img + (size_t)rank * p.a_bytes;
  const uint8_t* gb = p.b_img + (size_t)rank * p.b_bytes;
  for (uint32_t i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(ga + i);
  for (uint32_t i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(gb + i);
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_peer, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc2<256>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (rank == 1 && threadIdx.x == 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_peer), 0));
  if (rank == 0 && threadIdx.x == 0) {
    mbar_wait_cluster(&bar_peer, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(256, p.N, p.a_mn, p.b_mn);
    for (uint32_t k = 0; k < p.ksteps; ++k) {
      const uint64_t da = umma_smem_desc(smem_u32(sa) + k * p.a_kstep, p.a_lbo, p.a_sbo);
      const uint64_t db = umma_smem_desc(smem_u32(sb) + k * p.b_kstep, p.b_lbo, p.b_sbo);
      umma2_f16_ss(tmem_base, da, db, idesc, k != 0);
    }
    umma2_commit_multicast(&bar_done, 0x3);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row = rank * 128 + warp * 32 + lane;
  for (uint32_t c = 0; c < p.N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) p.d_out[row * p.N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2<256>(tmem_base);
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
  int grid_w, sw_grid;           // analytic sine weights (RENI_FLAG_GRID_SINEWEIGHT)
  const uint32_t* mask_bits;
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
        const float w0 = p.sw_grid ? grid_sineweight(0, p.grid_w, p.mask_bits)
                                   : p.sw[(size_t)b * p.sw_bstride + c];  // sine weight (x mask) of pixel 0
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
// fused-loss gradient scale written without the Cond-by-Concat prologue (FiLM core): scalars = [S, 1/S]
__global__ void reni_set_scale_kernel(float* scalars, float S) {
  scalars[0] = S;
  scalars[1] = 1.f / S;
}
__global__ void reni_scale_from_absmax_kernel(const unsigned int* slot, float* scalars) {
  const float m = __uint_as_float(*slot);
  const float S = (m > 0.f && isfinite(m)) ? 1.f / m : 1.f;
  scalars[0] = S;
  scalars[1] = 1.f / S;
}

// ------------------------------------------------------------------------------------------------
// Map-level backward of the hoisted first layer, as two small GEMMs + a per-map kernel.
//   dmc   (B,5,256) : rows 0..3 = dM_b, row 4 = dc_b          xfull (B,5,in): the input columns they multiply
//   (1) dW0[j,i] += sum_{(b,r)} dmc[b,r,j] * xfull[b,r,i]      (M=256, N=in, K=5B)   ; db0[j] += sum_b dc_b[j]
//   (2) E[(b,r),i] = sum_j dmc[b,r,j] * W0[j,i]                (M=5B, N=in, K=256)   = d(loss)/d(xfull)
//   (3) dZ from E (chain rule through G = Z Z^T etc.) + 2 alpha Z
// ------------------------------------------------------------------------------------------------
struct SmallGemmParams {
  const float* A;  // A(k, m) = A[k * a_sk + m * a_sm]
  const float* B;  // B(k, n) = B[k * ldb + n]
  float* C;        // C(m, n) = C[m * ldc + n]
  int M, N, K, a_sk, a_sm, ldb, ldc;
};

// 64 x 64 output tile x 32-deep K slice per block (grid.z = K slices), 256 threads, 4 x 4 outputs per thread; every
// block does ONE round of loads (latency paid once) and adds its partial result with atomics, so C must be
// zero-filled (or hold the value to accumulate into).  These GEMMs are <= 0.7 GFLOP even at N=100, B=256.
constexpr int kSmallGemmK = 32;
__global__ void __launch_bounds__(256) reni_small_gemm_kernel(const SmallGemmParams p) {
  __shared__ float As[kSmallGemmK][64 + 4];
  __shared__ float Bs[kSmallGemmK][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64, k0 = blockIdx.z * kSmallGemmK;
  for (int e = threadIdx.x; e < kSmallGemmK * 64; e += 256) {
    int ka = e >> 6, ma = e & 63;  // A tile: follow whichever index is contiguous in memory
    if (p.a_sk == 1) { ka = e % kSmallGemmK; ma = e / kSmallGemmK; }
    const int kA = k0 + ka, mA = m0 + ma;
    As[ka][ma] = (kA < p.K && mA < p.M) ? __ldg(p.A + (size_t)kA * p.a_sk + (size_t)mA * p.a_sm) : 0.f;
    const int kk = e >> 6, c = e & 63;
    const int k = k0 + kk, n = n0 + c;
    Bs[kk][c] = (k < p.K && n < p.N) ? __ldg(p.B + (size_t)k * p.ldb + n) : 0.f;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < kSmallGemmK; ++kk) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.N) atomicAdd(p.C + (size_t)m * p.ldc + n, acc[i][j]);
    }
  }
}

// db0[j] += sum_b dc_b[j]
__global__ void reni_db0_kernel(const float* dmc, float* db0, int B) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kH) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dmc[((size_t)b * 5 + 4) * kH + j];
  db0[j] += s;
}

struct MapBwdParams {
  const float* Z;
  const float* E;   // (B, 5, in): gradient w.r.t. xfull
  float* dZ;        // (B, N, 3) written (+= if accumulate)
  int B, N, in_features, equivariance;
  float alpha2;  // 2*alpha (prior gradient), 0 if none
  int accumulate;
};

__global__ void __launch_bounds__(128) reni_dz_kernel(const MapBwdParams p) {
  // grid (B): one thread per latent row n
  const int b = blockIdx.x, N = p.N;
  const float* Z = p.Z + (size_t)b * N * 3;
  const float* Eb = p.E + (size_t)b * 5 * p.in_features;
  const float* dxc = Eb + (size_t)4 * p.in_features;  // gradient of the constant columns
  const float* dip0 = Eb;                             // gradients of the inner-product columns, per feature row
  const float* dip1 = Eb + p.in_features;
  const float* dip2 = Eb + 2 * (size_t)p.in_features;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (p.equivariance == 1) {
      const float* dG = dxc + N;
      for (int m = 0; m < N; ++m) {
        const float s = dG[n * N + m] + dG[m * N + n];
        g0 = fmaf(s, Z[m * 3], g0);
        g2 = fmaf(s, Z[m * 3 + 2], g2);
      }
      g0 += dip0[n];
      g2 += dip1[n];
      g1 = dxc[N + N * N + 1 + n];
    } else if (p.equivariance == 2) {
      const float* dG = dxc + N;
      for (int m = 0; m < N; ++m) {
        const float s = dG[n * N + m] + dG[m * N + n];
        g0 = fmaf(s, Z[m * 3], g0);
        g1 = fmaf(s, Z[m * 3 + 1], g1);
        g2 = fmaf(s, Z[m * 3 + 2], g2);
      }
      g0 += dip0[n];
      g1 += dip1[n];
      g2 += dip2[n];
    } else {
      g0 = dxc[N + n * 3] + dip0[n];
      g1 = dxc[N + n * 3 + 1] + dip1[n];
      g2 = dxc[N + n * 3 + 2] + dip2[n];
    }
    g0 = fmaf(p.alpha2, Z[n * 3], g0);
    g1 = fmaf(p.alpha2, Z[n * 3 + 1], g1);
    g2 = fmaf(p.alpha2, Z[n * 3 + 2], g2);
    float* o = p.dZ + ((size_t)b * N + n) * 3;
    if (p.accumulate) { o[0] += g0; o[1] += g1; o[2] += g2; }
    else              { o[0] = g0;  o[1] = g1;  o[2] = g2; }
  }
}

// ------------------------------------------------------------------------------------------------
// FiLM backward, map level (RENI.py:515-524: a_l = freq_l[b] * (W_l h_{l-1} + b_l) + phase_l[b]).
// The weight-gradient kernel leaves, per map b and hidden layer l,
//     S[b][l]  = sum_{p in b} delta_l[p]^T h_{l-1}[p]   (256 x 256)      cs[b][l] = sum_{p in b} delta_l[p]   (256)
// with delta_l = dL/da_l.  From these
//     dphase_l[b][j] = cs[b][l][j]
//     dfreq_l[b][j]  = sum_p delta_l[p,j] (W_l h_{l-1}[p] + b_l)[j] = sum_k S[b][l][j,k] W_l[j,k] + b_l[j] cs[b][l][j]
//     dW_l[j,k]     += sum_b freq_l[b][j] S[b][l][j,k]         db_l[j] += sum_b freq_l[b][j] cs[b][l][j]
// ------------------------------------------------------------------------------------------------
struct FilmReduceParams {
  const float* S;     // (B, L, 256, 256)
  const float* cs;    // (B, L, 256)
  const float* film;  // (B, L, 2, 256)
  const float* W[kMaxHiddenLayers + 2];  // [1..L]: fp32 hidden weights (256, 256)
  const float* b[kMaxHiddenLayers + 2];
  float* dfilm;       // (B, L, 2, 256), written
  float* dW[kMaxHiddenLayers + 2];  // [1..L]: accumulated (+=); null entries are skipped
  float* db[kMaxHiddenLayers + 2];
  int B, L;
};

// grid (256 / 8, L, B), 256 threads: one warp per (map, layer, output feature j)
__global__ void __launch_bounds__(256) reni_film_dfilm_kernel(const FilmReduceParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp, l = blockIdx.y, b = blockIdx.z;
  const float4* srow = reinterpret_cast<const float4*>(p.S + (((size_t)b * p.L + l) * kH + j) * kH);
  const float4* wrow = reinterpret_cast<const float4*>(p.W[l + 1] + (size_t)j * kH);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kH / 128; ++i) {
    const float4 sv = __ldg(srow + i * 32 + lane), wv = __ldg(wrow + i * 32 + lane);
    acc = fmaf(sv.x, wv.x, fmaf(sv.y, wv.y, fmaf(sv.z, wv.z, fmaf(sv.w, wv.w, acc))));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    const float c = p.cs[((size_t)b * p.L + l) * kH + j];
    float* o = p.dfilm + ((size_t)b * p.L + l) * 2 * kH;
    o[j] = fmaf(p.b[l + 1][j], c, acc);
    o[kH + j] = c;
  }
}

// grid (256 / 4, L), 256 threads: block = (4 output features j, layer), thread = (j, 4 input features k); sums over the
// maps with eight 16-byte loads in flight per thread (32 MB of S at 32 maps: the loop is latency-bound otherwise)
__global__ void __launch_bounds__(256) reni_film_dw_kernel(const FilmReduceParams p) {
  const int j = blockIdx.x * 4 + (threadIdx.x >> 6), l = blockIdx.y, k4 = threadIdx.x & 63;
  if (p.dW[l + 1] == nullptr) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float accb = 0.f;
  const size_t sstride = (size_t)p.L * kH * kH, fstride = (size_t)p.L * 2 * kH, cstride = (size_t)p.L * kH;
  const float* sp = p.S + ((size_t)l * kH + j) * kH + k4 * 4;
  const float* fp = p.film + (size_t)l * 2 * kH + j;
  const float* cp = p.cs + (size_t)l * kH + j;
  int b = 0;
  for (; b + 8 <= p.B; b += 8) {
    float f[8];
    float4 sv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      f[u] = __ldg(fp + (b + u) * fstride);
      sv[u] = __ldcs(reinterpret_cast<const float4*>(sp + (b + u) * sstride));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x = fmaf(f[u], sv[u].x, acc.x); acc.y = fmaf(f[u], sv[u].y, acc.y);
      acc.z = fmaf(f[u], sv[u].z, acc.z); acc.w = fmaf(f[u], sv[u].w, acc.w);
    }
    if (k4 == 0) {
#pragma unroll
      for (int u = 0; u < 8; ++u) accb = fmaf(f[u], __ldg(cp + (b + u) * cstride), accb);
    }
  }
  for (; b < p.B; ++b) {
    const float f = __ldg(fp + b * fstride);
    const float4 sv = __ldcs(reinterpret_cast<const float4*>(sp + b * sstride));
    acc.x = fmaf(f, sv.x, acc.x); acc.y = fmaf(f, sv.y, acc.y); acc.z = fmaf(f, sv.z, acc.z); acc.w = fmaf(f, sv.w, acc.w);
    if (k4 == 0) accb = fmaf(f, __ldg(cp + b * cstride), accb);
  }
  float4* dst = reinterpret_cast<float4*>(p.dW[l + 1] + (size_t)j * kH + k4 * 4);
  float4 d = *dst;
  d.x += acc.x; d.y += acc.y; d.z += acc.z; d.w += acc.w;
  *dst = d;
  if (k4 == 0 && p.db[l + 1] != nullptr) p.db[l + 1][j] += accb;
}

// ------------------------------------------------------------------------------------------------
// Variational auto-decoder latents (RENIVariationalAutoDecoder, RENI.py:329-335; KLD, loss_functions.py:16-22;
// RENIVADTrainLoss with beta = KLD_WEIGHTING and Z_dims = 3N, RENI_module.py:312-315):
//   sample   : Z[b] = mu[idx[b]] + eps[b] * exp(0.5 * log_var[idx[b]])          (eps ~ N(0,1) drawn by the caller)
//   backward : given dLoss/dZ of the decoder step,
//                dmu[idx[b]]      += s * (dZ + kw * mu)
//                dlog_var[idx[b]] += s * (dZ * eps * 0.5 * exp(0.5 log_var) - 0.5 * kw * (1 - exp(log_var)))
//              kld_out            += kw * sum(-0.5 * (1 + log_var - mu^2 - exp(log_var)))       kw = beta / Z_dims
//              (s = 1 / world size: DDP averages the whole gradient of the replicated tables)
// One block per map; accumulation with atomics (an index may repeat inside a batch, as with index_add_).
// ------------------------------------------------------------------------------------------------
struct VadParams {
  const float* mu;
  const float* log_var;
  const int64_t* idx;
  const float* eps;    // (B, nz)
  float* Z;            // (B, nz) sample
  const float* dZ;     // (B, nz)
  float* dmu;          // (dataset, nz), accumulated
  float* dlog_var;     // (dataset, nz), accumulated
  float* kld_out;      // scalar, accumulated
  int B, nz;
  float kw, grad_scale;
};

__global__ void __launch_bounds__(128) reni_vad_sample_kernel(const VadParams p) {
  const int b = blockIdx.x;
  const int64_t row = p.idx[b];
  for (int i = threadIdx.x; i < p.nz; i += blockDim.x) {
    const float m = p.mu[row * p.nz + i], lv = p.log_var[row * p.nz + i];
    p.Z[(size_t)b * p.nz + i] = fmaf(p.eps[(size_t)b * p.nz + i], expf(0.5f * lv), m);
  }
}

__global__ void __launch_bounds__(128) reni_vad_backward_kernel(const VadParams p) {
  const int b = blockIdx.x;
  const int64_t row = p.idx[b];
  float part = 0.f;
  for (int i = threadIdx.x; i < p.nz; i += blockDim.x) {
    const float m = p.mu[row * p.nz + i], lv = p.log_var[row * p.nz + i];
    const float ev = expf(lv), sd = expf(0.5f * lv);
    const float g = p.dZ[(size_t)b * p.nz + i];
    atomicAdd(p.dmu + row * p.nz + i, p.grad_scale * fmaf(p.kw, m, g));
    atomicAdd(p.dlog_var + row * p.nz + i,
              p.grad_scale * (g * p.eps[(size_t)b * p.nz + i] * 0.5f * sd - 0.5f * p.kw * (1.f - ev)));
    part += -0.5f * (1.f + lv - m * m - ev);
  }
  __shared__ float s_part[4];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(p.kld_out, p.kw * (s_part[0] + s_part[1] + s_part[2] + s_part[3]));
}

// ------------------------------------------------------------------------------------------------
// Fused Adam over a list of fp32 segments (the flat decoder-weight buffer, the latent table, ...): ONE launch
// replaces the ~12 foreach kernels x tensors of torch.optim.Adam.  Same arithmetic as torch's default
// (non-amsgrad, no weight decay) single-tensor Adam, the optimiser the reference constructs
// (src/lightning/RENI_module.py:191-192: Adam(params, lr) -- the configured betas are never passed):
//     m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;
//     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step count t lives in device memory (the kernel uses *step + 1, reni_adam_advance_kernel increments it), so a
// captured CUDA graph replays correctly.  Dense semantics: rows of the latent table whose gradient is zero still move
// by their momentum, exactly as with the reference's dense Adam.
// ------------------------------------------------------------------------------------------------
constexpr int kAdamMaxSegments = 24;
struct AdamParams {
  float* p[kAdamMaxSegments];
  const float* g[kAdamMaxSegments];
  float* m[kAdamMaxSegments];
  float* v[kAdamMaxSegments];
  int64_t end[kAdamMaxSegments];  // exclusive prefix sums of the segment lengths
  int nseg;
  const int* step;                // device: completed steps
  double lr, beta1, beta2, eps;   // (doubles: torch forms 1 - beta^t and lr / (1 - beta1^t) in double before it casts)
};

__global__ void __launch_bounds__(256) reni_adam_kernel(const AdamParams a) {
  const int t = *a.step + 1;
  const float step_size = (float)(a.lr / (1.0 - pow(a.beta1, (double)t)));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2, (double)t));
  const float w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2), b2 = (float)a.beta2, eps = (float)a.eps;
  const int64_t total = a.end[a.nseg - 1];
  int seg = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    while (i >= a.end[seg]) ++seg;  // (monotone in i: grid-stride loop)
    const int64_t j = i - (seg ? a.end[seg - 1] : 0);
    const float g = a.g[seg][j];
    float m = a.m[seg][j], v = a.v[seg][j];
    m = m + (g - m) * w1;                              // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + w2 * g * g;                           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    a.m[seg][j] = m;
    a.v[seg][j] = v;
    a.p[seg][j] -= step_size * (m / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

__global__ void reni_adam_advance_kernel(int* step) { *step += 1; }

// ------------------------------------------------------------------------------------------------
// FiLM per-map stage, forward only (inference / no-grad decoding): everything RENIAutoDecoderFiLM.forward computes
// that is constant per map (RENI.py:405-452 mapping input, :481-512 mapping network, :667 freq = 15 raw + 30) plus the
// hoisted first FiLM layer.  A decode of a single latent otherwise spends its time in ~25 tiny torch launches; here it
// is 2 + n_linears launches that spread every dense layer over (out / 32) x B CTAs (a single CTA per map is bound by
// the bytes one SM can keep in flight: measured 390 us for one map).
//   mc   (B, 5, 256): rows 0..3 = freq_0 * M_b, row 4 = freq_0 * b_0 + phase_0      film (B, L, 2, 256), L = Lf - 1
// ------------------------------------------------------------------------------------------------
constexpr int kFilmMapMaxLinears = 8;

// mapping input per map (RENI.py:424-435 SO2: [vec(Z_xz Z_xz^T), Z_y]; :407-415 SO3: vec(Z Z^T)); grid (B)
__global__ void __launch_bounds__(256) reni_film_map_input_kernel(const float* Z, float* x, int N, int so2, int mn_in) {
  extern __shared__ float s_fz[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_fz[i] = Z[(size_t)b * 3 * N + i];
  __syncthreads();
  for (int i = threadIdx.x; i < mn_in; i += blockDim.x) {
    float v;
    if (i < N * N) {
      const int n = i / N, m = i % N;
      v = s_fz[n * 3] * s_fz[m * 3] + s_fz[n * 3 + 2] * s_fz[m * 3 + 2];
      if (!so2) v = fmaf(s_fz[n * 3 + 1], s_fz[m * 3 + 1], v);
    } else {
      v = s_fz[(i - N * N) * 3 + 1];
    }
    x[(size_t)b * mn_in + i] = v;
  }
}

// y[b, o] = act(b[o] + sum_k W[o, k] x[b, k]); grid (ceil(out / 32), ceil(B / kMaps)), 256 threads = 8 warps x 4 output
// rows each, lanes stride the input (coalesced weight reads, 4 rows x 2 loads in flight per lane), x of the block's kMaps
// maps staged in shared memory.  kMaps = 4 reads every weight once for four maps: the weight matrices are re-read from L2
// by every block row, 84 MB for the 2560 x 256 output layer at 32 maps with kMaps = 1.
template <int kMaps>
__global__ void __launch_bounds__(256) reni_film_map_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   int B, int in, int out, int leaky) {
  extern __shared__ float s_fx[];  // [kMaps][in]
  const int b0 = blockIdx.y * kMaps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kMaps * in; i += blockDim.x) {
    const int m = i / in, k = i - m * in;
    s_fx[i] = (b0 + m < B) ? x[(size_t)(b0 + m) * in + k] : 0.f;
  }
  __syncthreads();
  const int o0 = blockIdx.x * 32 + warp * 4;
  float acc[4][kMaps];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < kMaps; ++m) acc[r][m] = 0.f;
  const float* w[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = W + (size_t)min(o0 + r, out - 1) * in;
  int k = lane;
  for (; k + 32 < in; k += 64) {
    float a[4], c[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { a[r] = __ldg(w[r] + k); c[r] = __ldg(w[r] + k + 32); }
#pragma unroll
    for (int m = 0; m < kMaps; ++m) {
      const float x0 = s_fx[m * in + k], x1 = s_fx[m * in + k + 32];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][m] = fmaf(c[r], x1, fmaf(a[r], x0, acc[r][m]));
    }
  }
  for (; k < in; k += 32) {
    float a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = __ldg(w[r] + k);
#pragma unroll
    for (int m = 0; m < kMaps; ++m) {
      const float x0 = s_fx[m * in + k];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][m] = fmaf(a[r], x0, acc[r][m]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int m = 0; m < kMaps; ++m) {
      float v = acc[r][m];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
      if (lane == 0 && o0 + r < out && b0 + m < B) {
        v += bias[o0 + r];
        y[(size_t)(b0 + m) * out + o0 + r] = (leaky && v < 0.f) ? 0.2f * v : v;
      }
    }
  }
}

// (A fully batched variant -- warp = output row, 32 per-map partial sums per lane against the maps' inputs staged in
// shared memory in chunks, every weight read once for 32 maps -- measured SLOWER than the per-map kernel for 8 and 32 maps
// (decode of 32 latents 211 -> 254 us): the first layer's 1332-column rows leave it only out / 8 = 32 blocks that each
// walk six staged chunks in sequence.  Removed again; kMaps = 4 above keeps the grid wide.)
struct FilmMapFinishParams {
  const float* Z;    // (B, N, 3)
  const float* W0;   // (256, in0): in0 = 2 + N (SO2: [|d_xz|, d_y, innerprod]) or N (SO3)
  const float* b0;   // (256)
  const float* raw;  // (B, 2 * Lf * 256): mapping-network output [frequencies | phase_shifts]
  float* mc;
  float* film;
  int N, so2, Lf;
};

// freq = 15 raw + 30 (RENI.py:667), film for the hidden layers, hoisted + modulated first layer; grid (B), thread = feature
__global__ void __launch_bounds__(256) reni_film_map_finish_kernel(const FilmMapFinishParams p) {
  extern __shared__ float s_fz[];
  const int b = blockIdx.x, N = p.N, j = threadIdx.x;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_fz[i] = p.Z[(size_t)b * 3 * N + i];
  __syncthreads();
  const int half = p.Lf * kH;
  const float* raw = p.raw + (size_t)b * 2 * half;
  for (int l = 1; l < p.Lf; ++l) {
    float* o = p.film + ((size_t)b * (p.Lf - 1) + (l - 1)) * 2 * kH;
    o[j] = fmaf(raw[l * kH + j], 15.f, 30.f);
    o[kH + j] = raw[half + l * kH + j];
  }
  const float f0 = fmaf(raw[j], 15.f, 30.f), ph0 = raw[half + j];
  const int in0 = p.so2 ? N + 2 : N;
  const float* w = p.W0 + (size_t)j * in0;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  if (p.so2) {
    for (int n = 0; n < N; ++n) {
      const float wv = __ldg(w + 2 + n);
      m0 = fmaf(wv, s_fz[n * 3], m0);      // d_x
      m1 = fmaf(wv, s_fz[n * 3 + 2], m1);  // d_z
    }
    m2 = __ldg(w);      // |d_xz|
    m3 = __ldg(w + 1);  // d_y
  } else {
    for (int n = 0; n < N; ++n) {
      const float wv = __ldg(w + n);
      m0 = fmaf(wv, s_fz[n * 3], m0);
      m1 = fmaf(wv, s_fz[n * 3 + 1], m1);
      m2 = fmaf(wv, s_fz[n * 3 + 2], m2);
    }
  }
  float* o = p.mc + (size_t)b * 5 * kH + j;
  o[0] = f0 * m0;
  o[kH] = f0 * m1;
  o[2 * kH] = f0 * m2;
  o[3 * kH] = f0 * m3;
  o[4 * kH] = fmaf(f0, p.b0[j], ph0);
}

// ------------------------------------------------------------------------------------------------
// FiLM per-map stage, BACKWARD (training / latent fitting): gradients of everything reni_film_map_forward computes,
// from d_mc (B, 5, 256) and d_film (B, L, 2, 256) as the core returns them, hand-derived (checked against autograd of
// reni_b200.film.map_level in fp64) and run as 3 + 2 n_linears small launches instead of ~80 autograd kernels:
//   head   : d_raw = [15 d_freq | d_phase] (freq_0 collects the hoisted first layer: sum_r d_mc[r] M[r] + d_mc[4] b_0),
//            dM[r] = d_mc[r] freq_0
//   w0     : dW_0, db_0 += sums over the maps of dM (x) Z and d_mc[4] freq_0
//   per linear i = n-1..0 (y_i = act(W_i x_i + b_i), LeakyReLU(0.2) except the last, RENI.py:481-512):
//            dpre = dY * act'(y);  dW_i += dpre^T x_i;  db_i += sum_b dpre;  dX = dpre W_i
//   dz     : dZ = hoisted-layer part (dM W_ip) + mapping-input part ((dG + dG^T) Z_xz, d Z_y)
// Parameter gradients are ACCUMULATED into caller buffers (views of the flat all-reduce buffer), dZ is written.
// ------------------------------------------------------------------------------------------------
struct FilmMapBwdHeadParams {
  const float* Z;
  const float* W0;
  const float* b0;
  const float* raw;     // (B, 2 Lf 256) mapping-network output saved by the forward
  const float* d_mc;    // (B, 5, 256)
  const float* d_film;  // (B, Lf - 1, 2, 256)
  float* d_raw;         // (B, 2 Lf 256), written
  float* dM;            // (B, 4, 256), written
  int N, so2, Lf;
};

// grid (B), thread = feature j
__global__ void __launch_bounds__(256) reni_film_map_bwd_head_kernel(const FilmMapBwdHeadParams p) {
  extern __shared__ float s_fz[];
  const int b = blockIdx.x, N = p.N, j = threadIdx.x;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_fz[i] = p.Z[(size_t)b * 3 * N + i];
  __syncthreads();
  const int half = p.Lf * kH;
  const float* raw = p.raw + (size_t)b * 2 * half;
  float* d_raw = p.d_raw + (size_t)b * 2 * half;
  for (int l = 1; l < p.Lf; ++l) {
    const float* g = p.d_film + ((size_t)b * (p.Lf - 1) + (l - 1)) * 2 * kH;
    d_raw[l * kH + j] = 15.f * g[j];
    d_raw[half + l * kH + j] = g[kH + j];
  }
  const float f0 = fmaf(raw[j], 15.f, 30.f);
  const int in0 = p.so2 ? N + 2 : N;
  const float* w = p.W0 + (size_t)j * in0;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  if (p.so2) {
    for (int n = 0; n < N; ++n) {
      const float wv = __ldg(w + 2 + n);
      m0 = fmaf(wv, s_fz[n * 3], m0);
      m1 = fmaf(wv, s_fz[n * 3 + 2], m1);
    }
    m2 = __ldg(w);
    m3 = __ldg(w + 1);
  } else {
    for (int n = 0; n < N; ++n) {
      const float wv = __ldg(w + n);
      m0 = fmaf(wv, s_fz[n * 3], m0);
      m1 = fmaf(wv, s_fz[n * 3 + 1], m1);
      m2 = fmaf(wv, s_fz[n * 3 + 2], m2);
    }
  }
  const float* g = p.d_mc + (size_t)b * 5 * kH + j;
  const float g0 = g[0], g1 = g[kH], g2 = g[2 * kH], g3 = g[3 * kH], g4 = g[4 * kH];
  d_raw[j] = 15.f * fmaf(g0, m0, fmaf(g1, m1, fmaf(g2, m2, fmaf(g3, m3, g4 * p.b0[j]))));
  d_raw[half + j] = g4;
  float* dM = p.dM + (size_t)b * 4 * kH + j;
  dM[0] = g0 * f0;
  dM[kH] = g1 * f0;
  dM[2 * kH] = g2 * f0;
  dM[3 * kH] = g3 * f0;
}

struct FilmMapBwdW0Params {
  const float* Z;
  const float* raw;
  const float* d_mc;
  const float* dM;
  float* dW0;  // (256, in0), accumulated
  float* db0;  // (256), accumulated
  int B, N, so2, Lf;
};

// grid (256 / 8), 256 threads: warp = feature j, lanes stride the input columns of W_0; sums over the maps with the
// loads of eight maps in flight
__global__ void __launch_bounds__(256) reni_film_map_bwd_w0_kernel(const FilmMapBwdW0Params p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp, N = p.N;
  const int in0 = p.so2 ? N + 2 : N;
  for (int col = lane; col < in0; col += 32) {
    float acc = 0.f;
    for (int b0 = 0; b0 < p.B; b0 += 8) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u;
        t[u] = 0.f;
        if (b < p.B) {
          const float* dM = p.dM + (size_t)b * 4 * kH + j;
          const float* z = p.Z + (size_t)b * 3 * N;
          if (p.so2) {
            if (col == 0) t[u] = __ldg(dM + 2 * kH);
            else if (col == 1) t[u] = __ldg(dM + 3 * kH);
            else t[u] = fmaf(__ldg(dM), __ldg(z + (col - 2) * 3), __ldg(dM + kH) * __ldg(z + (col - 2) * 3 + 2));
          } else {
            t[u] = fmaf(__ldg(dM), __ldg(z + col * 3),
                        fmaf(__ldg(dM + kH), __ldg(z + col * 3 + 1), __ldg(dM + 2 * kH) * __ldg(z + col * 3 + 2)));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += t[u];
    }
    p.dW0[(size_t)j * in0 + col] += acc;
  }
  {  // db_0[j] += sum_b d_mc[b][4][j] * freq_0[b][j]: lanes stride the maps
    float acc = 0.f;
    const int half = p.Lf * kH;
    for (int b = lane; b < p.B; b += 32)
      acc = fmaf(__ldg(p.d_mc + (size_t)b * 5 * kH + 4 * kH + j), fmaf(__ldg(p.raw + (size_t)b * 2 * half + j), 15.f, 30.f),
                 acc);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (lane == 0) p.db0[j] += acc;
  }
}

// dW[o, k] += sum_b dpre[b, o] x[b, k],  db[o] += sum_b dpre[b, o],  dpre = dY * act'(y)
// grid (ceil(in / 256), ceil(out / 8)), 256 threads: thread = input column k with x[b, k] of 32 maps in registers (all
// 32 loads in flight at once), block = 8 output rows whose dpre is staged in shared memory
constexpr int kMapBwdRows = 8;
__global__ void __launch_bounds__(256) reni_film_map_bwd_dw_kernel(const float* __restrict__ dY, const float* __restrict__ y,
                                                                   const float* __restrict__ x, float* __restrict__ dW,
                                                                   float* __restrict__ db, int B, int in, int out,
                                                                   int leaky) {
  __shared__ float s_g[32][kMapBwdRows];
  const int o0 = blockIdx.y * kMapBwdRows, k = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[kMapBwdRows], accb = 0.f;
#pragma unroll
  for (int r = 0; r < kMapBwdRows; ++r) acc[r] = 0.f;
  for (int b0 = 0; b0 < B; b0 += 32) {
    __syncthreads();
    {
      const int bb = threadIdx.x / kMapBwdRows, r = threadIdx.x % kMapBwdRows;  // 32 x 8 = 256 entries
      float g = 0.f;
      if (b0 + bb < B && o0 + r < out) {
        g = __ldg(dY + (size_t)(b0 + bb) * out + o0 + r);
        if (leaky && __ldg(y + (size_t)(b0 + bb) * out + o0 + r) < 0.f) g *= 0.2f;
      }
      s_g[bb][r] = g;
    }
    float xr[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) xr[u] = (k < in && b0 + u < B) ? __ldg(x + (size_t)(b0 + u) * in + k) : 0.f;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 32; ++u) {
#pragma unroll
      for (int r = 0; r < kMapBwdRows; ++r) acc[r] = fmaf(s_g[u][r], xr[u], acc[r]);
    }
    if (blockIdx.x == 0 && threadIdx.x < kMapBwdRows) {
      for (int u = 0; u < 32; ++u) accb += s_g[u][threadIdx.x];
    }
  }
  if (k < in) {
#pragma unroll
    for (int r = 0; r < kMapBwdRows; ++r)
      if (o0 + r < out) dW[(size_t)(o0 + r) * in + k] += acc[r];
  }
  if (blockIdx.x == 0 && threadIdx.x < kMapBwdRows && o0 + (int)threadIdx.x < out) db[o0 + threadIdx.x] += accb;
}

// dX[b, k] += sum_{o in this block's slice of 32} dpre[b, o] W[o, k] for 32 maps at once: grid (ceil(in / 128),
// ceil(out / 32), ceil(B / 32)), 128 threads; thread = input column k (coalesced weight rows, each weight read once for
// all 32 maps), 32 accumulators in registers, dpre of the slice in shared memory; the slices add up with atomics
// (dX zeroed by the caller)
__global__ void __launch_bounds__(128) reni_film_map_bwd_dx_kernel(const float* __restrict__ dY, const float* __restrict__ y,
                                                                   const float* __restrict__ W, float* __restrict__ dX,
                                                                   int B, int in, int out, int leaky) {
  __shared__ __align__(16) float s_g[32][32];  // [o][b]
  const int o0 = blockIdx.y * 32, b0 = blockIdx.z * 32, k = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int bb = i >> 5, o = i & 31;  // consecutive threads read consecutive o of one map
    float g = 0.f;
    if (b0 + bb < B && o0 + o < out) {
      g = __ldg(dY + (size_t)(b0 + bb) * out + o0 + o);
      if (leaky && __ldg(y + (size_t)(b0 + bb) * out + o0 + o) < 0.f) g *= 0.2f;
    }
    s_g[o][bb] = g;
  }
  __syncthreads();
  if (k >= in) return;
  float acc[32];
#pragma unroll
  for (int u = 0; u < 32; ++u) acc[u] = 0.f;
  const int no = min(32, out - o0);
  for (int oc = 0; oc < no; oc += 8) {
    float w[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) w[t] = (oc + t < no) ? __ldg(W + (size_t)(o0 + oc + t) * in + k) : 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float4* g4 = reinterpret_cast<const float4*>(s_g[oc + t]);
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
  for (int u = 0; u < 32; ++u)
    if (b0 + u < B) atomicAdd(dX + (size_t)(b0 + u) * in + k, acc[u]);
}

struct FilmMapBwdDzParams {
  const float* Z;
  const float* W0;
  const float* dM;   // (B, 4, 256)
  const float* dx0;  // (B, mn_in): gradient w.r.t. the mapping input
  float* dZ;         // (B, N, 3), written
  int N, so2;
};

// grid (B, ceil(3 N / 8)), 256 threads: warp = (latent row n, component c), lanes split the two reductions
__global__ void __launch_bounds__(256) reni_film_map_bwd_dz_kernel(const FilmMapBwdDzParams p) {
  extern __shared__ float s_fz[];
  const int b = blockIdx.x, N = p.N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_fz[i] = p.Z[(size_t)b * 3 * N + i];
  __syncthreads();
  const int i = blockIdx.y * 8 + warp;
  if (i >= 3 * N) return;
  const int in0 = p.so2 ? N + 2 : N;
  const int mn_in = p.so2 ? N * N + N : N * N;
  const float* dx0 = p.dx0 + (size_t)b * mn_in;
  const float* dM = p.dM + (size_t)b * 4 * kH;
  const int n = i / 3, c = i % 3;
  float acc = 0.f;
  if (p.so2 && c == 1) {
    if (lane == 0) acc = dx0[N * N + n];  // Z_y enters the mapping input as it is
  } else {
    // hoisted first layer: dZ[n, c] = sum_j dM[r(c)][j] W_ip[j, n]
    const float* dm = dM + (p.so2 ? (c == 0 ? 0 : 1) : c) * kH;
    const float* w = p.W0 + (p.so2 ? 2 : 0) + n;
#pragma unroll
    for (int jj = 0; jj < kH / 32; ++jj) {
      const int j = jj * 32 + lane;
      acc = fmaf(__ldg(dm + j), __ldg(w + (size_t)j * in0), acc);
    }
    // mapping input G = sum over its components of Z_c Z_c^T: dZ[n, c] += sum_m (dG[n, m] + dG[m, n]) Z[m, c]
    for (int m = lane; m < N; m += 32) acc = fmaf(__ldg(dx0 + n * N + m) + __ldg(dx0 + m * N + n), s_fz[m * 3 + c], acc);
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
  if (lane == 0) p.dZ[(size_t)b * 3 * N + i] = acc;
}

}  // namespace reni
