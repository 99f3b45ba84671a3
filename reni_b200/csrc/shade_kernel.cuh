// Blinn-Phong shading of a surface by ALL texels of an environment map -- the downstream consumer of the decoder output
// on the reference's FIT_INVERSE path (src/utils/pytorch3d_envmap_shader.py:85-119, called from RENI_module.get_render,
// :386-396).  The reference materialises (B, H, W, J) weight tensors and a (B, H, W, J, 3) half-vector tensor (1.6 GB
// per map at a 128 x 128 render of a 64 x 128 map); here the weight of light j on pixel p
//     w[p, j] = kd * clamp(n_p . l_j, 0, 1) + c * ks * clamp(n_p . normalize(v_p + l_j), 0, 1)^s ,
//     c = (s + 2) / (4 (2 - exp(-s / 2)))
// lives in a register for the time it takes to use it:
//     forward  : colors[b, p, :] = sum_j w[p, j] * light[b, j, :]           (thread = pixel, texel tiles in smem)
//     backward : d_light[b, j, :] = sum_p w[p, j] * grad_colors[b, p, :]     (thread = texel, pixel tiles in smem)
// fp32 on the CUDA cores; up to kShadeMaps maps share one evaluation of w when they share the direction grid.  The
// reduction axis (texels forward, pixels backward) is split over blockIdx.z so that a 128 x 128 render of ONE map still
// fills the machine (the first version, one block per 128 pixels, ran 128 blocks on 148 SMs: 3.3 ms); the splits
// combine with fp32 reductions into the zeroed output.  HBM traffic is the operands once per split; the work is
// ~25 flops + sqrt + pow per pair.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reni {

constexpr int kShadeThreads = 128;
constexpr int kShadeTile = 128;
constexpr int kShadeMaps = 8;

struct ShadeParams {
  const float* normals;   // (n_pix, 3) unit surface normals
  const float* view;      // (n_pix, 3) unit directions surface -> camera
  const float* D;         // (B or 1, J, 3) unit light directions
  int64_t d_bstride;      // 0: one grid shared by all maps
  const float* light;     // (B, J, 3) light colours (environment map x sine weight)        [forward]
  const float* grad;      // (B, n_pix, 3) gradient w.r.t. the colours                      [backward]
  float* colors;          // (B, n_pix, 3)                                                  [forward]
  float* d_light;         // (B, J, 3)                                                      [backward]
  int B, J, n_pix;
  int split;                 // blocks along the reduction axis (gridDim.z); outputs are zeroed by the launcher
  float kd, cks, shininess;  // cks = c * ks
};

__device__ __forceinline__ float shade_weight(float nx, float ny, float nz, float vx, float vy, float vz, float lx,
                                              float ly, float lz, float kd, float cks, float s) {
  const float d = __saturatef(nx * lx + ny * ly + nz * lz);
  const float hx = vx + lx, hy = vy + ly, hz = vz + lz;
  const float inv = 1.f / fmaxf(sqrtf(hx * hx + hy * hy + hz * hz), 1e-6f);  // F.normalize(eps = 1e-6)
  const float nh = __saturatef((nx * hx + ny * hy + nz * hz) * inv);
  const float sp = cks != 0.f ? powf(nh, s) : 0.f;
  return kd * d + cks * sp;
}

// grid (ceil(n_pix / 128), map chunks); kShared: the chunk's maps share the direction grid and one evaluation of w
template <bool kShared>
__global__ void __launch_bounds__(kShadeThreads) reni_shade_fwd_kernel(const ShadeParams p) {
  constexpr int kM = kShared ? kShadeMaps : 1;
  __shared__ float s_l[kShadeTile * 3];
  __shared__ float s_e[kM * kShadeTile * 3];
  const int pix = blockIdx.x * kShadeThreads + threadIdx.x;
  const int b0 = blockIdx.y * kM;
  const int nm = min(kM, p.B - b0);
  float nx = 0.f, ny = 0.f, nz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
  if (pix < p.n_pix) {
    nx = p.normals[pix * 3]; ny = p.normals[pix * 3 + 1]; nz = p.normals[pix * 3 + 2];
    vx = p.view[pix * 3]; vy = p.view[pix * 3 + 1]; vz = p.view[pix * 3 + 2];
  }
  float acc[kM][3];
#pragma unroll
  for (int m = 0; m < kM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = 0.f;
  const float* Db = p.D + (size_t)b0 * p.d_bstride;
  const int jper = ((p.J + p.split - 1) / p.split + kShadeTile - 1) / kShadeTile * kShadeTile;
  const int jlo = blockIdx.z * jper, jhi = min(p.J, jlo + jper);
  for (int j0 = jlo; j0 < jhi; j0 += kShadeTile) {
    const int nj = min(kShadeTile, jhi - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nj * 3; i += kShadeThreads) s_l[i] = Db[(size_t)j0 * 3 + i];
    for (int m = 0; m < nm; ++m)
      for (int i = threadIdx.x; i < nj * 3; i += kShadeThreads)
        s_e[m * kShadeTile * 3 + i] = p.light[((size_t)(b0 + m) * p.J + j0) * 3 + i];
    __syncthreads();
    for (int j = 0; j < nj; ++j) {
      const float w = shade_weight(nx, ny, nz, vx, vy, vz, s_l[j * 3], s_l[j * 3 + 1], s_l[j * 3 + 2], p.kd, p.cks,
                                   p.shininess);
#pragma unroll
      for (int m = 0; m < kM; ++m) {
        if (m < nm) {
          const float* e = s_e + (m * kShadeTile + j) * 3;
          acc[m][0] = fmaf(w, e[0], acc[m][0]);
          acc[m][1] = fmaf(w, e[1], acc[m][1]);
          acc[m][2] = fmaf(w, e[2], acc[m][2]);
        }
      }
    }
  }
  if (pix < p.n_pix)
    for (int m = 0; m < nm; ++m) {
      float* o = p.colors + ((size_t)(b0 + m) * p.n_pix + pix) * 3;
      if (p.split == 1) { o[0] = acc[m][0]; o[1] = acc[m][1]; o[2] = acc[m][2]; }
      else { atomicAdd(o, acc[m][0]); atomicAdd(o + 1, acc[m][1]); atomicAdd(o + 2, acc[m][2]); }
    }
}

// grid (ceil(J / 128), map chunks): thread = texel, pixel tiles staged in shared memory
template <bool kShared>
__global__ void __launch_bounds__(kShadeThreads) reni_shade_bwd_kernel(const ShadeParams p) {
  constexpr int kM = kShared ? kShadeMaps : 1;
  __shared__ float s_n[kShadeTile * 3];
  __shared__ float s_v[kShadeTile * 3];
  __shared__ float s_g[kM * kShadeTile * 3];
  const int j = blockIdx.x * kShadeThreads + threadIdx.x;
  const int b0 = blockIdx.y * kM;
  const int nm = min(kM, p.B - b0);
  float lx = 0.f, ly = 0.f, lz = 0.f;
  if (j < p.J) {
    const float* d = p.D + (size_t)b0 * p.d_bstride + (size_t)j * 3;
    lx = d[0]; ly = d[1]; lz = d[2];
  }
  float acc[kM][3];
#pragma unroll
  for (int m = 0; m < kM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = 0.f;
  const int pper = ((p.n_pix + p.split - 1) / p.split + kShadeTile - 1) / kShadeTile * kShadeTile;
  const int plo = blockIdx.z * pper, phi = min(p.n_pix, plo + pper);
  for (int p0 = plo; p0 < phi; p0 += kShadeTile) {
    const int np = min(kShadeTile, phi - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < np * 3; i += kShadeThreads) {
      s_n[i] = p.normals[(size_t)p0 * 3 + i];
      s_v[i] = p.view[(size_t)p0 * 3 + i];
    }
    for (int m = 0; m < nm; ++m)
      for (int i = threadIdx.x; i < np * 3; i += kShadeThreads)
        s_g[m * kShadeTile * 3 + i] = p.grad[((size_t)(b0 + m) * p.n_pix + p0) * 3 + i];
    __syncthreads();
    for (int q = 0; q < np; ++q) {
      const float w = shade_weight(s_n[q * 3], s_n[q * 3 + 1], s_n[q * 3 + 2], s_v[q * 3], s_v[q * 3 + 1],
                                   s_v[q * 3 + 2], lx, ly, lz, p.kd, p.cks, p.shininess);
#pragma unroll
      for (int m = 0; m < kM; ++m) {
        if (m < nm) {
          const float* g = s_g + (m * kShadeTile + q) * 3;
          acc[m][0] = fmaf(w, g[0], acc[m][0]);
          acc[m][1] = fmaf(w, g[1], acc[m][1]);
          acc[m][2] = fmaf(w, g[2], acc[m][2]);
        }
      }
    }
  }
  if (j < p.J)
    for (int m = 0; m < nm; ++m) {
      float* o = p.d_light + ((size_t)(b0 + m) * p.J + j) * 3;
      if (p.split == 1) { o[0] = acc[m][0]; o[1] = acc[m][1]; o[2] = acc[m][2]; }
      else { atomicAdd(o, acc[m][0]); atomicAdd(o + 1, acc[m][1]); atomicAdd(o + 2, acc[m][2]); }
    }
}

}  // namespace reni
