// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (UBLKCP), tcgen05
// (TMEM alloc / mma / commit / ld), proxy fences.  No CUTLASS dependency.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace reni {

#define DEVINL __device__ __forceinline__

// Ablation switches for bottleneck attribution (tools/build_variants.sh builds timing-only variants; results are
// wrong by construction): 1 = no stash stores, 2 = no MUFU (sin/cos replaced by a multiply), 4 = no tcgen05.mma issue,
// 8 = no phase-stash loads in the delta chain.
#ifndef RENI_ABL
#define RENI_ABL 0
#endif
#ifndef RENI_REMOTE_RELAXED
#define RENI_REMOTE_RELAXED 1
#endif
DEVINL float abl_sin(float x) { return (RENI_ABL & 2) ? x * 0.5f : __sinf(x); }

// sin on the FMA pipe: range reduction to r = a/2pi - round(a/2pi) in [-0.5, 0.5] (magic-number rounding, two FFMA and
// an FADD) and an odd degree-9 minimax polynomial of sin(2 pi r) (max error 6.3e-6, fitted in tools/fit_sin_poly.py).
// Nine FMA-pipe instructions against FMUL + MUFU.SIN: used for a fraction of the epilogue's elements (RENI_POLY_SIN_MASK)
// because the MUFU pipe (16 results/clk/SM) is the busiest unit of the sin epilogue while the FMA pipe has slack.
DEVINL float poly_sin(float a) {
  const float z = fmaf(a, 0.15915494309189535f, 12582912.f);
  const float k = z - 12582912.f;
  const float r = fmaf(a, 0.15915494309189535f, -k);
  const float r2 = r * r;
  float p = 32.7813832286966f;
  p = fmaf(p, r2, -74.47799703355653f);
  p = fmaf(p, r2, 81.36681431178144f);
  p = fmaf(p, r2, -41.33121426664187f);
  p = fmaf(p, r2, 6.283055798352278f);
  return p * r;
}
#ifndef RENI_POLY_SIN_MASK
#define RENI_POLY_SIN_MASK 0  // bit i: element i of every group of 8 columns takes the polynomial
#endif
template <int kI>
DEVINL float epi_sin(float x) {
  if (RENI_ABL & 2) return x * 0.5f;
  return ((RENI_POLY_SIN_MASK >> kI) & 1) ? poly_sin(x) : __sinf(x);
}
DEVINL float abl_cos(float x) { return (RENI_ABL & 2) ? x * 0.5f : __cosf(x); }

DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

DEVINL uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------ bulk async copy (global -> smem)
DEVINL void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// the same with an L2 evict-first policy: for streams that are read exactly once (the weight-gradient GEMM's stash blocks)
DEVINL void bulk_g2s_stream(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// ------------------------------------------------------------------ CTA pairs (thread-block cluster of 2)
DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// every thread of every CTA in the cluster
DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
DEVINL uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// arrive on an mbarrier of another CTA of the cluster (release at cluster scope)
DEVINL void mbar_arrive_remote(uint32_t cluster_addr) {
#if RENI_REMOTE_RELAXED
  // the only data the consumer touches after this signal is read by the tensor core (async proxy) from THIS CTA's
  // shared memory, and fence.proxy.async has already ordered those writes; a cluster-scope release would in addition
  // wait for every earlier global store of the warp (the stash) to be acknowledged
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// wait on a local mbarrier that a peer CTA arrives on (acquire at cluster scope)
DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// TMA tile load of a CTA pair: box (x, y) of `tmap` -> this CTA's smem; the bytes complete on `bar_cluster_addr`,
// which may be an mbarrier of the pair's OTHER CTA (that is what .cta_group::2 adds over the plain bulk copy)
DEVINL void tma2_load_2d(void* smem_dst, const void* tmap, int32_t x, int32_t y, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(bar_cluster_addr)
      : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05 / TMEM
DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
DEVINL void tmem_alloc(uint32_t* smem_result) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
DEVINL void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16/bf16 in, fp32 accumulate), one thread issues.
DEVINL void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (RENI_ABL & 4) return;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- cta_group::2: one tcgen05.mma spans a CTA pair (M = 256: 128 rows from each CTA's A image and TMEM; each CTA
// holds N/2 rows of B).  Issued by one thread of the leader CTA (cluster rank 0); alloc/dealloc by one warp per CTA.
template <uint32_t kCols>
DEVINL void tmem_alloc2(uint32_t* smem_result) {  // whole warp, in both CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
DEVINL void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
DEVINL void umma2_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (RENI_ABL & 4) return;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
DEVINL void umma2_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" 8x16B core matrices).
//   K-major : LBO = byte stride between core matrices along K, SBO = along M/N
//   MN-major: LBO = byte stride between core matrices along K, SBO = along M/N
// (cute::UMMA::make_umma_desc, canonical INTERLEAVE layouts; version field = 1 on sm_100)
DEVINL uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// Instruction descriptor for kind::f16: fp16 A/B, fp32 D.
//   a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4)            // c_format = F32
         | (0u << 7)          // a_format = F16
         | (0u << 10)         // b_format = F16
         | (a_mn << 15)       // a_major
         | (b_mn << 16)       // b_major
         | ((N >> 3) << 17)   // n_dim
         | ((M >> 4) << 24);  // m_dim
}

// ------------------------------------------------------------------ misc
DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

DEVINL uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------ 16-bit phase stash
// The forward kernel stashes ONE 16-bit number per activation: the phase u = frac(a / 2pi) * 65536 of the
// pre-activation a.  sin and cos are both rebuilt from it to ~5e-5 (better than an fp16 rounding of either), so the
// backward kernels need neither an h stash nor a cos stash.
//   encode: mantissa of (a * 65536/2pi + 1.5*2^23) holds round(a * 65536/2pi) in two's complement -> low 16 bits
//   decode: bytes {u, 0x00, 0x4B} are the float 2^23 + u; one FFMA maps it to the angle 2pi*u/65536 in [0, 2pi)
DEVINL uint32_t phase_encode2(float a0, float a1) {
  const uint32_t z0 = __float_as_uint(fmaf(a0, 10430.378350470453f, 12582912.f));
  const uint32_t z1 = __float_as_uint(fmaf(a1, 10430.378350470453f, 12582912.f));
  return __byte_perm(z0, z1, 0x5410);
}
constexpr float kPhaseToAngle = 9.587379924285257e-05f;  // 2pi / 65536 (the offset below is its exact 2^23 multiple)
DEVINL float phase_angle_lo(uint32_t w) {  // angle of the low 16-bit phase of w
  const float f = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610));
  return fmaf(f, kPhaseToAngle, -8388608.f * kPhaseToAngle);
}
DEVINL float phase_angle_hi(uint32_t w) {  // angle of the high 16-bit phase of w
  const float f = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632));
  return fmaf(f, kPhaseToAngle, -8388608.f * kPhaseToAngle);
}

// ------------------------------------------------------------------ analytic equirectangular grid
// get_directions / get_sineweight (src/utils/utils.py:46-78) in closed form from the pixel index: u = (i + 1/2) / (W/2),
// v = (j + 1/2) / (W/2), theta = pi (u - 1), phi = pi v, d = (sin phi sin theta, cos phi, -sin phi cos theta), sw = sin phi.
// With RENI_FLAG_GRID_DIRECTIONS / RENI_FLAG_GRID_SINEWEIGHT the kernels call this instead of loading D / sw, and the
// mask (RENI_module.py:92-94) is one bit per pixel.
DEVINL void grid_point(int pix, int W, float& dx, float& dy, float& dz, float& sinphi) {
  const int j = pix / W, i = pix - j * W;
  const float inv = 2.f / (float)W;
  float st, ct, sp, cp;
  sincospif(((float)i + 0.5f) * inv - 1.f, &st, &ct);
  sincospif(((float)j + 0.5f) * inv, &sp, &cp);
  dx = sp * st;
  dy = cp;
  dz = -sp * ct;
  sinphi = sp;
}
DEVINL float grid_sineweight(int pix, int W, const uint32_t* mask_bits) {
  float sp, cp;
  sincospif(((float)(pix / W) + 0.5f) * (2.f / (float)W), &sp, &cp);
  if (mask_bits != nullptr && !((__ldg(mask_bits + (pix >> 5)) >> (pix & 31)) & 1u)) sp = 0.f;
  return sp;
}

DEVINL float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace reni
