// The phase stash: what the training forward leaves behind for the backward kernels -- ONE number per hidden
// pre-activation a, its phase frac(a / 2pi), from which both h = sin a (weight-gradient operand) and cos a (delta chain)
// are rebuilt (ptx.cuh: phase_encode2 / phase_angle_*).  This header hides its width and layout:
//
//   RENI_PHASE_BITS = 16: 8 columns of a row = one 16-byte record; a tile-layer = two 64-row half images
//                     [kg 32][64 rows][16 B] (stash_off), 64 KB.
//   RENI_PHASE_BITS = 12 (default): 8 columns = 12 bytes (three words); FOUR records (32 columns of a row) are stored together as
//                     three 16-byte planes -- plane p = word p of the four records -- so that every access is a 16-byte
//                     vector and a warp's accesses stay contiguous: half image [quad 8][plane 3][64 rows][16 B], 48 KB
//                     per tile-layer: 25 % fewer phase bytes through all three kernels; resolution 2pi/4096
//                     (rms 4.4e-4 rad, the size of the fp16 rounding of h).  Measured at cfg 2 against 16 bits on one
//                     box: delta chain 293 -> 262 us, weight-gradient GEMM 246 -> 238 us, forward 299 -> 306 us
//                     (packing), step 0.953 -> 0.917 ms, FiLM step 0.990 -> 0.92 ms; gradient error vs the fp64 oracle
//                     5-7e-4 -> 6-9e-4 (bar 1e-2); the forward output is untouched (the stash only feeds the backward).
//
// Every kernel goes through PhaseRec / phase_encode8 / phase_decode8 and the four-record accessors phase_store4 /
// phase_fetch4 / phase_fetch4_smem (kq = 32-column group of the row).
#pragma once
#include "layout.cuh"
#include "ptx.cuh"

#ifndef RENI_PHASE_BITS
#define RENI_PHASE_BITS 12
#endif
static_assert(RENI_PHASE_BITS == 16 || RENI_PHASE_BITS == 12, "RENI_PHASE_BITS: 16 or 12");

namespace reni {

constexpr int kPhaseRecBytes = RENI_PHASE_BITS == 16 ? 16 : 12;              // 8 columns of one row
constexpr int kPhaseHalfBytes = kHalfRows * (kH / 8) * kPhaseRecBytes;       // one 64-row half image: 32768 / 24576
constexpr int kPhaseTileBytes = 2 * kPhaseHalfBytes;                         // one tile-layer: 65536 / 49152

#if RENI_PHASE_BITS == 16
struct PhaseRec { uint4 v; };
DEVINL PhaseRec phase_encode8(const float (&a)[8]) {
  PhaseRec u;
  u.v.x = phase_encode2(a[0], a[1]);
  u.v.y = phase_encode2(a[2], a[3]);
  u.v.z = phase_encode2(a[4], a[5]);
  u.v.w = phase_encode2(a[6], a[7]);
  return u;
}
DEVINL void phase_decode8(const PhaseRec& u, float (&ang)[8]) {
  ang[0] = phase_angle_lo(u.v.x); ang[1] = phase_angle_hi(u.v.x);
  ang[2] = phase_angle_lo(u.v.y); ang[3] = phase_angle_hi(u.v.y);
  ang[4] = phase_angle_lo(u.v.z); ang[5] = phase_angle_hi(u.v.z);
  ang[6] = phase_angle_lo(u.v.w); ang[7] = phase_angle_hi(u.v.w);
}
template <int kI>  // angle of column kI of the record
DEVINL float phase_angle_of(const PhaseRec& u) {
  const uint32_t w = kI < 2 ? u.v.x : (kI < 4 ? u.v.y : (kI < 6 ? u.v.z : u.v.w));
  return (kI & 1) ? phase_angle_hi(w) : phase_angle_lo(w);
}
template <int kHint>  // 0 plain, 1 streaming (.cs), 2 write-through
DEVINL void phase_store4(uint8_t* tile_layer, uint32_t r, uint32_t kq, const PhaseRec (&u)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4* dst = reinterpret_cast<uint4*>(tile_layer + stash_off(r, kq * 4 + j, kH));
    if (kHint == 1) __stcs(dst, u[j].v);
    else if (kHint == 2) __stwt(dst, u[j].v);
    else *dst = u[j].v;
  }
}
template <int kHint>  // 0 ld.global.nc, 1 streaming (.cs), 2 last-use (.lu)
DEVINL void phase_fetch4(const uint8_t* tile_layer, uint32_t r, uint32_t kq, PhaseRec (&u)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4* src = reinterpret_cast<const uint4*>(tile_layer + stash_off(r, kq * 4 + j, kH));
    u[j].v = kHint == 1 ? __ldcs(src) : (kHint == 2 ? __ldlu(src) : __ldg(src));
  }
}
// producer side, one record at a time: record kg of the row (j = kg & 3 inside its 32-column group) is stored at once
template <int kHint>
DEVINL void phase_put(PhaseRec (&)[4], int, uint8_t* tile_layer, uint32_t r, uint32_t kg, const PhaseRec& u) {
  uint4* dst = reinterpret_cast<uint4*>(tile_layer + stash_off(r, kg, kH));
  if (kHint == 1) __stcs(dst, u.v);
  else if (kHint == 2) __stwt(dst, u.v);
  else *dst = u.v;
}
// the same four records of a 64-row half image that sits in shared memory (row r of the half)
DEVINL void phase_fetch4_smem(const uint8_t* half_image, uint32_t r, uint32_t kq, PhaseRec (&u)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) u[j].v = *reinterpret_cast<const uint4*>(half_image + ((kq * 4 + j) * kHalfRows + r) * 16);
}
#else
struct PhaseRec { uint32_t w0, w1, w2; };  // p0 | p1<<12 | p2<<24 ; p2>>8 | p3<<4 | p4<<16 | p5<<28 ; p5>>4 | p6<<8 | p7<<20
DEVINL uint32_t phase_bitsel(uint32_t a, uint32_t b, uint32_t m) { return (a & m) | (b & ~m); }
// encode: the mantissa of (a * 4096/2pi + 1.5*2^23) holds round(a * 4096/2pi) in two's complement -> its low 12 bits
DEVINL PhaseRec phase_encode8(const float (&a)[8]) {
  uint32_t z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(fmaf(a[i], 651.8986469044033f, 12582912.f));
  // bitsel(a, b, m) = (a & m) | (b & ~m) is one LOP3: every field is shifted into place and selected over the garbage
  // above the previous one (the bits of z above its low 12 are the integer part of the turn count)
  PhaseRec u;
  u.w0 = phase_bitsel(phase_bitsel(z[0], z[1] << 12, 0xFFFu), z[2] << 24, 0xFFFFFFu);
  u.w1 = phase_bitsel(phase_bitsel(phase_bitsel(z[2] >> 8, z[3] << 4, 0xFu), z[4] << 16, 0xFFFFu), z[5] << 28, 0xFFFFFFFu);
  u.w2 = phase_bitsel(phase_bitsel(z[5] >> 4, z[6] << 8, 0xFFu), z[7] << 20, 0xFFFFFu);
  return u;
}
constexpr float kPhase12ToAngle = 1.5339807878856412e-03f;  // 2pi / 4096
// bits {p, 0x4B000000} are the float 2^23 + p; one FFMA maps it to the angle 2pi*p/4096 in [0, 2pi)
DEVINL float phase12_angle(uint32_t p) {
  return fmaf(__uint_as_float((p & 0xFFFu) | 0x4B000000u), kPhase12ToAngle, -8388608.f * kPhase12ToAngle);
}
DEVINL void phase_decode8(const PhaseRec& u, float (&ang)[8]) {
  ang[0] = phase12_angle(u.w0);
  ang[1] = phase12_angle(u.w0 >> 12);
  ang[2] = phase12_angle(__funnelshift_r(u.w0, u.w1, 24));
  ang[3] = phase12_angle(u.w1 >> 4);
  ang[4] = phase12_angle(u.w1 >> 16);
  ang[5] = phase12_angle(__funnelshift_r(u.w1, u.w2, 28));
  ang[6] = phase12_angle(u.w2 >> 8);
  ang[7] = phase12_angle(u.w2 >> 20);
}
template <int kI>
DEVINL float phase_angle_of(const PhaseRec& u) {
  return kI == 0 ? phase12_angle(u.w0)
       : kI == 1 ? phase12_angle(u.w0 >> 12)
       : kI == 2 ? phase12_angle(__funnelshift_r(u.w0, u.w1, 24))
       : kI == 3 ? phase12_angle(u.w1 >> 4)
       : kI == 4 ? phase12_angle(u.w1 >> 16)
       : kI == 5 ? phase12_angle(__funnelshift_r(u.w1, u.w2, 28))
       : kI == 6 ? phase12_angle(u.w2 >> 8)
                 : phase12_angle(u.w2 >> 20);
}
// byte offset of plane 0 of quad kq (columns 32 kq .. 32 kq + 31) of row r inside a tile-layer image; planes are
// kHalfRows * 16 bytes apart
DEVINL uint32_t phase_quad_off(uint32_t r, uint32_t kq) {
  return (r / kHalfRows) * kPhaseHalfBytes + (kq * 3 * kHalfRows + (r % kHalfRows)) * 16u;
}
template <int kHint>
DEVINL void phase_store4(uint8_t* tile_layer, uint32_t r, uint32_t kq, const PhaseRec (&u)[4]) {
  uint4* d = reinterpret_cast<uint4*>(tile_layer + phase_quad_off(r, kq));
  const uint4 p0 = make_uint4(u[0].w0, u[1].w0, u[2].w0, u[3].w0);
  const uint4 p1 = make_uint4(u[0].w1, u[1].w1, u[2].w1, u[3].w1);
  const uint4 p2 = make_uint4(u[0].w2, u[1].w2, u[2].w2, u[3].w2);
  if (kHint == 1) { __stcs(d, p0); __stcs(d + kHalfRows, p1); __stcs(d + 2 * kHalfRows, p2); }
  else if (kHint == 2) { __stwt(d, p0); __stwt(d + kHalfRows, p1); __stwt(d + 2 * kHalfRows, p2); }
  else { d[0] = p0; d[kHalfRows] = p1; d[2 * kHalfRows] = p2; }
}
// producer side, one record at a time: the record waits in `buf` until its 32-column group is complete (j = kg & 3)
template <int kHint>
DEVINL void phase_put(PhaseRec (&buf)[4], int j, uint8_t* tile_layer, uint32_t r, uint32_t kg, const PhaseRec& u) {
  buf[j] = u;
  if (j == 3) phase_store4<kHint>(tile_layer, r, kg >> 2, buf);
}
DEVINL void phase_unplane(const uint4& p0, const uint4& p1, const uint4& p2, PhaseRec (&u)[4]) {
  u[0].w0 = p0.x; u[1].w0 = p0.y; u[2].w0 = p0.z; u[3].w0 = p0.w;
  u[0].w1 = p1.x; u[1].w1 = p1.y; u[2].w1 = p1.z; u[3].w1 = p1.w;
  u[0].w2 = p2.x; u[1].w2 = p2.y; u[2].w2 = p2.z; u[3].w2 = p2.w;
}
template <int kHint>
DEVINL void phase_fetch4(const uint8_t* tile_layer, uint32_t r, uint32_t kq, PhaseRec (&u)[4]) {
  const uint4* s = reinterpret_cast<const uint4*>(tile_layer + phase_quad_off(r, kq));
  uint4 p0, p1, p2;
  if (kHint == 1) { p0 = __ldcs(s); p1 = __ldcs(s + kHalfRows); p2 = __ldcs(s + 2 * kHalfRows); }
  else if (kHint == 2) { p0 = __ldlu(s); p1 = __ldlu(s + kHalfRows); p2 = __ldlu(s + 2 * kHalfRows); }
  else { p0 = __ldg(s); p1 = __ldg(s + kHalfRows); p2 = __ldg(s + 2 * kHalfRows); }
  phase_unplane(p0, p1, p2, u);
}
DEVINL void phase_fetch4_smem(const uint8_t* half_image, uint32_t r, uint32_t kq, PhaseRec (&u)[4]) {
  const uint4* s = reinterpret_cast<const uint4*>(half_image + (kq * 3 * kHalfRows + r) * 16u);
  phase_unplane(s[0], s[kHalfRows], s[2 * kHalfRows], u);
}
#endif

}  // namespace reni
