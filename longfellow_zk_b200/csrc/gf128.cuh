// GF(2^128) = GF(2)[x]/(x^128 + x^7 + x^2 + x + 1) on the sm_100a integer pipes.
//
// Same field, element encoding and results as the reference's GF2_128<>
// (lib/gf2k/gf2_128.h:35-246, lib/gf2k/sysdep.h:51-90): bit i of the 128-bit
// word is the coefficient of x^i, low word first; add = XOR.
//
// sm_100a has no carry-less multiply.  The 128x128 product is built from
// 32x32->64 carry-less products, each computed with 16 IMAD.WIDE on operands
// whose bits are spread into every 4th position ("holes"): an 8-bit-populated
// operand pair sums at most 8 partial products per bit position, which fits in
// the 4-bit hole, so the low bit of every hole is the GF(2) sum.  Two Karatsuba
// levels bring the 128x128 product to 9 such 32x32 products (144 IMAD.WIDE +
// LOP3 glue), followed by the two-step fold by 0x87.
#pragma once
#include <stdint.h>

namespace lf {

struct alignas(16) gf128 {
  uint32_t w[4];
};

__host__ __device__ __forceinline__ gf128 gf_zero() {
  gf128 r;
  r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0;
  return r;
}
__host__ __device__ __forceinline__ gf128 gf_one() {
  gf128 r = gf_zero();
  r.w[0] = 1;
  return r;
}
__host__ __device__ __forceinline__ gf128 gf_add(const gf128& a, const gf128& b) {
  gf128 r;
  r.w[0] = a.w[0] ^ b.w[0];
  r.w[1] = a.w[1] ^ b.w[1];
  r.w[2] = a.w[2] ^ b.w[2];
  r.w[3] = a.w[3] ^ b.w[3];
  return r;
}
__host__ __device__ __forceinline__ bool gf_is_zero(const gf128& a) {
  return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0;
}
__host__ __device__ __forceinline__ bool gf_eq(const gf128& a, const gf128& b) {
  return ((a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) | (a.w[3] ^ b.w[3])) == 0;
}

// 32x32 -> 64 carry-less product (16 wide multiplies).
__host__ __device__ __forceinline__ uint64_t clmul32(uint32_t x, uint32_t y) {
  const uint32_t m0 = 0x11111111u, m1 = 0x22222222u, m2 = 0x44444444u, m3 = 0x88888888u;
  uint32_t x0 = x & m0, x1 = x & m1, x2 = x & m2, x3 = x & m3;
  uint32_t y0 = y & m0, y1 = y & m1, y2 = y & m2, y3 = y & m3;
#define LF_WM(a, b) ((uint64_t)(a) * (uint64_t)(b))
  // x_i * y_j puts its partial products into the holes at bit 4h + (i + j): hole h holds at most
  // min(h + 1, 15 - h) <= 8 of them when i + j < 4, and the same profile one hole later when
  // i + j >= 4 (its hole h is the other's h - 1).  So a product of each kind can be ADDED: no hole
  // ever sums more than 8 + 7 = 15 terms and none carries into its neighbour -- and the add is the
  // multiplier's free accumulate input (IMAD.WIDE a * b + c), where an XOR is two LOP3.  Products
  // of the same kind could reach 16 in hole 7 (or 8) and still take the XOR.
  uint64_t z0 = (LF_WM(x0, y0) + LF_WM(x1, y3)) ^ LF_WM(x2, y2) ^ LF_WM(x3, y1);
  uint64_t z1 = (LF_WM(x0, y1) + LF_WM(x2, y3)) ^ (LF_WM(x1, y0) + LF_WM(x3, y2));
  uint64_t z2 = (LF_WM(x0, y2) + LF_WM(x3, y3)) ^ LF_WM(x1, y1) ^ LF_WM(x2, y0);
  uint64_t z3 = LF_WM(x0, y3) ^ LF_WM(x1, y2) ^ LF_WM(x2, y1) ^ LF_WM(x3, y0);
#undef LF_WM
  const uint64_t M0 = 0x1111111111111111ull, M1 = 0x2222222222222222ull,
                 M2 = 0x4444444444444444ull, M3 = 0x8888888888888888ull;
  return (z0 & M0) | (z1 & M1) | (z2 & M2) | (z3 & M3);
}

// 64x64 -> 128 by one Karatsuba level over 32-bit halves: r[0..3]
__host__ __device__ __forceinline__ void clmul64(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1,
                                                 uint32_t r[4]) {
  uint64_t lo = clmul32(a0, b0);
  uint64_t hi = clmul32(a1, b1);
  uint64_t mid = clmul32(a0 ^ a1, b0 ^ b1) ^ lo ^ hi;
  r[0] = (uint32_t)lo;
  r[1] = (uint32_t)(lo >> 32) ^ (uint32_t)mid;
  r[2] = (uint32_t)hi ^ (uint32_t)(mid >> 32);
  r[3] = (uint32_t)(hi >> 32);
}

// Unreduced 128x128 -> 256-bit product, t[0..7] (two Karatsuba levels).
__host__ __device__ __forceinline__ void gf_mul_wide_inl(const gf128& a, const gf128& b, uint32_t t[8]) {
  uint32_t lo[4], hi[4], mid[4];
  clmul64(a.w[0], a.w[1], b.w[0], b.w[1], lo);
  clmul64(a.w[2], a.w[3], b.w[2], b.w[3], hi);
  clmul64(a.w[0] ^ a.w[2], a.w[1] ^ a.w[3], b.w[0] ^ b.w[2], b.w[1] ^ b.w[3], mid);
#pragma unroll
  for (int i = 0; i < 4; ++i) mid[i] ^= lo[i] ^ hi[i];
  t[0] = lo[0];
  t[1] = lo[1];
  t[2] = lo[2] ^ mid[0];
  t[3] = lo[3] ^ mid[1];
  t[4] = hi[0] ^ mid[2];
  t[5] = hi[1] ^ mid[3];
  t[6] = hi[2];
  t[7] = hi[3];
}

// On the device the ~450-instruction product is ONE out-of-line function: the
// sumcheck kernel multiplies at a dozen sites, and inlining every one of them
// made the kernel body > 100 KB, i.e. instruction-cache bound (ncu: 36 % of the
// stall samples were "no instruction", profiles/r1_sumcheck_before.txt).
// Arguments and the 8-word result travel in registers (no local memory).
struct gf_wide {
  uint32_t t[8];
};
#if defined(__CUDA_ARCH__) && !defined(LF_GF_INLINE_MUL)
static __device__ __noinline__ gf_wide gf_mul_wide_fn(gf128 a, gf128 b) {
  gf_wide w;
  gf_mul_wide_inl(a, b, w.t);
  return w;
}
__device__ __forceinline__ void gf_mul_wide(const gf128& a, const gf128& b, uint32_t t[8]) {
  gf_wide w = gf_mul_wide_fn(a, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = w.t[i];
}
#else
__host__ __device__ __forceinline__ void gf_mul_wide(const gf128& a, const gf128& b, uint32_t t[8]) {
  gf_mul_wide_inl(a, b, t);
}
#endif

// Reduce a 256-bit polynomial modulo x^128 + x^7 + x^2 + x + 1
// (same two folds as lib/gf2k/sysdep.h:45-66).
__host__ __device__ __forceinline__ gf128 gf_reduce(const uint32_t t[8]) {
  // fold words 4..7 (h) : h * (1 + x + x^2 + x^7), a 135-bit value r[0..4]
  uint32_t h0 = t[4], h1 = t[5], h2 = t[6], h3 = t[7];
  uint32_t r0 = h0 ^ (h0 << 1) ^ (h0 << 2) ^ (h0 << 7);
  uint32_t r1 = h1 ^ (h1 << 1) ^ (h1 << 2) ^ (h1 << 7) ^ (h0 >> 31) ^ (h0 >> 30) ^ (h0 >> 25);
  uint32_t r2 = h2 ^ (h2 << 1) ^ (h2 << 2) ^ (h2 << 7) ^ (h1 >> 31) ^ (h1 >> 30) ^ (h1 >> 25);
  uint32_t r3 = h3 ^ (h3 << 1) ^ (h3 << 2) ^ (h3 << 7) ^ (h2 >> 31) ^ (h2 >> 30) ^ (h2 >> 25);
  uint32_t r4 = (h3 >> 31) ^ (h3 >> 30) ^ (h3 >> 25);  // < 2^7
  r0 ^= r4 ^ (r4 << 1) ^ (r4 << 2) ^ (r4 << 7);
  gf128 o;
  o.w[0] = t[0] ^ r0;
  o.w[1] = t[1] ^ r1;
  o.w[2] = t[2] ^ r2;
  o.w[3] = t[3] ^ r3;
  return o;
}

__host__ __device__ __forceinline__ gf128 gf_mul(const gf128& a, const gf128& b) {
  uint32_t t[8];
  gf_mul_wide(a, b, t);
  return gf_reduce(t);
}
__host__ __device__ __forceinline__ gf128 gf_mul_inl(const gf128& a, const gf128& b) {
  uint32_t t[8];
  gf_mul_wide_inl(a, b, t);
  return gf_reduce(t);
}

// acc (256-bit, unreduced) ^= a*b : lazy reduction for dot products
// (the reference's Accum/mac/reduce, lib/gf2k/sysdep.h:68-90).
__host__ __device__ __forceinline__ void gf_mac(uint32_t acc[8], const gf128& a, const gf128& b) {
  uint32_t t[8];
  gf_mul_wide(a, b, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] ^= t[i];
}

// a * x^-1 ... not needed; inversion by Fermat (a^(2^128-2)); the inverse is
// unique so it equals GF2_128::invertf (lib/gf2k/gf2_128.h:274-310).
__host__ __device__ inline gf128 gf_inv(const gf128& a) {
  gf128 r = gf_one(), s = a;
  for (int i = 1; i < 128; ++i) {
    s = gf_mul(s, s);
    r = gf_mul(r, s);
  }
  return r;
}

}  // namespace lf
