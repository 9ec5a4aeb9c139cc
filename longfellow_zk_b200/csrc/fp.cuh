// Montgomery prime fields on the sm_100a integer pipes.
//
// Same fields, element encoding and results as the reference's FpGeneric
// (lib/algebra/fp_generic.h:36-568): an element is W 32-bit little-endian limbs
// holding x * 2^(32 W) mod p ("Montgomery form", fp_generic.h:66-67 with 64-bit
// limbs -- the same integer); add/sub keep the canonical range [0, p)
// (fp_generic.h:161-182); mul is the CIOS Montgomery product with the final
// conditional subtraction (fp_generic.h:187-198,484-519).  REDC is a function,
// so the generic reduction step (lib/algebra/fp.h:32-53) and the special
// multiply-free steps for P-256 (lib/algebra/fp_p256.h:42-62) and
// 2^128-2^108+1 (lib/algebra/fp_p128.h:68-75) all return the same limbs.
// Wire encoding = de-Montgomerised little-endian bytes (fp_generic.h:378-380).
//
// Limb products are 32x32->64 (IMAD.WIDE) folded into 64-bit carry words; the
// chains live entirely in registers.
#pragma once
#include <stdint.h>

namespace lf {

template <int W>
struct alignas(16) fpw {
  uint32_t w[W];
};

// Per-field constants (host-computed, copied to __constant__ memory).
template <int W>
struct FpConsts {
  uint32_t m[W];        // modulus
  uint32_t mprime;      // -m^{-1} mod 2^32
  uint32_t rsq[W];      // 2^(64 W) mod m : to_montgomery multiplier
  uint32_t one[W];      // Montgomery one
  uint32_t evalpt[3][W];      // poly_evaluation_point(0..2) = 0,1,2 (fp_generic.h:117-124)
  uint32_t newton[3][3][W];   // newton_denominator(k,i) = 1/i   (fp_generic.h:407-413)
  uint32_t lag_id[3][3][W];   // Newton form of the Lagrange basis (poly.h:125-137)
  uint32_t exact_bits;        // fp_generic.h:93-96
};

#define LF_HDI __host__ __device__ __forceinline__

template <int W>
LF_HDI bool fp_geq(const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = W - 1; i >= 0; --i) {
    if (a[i] > b[i]) return true;
    if (a[i] < b[i]) return false;
  }
  return true;
}
// r = a - b, returns borrow
template <int W>
LF_HDI uint32_t fp_subn(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  int64_t c = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    c += (int64_t)a[i] - (int64_t)b[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
  return (uint32_t)(c & 1);
}
template <int W>
LF_HDI uint32_t fp_addn(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    c += (uint64_t)a[i] + b[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
  return (uint32_t)c;
}

template <int W>
LF_HDI fpw<W> fp_add(const fpw<W>& a, const fpw<W>& b, const uint32_t* m) {
  fpw<W> r, s;
  uint32_t c = fp_addn<W>(r.w, a.w, b.w);
  uint32_t br = fp_subn<W>(s.w, r.w, m);
  // r >= m  <=>  carry out or no borrow
  bool take = c != 0 || br == 0;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = take ? s.w[i] : r.w[i];
  return r;
}
template <int W>
LF_HDI fpw<W> fp_sub(const fpw<W>& a, const fpw<W>& b, const uint32_t* m) {
  fpw<W> r, s;
  uint32_t br = fp_subn<W>(r.w, a.w, b.w);
  fp_addn<W>(s.w, r.w, m);
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = br ? s.w[i] : r.w[i];
  return r;
}

// CIOS Montgomery product, generic modulus.
template <int W>
LF_HDI fpw<W> fp_mul_generic(const fpw<W>& a, const fpw<W>& b, const uint32_t* m, uint32_t mprime) {
  uint32_t t[W + 2];
#pragma unroll
  for (int i = 0; i < W + 2; ++i) t[i] = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      uint64_t s = (uint64_t)a.w[j] * b.w[i] + t[j] + c;
      t[j] = (uint32_t)s;
      c = s >> 32;
    }
    uint64_t s = (uint64_t)t[W] + c;
    t[W] = (uint32_t)s;
    t[W + 1] = (uint32_t)(s >> 32);
    uint32_t q = t[0] * mprime;
    c = ((uint64_t)q * m[0] + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < W; ++j) {
      uint64_t s2 = (uint64_t)q * m[j] + t[j] + c;
      t[j - 1] = (uint32_t)s2;
      c = s2 >> 32;
    }
    s = (uint64_t)t[W] + c;
    t[W - 1] = (uint32_t)s;
    t[W] = t[W + 1] + (uint32_t)(s >> 32);
  }
  fpw<W> r, sb;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = t[i];
  uint32_t br = fp_subn<W>(sb.w, r.w, m);
  bool take = t[W] != 0 || br == 0;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = take ? sb.w[i] : r.w[i];
  return r;
}

// P-256: p = 2^256 - 2^224 + 2^192 + 2^96 - 1, mprime = 1, so the reduction
// step t += t[0]*p needs no multiplication (lib/algebra/fp_p256.h:54-61):
//   t + q*p = t - q + q*2^96 + q*2^192 - q*2^224 + q*2^256,   q = t[0]
LF_HDI fpw<8> fp_mul_p256(const fpw<8>& a, const fpw<8>& b, const uint32_t* m) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint64_t s = (uint64_t)a.w[j] * b.w[i] + t[j] + c;
      t[j] = (uint32_t)s;
      c = s >> 32;
    }
    uint64_t s = (uint64_t)t[8] + c;
    t[8] = (uint32_t)s;
    t[9] = (uint32_t)(s >> 32);
    const int64_t q = t[0];
    int64_t cc = 0;  // limb 0 becomes t[0] - q = 0 with no borrow
#pragma unroll
    for (int k = 1; k < 10; ++k) {
      int64_t v = (int64_t)t[k] + cc;
      if (k == 3 || k == 6 || k == 8) v += q;
      if (k == 7) v -= q;
      t[k - 1] = (uint32_t)v;
      cc = v >> 32;
    }
    t[9] = (uint32_t)cc;
  }
  fpw<8> r, sb;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.w[i] = t[i];
  uint32_t br = fp_subn<8>(sb.w, r.w, m);
  bool take = t[8] != 0 || br == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.w[i] = take ? sb.w[i] : r.w[i];
  return r;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// Device product: W x W limbs -> 2W limbs with carry-chained wide multiply-adds.
// Within a row the products a_j * b_i for even j sit side by side as 64-bit
// values (and so do the ones for odd j, one limb higher), so one carry chain
//   mad.lo.cc / madc.hi.cc / madc.lo.cc / ... / addc
// adds W/2 of them; ptxas fuses each lo/hi pair into one IMAD.WIDE.U32(.X) with
// the carry in a predicate, i.e. W*W wide multiplies and almost no separate adds.
template <int N>
__device__ __forceinline__ void fp_mad_chain(uint32_t* acc, const uint32_t* a /* stride 2 */, uint32_t b);
template <>
__device__ __forceinline__ void fp_mad_chain<4>(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(acc[8])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
}
template <>
__device__ __forceinline__ void fp_mad_chain<2>(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
      : "r"(a[0]), "r"(a[2]), "r"(b));
}
template <>
__device__ __forceinline__ void fp_mad_chain<1>(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2])
      : "r"(a[0]), "r"(b));
}

// t[0..2W+1] = a * b (two spare limbs absorb chain carries; they end up zero).
// One accumulator: for every row the even-j chain covers limbs i..i+W-1 and
// the odd-j chain limbs i+1..i+W, so consecutive chains depend on each other
// and at most ~W/2 carries are live at a time.  (With separate even/odd
// accumulators ptxas interleaves twice as many chains, runs out of the seven
// predicate registers and spills carries through P2R/ISETP.)
template <int W>
__device__ __forceinline__ void fp_mul_wide_dev(const uint32_t* a, const uint32_t* b, uint32_t* t) {
#pragma unroll
  for (int i = 0; i < 2 * W + 2; ++i) t[i] = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    fp_mad_chain<W / 2>(t + i, a, b[i]);
    fp_mad_chain<W / 2>(t + i + 1, a + 1, b[i]);
  }
}

// P-256 Montgomery product: wide product, then the eight multiply-free
// reduction steps of fp_mul_p256 in closed form.  Step i adds q_i * p * 2^(32 i)
// with q_i = (limb i of the running value): -q_i at limb i (clearing it), +q_i
// at limbs i+3, i+6, i+8 and -q_i at limb i+7 (p = 2^256 - 2^224 + 2^192 + 2^96 - 1).
// REDC is a function, so the limbs equal the interleaved form's.
__device__ __forceinline__ fpw<8> fp_mul_p256_dev(const fpw<8>& a, const fpw<8>& b, const uint32_t* m) {
  uint32_t t[18];
  fp_mul_wide_dev<8>(a.w, b.w, t);
  uint32_t q[8];
  int64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t v = (int64_t)t[i] + c;
    if (i >= 3) v += q[i - 3];
    if (i >= 6) v += q[i - 6];
    if (i >= 7) v -= q[i - 7];
    q[i] = (uint32_t)v;
    c = v >> 32;  // (v - q_i) / 2^32
  }
  fpw<8> r, sb;
#pragma unroll
  for (int k = 8; k < 16; ++k) {
    int64_t v = (int64_t)t[k] + c + q[k - 8];
    if (k - 3 < 8) v += q[k - 3];
    if (k - 6 < 8) v += q[k - 6];
    if (k - 7 < 8) v -= q[k - 7];
    r.w[k - 8] = (uint32_t)v;
    c = v >> 32;
  }
  uint32_t br = fp_subn<8>(sb.w, r.w, m);
  bool take = c != 0 || br == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.w[i] = take ? sb.w[i] : r.w[i];
  return r;
}

// Generic modulus: wide product, then W reduction rows q_i = t[i] * mprime,
// t += q_i * m * 2^(32 i), with the same carry-chained rows; limb i becomes
// zero and its carry travels up the chain.
template <int W>
__device__ __forceinline__ fpw<W> fp_mul_generic_dev(const fpw<W>& a, const fpw<W>& b, const uint32_t* m,
                                                     uint32_t mprime) {
  uint32_t t[2 * W + 2];
  fp_mul_wide_dev<W>(a.w, b.w, t);
  uint32_t mm[W];
#pragma unroll
  for (int i = 0; i < W; ++i) mm[i] = m[i];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const uint32_t q = t[i] * mprime;
    fp_mad_chain<W / 2>(t + i, mm, q);
    fp_mad_chain<W / 2>(t + i + 1, mm + 1, q);
  }
  fpw<W> r, sb;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = t[W + i];
  uint32_t br = fp_subn<W>(sb.w, r.w, m);
  bool take = t[2 * W] != 0 || br == 0;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = take ? sb.w[i] : r.w[i];
  return r;
}
#endif  // __CUDACC__

}  // namespace lf

// ---------------------------------------------------------------------------
// host-side construction of the constant block (runs the same limb code)
// ---------------------------------------------------------------------------
namespace lf {

template <int W>
inline fpw<W> fp_pow_host(fpw<W> a, const uint32_t* e, const FpConsts<W>& C) {
  fpw<W> r;
  for (int i = 0; i < W; ++i) r.w[i] = C.one[i];
  for (int i = 0; i < 32 * W; ++i) {
    if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul_generic<W>(r, a, C.m, C.mprime);
    a = fp_mul_generic<W>(a, a, C.m, C.mprime);
  }
  return r;
}

template <int W>
inline void fp_build_consts(const uint32_t* modulus, FpConsts<W>* C) {
  for (int i = 0; i < W; ++i) C->m[i] = modulus[i];
  uint32_t inv = 1;
  for (int i = 0; i < 5; ++i) inv *= 2 - modulus[0] * inv;  // Newton: m^{-1} mod 2^32
  C->mprime = 0u - inv;
  C->exact_bits = 32 * W;
  while (((modulus[(C->exact_bits - 1) >> 5] >> ((C->exact_bits - 1) & 31)) & 1) == 0) --C->exact_bits;
  fpw<W> r;
  for (int i = 0; i < W; ++i) r.w[i] = 0;
  r.w[0] = 1;
  // (assumes m > 1) R mod m and R^2 mod m by repeated doubling (fp_generic.h:105-108)
  for (int i = 0; i < 32 * W; ++i) r = fp_add<W>(r, r, C->m);
  for (int i = 0; i < W; ++i) C->one[i] = r.w[i];
  for (int i = 0; i < 32 * W; ++i) r = fp_add<W>(r, r, C->m);
  for (int i = 0; i < W; ++i) C->rsq[i] = r.w[i];
  fpw<W> one, zero, two;
  for (int i = 0; i < W; ++i) {
    one.w[i] = C->one[i];
    zero.w[i] = 0;
  }
  two = fp_add<W>(one, one, C->m);
  const fpw<W> pts[3] = {zero, one, two};
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < W; ++i) C->evalpt[k][i] = pts[k].w[i];
  // inverses of 1 and 2 by Fermat
  uint32_t e[W], twoint[W];
  for (int i = 0; i < W; ++i) twoint[i] = 0;
  twoint[0] = 2;
  fp_subn<W>(e, modulus, twoint);
  fpw<W> inv_small[3] = {zero, one, fp_pow_host<W>(two, e, *C)};
  for (int k = 0; k < 3; ++k)
    for (int i2 = 0; i2 < 3; ++i2)
      for (int i = 0; i < W; ++i) C->newton[k][i2][i] = (i2 >= 1 && i2 <= k) ? inv_small[i2].w[i] : 0;
  for (int k = 0; k < 3; ++k) {
    fpw<W> t[3] = {zero, zero, zero};
    t[k] = one;
    for (int i2 = 1; i2 < 3; ++i2)
      for (int kk = 2; kk >= i2; --kk)
        t[kk] = fp_mul_generic<W>(fp_sub<W>(t[kk], t[kk - 1], C->m), inv_small[i2], C->m, C->mprime);
    for (int i2 = 0; i2 < 3; ++i2)
      for (int i = 0; i < W; ++i) C->lag_id[k][i2][i] = t[i2].w[i];
  }
}

}  // namespace lf
