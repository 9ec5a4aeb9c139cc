// Field policy classes used to template the kernels (the counterpart of the
// reference's "Field concept", lib/algebra/fp_generic.h:36-568 and
// lib/gf2k/gf2_128.h:35-509).  Each policy provides Elt, add/sub/mul, a lazy
// accumulator (the reference's Accum/mac/reduce), wire conversion and the small
// per-field constant block that the sumcheck needs.
#pragma once
#include <stdint.h>

#include "fp.cuh"
#include "gf128.cuh"
#include "hash.cuh"

namespace lf {

// Per-field constants, filled by the host (lf_engine.cu) and copied to
// __constant__ memory once per context.
struct GfConsts {
  gf128 beta[16];        // GF2_128::beta(i)           gf2_128.h:108-119
  gf128 evalpt[3];       // poly_evaluation_point(0..2) gf2_128.h:121-127
  gf128 newton[3][3];    // newton_denominator(k,i)     gf2_128.h:129-137
  gf128 lag_id[3][3];    // Newton form of the Lagrange basis polys (poly.h:125-137)
  uint32_t sub_u[16][4]; // row-echelon basis (gf2_128.h:451-493)
  uint32_t sub_linv[16];
  uint32_t sub_ldnz[16];
};

#ifdef __CUDACC__
static __constant__ GfConsts c_gf;
#endif

struct FGf128 {
  typedef FGf128 Compact;
  typedef gf128 Elt;
  static constexpr int kWords = 4;      // 32-bit words per element (memory and wire)
  static constexpr int kBytes = 16;
  static constexpr int kSubBytes = 2;
  static constexpr bool kChar2 = true;
  static constexpr int kFieldId = 4;    // proto/circuit_io.h GF2_128_ID

  struct Acc {
    uint32_t t[8];
  };

  __host__ __device__ static __forceinline__ Elt zero() { return gf_zero(); }
  __host__ __device__ static __forceinline__ Elt one() { return gf_one(); }
  __host__ __device__ static __forceinline__ Elt add(const Elt& a, const Elt& b) { return gf_add(a, b); }
  __host__ __device__ static __forceinline__ Elt sub(const Elt& a, const Elt& b) { return gf_add(a, b); }
  __host__ __device__ static __forceinline__ Elt neg(const Elt& a) { return a; }
  __host__ __device__ static __forceinline__ Elt mul(const Elt& a, const Elt& b) { return gf_mul(a, b); }
  __host__ __device__ static __forceinline__ bool is_zero(const Elt& a) { return gf_is_zero(a); }
  __host__ __device__ static __forceinline__ bool eq(const Elt& a, const Elt& b) { return gf_eq(a, b); }

  __host__ __device__ static __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a.t[i] = 0;
  }
  __host__ __device__ static __forceinline__ void mac(Acc& a, const Elt& x, const Elt& y) {
    gf_mac(a.t, x, y);
  }
  // (first ? a0 : a1) += x * y without a divergent call of the multiplier
  __host__ __device__ static __forceinline__ void mac_sel(Acc& a0, Acc& a1, bool first, const Elt& x, const Elt& y) {
    uint32_t t[8];
    gf_mul_wide(x, y, t);
    const uint32_t m0 = first ? 0xffffffffu : 0u, m1 = ~m0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a0.t[i] ^= t[i] & m0;
      a1.t[i] ^= t[i] & m1;
    }
  }
  // a += x (an element, i.e. x * 1)
  __host__ __device__ static __forceinline__ void acc_add_elt(Acc& a, const Elt& x) {
    a.t[0] ^= x.w[0]; a.t[1] ^= x.w[1]; a.t[2] ^= x.w[2]; a.t[3] ^= x.w[3];
  }
  __host__ __device__ static __forceinline__ void acc_add(Acc& a, const Acc& b) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a.t[i] ^= b.t[i];
  }
  __host__ __device__ static __forceinline__ Elt reduce(const Acc& a) { return gf_reduce(a.t); }

  // memory representation == wire representation (16 LE bytes)
  __host__ __device__ static __forceinline__ void to_wire(uint32_t out[4], const Elt& a) {
    out[0] = a.w[0]; out[1] = a.w[1]; out[2] = a.w[2]; out[3] = a.w[3];
  }
  __host__ __device__ static __forceinline__ Elt from_wire(const uint32_t in[4]) {
    Elt a;
    a.w[0] = in[0]; a.w[1] = in[1]; a.w[2] = in[2]; a.w[3] = in[3];
    return a;
  }

#ifdef __CUDACC__
  // ---- byte streams (wire bytes, caller randomness, transcript) ----
  // 16 little-endian bytes are an element (gf2_128.h:171-190); never fails
  __device__ static __forceinline__ Elt from_bytes(const uint8_t* p, bool* ok) {
    Elt e;
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
      e.w[0] = q[0]; e.w[1] = q[1]; e.w[2] = q[2]; e.w[3] = q[3];
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        e.w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
                 ((uint32_t)p[4 * k + 3] << 24);
    }
    return e;
  }
  // Field::sample on a fixed-size slot of the caller's random stream
  __device__ static __forceinline__ Elt sample_bytes(const uint8_t* p, bool* ok) { return from_bytes(p, ok); }
  __device__ static __forceinline__ bool sample_ok(const uint8_t*) { return true; }
  // RandomEngine::elt on the transcript (random.h:37-41, gf2_128.h:182-190)
  __device__ static __forceinline__ Elt ts_elt(Transcript* ts) {
    Elt e;
    ts->words(e.w, 4);
    return e;
  }
  __device__ static __forceinline__ Elt evalpt(int i) { return c_gf.evalpt[i]; }
  // a * poly_evaluation_point(2)  (= a * g)
  __device__ static __forceinline__ Elt mul_x2(const Elt& a) { return gf_mul(a, c_gf.evalpt[2]); }
  __device__ static __forceinline__ Elt newton(int k, int i) { return c_gf.newton[k][i]; }
  __device__ static __forceinline__ Elt lag_id(int k, int i) { return c_gf.lag_id[k][i]; }
  // GF2_128::of_scalar for a 16-bit subfield index (gf2_128.h:151-160)
  __device__ static __forceinline__ Elt of_sub16(uint32_t u) {
    Elt t = gf_zero();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      uint32_t m = 0u - ((u >> k) & 1u);
      t.w[0] ^= c_gf.beta[k].w[0] & m;
      t.w[1] ^= c_gf.beta[k].w[1] & m;
      t.w[2] ^= c_gf.beta[k].w[2] & m;
      t.w[3] ^= c_gf.beta[k].w[3] & m;
    }
    return t;
  }
  // GF2_128::solve (gf2_128.h:495-508): returns true iff e is in GF(2^16) and
  // then *u is its coordinate vector (the 2 wire bytes of to_bytes_subfield).
  __device__ static __forceinline__ bool solve_sub16(const Elt& e, uint32_t* u) {
    uint32_t x0 = e.w[0], x1 = e.w[1], x2 = e.w[2], x3 = e.w[3], acc = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      uint32_t j = c_gf.sub_ldnz[r];
      uint32_t word = (j < 32) ? x0 : (j < 64) ? x1 : (j < 96) ? x2 : x3;
      uint32_t m = 0u - ((word >> (j & 31)) & 1u);
      x0 ^= c_gf.sub_u[r][0] & m;
      x1 ^= c_gf.sub_u[r][1] & m;
      x2 ^= c_gf.sub_u[r][2] & m;
      x3 ^= c_gf.sub_u[r][3] & m;
      acc ^= c_gf.sub_linv[r] & m;
    }
    *u = acc;
    return (x0 | x1 | x2 | x3) == 0;
  }
#endif
};

#ifdef __CUDACC__
static __constant__ FpConsts<8> c_p256;   // lib/algebra/fp_p256.h
static __constant__ FpConsts<8> c_bn254;  // Fp<4>, lib/algebra/fft_test.cc:33-36
static __constant__ FpConsts<4> c_fp128;  // lib/algebra/fp_p128.h
static __constant__ FpConsts<2> c_gold;   // Fp<1> 2^64-2^32+1
#endif

struct P256Traits {
  static constexpr int W = 8;
  static constexpr int kFieldId = 1;  // proto/circuit_io.h P256_ID
  static constexpr bool kP256 = true;
#ifdef __CUDACC__
  __device__ static __forceinline__ const FpConsts<8>& C() { return c_p256; }
#endif
};
struct Bn254Traits {
  static constexpr int W = 8;
  static constexpr int kFieldId = 100;
  static constexpr bool kP256 = false;
#ifdef __CUDACC__
  __device__ static __forceinline__ const FpConsts<8>& C() { return c_bn254; }
#endif
};
struct Fp128Traits {
  static constexpr int W = 4;
  static constexpr int kFieldId = 101;
  static constexpr bool kP256 = false;
#ifdef __CUDACC__
  __device__ static __forceinline__ const FpConsts<4>& C() { return c_fp128; }
#endif
};
struct GoldTraits {
  static constexpr int W = 2;
  static constexpr int kFieldId = 102;
  static constexpr bool kP256 = false;
#ifdef __CUDACC__
  __device__ static __forceinline__ const FpConsts<2>& C() { return c_gold; }
#endif
};

// OOL: add/sub as out-of-line calls (used by the sumcheck kernel only, see below)
template <class T, bool OOL = false>
struct FFp {
  typedef FFp<T, true> Compact;  // the same field with the smallest code footprint
  static constexpr int W = T::W;
  typedef fpw<W> Elt;
  static constexpr int kWords = W;
  static constexpr int kBytes = 4 * W;
  static constexpr int kSubBytes = 4 * W;
  static constexpr bool kChar2 = false;
  static constexpr int kFieldId = T::kFieldId;

  struct Acc {
    Elt v;
  };

#ifdef __CUDACC__
  // one out-of-line Montgomery product per field (same reason as gf_mul_wide_fn)
  static __device__ __noinline__ Elt mul_fn(Elt a, Elt b) {
    if constexpr (T::kP256) {
      return fp_mul_p256_dev(a, b, T::C().m);
    } else {
      return fp_mul_generic_dev<W>(a, b, T::C().m, T::C().mprime);
    }
  }
  __device__ static __forceinline__ Elt zero() {
    Elt r;
#pragma unroll
    for (int i = 0; i < W; ++i) r.w[i] = 0;
    return r;
  }
  __device__ static __forceinline__ Elt cst(const uint32_t* src) {
    Elt r;
#pragma unroll
    for (int i = 0; i < W; ++i) r.w[i] = src[i];
    return r;
  }
  __device__ static __forceinline__ Elt one() { return cst(T::C().one); }
  // The sumcheck kernel instantiates FFp<T, true>: add/sub are ~35 instructions each and
  // sit at hundreds of sites; out of line they keep that kernel inside the instruction
  // cache (ncu: 19 % of the P-256 kernel's stall samples were "no instruction" with
  // everything inlined).  The FFT / RS kernels keep them inline (measured 13 % slower
  // with calls inside the butterflies).
  static __device__ __noinline__ Elt add_fn(Elt a, Elt b) { return fp_add<W>(a, b, T::C().m); }
  static __device__ __noinline__ Elt sub_fn(Elt a, Elt b) { return fp_sub<W>(a, b, T::C().m); }
  __device__ static __forceinline__ Elt add(const Elt& a, const Elt& b) {
    if constexpr (OOL) return add_fn(a, b);
    else return fp_add<W>(a, b, T::C().m);
  }
  __device__ static __forceinline__ Elt sub(const Elt& a, const Elt& b) {
    if constexpr (OOL) return sub_fn(a, b);
    else return fp_sub<W>(a, b, T::C().m);
  }
  __device__ static __forceinline__ Elt neg(const Elt& a) { return fp_sub<W>(zero(), a, T::C().m); }
  __device__ static __forceinline__ Elt mul(const Elt& a, const Elt& b) { return mul_fn(a, b); }
  __device__ static __forceinline__ bool is_zero(const Elt& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) o |= a.w[i];
    return o == 0;
  }
  __device__ static __forceinline__ bool eq(const Elt& a, const Elt& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) o |= a.w[i] ^ b.w[i];
    return o == 0;
  }
  // the reference's Accum is a wide integer (fp_generic.h:424-440); an eager
  // modular accumulator yields the same field element
  __device__ static __forceinline__ void acc_zero(Acc& a) { a.v = zero(); }
  __device__ static __forceinline__ void mac(Acc& a, const Elt& x, const Elt& y) { a.v = add(a.v, mul(x, y)); }
  __device__ static __forceinline__ void mac_sel(Acc& a0, Acc& a1, bool first, const Elt& x, const Elt& y) {
    const Elt prod = mul(x, y);
    Elt s;
#pragma unroll
    for (int i = 0; i < W; ++i) s.w[i] = first ? a0.v.w[i] : a1.v.w[i];
    s = add(s, prod);
#pragma unroll
    for (int i = 0; i < W; ++i) {
      a0.v.w[i] = first ? s.w[i] : a0.v.w[i];
      a1.v.w[i] = first ? a1.v.w[i] : s.w[i];
    }
  }
  __device__ static __forceinline__ void acc_add_elt(Acc& a, const Elt& x) { a.v = add(a.v, x); }
  __device__ static __forceinline__ void acc_add(Acc& a, const Acc& b) { a.v = add(a.v, b.v); }
  __device__ static __forceinline__ Elt reduce(const Acc& a) { return a.v; }

  // from_montgomery / to_montgomery (fp_generic.h:264-281)
  __device__ static __forceinline__ void to_wire(uint32_t out[W], const Elt& a) {
    Elt o = zero();
    o.w[0] = 1;
    Elt r = mul(a, o);
#pragma unroll
    for (int i = 0; i < W; ++i) out[i] = r.w[i];
  }
  __device__ static __forceinline__ Elt from_wire(const uint32_t in[W]) {
    Elt a;
#pragma unroll
    for (int i = 0; i < W; ++i) a.w[i] = in[i];
    return mul(a, cst(T::C().rsq));
  }
  // of_bytes_field (fp_generic.h:351-358): *ok = false if the value is >= p
  __device__ static __forceinline__ Elt from_bytes(const uint8_t* p, bool* ok) {
    uint32_t in[W];
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
      for (int k = 0; k < W; ++k) in[k] = q[k];
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k)
        in[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
                ((uint32_t)p[4 * k + 3] << 24);
    }
    if (fp_geq<W>(in, T::C().m)) *ok = false;
    return from_wire(in);
  }
  // Field::sample on one slot of the caller's stream (fp_generic.h:360-371): the
  // candidate is masked to exact_bits; a value >= p would make the reference
  // draw again -- reported through *ok.
  __device__ static __forceinline__ void mask_bits(uint32_t in[W]) {
    const uint32_t eb = T::C().exact_bits;
#pragma unroll
    for (int k = 0; k < W; ++k) {
      uint32_t lo = 32u * k;
      if (eb <= lo) in[k] = 0;
      else if (eb < lo + 32) in[k] &= (1u << (eb - lo)) - 1u;
    }
  }
  __device__ static __forceinline__ Elt sample_bytes(const uint8_t* p, bool* ok) {
    uint32_t in[W];
#pragma unroll
    for (int k = 0; k < W; ++k)
      in[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
              ((uint32_t)p[4 * k + 3] << 24);
    mask_bits(in);
    if (fp_geq<W>(in, T::C().m)) *ok = false;
    return from_wire(in);
  }
  // would Field::sample accept this slot of the stream, or draw again?
  __device__ static __forceinline__ bool sample_ok(const uint8_t* p) {
    uint32_t in[W];
#pragma unroll
    for (int k = 0; k < W; ++k)
      in[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
              ((uint32_t)p[4 * k + 3] << 24);
    mask_bits(in);
    return !fp_geq<W>(in, T::C().m);
  }
  __device__ static __forceinline__ Elt ts_elt(Transcript* ts) {
    uint32_t in[W];
    for (;;) {
      ts->words(in, W);
      mask_bits(in);
      if (!fp_geq<W>(in, T::C().m)) break;
    }
    return from_wire(in);
  }
  __device__ static __forceinline__ Elt evalpt(int i) { return cst(T::C().evalpt[i]); }
  // a * poly_evaluation_point(2)  (= a + a: the point is the integer 2, fp_generic.h:117-124)
  __device__ static __forceinline__ Elt mul_x2(const Elt& a) { return add(a, a); }
  __device__ static __forceinline__ Elt newton(int k, int i) { return cst(T::C().newton[k][i]); }
  __device__ static __forceinline__ Elt lag_id(int k, int i) { return cst(T::C().lag_id[k][i]); }
  __device__ static __forceinline__ Elt of_sub16(uint32_t) { return zero(); }        // never used
  __device__ static __forceinline__ bool solve_sub16(const Elt&, uint32_t* u) {       // never used
    *u = 0;
    return true;
  }
#endif
};

// Fp256Base = FpGeneric<4, true, Fp256Reduce> (lib/algebra/fp_p256.h:64-65, lib/ec/p256.h:42);
// sample_subfield == sample, in_subfield == true, kSubFieldBytes == kBytes (fp_generic.h:278,373-400)
typedef FFp<P256Traits> FFp256;
typedef FFp<Bn254Traits> FFpBn254;
typedef FFp<Fp128Traits> FFp128;
typedef FFp<GoldTraits> FFpGold;

// arrays/affine.h:25-52
template <class F>
__host__ __device__ __forceinline__ typename F::Elt affine(const typename F::Elt& r,
                                                         const typename F::Elt& f0,
                                                         const typename F::Elt& f1) {
  return F::add(f0, F::mul(F::sub(f1, f0), r));
}
template <class F>
__host__ __device__ __forceinline__ typename F::Elt affine_z_nz(const typename F::Elt& r,
                                                              const typename F::Elt& f1) {
  return F::mul(f1, r);
}
template <class F>
__host__ __device__ __forceinline__ typename F::Elt affine_nz_z(const typename F::Elt& r,
                                                              const typename F::Elt& f0) {
  return F::sub(f0, F::mul(f0, r));
}

}  // namespace lf
