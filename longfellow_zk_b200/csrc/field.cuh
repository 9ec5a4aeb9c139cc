// Field policy classes used to template the kernels (the counterpart of the
// reference's "Field concept", lib/algebra/fp_generic.h:36-568 and
// lib/gf2k/gf2_128.h:35-509).  Each policy provides Elt, add/sub/mul, a lazy
// accumulator (the reference's Accum/mac/reduce), wire conversion and the small
// per-field constant block that the sumcheck needs.
#pragma once
#include <stdint.h>

#include "gf128.cuh"

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
  __device__ static __forceinline__ Elt evalpt(int i) { return c_gf.evalpt[i]; }
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
