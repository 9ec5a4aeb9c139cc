// Reed-Solomon row extension over Fp256 (P-256) through a multi-modular (CRT) number-theoretic transform.
//
// P-256 has no 2-power roots of unity, so the reference runs the convolution of ReedSolomon::interpolate
// (lib/algebra/reed_solomon.h:93-110) over Fp2 (FFTExtConvolution, lib/algebra/convolution.h:128-191) -- and
// offers an exact alternative, CrtConvolution (lib/algebra/crt_convolution.h:30-83, lib/algebra/crt.h:133-289):
// the convolution of the integer representatives is computed modulo several word-size NTT primes and put back
// together with the Chinese remainder theorem.  A convolution is an exact integer identity, so both give the
// same field elements, bit for bit.  On the GPU the word-size route wins: k_rs_fp_fft_rows needs 128 registers
// and 128 KB of shared memory per row (one CTA per SM, 0.44 instructions per cycle and scheduler, ncu
// profiles/r2_rs_p256_final.txt); here a butterfly is a 32-bit Montgomery product and a row's working set is
// 16 KB, so four CTAs share an SM.
//
//   c[k] = sum_{i<n} x~[i] * t~[k-i],  k in [n, m)       (x~ = binom*y, t~ = the 1/i table, both as the
//                                                         integers < P of their Montgomery form)
//   0 <= c[k] < n * P^2 < 2^521;  17 primes p_j = k_j 2^13 + 1 just below 2^31, M = prod p_j ~ 2^527.
//   per prime: residues -> forward NTT (DIF) -> times the table's spectrum -> inverse NTT (DIT)
//   CRT:  a_j = c mod p_j times (M/p_j)^-1 (folded into the spectrum),  q = floor(sum a_j / p_j),
//         c R^-1 mod P = sum a_j C_j + E_q  mod P,   C_j = (M/p_j) R^-1 mod P,  E_q = -q M R^-1 mod P
//   (R = 2^256: the result is the Montgomery form of the field convolution), then times lead[k-(n-1)].
#pragma once
#include <stdint.h>

#include "field.cuh"
#include "fp.cuh"

namespace lf {

constexpr int kCrtPrimes = 17;
constexpr int kCrtQBits = 59;  // q = floor((sum a_j * floor(2^59 / p_j) + 2^39) / 2^59)

struct CrtPrime {
  uint32_t p, pinv;   // p^-1 mod 2^32
  uint32_t k[8];      // 2^(32 l) * 2^32 mod p: Montgomery factor of limb l
  uint32_t f;         // floor(2^59 / p)
  uint32_t pad[5];
};
struct CrtConsts {
  CrtPrime pr[kCrtPrimes];
  uint32_t C[kCrtPrimes][8];       // (M / p_j) R^-1 mod P
  uint32_t E[kCrtPrimes + 1][8];   // -q M R^-1 mod P
};

// a * b * 2^-32 mod p  (a < 2^32, b < p < 2^31), result in [0, p)
__device__ __forceinline__ uint32_t crt_mm(uint32_t a, uint32_t b, uint32_t p, uint32_t pinv) {
  const uint64_t t = (uint64_t)a * b;
  const uint32_t m = (uint32_t)t * pinv;
  const uint64_t u = (uint64_t)m * p;
  const uint32_t r = (uint32_t)(t >> 32) - (uint32_t)(u >> 32);  // in (-p, p), wrapped
  return min(r, r + p);
}
__device__ __forceinline__ uint32_t crt_add(uint32_t a, uint32_t b, uint32_t p) {
  const uint32_t s = a + b;
  return min(s, s - p);
}
__device__ __forceinline__ uint32_t crt_sub(uint32_t a, uint32_t b, uint32_t p) {
  const uint32_t d = a - b;
  return min(d, d + p);
}

// ---- transforms in shared memory, up to four stages per pass held in registers --------------------------
// A pass of R stages takes groups of 2^R elements into registers, runs the R butterfly stages there and writes
// them back: for N = 4096 a transform is three radix-16 passes (three barriers) instead of twelve stages.
// The array is skewed by one word per 32 (crt_sk) so that both the stride-N/16 and the contiguous groups are
// free of bank conflicts.
__device__ __forceinline__ uint32_t crt_sk(uint32_t i) { return i + (i >> 5); }

// forward (decimation in frequency) stages s0 .. s0+R-1; stage s has half = N >> (s+1) and twiddle
// w^(lo << s) for the butterfly whose lower element sits lo into its half block
template <int R, class Load, class Store>
__device__ __forceinline__ void crt_fwd_pass(uint32_t logN, uint32_t s0, uint32_t p, uint32_t pinv,
                                             const uint32_t* __restrict__ wf, Load load, Store store) {
  constexpr uint32_t G = 1u << R;
  const uint32_t N = 1u << logN;
  const uint32_t lhr = logN - s0 - R;  // log2 of the smallest half of the pass = stride between group elements
  const uint32_t hr = 1u << lhr;
  for (uint32_t g = threadIdx.x; g < (N >> R); g += blockDim.x) {
    const uint32_t lo = g & (hr - 1), base = ((g >> lhr) << (lhr + R)) + lo;
    uint32_t v[G];
#pragma unroll
    for (uint32_t k = 0; k < G; ++k) v[k] = load(base + (k << lhr));
#pragma unroll
    for (int q = 0; q < R; ++q) {
      constexpr uint32_t dummy = 0;
      (void)dummy;
      const uint32_t hk = 1u << (R - 1 - q);  // half, in units of group elements
      const uint32_t s = s0 + q;
#pragma unroll
      for (uint32_t k = 0; k < G; ++k) {
        if (k & hk) continue;
        const uint32_t los = lo + ((k & (hk - 1)) << lhr);
        const uint32_t a = v[k], b = v[k + hk];
        v[k] = crt_add(a, b, p);
        v[k + hk] = crt_mm(crt_sub(a, b, p), wf[los << s], p, pinv);
      }
    }
#pragma unroll
    for (uint32_t k = 0; k < G; ++k) store(base + (k << lhr), v[k]);
  }
}

// inverse (decimation in time) stages lh0 .. lh0+R-1; stage lh has half = 1 << lh and twiddle
// w^-(lo << (logN-1-lh))
template <int R, class Load, class Store>
__device__ __forceinline__ void crt_inv_pass(uint32_t logN, uint32_t lh0, uint32_t p, uint32_t pinv,
                                             const uint32_t* __restrict__ wi, Load load, Store store) {
  constexpr uint32_t G = 1u << R;
  const uint32_t N = 1u << logN;
  const uint32_t h0 = 1u << lh0;
  for (uint32_t g = threadIdx.x; g < (N >> R); g += blockDim.x) {
    const uint32_t lo = g & (h0 - 1), base = ((g >> lh0) << (lh0 + R)) + lo;
    uint32_t v[G];
#pragma unroll
    for (uint32_t k = 0; k < G; ++k) v[k] = load(base + (k << lh0));
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const uint32_t hk = 1u << q;
      const uint32_t sh = logN - 1 - (lh0 + q);
#pragma unroll
      for (uint32_t k = 0; k < G; ++k) {
        if (k & hk) continue;
        const uint32_t los = lo + ((k & (hk - 1)) << lh0);
        const uint32_t a = v[k], b = crt_mm(v[k + hk], wi[los << sh], p, pinv);
        v[k] = crt_add(a, b, p);
        v[k + hk] = crt_sub(a, b, p);
      }
    }
#pragma unroll
    for (uint32_t k = 0; k < G; ++k) store(base + (k << lh0), v[k]);
  }
}

template <class Load, class Store>
__device__ __forceinline__ void crt_fwd_pass_r(int r, uint32_t logN, uint32_t s0, uint32_t p, uint32_t pinv,
                                               const uint32_t* __restrict__ wf, Load load, Store store) {
  switch (r) {
    case 4: crt_fwd_pass<4>(logN, s0, p, pinv, wf, load, store); break;
    case 3: crt_fwd_pass<3>(logN, s0, p, pinv, wf, load, store); break;
    case 2: crt_fwd_pass<2>(logN, s0, p, pinv, wf, load, store); break;
    default: crt_fwd_pass<1>(logN, s0, p, pinv, wf, load, store); break;
  }
}
template <class Load, class Store>
__device__ __forceinline__ void crt_inv_pass_r(int r, uint32_t logN, uint32_t lh0, uint32_t p, uint32_t pinv,
                                               const uint32_t* __restrict__ wi, Load load, Store store) {
  switch (r) {
    case 4: crt_inv_pass<4>(logN, lh0, p, pinv, wi, load, store); break;
    case 3: crt_inv_pass<3>(logN, lh0, p, pinv, wi, load, store); break;
    case 2: crt_inv_pass<2>(logN, lh0, p, pinv, wi, load, store); break;
    default: crt_inv_pass<1>(logN, lh0, p, pinv, wi, load, store); break;
  }
}

// One CTA extends one row.  Tables: tw_f / tw_i [prime][N/2] forward / inverse twiddles w^k 2^32 mod p;
// spec [prime][N] = NTT(t~) (N^-1) ((M/p)^-1) 2^32 mod p in the bit-reversed order the DIF transform leaves;
// scratch [row][prime][m - n] residues.  Dynamic shared memory: 8 n + N + N/32 words.
template <class F>
__global__ void __launch_bounds__(256, 4)
k_rs_crt_rows(typename F::Elt* __restrict__ data, size_t row_stride, size_t batch_stride, uint32_t n, uint32_t m,
              uint32_t logN, const CrtConsts* __restrict__ cc, const uint32_t* __restrict__ tw_f,
              const uint32_t* __restrict__ tw_i, const uint32_t* __restrict__ spec,
              const typename F::Elt* __restrict__ lead, const typename F::Elt* __restrict__ binom,
              uint32_t* __restrict__ scratch) {
  typedef typename F::Elt Elt;
  extern __shared__ __align__(16) uint32_t crt_smem[];
  const uint32_t N = 1u << logN, tid = threadIdx.x, nth = blockDim.x;
  uint32_t* X = crt_smem;          // [8][n] limbs of x~, limb-major
  uint32_t* A = crt_smem + 8 * n;  // [N + N/32], skewed
  __shared__ CrtConsts s_cc;
  for (uint32_t i = tid; i < sizeof(CrtConsts) / 4; i += nth)
    reinterpret_cast<uint32_t*>(&s_cc)[i] = reinterpret_cast<const uint32_t*>(cc)[i];
  const size_t row = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  Elt* y = data + (size_t)blockIdx.y * batch_stride + (size_t)blockIdx.x * row_stride;
  uint32_t* S = scratch + row * (size_t)kCrtPrimes * (m - n);
  // x~ = binom * y (Montgomery product of two Montgomery forms = Montgomery form of the product)
  for (uint32_t i = tid; i < n; i += nth) {
    const Elt x = F::mul(binom[i], y[i]);
#pragma unroll
    for (int l = 0; l < 8; ++l) X[l * n + i] = x.w[l];
  }
  __syncthreads();
  const int npass = (int)((logN + 3) / 4), rlast = (int)(logN - 4 * (npass - 1));
  for (int j = 0; j < kCrtPrimes; ++j) {
    const uint32_t p = s_cc.pr[j].p, pinv = s_cc.pr[j].pinv;
    const uint32_t* wf = tw_f + (size_t)j * (N / 2);
    const uint32_t* wi = tw_i + (size_t)j * (N / 2);
    const uint32_t* sp = spec + (size_t)j * N;
    uint32_t* Sj = S + (size_t)j * (m - n);
    const uint32_t* kk = s_cc.pr[j].k;
    // the first pass reads the residues of x~ modulo p (zero beyond n), the last one multiplies by the spectrum
    auto ld_res = [&](uint32_t i) -> uint32_t {
      uint32_t r = 0;
      if (i < n) {
#pragma unroll
        for (int l = 0; l < 8; ++l) r = crt_add(r, crt_mm(X[l * n + i], kk[l], p, pinv), p);
      }
      return r;
    };
    auto ld_a = [&](uint32_t i) -> uint32_t { return A[crt_sk(i)]; };
    auto st_a = [&](uint32_t i, uint32_t v) { A[crt_sk(i)] = v; };
    auto st_spec = [&](uint32_t i, uint32_t v) { A[crt_sk(i)] = crt_mm(v, sp[i], p, pinv); };
    auto st_out = [&](uint32_t i, uint32_t v) {
      if (i >= n && i < m) Sj[i - n] = v;
    };
    // forward: natural order in, bit-reversed order out (the order spec is stored in)
    for (int ps = 0; ps < npass; ++ps) {
      const int r = ps + 1 == npass ? rlast : 4;
      const uint32_t s0 = 4 * ps;
      if (npass == 1) crt_fwd_pass_r(r, logN, s0, p, pinv, wf, ld_res, st_spec);
      else if (ps == 0) crt_fwd_pass_r(r, logN, s0, p, pinv, wf, ld_res, st_a);
      else if (ps + 1 == npass) crt_fwd_pass_r(r, logN, s0, p, pinv, wf, ld_a, st_spec);
      else crt_fwd_pass_r(r, logN, s0, p, pinv, wf, ld_a, st_a);
      __syncthreads();
    }
    // inverse: bit-reversed order in, natural order out; the last pass writes the residues of the wanted
    // outputs to the scratch array
    for (int ps = 0; ps < npass; ++ps) {
      const int r = ps + 1 == npass ? rlast : 4;
      const uint32_t lh0 = 4 * ps;
      if (ps + 1 == npass) crt_inv_pass_r(r, logN, lh0, p, pinv, wi, ld_a, st_out);
      else crt_inv_pass_r(r, logN, lh0, p, pinv, wi, ld_a, st_a);
      __syncthreads();
    }
  }
  // Chinese remaindering and the leading constants
  for (uint32_t k = n + tid; k < m; k += nth) {
    uint64_t lo[8], hi[8], z = (uint64_t)1 << (kCrtQBits - 20);
#pragma unroll
    for (int l = 0; l < 8; ++l) lo[l] = hi[l] = 0;
    for (int j = 0; j < kCrtPrimes; ++j) {
      const uint32_t a = S[(size_t)j * (m - n) + (k - n)];
      z += (uint64_t)a * s_cc.pr[j].f;
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const uint64_t pr = (uint64_t)a * s_cc.C[j][l];
        lo[l] += (uint32_t)pr;
        hi[l] += pr >> 32;
      }
    }
    const uint32_t q = (uint32_t)(z >> kCrtQBits);
    // V = sum a_j C_j + E_q as signed 64-bit limbs, then 2^256 = 2^224 - 2^192 - 2^96 + 1 (mod P) until the
    // top is gone
    int64_t v[9];
#pragma unroll
    for (int l = 0; l < 8; ++l) v[l] = (int64_t)(lo[l] + (l ? hi[l - 1] : 0) + s_cc.E[q][l]);
    v[8] = (int64_t)hi[7];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      int64_t c = 0;
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const int64_t t = v[l] + c;
        v[l] = t & 0xffffffffll;
        c = t >> 32;  // arithmetic shift: floor division, limbs end up in [0, 2^32)
      }
      const int64_t top = v[8] + c;  // multiples of 2^256
      v[8] = 0;
      v[0] += top;
      v[3] -= top;
      v[6] -= top;
      v[7] += top;
    }
    // value in [0, 2^256) with limbs normalised except for the last fold's carries: once more, exactly
    Elt e;
    {
      int64_t c = 0;
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const int64_t t = v[l] + c;
        e.w[l] = (uint32_t)t;
        c = t >> 32;
      }
    }
    if (fp_geq<8>(e.w, c_p256.m)) fp_subn<8>(e.w, e.w, c_p256.m);
    y[k] = F::mul(lead[k - (n - 1)], e);
  }
}

}  // namespace lf
