// Commitment-phase kernels: Reed-Solomon row extension over GF(2^128) (LCH14
// additive FFT) and the SHA-256 Merkle column commitment.
//
//   reference                                   here
//   LCH14ReedSolomon::interpolate               k_rs_gf_rows
//     (lib/gf2k/lch14_reed_solomon.h:49-103,
//      lib/gf2k/lch14.h:106-237)
//   MerkleCommitment::commit leaf loop          k_merkle_leaves
//     (lib/merkle/merkle_commitment.h:50-64,
//      lib/ligero/ligero_param.h:432-439)
//   MerkleTree::build_tree                      k_merkle_tree
//     (lib/merkle/merkle_tree.h:109-114)
#pragma once
#include <stdint.h>

#include "field.cuh"
#include "hash.cuh"

namespace lf {

// ---------------------------------------------------------------------------
// LCH14 twiddles.  The twiddle of the butterfly at global evaluation index p
// (bit i of p clear) in stage i is  twiddle(i, p with its low i+1 bits cleared)
// (lch14.h:81-100), which is GF(2)-linear in the index bits, so one table
//   T_i[u] = sum_{k : bit k of u} w_hat[i][i+1+k],   u < 2^(15-i)
// serves every FFT size, every coset and the truncated ("bidirectional")
// transform alike.  T_i starts at offset 65536 - 2^(16-i) of d_tw.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tw_offset(uint32_t i) { return 65536u - (1u << (16 - i)); }

// One step of the precomputed butterfly schedule of an (n -> m) extension.
// Butterflies t in [t0,t1) of the step act at local position
//   q = base + ((t >> stage) << (stage+1)) + (t & ((1<<stage)-1))   and q + 2^stage.
struct RsStep {
  uint32_t kind;   // 0 fwd, 1 bwd, 2 diag   (lch14.h:219-237)
  uint32_t stage;
  uint32_t base;
  uint32_t t0, t1;
  uint32_t wsync;  // in-CTA kernel: the barrier after this step may be __syncwarp() (set by rs_mark_wsync)
};

struct RsPlan {
  uint32_t n, m, l, fftn;
  uint32_t nsteps;       // schedule of BidirectionalFFT(l, n) on the first coset
  const RsStep* steps;   // device
};

template <class F>
__device__ __forceinline__ void rs_butterfly(typename F::Elt* B, uint32_t q, uint32_t s, uint32_t kind,
                                             const typename F::Elt& tw) {
  typename F::Elt b0 = B[q], b1 = B[q + s];
  if (kind == 0) {  // fwd
    b0 = F::add(b0, F::mul(tw, b1));
    b1 = F::add(b1, b0);
  } else if (kind == 1) {  // bwd
    b1 = F::sub(b1, b0);
    b0 = F::sub(b0, F::mul(tw, b1));
  } else {  // diag: forward at [q+s], backward at [q]
    typename F::Elt t = b1;
    b1 = F::add(b1, b0);
    b0 = F::sub(b0, F::mul(tw, t));
  }
  B[q] = b0;
  B[q + s] = b1;
}

#ifndef LF_RS_GF_MIN_CTAS
#define LF_RS_GF_MIN_CTAS 4  // 64 registers (66 uncapped): 32 warps per SM, i.e. 8 CTAs of the default 128 threads
#endif
// One CTA extends one row: y[0..n) given, y[n..m) produced.  rows are
// `row_stride` elements apart; blockIdx.y selects the batch instance
// (batch_stride elements apart).  Dynamic shared memory: 2 * fftn elements.
template <class F>
__global__ void __launch_bounds__(256, LF_RS_GF_MIN_CTAS)
k_rs_gf_rows(typename F::Elt* __restrict__ data, size_t row_stride, size_t batch_stride, RsPlan plan,
             const typename F::Elt* __restrict__ d_tw) {
  typedef typename F::Elt Elt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Elt* C = reinterpret_cast<Elt*>(smem_raw);
  Elt* D = C + plan.fftn;
  Elt* y = data + (size_t)blockIdx.y * batch_stride + (size_t)blockIdx.x * row_stride;
  const uint32_t n = plan.n, m = plan.m, l = plan.l, fftn = plan.fftn;

  for (uint32_t i = threadIdx.x; i < fftn; i += blockDim.x) C[i] = (i < n) ? y[i] : F::zero();
  __syncthreads();

  // truncated transform on the first coset (lch14.h:185-217), flattened into
  // a host-generated schedule: after it C[0..n) are novel-basis coefficients
  // and C[n..fftn) the missing evaluations.
  for (uint32_t s = 0; s < plan.nsteps; ++s) {
    const RsStep st = plan.steps[s];
    const uint32_t half = 1u << st.stage;
    for (uint32_t t = st.t0 + threadIdx.x; t < st.t1; t += blockDim.x) {
      uint32_t q = st.base + ((t >> st.stage) << (st.stage + 1)) + (t & (half - 1));
      Elt tw = d_tw[tw_offset(st.stage) + (q >> (st.stage + 1))];
      rs_butterfly<F>(C, q, half, st.kind, tw);
    }
    if (st.wsync) __syncwarp();
    else __syncthreads();
  }
  for (uint32_t i = n + threadIdx.x; i < fftn && i < m; i += blockDim.x) y[i] = C[i];
  __syncthreads();
  for (uint32_t i = n + threadIdx.x; i < fftn; i += blockDim.x) C[i] = F::zero();
  __syncthreads();

  // remaining cosets: forward FFT of the coefficients at offset b (lch14.h:106-123)
  // The 32 butterflies a warp owns in one pass of a stage st <= 5 lie in the warp's
  // own 64-element block (t is warp-aligned), for this stage and every later one:
  // those stages are separated by __syncwarp() instead of a CTA barrier.  The top
  // stage reads the coefficients C and writes D, so no copy pass is needed.
  const bool warp_private = fftn >= 64;
  if (l == 0) {  // n == 1: the constant polynomial
    if (threadIdx.x == 0) D[0] = C[0];
    __syncthreads();
  }
  for (uint32_t b = fftn; b < m; b += fftn) {
    for (uint32_t st = l; st-- > 0;) {
      const uint32_t half = 1u << st;
      const Elt* src = (st + 1 == l) ? C : D;
      for (uint32_t t = threadIdx.x; t < fftn / 2; t += blockDim.x) {
        uint32_t q = ((t >> st) << (st + 1)) + (t & (half - 1));
        Elt tw = d_tw[tw_offset(st) + ((b + q) >> (st + 1))];
        Elt b0 = src[q], b1 = src[q + half];
        b0 = F::add(b0, F::mul(tw, b1));
        b1 = F::add(b1, b0);
        D[q] = b0;
        D[q + half] = b1;
      }
      if (warp_private && st >= 1 && st <= 5) __syncwarp();
      else __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < fftn && b + i < m; i += blockDim.x) y[b + i] = D[i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Large rows (2 * fftn elements do not fit in shared memory, e.g. the
// 16384 -> 65536 shape of lch14_reed_solomon_test.cc:72-107) and the
// stand-alone LCH14::FFT / IFFT (lch14.h:106-146): the same schedule, one
// launch per step over a global-memory work array.  blockIdx.y = row.
// ---------------------------------------------------------------------------
template <class F>
__global__ void k_rs_gf_gstep(typename F::Elt* __restrict__ work, size_t work_stride, RsStep st, uint32_t coset,
                              const typename F::Elt* __restrict__ d_tw) {
  const uint32_t t = st.t0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= st.t1) return;
  const uint32_t half = 1u << st.stage;
  const uint32_t q = st.base + ((t >> st.stage) << (st.stage + 1)) + (t & (half - 1));
  typename F::Elt tw = d_tw[tw_offset(st.stage) + ((coset + q) >> (st.stage + 1))];
  rs_butterfly<F>(work + (size_t)blockIdx.y * work_stride, q, half, st.kind, tw);
}
// dst[row][i] = i in [lo, hi) ? src[row][i] : 0   for i < count (zero-extending copy)
template <class F>
__global__ void k_rs_gf_copy(typename F::Elt* __restrict__ dst, size_t dst_stride,
                             const typename F::Elt* __restrict__ src, size_t src_stride, uint32_t count, uint32_t lo,
                             uint32_t hi) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  dst[(size_t)blockIdx.y * dst_stride + i] =
      (i >= lo && i < hi) ? src[(size_t)blockIdx.y * src_stride + i] : F::zero();
}

// ---------------------------------------------------------------------------
// Prime-field RS extension (lib/algebra/reed_solomon.h:27-41,93-110):
//   p(k) = lead[k-d] * sum_{i<n} x_i / (k - i),  x_i = (-1)^i C(d,i) p(i),  d = n-1
// The reference evaluates the sum as an FFT convolution with the table 1/i
// (lib/algebra/convolution.h:80-91,156-175); arithmetic is exact, so the direct
// Toeplitz sum gives the same values.  One CTA per row keeps x[] in shared
// memory; a warp reads 32 consecutive entries of the 1/i table per step.
// ---------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_rs_fp_rows(typename F::Elt* __restrict__ data, size_t row_stride, size_t batch_stride, uint32_t n, uint32_t m,
             const typename F::Elt* __restrict__ inv, const typename F::Elt* __restrict__ lead,
             const typename F::Elt* __restrict__ binom) {
  typedef typename F::Elt Elt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Elt* x = reinterpret_cast<Elt*>(smem_raw);
  Elt* y = data + (size_t)blockIdx.y * batch_stride + (size_t)blockIdx.x * row_stride;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) x[i] = F::mul(binom[i], y[i]);
  __syncthreads();
  for (uint32_t k = n + threadIdx.x; k < m; k += blockDim.x) {
    typename F::Acc acc;
    F::acc_zero(acc);
    const Elt* iv = inv + k;
    for (uint32_t i = 0; i < n; ++i) F::mac(acc, x[i], iv[-(int)i]);
    y[k] = F::mul(lead[k - (n - 1)], F::reduce(acc));
  }
}

// ---------------------------------------------------------------------------
// The same extension through an exact real-input FFT convolution over
// Fp2 = Fp[i]/(i^2+1), the counterpart of FFTExtConvolution + RFFT
// (lib/algebra/convolution.h:128-191, lib/algebra/rfft.h:46-409): P-256 has no
// 2-power roots of unity in Fp, but the norm-1 subgroup of Fp2 does (the root
// of order 2^31 of lib/circuits/mdoc/mdoc_zk.cc:83-88).  With W a primitive
// N-th root on the unit circle (conj(W) = 1/W), M = N/2 and V = W^2:
//   z[j] = x[2j] + i x[2j+1];  Z = DFT_M(z) (decimation in frequency, output
//   bit-reversed);  E2 = Z[k] + conj Z[M-k], O2 = -i (Z[k] - conj Z[M-k]),
//   X2[k] = E2 + W^-k O2 (= 2 DFT_N(x)[k]);  P = X2 * Yh with Yh = DFT_N(1/i
//   table)/(2N) precomputed;  Pe = P[k] + conj P[M-k], Po = (P[k] - conj P[M-k]) W^k,
//   Q[k] = Pe + i Po;  q = inverse DFT_M(Q) (decimation in time, bit-reversed
//   input)  ==>  conv[2j] = Re q[j], conv[2j+1] = Im q[j].
// Every step is exact field arithmetic, so the result equals the direct sum
// (k_rs_fp_rows) and the reference's output bit for bit.  One CTA per row; the
// M complex points (64 B each) live in shared memory (128 KB for N = 4096).
// ---------------------------------------------------------------------------
template <class F>
struct Cx {
  typename F::Elt re, im;
};
template <class F>
__device__ __forceinline__ Cx<F> cx_add(const Cx<F>& a, const Cx<F>& b) {
  return Cx<F>{F::add(a.re, b.re), F::add(a.im, b.im)};
}
template <class F>
__device__ __forceinline__ Cx<F> cx_sub(const Cx<F>& a, const Cx<F>& b) {
  return Cx<F>{F::sub(a.re, b.re), F::sub(a.im, b.im)};
}
template <class F>
__device__ __forceinline__ Cx<F> cx_neg(const Cx<F>& a) {
  return Cx<F>{F::neg(a.re), F::neg(a.im)};
}
template <class F>
__device__ __forceinline__ Cx<F> cx_conj(const Cx<F>& a) {
  return Cx<F>{a.re, F::neg(a.im)};
}
template <class F>
__device__ __forceinline__ Cx<F> cx_muli(const Cx<F>& a) {  // a * i
  return Cx<F>{F::neg(a.im), a.re};
}
template <class F>
__device__ __forceinline__ Cx<F> cx_mulmi(const Cx<F>& a) {  // a * (-i)
  return Cx<F>{a.im, F::neg(a.re)};
}
// Fp2::mul with three base multiplications (lib/algebra/fp2.h:87-101)
template <class F>
__device__ __forceinline__ Cx<F> cx_mul(const Cx<F>& a, const Cx<F>& b) {
  typename F::Elt t1 = F::mul(a.re, b.re), t2 = F::mul(a.im, b.im);
  typename F::Elt t3 = F::mul(F::add(a.re, a.im), F::add(b.re, b.im));
  return Cx<F>{F::sub(t1, t2), F::sub(F::sub(t3, t1), t2)};
}

template <class F>
__global__ void __launch_bounds__(512, 1)
k_rs_fp_fft_rows(typename F::Elt* __restrict__ data, size_t row_stride, size_t batch_stride, uint32_t n,
                 uint32_t m, uint32_t logM, const Cx<F>* __restrict__ Wk /* [M] W^k */,
                 const Cx<F>* __restrict__ Yh /* [M+1] */, const typename F::Elt* __restrict__ lead,
                 const typename F::Elt* __restrict__ binom) {
  typedef typename F::Elt Elt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<F>* a = reinterpret_cast<Cx<F>*>(smem_raw);
  const uint32_t M = 1u << logM, tid = threadIdx.x, nth = blockDim.x;
  Elt* y = data + (size_t)blockIdx.y * batch_stride + (size_t)blockIdx.x * row_stride;
  auto brev = [logM](uint32_t k) { return logM ? (__brev(k) >> (32 - logM)) : 0u; };

  for (uint32_t j = tid; j < M; j += nth) {
    Cx<F> z;
    z.re = (2 * j < n) ? F::mul(binom[2 * j], y[2 * j]) : F::zero();
    z.im = (2 * j + 1 < n) ? F::mul(binom[2 * j + 1], y[2 * j + 1]) : F::zero();
    a[j] = z;
  }
  __syncthreads();
  // forward: decimation in frequency, twiddle V^-e = conj(W^(2e)).
  // The input is zero beyond nz = ceil(n/2) complex points, so while a block's
  // lower half still holds all of them (half >= nz) the upper input v is zero:
  // the butterfly degenerates to (u, u * w) on the nz live points of each block.
  const uint32_t nz = (n + 1) / 2;
  uint32_t lg = logM;
  for (; lg >= 1 && (1u << (lg - 1)) >= nz; --lg) {
    const uint32_t half = 1u << (lg - 1), step = M >> lg, nblk = M >> lg;
    for (uint32_t t = tid; t < nblk * nz; t += nth) {
      const uint32_t j = t % nz, i = (t / nz) << lg;
      const Cx<F> u = a[i + j];
      a[i + j + half] = (j == 0) ? u : cx_mul<F>(u, cx_conj<F>(Wk[2 * j * step]));
    }
    __syncthreads();
  }
  // Remaining stages two at a time (radix 4: three twiddle multiplications per four
  // points instead of four, half the barriers).  W^(e) for e in [0, 2M): W^M = -1.
  // J = W^(M/2) is the element of order 4, i.e. +i or -i; multiplying by it is free.
  const bool j_is_i = (M >= 2) && F::eq(Wk[M / 2].im, F::one());
  auto Wpow = [&](uint32_t e) { return e < M ? Wk[e] : cx_neg<F>(Wk[e - M]); };
  auto mulJ = [&](const Cx<F>& z) { return j_is_i ? cx_muli<F>(z) : cx_mulmi<F>(z); };     // z * J
  auto mulJinv = [&](const Cx<F>& z) { return j_is_i ? cx_mulmi<F>(z) : cx_muli<F>(z); };  // z / J
  for (; lg >= 2; lg -= 2) {
    const uint32_t q = 1u << (lg - 2), step = M >> lg;  // block 4q, twiddle base e1 = 2 j step
    for (uint32_t t = tid; t < M / 4; t += nth) {
      const uint32_t j = t & (q - 1), i = (t >> (lg - 2)) << lg, e1 = 2 * j * step;
      const Cx<F> a0 = a[i + j], a1 = a[i + j + q], a2 = a[i + j + 2 * q], a3 = a[i + j + 3 * q];
      const Cx<F> t0 = cx_add<F>(a0, a2), t1 = cx_add<F>(a1, a3), t2 = cx_sub<F>(a0, a2),
                  t3 = mulJinv(cx_sub<F>(a1, a3));
      const Cx<F> c1 = cx_sub<F>(t0, t1), c2 = cx_add<F>(t2, t3), c3 = cx_sub<F>(t2, t3);
      a[i + j] = cx_add<F>(t0, t1);
      if (j == 0) {
        a[i + j + q] = c1;
        a[i + j + 2 * q] = c2;
        a[i + j + 3 * q] = c3;
      } else {
        a[i + j + q] = cx_mul<F>(c1, cx_conj<F>(Wpow(2 * e1)));
        a[i + j + 2 * q] = cx_mul<F>(c2, cx_conj<F>(Wpow(e1)));
        a[i + j + 3 * q] = cx_mul<F>(c3, cx_conj<F>(Wpow(3 * e1)));
      }
    }
    __syncthreads();
  }
  if (lg == 1) {  // one radix-2 stage left (distance 1, twiddle 1)
    for (uint32_t t = tid; t < M / 2; t += nth) {
      Cx<F> u = a[2 * t], v = a[2 * t + 1];
      a[2 * t] = cx_add<F>(u, v);
      a[2 * t + 1] = cx_sub<F>(u, v);
    }
    __syncthreads();
  }
  // spectrum of the real sequence, multiply by Yh, re-pack for the inverse
  for (uint32_t k = tid; k <= M / 2; k += nth) {
    if (k == 0) {
      Cx<F> z0 = a[0];
      Elt e2 = F::add(z0.re, z0.re), o2 = F::add(z0.im, z0.im);
      Cx<F> x0{F::add(e2, o2), F::zero()}, xm{F::sub(e2, o2), F::zero()};
      Cx<F> p0 = cx_mul<F>(x0, Yh[0]), pm = cx_mul<F>(xm, Yh[M]);
      a[0] = cx_add<F>(cx_add<F>(p0, pm), cx_muli<F>(cx_sub<F>(p0, pm)));
    } else {
      const uint32_t k2 = M - k, bk = brev(k), bk2 = brev(k2);
      Cx<F> zk = a[bk], zk2c = cx_conj<F>(a[bk2]);
      Cx<F> e2 = cx_add<F>(zk, zk2c), o2 = cx_mulmi<F>(cx_sub<F>(zk, zk2c));
      const Cx<F> w = Wk[k];
      const Cx<F> wo = cx_mul<F>(cx_conj<F>(w), o2);  // and w * conj(o2) = conj(wo)
      Cx<F> xk = cx_add<F>(e2, wo);
      Cx<F> xk2 = cx_sub<F>(cx_conj<F>(e2), cx_conj<F>(wo));
      Cx<F> pk = cx_mul<F>(xk, Yh[k]), pk2c = cx_conj<F>(cx_mul<F>(xk2, Yh[k2]));
      Cx<F> pe = cx_add<F>(pk, pk2c), po = cx_mul<F>(cx_sub<F>(pk, pk2c), w);
      Cx<F> qk = cx_add<F>(pe, cx_muli<F>(po));
      a[bk] = qk;
      if (k2 != k) a[bk2] = cx_add<F>(cx_conj<F>(pe), cx_muli<F>(cx_conj<F>(po)));
    }
  }
  __syncthreads();
  // inverse: decimation in time on the bit-reversed array, twiddle V^e = W^(2e)
  {
    uint32_t li = 1;
    if (logM & 1) {  // odd number of stages: the first one (distance 1, twiddle 1) alone
      for (uint32_t t = tid; t < M / 2; t += nth) {
        Cx<F> u = a[2 * t], v = a[2 * t + 1];
        a[2 * t] = cx_add<F>(u, v);
        a[2 * t + 1] = cx_sub<F>(u, v);
      }
      __syncthreads();
      li = 2;
    }
    for (; li + 1 <= logM; li += 2) {  // stages li and li+1 together: block 4q = 2^(li+1)
      const uint32_t lgb = li + 1, q = 1u << (li - 1), step = M >> lgb;
      for (uint32_t t = tid; t < M / 4; t += nth) {
        const uint32_t j = t & (q - 1), i = (t >> (li - 1)) << lgb, e1 = 2 * j * step;
        const Cx<F> x0 = a[i + j];
        Cx<F> A = a[i + j + q], B = a[i + j + 2 * q], C = a[i + j + 3 * q];
        if (j != 0) {
          A = cx_mul<F>(A, Wpow(2 * e1));
          B = cx_mul<F>(B, Wpow(e1));
          C = cx_mul<F>(C, Wpow(3 * e1));
        }
        const Cx<F> s0 = cx_add<F>(x0, A), s1 = cx_sub<F>(x0, A), s2 = cx_add<F>(B, C),
                    s3 = mulJ(cx_sub<F>(B, C));
        a[i + j] = cx_add<F>(s0, s2);
        a[i + j + q] = cx_add<F>(s1, s3);
        a[i + j + 2 * q] = cx_sub<F>(s0, s2);
        a[i + j + 3 * q] = cx_sub<F>(s1, s3);
      }
      __syncthreads();
    }
  }
  for (uint32_t k = n + tid; k < m; k += nth) {
    const Cx<F>& q = a[k >> 1];
    y[k] = F::mul(lead[k - (n - 1)], (k & 1) ? q.im : q.re);
  }
}

// wire <-> Montgomery conversion of flat element arrays (boundary of the C ABI)
template <class F>
__global__ void k_from_wire(const uint8_t* __restrict__ in, typename F::Elt* __restrict__ out, size_t n,
                            int* __restrict__ bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok = true;
  out[i] = F::from_bytes(in + i * F::kBytes, &ok);
  if (!ok) *bad = 1;
}
template <class F>
__global__ void k_to_wire(const typename F::Elt* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[F::kWords];
  F::to_wire(w, in[i]);
#pragma unroll
  for (int k = 0; k < F::kWords; ++k) out[i * F::kWords + k] = w[k];
}

// ---------------------------------------------------------------------------
// Merkle leaves: one thread per (instance, column).  leaf_j =
// SHA256(nonce_j || bytes(T[0][dblock+j]) || ... || bytes(T[nrow-1][dblock+j])).
// A warp reads 32 consecutive elements of a row: 128-bit coalesced loads.
// ---------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_merkle_leaves(const typename F::Elt* __restrict__ tableau, size_t tab_batch_stride, uint32_t nrow,
                uint32_t block_enc, uint32_t dblock, uint32_t block_ext,
                const uint8_t* __restrict__ nonces, size_t nonce_batch_stride,
                uint32_t* __restrict__ nodes /* [batch][2*block_ext][8] big-endian digest words */,
                size_t nodes_batch_stride,
                const uint32_t* __restrict__ rej /* or null: per-instance count of redrawn sample slots */,
                size_t rej_stride) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= block_ext) return;
  const typename F::Elt* T = tableau + (size_t)blockIdx.y * tab_batch_stride + dblock + j;
  // prime fields: every redrawn sample (k_zk_rng_scan) pushes the nonces one slot back in the caller's stream
  const size_t shift = rej ? (size_t)rej[(size_t)blockIdx.y * rej_stride] * F::kBytes : 0;
  const uint8_t* nzb = nonces + (size_t)blockIdx.y * nonce_batch_stride + shift + 32ull * j;
  uint32_t h[8], w[16];
  sha256_iv(h);
  if ((reinterpret_cast<uintptr_t>(nzb) & 3) == 0) {
    const uint32_t* nz = reinterpret_cast<const uint32_t*>(nzb);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = bswap32(nz[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      w[k] = ((uint32_t)nzb[4 * k] << 24) | ((uint32_t)nzb[4 * k + 1] << 16) | ((uint32_t)nzb[4 * k + 2] << 8) |
             (uint32_t)nzb[4 * k + 3];
  }
  constexpr int EW = F::kWords;        // words per element
  constexpr int PER = 16 / EW;         // elements per 64-byte block
  uint32_t pos = 8;                    // words filled in w (always a multiple of EW)
  uint64_t total = 32 + (uint64_t)nrow * F::kBytes;
  uint32_t i = 0;
  // first block: nonce + (8/EW) elements
  {
#pragma unroll
    for (int e = 0; e < 8 / EW; ++e) {
      if (i < nrow) {
        uint32_t ww[EW];
        F::to_wire(ww, T[(size_t)i * block_enc]);
#pragma unroll
        for (int k = 0; k < EW; ++k) w[8 + e * EW + k] = bswap32(ww[k]);
        ++i;
        pos += EW;
      }
    }
  }
  while (pos == 16) {
    sha256_compress(h, w);
    pos = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      if (i < nrow) {
        uint32_t ww[EW];
        F::to_wire(ww, T[(size_t)i * block_enc]);
#pragma unroll
        for (int k = 0; k < EW; ++k) w[e * EW + k] = bswap32(ww[k]);
        ++i;
        pos += EW;
      }
    }
  }
  // padding: pos in {0, EW, .., 16-EW} words are valid
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if ((uint32_t)k >= pos) w[k] = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if ((uint32_t)k == pos) w[k] = 0x80000000u;
  if (pos >= 14) {
    sha256_compress(h, w);
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = 0;
  }
  w[14] = (uint32_t)((total * 8) >> 32);
  w[15] = (uint32_t)(total * 8);
  sha256_compress(h, w);
  uint32_t* out = nodes + (size_t)blockIdx.y * nodes_batch_stride + 8ull * (block_ext + j);
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = h[k];
}

// node i = SHA256(node 2i || node 2i+1); digests kept as 8 big-endian words
__device__ __forceinline__ void merkle_hash2(uint32_t* __restrict__ nodes, uint32_t i) {
  uint32_t h[8], w[16];
  sha256_iv(h);
  const uint4* c = reinterpret_cast<const uint4*>(nodes + 16ull * i);
  uint4 a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  w[0] = a0.x; w[1] = a0.y; w[2] = a0.z; w[3] = a0.w;
  w[4] = a1.x; w[5] = a1.y; w[6] = a1.z; w[7] = a1.w;
  w[8] = a2.x; w[9] = a2.y; w[10] = a2.z; w[11] = a2.w;
  w[12] = a3.x; w[13] = a3.y; w[14] = a3.z; w[15] = a3.w;
  sha256_compress(h, w);
  w[0] = 0x80000000u;
#pragma unroll
  for (int k = 1; k < 15; ++k) w[k] = 0;
  w[15] = 512;
  sha256_compress(h, w);
  uint4* o = reinterpret_cast<uint4*>(nodes + 8ull * i);
  o[0] = make_uint4(h[0], h[1], h[2], h[3]);
  o[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// One CTA per instance walks the heap level by level (arbitrary n: the nodes
// of one level are i in [2^d, 2^(d+1)) below n; their children are either
// deeper inner nodes or leaves).
__global__ void __launch_bounds__(256)
k_merkle_tree(uint32_t* __restrict__ nodes, size_t nodes_batch_stride, uint32_t n) {
  uint32_t* N = nodes + (size_t)blockIdx.x * nodes_batch_stride;
  if (n < 2) return;
  int top = 31 - __clz(n - 1);  // deepest level that has an inner node
  for (int d = top; d >= 0; --d) {
    uint32_t lo = 1u << d, hi = min(2u << d, n);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) merkle_hash2(N, i);
    __syncthreads();
  }
}

}  // namespace lf
