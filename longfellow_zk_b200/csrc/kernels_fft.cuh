// Stand-alone radix-2 FFT over prime fields (and over Fp2 of P-256), the
// counterpart of FFT<Field>::fftb / fftf (lib/algebra/fft.h:47-202):
//   fftb: T[j] = sum_k F[k] w^{jk}   (unnormalised),   fftf: the same with w^-1.
// The reference bit-reverses and runs log2(n) butterfly passes (recursing
// above 16384 points, fft.h:125-153); any exact DFT algorithm produces the same
// array.  Here a transform is at most three launches: an in-place bit reversal
// and one or two "stage group" kernels, each of which loads a tile into shared
// memory, runs up to kFftTileLog consecutive stages there and stores it back:
//   element index = hi * 2^(s0+nst) + mid * 2^s0 + lo,   mid in [0, 2^nst)
// a CTA owns one hi, 2^nst mids and a run of consecutive lo (coalesced rows).
// Twiddles come from one table w^k, k < n/2, per (field, n); the inverse root
// uses w^-k = -w^(n/2-k).
#pragma once
#include <stdint.h>

#include "field.cuh"
#include "kernels_commit.cuh"

namespace lf {

// algebra adaptors: scalar field elements and Fp2 elements share the kernels
template <class F>
struct AlgF {
  typedef typename F::Elt Elt;
  __device__ static __forceinline__ Elt add(const Elt& a, const Elt& b) { return F::add(a, b); }
  __device__ static __forceinline__ Elt sub(const Elt& a, const Elt& b) { return F::sub(a, b); }
  __device__ static __forceinline__ Elt mul(const Elt& a, const Elt& b) { return F::mul(a, b); }
  __device__ static __forceinline__ Elt neg(const Elt& a) { return F::neg(a); }
};
template <class F>
struct AlgCx {
  typedef Cx<F> Elt;
  __device__ static __forceinline__ Elt add(const Elt& a, const Elt& b) { return cx_add<F>(a, b); }
  __device__ static __forceinline__ Elt sub(const Elt& a, const Elt& b) { return cx_sub<F>(a, b); }
  __device__ static __forceinline__ Elt mul(const Elt& a, const Elt& b) { return cx_mul<F>(a, b); }
  __device__ static __forceinline__ Elt neg(const Elt& a) { return Elt{F::neg(a.re), F::neg(a.im)}; }
};

template <class A>
__global__ void k_fft_bitrev(typename A::Elt* __restrict__ a, uint32_t logn, size_t batch_stride) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logn)) return;
  a += (size_t)blockIdx.y * batch_stride;
  uint32_t j = logn ? (__brev(i) >> (32 - logn)) : 0;
  if (i < j) {
    typename A::Elt t = a[i];
    a[i] = a[j];
    a[j] = t;
  }
}

// twiddle w^(+-e) for e < n/2 from the positive table tw[k] = w^k, k < n/2
template <class A>
__device__ __forceinline__ typename A::Elt fft_tw(const typename A::Elt* __restrict__ tw, uint32_t e, uint32_t half_n,
                                                 bool inverse_root) {
  if (!inverse_root || e == 0) return tw[e];
  return A::neg(tw[half_n - e]);
}

// nst stages starting at global stage s0 (butterfly distance 2^s0 .. 2^(s0+nst-1)).
// dif == 0: decimation in time (stages ascending, v *= w first);
// dif == 1: decimation in frequency (stages descending, (u-v) *= w after).
// grid.x = number of tiles = n / (2^nst * L); dynamic smem = 2^nst * L elements.
template <class A>
__global__ void __launch_bounds__(256)
k_fft_stages(typename A::Elt* __restrict__ a, size_t batch_stride, uint32_t logn, uint32_t s0, uint32_t nst,
             uint32_t logL, const typename A::Elt* __restrict__ tw, int inverse_root, int dif) {
  typedef typename A::Elt Elt;
  a += (size_t)blockIdx.y * batch_stride;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Elt* t = reinterpret_cast<Elt*>(smem_raw);
  const uint32_t L = 1u << logL, nm = 1u << nst, tile = nm << logL;
  const uint32_t lo_tiles = (1u << s0) >> logL;  // tiles per hi along lo
  const uint32_t hi = blockIdx.x / lo_tiles, lo0 = (blockIdx.x % lo_tiles) << logL;
  const size_t base = ((size_t)hi << (s0 + nst)) + lo0;
  for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) {
    uint32_t mid = i >> logL, l = i & (L - 1);
    t[i] = a[base + ((size_t)mid << s0) + l];
  }
  __syncthreads();
  const uint32_t half_n = 1u << (logn - 1);
  for (uint32_t qq = 0; qq < nst; ++qq) {
    const uint32_t q = dif ? (nst - 1 - qq) : qq;
    const uint32_t s = s0 + q;  // global stage: distance 2^s
    for (uint32_t b = threadIdx.x; b < tile / 2; b += blockDim.x) {
      uint32_t l = b & (L - 1), mb = b >> logL;  // mb in [0, nm/2)
      uint32_t mlow = mb & ((1u << q) - 1), mhigh = mb >> q;
      uint32_t m0 = (mhigh << (q + 1)) | mlow, m1 = m0 | (1u << q);
      // position inside the length-2^(s+1) block, then the exponent of w_n
      uint32_t j = (mlow << s0) + lo0 + l;
      uint32_t e = j << (logn - s - 1);
      Elt w = fft_tw<A>(tw, e, half_n, inverse_root != 0);
      Elt u = t[(m0 << logL) + l], v = t[(m1 << logL) + l];
      if (dif) {
        t[(m0 << logL) + l] = A::add(u, v);
        Elt dlt = A::sub(u, v);
        t[(m1 << logL) + l] = e == 0 ? dlt : A::mul(dlt, w);
      } else {
        Elt tv = e == 0 ? v : A::mul(v, w);
        t[(m0 << logL) + l] = A::add(u, tv);
        t[(m1 << logL) + l] = A::sub(u, tv);
      }
    }
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) {
    uint32_t mid = i >> logL, l = i & (L - 1);
    a[base + ((size_t)mid << s0) + l] = t[i];
  }
}

// out[i] = a[i] * b[i]
template <class A>
__global__ void k_fft_pointwise(typename A::Elt* __restrict__ a, const typename A::Elt* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[(size_t)blockIdx.y * n + i] = A::mul(a[(size_t)blockIdx.y * n + i], b[i]);
}

// RS through FFTConvolution (lib/algebra/convolution.h:55-106) for fields with
// 2-power roots: x = binom * y zero-padded to N.  Work row blockIdx.y is row (row0 + blockIdx.y) of the
// caller's rows: row r of batch b sits at rows + b * batch_stride + r * row_stride (rpb rows per batch).
struct RsRows {
  size_t row_stride, batch_stride;
  uint32_t rpb, row0;
  __device__ __forceinline__ size_t at(uint32_t y) const {
    const uint32_t g = row0 + y;
    return (size_t)(g / rpb) * batch_stride + (size_t)(g % rpb) * row_stride;
  }
};
template <class F>
__global__ void k_rs_pad(const typename F::Elt* __restrict__ rows, RsRows R, typename F::Elt* __restrict__ x,
                         uint32_t n, uint32_t N, const typename F::Elt* __restrict__ binom) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const typename F::Elt* y = rows + R.at(blockIdx.y);
  x[(size_t)blockIdx.y * N + i] = i < n ? F::mul(binom[i], y[i]) : F::zero();
}
// y[k] = lead[k-d] * conv[k], k in [n, m); conv is stored bit-reversed or natural
template <class F>
__global__ void k_rs_finish(typename F::Elt* __restrict__ rows, RsRows R, const typename F::Elt* __restrict__ x,
                            uint32_t n, uint32_t m, uint32_t N, const typename F::Elt* __restrict__ lead) {
  uint32_t k = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  typename F::Elt* y = rows + R.at(blockIdx.y);
  y[k] = F::mul(lead[k - (n - 1)], x[(size_t)blockIdx.y * N + k]);
}

// the same pad / finish steps for a real row carried in the real parts of Fp2
// elements (P-256 rows too long for the shared-memory real-FFT kernel)
template <class F>
__global__ void k_rs_pad_cx(const typename F::Elt* __restrict__ rows, RsRows R, Cx<F>* __restrict__ x,
                            uint32_t n, uint32_t N, const typename F::Elt* __restrict__ binom) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const typename F::Elt* y = rows + R.at(blockIdx.y);
  x[(size_t)blockIdx.y * N + i] = Cx<F>{i < n ? F::mul(binom[i], y[i]) : F::zero(), F::zero()};
}
template <class F>
__global__ void k_rs_finish_cx(typename F::Elt* __restrict__ rows, RsRows R, const Cx<F>* __restrict__ x,
                               uint32_t n, uint32_t m, uint32_t N, const typename F::Elt* __restrict__ lead) {
  uint32_t k = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  typename F::Elt* y = rows + R.at(blockIdx.y);
  y[k] = F::mul(lead[k - (n - 1)], x[(size_t)blockIdx.y * N + k].re);
}

}  // namespace lf
