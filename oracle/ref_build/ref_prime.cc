// oracle/_ref/libref.so -- TEST INFRASTRUCTURE (see ref_common.cc).
// Prime-field transforms of the reference: FFT<F>::fftb/fftf
// (algebra/fft.h:185-201), ReedSolomon::interpolate (algebra/reed_solomon.h:93-110)
// over the fields of BASELINE config 1 (algebra/fft_test.cc:33-44,168-172,
// algebra/reed_solomon_test.cc:337-401).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "algebra/bogorng.h"
#include "algebra/convolution.h"
#include "algebra/fft.h"
#include "algebra/fp.h"
#include "algebra/fp2.h"
#include "algebra/fp_p128.h"
#include "algebra/reed_solomon.h"
#include "algebra/static_string.h"
#include "ec/p256.h"

namespace proofs {
namespace {
// field ids local to this file (NOT proto FieldID): see oracle/refapi.py
//  100 = Fp<4> BN254-size prime of fft_test.cc:33-36 (omega order 2^28)
//  101 = Fp128 (fp_p128.h) omega order 2^32 (reed_solomon_test.cc:358-360)
//  102 = Fp<1> Goldilocks 2^64-2^32+1, omega 1753635133440165772 order 2^32
//  1   = Fp256 P-256 base field (no FFT; RS via Fp2)
using Fp4 = Fp<4>;
const Fp4& bn() {
  static const Fp4 f(
      "2188824287183927522224640574525727508854836440041603434369820418657580"
      "8495617");
  return f;
}
const Fp128<>& f128() {
  static const Fp128<> f;
  return f;
}
const Fp<1>& gold() {
  static const Fp<1> f("18446744069414584321");
  return f;
}

template <class Field>
void fft_t(const Field& F, const StaticString omega_s, uint64_t order, uint8_t* data,
           size_t n, int fwd) {
  using Elt = typename Field::Elt;
  Elt omega = F.of_string(omega_s);
  std::vector<Elt> A(n);
  for (size_t i = 0; i < n; ++i) A[i] = F.of_bytes_field(data + i * Field::kBytes).value();
  if (fwd) FFT<Field>::fftf(A.data(), n, omega, order, F);
  else FFT<Field>::fftb(A.data(), n, omega, order, F);
  for (size_t i = 0; i < n; ++i) F.to_bytes_field(data + i * Field::kBytes, A[i]);
}

template <class Field>
void rs_t(const Field& F, const StaticString omega_s, uint64_t order, uint8_t* rows,
          size_t n, size_t m, size_t nrows) {
  using Elt = typename Field::Elt;
  using Conv = FFTConvolutionFactory<Field>;
  Elt omega = F.of_string(omega_s);
  Conv conv(F, omega, order);
  ReedSolomon<Field, Conv> rs(n, m, F, conv);
  std::vector<Elt> y(m);
  for (size_t r = 0; r < nrows; ++r) {
    uint8_t* p = rows + r * m * Field::kBytes;
    for (size_t i = 0; i < n; ++i) y[i] = F.of_bytes_field(p + i * Field::kBytes).value();
    rs.interpolate(y.data());
    for (size_t i = 0; i < m; ++i) F.to_bytes_field(p + i * Field::kBytes, y[i]);
  }
}
// timing of the same two calls on one host thread (the reference's BM_FFT_* / BM_ReedSolomon* bodies,
// algebra/fft_test.cc:185-251, algebra/reed_solomon_test.cc:337-401): seconds per call, data prepared outside
template <class Field>
double fft_bench_t(const Field& F, const StaticString omega_s, uint64_t order, size_t n, int reps) {
  using Elt = typename Field::Elt;
  Elt omega = F.of_string(omega_s);
  std::vector<Elt> A(n);
  for (size_t i = 0; i < n; ++i) A[i] = F.of_scalar(i * 2654435761u + 1);
  FFT<Field>::fftb(A.data(), n, omega, order, F);
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) FFT<Field>::fftb(A.data(), n, omega, order, F);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}
template <class Field>
double rs_bench_t(const Field& F, const StaticString omega_s, uint64_t order, size_t n, size_t m, int reps) {
  using Elt = typename Field::Elt;
  using Conv = FFTConvolutionFactory<Field>;
  Elt omega = F.of_string(omega_s);
  Conv conv(F, omega, order);
  ReedSolomon<Field, Conv> rs(n, m, F, conv);
  std::vector<Elt> y(m);
  for (size_t i = 0; i < n; ++i) y[i] = F.of_scalar(i * 2654435761u + 1);
  rs.interpolate(y.data());
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) rs.interpolate(y.data());
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}
const char kBnOmega[] =
    "1910321906792171394429139282769207003614565195732928631530564200482146216"
    "1904";
const char kF128Omega[] = "164956748514267535023998284330560247862";
const char kGoldOmega[] = "1753635133440165772";
static constexpr char kRootX[] =
    "112649224146410281873500457609690258373018840430489408729223714171582664"
    "680802";
static constexpr char kRootY[] =
    "840879943585409076957404614278186605601821689971823787493130182544504602"
    "12908";
}  // namespace
}  // namespace proofs

using namespace proofs;
extern "C" {
// in-place, data = n elements in to_bytes_field encoding. fwd=0: fftb, 1: fftf
int ref_fft(int fid, uint8_t* data, size_t n, int fwd) {
  if (fid == 100) fft_t(bn(), StaticString(kBnOmega), 1ull << 28, data, n, fwd);
  else if (fid == 101) fft_t(f128(), StaticString(kF128Omega), 1ull << 32, data, n, fwd);
  else if (fid == 102) fft_t(gold(), StaticString(kGoldOmega), 1ull << 32, data, n, fwd);
  else return -1;
  return 0;
}
// FFT over Fp2(P-256) (fft_test.cc:168-172): data = n x (re 32B, im 32B)
int ref_fft_p256_2(uint8_t* data, size_t n, int fwd) {
  using F2 = Fp2<Fp256Base>;
  static const F2 f2(p256_base);
  auto omega = f2.of_string(kRootX, kRootY);
  std::vector<F2::Elt> A(n);
  for (size_t i = 0; i < n; ++i) {
    A[i].re = p256_base.of_bytes_field(data + 64 * i).value();
    A[i].im = p256_base.of_bytes_field(data + 64 * i + 32).value();
  }
  if (fwd) FFT<F2>::fftf(A.data(), n, omega, 1ull << 31, f2);
  else FFT<F2>::fftb(A.data(), n, omega, 1ull << 31, f2);
  for (size_t i = 0; i < n; ++i) {
    p256_base.to_bytes_field(data + 64 * i, A[i].re);
    p256_base.to_bytes_field(data + 64 * i + 32, A[i].im);
  }
  return 0;
}
// rows: nrows x m elements, first n valid
int ref_rs(int fid, uint8_t* rows, size_t n, size_t m, size_t nrows) {
  if (fid == 100) rs_t(bn(), StaticString(kBnOmega), 1ull << 28, rows, n, m, nrows);
  else if (fid == 101) rs_t(f128(), StaticString(kF128Omega), 1ull << 32, rows, n, m, nrows);
  else if (fid == 102) rs_t(gold(), StaticString(kGoldOmega), 1ull << 32, rows, n, m, nrows);
  else if (fid == 1) {
    using F2 = Fp2<Fp256Base>;
    using Conv = FFTExtConvolutionFactory<Fp256Base, F2>;
    static const F2 f2(p256_base);
    static const Conv conv(p256_base, f2, f2.of_string(kRootX, kRootY), 1ull << 31);
    ReedSolomon<Fp256Base, Conv> rs(n, m, p256_base, conv);
    std::vector<Fp256Base::Elt> y(m);
    for (size_t r = 0; r < nrows; ++r) {
      uint8_t* p = rows + r * m * 32;
      for (size_t i = 0; i < n; ++i) y[i] = p256_base.of_bytes_field(p + 32 * i).value();
      rs.interpolate(y.data());
      for (size_t i = 0; i < m; ++i) p256_base.to_bytes_field(p + 32 * i, y[i]);
    }
  } else return -1;
  return 0;
}
// seconds per FFT<Field>::fftb of n points on one thread (fid 1: Fp2 over P-256, BM_FFT_Fp256_2)
double ref_fft_bench(int fid, size_t n, int reps) {
  if (fid == 100) return fft_bench_t(bn(), StaticString(kBnOmega), 1ull << 28, n, reps);
  if (fid == 101) return fft_bench_t(f128(), StaticString(kF128Omega), 1ull << 32, n, reps);
  if (fid == 102) return fft_bench_t(gold(), StaticString(kGoldOmega), 1ull << 32, n, reps);
  if (fid == 1) {
    using F2 = Fp2<Fp256Base>;
    static const F2 f2(p256_base);
    auto omega = f2.of_string(kRootX, kRootY);
    std::vector<F2::Elt> A(n);
    for (size_t i = 0; i < n; ++i) {
      A[i].re = p256_base.of_scalar(i * 2654435761u + 1);
      A[i].im = p256_base.of_scalar(i + 7);
    }
    FFT<F2>::fftb(A.data(), n, omega, 1ull << 31, f2);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) FFT<F2>::fftb(A.data(), n, omega, 1ull << 31, f2);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  }
  return -1;
}
// seconds per ReedSolomon(n, m)::interpolate of one row on one thread (BM_ReedSolomonFp*)
double ref_rs_bench(int fid, size_t n, size_t m, int reps) {
  if (fid == 100) return rs_bench_t(bn(), StaticString(kBnOmega), 1ull << 28, n, m, reps);
  if (fid == 101) return rs_bench_t(f128(), StaticString(kF128Omega), 1ull << 32, n, m, reps);
  if (fid == 102) return rs_bench_t(gold(), StaticString(kGoldOmega), 1ull << 32, n, m, reps);
  if (fid == 1) {
    using F2 = Fp2<Fp256Base>;
    using Conv = FFTExtConvolutionFactory<Fp256Base, F2>;
    static const F2 f2(p256_base);
    static const Conv conv(p256_base, f2, f2.of_string(kRootX, kRootY), 1ull << 31);
    ReedSolomon<Fp256Base, Conv> rs(n, m, p256_base, conv);
    std::vector<Fp256Base::Elt> y(m);
    for (size_t i = 0; i < n; ++i) y[i] = p256_base.of_scalar(i * 2654435761u + 1);
    rs.interpolate(y.data());
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) rs.interpolate(y.data());
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  }
  return -1;
}
// elementwise Montgomery-domain-free multiply in byte encoding
int ref_fp_mul(int fid, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  auto go = [&](const auto& F) {
    using Field = std::decay_t<decltype(F)>;
    for (size_t i = 0; i < n; ++i) {
      auto x = F.of_bytes_field(a + i * Field::kBytes).value();
      auto y = F.of_bytes_field(b + i * Field::kBytes).value();
      F.to_bytes_field(out + i * Field::kBytes, F.mulf(x, y));
    }
  };
  if (fid == 100) go(bn());
  else if (fid == 101) go(f128());
  else if (fid == 102) go(gold());
  else if (fid == 1) go(p256_base);
  else return -1;
  return 0;
}
}
