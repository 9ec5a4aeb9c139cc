// liblongfellow_b200.so -- single translation unit: device kernels + host
// engine + C ABI (include/longfellow_b200.h).  No CPU fallback: every compute
// entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <set>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/longfellow_b200.h"
#include "field.cuh"
#include "hash.cuh"
#include "kernels_commit.cuh"
#include "kernels_fft.cuh"
#include "kernels_rscrt.cuh"
#include "kernels_zk.cuh"
#include "kernels_scflat.cuh"
#include "kernels_verify.cuh"
#include "zk_types.cuh"

namespace lf {

// ---------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define LF_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return fail(LF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
  } while (0)

// ---------------------------------------------------------------- GF host tables
// All computed with the same gf128.cuh code the kernels run.
struct GfHost {
  gf128 g;
  gf128 beta[16];
  gf128 what[16][16];
  GfConsts consts;
  std::vector<gf128> tw;  // 65535 twiddles, see kernels_commit.cuh

  GfHost() {
    // gf2_128.h:369-391 subfield_generator
    gf128 x = gf_zero();
    x.w[0] = 2;
    gf128 r = x;
    for (int i = 4; i < 7; ++i) {
      gf128 s = r;
      for (int j = 0; j < (1 << i); ++j) s = gf_mul(s, s);
      r = gf_mul(r, s);
    }
    g = r;
    beta[0] = gf_one();
    for (int i = 1; i < 16; ++i) beta[i] = gf_mul(beta[i - 1], g);
    memset(&consts, 0, sizeof(consts));
    for (int i = 0; i < 16; ++i) consts.beta[i] = beta[i];
    // gf2_128.h:451-493 beta_ref
    uint32_t u[16][4];
    uint32_t linv[16];
    for (int i = 0; i < 16; ++i) {
      for (int k = 0; k < 4; ++k) u[i][k] = beta[i].w[k];
      linv[i] = 1u << i;
    }
    auto bit = [](const uint32_t* v, int j) { return (v[j >> 5] >> (j & 31)) & 1u; };
    int rnk = 0;
    for (int j = 0; rnk < 16 && j < 128; ++j) {
      int piv = -1;
      for (int i = rnk; i < 16; ++i)
        if (bit(u[i], j)) {
          piv = i;
          break;
        }
      if (piv < 0) continue;
      for (int k = 0; k < 4; ++k) std::swap(u[rnk][k], u[piv][k]);
      std::swap(linv[rnk], linv[piv]);
      consts.sub_ldnz[rnk] = (uint32_t)j;
      for (int i1 = rnk + 1; i1 < 16; ++i1)
        if (bit(u[i1], j)) {
          for (int k = 0; k < 4; ++k) u[i1][k] ^= u[rnk][k];
          linv[i1] ^= linv[rnk];
        }
      ++rnk;
    }
    for (int i = 0; i < 16; ++i) {
      for (int k = 0; k < 4; ++k) consts.sub_u[i][k] = u[i][k];
      consts.sub_linv[i] = linv[i];
    }
    // gf2_128.h:121-137 evaluation points 0, 1, g and Newton denominators
    consts.evalpt[0] = gf_zero();
    consts.evalpt[1] = gf_one();
    consts.evalpt[2] = g;
    for (int i = 1; i < 3; ++i)
      for (int k = 2; k >= i; --k)
        consts.newton[k][i] = gf_inv(gf_add(consts.evalpt[k], consts.evalpt[k - i]));
    // poly.h:125-137 dot_interpolation identity_[k] (Newton form of Lagrange basis)
    for (int k = 0; k < 3; ++k) {
      gf128 t[3] = {gf_zero(), gf_zero(), gf_zero()};
      t[k] = gf_one();
      for (int i = 1; i < 3; ++i)
        for (int kk = 2; kk >= i; --kk) t[kk] = gf_mul(gf_add(t[kk], t[kk - 1]), consts.newton[kk][i]);
      for (int i = 0; i < 3; ++i) consts.lag_id[k][i] = t[i];
    }
    // lch14.h:45-77 w_hat
    for (int j = 0; j < 16; ++j) what[0][j] = beta[j];
    for (int i = 0; i + 1 < 16; ++i)
      for (int j = 0; j < 16; ++j) what[i + 1][j] = gf_mul(what[i][j], gf_add(what[i][j], what[i][i]));
    for (int i = 0; i < 16; ++i) {
      gf128 sc = gf_inv(what[i][i]);
      for (int j = 0; j < 16; ++j) what[i][j] = gf_mul(sc, what[i][j]);
    }
    // T_i[u] = sum_k bit_k(u) w_hat[i][i+1+k]   (lch14.h:81-100)
    tw.assign(65536, gf_zero());
    for (uint32_t i = 0; i < 16; ++i) {
      uint32_t off = 65536u - (1u << (16 - i));
      uint32_t cnt = 1u << (15 - i);
      tw[off] = gf_zero();
      for (uint32_t v = 1; v < cnt; ++v) {
        uint32_t low = v & (0u - v);
        uint32_t k = (uint32_t)__builtin_ctz(low);
        tw[off + v] = gf_add(tw[off + (v ^ low)], what[i][i + 1 + k]);
      }
    }
  }
};
static const GfHost& gf_host() {
  static const GfHost h;
  return h;
}

// LCH14 BidirectionalFFT(l, k) flattened into butterfly steps (lch14.h:185-217)
static void emit(std::vector<RsStep>& out, uint32_t kind, uint32_t stage, uint32_t base, uint32_t t0,
                 uint32_t t1) {
  if (t0 < t1) out.push_back(RsStep{kind, stage, base, t0, t1});
}
static void gen_bidir(std::vector<RsStep>& out, uint32_t i, uint32_t base, uint32_t k) {
  if (i-- > 0) {
    uint32_t s = 1u << i;
    if (k < s) {
      emit(out, 0, i, base, k, s);
      gen_bidir(out, i, base, k);
      emit(out, 2, i, base, 0, k);
      for (uint32_t j = i; j-- > 0;) emit(out, 0, j, base + s, 0, s / 2);  // FFT(i, coset+s, B+s)
    } else {
      for (uint32_t j = 0; j < i; ++j) emit(out, 1, j, base, 0, s / 2);  // IFFT(i, coset, B)
      emit(out, 2, i, base, k - s, s);
      gen_bidir(out, i, base + s, k - s);
      emit(out, 1, i, base, 0, k - s);
    }
  }
}

// Consecutive steps that both start at t0 == 0 on the same base with stages <= 5: a warp's 32
// aligned butterflies touch only its own 64-element block in both, so the barrier between them
// can be a __syncwarp() (k_rs_gf_rows; CTA sizes are multiples of 32).
static void rs_mark_wsync(std::vector<RsStep>& steps) {
  for (size_t k = 0; k + 1 < steps.size(); ++k) {
    const RsStep &a = steps[k], &b = steps[k + 1];
    steps[k].wsync = (a.t0 == 0 && b.t0 == 0 && a.base == b.base && a.stage <= 5 && b.stage <= 5) ? 1u : 0u;
  }
}

struct RsPlanHost {
  RsPlan plan;
  RsStep* d_steps = nullptr;
  std::vector<RsStep> steps;  // host copy: the large-row path launches one kernel per step
};

static void parse_dec_words(const char* s, uint32_t* out, int W) {
  for (int i = 0; i < W; ++i) out[i] = 0;
  for (; *s; ++s) {
    uint64_t c = (uint64_t)(*s - '0');
    for (int i = 0; i < W; ++i) {
      c += (uint64_t)out[i] * 10;
      out[i] = (uint32_t)c;
      c >>= 32;
    }
  }
}

// host arithmetic of a Montgomery field with W 32-bit limbs (same limb code as the device)
template <int W>
struct FpHostT {
  FpConsts<W> C;
  fpw<W> omega;            // root of unity of order 2^omega_log (Montgomery form), if any
  uint32_t omega_log = 0;
  explicit FpHostT(const char* modulus_dec, const char* omega_dec = nullptr, uint32_t olog = 0) {
    uint32_t m[W];
    parse_dec_words(modulus_dec, m, W);
    fp_build_consts<W>(m, &C);
    if (omega_dec) {
      uint32_t o[W];
      parse_dec_words(omega_dec, o, W);
      from_wire((const uint8_t*)o, &omega);
      omega_log = olog;
    }
  }
  fpw<W> mul(const fpw<W>& a, const fpw<W>& b) const { return fp_mul_generic<W>(a, b, C.m, C.mprime); }
  fpw<W> add(const fpw<W>& a, const fpw<W>& b) const { return fp_add<W>(a, b, C.m); }
  fpw<W> sub(const fpw<W>& a, const fpw<W>& b) const { return fp_sub<W>(a, b, C.m); }
  fpw<W> one() const {
    fpw<W> r;
    for (int i = 0; i < W; ++i) r.w[i] = C.one[i];
    return r;
  }
  fpw<W> zero() const {
    fpw<W> r;
    for (int i = 0; i < W; ++i) r.w[i] = 0;
    return r;
  }
  fpw<W> inv(const fpw<W>& a) const {
    uint32_t e[W], two[W];
    for (int i = 0; i < W; ++i) two[i] = 0;
    two[0] = 2;
    fp_subn<W>(e, C.m, two);
    return fp_pow_host<W>(a, e, C);
  }
  bool from_wire(const uint8_t* p, fpw<W>* out) const {
    fpw<W> a, q;
    memcpy(a.w, p, 4 * W);
    if (fp_geq<W>(a.w, C.m)) return false;
    for (int i = 0; i < W; ++i) q.w[i] = C.rsq[i];
    *out = mul(a, q);
    return true;
  }
  void to_wire(uint8_t* p, const fpw<W>& a) const {
    fpw<W> o = zero();
    o.w[0] = 1;
    fpw<W> r = mul(a, o);
    memcpy(p, r.w, 4 * W);
  }
  // primitive 2^logn-th root
  fpw<W> root(uint32_t logn) const {
    fpw<W> w = omega;
    for (uint32_t o = omega_log; o > logn; --o) w = mul(w, w);
    return w;
  }
};
// the benchmark fields of lib/algebra/fft_test.cc:33-44 and reed_solomon_test.cc:337-401
static const FpHostT<8>& bn254_host() {
  static const FpHostT<8> h("21888242871839275222246405745257275088548364400416034343698204186575808495617",
                            "19103219067921713944291392827692070036145651957329286315305642004821462161904", 28);
  return h;
}
static const FpHostT<4>& fp128_host() {
  static const FpHostT<4> h("340282042402384805036647824275747635201", "164956748514267535023998284330560247862", 32);
  return h;
}
static const FpHostT<2>& gold_host() {
  static const FpHostT<2> h("18446744069414584321", "1753635133440165772", 32);
  return h;
}

// host copy of the P-256 constant block and helpers on Montgomery limbs
struct P256Host {
  FpConsts<8> C;
  P256Host() {
    // lib/algebra/fp_p256.h:34-39
    const uint32_t m[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0, 1, 0xFFFFFFFFu};
    fp_build_consts<8>(m, &C);
  }
  fpw<8> mul(const fpw<8>& a, const fpw<8>& b) const { return fp_mul_generic<8>(a, b, C.m, C.mprime); }
  fpw<8> add(const fpw<8>& a, const fpw<8>& b) const { return fp_add<8>(a, b, C.m); }
  fpw<8> sub(const fpw<8>& a, const fpw<8>& b) const { return fp_sub<8>(a, b, C.m); }
  fpw<8> one() const {
    fpw<8> r;
    for (int i = 0; i < 8; ++i) r.w[i] = C.one[i];
    return r;
  }
  fpw<8> zero() const {
    fpw<8> r;
    for (int i = 0; i < 8; ++i) r.w[i] = 0;
    return r;
  }
  fpw<8> inv(const fpw<8>& a) const {
    uint32_t e[8], two[8] = {2, 0, 0, 0, 0, 0, 0, 0};
    fp_subn<8>(e, C.m, two);
    return fp_pow_host<8>(a, e, C);
  }
  bool from_wire(const uint8_t* p, fpw<8>* out) const {
    fpw<8> a, q;
    memcpy(a.w, p, 32);
    if (fp_geq<8>(a.w, C.m)) return false;
    for (int i = 0; i < 8; ++i) q.w[i] = C.rsq[i];
    *out = mul(a, q);
    return true;
  }
  void to_wire(uint8_t* p, const fpw<8>& a) const {
    fpw<8> o = zero();
    o.w[0] = 1;
    fpw<8> r = mul(a, o);
    memcpy(p, r.w, 32);
  }
};
static const P256Host& p256_host() {
  static const P256Host h;
  return h;
}

// tables of ReedSolomon(n, m) (lib/algebra/reed_solomon.h:51-88) in Montgomery form
struct RsFpTables {
  fpw<8>* d_inv = nullptr;    // [m] 1/i, inv[0] = 0
  fpw<8>* d_lead = nullptr;   // [m-n+1]
  fpw<8>* d_binom = nullptr;  // [n]
  // FFT path (k_rs_fp_fft_rows): N = padding, M = N/2
  uint32_t logM = 0;
  Cx<FFp256>* d_wk = nullptr;  // [M] W^k
  Cx<FFp256>* d_yh = nullptr;  // [M+1] DFT_N(1/i table) / (2N)
};

// host Fp2 over P-256 (Montgomery limbs)
struct HCx {
  fpw<8> re, im;
};
static HCx hcx_mul(const P256Host& H, const HCx& a, const HCx& b) {
  return HCx{H.sub(H.mul(a.re, b.re), H.mul(a.im, b.im)), H.add(H.mul(a.re, b.im), H.mul(a.im, b.re))};
}
static void parse_dec256(const char* s, uint32_t out[8]) {
  for (int i = 0; i < 8; ++i) out[i] = 0;
  for (; *s; ++s) {
    uint64_t c = (uint64_t)(*s - '0');
    for (int i = 0; i < 8; ++i) {
      c += (uint64_t)out[i] * 10;
      out[i] = (uint32_t)c;
      c >>= 32;
    }
  }
}
// primitive N-th root of unity on the unit circle of Fp2(P-256): the element of
// order 2^31 of lib/circuits/mdoc/mdoc_zk.cc:83-88 / ecdsa/verify_test.cc:519-530, squared down
static HCx p256_root(const P256Host& H, uint32_t logN) {
  uint32_t x[8], y[8];
  parse_dec256("112649224146410281873500457609690258373018840430489408729223714171582664680802", x);
  parse_dec256("84087994358540907695740461427818660560182168997182378749313018254450460212908", y);
  HCx w;
  H.from_wire((const uint8_t*)x, &w.re);
  H.from_wire((const uint8_t*)y, &w.im);
  for (uint32_t o = 31; o > logN; --o) w = hcx_mul(H, w, w);
  return w;
}

}  // namespace lf

using namespace lf;

// ---------------------------------------------------------------- context
struct lf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  gf128* d_tw = nullptr;
  std::map<std::pair<size_t, size_t>, RsPlanHost> rs_plans;
  std::map<std::pair<size_t, size_t>, RsFpTables> rs_fp;
  std::map<std::pair<int, uint32_t>, void*> fft_tw;             // (field, logn) -> w^k, k < n/2
  std::map<std::pair<int, std::pair<size_t, size_t>>, void*> rs_conv;  // (field,(n,m)) -> RsConvTables*
  uint64_t launches = 0;
  int sm_count = 0;
  // cudaFuncSetAttribute acts on the CURRENT device: which kernels already have their attributes on
  // this context's device (a process may hold contexts on several GPUs)
  uint32_t attr_mask = 0;
  std::set<const void*> fft_attr;
  // grow-only work arrays of the large-row RS / convolution paths (stream ordered: a later call on this
  // context's stream reuses them only after the earlier kernels are done)
  // CRT Reed-Solomon over P-256 (kernels_rscrt.cuh): constants, twiddles per transform size, spectra per (n, m)
  void* d_crt_consts = nullptr;
  std::vector<uint32_t> crt_p, crt_w;                  // primes, a generator-derived root of order 2^13 each
  std::map<uint32_t, std::pair<uint32_t*, uint32_t*>> crt_tw;        // logN -> (forward, inverse) [17][N/2]
  std::map<std::pair<size_t, size_t>, uint32_t*> crt_spec;           // (n, m) -> [17][N]
  int cluster_fit[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}};  // resident clusters of 16 / 8 / 4 / 2 CTAs, per field
  void* work[2] = {nullptr, nullptr};
  size_t work_cap[2] = {0, 0};
};
enum { kAttrRsGf = 1, kAttrRsFp = 2, kAttrScCluster = 4 /* and 8: the prime-field cluster kernel */, kAttrRsCrt = 16, kAttrBindTab = 32 };

namespace lf {

static const size_t kWorkChunkBytes = (size_t)2 << 30;  // rows are taken in chunks whose work array stays below this
static int ctx_work(lf_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes > ctx->work_cap[slot]) {
    LF_CUDA(cudaStreamSynchronize(ctx->stream));  // kernels of an earlier call may still use the old array
    cudaFree(ctx->work[slot]);
    ctx->work[slot] = nullptr;
    ctx->work_cap[slot] = 0;
    LF_CUDA(cudaMalloc(&ctx->work[slot], bytes));
    ctx->work_cap[slot] = bytes;
  }
  *out = ctx->work[slot];
  return 0;
}

static const int kRsGfThreads = 128;  // measured: 64 / 128 / 192 / 256 threads -> 11.77 / 11.11 / 11.84 / 11.76 ms per 1024 SHA proofs

static int ctx_rs_plan(lf_ctx* ctx, size_t n, size_t m, RsPlanHost** out) {
  auto key = std::make_pair(n, m);
  auto it = ctx->rs_plans.find(key);
  if (it == ctx->rs_plans.end()) {
    RsPlanHost ph;
    uint32_t l = 0, fftn = 1;
    while (fftn < n) {
      fftn <<= 1;
      ++l;
    }
    std::vector<RsStep> steps;
    gen_bidir(steps, l, 0, (uint32_t)n);
    rs_mark_wsync(steps);
    ph.plan.n = (uint32_t)n;
    ph.plan.m = (uint32_t)m;
    ph.plan.l = l;
    ph.plan.fftn = fftn;
    ph.plan.nsteps = (uint32_t)steps.size();
    if (!steps.empty()) {
      LF_CUDA(cudaMalloc(&ph.d_steps, steps.size() * sizeof(RsStep)));
      LF_CUDA(cudaMemcpyAsync(ph.d_steps, steps.data(), steps.size() * sizeof(RsStep),
                              cudaMemcpyHostToDevice, ctx->stream));
      LF_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ph.plan.steps = ph.d_steps;
    ph.steps = steps;
    it = ctx->rs_plans.emplace(key, ph).first;
  }
  *out = &it->second;
  return 0;
}

// One step of an LCH14 schedule over `rows` global-memory work arrays.
static void gf_gstep(lf_ctx* ctx, gf128* work, size_t work_stride, size_t rows, const RsStep& st, uint32_t coset) {
  if (st.t1 <= st.t0) return;
  dim3 grid((st.t1 - st.t0 + 255) / 256, (unsigned)rows);
  k_rs_gf_gstep<FGf128><<<grid, 256, 0, ctx->stream>>>(work, work_stride, st, coset, ctx->d_tw);
  ctx->launches++;
}

// Rows whose first coset does not fit in shared memory: the schedule of
// k_rs_gf_rows with one launch per step (lch14_reed_solomon.h:49-103).
static int launch_rs_gf_global(lf_ctx* ctx, gf128* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                               size_t nbatch, const RsPlanHost& ph) {
  const uint32_t n = ph.plan.n, m = ph.plan.m, l = ph.plan.l, fftn = ph.plan.fftn;
  gf128 *C = nullptr, *D = nullptr;
  // rows of one batch in chunks that keep the two work arrays bounded
  const size_t chunk = std::max<size_t>(1, std::min(nrows, kWorkChunkBytes / ((size_t)fftn * sizeof(gf128))));
  int rc = ctx_work(ctx, 0, chunk * (size_t)fftn * sizeof(gf128), (void**)&C);
  if (rc) return rc;
  if (m > fftn && (rc = ctx_work(ctx, 1, chunk * (size_t)fftn * sizeof(gf128), (void**)&D))) return rc;
  for (size_t bi = 0; bi < nbatch; ++bi)
    for (size_t r0 = 0; r0 < nrows; r0 += chunk) {
      const size_t nr = std::min(chunk, nrows - r0);
      const dim3 cgrid((fftn + 255) / 256, (unsigned)nr);
      gf128* y = d_rows + bi * batch_stride + r0 * row_stride;
      // C = y[0..n) zero-extended; truncated transform on the first coset
      k_rs_gf_copy<FGf128><<<cgrid, 256, 0, ctx->stream>>>(C, fftn, y, row_stride, fftn, 0, n);
      ctx->launches++;
      for (const RsStep& st : ph.steps) gf_gstep(ctx, C, fftn, nr, st, 0);
      // evaluations n..fftn of the first coset go out; C keeps the n coefficients
      if (fftn > n) {
        const uint32_t hi = std::min(fftn, m);
        LF_CUDA(cudaMemcpy2DAsync(y + n, row_stride * sizeof(gf128), C + n, (size_t)fftn * sizeof(gf128),
                                  (size_t)(hi - n) * sizeof(gf128), nr, cudaMemcpyDeviceToDevice, ctx->stream));
      }
      for (uint32_t b = fftn; b < m; b += fftn) {
        k_rs_gf_copy<FGf128><<<cgrid, 256, 0, ctx->stream>>>(D, fftn, C, fftn, fftn, 0, n);
        ctx->launches++;
        for (uint32_t st = l; st-- > 0;) gf_gstep(ctx, D, fftn, nr, RsStep{0, st, 0, 0, fftn / 2}, b);
        const uint32_t cnt = std::min(fftn, m - b);
        LF_CUDA(cudaMemcpy2DAsync(y + b, row_stride * sizeof(gf128), D, (size_t)fftn * sizeof(gf128),
                                  (size_t)cnt * sizeof(gf128), nr, cudaMemcpyDeviceToDevice, ctx->stream));
      }
    }
  LF_CUDA(cudaGetLastError());
  return 0;
}

// LCH14::FFT(l, coset, B) / IFFT (lch14.h:106-146) on one device array of 2^l elements
static int launch_lch14_fft(lf_ctx* ctx, gf128* d, uint32_t l, uint32_t coset, bool forward, size_t nrows = 1) {
  const uint32_t n = 1u << l;
  if (forward) {
    for (uint32_t st = l; st-- > 0;) gf_gstep(ctx, d, n, nrows, RsStep{0, st, 0, 0, n / 2}, coset);
  } else {
    for (uint32_t st = 0; st < l; ++st) gf_gstep(ctx, d, n, nrows, RsStep{1, st, 0, 0, n / 2}, coset);
  }
  LF_CUDA(cudaGetLastError());
  return 0;
}

// rows: device, GF(2^128).  grid = (nrows, nbatch)
static int launch_rs_gf(lf_ctx* ctx, gf128* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                        size_t nbatch, size_t n, size_t m) {
  if (n == 0 || m < n || m > 65536) return fail(LF_ERR_ARG, "rs: need 0 < n <= m <= 65536");
  if (nrows == 0 || nbatch == 0 || m == n) return 0;
  RsPlanHost* ph;
  int rc = ctx_rs_plan(ctx, n, m, &ph);
  if (rc) return rc;
  size_t smem = 2 * (size_t)ph->plan.fftn * sizeof(gf128);
  if (smem > 200 * 1024) return launch_rs_gf_global(ctx, d_rows, row_stride, nrows, batch_stride, nbatch, *ph);
  if (!(ctx->attr_mask & kAttrRsGf)) {
    LF_CUDA(cudaFuncSetAttribute(k_rs_gf_rows<FGf128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    ctx->attr_mask |= kAttrRsGf;
  }
  dim3 grid((unsigned)nrows, (unsigned)nbatch);
  // CTA size: LF_RS_GF_THREADS (32..256, multiple of 32) is a tuning knob; the kernel strides by blockDim.x
  static int rs_threads = 0;
  if (!rs_threads) {
    const char* e = getenv("LF_RS_GF_THREADS");
    int v = e ? atoi(e) : 0;
    rs_threads = (v >= 32 && v <= 256 && v % 32 == 0) ? v : kRsGfThreads;
  }
  k_rs_gf_rows<FGf128><<<grid, rs_threads, smem, ctx->stream>>>(d_rows, row_stride, batch_stride, ph->plan,
                                                                ctx->d_tw);
  ctx->launches++;
  LF_CUDA(cudaGetLastError());
  return 0;
}

// inv[i] = 1/i, lead, binom of ReedSolomon(n, m) (lib/algebra/reed_solomon.h:51-88,
// lib/algebra/utility.h:51-72) in Montgomery form, for any host field H
template <class H, class E>
static void rs_tables_host(const H& Hf, size_t n, size_t m, std::vector<E>& inv, std::vector<E>& lead,
                           std::vector<E>& binom) {
  const size_t d = n - 1;
  std::vector<E> sc(m);
  inv.resize(m);
  lead.resize(m - n + 1);
  binom.resize(n);
  sc[0] = Hf.zero();
  for (size_t i = 1; i < m; ++i) sc[i] = Hf.add(sc[i - 1], Hf.one());
  {
    std::vector<E> pre(m);
    E p = Hf.one();
    inv[0] = Hf.zero();
    for (size_t i = 1; i < m; ++i) {
      pre[i] = p;
      p = Hf.mul(p, sc[i]);
    }
    p = Hf.inv(p);
    for (size_t i = m; i-- > 1;) {
      inv[i] = Hf.mul(pre[i], p);
      p = Hf.mul(p, sc[i]);
    }
  }
  // reed_solomon.h:62-79 leading constants (-1)^d (k-d) C(k,d)
  lead[0] = Hf.one();
  for (size_t i = 1; i + d < m; ++i) lead[i] = Hf.mul(lead[i - 1], Hf.mul(sc[d + i], inv[i]));
  for (size_t k = d; k < m; ++k) {
    lead[k - d] = Hf.mul(lead[k - d], sc[k - d]);
    if (d % 2 == 1) lead[k - d] = Hf.sub(Hf.zero(), lead[k - d]);
  }
  // reed_solomon.h:81-87 (-1)^i C(d, i)
  binom[0] = Hf.one();
  for (size_t i = 1; i < n; ++i) binom[i] = Hf.mul(binom[i - 1], Hf.mul(sc[n - i], inv[i]));
  for (size_t i = 1; i < n; i += 2) binom[i] = Hf.sub(Hf.zero(), binom[i]);
}

static int ctx_rs_fp_tables(lf_ctx* ctx, size_t n, size_t m, RsFpTables** out) {
  auto key = std::make_pair(n, m);
  auto it = ctx->rs_fp.find(key);
  if (it == ctx->rs_fp.end()) {
    const P256Host& H = p256_host();
    std::vector<fpw<8>> inv, lead, binom;
    rs_tables_host(H, n, m, inv, lead, binom);
    RsFpTables t;
    {
      // FFT tables: N = smallest power of two >= m (convolution.h:47-53 choose_padding)
      uint32_t logN = 1;
      while (((size_t)1 << logN) < m) ++logN;
      const uint32_t N = 1u << logN, M = N / 2;
      t.logM = logN - 1;
      HCx W = p256_root(H, logN);
      std::vector<HCx> wn(N);
      wn[0] = HCx{H.one(), H.zero()};
      for (uint32_t k = 1; k < N; ++k) wn[k] = hcx_mul(H, wn[k - 1], W);
      // Yh = fftf_N(inv padded): radix-2 DIT with twiddles W^-e = conj(W^e)
      std::vector<HCx> a(N);
      auto brev = [logN](uint32_t k) {
        uint32_t r = 0;
        for (uint32_t i = 0; i < logN; ++i) r |= ((k >> i) & 1u) << (logN - 1 - i);
        return r;
      };
      for (uint32_t j = 0; j < N; ++j) a[brev(j)] = HCx{j < m ? inv[j] : H.zero(), H.zero()};
      for (uint32_t len = 2; len <= N; len <<= 1) {
        uint32_t half = len / 2, step = N / len;
        for (uint32_t i = 0; i < N; i += len)
          for (uint32_t j = 0; j < half; ++j) {
            HCx tw = wn[j * step];
            tw.im = H.sub(H.zero(), tw.im);
            HCx tv = hcx_mul(H, a[i + j + half], tw), u = a[i + j];
            a[i + j] = HCx{H.add(u.re, tv.re), H.add(u.im, tv.im)};
            a[i + j + half] = HCx{H.sub(u.re, tv.re), H.sub(u.im, tv.im)};
          }
      }
      fpw<8> twoN = H.zero();
      for (uint32_t i = 0; i < 2 * N; ++i) twoN = H.add(twoN, H.one());
      fpw<8> s = H.inv(twoN);
      std::vector<HCx> yh(M + 1), wk(M);
      for (uint32_t k = 0; k <= M; ++k) yh[k] = HCx{H.mul(a[k].re, s), H.mul(a[k].im, s)};
      for (uint32_t k = 0; k < M; ++k) wk[k] = wn[k];
      static_assert(sizeof(HCx) == sizeof(Cx<FFp256>), "Cx layout");
      LF_CUDA(cudaMalloc(&t.d_wk, wk.size() * sizeof(HCx)));
      LF_CUDA(cudaMalloc(&t.d_yh, yh.size() * sizeof(HCx)));
      LF_CUDA(cudaMemcpy(t.d_wk, wk.data(), wk.size() * sizeof(HCx), cudaMemcpyHostToDevice));
      LF_CUDA(cudaMemcpy(t.d_yh, yh.data(), yh.size() * sizeof(HCx), cudaMemcpyHostToDevice));
    }
    LF_CUDA(cudaMalloc(&t.d_inv, m * 32));
    LF_CUDA(cudaMalloc(&t.d_lead, lead.size() * 32));
    LF_CUDA(cudaMalloc(&t.d_binom, n * 32));
    LF_CUDA(cudaMemcpy(t.d_inv, inv.data(), m * 32, cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(t.d_lead, lead.data(), lead.size() * 32, cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(t.d_binom, binom.data(), n * 32, cudaMemcpyHostToDevice));
    it = ctx->rs_fp.emplace(key, t).first;
  }
  *out = &it->second;
  return 0;
}


// ---- CRT Reed-Solomon over P-256: host tables (kernels_rscrt.cuh) ---------------------------------
namespace crt {
static uint32_t mulmod(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t)((uint64_t)a * b % p); }
static uint32_t powmod(uint32_t a, uint64_t e, uint32_t p) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r = mulmod(r, a, p);
    a = mulmod(a, a, p);
    e >>= 1;
  }
  return r;
}
static bool is_prime(uint32_t n) {
  if (n < 2 || n % 2 == 0) return n == 2;
  for (uint32_t i = 3; (uint64_t)i * i <= n; i += 2)
    if (n % i == 0) return false;
  return true;
}
constexpr uint32_t kLogMax = 13;  // transforms up to 2^13 points
// x mod p for a 256-bit integer given as 8 little-endian limbs
static uint32_t limbs_mod(const uint32_t* w, uint32_t p) {
  uint64_t r = 0;
  for (int l = 7; l >= 0; --l) r = ((r << 32) | w[l]) % p;
  return (uint32_t)r;
}
}  // namespace crt

static int ctx_crt_consts(lf_ctx* ctx) {
  if (ctx->d_crt_consts) return 0;
  const P256Host& H = p256_host();
  // the 17 largest primes k 2^13 + 1 below 2^31 and, for each, an element of order 2^13
  std::vector<uint32_t> pr, wr;
  for (uint32_t k = (1u << (31 - crt::kLogMax)) - 1; pr.size() < (size_t)kCrtPrimes; --k) {
    const uint32_t p = (k << crt::kLogMax) + 1;
    if (!crt::is_prime(p)) continue;
    uint32_t g = 2;
    for (;; ++g) {
      if (crt::powmod(g, (p - 1) / 2, p) == 1) continue;  // a quadratic residue cannot have full 2-power order
      const uint32_t w = crt::powmod(g, (p - 1) >> crt::kLogMax, p);
      if (crt::powmod(w, 1u << (crt::kLogMax - 1), p) == p - 1) {
        wr.push_back(w);
        break;
      }
    }
    pr.push_back(p);
  }
  auto small = [&](uint32_t v) {
    fpw<8> r = H.zero();
    r.w[0] = v;
    return r;
  };
  fpw<8> rsq;
  for (int i = 0; i < 8; ++i) rsq.w[i] = H.C.rsq[i];
  // raw integers modulo P: a*b = mul(mul(a, R^2), b);  a*R^-1 = mul(a, 1)
  auto mulraw = [&](const fpw<8>& a, const fpw<8>& b) { return H.mul(H.mul(a, rsq), b); };
  CrtConsts cc;
  memset(&cc, 0, sizeof(cc));
  for (int j = 0; j < kCrtPrimes; ++j) {
    const uint32_t p = pr[j];
    CrtPrime& q = cc.pr[j];
    q.p = p;
    uint32_t inv = p;  // Newton: p * inv = 1 mod 2^32
    for (int it = 0; it < 5; ++it) inv *= 2 - p * inv;
    q.pinv = inv;
    uint64_t t = ((uint64_t)1 << 32) % p;  // 2^32 mod p
    const uint32_t r32 = (uint32_t)t;
    uint32_t cur = r32;  // 2^(32 l) * 2^32 for l = 0
    for (int l = 0; l < 8; ++l) {
      q.k[l] = cur;
      cur = crt::mulmod(cur, r32, p);
    }
    q.f = (uint32_t)(((uint64_t)1 << kCrtQBits) / p);
    fpw<8> x = small(1);
    for (int i = 0; i < kCrtPrimes; ++i)
      if (i != j) x = mulraw(x, small(pr[i]));
    const fpw<8> c = H.mul(x, small(1));  // (M / p_j) R^-1 mod P
    for (int l = 0; l < 8; ++l) cc.C[j][l] = c.w[l];
  }
  {
    fpw<8> x = small(1);
    for (int i = 0; i < kCrtPrimes; ++i) x = mulraw(x, small(pr[i]));
    const fpw<8> dd = H.mul(x, small(1));  // M R^-1 mod P
    fpw<8> e = H.zero();
    for (int q = 0; q <= kCrtPrimes; ++q) {
      for (int l = 0; l < 8; ++l) cc.E[q][l] = e.w[l];
      e = H.sub(e, dd);
    }
  }
  LF_CUDA(cudaMalloc(&ctx->d_crt_consts, sizeof(cc)));
  LF_CUDA(cudaMemcpy(ctx->d_crt_consts, &cc, sizeof(cc), cudaMemcpyHostToDevice));
  ctx->crt_p = pr;
  ctx->crt_w = wr;
  return 0;
}

// forward / inverse twiddles w^k 2^32 mod p, k < N/2, for every prime
static int ctx_crt_twiddles(lf_ctx* ctx, uint32_t logN, uint32_t** f, uint32_t** b) {
  auto it = ctx->crt_tw.find(logN);
  if (it == ctx->crt_tw.end()) {
    const uint32_t N = 1u << logN, half = N / 2;
    std::vector<uint32_t> tf((size_t)kCrtPrimes * half), tb((size_t)kCrtPrimes * half);
    for (int j = 0; j < kCrtPrimes; ++j) {
      const uint32_t p = ctx->crt_p[j];
      const uint32_t w = crt::powmod(ctx->crt_w[j], 1u << (crt::kLogMax - logN), p), wi = crt::powmod(w, p - 2, p);
      const uint32_t r32 = (uint32_t)(((uint64_t)1 << 32) % p);
      uint32_t a = r32, c = r32;
      for (uint32_t k = 0; k < half; ++k) {
        tf[(size_t)j * half + k] = a;
        tb[(size_t)j * half + k] = c;
        a = crt::mulmod(a, w, p);
        c = crt::mulmod(c, wi, p);
      }
    }
    uint32_t *df = nullptr, *db = nullptr;
    LF_CUDA(cudaMalloc(&df, tf.size() * 4));
    LF_CUDA(cudaMalloc(&db, tb.size() * 4));
    LF_CUDA(cudaMemcpy(df, tf.data(), tf.size() * 4, cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(db, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice));
    it = ctx->crt_tw.emplace(logN, std::make_pair(df, db)).first;
  }
  *f = it->second.first;
  *b = it->second.second;
  return 0;
}

// spectrum of the 1/i table modulo every prime, with N^-1, (M/p)^-1 and the Montgomery factor folded in,
// in the bit-reversed order the DIF transform leaves
static int ctx_crt_spectrum(lf_ctx* ctx, size_t n, size_t m, uint32_t logN, const std::vector<fpw<8>>& inv,
                            uint32_t** out) {
  auto key = std::make_pair(n, m);
  auto it = ctx->crt_spec.find(key);
  if (it == ctx->crt_spec.end()) {
    const uint32_t N = 1u << logN;
    std::vector<uint32_t> sp((size_t)kCrtPrimes * N);
    std::vector<uint32_t> a(N);
    for (int j = 0; j < kCrtPrimes; ++j) {
      const uint32_t p = ctx->crt_p[j];
      const uint32_t w = crt::powmod(ctx->crt_w[j], 1u << (crt::kLogMax - logN), p);
      for (uint32_t i = 0; i < N; ++i) a[i] = i < m ? crt::limbs_mod(inv[i].w, p) : 0;
      // the same decimation-in-frequency network as the kernel's
      for (uint32_t s = 0; s < logN; ++s) {
        const uint32_t lh = logN - 1 - s, half = 1u << lh;
        std::vector<uint32_t> wp(half);
        const uint32_t ws = crt::powmod(w, 1u << s, p);
        wp[0] = 1;
        for (uint32_t k = 1; k < half; ++k) wp[k] = crt::mulmod(wp[k - 1], ws, p);
        for (uint32_t t = 0; t < N / 2; ++t) {
          const uint32_t lo = t & (half - 1), i0 = ((t >> lh) << (lh + 1)) + lo, i1 = i0 + half;
          const uint32_t x = a[i0], y = a[i1];
          a[i0] = (x + y) % p;
          a[i1] = crt::mulmod((x + p - y) % p, wp[lo], p);
        }
      }
      uint32_t mj = 1;  // (M / p_j) mod p_j
      for (int i = 0; i < kCrtPrimes; ++i)
        if (i != j) mj = crt::mulmod(mj, ctx->crt_p[i] % p, p);
      uint32_t scale = crt::powmod(mj, p - 2, p);
      scale = crt::mulmod(scale, crt::powmod(N % p, p - 2, p), p);
      scale = crt::mulmod(scale, (uint32_t)(((uint64_t)1 << 32) % p), p);
      for (uint32_t i = 0; i < N; ++i) sp[(size_t)j * N + i] = crt::mulmod(a[i], scale, p);
    }
    uint32_t* d = nullptr;
    LF_CUDA(cudaMalloc(&d, sp.size() * 4));
    LF_CUDA(cudaMemcpy(d, sp.data(), sp.size() * 4, cudaMemcpyHostToDevice));
    it = ctx->crt_spec.emplace(key, d).first;
  }
  *out = it->second;
  return 0;
}

static int launch_rs_p256_crt(lf_ctx* ctx, fpw<8>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                              size_t nbatch, size_t n, size_t m, const RsFpTables* t) {
  uint32_t logN = 1;
  while (((size_t)1 << logN) < m) ++logN;
  int rc = ctx_crt_consts(ctx);
  if (rc) return rc;
  uint32_t *twf, *twb, *spec;
  if ((rc = ctx_crt_twiddles(ctx, logN, &twf, &twb))) return rc;
  auto sit = ctx->crt_spec.find(std::make_pair(n, m));
  if (sit == ctx->crt_spec.end()) {
    std::vector<fpw<8>> inv, lead, binom;
    rs_tables_host(p256_host(), n, m, inv, lead, binom);
    if ((rc = ctx_crt_spectrum(ctx, n, m, logN, inv, &spec))) return rc;
  } else {
    spec = sit->second;
  }
  const size_t smem = (8 * n + ((size_t)1 << logN) + ((size_t)1 << logN) / 32 + 32) * 4;
  if (!(ctx->attr_mask & kAttrRsCrt)) {
    LF_CUDA(cudaFuncSetAttribute(k_rs_crt_rows<FFp256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ctx->attr_mask |= kAttrRsCrt;
  }
  // residue scratch: rows in chunks that keep it bounded (batches, then rows of a batch)
  const size_t per_row = (size_t)kCrtPrimes * (m - n) * 4;
  const size_t max_rows = std::max<size_t>(1, kWorkChunkBytes / per_row);
  uint32_t* scratch;
  const size_t total = nrows * nbatch;
  if ((rc = ctx_work(ctx, 0, std::min(total, max_rows) * per_row, (void**)&scratch))) return rc;
  const size_t bstep = std::max<size_t>(1, max_rows / nrows);  // whole batches per launch
  if (nrows > max_rows) return fail(LF_ERR_UNSUPPORTED, "rs: too many rows per batch for the CRT path");
  for (size_t b0 = 0; b0 < nbatch; b0 += bstep) {
    const size_t nb = std::min(bstep, nbatch - b0);
    k_rs_crt_rows<FFp256><<<dim3((unsigned)nrows, (unsigned)nb), 256, smem, ctx->stream>>>(
        d_rows + b0 * batch_stride, row_stride, batch_stride, (uint32_t)n, (uint32_t)m, logN,
        (const CrtConsts*)ctx->d_crt_consts, twf, twb, spec, t->d_lead, t->d_binom, scratch);
    ctx->launches++;
  }
  LF_CUDA(cudaGetLastError());
  return 0;
}

static int rs_conv_run_p256(lf_ctx* ctx, fpw<8>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                            size_t nbatch, size_t n, size_t m);
static int launch_rs_p256(lf_ctx* ctx, fpw<8>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                          size_t nbatch, size_t n, size_t m) {
  if (n == 0 || m < n || m > (1u << 24)) return fail(LF_ERR_ARG, "rs: need 0 < n <= m <= 2^24");
  if (nrows == 0 || nbatch == 0 || m == n) return 0;
  if (m > 6400) return rs_conv_run_p256(ctx, d_rows, row_stride, nrows, batch_stride, nbatch, n, m);
  RsFpTables* t;
  int rc = ctx_rs_fp_tables(ctx, n, m, &t);
  if (rc) return rc;
  if (!(ctx->attr_mask & kAttrRsFp)) {
    LF_CUDA(cudaFuncSetAttribute(k_rs_fp_rows<FFp256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LF_CUDA(cudaFuncSetAttribute(k_rs_fp_fft_rows<FFp256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ctx->attr_mask |= kAttrRsFp;
  }
  dim3 grid((unsigned)nrows, (unsigned)nbatch);
  static const bool force_direct = getenv("LF_RS_DIRECT") != nullptr;
  const size_t fft_smem = ((size_t)64) << t->logM;
  // the direct Toeplitz sum costs n*(m-n) multiplications, the FFT ~ 10 N log N
  // LF_RS_P256=fft keeps the Fp2 real-FFT kernel of round 1 (for comparison); default: the CRT transform
  static const bool use_fft = getenv("LF_RS_P256") && !strcmp(getenv("LF_RS_P256"), "fft");
  if (!force_direct && !use_fft && n * (m - n) > 4096 && m <= (1u << crt::kLogMax) && (8 * n + 3 * m) * 4 <= 200 * 1024)
    return launch_rs_p256_crt(ctx, d_rows, row_stride, nrows, batch_stride, nbatch, n, m, t);
  if (!force_direct && n * (m - n) > 4096 && fft_smem <= 200 * 1024) {
    // one radix-4 group per thread and pass: M/4 threads (the 455 -> 909 rows of the Ligero
    // prove have M = 512: 128 threads and four CTAs per SM instead of 512 mostly idle threads)
    const unsigned fft_threads = (unsigned)std::min<size_t>(512, std::max<size_t>(64, ((size_t)1 << t->logM) / 4));
    k_rs_fp_fft_rows<FFp256><<<grid, fft_threads, fft_smem, ctx->stream>>>(d_rows, row_stride, batch_stride, (uint32_t)n,
                                                                (uint32_t)m, t->logM, t->d_wk, t->d_yh, t->d_lead,
                                                                t->d_binom);
  } else {
    k_rs_fp_rows<FFp256><<<grid, 256, n * 32, ctx->stream>>>(d_rows, row_stride, batch_stride, (uint32_t)n,
                                                            (uint32_t)m, t->d_inv, t->d_lead, t->d_binom);
  }
  ctx->launches++;
  LF_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- stand-alone FFT
template <class A>
static int fft_run(lf_ctx* ctx, typename A::Elt* d, size_t batch_stride, size_t nbatch, uint32_t logn,
                   const typename A::Elt* d_tw, bool inverse_root, bool dif, bool bitrev) {
  typedef typename A::Elt Elt;
  if (logn == 0 || nbatch == 0) return 0;
  // one flag per (context, instantiation): set on the context's device, which is current here
  {
    static const char kTag = 0;  // address unique per instantiation of fft_run<A>
    if (!ctx->fft_attr.count(&kTag)) {
      LF_CUDA(cudaFuncSetAttribute(k_fft_stages<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      ctx->fft_attr.insert(&kTag);
    }
  }
  uint32_t tile_log = 0;
  while ((sizeof(Elt) << (tile_log + 1)) <= 64 * 1024) ++tile_log;
  struct Grp {
    uint32_t s0, nst, logL;
  };
  std::vector<Grp> groups;
  uint32_t s = std::min(tile_log, logn);
  groups.push_back(Grp{0, s, 0});
  while (s < logn) {
    uint32_t logL = std::min<uint32_t>(s, 5);
    uint32_t nst = std::min(logn - s, tile_log - logL);
    groups.push_back(Grp{s, nst, logL});
    s += nst;
  }
  const unsigned nb = (unsigned)nbatch;
  auto do_bitrev = [&]() {
    k_fft_bitrev<A><<<dim3(((1u << logn) + 255) / 256, nb), 256, 0, ctx->stream>>>(d, logn, batch_stride);
    ctx->launches++;
  };
  if (!dif && bitrev) do_bitrev();
  for (size_t gi = 0; gi < groups.size(); ++gi) {
    const Grp& g = dif ? groups[groups.size() - 1 - gi] : groups[gi];
    const uint32_t tile = 1u << (g.nst + g.logL);
    k_fft_stages<A><<<dim3((1u << logn) / tile, nb), 256, (size_t)tile * sizeof(Elt), ctx->stream>>>(
        d, batch_stride, logn, g.s0, g.nst, g.logL, d_tw, inverse_root ? 1 : 0, dif ? 1 : 0);
    ctx->launches++;
  }
  if (dif && bitrev) do_bitrev();
  LF_CUDA(cudaGetLastError());
  return 0;
}

// w^k, k < n/2 for the scalar fields
template <class H, int W>
static int fft_twiddles(lf_ctx* ctx, int field_id, const H& Hf, uint32_t logn, fpw<W>** out) {
  auto key = std::make_pair(field_id, logn);
  auto it = ctx->fft_tw.find(key);
  if (it == ctx->fft_tw.end()) {
    if (logn > Hf.omega_log) return fail(LF_ERR_ARG, "fft: n exceeds the order of the field's root of unity");
    const size_t half = logn ? ((size_t)1 << (logn - 1)) : 1;
    std::vector<fpw<W>> tw(half);
    fpw<W> w = Hf.root(logn);
    tw[0] = Hf.one();
    for (size_t k = 1; k < half; ++k) tw[k] = Hf.mul(tw[k - 1], w);
    void* dptr;
    LF_CUDA(cudaMalloc(&dptr, half * sizeof(fpw<W>)));
    LF_CUDA(cudaMemcpy(dptr, tw.data(), half * sizeof(fpw<W>), cudaMemcpyHostToDevice));
    it = ctx->fft_tw.emplace(key, dptr).first;
  }
  *out = (fpw<W>*)it->second;
  return 0;
}
// W^k, k < n/2 in Fp2 of P-256 (key: field id 1)
static int fft_twiddles_p256(lf_ctx* ctx, uint32_t logn, Cx<FFp256>** out) {
  auto key = std::make_pair((int)LF_FIELD_P256, logn);
  auto it = ctx->fft_tw.find(key);
  if (it == ctx->fft_tw.end()) {
    if (logn > 31) return fail(LF_ERR_ARG, "fft: n exceeds 2^31");
    const P256Host& H = p256_host();
    const size_t half = logn ? ((size_t)1 << (logn - 1)) : 1;
    std::vector<HCx> tw(half);
    HCx w = p256_root(H, logn);
    tw[0] = HCx{H.one(), H.zero()};
    for (size_t k = 1; k < half; ++k) tw[k] = hcx_mul(H, tw[k - 1], w);
    void* dptr;
    LF_CUDA(cudaMalloc(&dptr, half * sizeof(HCx)));
    LF_CUDA(cudaMemcpy(dptr, tw.data(), half * sizeof(HCx), cudaMemcpyHostToDevice));
    it = ctx->fft_tw.emplace(key, dptr).first;
  }
  *out = (Cx<FFp256>*)it->second;
  return 0;
}

// ReedSolomon over a field with 2-power roots = FFTConvolution
// (lib/algebra/convolution.h:55-106): fftf(x) * fftf(1/i table) / N -> fftb
struct RsConvTables {
  uint32_t logN = 0;
  void *d_yh = nullptr, *d_lead = nullptr, *d_binom = nullptr;
};
template <class F, class H>
static int rs_conv_run(lf_ctx* ctx, const H& Hf, typename F::Elt* d_rows, size_t row_stride, size_t nrows,
                       size_t batch_stride, size_t nbatch, size_t n, size_t m) {
  typedef typename F::Elt Elt;
  typedef AlgF<F> A;
  if (n == 0 || m < n) return fail(LF_ERR_ARG, "rs: need 0 < n <= m");
  if (nrows == 0 || nbatch == 0 || m == n) return 0;
  uint32_t logN = 0;
  while (((size_t)1 << logN) < m) ++logN;
  const size_t N = (size_t)1 << logN;
  Elt* d_tw;
  int rc = fft_twiddles<H, F::W>(ctx, F::kFieldId, Hf, logN, &d_tw);
  if (rc) return rc;
  auto key = std::make_pair((int)F::kFieldId, std::make_pair(n, m));
  auto it = ctx->rs_conv.find(key);
  if (it == ctx->rs_conv.end()) {
    std::vector<Elt> inv, lead, binom;
    rs_tables_host(Hf, n, m, inv, lead, binom);
    // fold the 1/N of the inverse transform into the leading constants
    Elt nn = Hf.zero();
    {
      // N as a field element by doubling
      nn = Hf.one();
      for (uint32_t i = 0; i < logN; ++i) nn = Hf.add(nn, nn);
    }
    Elt ninv = Hf.inv(nn);
    for (auto& v : lead) v = Hf.mul(v, ninv);
    auto* t = new RsConvTables;
    t->logN = logN;
    std::vector<Elt> ypad(N, Hf.zero());
    std::copy(inv.begin(), inv.end(), ypad.begin());
    LF_CUDA(cudaMalloc(&t->d_yh, N * sizeof(Elt)));
    LF_CUDA(cudaMalloc(&t->d_lead, lead.size() * sizeof(Elt)));
    LF_CUDA(cudaMalloc(&t->d_binom, n * sizeof(Elt)));
    LF_CUDA(cudaMemcpy(t->d_yh, ypad.data(), N * sizeof(Elt), cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(t->d_lead, lead.data(), lead.size() * sizeof(Elt), cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(t->d_binom, binom.data(), n * sizeof(Elt), cudaMemcpyHostToDevice));
    // spectrum of the 1/i table, kept in the bit-reversed order the DIF transform leaves
    if ((rc = fft_run<A>(ctx, (Elt*)t->d_yh, 0, 1, logN, d_tw, /*inverse_root=*/true, /*dif=*/true, false))) return rc;
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ctx->rs_conv.emplace(key, (void*)t).first;
  }
  auto* t = (RsConvTables*)it->second;
  // all rows of all batches, in chunks whose work array stays bounded; asynchronous on the context's stream
  const size_t total = nrows * nbatch;
  const size_t chunk = std::max<size_t>(1, std::min(total, kWorkChunkBytes / (N * sizeof(Elt))));
  Elt* x;
  if ((rc = ctx_work(ctx, 0, chunk * N * sizeof(Elt), (void**)&x))) return rc;
  for (size_t r0 = 0; r0 < total && rc == 0; r0 += chunk) {
    const size_t nr = std::min(chunk, total - r0);
    const RsRows R{row_stride, batch_stride, (uint32_t)nrows, (uint32_t)r0};
    k_rs_pad<F><<<dim3((unsigned)((N + 255) / 256), (unsigned)nr), 256, 0, ctx->stream>>>(
        d_rows, R, x, (uint32_t)n, (uint32_t)N, (const Elt*)t->d_binom);
    ctx->launches++;
    if ((rc = fft_run<A>(ctx, x, N, nr, logN, d_tw, true, true, false)) == 0) {
      k_fft_pointwise<A><<<dim3((unsigned)((N + 255) / 256), (unsigned)nr), 256, 0, ctx->stream>>>(
          x, (const Elt*)t->d_yh, N);
      ctx->launches++;
      rc = fft_run<A>(ctx, x, N, nr, logN, d_tw, false, false, false);
    }
    if (rc == 0) {
      k_rs_finish<F><<<dim3((unsigned)((m - n + 255) / 256), (unsigned)nr), 256, 0, ctx->stream>>>(
          d_rows, R, x, (uint32_t)n, (uint32_t)m, (uint32_t)N, (const Elt*)t->d_lead);
      ctx->launches++;
      cudaError_t ce = cudaGetLastError();
      if (ce != cudaSuccess) rc = fail(LF_ERR_CUDA, cudaGetErrorString(ce));
    }
  }
  return rc;
}

// P-256 rows too long for k_rs_fp_fft_rows: the same convolution with the real
// row embedded in Fp2 (imaginary parts zero) and global-memory stage-group FFTs
// over Fp2 (FFTExtConvolution, lib/algebra/convolution.h:128-191; exact, so the
// real part of the result is the convolution).
static int rs_conv_run_p256(lf_ctx* ctx, fpw<8>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                            size_t nbatch, size_t n, size_t m) {
  typedef FFp256 F;
  typedef AlgCx<F> A;
  typedef Cx<F> CxE;
  const P256Host& H = p256_host();
  uint32_t logN = 0;
  while (((size_t)1 << logN) < m) ++logN;
  const size_t N = (size_t)1 << logN;
  CxE* d_tw;
  int rc = fft_twiddles_p256(ctx, logN, &d_tw);
  if (rc) return rc;
  auto key = std::make_pair((int)LF_FIELD_P256, std::make_pair(n, m));
  auto it = ctx->rs_conv.find(key);
  if (it == ctx->rs_conv.end()) {
    std::vector<fpw<8>> inv, lead, binom;
    rs_tables_host(H, n, m, inv, lead, binom);
    fpw<8> nn = H.one();
    for (uint32_t i = 0; i < logN; ++i) nn = H.add(nn, nn);
    const fpw<8> ninv = H.inv(nn);
    for (auto& v : lead) v = H.mul(v, ninv);
    auto* t = new RsConvTables;
    t->logN = logN;
    std::vector<HCx> ypad(N, HCx{H.zero(), H.zero()});
    for (size_t i = 0; i < inv.size(); ++i) ypad[i].re = inv[i];
    LF_CUDA(cudaMalloc(&t->d_yh, N * sizeof(HCx)));
    LF_CUDA(cudaMalloc(&t->d_lead, lead.size() * 32));
    LF_CUDA(cudaMalloc(&t->d_binom, n * 32));
    LF_CUDA(cudaMemcpy(t->d_yh, ypad.data(), N * sizeof(HCx), cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(t->d_lead, lead.data(), lead.size() * 32, cudaMemcpyHostToDevice));
    LF_CUDA(cudaMemcpy(t->d_binom, binom.data(), n * 32, cudaMemcpyHostToDevice));
    if ((rc = fft_run<A>(ctx, (CxE*)t->d_yh, 0, 1, logN, d_tw, /*inverse_root=*/true, /*dif=*/true, false))) return rc;
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ctx->rs_conv.emplace(key, (void*)t).first;
  }
  auto* t = (RsConvTables*)it->second;
  const size_t total = nrows * nbatch;
  const size_t chunk = std::max<size_t>(1, std::min(total, kWorkChunkBytes / (N * sizeof(CxE))));
  CxE* x;
  if ((rc = ctx_work(ctx, 0, chunk * N * sizeof(CxE), (void**)&x))) return rc;
  for (size_t r0 = 0; r0 < total && rc == 0; r0 += chunk) {
    const size_t nr = std::min(chunk, total - r0);
    const RsRows R{row_stride, batch_stride, (uint32_t)nrows, (uint32_t)r0};
    k_rs_pad_cx<F><<<dim3((unsigned)((N + 255) / 256), (unsigned)nr), 256, 0, ctx->stream>>>(
        d_rows, R, x, (uint32_t)n, (uint32_t)N, (const fpw<8>*)t->d_binom);
    ctx->launches++;
    if ((rc = fft_run<A>(ctx, x, N, nr, logN, d_tw, true, true, false)) == 0) {
      k_fft_pointwise<A><<<dim3((unsigned)((N + 255) / 256), (unsigned)nr), 256, 0, ctx->stream>>>(
          x, (const CxE*)t->d_yh, N);
      ctx->launches++;
      rc = fft_run<A>(ctx, x, N, nr, logN, d_tw, false, false, false);
    }
    if (rc == 0) {
      k_rs_finish_cx<F><<<dim3((unsigned)((m - n + 255) / 256), (unsigned)nr), 256, 0, ctx->stream>>>(
          d_rows, R, x, (uint32_t)n, (uint32_t)m, (uint32_t)N, (const fpw<8>*)t->d_lead);
      ctx->launches++;
      cudaError_t ce = cudaGetLastError();
      if (ce != cudaSuccess) rc = fail(LF_ERR_CUDA, cudaGetErrorString(ce));
    }
  }
  return rc;
}

// field-generic front ends used by the ZK pipeline
template <class F>
static int launch_rs(lf_ctx* ctx, typename F::Elt* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                     size_t nbatch, size_t n, size_t m);
template <>
int launch_rs<FGf128>(lf_ctx* ctx, gf128* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                      size_t nbatch, size_t n, size_t m) {
  return launch_rs_gf(ctx, d_rows, row_stride, nrows, batch_stride, nbatch, n, m);
}
template <>
int launch_rs<FFp256>(lf_ctx* ctx, fpw<8>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                      size_t nbatch, size_t n, size_t m) {
  return launch_rs_p256(ctx, d_rows, row_stride, nrows, batch_stride, nbatch, n, m);
}

template <>
int launch_rs<FFpBn254>(lf_ctx* ctx, fpw<8>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                        size_t nbatch, size_t n, size_t m) {
  return rs_conv_run<FFpBn254>(ctx, bn254_host(), d_rows, row_stride, nrows, batch_stride, nbatch, n, m);
}
template <>
int launch_rs<FFp128>(lf_ctx* ctx, fpw<4>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                      size_t nbatch, size_t n, size_t m) {
  return rs_conv_run<FFp128>(ctx, fp128_host(), d_rows, row_stride, nrows, batch_stride, nbatch, n, m);
}
template <>
int launch_rs<FFpGold>(lf_ctx* ctx, fpw<2>* d_rows, size_t row_stride, size_t nrows, size_t batch_stride,
                       size_t nbatch, size_t n, size_t m) {
  return rs_conv_run<FFpGold>(ctx, gold_host(), d_rows, row_stride, nrows, batch_stride, nbatch, n, m);
}

template <class F>
static int launch_merkle(lf_ctx* ctx, const typename F::Elt* d_tab, size_t tab_batch_stride, uint32_t nrow,
                         uint32_t block_enc, uint32_t dblock, const uint8_t* d_nonces,
                         size_t nonce_batch_stride, uint32_t* d_nodes, size_t nodes_batch_stride,
                         size_t nbatch, const uint32_t* d_rej = nullptr, size_t rej_stride = 0) {
  uint32_t block_ext = block_enc - dblock;
  dim3 grid((block_ext + 127) / 128, (unsigned)nbatch);
  k_merkle_leaves<F><<<grid, 128, 0, ctx->stream>>>(d_tab, tab_batch_stride, nrow, block_enc, dblock,
                                                         block_ext, d_nonces, nonce_batch_stride, d_nodes,
                                                         nodes_batch_stride, d_rej, rej_stride);
  ctx->launches++;
  LF_CUDA(cudaGetLastError());
  k_merkle_tree<<<(unsigned)nbatch, 256, 0, ctx->stream>>>(d_nodes, nodes_batch_stride, block_ext);
  ctx->launches++;
  LF_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- micro kernels
template <class F>
__global__ void k_elt_mul(const typename F::Elt* a, const typename F::Elt* b, typename F::Elt* out,
                          size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = F::mul(a[i], b[i]);
}

__global__ void k_bench_imad(uint64_t* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
  uint64_t acc0 = a, acc1 = b, acc2 = a ^ b, acc3 = a + b;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc0 = (uint64_t)(uint32_t)acc0 * b + acc0;
      acc1 = (uint64_t)(uint32_t)acc1 * a + acc1;
      acc2 = (uint64_t)(uint32_t)acc2 * b + acc2;
      acc3 = (uint64_t)(uint32_t)acc3 * a + acc3;
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc0 ^ acc1 ^ acc2 ^ acc3;
}
__global__ void k_bench_lop3(uint32_t* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7, c = a ^ 0x5bd1e995u;
  uint32_t x0 = a, x1 = b, x2 = c, x3 = a + b;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = (x0 & x1) ^ (~x0 & x2) ^ b;
      x1 = (x1 & x2) ^ (~x1 & x3) ^ c;
      x2 = (x2 & x3) ^ (~x2 & x0) ^ a;
      x3 = (x3 & x0) ^ (~x3 & x1) ^ b;
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3;
}
__global__ void k_bench_gfmul(gf128* out, int iters) {
  gf128 a, b;
  a.w[0] = threadIdx.x * 2654435761u + 1; a.w[1] = blockIdx.x + 3; a.w[2] = 0x9e3779b9u; a.w[3] = threadIdx.x;
  b.w[0] = 0x85ebca6bu; b.w[1] = threadIdx.x ^ 0xc2b2ae35u; b.w[2] = blockIdx.x; b.w[3] = 0x27d4eb2fu;
  for (int i = 0; i < iters; ++i) {
    a = gf_mul_inl(a, b);
    b = gf_mul_inl(b, a);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = gf_add(a, b);
}
__global__ void k_bench_p256mul(fpw<8>* out, int iters) {
  typedef FFp256 F;
  fpw<8> a, b;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a.w[i] = threadIdx.x * 2654435761u + i;
    b.w[i] = blockIdx.x * 40503u + 7 * i + 1;
  }
  a.w[7] &= 0x7fffffffu;
  b.w[7] &= 0x7fffffffu;
  for (int i = 0; i < iters; ++i) {
    a = fp_mul_p256_dev(a, b, c_p256.m);
    b = fp_mul_p256_dev(b, a, c_p256.m);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = fp_add<8>(a, b, c_p256.m);
}
__global__ void k_bench_sha(uint32_t* out, int iters) {
  uint32_t h[8], w[16];
  sha256_iv(h);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = h[k & 7] + threadIdx.x + k;
    sha256_compress(h, w);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = h[0] ^ h[7];
}

}  // namespace lf

// ================================================================ C ABI
extern "C" {

const char* lf_last_error(void) { return g_err.c_str(); }
const char* lf_version(void) { return "longfellow_b200 0.1 (sm_100a)"; }

int lf_ctx_create(int device, void* stream, lf_ctx** out) {
  if (!out) return fail(LF_ERR_ARG, "lf_ctx_create: out is null");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LF_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                 " (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(LF_ERR_ARG, "lf_ctx_create: bad device ordinal");
  LF_CUDA(cudaSetDevice(device));
  std::unique_ptr<lf_ctx> c(new lf_ctx);
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    LF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  LF_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  const GfHost& h = gf_host();
  LF_CUDA(cudaMalloc(&c->d_tw, h.tw.size() * sizeof(gf128)));
  LF_CUDA(cudaMemcpy(c->d_tw, h.tw.data(), h.tw.size() * sizeof(gf128), cudaMemcpyHostToDevice));
  LF_CUDA(cudaMemcpyToSymbol(c_gf, &h.consts, sizeof(GfConsts)));
  LF_CUDA(cudaMemcpyToSymbol(c_p256, &p256_host().C, sizeof(FpConsts<8>)));
  LF_CUDA(cudaMemcpyToSymbol(c_bn254, &bn254_host().C, sizeof(FpConsts<8>)));
  LF_CUDA(cudaMemcpyToSymbol(c_fp128, &fp128_host().C, sizeof(FpConsts<4>)));
  LF_CUDA(cudaMemcpyToSymbol(c_gold, &gold_host().C, sizeof(FpConsts<2>)));
  *out = c.release();
  return 0;
}

void lf_ctx_destroy(lf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->rs_plans) cudaFree(kv.second.d_steps);
  for (auto& kv : ctx->fft_tw) cudaFree(kv.second);
  for (auto& kv : ctx->rs_conv) {
    auto* t = (RsConvTables*)kv.second;
    cudaFree(t->d_yh);
    cudaFree(t->d_lead);
    cudaFree(t->d_binom);
    delete t;
  }
  for (auto& kv : ctx->rs_fp) {
    cudaFree(kv.second.d_inv);
    cudaFree(kv.second.d_lead);
    cudaFree(kv.second.d_binom);
    cudaFree(kv.second.d_wk);
    cudaFree(kv.second.d_yh);
  }
  cudaFree(ctx->d_tw);
  cudaFree(ctx->work[0]);
  cudaFree(ctx->work[1]);
  cudaFree(ctx->d_crt_consts);
  for (auto& kv : ctx->crt_tw) {
    cudaFree(kv.second.first);
    cudaFree(kv.second.second);
  }
  for (auto& kv : ctx->crt_spec) cudaFree(kv.second);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int lf_ctx_synchronize(lf_ctx* ctx) {
  if (!ctx) return fail(LF_ERR_ARG, "null ctx");
  LF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

uint64_t lf_ctx_launch_count(const lf_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

namespace {
// a device allocation that is freed on every way out of a function (LF_CUDA returns early on errors)
struct DevBuf {
  void* p = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
  void* release() {
    void* q = p;
    p = nullptr;
    return q;
  }
};

// host wire buffer -> device elements (Montgomery for prime fields)
template <class F>
int upload_elts(lf_ctx* ctx, const void* host, size_t n, typename F::Elt** d_out) {
  DevBuf d, raw, bad;
  LF_CUDA(d.alloc(std::max<size_t>(n, 1) * sizeof(typename F::Elt)));
  if (F::kChar2) {
    LF_CUDA(cudaMemcpyAsync(d.p, host, n * F::kBytes, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    LF_CUDA(raw.alloc(std::max<size_t>(n, 1) * F::kBytes));
    LF_CUDA(bad.alloc(sizeof(int)));
    LF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
    LF_CUDA(cudaMemcpyAsync(raw.p, host, n * F::kBytes, cudaMemcpyHostToDevice, ctx->stream));
    k_from_wire<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(raw.as<uint8_t>(), d.as<typename F::Elt>(), n,
                                                                      bad.as<int>());
    ctx->launches++;
    int hbad = 0;
    LF_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hbad) return fail(LF_ERR_FORMAT, "element is not canonical (>= modulus)");
  }
  *d_out = static_cast<typename F::Elt*>(d.release());
  return 0;
}
template <class F>
int download_elts(lf_ctx* ctx, const typename F::Elt* d, size_t n, void* host) {
  if (F::kChar2) {
    LF_CUDA(cudaMemcpyAsync(host, d, n * F::kBytes, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    DevBuf raw;
    LF_CUDA(raw.alloc(std::max<size_t>(n, 1) * F::kBytes));
    k_to_wire<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, raw.as<uint32_t>(), n);
    ctx->launches++;
    LF_CUDA(cudaMemcpyAsync(host, raw.p, n * F::kBytes, cudaMemcpyDeviceToHost, ctx->stream));
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  LF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

template <class F>
int elt_mul_t(lf_ctx* ctx, const void* a, const void* b, void* out, size_t n) {
  typename F::Elt *da = nullptr, *db = nullptr, *dc = nullptr;
  int rc = upload_elts<F>(ctx, a, n, &da);
  if (!rc) rc = upload_elts<F>(ctx, b, n, &db);
  if (!rc) {
    cudaError_t e = cudaMalloc(&dc, n * sizeof(typename F::Elt));
    if (e != cudaSuccess) rc = fail(LF_ERR_CUDA, cudaGetErrorString(e));
  }
  if (!rc) {
    k_elt_mul<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(da, db, dc, n);
    ctx->launches++;
    rc = download_elts<F>(ctx, dc, n, out);
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dc);
  return rc;
}

template <class F>
int rs_interpolate_t(lf_ctx* ctx, size_t n, size_t m, void* rows, size_t nrows) {
  typename F::Elt* d = nullptr;
  int rc = upload_elts<F>(ctx, rows, nrows * m, &d);
  if (!rc) rc = launch_rs<F>(ctx, d, m, nrows, 0, 1, n, m);
  if (!rc) rc = download_elts<F>(ctx, d, nrows * m, rows);
  cudaFree(d);
  return rc;
}

template <class F>
int merkle_commit_t(lf_ctx* ctx, size_t nrow, size_t block_enc, size_t dblock, const void* tableau,
                    const uint8_t* nonces, uint8_t root_out[32], uint8_t* nodes_out) {
  size_t block_ext = block_enc - dblock;
  typename F::Elt* d_tab = nullptr;
  int rc = upload_elts<F>(ctx, tableau, nrow * block_enc, &d_tab);
  if (rc) return rc;
  DevBuf tab_owner, nonce_buf, nodes_buf;
  tab_owner.p = d_tab;
  LF_CUDA(nonce_buf.alloc(block_ext * 32));
  LF_CUDA(nodes_buf.alloc(2 * block_ext * 32));
  uint8_t* d_nonce = nonce_buf.as<uint8_t>();
  uint32_t* d_nodes = nodes_buf.as<uint32_t>();
  LF_CUDA(cudaMemsetAsync(d_nodes, 0, 2 * block_ext * 32, ctx->stream));
  LF_CUDA(cudaMemcpyAsync(d_nonce, nonces, block_ext * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_merkle<F>(ctx, d_tab, 0, (uint32_t)nrow, (uint32_t)block_enc, (uint32_t)dblock, d_nonce, 0, d_nodes,
                        0, 1);
  if (rc == 0) {
    std::vector<uint32_t> nodes(2 * block_ext * 8);
    LF_CUDA(cudaMemcpyAsync(nodes.data(), d_nodes, nodes.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
    // digests are kept as big-endian words on the device
    auto put = [&](uint8_t* dst, size_t node) {
      for (int k = 0; k < 8; ++k) {
        uint32_t x = nodes[8 * node + k];
        dst[4 * k] = (uint8_t)(x >> 24);
        dst[4 * k + 1] = (uint8_t)(x >> 16);
        dst[4 * k + 2] = (uint8_t)(x >> 8);
        dst[4 * k + 3] = (uint8_t)x;
      }
    };
    put(root_out, 1);
    if (nodes_out)
      for (size_t i = 0; i < 2 * block_ext; ++i) put(nodes_out + 32 * i, i);
  }
  return rc;
}
}  // namespace

extern "C" {

#define LF_DISPATCH_FIELD(field_id, CALL)                                      \
  switch (field_id) {                                                          \
    case LF_FIELD_GF2_128: { typedef FGf128 F; return CALL; }                  \
    case LF_FIELD_P256: { typedef FFp256 F; return CALL; }                     \
    case LF_FIELD_BN254: { typedef FFpBn254 F; return CALL; }                  \
    case LF_FIELD_FP128: { typedef FFp128 F; return CALL; }                    \
    case LF_FIELD_GOLDILOCKS: { typedef FFpGold F; return CALL; }              \
    default: return fail(LF_ERR_UNSUPPORTED, "field not built yet");           \
  }
#define LF_DISPATCH_ZK_FIELD(field_id, CALL)                                   \
  switch (field_id) {                                                          \
    case LF_FIELD_GF2_128: { typedef FGf128 F; return CALL; }                  \
    case LF_FIELD_P256: { typedef FFp256 F; return CALL; }                     \
    default: return fail(LF_ERR_UNSUPPORTED, "field not built yet");           \
  }

int lf_elt_mul(lf_ctx* ctx, int field_id, const void* a, const void* b, void* out, size_t n) {
  if (!ctx || !a || !b || !out) return fail(LF_ERR_ARG, "lf_elt_mul: null argument");
  if (n == 0) return 0;
  LF_CUDA(cudaSetDevice(ctx->device));
  LF_DISPATCH_FIELD(field_id, elt_mul_t<F>(ctx, a, b, out, n));
}

int lf_rs_interpolate_dev(lf_ctx* ctx, int field_id, size_t n, size_t m, void* d_rows, size_t row_stride,
                          size_t nrows) {
  if (!ctx || !d_rows) return fail(LF_ERR_ARG, "lf_rs_interpolate_dev: null argument");
  if (row_stride < m) return fail(LF_ERR_ARG, "rs: row_stride < m");
  LF_CUDA(cudaSetDevice(ctx->device));
  // device rows are in the in-memory (Montgomery) representation
  LF_DISPATCH_FIELD(field_id, launch_rs<F>(ctx, (typename F::Elt*)d_rows, row_stride, nrows, 0, 1, n, m));
}

int lf_rs_interpolate(lf_ctx* ctx, int field_id, size_t n, size_t m, void* rows, size_t nrows) {
  if (!ctx || !rows) return fail(LF_ERR_ARG, "lf_rs_interpolate: null argument");
  if (nrows == 0) return 0;
  if (n == 0 || m < n) return fail(LF_ERR_ARG, "rs: need 0 < n <= m");
  LF_CUDA(cudaSetDevice(ctx->device));
  LF_DISPATCH_FIELD(field_id, rs_interpolate_t<F>(ctx, n, m, rows, nrows));
}

int lf_merkle_commit(lf_ctx* ctx, int field_id, size_t nrow, size_t block_enc, size_t dblock,
                     const void* tableau, const uint8_t* nonces, uint8_t root_out[32], uint8_t* nodes_out) {
  if (!ctx || !tableau || !nonces || !root_out) return fail(LF_ERR_ARG, "lf_merkle_commit: null argument");
  if (block_enc <= dblock || nrow == 0) return fail(LF_ERR_ARG, "merkle: need block_enc > dblock, nrow > 0");
  LF_CUDA(cudaSetDevice(ctx->device));
  LF_DISPATCH_ZK_FIELD(field_id, merkle_commit_t<F>(ctx, nrow, block_enc, dblock, tableau, nonces, root_out, nodes_out));
}

}  // extern "C"

namespace {
template <class F, class H>
int fft_scalar_t(lf_ctx* ctx, const H& Hf, void* elts, size_t n, uint32_t logn, int forward, int reps, double* ms,
                 size_t nrows = 1) {
  typedef typename F::Elt Elt;
  Elt* d = nullptr;
  Elt* d_tw;
  int rc = fft_twiddles<H, F::W>(ctx, F::kFieldId, Hf, logn, &d_tw);
  if (rc) return rc;
  if (elts) {
    rc = upload_elts<F>(ctx, elts, n, &d);
  } else {
    LF_CUDA(cudaMalloc(&d, nrows * n * sizeof(Elt)));
    LF_CUDA(cudaMemsetAsync(d, 0x11, nrows * n * sizeof(Elt), ctx->stream));  // timing only: any limbs do
  }
  if (rc) return rc;
  cudaEvent_t e0, e1;
  LF_CUDA(cudaEventCreate(&e0));
  LF_CUDA(cudaEventCreate(&e1));
  LF_CUDA(cudaEventRecord(e0, ctx->stream));
  for (int r = 0; r < reps && !rc; ++r) rc = fft_run<AlgF<F>>(ctx, d, n, nrows, logn, d_tw, forward != 0, false, true);
  LF_CUDA(cudaEventRecord(e1, ctx->stream));
  LF_CUDA(cudaEventSynchronize(e1));
  if (ms) {
    float t;
    LF_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = t / reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!rc && elts) rc = download_elts<F>(ctx, d, n, elts);
  cudaFree(d);
  return rc;
}
int fft_p256_t(lf_ctx* ctx, void* elts, size_t n, uint32_t logn, int forward, int reps, double* ms, size_t nrows = 1) {
  typedef FFp256 F;
  fpw<8>* d = nullptr;
  Cx<F>* d_tw;
  int rc = fft_twiddles_p256(ctx, logn, &d_tw);
  if (rc) return rc;
  if (elts) {
    rc = upload_elts<F>(ctx, elts, 2 * n, &d);  // (re, im) pairs
  } else {
    LF_CUDA(cudaMalloc(&d, nrows * 2 * n * sizeof(fpw<8>)));
    LF_CUDA(cudaMemsetAsync(d, 0x5a, nrows * 2 * n * sizeof(fpw<8>), ctx->stream));
  }
  if (rc) return rc;
  cudaEvent_t e0, e1;
  LF_CUDA(cudaEventCreate(&e0));
  LF_CUDA(cudaEventCreate(&e1));
  LF_CUDA(cudaEventRecord(e0, ctx->stream));
  for (int r = 0; r < reps && !rc; ++r)
    rc = fft_run<AlgCx<F>>(ctx, (Cx<F>*)d, n, nrows, logn, d_tw, forward != 0, false, true);
  LF_CUDA(cudaEventRecord(e1, ctx->stream));
  LF_CUDA(cudaEventSynchronize(e1));
  if (ms) {
    float t;
    LF_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = t / reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!rc && elts) rc = download_elts<F>(ctx, d, 2 * n, elts);
  cudaFree(d);
  return rc;
}
// GF(2^128): the additive FFT of LCH14 on the span of beta_0..beta_{l-1}, coset 0
// (lch14.h:106-146): forward = evaluate the novel-basis coefficients, else IFFT
int fft_gf_t(lf_ctx* ctx, void* elts, size_t n, uint32_t logn, int forward, int reps, double* ms, size_t nrows = 1) {
  if (logn > 16) return fail(LF_ERR_ARG, "fft: the LCH14 subspace has dimension 16 (lch14.h:45-47), n <= 65536");
  gf128* d = nullptr;
  LF_CUDA(cudaMalloc(&d, nrows * n * sizeof(gf128)));
  if (elts) {
    LF_CUDA(cudaMemcpyAsync(d, elts, n * sizeof(gf128), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    LF_CUDA(cudaMemsetAsync(d, 0x5a, nrows * n * sizeof(gf128), ctx->stream));
  }
  cudaEvent_t e0, e1;
  LF_CUDA(cudaEventCreate(&e0));
  LF_CUDA(cudaEventCreate(&e1));
  LF_CUDA(cudaEventRecord(e0, ctx->stream));
  int rc = 0;
  for (int r = 0; r < reps && !rc; ++r) rc = launch_lch14_fft(ctx, d, logn, 0, forward != 0, nrows);
  LF_CUDA(cudaEventRecord(e1, ctx->stream));
  LF_CUDA(cudaEventSynchronize(e1));
  if (ms) {
    float t;
    LF_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = t / reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!rc && elts) {
    LF_CUDA(cudaMemcpyAsync(elts, d, n * sizeof(gf128), cudaMemcpyDeviceToHost, ctx->stream));
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(d);
  return rc;
}
int fft_dispatch(lf_ctx* ctx, int field_id, void* elts, size_t n, int forward, int reps, double* ms, size_t nrows = 1) {
  if (n == 0 || (n & (n - 1))) return fail(LF_ERR_ARG, "fft: n must be a power of two");
  uint32_t logn = 0;
  while (((size_t)1 << logn) < n) ++logn;
  if (logn > 26) return fail(LF_ERR_ARG, "fft: n too large");
  switch (field_id) {
    case LF_FIELD_BN254: return fft_scalar_t<FFpBn254>(ctx, bn254_host(), elts, n, logn, forward, reps, ms, nrows);
    case LF_FIELD_FP128: return fft_scalar_t<FFp128>(ctx, fp128_host(), elts, n, logn, forward, reps, ms, nrows);
    case LF_FIELD_GOLDILOCKS: return fft_scalar_t<FFpGold>(ctx, gold_host(), elts, n, logn, forward, reps, ms, nrows);
    case LF_FIELD_P256: return fft_p256_t(ctx, elts, n, logn, forward, reps, ms, nrows);
    case LF_FIELD_GF2_128: return fft_gf_t(ctx, elts, n, logn, forward, reps, ms, nrows);
    default: return fail(LF_ERR_UNSUPPORTED, "fft: unknown field");
  }
}
}  // namespace

extern "C" {

int lf_fft(lf_ctx* ctx, int field_id, void* elts, size_t n, int forward) {
  if (!ctx || !elts) return fail(LF_ERR_ARG, "lf_fft: null argument");
  LF_CUDA(cudaSetDevice(ctx->device));
  return fft_dispatch(ctx, field_id, elts, n, forward, 1, nullptr);
}

int lf_fft_time(lf_ctx* ctx, int field_id, size_t n, int reps, double* ms_per_fft) {
  if (!ctx || !ms_per_fft || reps <= 0) return fail(LF_ERR_ARG, "lf_fft_time: bad argument");
  LF_CUDA(cudaSetDevice(ctx->device));
  int rc = fft_dispatch(ctx, field_id, nullptr, n, 0, 1, nullptr);  // warm-up (tables, attributes)
  if (rc) return rc;
  return fft_dispatch(ctx, field_id, nullptr, n, 0, reps, ms_per_fft);
}

int lf_fft_time_rows(lf_ctx* ctx, int field_id, size_t n, size_t nrows, int reps, double* ms_per_call) {
  if (!ctx || !ms_per_call || reps <= 0 || nrows == 0) return fail(LF_ERR_ARG, "lf_fft_time_rows: bad argument");
  LF_CUDA(cudaSetDevice(ctx->device));
  int rc = fft_dispatch(ctx, field_id, nullptr, n, 0, 1, nullptr, nrows);
  if (rc) return rc;
  return fft_dispatch(ctx, field_id, nullptr, n, 0, reps, ms_per_call, nrows);
}

}  // extern "C"
namespace {
template <class F>
int rs_time_t(lf_ctx* ctx, size_t n, size_t m, size_t nrows, int reps, double* ms) {
  typedef typename F::Elt Elt;
  Elt* d = nullptr;
  LF_CUDA(cudaMalloc(&d, nrows * m * sizeof(Elt)));
  LF_CUDA(cudaMemsetAsync(d, 0x11, nrows * m * sizeof(Elt), ctx->stream));  // timing only: any limbs do
  int rc = launch_rs<F>(ctx, d, m, nrows, 0, 1, n, m);                       // warm-up (tables, attributes)
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, ctx->stream);
  for (int r = 0; r < reps && !rc; ++r) rc = launch_rs<F>(ctx, d, m, nrows, 0, 1, n, m);
  cudaEventRecord(e1, ctx->stream);
  cudaError_t ce = cudaEventSynchronize(e1);
  if (!rc && ce != cudaSuccess) rc = fail(LF_ERR_CUDA, cudaGetErrorString(ce));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  *ms = t / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return rc;
}
}  // namespace
extern "C" {

int lf_rs_time(lf_ctx* ctx, int field_id, size_t n, size_t m, size_t nrows, int reps, double* ms_per_call) {
  if (!ctx || !ms_per_call || reps <= 0 || nrows == 0 || n == 0 || m < n)
    return fail(LF_ERR_ARG, "lf_rs_time: bad argument");
  LF_CUDA(cudaSetDevice(ctx->device));
  LF_DISPATCH_FIELD(field_id, rs_time_t<F>(ctx, n, m, nrows, reps, ms_per_call));
}

}  // extern "C"
// single-thread latency probes of the transcript primitives (cycles per call);
// what = 100 + k:  0 compression  1 snapshot  2 AES-256 key schedule  3 AES block
//                  4 write of one 16-byte element  5 write + 16 challenge bytes
__global__ void k_probe_serial(long long* out) {
  __shared__ lf::Transcript ts;
  __shared__ lf::AesTables aes;
  lf::aes_stage_tables(&aes);
  const uint8_t* sbox = aes.sbox;
  if (threadIdx.x != 0) return;
  const uint8_t seed[4] = {1, 2, 3, 4};
  ts.init(seed, 4);
  ts.use_tables(&aes);
  uint32_t acc = 0;
  const int R = 16;
  // the very first write + challenge of the launch: instruction caches cold, as in a round of the prover
  long long c0 = clock64();
  {
    uint32_t e[4] = {acc, 7, 8, 9}, o[4];
    ts.write_elt_words(e, 4);
    ts.words(o, 4);
    acc += o[1];
  }
  long long c1 = clock64();
  out[7] = c1 - c0;
  long long t0 = clock64();
  for (int i = 0; i < R; ++i) {
    ts.sha.buf[3] = i + acc;
    ts.sha.compress_block();
  }
  long long t1 = clock64();
  for (int i = 0; i < R; ++i) {
    uint32_t d[8];
    ts.sha.buf[1] = i + acc;
    ts.sha.snapshot(d);
    acc += d[0];
  }
  long long t2 = clock64();
  for (int i = 0; i < R; ++i) {
    uint32_t key[8] = {acc, 1, 2, 3, 4, 5, 6, (uint32_t)i};
    ts.prf.init_unrolled(key, sbox);
    acc += ts.prf.rk[59];
  }
  long long t3 = clock64();
  for (int i = 0; i < R; ++i) {
    uint32_t in[4] = {acc, 0, 0, (uint32_t)i}, o[4];
    ts.prf.encrypt_te(in, o, &aes);
    acc += o[0];
  }
  long long t4 = clock64();
  for (int i = 0; i < R; ++i) {
    uint32_t e[4] = {acc, 7, 8, (uint32_t)i};
    ts.write_elt_words(e, 4);
  }
  long long t5 = clock64();
  for (int i = 0; i < R; ++i) {
    uint32_t e[4] = {acc, 7, 8, (uint32_t)i}, o[4];
    ts.write_elt_words(e, 4);
    ts.words(o, 4);
    acc += o[1];
  }
  long long t6 = clock64();
  out[0] = (t1 - t0) / R;
  out[1] = (t2 - t1) / R;
  out[2] = (t3 - t2) / R;
  out[3] = (t4 - t3) / R;
  out[4] = (t5 - t4) / R;
  out[5] = (t6 - t5) / R;
  out[6] = acc;
}
extern "C" {

int lf_microbench(lf_ctx* ctx, int what, double* gops) {
  if (!ctx || !gops) return fail(LF_ERR_ARG, "lf_microbench: null argument");
  LF_CUDA(cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, threads = 256;
  void* d;
  LF_CUDA(cudaMalloc(&d, (size_t)blocks * threads * 32));
  if (what >= 100 && what < 108) {
    long long h[8];
    k_probe_serial<<<1, 32, 0, ctx->stream>>>((long long*)d);
    ctx->launches++;
    LF_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    LF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    *gops = (double)h[what - 100];
    return 0;
  }
  cudaEvent_t e0, e1;
  LF_CUDA(cudaEventCreate(&e0));
  LF_CUDA(cudaEventCreate(&e1));
  int iters = what == 0 ? 4096 : what == 1 ? 4096 : 256;
  double ops_per_thread = what == 0 ? 32.0 * iters : what == 1 ? 32.0 * 3 * iters
                          : (what == 2 || what == 4) ? 2.0 * iters : 1.0 * iters;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    LF_CUDA(cudaEventRecord(e0, ctx->stream));
    if (what == 0) k_bench_imad<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, iters);
    else if (what == 1) k_bench_lop3<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)d, iters);
    else if (what == 2) k_bench_gfmul<<<blocks, threads, 0, ctx->stream>>>((gf128*)d, iters);
    else if (what == 3) k_bench_sha<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)d, iters);
    else if (what == 4) k_bench_p256mul<<<blocks, threads, 0, ctx->stream>>>((fpw<8>*)d, iters);
    else return fail(LF_ERR_ARG, "lf_microbench: unknown benchmark");
    ctx->launches++;
    LF_CUDA(cudaEventRecord(e1, ctx->stream));
    LF_CUDA(cudaEventSynchronize(e1));
    float ms;
    LF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  *gops = ops_per_thread * blocks * threads / (best * 1e-3) / 1e9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}

}  // extern "C"
#include "zk_host.inc"
