// Host-side unit check of the header-only device primitives (gf128.cuh,
// hash.cuh compiled as plain C++).  Reads test vectors on stdin, writes results
// on stdout; driven by tests/test_device_primitives_host.py which compares the
// output with the oracle.  This exercises the SAME source the kernels compile,
// here on the CPU where there is no GPU to debug on.
#define __host__
#define __device__
#define __forceinline__ inline
#define LF_HD
#include <cstdio>
#include <cstring>
#include <vector>
#include <type_traits>

#include "../../longfellow_zk_b200/csrc/gf128.cuh"
#include "../../longfellow_zk_b200/csrc/hash.cuh"
#define LF_HDI inline
#include "../../longfellow_zk_b200/csrc/fp.cuh"

using namespace lf;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::vector<uint8_t> in;
  uint8_t tmp[4096];
  size_t k;
  while ((k = fread(tmp, 1, sizeof(tmp), stdin)) > 0) in.insert(in.end(), tmp, tmp + k);
  if (!strcmp(argv[1], "gfmul")) {  // pairs of 16-byte elements -> products
    for (size_t i = 0; i + 32 <= in.size(); i += 32) {
      gf128 a, b;
      memcpy(a.w, &in[i], 16);
      memcpy(b.w, &in[i + 16], 16);
      gf128 c = gf_mul(a, b);
      fwrite(c.w, 1, 16, stdout);
    }
  } else if (!strcmp(argv[1], "gfinv")) {
    for (size_t i = 0; i + 16 <= in.size(); i += 16) {
      gf128 a;
      memcpy(a.w, &in[i], 16);
      gf128 c = gf_inv(a);
      fwrite(c.w, 1, 16, stdout);
    }
  } else if (!strncmp(argv[1], "fp", 2)) {
    // fpmul8 / fpmul8p256 / fpmul4 / fpmul2 : stdin = modulus (W words) || pairs of
    // wire elements; stdout = wire products (through to/from Montgomery)
    auto run = [&](auto tag, bool p256) {
      constexpr int W = decltype(tag)::value;
      FpConsts<W> C;
      uint32_t mod[W];
      memcpy(mod, in.data(), 4 * W);
      fp_build_consts<W>(mod, &C);
      fpw<W> rsq, one_int;
      for (int i = 0; i < W; ++i) { rsq.w[i] = C.rsq[i]; one_int.w[i] = 0; }
      one_int.w[0] = 1;
      auto mul = [&](const fpw<W>& x, const fpw<W>& y) {
        if constexpr (W == 8) { if (p256) return fp_mul_p256(x, y, C.m); }
        return fp_mul_generic<W>(x, y, C.m, C.mprime);
      };
      for (size_t i = 4 * W; i + 8 * W <= in.size(); i += 8 * W) {
        fpw<W> a, b;
        memcpy(a.w, &in[i], 4 * W);
        memcpy(b.w, &in[i + 4 * W], 4 * W);
        fpw<W> am = mul(a, rsq), bm = mul(b, rsq);
        fpw<W> c = mul(mul(am, bm), one_int);
        fpw<W> s = fp_add<W>(am, bm, C.m), d = fp_sub<W>(am, bm, C.m);
        fpw<W> sw = mul(s, one_int), dw = mul(d, one_int);
        fwrite(c.w, 1, 4 * W, stdout);
        fwrite(sw.w, 1, 4 * W, stdout);
        fwrite(dw.w, 1, 4 * W, stdout);
      }
    };
    if (!strcmp(argv[1], "fpmul8")) run(std::integral_constant<int, 8>{}, false);
    else if (!strcmp(argv[1], "fpmul8p256")) run(std::integral_constant<int, 8>{}, true);
    else if (!strcmp(argv[1], "fpmul4")) run(std::integral_constant<int, 4>{}, false);
    else if (!strcmp(argv[1], "fpmul2")) run(std::integral_constant<int, 2>{}, false);
    else return 2;
  } else if (!strcmp(argv[1], "sha")) {  // digest of stdin
    Sha256 s;
    s.init();
    s.update(in.data(), (uint32_t)in.size());
    uint32_t d[8];
    s.snapshot(d);
    for (int i = 0; i < 8; ++i) {
      uint32_t x = bswap32(d[i]);
      fwrite(&x, 1, 4, stdout);
    }
  } else if (!strcmp(argv[1], "aes")) {  // key(32) || blocks
    Aes256 a;
    uint32_t key[8];
    memcpy(key, in.data(), 32);
    a.init(key, aes_default_sbox());
    for (size_t i = 32; i + 16 <= in.size(); i += 16) {
      uint32_t x[4], y[4];
      memcpy(x, &in[i], 16);
      a.encrypt(x, y, aes_default_sbox());
      fwrite(y, 1, 16, stdout);
    }
  } else if (!strcmp(argv[1], "transcript8")) {
    // the same script language with 32-byte elements (P-256 wire bytes): the 8-word element writes
    Transcript t;
    t.init((const uint8_t*)"test", 4);
    size_t p = 0;
    auto rd32 = [&]() { uint32_t v; memcpy(&v, &in[p], 4); p += 4; return v; };
    while (p < in.size()) {
      char op = (char)in[p++];
      if (op == 'B') { uint32_t n = rd32(); t.write_bytes(&in[p], n); p += n; }
      else if (op == 'Z') { t.write0(rd32()); }
      else if (op == 'E') { uint32_t w[8]; memcpy(w, &in[p], 32); p += 32; t.write_elt_words(w, 8); }
      else if (op == 'A') { uint32_t n = rd32(); t.begin_array(n);
        for (uint32_t i = 0; i < n; ++i) { uint32_t w[8]; memcpy(w, &in[p], 32); p += 32; t.elt_words(w, 8); } }
      else if (op == 'R') { uint32_t n = rd32(); std::vector<uint8_t> o(n); t.bytes(o.data(), n); fwrite(o.data(), 1, n, stdout); }
      else return 3;
    }
  } else if (!strcmp(argv[1], "transcript")) {
    // same script language as oracle (GF(2^128) elements); init = "test"
    Transcript t;
    t.init((const uint8_t*)"test", 4);
    size_t p = 0;
    auto rd32 = [&]() { uint32_t v; memcpy(&v, &in[p], 4); p += 4; return v; };
    while (p < in.size()) {
      char op = (char)in[p++];
      if (op == 'B') { uint32_t n = rd32(); t.write_bytes(&in[p], n); p += n; }
      else if (op == 'Z') { t.write0(rd32()); }
      else if (op == 'E') { uint32_t w[4]; memcpy(w, &in[p], 16); p += 16; t.write_elt_words(w, 4); }
      else if (op == 'A') { uint32_t n = rd32(); t.begin_array(n);
        for (uint32_t i = 0; i < n; ++i) { uint32_t w[4]; memcpy(w, &in[p], 16); p += 16; t.elt_words(w, 4); } }
      else if (op == 'R') { uint32_t n = rd32(); std::vector<uint8_t> o(n); t.bytes(o.data(), n); fwrite(o.data(), 1, n, stdout); }
      else if (op == 'N') { uint32_t r = t.nat(rd32()); fwrite(&r, 1, 4, stdout); }
      else if (op == 'G') { uint32_t n = rd32(); for (uint32_t i = 0; i < n; ++i) { uint32_t w[4]; t.words(w, 4); fwrite(w, 1, 16, stdout); } }
      else return 3;
    }
  } else {
    return 2;
  }
  return 0;
}
