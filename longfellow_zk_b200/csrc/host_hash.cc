// Host-side SHA-256 block function of the back end (plain C++, compiled by g++).
//
// The prover's Fiat-Shamir transcript is a SHA-256 chain (lib/random/transcript.h:33-190,
// lib/util/crypto.h:40-69).  Almost all of it runs on the device next to the data it binds,
// but two pieces are long chains over bytes the HOST already knows and that a single GPU
// thread hashes ~10x slower than one x86 core with the SHA extensions:
//   * Transcript::write0(nterms) of initialize_sumcheck_fiat_shamir (lib/zk/zk_common.h:177-179):
//     nterms zero bytes, 2 425 compressions for the SHA-256 circuit, 121 k for the mdoc hash circuit;
//   * circuit_id (lib/sumcheck/circuit_id.h:30-67): a digest over every quad term.
// lf_host_sha256_blocks / _zero_blocks advance a chaining value over whole 64-byte blocks, with the
// SHA extensions when the CPU has them (checked once with CPUID) and portable C otherwise.
#include <cpuid.h>
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace {

const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void blocks_portable(uint32_t h[8], const uint8_t* data, size_t nblocks) {
  static const uint8_t zero[64] = {0};
  for (size_t blk = 0; blk < nblocks; ++blk) {
    const uint8_t* p = data ? data + 64 * blk : zero;
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
      w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
      uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
}

// SHA-NI: two rounds per sha256rnds2, state kept as (ABEF, CDGH)
__attribute__((target("sha,sse4.1,ssse3"))) void blocks_shani(uint32_t h[8], const uint8_t* data, size_t nblocks) {
  const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
  __m128i tmp = _mm_loadu_si128((const __m128i*)&h[0]);     // DCBA
  __m128i st1 = _mm_loadu_si128((const __m128i*)&h[4]);     // HGFE
  tmp = _mm_shuffle_epi32(tmp, 0xB1);                        // CDAB
  st1 = _mm_shuffle_epi32(st1, 0x1B);                        // EFGH
  __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                // ABEF
  st1 = _mm_blend_epi16(st1, tmp, 0xF0);                     // CDGH
  const __m128i* KV = (const __m128i*)K;
  for (size_t blk = 0; blk < nblocks; ++blk) {
    const __m128i save0 = st0, save1 = st1;
    __m128i m0, m1, m2, m3, msg;
    if (data) {
      const __m128i* p = (const __m128i*)(data + 64 * blk);
      m0 = _mm_shuffle_epi8(_mm_loadu_si128(p + 0), bswap);
      m1 = _mm_shuffle_epi8(_mm_loadu_si128(p + 1), bswap);
      m2 = _mm_shuffle_epi8(_mm_loadu_si128(p + 2), bswap);
      m3 = _mm_shuffle_epi8(_mm_loadu_si128(p + 3), bswap);
    } else {
      m0 = m1 = m2 = m3 = _mm_setzero_si128();
    }
#define LF_RND4(M, i)                                   \
  msg = _mm_add_epi32(M, _mm_loadu_si128(KV + (i)));    \
  st1 = _mm_sha256rnds2_epu32(st1, st0, msg);           \
  msg = _mm_shuffle_epi32(msg, 0x0E);                   \
  st0 = _mm_sha256rnds2_epu32(st0, st1, msg);
#define LF_SCHED(Ma, Mb, Mc, Md)                                            \
  Ma = _mm_sha256msg1_epu32(Ma, Mb);                                        \
  Ma = _mm_add_epi32(Ma, _mm_alignr_epi8(Md, Mc, 4));                       \
  Ma = _mm_sha256msg2_epu32(Ma, Md);
    LF_RND4(m0, 0) LF_RND4(m1, 1) LF_RND4(m2, 2) LF_RND4(m3, 3)
    for (int i = 4; i < 16; i += 4) {
      LF_SCHED(m0, m1, m2, m3) LF_RND4(m0, i)
      LF_SCHED(m1, m2, m3, m0) LF_RND4(m1, i + 1)
      LF_SCHED(m2, m3, m0, m1) LF_RND4(m2, i + 2)
      LF_SCHED(m3, m0, m1, m2) LF_RND4(m3, i + 3)
    }
#undef LF_RND4
#undef LF_SCHED
    st0 = _mm_add_epi32(st0, save0);
    st1 = _mm_add_epi32(st1, save1);
  }
  tmp = _mm_shuffle_epi32(st0, 0x1B);                        // FEBA
  st1 = _mm_shuffle_epi32(st1, 0xB1);                        // DCHG
  st0 = _mm_blend_epi16(tmp, st1, 0xF0);                     // DCBA
  st1 = _mm_alignr_epi8(st1, tmp, 8);                        // HGFE
  _mm_storeu_si128((__m128i*)&h[0], st0);
  _mm_storeu_si128((__m128i*)&h[4], st1);
}

bool cpu_has_shani() {
  static const bool has = [] {
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
    const bool sha = (b >> 29) & 1;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    const bool sse41 = (c >> 19) & 1, ssse3 = (c >> 9) & 1;
    return sha && sse41 && ssse3;
  }();
  return has;
}

}  // namespace

extern "C" {

// h: SHA-256 chaining value (8 host-order words); data: nblocks * 64 message bytes
void lf_host_sha256_blocks(uint32_t h[8], const uint8_t* data, size_t nblocks) {
  if (nblocks == 0) return;
  if (cpu_has_shani()) blocks_shani(h, data, nblocks);
  else blocks_portable(h, data, nblocks);
}
// the same over nblocks all-zero blocks
void lf_host_sha256_zero_blocks(uint32_t h[8], size_t nblocks) {
  if (nblocks == 0) return;
  if (cpu_has_shani()) blocks_shani(h, nullptr, nblocks);
  else blocks_portable(h, nullptr, nblocks);
}
int lf_host_sha256_accelerated(void) { return cpu_has_shani() ? 1 : 0; }
// for tests: force the portable path
void lf_host_sha256_blocks_portable(uint32_t h[8], const uint8_t* data, size_t nblocks) {
  blocks_portable(h, data, nblocks);
}

}  // extern "C"
