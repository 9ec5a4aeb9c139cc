// SHA-256, AES-256 and the Fiat-Shamir transcript, device side.
//
// Mirrors the reference's Transcript / FSPRF (lib/random/transcript.h:33-190)
// which wraps OpenSSL SHA-256 and AES-256-ECB (lib/util/crypto.h:41-103):
//   * every write is tagged (1 byte) and, for byte strings / arrays, prefixed by
//     an 8-byte little-endian length (transcript.h:116-153,160-171);
//   * challenge bytes are AES-256-ECB(key = SHA-256 of everything written so
//     far, block i = LE64(i) || 0^8), consumed sequentially; any write drops
//     the key (transcript.h:46-62,89-105,174-178).
// The transcript is inherently sequential: one thread per proof runs it, the
// state lives in that thread's registers / local memory and is parked in HBM
// between kernels.
#pragma once
#include <stdint.h>

#ifndef LF_HD
#ifdef __CUDACC__
#define LF_HD __host__ __device__
#else
#define LF_HD
#endif
#endif

namespace lf {

// ---------------------------------------------------------------- SHA-256
#define LF_SHA256_K \
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, \
    0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, \
    0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, \
    0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, \
    0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, \
    0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, \
    0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, \
    0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, \
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, \
    0xc67178f2
#ifdef __CUDACC__
static __constant__ uint32_t kSha256K_dev[64] = {LF_SHA256_K};
#endif
static const uint32_t kSha256K_host[64] = {LF_SHA256_K};
LF_HD __forceinline__ uint32_t sha256_k(int i) {
#ifdef __CUDA_ARCH__
  return kSha256K_dev[i];
#else
  return kSha256K_host[i];
#endif
}

LF_HD __forceinline__ uint32_t rotr32(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(x, x, n);
#else
  return (x >> n) | (x << (32 - n));
#endif
}
LF_HD __forceinline__ uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __byte_perm(x, 0, 0x0123);
#else
  return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

LF_HD __forceinline__ void sha256_iv(uint32_t h[8]) {
  h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
  h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
}

// One compression; w[16] are the big-endian message words (clobbered).
LF_HD __forceinline__ void sha256_compress(uint32_t h[8], uint32_t w[16]) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i];
    } else {
      uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + sha256_k(i) + wi;
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// Compression of an all-zero block: the message schedule is identically zero,
// so only the round constants enter (used by Transcript::write0, the
// sequential nterms-zero-bytes write of lib/zk/zk_common.h:177-179).
LF_HD __forceinline__ void sha256_compress_zero(uint32_t h[8]) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + sha256_k(i);
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// The transcript thread's compression: ONE out-of-line copy, 16 rounds per loop
// iteration (~10 KB of SASS, inside the 32 KB L1.5 instruction cache).  The
// serial Fiat-Shamir path calls a compression from ~50 sites; inlining the
// fully unrolled rounds at each of them made the sumcheck kernel > 1 MB of
// straight-line code that a single warp executes once per call, i.e. every
// call ran at instruction-fetch speed (profiles/r1_sumcheck_occ8_full.txt).
// hs/ws may point to shared or local memory.
#ifdef __CUDACC__
#define LF_SHA_ROUND(K, W)                                              \
  {                                                                     \
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);         \
    uint32_t ch = (e & f) ^ (~e & g);                                   \
    uint32_t t1 = hh + S1 + ch + (K) + (W);                             \
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);         \
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);                          \
    uint32_t t2 = S0 + mj;                                              \
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2; \
  }
static __device__ __noinline__ void sha256_compress_fn(uint32_t* hs, const uint32_t* ws) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = ws[i];
  uint32_t a = hs[0], b = hs[1], c = hs[2], d = hs[3], e = hs[4], f = hs[5], g = hs[6], hh = hs[7];
  // branch-free bodies: rounds 0..15 straight, then three passes of 16 rounds
  // with the message schedule (a test inside the body splits it into one basic
  // block per round and the rounds no longer overlap)
#pragma unroll
  for (int i = 0; i < 16; ++i) LF_SHA_ROUND(kSha256K_dev[i], w[i])
#pragma unroll 1
  for (int j = 16; j < 64; j += 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
      LF_SHA_ROUND(kSha256K_dev[j + i], w[i])
    }
  }
  hs[0] += a; hs[1] += b; hs[2] += c; hs[3] += d; hs[4] += e; hs[5] += f; hs[6] += g; hs[7] += hh;
}
// The 64 rounds alone, on an already expanded message schedule w64[64] (for long
// messages whose bytes are known up front other threads expand the schedules, so
// that only this part stays serial)
static __device__ __noinline__ void sha256_rounds_fn(uint32_t* hs, const uint32_t* w64) {
  uint32_t a = hs[0], b = hs[1], c = hs[2], d = hs[3], e = hs[4], f = hs[5], g = hs[6], hh = hs[7];
#pragma unroll 1
  for (int j = 0; j < 64; j += 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) LF_SHA_ROUND(kSha256K_dev[j + i], w64[j + i])
  }
  hs[0] += a; hs[1] += b; hs[2] += c; hs[3] += d; hs[4] += e; hs[5] += f; hs[6] += g; hs[7] += hh;
}
// message schedule of one block given as 16 big-endian words
static __device__ __forceinline__ void sha256_expand(const uint32_t* w16, uint32_t* w64) {
#pragma unroll
  for (int i = 0; i < 16; ++i) w64[i] = w16[i];
#pragma unroll 4
  for (int i = 16; i < 64; ++i) {
    uint32_t w15 = w64[i - 15], w2 = w64[i - 2];
    uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
    uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
    w64[i] = w64[i - 16] + s0 + w64[i - 7] + s1;
  }
}
#undef LF_SHA_ROUND
#endif

// Incremental SHA-256 with a byte-granular buffer (transcript writes are 1, 8,
// 16 and 32 bytes long).
struct Sha256 {
  uint32_t h[8];
  uint32_t buf[16];  // big-endian words being filled
  uint64_t len;      // bytes absorbed

  // One compression of the buffered block.
  LF_HD void compress_block() {
#ifdef __CUDA_ARCH__
    sha256_compress_fn(h, buf);
#else
    uint32_t ww[16];
    for (int i = 0; i < 16; ++i) ww[i] = buf[i];
    sha256_compress(h, ww);
#endif
#pragma unroll
    for (int i = 0; i < 16; ++i) buf[i] = 0;
  }
  LF_HD static void compress_any(uint32_t* hh, uint32_t* w) {
#ifdef __CUDA_ARCH__
    sha256_compress_fn(hh, w);
#else
    sha256_compress(hh, w);
#endif
  }
  LF_HD void compress_zero_block() {
    uint32_t hh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hh[i] = h[i];
    sha256_compress_zero(hh);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = hh[i];
  }
  LF_HD void init() {
    sha256_iv(h);
    len = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) buf[i] = 0;
  }
  LF_HD void put_byte(uint8_t b) {
    uint32_t pos = (uint32_t)(len & 63);
    uint32_t wi = pos >> 2, sh = 24 - 8 * (pos & 3);
    buf[wi] |= (uint32_t)b << sh;
    ++len;
    if (pos == 63) compress_block();
  }
  LF_HD void update(const uint8_t* p, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) put_byte(p[i]);
  }
  // n zero bytes
  LF_HD void update_zero(uint64_t n) {
    while (n > 0 && (len & 63) != 0) {
      put_byte(0);
      --n;
    }
    if (n >= 64) {
      uint32_t hh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hh[i] = h[i];
      while (n >= 64) {
        sha256_compress_zero(hh);
        len += 64;
        n -= 64;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = hh[i];
    }
    while (n > 0) {
      put_byte(0);
      --n;
    }
  }
  // one 32-bit word whose bytes enter the stream most-significant first; the
  // stream position may be unaligned (tags are single bytes), so the word can
  // straddle two buffer words or two blocks
  LF_HD void put_word_be(uint32_t x) {
    const uint32_t pos = (uint32_t)(len & 63), k = pos & 3, wi = pos >> 2;
    len += 4;
    if (k == 0) {
      buf[wi] = x;
      if (wi == 15) compress_block();
    } else {
      const uint32_t sh = 8 * k;
      buf[wi] |= x >> sh;
      const uint32_t rest = x << (32 - sh);
      if (wi == 15) {
        compress_block();
        buf[0] = rest;
      } else {
        buf[wi + 1] |= rest;
      }
    }
  }
  // little-endian 32-bit words of a field element (wire bytes = LE)
  LF_HD void update_le_words(const uint32_t* w, int nwords) {
    for (int i = 0; i < nwords; ++i) put_word_be(bswap32(w[i]));
  }
  // nbytes (<= 4 * NS) bytes given as big-endian stream words S[0..NS) (unused low bytes of the last word
  // zero), appended at any byte position in one pass: the words are shifted to the buffer's alignment with
  // funnel shifts and OR-ed in (the buffer is zero beyond the write position), with at most one compression
  // in between.  The element writes of a Fiat-Shamir round (tag + 16 or 32 bytes, transcript.h:136-153) take
  // this path instead of five or nine byte-granular calls.
  template <int NS>
  LF_HD void put_stream_words(const uint32_t* S, uint32_t nbytes) {
    const uint32_t pos = (uint32_t)(len & 63), k = pos & 3, wi = pos >> 2, sh = 8 * k;
    len += nbytes;
    uint32_t T[NS + 1];
#ifdef __CUDA_ARCH__
    T[0] = S[0] >> sh;
#pragma unroll
    for (int j = 1; j < NS; ++j) T[j] = __funnelshift_r(S[j], S[j - 1], sh);
    T[NS] = __funnelshift_r(0u, S[NS - 1], sh);
#else
    T[0] = S[0] >> sh;
    for (int j = 1; j < NS; ++j) T[j] = (uint32_t)(((((uint64_t)S[j - 1]) << 32) | S[j]) >> sh);
    T[NS] = (uint32_t)((((uint64_t)S[NS - 1]) << 32) >> sh);
#endif
    const uint32_t nT = (k + nbytes + 3) >> 2;        // buffer words touched
    const uint32_t end = pos + nbytes;                // >= 64: the block fills
    const uint32_t nb = end >= 64 ? 16 - wi : nT;     // words that go in before the compression
#pragma unroll
    for (int j = 0; j <= NS; ++j)
      if ((uint32_t)j < nb) buf[wi + j] |= T[j];
    if (end >= 64) {
      compress_block();
#pragma unroll
      for (int j = 0; j <= NS; ++j)
        if ((uint32_t)j >= nb && (uint32_t)j < nT) buf[j - nb] = T[j];
    }
  }
  // tag byte + N little-endian element words
  template <int N>
  LF_HD void put_tagged_le_words(uint32_t tag, const uint32_t* w) {
    uint32_t S[N + 1];
    uint32_t prev = tag << 24;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const uint32_t x = bswap32(w[j]);
      S[j] = prev | (x >> 8);
      prev = x << 24;
    }
    S[N] = prev;
    put_stream_words<N + 1>(S, 4 * N + 1);
  }
  // N little-endian element words, no tag (array elements)
  template <int N>
  LF_HD void put_le_words(const uint32_t* w) {
    uint32_t S[N];
#pragma unroll
    for (int j = 0; j < N; ++j) S[j] = bswap32(w[j]);
    put_stream_words<N>(S, 4 * N);
  }
  // digest of the bytes so far, without disturbing the running state
  // (Transcript::get, lib/random/transcript.h:99-105). out = 8 big-endian words
  LF_HD void snapshot(uint32_t out[8]) const {
    uint32_t hh[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) hh[i] = h[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = buf[i];
    uint32_t pos = (uint32_t)(len & 63);
    w[pos >> 2] |= 0x80u << (24 - 8 * (pos & 3));
    if (pos >= 56) {
      compress_any(hh, w);
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = 0;
    }
    uint64_t bits = len * 8;
    w[14] = (uint32_t)(bits >> 32);
    w[15] = (uint32_t)bits;
    compress_any(hh, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = hh[i];
  }
};

// ---------------------------------------------------------------- AES-256
#define LF_AES_SBOX \
    0x63, 0x7c, 0x77, 0x7b, 0xf2, 0x6b, 0x6f, 0xc5, 0x30, 0x01, 0x67, 0x2b, 0xfe, 0xd7, 0xab, 0x76, \
    0xca, 0x82, 0xc9, 0x7d, 0xfa, 0x59, 0x47, 0xf0, 0xad, 0xd4, 0xa2, 0xaf, 0x9c, 0xa4, 0x72, 0xc0, \
    0xb7, 0xfd, 0x93, 0x26, 0x36, 0x3f, 0xf7, 0xcc, 0x34, 0xa5, 0xe5, 0xf1, 0x71, 0xd8, 0x31, 0x15, \
    0x04, 0xc7, 0x23, 0xc3, 0x18, 0x96, 0x05, 0x9a, 0x07, 0x12, 0x80, 0xe2, 0xeb, 0x27, 0xb2, 0x75, \
    0x09, 0x83, 0x2c, 0x1a, 0x1b, 0x6e, 0x5a, 0xa0, 0x52, 0x3b, 0xd6, 0xb3, 0x29, 0xe3, 0x2f, 0x84, \
    0x53, 0xd1, 0x00, 0xed, 0x20, 0xfc, 0xb1, 0x5b, 0x6a, 0xcb, 0xbe, 0x39, 0x4a, 0x4c, 0x58, 0xcf, \
    0xd0, 0xef, 0xaa, 0xfb, 0x43, 0x4d, 0x33, 0x85, 0x45, 0xf9, 0x02, 0x7f, 0x50, 0x3c, 0x9f, 0xa8, \
    0x51, 0xa3, 0x40, 0x8f, 0x92, 0x9d, 0x38, 0xf5, 0xbc, 0xb6, 0xda, 0x21, 0x10, 0xff, 0xf3, 0xd2, \
    0xcd, 0x0c, 0x13, 0xec, 0x5f, 0x97, 0x44, 0x17, 0xc4, 0xa7, 0x7e, 0x3d, 0x64, 0x5d, 0x19, 0x73, \
    0x60, 0x81, 0x4f, 0xdc, 0x22, 0x2a, 0x90, 0x88, 0x46, 0xee, 0xb8, 0x14, 0xde, 0x5e, 0x0b, 0xdb, \
    0xe0, 0x32, 0x3a, 0x0a, 0x49, 0x06, 0x24, 0x5c, 0xc2, 0xd3, 0xac, 0x62, 0x91, 0x95, 0xe4, 0x79, \
    0xe7, 0xc8, 0x37, 0x6d, 0x8d, 0xd5, 0x4e, 0xa9, 0x6c, 0x56, 0xf4, 0xea, 0x65, 0x7a, 0xae, 0x08, \
    0xba, 0x78, 0x25, 0x2e, 0x1c, 0xa6, 0xb4, 0xc6, 0xe8, 0xdd, 0x74, 0x1f, 0x4b, 0xbd, 0x8b, 0x8a, \
    0x70, 0x3e, 0xb5, 0x66, 0x48, 0x03, 0xf6, 0x0e, 0x61, 0x35, 0x57, 0xb9, 0x86, 0xc1, 0x1d, 0x9e, \
    0xe1, 0xf8, 0x98, 0x11, 0x69, 0xd9, 0x8e, 0x94, 0x9b, 0x1e, 0x87, 0xe9, 0xce, 0x55, 0x28, 0xdf, \
    0x8c, 0xa1, 0x89, 0x0d, 0xbf, 0xe6, 0x42, 0x68, 0x41, 0x99, 0x2d, 0x0f, 0xb0, 0x54, 0xbb, 0x16
#ifdef __CUDACC__
static __constant__ uint8_t kAesSbox_dev[256] = {LF_AES_SBOX};
#endif
static const uint8_t kAesSbox_host[256] = {LF_AES_SBOX};
// The S-box is read through a pointer: kernels stage the 256 bytes in shared
// memory (aes_stage_sbox), where a dependent byte lookup costs ~30 cycles
// instead of a divergent constant-bank access; host code uses the static table.
LF_HD __forceinline__ const uint8_t* aes_default_sbox() {
#ifdef __CUDA_ARCH__
  return kAesSbox_dev;
#else
  return kAesSbox_host;
#endif
}
// S-box plus the four round tables te[k][x] = rotl(Te0[x], 8k), Te0[x] = the
// MixColumns image (2s, s, s, 3s) of s = S[x] in a little-endian column word.
struct AesTables {
  uint32_t te[4][256];
  uint8_t sbox[256];
};
#ifdef __CUDACC__
// call from all threads of a CTA; sb = __shared__ uint8_t[256]
__device__ __forceinline__ void aes_stage_sbox(uint8_t* sb) {
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) sb[i] = kAesSbox_dev[i];
  __syncthreads();
}
// call from all threads of a CTA; t = __shared__ AesTables
__device__ __forceinline__ void aes_stage_tables(AesTables* t) {
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t s = kAesSbox_dev[i];
    uint32_t s2 = ((s << 1) ^ ((s >> 7) * 0x11bu)) & 0xffu, s3 = s2 ^ s;
    uint32_t t0 = s2 | (s << 8) | (s << 16) | (s3 << 24);
    t->sbox[i] = (uint8_t)s;
    t->te[0][i] = t0;
    t->te[1][i] = (t0 << 8) | (t0 >> 24);
    t->te[2][i] = (t0 << 16) | (t0 >> 16);
    t->te[3][i] = (t0 << 24) | (t0 >> 8);
  }
  __syncthreads();
}
#endif

// AES works on columns held as little-endian words: byte r of word c = state[r][c].
LF_HD __forceinline__ uint32_t aes_subword(const uint8_t* sb, uint32_t x) {
  return (uint32_t)sb[x & 0xff] | ((uint32_t)sb[(x >> 8) & 0xff] << 8) |
         ((uint32_t)sb[(x >> 16) & 0xff] << 16) | ((uint32_t)sb[x >> 24] << 24);
}
LF_HD __forceinline__ uint32_t aes_xtime4(uint32_t x) {  // xtime on 4 packed bytes
  return ((x & 0x7f7f7f7fu) << 1) ^ (((x >> 7) & 0x01010101u) * 0x1bu);
}

struct Aes256 {
  uint32_t rk[60];
  // key: 8 words, byte i of the key = byte (i&3) of word i>>2
  LF_HD void init(const uint32_t key[8], const uint8_t* sb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) rk[i] = key[i];
    uint32_t rcon = 1;
    for (int i = 8; i < 60; ++i) {
      uint32_t t = rk[i - 1];
      if ((i & 7) == 0) {
        t = aes_subword(sb, (t >> 8) | (t << 24)) ^ rcon;  // RotWord on LE-packed bytes
        rcon = (rcon << 1) ^ ((rcon >> 7) * 0x11bu);
      } else if ((i & 7) == 4) {
        t = aes_subword(sb, t);
      }
      rk[i] = rk[i - 8] ^ t;
    }
  }
#ifdef __CUDACC__
  // Device forms used by the transcript thread.  Key schedule: fully unrolled
  // over a sliding window of the last eight words (no dependent shared-memory
  // round trips); block: one table lookup per state byte.
  __device__ __forceinline__ void init_unrolled(const uint32_t key[8], const uint8_t* sb) {
    uint32_t w0 = key[0], w1 = key[1], w2 = key[2], w3 = key[3], w4 = key[4], w5 = key[5], w6 = key[6], w7 = key[7];
    rk[0] = w0; rk[1] = w1; rk[2] = w2; rk[3] = w3; rk[4] = w4; rk[5] = w5; rk[6] = w6; rk[7] = w7;
    uint32_t rcon = 1;
#pragma unroll
    for (int i = 8; i < 60; i += 8) {
      w0 ^= aes_subword(sb, (w7 >> 8) | (w7 << 24)) ^ rcon;
      rcon <<= 1;  // seven doublings from 1: stays below 0x80, no reduction
      w1 ^= w0; w2 ^= w1; w3 ^= w2;
      rk[i] = w0; rk[i + 1] = w1; rk[i + 2] = w2; rk[i + 3] = w3;
      if (i + 4 < 60) {
        w4 ^= aes_subword(sb, w3);
        w5 ^= w4; w6 ^= w5; w7 ^= w6;
        rk[i + 4] = w4; rk[i + 5] = w5; rk[i + 6] = w6; rk[i + 7] = w7;
      }
    }
  }
  __device__ __forceinline__ void encrypt_te(const uint32_t in[4], uint32_t out[4], const AesTables* T) const {
    uint32_t s0 = in[0] ^ rk[0], s1 = in[1] ^ rk[1], s2 = in[2] ^ rk[2], s3 = in[3] ^ rk[3];
#pragma unroll 1
    for (int r = 1; r < 14; ++r) {
      uint32_t t0 = T->te[0][s0 & 0xff] ^ T->te[1][(s1 >> 8) & 0xff] ^ T->te[2][(s2 >> 16) & 0xff] ^ T->te[3][s3 >> 24];
      uint32_t t1 = T->te[0][s1 & 0xff] ^ T->te[1][(s2 >> 8) & 0xff] ^ T->te[2][(s3 >> 16) & 0xff] ^ T->te[3][s0 >> 24];
      uint32_t t2 = T->te[0][s2 & 0xff] ^ T->te[1][(s3 >> 8) & 0xff] ^ T->te[2][(s0 >> 16) & 0xff] ^ T->te[3][s1 >> 24];
      uint32_t t3 = T->te[0][s3 & 0xff] ^ T->te[1][(s0 >> 8) & 0xff] ^ T->te[2][(s1 >> 16) & 0xff] ^ T->te[3][s2 >> 24];
      s0 = t0 ^ rk[4 * r];
      s1 = t1 ^ rk[4 * r + 1];
      s2 = t2 ^ rk[4 * r + 2];
      s3 = t3 ^ rk[4 * r + 3];
    }
    const uint8_t* sb = T->sbox;
    out[0] = ((uint32_t)sb[s0 & 0xff] | ((uint32_t)sb[(s1 >> 8) & 0xff] << 8) | ((uint32_t)sb[(s2 >> 16) & 0xff] << 16) |
              ((uint32_t)sb[s3 >> 24] << 24)) ^ rk[56];
    out[1] = ((uint32_t)sb[s1 & 0xff] | ((uint32_t)sb[(s2 >> 8) & 0xff] << 8) | ((uint32_t)sb[(s3 >> 16) & 0xff] << 16) |
              ((uint32_t)sb[s0 >> 24] << 24)) ^ rk[57];
    out[2] = ((uint32_t)sb[s2 & 0xff] | ((uint32_t)sb[(s3 >> 8) & 0xff] << 8) | ((uint32_t)sb[(s0 >> 16) & 0xff] << 16) |
              ((uint32_t)sb[s1 >> 24] << 24)) ^ rk[58];
    out[3] = ((uint32_t)sb[s3 & 0xff] | ((uint32_t)sb[(s0 >> 8) & 0xff] << 8) | ((uint32_t)sb[(s1 >> 16) & 0xff] << 16) |
              ((uint32_t)sb[s2 >> 24] << 24)) ^ rk[59];
  }
#endif
  LF_HD void encrypt(const uint32_t in[4], uint32_t out[4], const uint8_t* sb) const {
    uint32_t s0 = in[0] ^ rk[0], s1 = in[1] ^ rk[1], s2 = in[2] ^ rk[2], s3 = in[3] ^ rk[3];
    for (int r = 1; r <= 14; ++r) {
      // SubBytes + ShiftRows: new column c takes row k from column (c+k)&3
      uint32_t t0 = (uint32_t)sb[s0 & 0xff] | ((uint32_t)sb[(s1 >> 8) & 0xff] << 8) |
                    ((uint32_t)sb[(s2 >> 16) & 0xff] << 16) | ((uint32_t)sb[s3 >> 24] << 24);
      uint32_t t1 = (uint32_t)sb[s1 & 0xff] | ((uint32_t)sb[(s2 >> 8) & 0xff] << 8) |
                    ((uint32_t)sb[(s3 >> 16) & 0xff] << 16) | ((uint32_t)sb[s0 >> 24] << 24);
      uint32_t t2 = (uint32_t)sb[s2 & 0xff] | ((uint32_t)sb[(s3 >> 8) & 0xff] << 8) |
                    ((uint32_t)sb[(s0 >> 16) & 0xff] << 16) | ((uint32_t)sb[s1 >> 24] << 24);
      uint32_t t3 = (uint32_t)sb[s3 & 0xff] | ((uint32_t)sb[(s0 >> 8) & 0xff] << 8) |
                    ((uint32_t)sb[(s1 >> 16) & 0xff] << 16) | ((uint32_t)sb[s2 >> 24] << 24);
      if (r < 14) {
        // MixColumns on packed columns: out = 2*t ^ 3*rot8(t) ^ rot16(t) ^ rot24(t)
#define LF_MIX(t)                                                      \
  {                                                                    \
    uint32_t r1 = ((t) >> 8) | ((t) << 24);                            \
    uint32_t r2 = ((t) >> 16) | ((t) << 16);                           \
    uint32_t r3 = ((t) >> 24) | ((t) << 8);                            \
    (t) = aes_xtime4((t) ^ r1) ^ r1 ^ r2 ^ r3;                         \
  }
        LF_MIX(t0) LF_MIX(t1) LF_MIX(t2) LF_MIX(t3)
#undef LF_MIX
      }
      s0 = t0 ^ rk[4 * r];
      s1 = t1 ^ rk[4 * r + 1];
      s2 = t2 ^ rk[4 * r + 2];
      s3 = t3 ^ rk[4 * r + 3];
    }
    out[0] = s0; out[1] = s1; out[2] = s2; out[3] = s3;
  }
};

// ------------------------------------------------------------- Transcript
// What crosses the C ABI (lf_transcript in longfellow_b200.h, same layout): the
// SHA-256 state of everything written so far plus the read position of the
// challenge stream.  The AES key schedule is re-derived from the hash on import.
struct TranscriptState {
  uint32_t h[8];
  uint32_t buf[16];  // buffered message bytes as big-endian words
  uint64_t len;
  uint64_t nblock;
  uint32_t rdptr, have_prf;
  uint32_t saved[4];
};

struct Transcript {
  Sha256 sha;
  Aes256 prf;
  uint64_t nblock;
  uint32_t rdptr;     // byte read pointer into saved[]
  uint32_t have_prf;
  uint32_t saved[4];  // LE-packed bytes of the current PRF block
  const uint8_t* sbox;  // AES S-box to use (re-pointed by every kernel after loading the state)
  const AesTables* tab; // round tables in shared memory, or null: S-box only

  LF_HD void export_state(TranscriptState* o) const {
    for (int i = 0; i < 8; ++i) o->h[i] = sha.h[i];
    for (int i = 0; i < 16; ++i) o->buf[i] = sha.buf[i];
    o->len = sha.len;
    o->nblock = nblock;
    o->rdptr = rdptr;
    o->have_prf = have_prf;
    for (int i = 0; i < 4; ++i) o->saved[i] = saved[i];
  }
  // sbox (and tab) must be set before a state with a live challenge stream is imported
  LF_HD void import_state(const TranscriptState* s) {
    for (int i = 0; i < 8; ++i) sha.h[i] = s->h[i];
    for (int i = 0; i < 16; ++i) sha.buf[i] = s->buf[i];
    sha.len = s->len;
    have_prf = 0;
    nblock = 0;
    rdptr = 16;
    if (s->have_prf) {
      refill();  // re-keys from the hash (block 0 is recomputed and dropped)
      nblock = s->nblock;
      rdptr = s->rdptr;
      for (int i = 0; i < 4; ++i) saved[i] = s->saved[i];
    }
  }
  LF_HD void use_tables(const AesTables* t) {
    tab = t;
    sbox = t->sbox;
  }
  LF_HD void raw_byte(uint8_t b) {
    have_prf = 0;
    sha.put_byte(b);
  }
  LF_HD void raw_len(uint64_t x) {
    sha.put_word_be(bswap32((uint32_t)x));
    sha.put_word_be(bswap32((uint32_t)(x >> 32)));
  }
  // transcript.h:76-79
  LF_HD void init(const uint8_t* seed, uint32_t n) {
    sha.init();
    have_prf = 0;
    nblock = 0;
    rdptr = 16;
    sbox = aes_default_sbox();
    tab = nullptr;
    write_bytes(seed, n);
  }
  // transcript.h:116-121
  LF_HD void write_bytes(const uint8_t* p, uint32_t n) {
    raw_byte(0);
    raw_len(n);
    sha.update(p, n);
  }
  // byte string given as LE-packed 32-bit words (e.g. a digest kept in words)
  LF_HD void write_bytes_words(const uint32_t* w, uint32_t nwords) {
    raw_byte(0);
    raw_len(4ull * nwords);
    sha.update_le_words(w, (int)nwords);
  }
  // transcript.h:124-133
  LF_HD void write0(uint64_t n) {
    raw_byte(0);
    raw_len(n);
    sha.update_zero(n);
  }
  // transcript.h:136-141 ; e = wire-order little-endian words of the element
  LF_HD void write_elt_words(const uint32_t* e, int nwords) {
    have_prf = 0;
    if (nwords == 4) {
      sha.put_tagged_le_words<4>(1, e);
    } else if (nwords == 8) {
      sha.put_tagged_le_words<8>(1, e);
    } else {
      sha.put_byte(1);
      sha.update_le_words(e, nwords);
    }
  }
  // transcript.h:144-153 header of an array write; follow with n x elt_words()
  LF_HD void begin_array(uint64_t n) {
    raw_byte(2);
    raw_len(n);
  }
  LF_HD void elt_words(const uint32_t* e, int nwords) {
    if (nwords == 4) sha.put_le_words<4>(e);
    else if (nwords == 8) sha.put_le_words<8>(e);
    else sha.update_le_words(e, nwords);
  }

  // transcript.h:46-62,89-96
  // (re)key the PRF if a write intervened, then produce the next 16-byte block;
  // one out-of-line copy on the device (see sha256_compress_fn)
#ifdef __CUDACC__
  __host__ __device__ __noinline__
#endif
  void refill() {
    if (!have_prf) {
      uint32_t d[8], key[8];
      sha.snapshot(d);
#pragma unroll
      for (int i = 0; i < 8; ++i) key[i] = bswap32(d[i]);  // digest bytes, LE-packed
#ifdef __CUDA_ARCH__
      prf.init_unrolled(key, sbox);
#else
      prf.init(key, sbox);
#endif
      have_prf = 1;
      nblock = 0;
    }
    uint32_t in[4] = {(uint32_t)nblock, (uint32_t)(nblock >> 32), 0, 0};
    ++nblock;
#ifdef __CUDA_ARCH__
    if (tab)
      prf.encrypt_te(in, saved, tab);
    else
#endif
      prf.encrypt(in, saved, sbox);
    rdptr = 0;
  }
  LF_HD uint8_t next_byte() {
    if (!have_prf || rdptr == 16) refill();
    uint8_t b = (uint8_t)(saved[rdptr >> 2] >> (8 * (rdptr & 3)));
    ++rdptr;
    return b;
  }
  LF_HD void bytes(uint8_t* out, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) out[i] = next_byte();
  }
  // nwords LE words of challenge bytes
  LF_HD void words(uint32_t* out, int nwords) {
    for (int i = 0; i < nwords; ++i) {
      if (!have_prf || rdptr == 16) refill();
      if ((rdptr & 3) == 0) {
        out[i] = saved[rdptr >> 2];
        rdptr += 4;
      } else {
        uint32_t x = 0;
        for (int k = 0; k < 4; ++k) x |= (uint32_t)next_byte() << (8 * k);
        out[i] = x;
      }
    }
  }
  // RandomEngine::nat (lib/random/random.h:57-88)
  LF_HD uint32_t nat(uint32_t n) {
    uint32_t l = 0, nn = n;
    while (nn != 0) {
      nn >>= 8;
      ++l;
    }
    uint32_t msk = 0;
    while ((n & msk) != n) msk = (msk << 1) | 1u;
    uint32_t r;
    do {
      r = 0;
      for (uint32_t i = 0; i < l; ++i) r |= (uint32_t)next_byte() << (8 * i);
      r &= msk;
    } while (r >= n);
    return r;
  }
};

}  // namespace lf
