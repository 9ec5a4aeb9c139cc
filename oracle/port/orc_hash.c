/* oracle/port/orc_hash.c -- SHA-256, AES-256 and the Fiat-Shamir transcript
 * (TEST INFRASTRUCTURE).  The reference calls OpenSSL (util/crypto.h:41-103,
 * system libcrypto, unpinned); the primitives are FIPS 180-4 / FIPS 197 and
 * parity at that boundary is pinned by random/transcript_test.cc:131-341 and
 * merkle/merkle_tree_test.cc:186-236 (see tests/test_oracle_golden.py). */
#include <string.h>

#include "orc.h"

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4,
    0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe,
    0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f,
    0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7,
    0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc,
    0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b,
    0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116,
    0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(uint32_t h[8], const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; ++i)
    w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) |
           ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; ++i) {
    uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; ++i) {
    uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
    uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void sha256_init(sha256* s) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(s->h, iv, sizeof(iv));
  s->len = 0;
}
void sha256_update(sha256* s, const uint8_t* p, size_t n) {
  size_t fill = (size_t)(s->len & 63);
  s->len += n;
  if (fill) {
    size_t take = 64 - fill < n ? 64 - fill : n;
    memcpy(s->buf + fill, p, take);
    p += take;
    n -= take;
    if (fill + take < 64) return;
    sha256_block(s->h, s->buf);
  }
  while (n >= 64) {
    sha256_block(s->h, p);
    p += 64;
    n -= 64;
  }
  if (n) memcpy(s->buf, p, n);
}
void sha256_final(const sha256* s0, uint8_t out[32]) {
  sha256 s = *s0;
  uint64_t bits = s.len * 8;
  uint8_t pad[72];
  size_t fill = (size_t)(s.len & 63);
  size_t padlen = (fill < 56 ? 56 : 120) - fill;
  memset(pad, 0, sizeof(pad));
  pad[0] = 0x80;
  for (int i = 0; i < 8; ++i) pad[padlen + i] = (uint8_t)(bits >> (56 - 8 * i));
  sha256_update(&s, pad, padlen + 8);
  for (int i = 0; i < 8; ++i) {
    out[4 * i] = (uint8_t)(s.h[i] >> 24);
    out[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
    out[4 * i + 2] = (uint8_t)(s.h[i] >> 8);
    out[4 * i + 3] = (uint8_t)s.h[i];
  }
}

/* ---------------- AES-256 (FIPS 197), table-free ---------------- */
static uint8_t SBOX[256];
static int sbox_ready;
static uint8_t xt(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }
static uint8_t gmul(uint8_t a, uint8_t b) {
  uint8_t r = 0;
  while (b) {
    if (b & 1) r ^= a;
    a = xt(a);
    b >>= 1;
  }
  return r;
}
static void sbox_init(void) {
  if (sbox_ready) return;
  for (int x = 0; x < 256; ++x) {
    /* multiplicative inverse by brute force, then the affine map */
    uint8_t inv = 0;
    if (x)
      for (int y = 1; y < 256; ++y)
        if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
    uint8_t s = inv;
    uint8_t r = inv;
    for (int i = 0; i < 4; ++i) {
      s = (uint8_t)((s << 1) | (s >> 7));
      r ^= s;
    }
    SBOX[x] = r ^ 0x63;
  }
  sbox_ready = 1;
}
void aes256_init(aes256* a, const uint8_t key[32]) {
  sbox_init();
  uint8_t w[240];
  memcpy(w, key, 32);
  uint8_t rcon = 1;
  for (int i = 32; i < 240; i += 4) {
    uint8_t t[4] = {w[i - 4], w[i - 3], w[i - 2], w[i - 1]};
    if (i % 32 == 0) {
      uint8_t t0 = t[0];
      t[0] = SBOX[t[1]] ^ rcon;
      t[1] = SBOX[t[2]];
      t[2] = SBOX[t[3]];
      t[3] = SBOX[t0];
      rcon = xt(rcon);
    } else if (i % 32 == 16) {
      for (int k = 0; k < 4; ++k) t[k] = SBOX[t[k]];
    }
    for (int k = 0; k < 4; ++k) w[i + k] = w[i - 32 + k] ^ t[k];
  }
  memcpy(a->rk, w, 240);
}
void aes256_encrypt(const aes256* a, const uint8_t in[16], uint8_t out[16]) {
  uint8_t s[16], t[16];
  for (int i = 0; i < 16; ++i) s[i] = in[i] ^ a->rk[0][i];
  for (int r = 1; r <= 14; ++r) {
    /* SubBytes + ShiftRows: state is column major, s[4*c + row] */
    for (int c = 0; c < 4; ++c)
      for (int row = 0; row < 4; ++row) t[4 * c + row] = SBOX[s[4 * ((c + row) & 3) + row]];
    if (r < 14) {
      for (int c = 0; c < 4; ++c) {
        uint8_t* p = t + 4 * c;
        uint8_t a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
        uint8_t all = a0 ^ a1 ^ a2 ^ a3;
        p[0] = a0 ^ all ^ xt(a0 ^ a1);
        p[1] = a1 ^ all ^ xt(a1 ^ a2);
        p[2] = a2 ^ all ^ xt(a2 ^ a3);
        p[3] = a3 ^ all ^ xt(a3 ^ a0);
      }
    }
    for (int i = 0; i < 16; ++i) s[i] = t[i] ^ a->rk[r][i];
  }
  memcpy(out, s, 16);
}

/* ---------------- Transcript (random/transcript.h) ---------------- */
/* transcript.h:46-62 FSPRF::bytes/refill, :89-96 Transcript::bytes */
static void ts_bytes(rng* b, uint8_t* out, size_t n) {
  transcript* t = (transcript*)b;
  if (!t->have_prf) {
    uint8_t key[32];
    sha256_final(&t->sha, key); /* :99-105 get() */
    aes256_init(&t->prf, key);
    t->have_prf = 1;
    t->nblock = 0;
    t->rdptr = 16;
  }
  while (n-- > 0) {
    if (t->rdptr == 16) {
      uint8_t in[16];
      memset(in, 0, 16);
      for (int i = 0; i < 8; ++i) in[i] = (uint8_t)(t->nblock >> (8 * i));
      t->nblock++;
      aes256_encrypt(&t->prf, in, t->saved);
      t->rdptr = 0;
    }
    *out++ = t->saved[t->rdptr++];
  }
}
/* :174-178 write_untyped */
static void ts_raw(transcript* t, const uint8_t* p, size_t n) {
  t->have_prf = 0;
  sha256_update(&t->sha, p, n);
}
static void ts_tag(transcript* t, uint8_t tag) { ts_raw(t, &tag, 1); }
static void ts_len(transcript* t, uint64_t x) {
  uint8_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = (uint8_t)(x >> (8 * i));
  ts_raw(t, a, 8);
}
/* :76-79 */
void ts_init(transcript* t, const uint8_t* init, size_t n) {
  t->base.bytes = ts_bytes;
  sha256_init(&t->sha);
  t->have_prf = 0;
  ts_write_bytes(t, init, n);
}
/* :116-121 */
void ts_write_bytes(transcript* t, const uint8_t* p, size_t n) {
  ts_tag(t, 0);
  ts_len(t, n);
  ts_raw(t, p, n);
}
/* :124-133 */
void ts_write0(transcript* t, size_t n) {
  ts_tag(t, 0);
  ts_len(t, n);
  uint8_t z[32] = {0};
  for (; n > 32; n -= 32) ts_raw(t, z, 32);
  ts_raw(t, z, n);
}
/* :136-141 */
void ts_write_elt(transcript* t, const field* F, elt e) {
  uint8_t buf[32];
  ts_tag(t, 1);
  f_to_bytes(F, buf, e);
  ts_raw(t, buf, F->kbytes);
}
/* :144-153 */
void ts_write_array(transcript* t, const field* F, const elt* e, size_t ince, size_t n) {
  uint8_t buf[32];
  ts_tag(t, 2);
  ts_len(t, n);
  for (size_t i = 0; i < n; ++i) {
    f_to_bytes(F, buf, e[i * ince]);
    ts_raw(t, buf, F->kbytes);
  }
}
