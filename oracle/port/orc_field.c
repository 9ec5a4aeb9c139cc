/* oracle/port/orc_field.c -- field arithmetic restatement (TEST INFRASTRUCTURE).
 * GF(2^128): gf2k/gf2_128.h, gf2k/sysdep.h.  Prime fields: algebra/fp_generic.h,
 * fp.h, fp_p256.h, fp_p128.h.  See orc.h for the rules on who may call this. */
#include "orc.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ */
/* GF(2)[x]/(x^128+x^7+x^2+x+1)                                        */
/* ------------------------------------------------------------------ */

/* gf2k/sysdep.h:391-400 (portable gf2_128_mul) restated with a 4-bit window
 * instead of Kronecker substitution; the product is the same polynomial. */
static elt gf_mul(elt a, elt b) {
  uint64_t tab[16][3];
  memset(tab, 0, sizeof(tab));
  tab[1][0] = a.l[0];
  tab[1][1] = a.l[1];
  for (int i = 2; i < 16; ++i) {
    if (i & 1) {
      for (int k = 0; k < 3; ++k) tab[i][k] = tab[i - 1][k] ^ tab[1][k];
    } else {
      const uint64_t* s = tab[i / 2];
      tab[i][0] = s[0] << 1;
      tab[i][1] = (s[1] << 1) | (s[0] >> 63);
      tab[i][2] = (s[2] << 1) | (s[1] >> 63);
    }
  }
  uint64_t t[4] = {0, 0, 0, 0};
  for (int i = 31; i >= 0; --i) {
    t[3] = (t[3] << 4) | (t[2] >> 60);
    t[2] = (t[2] << 4) | (t[1] >> 60);
    t[1] = (t[1] << 4) | (t[0] >> 60);
    t[0] <<= 4;
    unsigned nib = (unsigned)((b.l[i / 16] >> (4 * (i % 16))) & 15u);
    t[0] ^= tab[nib][0];
    t[1] ^= tab[nib][1];
    t[2] ^= tab[nib][2];
  }
  /* reduce: x^128 = x^7 + x^2 + x + 1 (gf2k/sysdep.h:378-389) */
  uint64_t h0 = t[2], h1 = t[3];
  uint64_t r0 = h0 ^ (h0 << 1) ^ (h0 << 2) ^ (h0 << 7);
  uint64_t r1 = h1 ^ (h1 << 1) ^ (h1 << 2) ^ (h1 << 7) ^ (h0 >> 63) ^ (h0 >> 62) ^ (h0 >> 57);
  uint64_t r2 = (h1 >> 63) ^ (h1 >> 62) ^ (h1 >> 57);
  r0 ^= r2 ^ (r2 << 1) ^ (r2 << 2) ^ (r2 << 7);
  elt o = {{t[0] ^ r0, t[1] ^ r1, 0, 0}};
  return o;
}

static int gf_bit(const uint64_t u[2], int j) { return (int)((u[j >> 6] >> (j & 63)) & 1); }

/* gf2k/gf2_128.h:274-310 computes the inverse by a binary Euclid; the inverse
 * is unique, so Fermat a^(2^128-2) yields the same element. */
static elt gf_inv(elt a) {
  /* a^(2^128-2) = prod_{i=1..127} a^(2^i) */
  elt r = {{1, 0, 0, 0}};
  elt s = a;
  for (int i = 1; i < 128; ++i) {
    s = gf_mul(s, s);
    r = gf_mul(r, s);
  }
  return r;
}

/* gf2k/gf2_128.h:369-391 subfield_generator, :108-119 beta, :451-493 beta_ref */
static void gf_setup(field* F) {
  elt x = {{2, 0, 0, 0}};
  elt r = x;
  for (int i = 4; i < 7; ++i) {
    elt s = r;
    for (int j = 0; j < (1 << i); ++j) s = gf_mul(s, s);
    r = gf_mul(r, s);
  }
  elt g = r;
  F->beta[0] = F->one;
  for (int i = 1; i < 16; ++i) F->beta[i] = gf_mul(F->beta[i - 1], g);
  for (int i = 0; i < 16; ++i) {
    F->sub_u[i][0] = F->beta[i].l[0];
    F->sub_u[i][1] = F->beta[i].l[1];
    F->sub_linv[i] = (uint64_t)1 << i;
  }
  int rnk = 0;
  for (int j = 0; rnk < 16 && j < 128; ++j) {
    int piv = -1;
    for (int i = rnk; i < 16; ++i) {
      if (gf_bit(F->sub_u[i], j)) { piv = i; break; }
    }
    if (piv < 0) continue;
    for (int k = 0; k < 2; ++k) {
      uint64_t t = F->sub_u[rnk][k]; F->sub_u[rnk][k] = F->sub_u[piv][k]; F->sub_u[piv][k] = t;
    }
    uint64_t t = F->sub_linv[rnk]; F->sub_linv[rnk] = F->sub_linv[piv]; F->sub_linv[piv] = t;
    F->sub_ldnz[rnk] = j;
    for (int i1 = rnk + 1; i1 < 16; ++i1) {
      if (gf_bit(F->sub_u[i1], j)) {
        F->sub_u[i1][0] ^= F->sub_u[rnk][0];
        F->sub_u[i1][1] ^= F->sub_u[rnk][1];
        F->sub_linv[i1] ^= F->sub_linv[rnk];
      }
    }
    ++rnk;
  }
  if (rnk != 16) abort();
  /* gf2_128.h:121-127 evaluation points 0, g^0, g^1, ...; :129-137 Newton */
  F->nevalpt = 6;
  F->evalpt[0] = F->zero;
  elt gi = F->one;
  for (int i = 1; i < 6; ++i) {
    F->evalpt[i] = gi;
    gi = gf_mul(gi, g);
  }
  for (int i = 1; i < 6; ++i) {
    for (int k = 5; k >= i; --k) {
      elt dx = f_sub(F, F->evalpt[k], F->evalpt[k - i]);
      F->newton[k][i] = gf_inv(dx);
    }
  }
}

/* gf2_128.h:495-508 solve(): returns residual in ue, coordinates in *u */
static void gf_solve(const field* F, elt e, uint64_t ue[2], uint64_t* u) {
  ue[0] = e.l[0];
  ue[1] = e.l[1];
  *u = 0;
  for (int rnk = 0; rnk < 16; ++rnk) {
    if (gf_bit(ue, F->sub_ldnz[rnk])) {
      ue[0] ^= F->sub_u[rnk][0];
      ue[1] ^= F->sub_u[rnk][1];
      *u ^= F->sub_linv[rnk];
    }
  }
}

/* ------------------------------------------------------------------ */
/* Montgomery prime fields (algebra/fp_generic.h)                      */
/* ------------------------------------------------------------------ */

static int geq(const uint64_t* a, const uint64_t* b, int w) {
  for (int i = w - 1; i >= 0; --i) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static uint64_t addn(uint64_t* r, const uint64_t* a, const uint64_t* b, int w) {
  u128 c = 0;
  for (int i = 0; i < w; ++i) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
static uint64_t subn(uint64_t* r, const uint64_t* a, const uint64_t* b, int w) {
  uint64_t br = 0;
  for (int i = 0; i < w; ++i) {
    u128 d = (u128)a[i] - b[i] - br;
    r[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
/* fp_generic.h:161-169 */
static elt fp_add(const field* F, elt a, elt b) {
  elt r = {{0, 0, 0, 0}};
  uint64_t c = addn(r.l, a.l, b.l, F->w64);
  if (c || geq(r.l, F->m, F->w64)) subn(r.l, r.l, F->m, F->w64);
  return r;
}
/* fp_generic.h:175-182 */
static elt fp_sub(const field* F, elt a, elt b) {
  elt r = {{0, 0, 0, 0}};
  if (subn(r.l, a.l, b.l, F->w64)) addn(r.l, r.l, F->m, F->w64);
  return r;
}
/* fp_generic.h:187-198,484-519 (CIOS Montgomery product, canonical result).
 * Fp256Reduce / Fp128Reduce (fp_p256.h:42-52, fp_p128.h:68-75) are special
 * reduction steps computing the same REDC value. */
static elt fp_mul(const field* F, elt a, elt b) {
  int w = F->w64;
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < w; ++i) {
    u128 c = 0;
    for (int j = 0; j < w; ++j) {
      c += (u128)a.l[j] * b.l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[w];
    t[w] = (uint64_t)c;
    t[w + 1] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * F->mprime;
    c = (u128)q * F->m[0] + t[0];
    c >>= 64;
    for (int j = 1; j < w; ++j) {
      c += (u128)q * F->m[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[w];
    t[w - 1] = (uint64_t)c;
    t[w] = t[w + 1] + (uint64_t)(c >> 64);
  }
  elt r = {{0, 0, 0, 0}};
  for (int i = 0; i < w; ++i) r.l[i] = t[i];
  if (t[w] || geq(r.l, F->m, w)) subn(r.l, r.l, F->m, w);
  return r;
}
static elt fp_from_mont(const field* F, elt a) {
  elt one = {{1, 0, 0, 0}};
  return fp_mul(F, a, one);
}
static elt fp_to_mont(const field* F, elt a) { return fp_mul(F, a, F->rsq); }

/* fp_generic.h:233-253 uses a binary xgcd; the inverse is unique so Fermat
 * a^(p-2) gives the same Montgomery representative. */
static elt fp_inv(const field* F, elt a) {
  uint64_t e[4] = {0, 0, 0, 0}, two[4] = {2, 0, 0, 0};
  subn(e, F->m, two, F->w64);
  elt r = F->one, s = a;
  for (int i = 0; i < 64 * F->w64; ++i) {
    if ((e[i >> 6] >> (i & 63)) & 1) r = fp_mul(F, r, s);
    s = fp_mul(F, s, s);
  }
  return r;
}

static void parse_dec(const char* s, uint64_t out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
  for (; *s; ++s) {
    u128 c = (u128)(*s - '0');
    for (int i = 0; i < 4; ++i) {
      c += (u128)out[i] * 10;
      out[i] = (uint64_t)c;
      c >>= 64;
    }
  }
}

static void fp_setup(field* F, int w64, const uint64_t m[4]) {
  F->char2 = 0;
  F->w64 = w64;
  F->kbytes = F->ksubbytes = 8 * (size_t)w64;
  memcpy(F->m, m, 32);
  /* fp_generic.h:93-96 exact_bits */
  F->exact_bits = 64 * w64;
  while (((m[(F->exact_bits - 1) >> 6] >> ((F->exact_bits - 1) & 63)) & 1) == 0) --F->exact_bits;
  /* mprime = -m^{-1} mod 2^64 (fp_generic.h:104) */
  uint64_t inv = 1;
  for (int i = 0; i < 6; ++i) inv *= 2 - m[0] * inv;
  F->mprime = (uint64_t)0 - inv;
  memset(&F->zero, 0, sizeof(elt));
  /* rsquare = 2^(2*kBits) mod m by repeated doubling (fp_generic.h:105-108) */
  elt r = {{1, 0, 0, 0}};
  for (int i = 0; i < 2 * 64 * w64; ++i) r = fp_add(F, r, r);
  F->rsq = r;
  elt one = {{1, 0, 0, 0}};
  F->one = fp_to_mont(F, one);
  F->nevalpt = 6;
  for (int i = 0; i < 6; ++i) {
    elt s = {{(uint64_t)i, 0, 0, 0}};
    F->evalpt[i] = fp_to_mont(F, s);
  }
  /* newton_denominator(k,i) = 1/i (fp_generic.h:407-413) */
  for (int i = 1; i < 6; ++i) {
    elt inv_i = fp_inv(F, F->evalpt[i]);
    for (int k = i; k < 6; ++k) F->newton[k][i] = inv_i;
  }
}

/* ------------------------------------------------------------------ */
/* field table                                                         */
/* ------------------------------------------------------------------ */
static field g_fields[6];
static int g_init[6];

const field* orc_field(int id) {
  int slot;
  switch (id) {
    case ORC_GF2_128: slot = 0; break;
    case ORC_P256: slot = 1; break;
    case ORC_BN254: slot = 2; break;
    case ORC_FP128: slot = 3; break;
    case ORC_GOLDILOCKS: slot = 4; break;
    case ORC_SECP256K1: slot = 5; break;
    default: return NULL;
  }
  field* F = &g_fields[slot];
  if (g_init[slot]) return F;
  memset(F, 0, sizeof(*F));
  F->id = id;
  if (id == ORC_GF2_128) {
    F->char2 = 1;
    F->w64 = 2;
    F->kbytes = 16;
    F->ksubbytes = 2;
    F->one.l[0] = 1;
    gf_setup(F);
  } else if (id == ORC_P256) {
    /* fp_p256.h:34-39 */
    uint64_t m[4] = {0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFull, 0, 0xFFFFFFFF00000001ull};
    fp_setup(F, 4, m);
    /* circuits/mdoc/mdoc_zk.cc:83-88, ecdsa/verify_test.cc:519-530: root of
     * unity of order 2^31 in Fp2 = Fp[i]/(i^2+1) */
    uint64_t x[4], y[4];
    parse_dec("112649224146410281873500457609690258373018840430489408729223714171582664680802", x);
    parse_dec("84087994358540907695740461427818660560182168997182378749313018254450460212908", y);
    elt ex, ey;
    memcpy(ex.l, x, 32);
    memcpy(ey.l, y, 32);
    F->omega2[0] = fp_to_mont(F, ex);
    F->omega2[1] = fp_to_mont(F, ey);
    F->omega_order = 1ull << 31;
    F->has_omega = 2;
  } else if (id == ORC_SECP256K1) {
    uint64_t m[4];
    parse_dec("115792089237316195423570985008687907853269984665640564039457584007908834671663", m);
    fp_setup(F, 4, m);
  } else if (id == ORC_BN254) {
    uint64_t m[4], o[4];
    parse_dec("21888242871839275222246405745257275088548364400416034343698204186575808495617", m);
    fp_setup(F, 4, m);
    parse_dec("19103219067921713944291392827692070036145651957329286315305642004821462161904", o);
    elt eo;
    memcpy(eo.l, o, 32);
    F->omega = fp_to_mont(F, eo);
    F->omega_order = 1ull << 28;
    F->has_omega = 1;
  } else if (id == ORC_FP128) {
    /* fp_p128.h: p = 2^128 - 2^108 + 1 */
    uint64_t m[4] = {1, 0xFFFFF00000000000ull, 0, 0}, o[4];
    fp_setup(F, 2, m);
    parse_dec("164956748514267535023998284330560247862", o);
    elt eo;
    memcpy(eo.l, o, 32);
    F->omega = fp_to_mont(F, eo);
    F->omega_order = 1ull << 32;
    F->has_omega = 1;
  } else if (id == ORC_GOLDILOCKS) {
    uint64_t m[4] = {0xFFFFFFFF00000001ull, 0, 0, 0}, o[4];
    fp_setup(F, 1, m);
    parse_dec("1753635133440165772", o);
    elt eo;
    memcpy(eo.l, o, 32);
    F->omega = fp_to_mont(F, eo);
    F->omega_order = 1ull << 32;
    F->has_omega = 1;
  }
  g_init[slot] = 1;
  return F;
}

/* ------------------------------------------------------------------ */
/* generic dispatch                                                    */
/* ------------------------------------------------------------------ */
elt f_add(const field* F, elt a, elt b) {
  if (F->char2) {
    elt r = {{a.l[0] ^ b.l[0], a.l[1] ^ b.l[1], 0, 0}};
    return r;
  }
  return fp_add(F, a, b);
}
elt f_sub(const field* F, elt a, elt b) {
  if (F->char2) return f_add(F, a, b);
  return fp_sub(F, a, b);
}
elt f_mul(const field* F, elt a, elt b) { return F->char2 ? gf_mul(a, b) : fp_mul(F, a, b); }
elt f_neg(const field* F, elt a) { return F->char2 ? a : fp_sub(F, F->zero, a); }
elt f_inv(const field* F, elt a) { return F->char2 ? gf_inv(a) : fp_inv(F, a); }
int f_eq(const field* F, elt a, elt b) {
  for (int i = 0; i < F->w64; ++i)
    if (a.l[i] != b.l[i]) return 0;
  return 1;
}
int f_is_zero(const field* F, elt a) { return f_eq(F, a, F->zero); }

/* gf2_128.h:151-160 / fp_generic.h:285 */
elt f_of_scalar(const field* F, uint64_t u) {
  if (F->char2) {
    elt t = F->zero;
    for (int k = 0; k < 16; ++k, u >>= 1)
      if (u & 1) t = f_add(F, t, F->beta[k]);
    if (u != 0) abort();
    return t;
  }
  elt s = {{u, 0, 0, 0}};
  if (F->w64 == 1 && u >= F->m[0]) abort();
  return fp_to_mont(F, s);
}

static void le_load(uint64_t* l, const uint8_t* b, size_t nbytes) {
  for (size_t i = 0; i < nbytes; ++i) l[i >> 3] |= (uint64_t)b[i] << (8 * (i & 7));
}
/* gf2_128.h:171-176 / fp_generic.h:351-358 */
int f_of_bytes(const field* F, const uint8_t* b, elt* out) {
  elt s = {{0, 0, 0, 0}};
  le_load(s.l, b, F->kbytes);
  if (F->char2) {
    *out = s;
    return 0;
  }
  if (geq(s.l, F->m, F->w64)) return -1;
  *out = fp_to_mont(F, s);
  return 0;
}
/* gf2_128.h:178-180 / fp_generic.h:378-380 */
void f_to_bytes(const field* F, uint8_t* b, elt a) {
  elt s = F->char2 ? a : fp_from_mont(F, a);
  for (size_t i = 0; i < F->kbytes; ++i) b[i] = (uint8_t)(s.l[i >> 3] >> (8 * (i & 7)));
}
/* gf2_128.h:201-204 / fp_generic.h:278 */
int f_in_subfield(const field* F, elt a) {
  if (!F->char2) return 1;
  uint64_t ue[2], u;
  gf_solve(F, a, ue, &u);
  return ue[0] == 0 && ue[1] == 0;
}
/* gf2_128.h:216-224 */
void f_to_bytes_subfield(const field* F, uint8_t* b, elt a) {
  if (!F->char2) {
    f_to_bytes(F, b, a);
    return;
  }
  uint64_t ue[2], u;
  gf_solve(F, a, ue, &u);
  if (ue[0] || ue[1]) abort();
  b[0] = (uint8_t)u;
  b[1] = (uint8_t)(u >> 8);
}
/* gf2_128.h:206-214 */
elt f_of_bytes_subfield(const field* F, const uint8_t* b) {
  if (!F->char2) {
    elt e = F->zero;
    if (f_of_bytes(F, b, &e)) abort();
    return e;
  }
  return f_of_scalar(F, (uint64_t)b[0] | ((uint64_t)b[1] << 8));
}

/* ------------------------------------------------------------------ */
/* RandomEngine helpers (random/random.h:32-116)                       */
/* ------------------------------------------------------------------ */
/* gf2_128.h:182-190 / fp_generic.h:360-371 */
elt rng_elt(rng* r, const field* F) {
  if (F->char2) {
    uint8_t buf[16];
    r->bytes(r, buf, 16);
    elt e;
    f_of_bytes(F, buf, &e);
    return e;
  }
  size_t total = ((size_t)F->exact_bits + 7) / 8;
  uint8_t buf[32];
  memset(buf, 0, sizeof(buf));
  for (;;) {
    r->bytes(r, buf, total);
    elt s = {{0, 0, 0, 0}};
    le_load(s.l, buf, F->kbytes);
    size_t nbits = (size_t)F->exact_bits;
    for (int i = 0; i < F->w64; ++i) {
      if (nbits >= 64) {
        nbits -= 64;
      } else {
        s.l[i] &= ((uint64_t)1 << nbits) - 1;
        nbits = 0;
      }
    }
    if (!geq(s.l, F->m, F->w64)) return fp_to_mont(F, s);
  }
}
/* gf2_128.h:192-199 / fp_generic.h:373-376 */
elt rng_subfield_elt(rng* r, const field* F) {
  if (!F->char2) return rng_elt(r, F);
  uint8_t buf[2];
  r->bytes(r, buf, 2);
  return f_of_bytes_subfield(F, buf);
}
/* random.h:57-88 */
size_t rng_nat(rng* r, size_t n) {
  size_t l = 0, nn = n;
  while (nn != 0) {
    nn >>= 8;
    ++l;
  }
  size_t msk = 0;
  while ((n & msk) != n) msk = (msk << 1) | 1u;
  size_t v;
  uint8_t buf[8];
  do {
    r->bytes(r, buf, l);
    v = 0;
    for (size_t i = l; i-- > 0;) v = (v << 8) | buf[i];
    v &= msk;
  } while (v >= n);
  return v;
}
/* random.h:92-105 */
void rng_choose(rng* r, size_t* res, size_t n, size_t k) {
  size_t* A = (size_t*)malloc(n * sizeof(size_t));
  for (size_t i = 0; i < n; ++i) A[i] = i;
  for (size_t i = 0; i < k; ++i) {
    size_t j = i + rng_nat(r, n - i);
    size_t t = A[i];
    A[i] = A[j];
    A[j] = t;
    res[i] = A[i];
  }
  free(A);
}

static void bufrng_bytes(rng* b, uint8_t* out, size_t n) {
  bufrng* r = (bufrng*)b;
  if (r->pos + n > r->n) {
    r->overrun = 1;
    memset(out, 0, n);
    return;
  }
  memcpy(out, r->p + r->pos, n);
  r->pos += n;
}
void bufrng_init(bufrng* r, const uint8_t* p, size_t n) {
  r->base.bytes = bufrng_bytes;
  r->p = p;
  r->n = n;
  r->pos = 0;
  r->overrun = 0;
}
