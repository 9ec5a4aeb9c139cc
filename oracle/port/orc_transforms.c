/* oracle/port/orc_transforms.c -- LCH14 additive FFT, prime-field FFT and the
 * Reed-Solomon extension (TEST INFRASTRUCTURE; see orc.h). */
#include <stdlib.h>
#include <string.h>

#include "orc.h"

/* ------------------------------------------------------------------ */
/* LCH14 (gf2k/lch14.h)                                                */
/* ------------------------------------------------------------------ */
static elt WHAT[16][16];
static int what_ready;

/* lch14.h:45-77 */
static void what_init(void) {
  if (what_ready) return;
  const field* F = orc_field(ORC_GF2_128);
  for (int j = 0; j < 16; ++j) WHAT[0][j] = F->beta[j];
  for (int i = 0; i + 1 < 16; ++i)
    for (int j = 0; j < 16; ++j)
      WHAT[i + 1][j] = f_mul(F, WHAT[i][j], f_add(F, WHAT[i][j], WHAT[i][i]));
  for (int i = 0; i < 16; ++i) {
    elt scale = f_inv(F, WHAT[i][i]);
    for (int j = 0; j < 16; ++j) WHAT[i][j] = f_mul(F, scale, WHAT[i][j]);
  }
  what_ready = 1;
}
elt lch14_what(size_t i, size_t j) {
  what_init();
  return WHAT[i][j];
}
/* lch14.h:81-89 */
static elt twiddle(size_t i, size_t u) {
  const field* F = orc_field(ORC_GF2_128);
  elt t = F->zero;
  for (size_t k = 0; u != 0; ++k, u >>= 1)
    if (u & 1) t = f_add(F, t, WHAT[i][k]);
  return t;
}
/* lch14.h:219-237 */
static void bf_fwd(const field* F, elt* B, size_t uv, size_t s, elt tw) {
  B[uv] = f_add(F, B[uv], f_mul(F, tw, B[uv + s]));
  B[uv + s] = f_add(F, B[uv + s], B[uv]);
}
static void bf_bwd(const field* F, elt* B, size_t uv, size_t s, elt tw) {
  B[uv + s] = f_add(F, B[uv + s], B[uv]);
  B[uv] = f_add(F, B[uv], f_mul(F, tw, B[uv + s]));
}
static void bf_diag(const field* F, elt* B, size_t uv, size_t s, elt tw) {
  elt b1 = B[uv + s];
  B[uv + s] = f_add(F, B[uv + s], B[uv]);
  B[uv] = f_add(F, B[uv], f_mul(F, tw, b1));
}
/* lch14.h:106-123.  The twiddle of block u at stage i with coset c is
 * twiddle(i, c) + sum over bits k of u of w_hat[i][i+1+k]  (lch14.h:92-100),
 * i.e. twiddle(i, c ^ (u << (i+1))) when c is a multiple of 2^l. We compute
 * it from the definition tw[u] = twiddle(i,coset) + sum_k bit_k(u) w_hat[i][i+1+k]. */
static elt stage_tw(const field* F, size_t i, size_t coset, size_t u) {
  elt t = twiddle(i, coset);
  for (size_t k = 0; u != 0; ++k, u >>= 1)
    if (u & 1) t = f_add(F, t, WHAT[i][(i + 1) + k]);
  return t;
}
void lch14_fft(size_t l, size_t coset, elt* B) {
  what_init();
  const field* F = orc_field(ORC_GF2_128);
  for (size_t i = l; i-- > 0;) {
    size_t s = (size_t)1 << i;
    for (size_t u = 0; (u << (i + 1)) < ((size_t)1 << l); ++u) {
      elt tw = stage_tw(F, i, coset, u);
      for (size_t v = 0; v < s; ++v) bf_fwd(F, B, (u << (i + 1)) + v, s, tw);
    }
  }
}
/* lch14.h:125-143 */
void lch14_ifft(size_t l, size_t coset, elt* B) {
  what_init();
  const field* F = orc_field(ORC_GF2_128);
  for (size_t i = 0; i < l; ++i) {
    size_t s = (size_t)1 << i;
    for (size_t u = 0; (u << (i + 1)) < ((size_t)1 << l); ++u) {
      elt tw = stage_tw(F, i, coset, u);
      for (size_t v = 0; v < s; ++v) bf_bwd(F, B, (u << (i + 1)) + v, s, tw);
    }
  }
}
/* lch14.h:185-217 */
static void bidir_recur(const field* F, size_t i, size_t coset, size_t k, elt* B) {
  if (i-- > 0) {
    size_t s = (size_t)1 << i;
    elt tw = twiddle(i, coset);
    if (k < s) {
      for (size_t uv = k; uv < s; ++uv) bf_fwd(F, B, uv, s, tw);
      bidir_recur(F, i, coset, k, B);
      for (size_t uv = 0; uv < k; ++uv) bf_diag(F, B, uv, s, tw);
      lch14_fft(i, coset + s, B + s);
    } else {
      lch14_ifft(i, coset, B);
      for (size_t uv = k - s; uv < s; ++uv) bf_diag(F, B, uv, s, tw);
      bidir_recur(F, i, coset + s, k - s, B + s);
      for (size_t uv = 0; uv < k - s; ++uv) bf_bwd(F, B, uv, s, tw);
    }
  }
}
void lch14_bidir(size_t l, size_t k, elt* B) {
  what_init();
  bidir_recur(orc_field(ORC_GF2_128), l, 0, k, B);
}

/* gf2k/lch14_reed_solomon.h:49-103 */
static void lch14_interpolate(const field* F, size_t n, size_t m, elt* y) {
  size_t l = 0, fftn = 1;
  while (fftn < n) {
    fftn <<= 1;
    ++l;
  }
  elt* C = (elt*)calloc(fftn, sizeof(elt));
  for (size_t i = 0; i < n; ++i) C[i] = y[i];
  lch14_bidir(l, n, C);
  for (size_t i = n; i < (m < fftn ? m : fftn); ++i) y[i] = C[i];
  for (size_t i = n; i < fftn; ++i) C[i] = F->zero;
  for (size_t coset = 1; (coset << l) < m; ++coset) {
    size_t b = coset << l;
    if (b + fftn <= m) {
      for (size_t i = 0; i < fftn; ++i) y[i + b] = C[i];
      lch14_fft(l, b, &y[b]);
    } else {
      lch14_fft(l, b, C);
      for (size_t i = 0; i + b < m; ++i) y[i + b] = C[i];
    }
  }
  free(C);
}

/* ------------------------------------------------------------------ */
/* prime-field FFT (algebra/fft.h)                                     */
/* ------------------------------------------------------------------ */
static size_t bitrev(size_t x, size_t lg) {
  size_t r = 0;
  for (size_t i = 0; i < lg; ++i) r |= ((x >> i) & 1) << (lg - 1 - i);
  return r;
}
/* fft.h:27-46: fftb computes T[j] = sum_k F[k] w^{jk}; fftf uses w^{-1}.
 * fft.h:70-89 is an iterative radix-2 DIT after bit reversal; any exact DFT
 * algorithm yields the same array. */
void fp_fft(const field* F, elt* A, size_t n, elt omega, uint64_t order, int fwd) {
  if (n <= 1) return;
  size_t lg = 0;
  while (((size_t)1 << lg) < n) ++lg;
  if (fwd) omega = f_inv(F, omega);
  /* twiddle.h:46-54 reroot: omega_n = omega^(order/n) */
  for (uint64_t r = n; r < order; r += r) omega = f_mul(F, omega, omega);
  for (size_t i = 0; i < n; ++i) {
    size_t j = bitrev(i, lg);
    if (i < j) {
      elt t = A[i];
      A[i] = A[j];
      A[j] = t;
    }
  }
  elt* w = (elt*)malloc((n / 2) * sizeof(elt));
  w[0] = F->one;
  for (size_t i = 1; i < n / 2; ++i) w[i] = f_mul(F, w[i - 1], omega);
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t half = len / 2, step = n / len;
    for (size_t i = 0; i < n; i += len)
      for (size_t j = 0; j < half; ++j) {
        elt t = f_mul(F, A[i + j + half], w[j * step]);
        elt u = A[i + j];
        A[i + j] = f_add(F, u, t);
        A[i + j + half] = f_sub(F, u, t);
      }
  }
  free(w);
}

/* Fp2 = Fp[i]/(i^2+1) (algebra/fp2.h:79-123) */
static void c_mul(const field* F, const elt a[2], const elt b[2], elt r[2]) {
  elt rr = f_sub(F, f_mul(F, a[0], b[0]), f_mul(F, a[1], b[1]));
  elt ri = f_add(F, f_mul(F, a[0], b[1]), f_mul(F, a[1], b[0]));
  r[0] = rr;
  r[1] = ri;
}
void fp2_fft(const field* F, elt* A, size_t n, const elt omega_in[2], uint64_t order, int fwd) {
  if (n <= 1) return;
  size_t lg = 0;
  while (((size_t)1 << lg) < n) ++lg;
  elt om[2] = {omega_in[0], omega_in[1]};
  if (fwd) {
    /* 1/(a+bi) = (a-bi)/(a^2+b^2) */
    elt nrm = f_inv(F, f_add(F, f_mul(F, om[0], om[0]), f_mul(F, om[1], om[1])));
    om[0] = f_mul(F, om[0], nrm);
    om[1] = f_neg(F, f_mul(F, om[1], nrm));
  }
  for (uint64_t r = n; r < order; r += r) c_mul(F, om, om, om);
  for (size_t i = 0; i < n; ++i) {
    size_t j = bitrev(i, lg);
    if (i < j) {
      elt t0 = A[2 * i], t1 = A[2 * i + 1];
      A[2 * i] = A[2 * j];
      A[2 * i + 1] = A[2 * j + 1];
      A[2 * j] = t0;
      A[2 * j + 1] = t1;
    }
  }
  elt* w = (elt*)malloc(n * sizeof(elt));
  w[0] = F->one;
  w[1] = F->zero;
  for (size_t i = 1; i < n / 2; ++i) c_mul(F, &w[2 * (i - 1)], om, &w[2 * i]);
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t half = len / 2, step = n / len;
    for (size_t i = 0; i < n; i += len)
      for (size_t j = 0; j < half; ++j) {
        elt t[2];
        c_mul(F, &A[2 * (i + j + half)], &w[2 * j * step], t);
        elt u0 = A[2 * (i + j)], u1 = A[2 * (i + j) + 1];
        A[2 * (i + j)] = f_add(F, u0, t[0]);
        A[2 * (i + j) + 1] = f_add(F, u1, t[1]);
        A[2 * (i + j + half)] = f_sub(F, u0, t[0]);
        A[2 * (i + j + half) + 1] = f_sub(F, u1, t[1]);
      }
  }
  free(w);
}

/* ------------------------------------------------------------------ */
/* Reed-Solomon extension over prime fields (algebra/reed_solomon.h)   */
/* ------------------------------------------------------------------ */
/* reed_solomon.h:27-41,51-110:
 *   p(k) = (-1)^d (k-d) C(k,d) * sum_{j<=d} (-1)^j C(d,j) p(j) / (k-j),  d=n-1
 * The reference evaluates the inner sum as a convolution with 1/i through
 * FFT (convolution.h:80-91) or a real FFT over Fp2 (convolution.h:156-175);
 * all arithmetic is exact, so the direct sum (small n) or any FFT gives the
 * same values. */
static void fp_rs(const field* F, size_t n, size_t m, elt* y) {
  size_t d = n - 1;
  elt* inv = (elt*)malloc(m * sizeof(elt));
  /* utility.h:51-72 batch_inverse_arithmetic: inv[i] = 1/i */
  {
    elt p = F->one, bi = F->zero;
    inv[0] = F->zero;
    for (size_t i = 1; i < m; ++i) {
      bi = f_add(F, bi, F->one);
      inv[i] = p;
      p = f_mul(F, p, bi);
    }
    p = f_inv(F, p);
    for (size_t i = m; i-- > 1;) {
      inv[i] = f_mul(F, inv[i], p);
      p = f_mul(F, p, bi);
      bi = f_sub(F, bi, F->one);
    }
  }
  elt* lead = (elt*)malloc((m - n + 1) * sizeof(elt));
  elt* x = (elt*)malloc(n * sizeof(elt));
  /* small integers as field elements by repeated addition of one */
  elt* sc = (elt*)malloc(m * sizeof(elt));
  sc[0] = F->zero;
  for (size_t i = 1; i < m; ++i) sc[i] = f_add(F, sc[i - 1], F->one);
  lead[0] = F->one;
  for (size_t i = 1; i + d < m; ++i) lead[i] = f_mul(F, lead[i - 1], f_mul(F, sc[d + i], inv[i]));
  for (size_t k = d; k < m; ++k) {
    lead[k - d] = f_mul(F, lead[k - d], sc[k - d]);
    if (d % 2 == 1) lead[k - d] = f_neg(F, lead[k - d]);
  }
  elt b = F->one;
  x[0] = f_mul(F, b, y[0]);
  for (size_t i = 1; i < n; ++i) {
    b = f_mul(F, b, f_mul(F, sc[n - i], inv[i]));
    x[i] = f_mul(F, (i & 1) ? f_neg(F, b) : b, y[i]);
  }
  size_t P = 1;
  while (P < m) P <<= 1;
  if (n * (m - n) <= ((size_t)1 << 22) || !F->has_omega) {
    for (size_t k = n; k < m; ++k) {
      elt t = F->zero;
      for (size_t i = 0; i < n; ++i) t = f_add(F, t, f_mul(F, x[i], inv[k - i]));
      y[k] = f_mul(F, lead[k - d], t);
    }
  } else if (F->has_omega == 1) {
    elt* X = (elt*)calloc(P, sizeof(elt));
    elt* Y = (elt*)calloc(P, sizeof(elt));
    memcpy(X, x, n * sizeof(elt));
    memcpy(Y, inv, m * sizeof(elt));
    fp_fft(F, X, P, F->omega, F->omega_order, 1);
    fp_fft(F, Y, P, F->omega, F->omega_order, 1);
    for (size_t i = 0; i < P; ++i) X[i] = f_mul(F, X[i], Y[i]);
    fp_fft(F, X, P, F->omega, F->omega_order, 0);
    elt pinv = f_inv(F, f_of_scalar(F, P));
    for (size_t k = n; k < m; ++k) y[k] = f_mul(F, lead[k - d], f_mul(F, X[k], pinv));
    free(X);
    free(Y);
  } else {
    elt* X = (elt*)calloc(2 * P, sizeof(elt));
    elt* Y = (elt*)calloc(2 * P, sizeof(elt));
    for (size_t i = 0; i < n; ++i) X[2 * i] = x[i];
    for (size_t i = 0; i < m; ++i) Y[2 * i] = inv[i];
    fp2_fft(F, X, P, F->omega2, F->omega_order, 1);
    fp2_fft(F, Y, P, F->omega2, F->omega_order, 1);
    for (size_t i = 0; i < P; ++i) c_mul(F, &X[2 * i], &Y[2 * i], &X[2 * i]);
    fp2_fft(F, X, P, F->omega2, F->omega_order, 0);
    elt pinv = f_inv(F, f_of_scalar(F, P));
    for (size_t k = n; k < m; ++k) y[k] = f_mul(F, lead[k - d], f_mul(F, X[2 * k], pinv));
    free(X);
    free(Y);
  }
  free(inv);
  free(lead);
  free(x);
  free(sc);
}

void rs_interpolate(const field* F, size_t n, size_t m, elt* y) {
  if (F->char2) lch14_interpolate(F, n, m, y);
  else fp_rs(F, n, m, y);
}
