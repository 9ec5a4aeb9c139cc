/* oracle/port/orc_zk.c -- circuit parsing, sumcheck layer prover, Ligero
 * commit/prove and the ZK composition + wire format (TEST INFRASTRUCTURE; see
 * orc.h).  Scalar restatement; every block cites the reference lines. */
#include <stdlib.h>
#include <string.h>

#include <stdio.h>
#include "orc.h"

#define KMAXB 40 /* sumcheck/circuit.h:77 kMaxBindings */

/* util/ceildiv.h:36-44 */
static size_t lg(size_t n) {
  size_t r = 0;
  while (n > 1) {
    n = (n / 2) + (n % 2);
    r += 1;
  }
  return r;
}
static size_t ceildiv(size_t a, size_t b) { return (a + b - 1) / b; }

/* ------------------------------------------------------------------ */
/* LigeroParam (ligero/ligero_param.h:171-293)                         */
/* ------------------------------------------------------------------ */
static size_t ligero_layout(ligero_param* p, const field* F, size_t e) {
  const size_t max_lg = 28, max_size = (size_t)1 << max_lg;
  p->block_enc = e;
  size_t subfield_bits = 8 * F->ksubbytes;
  if (subfield_bits <= max_lg && p->block_enc >= ((size_t)1 << subfield_bits)) return SIZE_MAX;
  if (p->block_enc > max_size || p->rateinv > max_size || (p->block_enc + 1) < (2 + p->rateinv))
    return SIZE_MAX;
  p->block = (p->block_enc + 1) / (2 + p->rateinv);
  if (p->block < p->r) return SIZE_MAX;
  p->w = p->block - p->r;
  if (p->w < p->r) return SIZE_MAX;
  p->dblock = 2 * p->block - 1;
  if (p->block_enc < p->dblock) return SIZE_MAX;
  p->block_ext = p->block_enc - p->dblock;
  p->nwrow = ceildiv(p->nw, p->w);
  p->nqtriples = ceildiv(p->nq, p->w);
  p->nwqrow = p->nwrow + 3 * p->nqtriples;
  p->nrow = p->nwqrow + 3;
  if (p->nrow >= max_size / p->block_enc) return SIZE_MAX;
  p->mc_pathlen = merkle_tree_len(p->block_ext);
  uint64_t sz = 32;
  sz += (uint64_t)p->mc_pathlen / 2 * p->nreq * 32;
  sz += (uint64_t)p->block * F->kbytes;
  sz += (uint64_t)p->dblock * F->kbytes;
  sz += (uint64_t)(p->dblock - p->w) * F->kbytes;
  sz += (uint64_t)p->nreq * 32;
  sz += (uint64_t)p->nrow * p->nreq * F->ksubbytes;
  return (size_t)sz;
}
int ligero_param_init(ligero_param* p, const field* F, size_t nw, size_t nq, size_t rateinv,
                      size_t nreq, size_t block_enc) {
  memset(p, 0, sizeof(*p));
  p->nw = nw;
  p->nq = nq;
  p->rateinv = rateinv;
  p->nreq = nreq;
  p->r = nreq;
  if (block_enc == 0) {
    /* ligero_param.h:152-169 */
    size_t best = SIZE_MAX, best_e = 1;
    for (size_t e = 1; e <= ((size_t)1 << 28); e *= 2) {
      size_t s = ligero_layout(p, F, e);
      if (s < best) {
        best = s;
        best_e = e;
      }
    }
    block_enc = best_e;
  }
  if (ligero_layout(p, F, block_enc) == SIZE_MAX) return -1;
  if (!(p->block_enc > p->block)) return -1;
  p->ildt = 0;
  p->idot = 1;
  p->iquad = 2;
  p->iw = 3;
  p->iq = p->iw + p->nwrow;
  return 0;
}

/* ------------------------------------------------------------------ */
/* Circuit: LFC1 (proto/circuit_reader.h:83-252)                        */
/* ------------------------------------------------------------------ */
typedef struct {
  const uint8_t* b;
  size_t n, pos;
  int bad;
} rdbuf;
static size_t rd_num(rdbuf* r) {
  if (r->pos + 3 > r->n) {
    r->bad = 1;
    return 0;
  }
  size_t v = (size_t)r->b[r->pos] | ((size_t)r->b[r->pos + 1] << 8) | ((size_t)r->b[r->pos + 2] << 16);
  r->pos += 3;
  return v;
}
static size_t rd_index(rdbuf* r, size_t prev) {
  size_t d = rd_num(r);
  return (d & 1) ? prev - (d >> 1) : prev + (d >> 1);
}
void circuit_free(circuit* c) {
  if (!c) return;
  if (c->l) {
    for (size_t i = 0; i < c->nl; ++i) {
      free(c->l[i].g);
      free(c->l[i].h0);
      free(c->l[i].h1);
      free(c->l[i].vi);
    }
    free(c->l);
  }
  free(c->consts);
  free(c);
}
circuit* circuit_parse(const field* F, const uint8_t* b, size_t n) {
  rdbuf r = {b, n, 0, 0};
  if (n < 25 || b[0] != 1) return NULL;
  r.pos = 1;
  circuit* c = (circuit*)calloc(1, sizeof(circuit));
  c->F = F;
  size_t fid = rd_num(&r);
  c->nv = rd_num(&r);
  c->nc = rd_num(&r);
  c->npub_in = rd_num(&r);
  c->subfield_boundary = rd_num(&r);
  c->ninputs = rd_num(&r);
  c->nl = rd_num(&r);
  c->nconst = rd_num(&r);
  if (r.bad || fid != (size_t)F->id || c->nv == 0 || c->nc == 0 || c->nl == 0 || c->nl > 10000 ||
      c->npub_in > c->ninputs || c->subfield_boundary > c->ninputs ||
      r.pos + c->nconst * F->kbytes > n) {
    free(c);
    return NULL;
  }
  c->logv = lg(c->nv);
  c->logc = lg(c->nc);
  c->consts = (elt*)calloc(c->nconst ? c->nconst : 1, sizeof(elt));
  for (size_t i = 0; i < c->nconst; ++i) {
    if (f_of_bytes(F, b + r.pos, &c->consts[i])) {
      circuit_free(c);
      return NULL;
    }
    r.pos += F->kbytes;
  }
  c->l = (layer*)calloc(c->nl, sizeof(layer));
  size_t max_g = c->nv;
  for (size_t ly = 0; ly < c->nl; ++ly) {
    layer* L = &c->l[ly];
    L->logw = rd_num(&r);
    L->nw = rd_num(&r);
    L->nterms = rd_num(&r);
    if (r.bad || L->logw == 0 || L->logw > KMAXB || L->nw == 0 || L->nw < L->logw ||
        L->nw > ((size_t)1 << L->logw) || L->nterms == 0 || r.pos + 12 * L->nterms > n) {
      circuit_free(c);
      return NULL;
    }
    L->g = (uint32_t*)malloc(4 * L->nterms);
    L->h0 = (uint32_t*)malloc(4 * L->nterms);
    L->h1 = (uint32_t*)malloc(4 * L->nterms);
    L->vi = (uint32_t*)malloc(4 * L->nterms);
    size_t pg = 0, p0 = 0, p1 = 0;
    for (size_t i = 0; i < L->nterms; ++i) {
      size_t g = rd_index(&r, pg), h0 = rd_index(&r, p0), h1 = rd_index(&r, p1), vi = rd_num(&r);
      if (g >= max_g || h0 >= L->nw || h1 >= L->nw || vi >= c->nconst) {
        circuit_free(c);
        return NULL;
      }
      L->g[i] = (uint32_t)g;
      L->h0[i] = (uint32_t)h0;
      L->h1[i] = (uint32_t)h1;
      L->vi[i] = (uint32_t)vi;
      pg = g;
      p0 = h0;
      p1 = h1;
    }
    c->nterms += L->nterms;
    max_g = L->nw;
  }
  if (r.pos + 32 > n) {
    circuit_free(c);
    return NULL;
  }
  memcpy(c->id, b + r.pos, 32);
  return c;
}

/* sumcheck/circuit_id.h:30-67 */
static void upd8(sha256* s, uint64_t x) {
  uint8_t b[8];
  for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(x >> (8 * i));
  sha256_update(s, b, 8);
}
void circuit_id(const circuit* c, uint8_t id[32]) {
  const field* F = c->F;
  sha256 s;
  uint8_t tmp[32];
  sha256_init(&s);
  if (F->char2) {
    upd8(&s, 2);
    upd8(&s, 128);
  } else {
    upd8(&s, 1);
    f_to_bytes(F, tmp, f_neg(F, F->one));
    sha256_update(&s, tmp, F->kbytes);
  }
  upd8(&s, c->nv);
  upd8(&s, c->logv);
  upd8(&s, c->nc);
  upd8(&s, c->logc);
  upd8(&s, c->nl);
  upd8(&s, c->ninputs);
  upd8(&s, c->npub_in);
  upd8(&s, c->subfield_boundary);
  for (size_t ly = 0; ly < c->nl; ++ly) {
    const layer* L = &c->l[ly];
    upd8(&s, L->nw);
    upd8(&s, L->logw);
    upd8(&s, L->nterms);
    for (size_t i = 0; i < L->nterms; ++i) {
      upd8(&s, L->g[i]);
      upd8(&s, L->h0[i]);
      upd8(&s, L->h1[i]);
      f_to_bytes(F, tmp, c->consts[L->vi[i]]);
      sha256_update(&s, tmp, F->kbytes);
    }
  }
  sha256_final(&s, id);
}

/* ------------------------------------------------------------------ */
/* small helpers restating arrays/, algebra/poly.h                     */
/* ------------------------------------------------------------------ */
/* arrays/affine.h:25-52 */
static elt affine(const field* F, elt r, elt f0, elt f1) {
  return f_add(F, f0, f_mul(F, f_sub(F, f1, f0), r));
}
static elt affine_z_nz(const field* F, elt r, elt f1) { return f_mul(F, f1, r); }
static elt affine_nz_z(const field* F, elt r, elt f0) { return f_sub(F, f0, f_mul(F, f0, r)); }

/* arrays/dense.h:70-89 (n1 == 1) */
static size_t dense_bind(const field* F, elt* out, const elt* in, size_t n0, elt r) {
  size_t i0 = 0, rd = 0, wr = 0;
  while (2 * i0 + 1 < n0) {
    out[wr] = affine(F, r, in[rd], in[rd + 1]);
    i0++, rd += 2, wr += 1;
  }
  if (2 * i0 < n0) {
    out[wr] = affine_nz_z(F, r, in[rd]);
    wr++;
  }
  return (n0 + 1) / 2;
}

/* arrays/eqs.h:46-78 raw_eq2 / fill_recursive */
static void fill_recursive(const field* F, elt* eq, size_t l, size_t n, const elt* G0, const elt* G1,
                           elt w0, elt w1) {
  if (l > 0) {
    size_t nl = l - 1, s = (size_t)1 << nl;
    elt w0hi = f_mul(F, w0, G0[nl]), w1hi = f_mul(F, w1, G1[nl]);
    elt w0lo = f_sub(F, w0, w0hi), w1lo = f_sub(F, w1, w1hi);
    if (n <= s) {
      fill_recursive(F, eq, nl, n, G0, G1, w0lo, w1lo);
    } else {
      fill_recursive(F, eq, nl, s, G0, G1, w0lo, w1lo);
      fill_recursive(F, eq + s, nl, n - s, G0, G1, w0hi, w1hi);
    }
  } else {
    eq[0] = f_add(F, w0, w1);
  }
}
/* arrays/eqs.h:104-134 filleq */
static size_t ceilshr(size_t a, size_t n) { return 1u + ((a - 1u) >> n); }
static void filleq(const field* F, elt* eq, size_t logn, size_t n, const elt* Q) {
  eq[0] = F->one;
  for (size_t l = logn; l-- > 0;) {
    size_t nl = ceilshr(n, l), i = ceilshr(nl, 1);
    if (2 * i - 1 >= nl) {
      i--;
      elt v = eq[i], qv = f_mul(F, Q[l], v);
      eq[2 * i] = f_sub(F, v, qv);
    }
    while (i-- > 0) {
      elt v = eq[i], qv = f_mul(F, Q[l], v);
      eq[2 * i] = f_sub(F, v, qv);
      eq[2 * i + 1] = qv;
    }
  }
}

/* algebra/poly.h:59-98 for N = 3 */
static void newton_of_lagrange3(const field* F, elt t[3]) {
  for (size_t i = 1; i < 3; i++)
    for (size_t k = 3; k-- > i;) t[k] = f_mul(F, f_sub(F, t[k], t[k - 1]), F->newton[k][i]);
}
static elt eval_newton3(const field* F, const elt t[3], elt x) {
  elt e = t[2];
  for (size_t i = 2; i-- > 0;) e = f_add(F, f_mul(F, e, f_sub(F, x, F->evalpt[i])), t[i]);
  return e;
}
static elt eval_lagrange3(const field* F, const elt t[3], elt x) {
  elt tmp[3] = {t[0], t[1], t[2]};
  newton_of_lagrange3(F, tmp);
  return eval_newton3(F, tmp, x);
}
static elt eval_monomial3(const field* F, const elt t[3], elt x) {
  elt e = t[2];
  for (size_t i = 2; i-- > 0;) e = f_add(F, f_mul(F, e, x), t[i]);
  return e;
}

/* ------------------------------------------------------------------ */
/* sumcheck proof storage (sumcheck/circuit.h:68-93, logc == 0)        */
/* ------------------------------------------------------------------ */
typedef struct {
  elt hp[2][KMAXB][3];
  elt wc[2];
} layer_proof;

/* ------------------------------------------------------------------ */
/* ProverLayers::eval_circuit (sumcheck/prover_layers.h:52-98,278-305)  */
/* ------------------------------------------------------------------ */
static int eval_quad(const circuit* c, const layer* L, elt* V, size_t nvout, const elt* W) {
  const field* F = c->F;
  for (size_t i = 0; i < nvout; ++i) V[i] = F->zero;
  for (size_t i = 0; i < L->nterms; ++i) {
    size_t g = L->g[i], r = L->h0[i], l = L->h1[i];
    elt v = c->consts[L->vi[i]];
    if (f_is_zero(F, v)) {
      elt y = f_mul(F, W[l], W[r]);
      if (!f_is_zero(F, y)) return 0;
    } else {
      elt x = f_mul(F, f_mul(F, v, W[l]), W[r]);
      V[g] = f_add(F, V[g], x);
    }
  }
  return 1;
}

/* ------------------------------------------------------------------ */
/* HQuad (sumcheck/hquad.h) and Quad::bind_g (sumcheck/quad.h:152-185)  */
/* ------------------------------------------------------------------ */
typedef struct {
  size_t n;
  uint32_t (*h)[2];
  elt* v;
} hquad;

static void bind_g(const circuit* c, const layer* L, size_t logv, const elt* G0, const elt* G1,
                   elt alpha, elt beta, hquad* s) {
  const field* F = c->F;
  size_t nv = (size_t)1 << logv;
  elt* dot = (elt*)malloc(nv * sizeof(elt));
  fill_recursive(F, dot, logv, nv, G0, G1, F->one, alpha);
  s->h = malloc(L->nterms * sizeof(*s->h));
  s->v = (elt*)malloc(L->nterms * sizeof(elt));
  size_t wr = 0;
  for (size_t i = 0; i < L->nterms; ++i) {
    elt v = c->consts[L->vi[i]];
    /* quad.h:213-220 prep_v */
    elt vc = f_mul(F, f_is_zero(F, v) ? beta : v, dot[L->g[i]]);
    if (wr > 0 && s->h[wr - 1][0] == L->h0[i] && s->h[wr - 1][1] == L->h1[i]) {
      s->v[wr - 1] = f_add(F, s->v[wr - 1], vc);
    } else {
      s->h[wr][0] = L->h0[i];
      s->h[wr][1] = L->h1[i];
      s->v[wr] = vc;
      ++wr;
    }
  }
  s->n = wr;
  free(dot);
}
/* hquad.h:89-123 */
static void bind_h(const field* F, hquad* q, elt r, size_t hand) {
  size_t rd = 0, wr = 0, o = 1 - hand;
  while (rd < q->n) {
    uint32_t hh = q->h[rd][hand] >> 1, ho = q->h[rd][o];
    elt vcc;
    size_t rd1 = rd + 1;
    if (rd1 < q->n && q->h[rd][o] == q->h[rd1][o] && (q->h[rd][hand] >> 1) == (q->h[rd1][hand] >> 1) &&
        q->h[rd1][hand] == q->h[rd][hand] + 1) {
      vcc = affine(F, r, q->v[rd], q->v[rd1]);
      rd += 2;
    } else {
      if ((q->h[rd][hand] & 1) == 0) vcc = affine_nz_z(F, r, q->v[rd]);
      else vcc = affine_z_nz(F, r, q->v[rd]);
      rd = rd1;
    }
    q->h[wr][hand] = hh;
    q->h[wr][o] = ho;
    q->v[wr] = vcc;
    ++wr;
  }
  q->n = wr;
}

/* prover_layers.h:357-402 */
static void evaluations(const field* F, size_t n, elt eq0, const elt* QW, const elt* W, elt sum,
                        elt evals[3]) {
  size_t nodd = n / 2;
  elt a0 = F->zero, a2 = F->zero;
  for (size_t i = 0; i < nodd; i++) {
    a0 = f_add(F, a0, f_mul(F, QW[2 * i], W[2 * i]));
    a2 = f_add(F, a2, f_mul(F, f_sub(F, QW[2 * i + 1], QW[2 * i]), f_sub(F, W[2 * i + 1], W[2 * i])));
  }
  if (2 * nodd < n) {
    elt p = f_mul(F, QW[2 * nodd], W[2 * nodd]);
    a0 = f_add(F, a0, p);
    a2 = f_add(F, a2, p);
  }
  elt coef[3];
  coef[0] = f_mul(F, eq0, a0);
  coef[2] = f_mul(F, eq0, a2);
  coef[1] = f_sub(F, f_sub(F, f_sub(F, sum, coef[0]), coef[0]), coef[2]);
  for (int k = 0; k < 3; ++k) evals[k] = eval_monomial3(F, coef, F->evalpt[k]);
}

/* TranscriptSumcheck::round (sumcheck/transcript_sumcheck.h:63-79) */
static elt ts_round(transcript* t, const field* F, const elt poly[3]) {
  ts_write_elt(t, F, poly[0]);
  ts_write_elt(t, F, poly[2]);
  return rng_elt(&t->base, F);
}

typedef struct {
  size_t logv;
  elt q[KMAXB];
  elt g[2][KMAXB];
} bindings;

/* prover_layers.h:185-271 (logc == 0) */
static int prove_layer(const circuit* c, size_t ly, layer_proof* pr, const layer_proof* pad,
                       transcript* ts, bindings* bnd, hquad* Q, elt* W /* destroyed */, elt alpha,
                       elt WC[2]) {
  const field* F = c->F;
  const layer* L = &c->l[ly];
  size_t logw = L->logw;
  bnd->logv = logw;
  elt sum = f_add(F, WC[0], f_mul(F, alpha, WC[1]));
  elt eq0 = F->one; /* Eqs(logc=0, nc=1): eqs.h:104-106 */
  size_t n[2] = {L->nw, L->nw};
  elt* WH[2];
  WH[0] = (elt*)malloc(L->nw * sizeof(elt));
  WH[1] = W;
  memcpy(WH[0], W, L->nw * sizeof(elt));
  elt* QW = (elt*)malloc(L->nw * sizeof(elt));
  for (size_t round = 0; round < logw; ++round) {
    for (size_t hand = 0; hand < 2; hand++) {
      size_t o = 1 - hand;
      for (size_t i = 0; i < n[hand]; ++i) QW[i] = F->zero;
      for (size_t i = 0; i < Q->n; ++i) {
        size_t p0 = Q->h[i][hand], p1 = Q->h[i][o];
        QW[p0] = f_add(F, QW[p0], f_mul(F, Q->v[i], WH[o][p1]));
      }
      elt evals[3], poly[3];
      evaluations(F, n[hand], eq0, QW, WH[hand], sum, evals);
      /* round_h, prover_layers.h:320-329 */
      for (int k = 0; k < 3; ++k) poly[k] = f_sub(F, evals[k], pad->hp[hand][round][k]);
      for (int k = 0; k < 3; ++k) pr->hp[hand][round][k] = poly[k];
      elt rnd = ts_round(ts, F, poly);
      bnd->g[hand][round] = rnd;
      sum = eval_lagrange3(F, evals, rnd);
      n[hand] = dense_bind(F, WH[hand], WH[hand], n[hand], rnd);
      bind_h(F, Q, rnd, hand);
    }
  }
  int ok = (Q->n == 1 && Q->h[0][0] == 0 && Q->h[0][1] == 0);
  WC[0] = WH[0][0];
  WC[1] = WH[1][0];
  elt expect = f_mul(F, eq0, f_mul(F, Q->v[0], f_mul(F, WC[0], WC[1])));
  ok = ok && f_eq(F, sum, expect);
  /* end_layer, prover_layers.h:331-344 */
  elt tt[2] = {f_sub(F, WC[0], pad->wc[0]), f_sub(F, WC[1], pad->wc[1])};
  pr->wc[0] = tt[0];
  pr->wc[1] = tt[1];
  ts_write_array(ts, F, tt, 1, 2);
  free(WH[0]);
  free(QW);
  return ok;
}

/* ------------------------------------------------------------------ */
/* ZkCommon::verifier_constraints (zk/zk_common.h:49-136,291-439)       */
/* ------------------------------------------------------------------ */
typedef struct {
  size_t c, w;
  elt k;
} llc;
typedef struct {
  llc* a;
  size_t n, cap;
} llvec;
static void ll_push(llvec* v, size_t c, size_t w, elt k) {
  if (v->n == v->cap) {
    v->cap = v->cap ? 2 * v->cap : 1024;
    v->a = (llc*)realloc(v->a, v->cap * sizeof(llc));
  }
  v->a[v->n].c = c;
  v->a[v->n].w = w;
  v->a[v->n].k = k;
  v->n++;
}

/* Poly<3>::dot_interpolation (algebra/poly.h:125-150) */
static void lagrange_coef3(const field* F, elt x, elt lag[3]) {
  for (int k = 0; k < 3; ++k) {
    elt id[3] = {F->zero, F->zero, F->zero};
    id[k] = F->one;
    newton_of_lagrange3(F, id);
    lag[k] = eval_newton3(F, id, x);
  }
}

static size_t verifier_constraints(const circuit* c, const elt* pub, const layer_proof* proof,
                                   const elt* bound_quad, llvec* a, transcript* tsv, size_t pi) {
  const field* F = c->F;
  elt q[KMAXB], g[KMAXB];
  for (int i = 0; i < KMAXB; ++i) q[i] = rng_elt(&tsv->base, F);
  for (int i = 0; i < KMAXB; ++i) g[i] = rng_elt(&tsv->base, F);
  (void)q;
  size_t cla_logv = c->logv;
  elt claim[2] = {F->zero, F->zero};
  elt hb[2][KMAXB];
  const elt* cg[2] = {g, g};
  elt hb_prev[2][KMAXB];
  size_t ci = 0;
  for (size_t ly = 0; ly < c->nl; ++ly) {
    const layer* L = &c->l[ly];
    const layer_proof* plr = &proof[ly];
    elt alpha = rng_elt(&tsv->base, F);
    elt beta = rng_elt(&tsv->base, F);
    (void)beta;
    size_t logw = L->logw;
    /* PadLayout (zk_common.h:193-243): ovp indices */
    size_t nvar = 3 + 4 * logw + 3;
    elt known = F->zero;
    elt* sym = (elt*)calloc(nvar, sizeof(elt));
    for (size_t i = 0; i < nvar; ++i) sym[i] = F->zero;
    /* cb.first (zk_common.h:330-335) */
    known = f_add(F, known, f_mul(F, F->one, claim[0]));
    sym[0] = f_add(F, sym[0], F->one);
    known = f_add(F, known, f_mul(F, alpha, claim[1]));
    sym[1] = f_add(F, sym[1], alpha);
    for (size_t round = 0; round < logw; ++round) {
      for (size_t hand = 0; hand < 2; ++hand) {
        size_t r = 2 * round + hand;
        const elt* hp = plr->hp[hand][round];
        hb[hand][round] = ts_round(tsv, F, hp);
        elt lag[3];
        lagrange_coef3(F, hb[hand][round], lag);
        /* cb.next (zk_common.h:338-351) */
        size_t i0 = 3 + 2 * r, i2 = 3 + 2 * r + 1;
        known = f_sub(F, known, f_mul(F, F->one, hp[0]));
        sym[i0] = f_sub(F, sym[i0], F->one);
        known = f_mul(F, known, lag[1]);
        for (size_t i = 0; i < nvar; ++i) sym[i] = f_mul(F, sym[i], lag[1]);
        known = f_add(F, known, f_mul(F, lag[0], hp[0]));
        sym[i0] = f_add(F, sym[i0], lag[0]);
        known = f_add(F, known, f_mul(F, lag[2], hp[2]));
        sym[i2] = f_add(F, sym[i2], lag[2]);
      }
    }
    elt quad = bound_quad[ly];
    elt eqq = f_mul(F, F->one /* Eq::eval(logc=0) eq.h:55-75 */, quad);
    /* cb.finalize (zk_common.h:374-399) */
    size_t cp = 3 + 4 * logw;
    sym[cp + 0] = f_sub(F, sym[cp + 0], f_mul(F, eqq, plr->wc[1]));
    sym[cp + 1] = f_sub(F, sym[cp + 1], f_mul(F, eqq, plr->wc[0]));
    sym[cp + 2] = f_sub(F, sym[cp + 2], eqq);
    size_t i0 = (ly == 0) ? 3 : 0;
    for (size_t i = i0; i < nvar; ++i) ll_push(a, ci, (pi + i) - 3, sym[i]);
    ci++;
    free(sym);
    ts_write_array(tsv, F, plr->wc, 1, 2);
    cla_logv = logw;
    claim[0] = plr->wc[0];
    claim[1] = plr->wc[1];
    memcpy(hb_prev, hb, sizeof(hb));
    cg[0] = hb_prev[0];
    cg[1] = hb_prev[1];
    pi += 4 * logw + 3;
  }
  /* input constraint (zk_common.h:119-135,406-439) */
  elt alpha = rng_elt(&tsv->base, F);
  size_t ninp = c->ninputs, npub = c->npub_in;
  elt* eq0 = (elt*)malloc(ninp * sizeof(elt));
  elt* eq1 = (elt*)malloc(ninp * sizeof(elt));
  filleq(F, eq0, cla_logv, ninp, cg[0]);
  filleq(F, eq1, cla_logv, ninp, cg[1]);
  for (size_t i = 0; i < ninp; ++i) {
    elt b_i = f_add(F, eq0[i], f_mul(F, alpha, eq1[i]));
    if (i >= npub) ll_push(a, ci, i - npub, b_i);
  }
  (void)pub; /* the public binding only enters b, which the prover does not need */
  size_t m1 = pi - 3;
  ll_push(a, ci, m1 + 0, f_neg(F, F->one));
  ll_push(a, ci, m1 + 1, f_neg(F, alpha));
  free(eq0);
  free(eq1);
  return ++ci;
}

/* ------------------------------------------------------------------ */
/* ZK prover                                                           */
/* ------------------------------------------------------------------ */
typedef struct {
  uint8_t* p;
  size_t n, cap;
  int overflow;
} outbuf;
static void ob_put(outbuf* o, const uint8_t* b, size_t n) {
  if (o->n + n > o->cap) {
    o->overflow = 1;
    o->n += n;
    return;
  }
  memcpy(o->p + o->n, b, n);
  o->n += n;
}
static void ob_elt(outbuf* o, const field* F, elt e) {
  uint8_t t[32];
  f_to_bytes(F, t, e);
  ob_put(o, t, F->kbytes);
}
static void ob_u32(outbuf* o, size_t g) {
  uint8_t t[4] = {(uint8_t)g, (uint8_t)(g >> 8), (uint8_t)(g >> 16), (uint8_t)(g >> 24)};
  ob_put(o, t, 4);
}

/* G != NULL: LigeroProver::commit + ::prove alone (ligero/ligero_prover.h:58-146) on a caller-given
 * statement -- witness vector, quadratic constraints, linear constraint terms, hash of the statement --
 * with no circuit and no sumcheck in front; the output is the commitment root (32 bytes) followed by
 * LigeroProof's serialisation (what ZkProof::write_com_proof emits, zk_proof.h:137-184).  This is the shape
 * of the reference's C++-generated Ligero vector (rust/runtime/ligero/tests/ligero.rs:592-760). */
static int zk_prove_ex(const circuit* c, const ligero_generic* G, const uint8_t* wbytes, rng* rg,
                       const uint8_t* tinit, size_t tinit_len, size_t rate, size_t nreq, size_t block_enc,
                       uint8_t* out, size_t out_cap, size_t* out_len, zk_dump* dump) {
  const field* F = G ? G->F : c->F;
  int rc = 0;
  if (!G && c->logc != 0) return -10;
  size_t ninp = G ? 0 : c->ninputs, npub = G ? 0 : c->npub_in, nl = G ? 0 : c->nl;
  size_t n_witness = G ? G->nw : ninp - npub;
  size_t pad_size = 0;
  for (size_t i = 0; i < nl; ++i) pad_size += 4 * c->l[i].logw + 3; /* zk_common.h:139-146 */
  ligero_param P;
  if (ligero_param_init(&P, F, n_witness + pad_size, G ? G->nq : nl, rate, nreq, block_enc)) return -11;

  elt* W = (elt*)malloc(ninp * sizeof(elt));
  for (size_t i = 0; i < ninp; ++i)
    if (f_of_bytes(F, wbytes + i * F->kbytes, &W[i])) {
      free(W);
      return -2;
    }

  /* ---- ZkProver::commit (zk/zk_prover.h:72-100) ---- */
  elt* wit = (elt*)malloc(P.nw * sizeof(elt));
  for (size_t i = 0; i < n_witness; ++i) wit[i] = G ? G->W[i] : W[i + npub];
  size_t sb = G ? G->subfield_boundary : (c->subfield_boundary >= npub ? c->subfield_boundary - npub : 0);
  /* fill_pad (zk_prover.h:152-188) */
  layer_proof* pad = (layer_proof*)calloc(nl, sizeof(layer_proof));
  layer_proof* proof = (layer_proof*)calloc(nl, sizeof(layer_proof));
  size_t wp = n_witness;
  for (size_t i = 0; i < nl; ++i) {
    for (size_t j = 0; j < c->l[i].logw; ++j)
      for (size_t h = 0; h < 2; ++h)
        for (size_t k = 0; k < 3; ++k) {
          if (k != 1) {
            elt r = rng_elt(rg, F);
            pad[i].hp[h][j][k] = r;
            wit[wp++] = r;
          } else {
            pad[i].hp[h][j][k] = F->zero;
          }
        }
    for (size_t k = 0; k < 2; ++k) {
      elt r = rng_elt(rg, F);
      pad[i].wc[k] = r;
      wit[wp++] = r;
    }
    wit[wp++] = f_mul(F, pad[i].wc[0], pad[i].wc[1]);
  }
  /* setup_lqc (zk_common.h:149-160) */
  size_t(*lqc)[3] = malloc((P.nq ? P.nq : 1) * sizeof(*lqc));
  if (G) {
    for (size_t i = 0; i < G->nq; ++i)
      for (int k = 0; k < 3; ++k) lqc[i][k] = G->lqc[3 * i + k];
  } else {
    size_t pi = n_witness;
    for (size_t i = 0; i < nl; ++i) {
      size_t cp = 4 * c->l[i].logw;
      lqc[i][0] = pi + cp;
      lqc[i][1] = pi + cp + 1;
      lqc[i][2] = pi + cp + 2;
      pi += cp + 3;
    }
  }
  /* ---- LigeroProver::commit (ligero/ligero_prover.h:58-79,171-279) ---- */
  size_t ld = P.block_enc;
  elt* T = (elt*)calloc(P.nrow * ld, sizeof(elt));
#define TAB(i, j) T[(i) * ld + (j)]
  for (size_t i = 0; i < sb; ++i)
    if (!f_in_subfield(F, wit[i])) rc = -12;
  /* layout_blinding_rows :171-205 */
  for (size_t j = 0; j < P.block; ++j) TAB(P.ildt, j) = rng_elt(rg, F);
  rs_interpolate(F, P.block, P.block_enc, &TAB(P.ildt, 0));
  for (size_t j = 0; j < P.dblock; ++j) TAB(P.idot, j) = rng_elt(rg, F);
  {
    elt s = F->zero;
    for (size_t j = 0; j < P.w; ++j) s = f_add(F, s, TAB(P.idot, P.r + j));
    TAB(P.idot, P.r) = f_sub(F, TAB(P.idot, P.r), s);
  }
  rs_interpolate(F, P.dblock, P.block_enc, &TAB(P.idot, 0));
  for (size_t j = 0; j < P.dblock; ++j) TAB(P.iquad, j) = rng_elt(rg, F);
  for (size_t j = 0; j < P.w; ++j) TAB(P.iquad, P.r + j) = F->zero;
  rs_interpolate(F, P.dblock, P.block_enc, &TAB(P.iquad, 0));
  /* layout_witness_rows :207-231 */
  for (size_t i = 0; i < P.nwrow; ++i) {
    int sub_only = ((i + 1) * P.w <= sb);
    for (size_t j = 0; j < P.r; ++j) TAB(i + P.iw, j) = sub_only ? rng_subfield_elt(rg, F) : rng_elt(rg, F);
    for (size_t j = 0; j < P.w; ++j) TAB(i + P.iw, P.r + j) = F->zero;
    size_t mx = P.w < P.nw - i * P.w ? P.w : P.nw - i * P.w;
    for (size_t j = 0; j < mx; ++j) TAB(i + P.iw, P.r + j) = wit[i * P.w + j];
    rs_interpolate(F, P.block, P.block_enc, &TAB(i + P.iw, 0));
  }
  /* layout_quadratic_rows :233-270 */
  {
    size_t iqx = P.iq, iqy = iqx + P.nqtriples, iqz = iqy + P.nqtriples;
    for (size_t i = 0; i < P.nqtriples; ++i) {
      for (size_t j = 0; j < P.r; ++j) TAB(iqx + i, j) = rng_elt(rg, F);
      for (size_t j = 0; j < P.r; ++j) TAB(iqy + i, j) = rng_elt(rg, F);
      for (size_t j = 0; j < P.r; ++j) TAB(iqz + i, j) = rng_elt(rg, F);
      for (size_t j = 0; j < P.w; ++j) TAB(iqx + i, P.r + j) = TAB(iqy + i, P.r + j) = TAB(iqz + i, P.r + j) = F->zero;
      for (size_t j = 0; j < P.w && j + i * P.w < P.nq; ++j) {
        const size_t* l = lqc[j + i * P.w];
        TAB(iqx + i, j + P.r) = wit[l[0]];
        TAB(iqy + i, j + P.r) = wit[l[1]];
        TAB(iqz + i, j + P.r) = wit[l[2]];
      }
      rs_interpolate(F, P.block, P.block_enc, &TAB(iqx + i, 0));
      rs_interpolate(F, P.block, P.block_enc, &TAB(iqy + i, 0));
      rs_interpolate(F, P.block, P.block_enc, &TAB(iqz + i, 0));
    }
  }
  /* MerkleCommitment::commit (merkle/merkle_commitment.h:50-64) with
   * LigeroCommon::column_hash (ligero_param.h:432-439) */
  size_t nleaf = P.block_ext;
  uint8_t* nodes = (uint8_t*)calloc(2 * nleaf, 32);
  uint8_t* nonces = (uint8_t*)malloc(nleaf * 32);
  for (size_t j = 0; j < nleaf; ++j) {
    sha256 s;
    uint8_t buf[32];
    sha256_init(&s);
    rg->bytes(rg, nonces + 32 * j, 32);
    sha256_update(&s, nonces + 32 * j, 32);
    for (size_t i = 0; i < P.nrow; ++i) {
      f_to_bytes(F, buf, TAB(i, j + P.dblock));
      sha256_update(&s, buf, F->kbytes);
    }
    sha256_final(&s, nodes + 32 * (nleaf + j));
  }
  merkle_build(nleaf, nodes);
  const uint8_t* root = nodes + 32;
  transcript tp;
  ts_init(&tp, tinit, tinit_len);
  ts_write_bytes(&tp, root, 32); /* ligero_transcript.h:31-34 */

  if (dump) {
    if (dump->witness && dump->witness_cap >= P.nw * F->kbytes)
      for (size_t i = 0; i < P.nw; ++i) f_to_bytes(F, dump->witness + i * F->kbytes, wit[i]);
    if (dump->tableau && dump->tableau_cap >= P.nrow * ld * F->kbytes)
      for (size_t i = 0; i < P.nrow * ld; ++i) f_to_bytes(F, dump->tableau + i * F->kbytes, T[i]);
    if (dump->root) memcpy(dump->root, root, 32);
  }

  /* ---- ZkProver::prove (zk/zk_prover.h:102-149) ---- */
  transcript tst;
  elt** in = (elt**)calloc(nl ? nl : 1, sizeof(elt*));
  elt* finalV = NULL;
  if (!G) {
  /* initialize_sumcheck_fiat_shamir (zk_common.h:163-180) */
  ts_write_bytes(&tp, c->id, 32);
  for (size_t i = 0; i < npub; ++i) ts_write_elt(&tp, F, W[i]);
  ts_write_elt(&tp, F, F->zero);
  ts_write0(&tp, c->nterms);
  tst = tp; /* clone(): transcript.h:86 copies only the hash */
  tst.have_prf = 0;

  /* eval_circuit */
  in[nl - 1] = (elt*)malloc(ninp * sizeof(elt));
  memcpy(in[nl - 1], W, ninp * sizeof(elt));
  finalV = (elt*)malloc(c->nv * sizeof(elt));
  for (size_t l = nl; l-- > 0;) {
    elt* V;
    size_t nvout;
    if (l > 0) {
      nvout = c->l[l - 1].nw;
      in[l - 1] = (elt*)malloc(nvout * sizeof(elt));
      V = in[l - 1];
    } else {
      nvout = c->nv;
      V = finalV;
    }
    if (!eval_quad(c, &c->l[l], V, nvout, in[l])) rc = -3;
  }
  if (getenv("ORC_WIRE_STATS"))
    for (size_t l = 0; l < nl; ++l) {
      size_t nb = 0, nwl = c->l[l].nw;
      for (size_t i = 0; i < nwl; ++i)
        if (f_is_zero(F, in[l][i]) || memcmp(&in[l][i], &F->one, sizeof(elt)) == 0) ++nb;
      fprintf(stderr, "layer %zu nw %zu bits %zu nterms %zu\n", l, nwl, nb, (size_t)c->l[l].nterms);
    }
  for (size_t i = 0; i < c->nv; ++i)
    if (!f_is_zero(F, finalV[i])) rc = -3;
  }  /* !G */

  elt* bound_quad = (elt*)calloc(nl ? nl : 1, sizeof(elt));
  llvec A = {0, 0, 0};
  elt *y_ldt = NULL, *y_dot = NULL, *y_q = NULL, *Avec = NULL;
  size_t* idx = NULL;
  uint8_t* path = NULL;
  size_t pathlen = 0;
  outbuf ob = {out, 0, out_cap, 0};
  if (rc == 0) {
    size_t ncons = 0;
    uint8_t hashA[32] = {0xde, 0xad, 0xbe, 0xef};
    if (G) {
      for (size_t l = 0; l < G->nterms; ++l) ll_push(&A, G->term_c[l], G->term_w[l], G->term_k[l]);
      ncons = G->ncons;
      memcpy(hashA, G->hash, 32);
    } else {
    /* ProverLayers::prove (prover_layers.h:114-166) */
    bindings bnd;
    bnd.logv = c->logv;
    for (int i = 0; i < KMAXB; ++i) bnd.q[i] = rng_elt(&tst.base, F);
    for (int i = 0; i < KMAXB; ++i) bnd.g[0][i] = rng_elt(&tst.base, F);
    for (size_t i = 0; i < bnd.logv; ++i) bnd.g[1][i] = bnd.g[0][i];
    elt WC[2] = {F->zero, F->zero};
    for (size_t ly = 0; ly < nl; ++ly) {
      elt alpha = rng_elt(&tst.base, F);
      elt beta = rng_elt(&tst.base, F);
      hquad Q;
      bind_g(c, &c->l[ly], bnd.logv, bnd.g[0], bnd.g[1], alpha, beta, &Q);
      if (!prove_layer(c, ly, &proof[ly], &pad[ly], &tst, &bnd, &Q, in[ly], alpha, WC)) rc = -13;
      bound_quad[ly] = Q.v[0];
      free(Q.h);
      free(Q.v);
    }
    ncons = verifier_constraints(c, W, proof, bound_quad, &A, &tp, n_witness);
    }  /* !G */

    /* ---- LigeroProver::prove (ligero_prover.h:84-146) ---- */
    ts_write_bytes(&tp, hashA, 32);
    /* low_degree_proof :281-291 */
    y_ldt = (elt*)malloc(P.block * sizeof(elt));
    {
      elt* u = (elt*)malloc(P.nwqrow * sizeof(elt));
      for (size_t i = 0; i < P.nwqrow; ++i) u[i] = rng_elt(&tp.base, F);
      for (size_t j = 0; j < P.block; ++j) y_ldt[j] = TAB(P.ildt, j);
      for (size_t i = 0; i < P.nwqrow; ++i)
        for (size_t j = 0; j < P.block; ++j)
          y_ldt[j] = f_add(F, y_ldt[j], f_mul(F, TAB(i + P.iw, j), u[i]));
      free(u);
    }
    /* gen_alphal, gen_alphaq, inner_product_vector (ligero_param.h:382-421) */
    {
      elt* alphal = (elt*)malloc(ncons * sizeof(elt));
      elt* alphaq = (elt*)malloc(3 * P.nq * sizeof(elt));
      for (size_t i = 0; i < ncons; ++i) alphal[i] = rng_elt(&tp.base, F);
      for (size_t i = 0; i < 3 * P.nq; ++i) alphaq[i] = rng_elt(&tp.base, F);
      Avec = (elt*)calloc(P.nwqrow * P.w, sizeof(elt));
      for (size_t i = 0; i < P.nwqrow * P.w; ++i) Avec[i] = F->zero;
      for (size_t l = 0; l < A.n; ++l)
        Avec[A.a[l].w] = f_add(F, Avec[A.a[l].w], f_mul(F, A.a[l].k, alphal[A.a[l].c]));
      elt* Ax = &Avec[P.nwrow * P.w];
      elt* Ay = Ax + P.nqtriples * P.w;
      elt* Az = Ay + P.nqtriples * P.w;
      for (size_t i = 0; i < P.nqtriples; ++i)
        for (size_t j = 0; j < P.w && j + i * P.w < P.nq; ++j) {
          size_t iw = j + i * P.w;
          const size_t* l = lqc[iw];
          Ax[iw] = f_add(F, Ax[iw], alphaq[3 * iw + 0]);
          Avec[l[0]] = f_sub(F, Avec[l[0]], alphaq[3 * iw + 0]);
          Ay[iw] = f_add(F, Ay[iw], alphaq[3 * iw + 1]);
          Avec[l[1]] = f_sub(F, Avec[l[1]], alphaq[3 * iw + 1]);
          Az[iw] = f_add(F, Az[iw], alphaq[3 * iw + 2]);
          Avec[l[2]] = f_sub(F, Avec[l[2]], alphaq[3 * iw + 2]);
        }
      free(alphal);
      free(alphaq);
    }
    /* dot_proof :293-309 */
    y_dot = (elt*)malloc(P.dblock * sizeof(elt));
    {
      elt* Aext = (elt*)malloc(P.dblock * sizeof(elt));
      for (size_t j = 0; j < P.dblock; ++j) y_dot[j] = TAB(P.idot, j);
      for (size_t i = 0; i < P.nwqrow; ++i) {
        for (size_t j = 0; j < P.r; ++j) Aext[j] = F->zero;
        for (size_t j = 0; j < P.w; ++j) Aext[P.r + j] = Avec[i * P.w + j];
        rs_interpolate(F, P.block, P.dblock, Aext);
        for (size_t j = 0; j < P.dblock; ++j)
          y_dot[j] = f_add(F, y_dot[j], f_mul(F, TAB(i + P.iw, j), Aext[j]));
      }
      free(Aext);
    }
    /* quadratic_proof :311-344 */
    y_q = (elt*)malloc(P.dblock * sizeof(elt));
    {
      elt* u = (elt*)malloc((P.nqtriples ? P.nqtriples : 1) * sizeof(elt));
      for (size_t i = 0; i < P.nqtriples; ++i) u[i] = rng_elt(&tp.base, F);
      size_t iqx = P.iq, iqy = iqx + P.nqtriples, iqz = iqy + P.nqtriples;
      for (size_t j = 0; j < P.dblock; ++j) y_q[j] = TAB(P.iquad, j);
      for (size_t i = 0; i < P.nqtriples; ++i)
        for (size_t j = 0; j < P.dblock; ++j) {
          elt tmp = f_sub(F, TAB(iqz + i, j), f_mul(F, TAB(iqy + i, j), TAB(iqx + i, j)));
          y_q[j] = f_add(F, y_q[j], f_mul(F, tmp, u[i]));
        }
      for (size_t j = 0; j < P.w; ++j)
        if (!f_is_zero(F, y_q[P.r + j])) rc = -14;
      free(u);
    }
    ts_write_array(&tp, F, y_ldt, 1, P.block);
    ts_write_array(&tp, F, y_dot, 1, P.dblock);
    ts_write_array(&tp, F, y_q, 1, P.r);
    ts_write_array(&tp, F, y_q + P.block, 1, P.dblock - P.block);
    /* gen_idx (ligero_transcript.h:64-69) */
    idx = (size_t*)malloc(P.nreq * sizeof(size_t));
    rng_choose(&tp.base, idx, P.block_ext, P.nreq);
    path = (uint8_t*)malloc(P.nreq * P.mc_pathlen * 32);
    pathlen = merkle_open(nleaf, nodes, idx, P.nreq, path);

    /* ---- ZkProof::write (zk/zk_proof.h:90-184) ---- */
    ob_put(&ob, root, 32);
    for (size_t i = 0; i < nl; ++i) {
      for (size_t wi = 0; wi < c->l[i].logw; ++wi)
        for (size_t k = 0; k < 3; ++k)
          if (k != 1) {
            ob_elt(&ob, F, proof[i].hp[0][wi][k]);
            ob_elt(&ob, F, proof[i].hp[1][wi][k]);
          }
      ob_elt(&ob, F, proof[i].wc[0]);
      ob_elt(&ob, F, proof[i].wc[1]);
    }
    if (dump && dump->sumcheck && dump->sumcheck_cap >= ob.n - 32 && !ob.overflow)
      memcpy(dump->sumcheck, out + 32, ob.n - 32);
    for (size_t i = 0; i < P.block; ++i) ob_elt(&ob, F, y_ldt[i]);
    for (size_t i = 0; i < P.dblock; ++i) ob_elt(&ob, F, y_dot[i]);
    for (size_t i = 0; i < P.r; ++i) ob_elt(&ob, F, y_q[i]);
    for (size_t i = 0; i < P.dblock - P.block; ++i) ob_elt(&ob, F, y_q[P.block + i]);
    for (size_t i = 0; i < P.nreq; ++i) ob_put(&ob, nonces + 32 * idx[i], 32);
    /* compute_req :346-351 + run-length coding zk_proof.h:157-178 */
    {
      size_t total = P.nreq * P.nrow, ci = 0;
      int subrun = 0;
      while (ci < total) {
        size_t runlen = 0;
        while (ci + runlen < total && runlen < ((size_t)1 << 25)) {
          size_t k = ci + runlen;
          elt e = TAB(k / P.nreq, P.dblock + idx[k % P.nreq]);
          if (f_in_subfield(F, e) != subrun) break;
          ++runlen;
        }
        ob_u32(&ob, runlen);
        for (size_t k = ci; k < ci + runlen; ++k) {
          elt e = TAB(k / P.nreq, P.dblock + idx[k % P.nreq]);
          if (subrun) {
            uint8_t t[32];
            f_to_bytes_subfield(F, t, e);
            ob_put(&ob, t, F->ksubbytes);
          } else {
            ob_elt(&ob, F, e);
          }
        }
        ci += runlen;
        subrun = !subrun;
      }
    }
    ob_u32(&ob, pathlen);
    ob_put(&ob, path, 32 * pathlen);
    *out_len = ob.n;
    if (ob.overflow) rc = -4;
  }
#undef TAB
  free(W);
  free(wit);
  free(pad);
  free(proof);
  free(lqc);
  free(T);
  free(nodes);
  free(nonces);
  for (size_t i = 0; i < nl; ++i) free(in[i]);
  free(in);
  free(finalV);
  free(bound_quad);
  free(A.a);
  free(y_ldt);
  free(y_dot);
  free(y_q);
  free(Avec);
  free(idx);
  free(path);
  return rc;
}
int zk_prove(const circuit* c, const uint8_t* wbytes, rng* rg, const uint8_t* tinit, size_t tinit_len,
             size_t rate, size_t nreq, size_t block_enc, uint8_t* out, size_t out_cap, size_t* out_len,
             zk_dump* dump) {
  return zk_prove_ex(c, NULL, wbytes, rg, tinit, tinit_len, rate, nreq, block_enc, out, out_cap, out_len, dump);
}
int ligero_prove_generic(const ligero_generic* G, rng* rg, const uint8_t* tinit, size_t tinit_len, size_t rate,
                         size_t nreq, size_t block_enc, uint8_t* out, size_t out_cap, size_t* out_len) {
  return zk_prove_ex(NULL, G, NULL, rg, tinit, tinit_len, rate, nreq, block_enc, out, out_cap, out_len, NULL);
}
