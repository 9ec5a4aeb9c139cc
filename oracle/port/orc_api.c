/* oracle/port/orc_api.c -- byte-level entry points for ctypes (TEST
 * INFRASTRUCTURE; see orc.h).  All elements cross this boundary in the
 * reference's wire encoding (to_bytes_field). */
#include <stdlib.h>
#include <string.h>

#include "orc.h"

int orc_elt_op(int fid, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  const field* F = orc_field(fid);
  if (!F) return -1;
  for (size_t i = 0; i < n; ++i) {
    elt x, y, r;
    if (f_of_bytes(F, a + i * F->kbytes, &x)) return -2;
    if (b && f_of_bytes(F, b + i * F->kbytes, &y)) return -2;
    switch (op) {
      case 0: r = f_add(F, x, y); break;
      case 1: r = f_sub(F, x, y); break;
      case 2: r = f_mul(F, x, y); break;
      case 3: r = f_inv(F, x); break;
      default: return -3;
    }
    f_to_bytes(F, out + i * F->kbytes, r);
  }
  return 0;
}
void orc_gf128_of_scalar(const uint64_t* u, uint8_t* out, size_t n) {
  const field* F = orc_field(ORC_GF2_128);
  for (size_t i = 0; i < n; ++i) f_to_bytes(F, out + 16 * i, f_of_scalar(F, u[i]));
}
void orc_gf128_subfield_index(const uint8_t* a, uint32_t* out, size_t n) {
  const field* F = orc_field(ORC_GF2_128);
  for (size_t i = 0; i < n; ++i) {
    elt x;
    f_of_bytes(F, a + 16 * i, &x);
    if (f_in_subfield(F, x)) {
      uint8_t b[2];
      f_to_bytes_subfield(F, b, x);
      out[i] = b[0] | (b[1] << 8);
    } else {
      out[i] = 0xFFFFFFFFu;
    }
  }
}
void orc_gf128_constants(uint8_t* beta, uint8_t* pts, uint8_t* newton) {
  const field* F = orc_field(ORC_GF2_128);
  for (int i = 0; i < 16; ++i) f_to_bytes(F, beta + 16 * i, F->beta[i]);
  for (int i = 0; i < 6; ++i) f_to_bytes(F, pts + 16 * i, F->evalpt[i]);
  memset(newton, 0, 6 * 6 * 16);
  for (int k = 1; k < 6; ++k)
    for (int i = 1; i <= k; ++i) f_to_bytes(F, newton + 16 * (k * 6 + i), F->newton[k][i]);
}
void orc_lch14(int op, size_t l, size_t coset_or_k, uint8_t* B) {
  const field* F = orc_field(ORC_GF2_128);
  size_t n = (size_t)1 << l;
  elt* v = (elt*)calloc(n, sizeof(elt));
  for (size_t i = 0; i < n; ++i) f_of_bytes(F, B + 16 * i, &v[i]);
  if (op == 0) lch14_fft(l, coset_or_k, v);
  else if (op == 1) lch14_ifft(l, coset_or_k, v);
  else lch14_bidir(l, coset_or_k, v);
  for (size_t i = 0; i < n; ++i) f_to_bytes(F, B + 16 * i, v[i]);
  free(v);
}
void orc_lch14_what(uint8_t* out) {
  const field* F = orc_field(ORC_GF2_128);
  for (size_t i = 0; i < 16; ++i)
    for (size_t j = 0; j < 16; ++j) f_to_bytes(F, out + 16 * (i * 16 + j), lch14_what(i, j));
}
/* rows: nrows x m elements (first n valid), extended in place */
int orc_rs_interpolate(int fid, size_t n, size_t m, uint8_t* rows, size_t nrows) {
  const field* F = orc_field(fid);
  if (!F) return -1;
  elt* y = (elt*)calloc(m, sizeof(elt));
  for (size_t r = 0; r < nrows; ++r) {
    uint8_t* p = rows + r * m * F->kbytes;
    for (size_t i = 0; i < n; ++i)
      if (f_of_bytes(F, p + i * F->kbytes, &y[i])) { free(y); return -2; }
    rs_interpolate(F, n, m, y);
    for (size_t i = 0; i < m; ++i) f_to_bytes(F, p + i * F->kbytes, y[i]);
  }
  free(y);
  return 0;
}
/* FFT over F (fid 100/101/102) or over Fp2(P-256) (fid 1; data = n x (re,im)) */
int orc_fft(int fid, uint8_t* data, size_t n, int fwd) {
  const field* F = orc_field(fid);
  if (!F || !F->has_omega) return -1;
  size_t per = F->has_omega == 2 ? 2 : 1;
  elt* A = (elt*)calloc(n * per, sizeof(elt));
  for (size_t i = 0; i < n * per; ++i)
    if (f_of_bytes(F, data + i * F->kbytes, &A[i])) { free(A); return -2; }
  if (per == 1) fp_fft(F, A, n, F->omega, F->omega_order, fwd);
  else fp2_fft(F, A, n, F->omega2, F->omega_order, fwd);
  for (size_t i = 0; i < n * per; ++i) f_to_bytes(F, data + i * F->kbytes, A[i]);
  free(A);
  return 0;
}
void orc_sha256(const uint8_t* p, size_t n, uint8_t out[32]) {
  sha256 s;
  sha256_init(&s);
  sha256_update(&s, p, n);
  sha256_final(&s, out);
}
void orc_aes256_ecb(const uint8_t key[32], const uint8_t* in, uint8_t* out, size_t nblocks) {
  aes256 a;
  aes256_init(&a, key);
  for (size_t i = 0; i < nblocks; ++i) aes256_encrypt(&a, in + 16 * i, out + 16 * i);
}
void orc_merkle_build(size_t n, const uint8_t* leaves, uint8_t* nodes_out, uint8_t* root_out) {
  memset(nodes_out, 0, 32 * n);
  memcpy(nodes_out + 32 * n, leaves, 32 * n);
  merkle_build(n, nodes_out);
  memcpy(root_out, nodes_out + 32, 32);
}
size_t orc_merkle_tree_len(size_t n) { return merkle_tree_len(n); }
/* merkle/merkle_commitment.h:50-73 */
size_t orc_merkle_commit_open(size_t n, const uint8_t* payload, size_t len, const uint8_t* rngb,
                              uint8_t* root_out, const size_t* pos, size_t np, uint8_t* nonce_out,
                              uint8_t* path_out) {
  uint8_t* nodes = (uint8_t*)calloc(2 * n, 32);
  for (size_t i = 0; i < n; ++i) {
    sha256 s;
    sha256_init(&s);
    sha256_update(&s, rngb + 32 * i, 32);
    sha256_update(&s, payload + i * len, len);
    sha256_final(&s, nodes + 32 * (n + i));
  }
  merkle_build(n, nodes);
  memcpy(root_out, nodes + 32, 32);
  size_t k = 0;
  if (np) {
    for (size_t i = 0; i < np; ++i) memcpy(nonce_out + 32 * i, rngb + 32 * pos[i], 32);
    k = merkle_open(n, nodes, pos, np, path_out);
  }
  free(nodes);
  return k;
}
/* same script language as oracle/ref_build/ref_common.cc:ref_transcript_script,
 * with 'E'/'A'/'G' using field `fid` */
size_t orc_transcript_script(int fid, const uint8_t* init, size_t init_len, const uint8_t* script,
                             size_t script_len, uint8_t* out, size_t out_cap) {
  const field* F = orc_field(fid);
  transcript ts;
  ts_init(&ts, init, init_len);
  size_t p = 0, o = 0;
#define RD32(v) do { memcpy(&(v), script + p, 4); p += 4; } while (0)
  while (p < script_len) {
    char op = (char)script[p++];
    uint32_t n, k;
    if (op == 'B') {
      RD32(n);
      ts_write_bytes(&ts, script + p, n);
      p += n;
    } else if (op == 'Z') {
      RD32(n);
      ts_write0(&ts, n);
    } else if (op == 'E') {
      elt e;
      f_of_bytes(F, script + p, &e);
      p += F->kbytes;
      ts_write_elt(&ts, F, e);
    } else if (op == 'A') {
      RD32(n);
      elt* v = (elt*)calloc(n ? n : 1, sizeof(elt));
      for (uint32_t i = 0; i < n; ++i) f_of_bytes(F, script + p + F->kbytes * i, &v[i]);
      p += F->kbytes * n;
      ts_write_array(&ts, F, v, 1, n);
      free(v);
    } else if (op == 'R') {
      RD32(n);
      if (o + n > out_cap) return 0;
      ts.base.bytes(&ts.base, out + o, n);
      o += n;
    } else if (op == 'N') {
      RD32(n);
      uint32_t r = (uint32_t)rng_nat(&ts.base, n);
      memcpy(out + o, &r, 4);
      o += 4;
    } else if (op == 'C') {
      RD32(n);
      RD32(k);
      size_t* res = (size_t*)malloc(k * sizeof(size_t));
      rng_choose(&ts.base, res, n, k);
      for (uint32_t i = 0; i < k; ++i) {
        uint32_t r = (uint32_t)res[i];
        memcpy(out + o, &r, 4);
        o += 4;
      }
      free(res);
    } else if (op == 'G') {
      RD32(n);
      for (uint32_t i = 0; i < n; ++i) {
        f_to_bytes(F, out + o, rng_elt(&ts.base, F));
        o += F->kbytes;
      }
    } else if (op == 'K') { /* Transcript::get: 32-byte snapshot */
      sha256_final(&ts.sha, out + o);
      o += 32;
    } else {
      return 0;
    }
  }
#undef RD32
  return o;
}
int orc_ligero_param(int fid, size_t nw, size_t nq, size_t rate, size_t nreq, size_t block_enc,
                     size_t* out) {
  const field* F = orc_field(fid);
  ligero_param p;
  if (!F || ligero_param_init(&p, F, nw, nq, rate, nreq, block_enc)) return -1;
  size_t v[12] = {p.block_enc, p.block, p.dblock, p.block_ext, p.r, p.w, p.nwrow,
                  p.nqtriples, p.nwqrow, p.nrow, p.mc_pathlen, p.iq};
  memcpy(out, v, sizeof(v));
  return 0;
}
void* orc_circuit_load(int fid, const uint8_t* b, size_t n) {
  const field* F = orc_field(fid);
  return F ? circuit_parse(F, b, n) : NULL;
}
void orc_circuit_free(void* c) { circuit_free((circuit*)c); }
void orc_circuit_id(void* c, uint8_t id[32]) { circuit_id((const circuit*)c, id); }
/* the generic Ligero prover on a statement given in wire encoding (elements: to_bytes_field; indices: u64) */
int orc_ligero_prove(int fid, size_t nw, size_t nq, size_t ncons, size_t nterms, size_t subfield_boundary,
                     const uint8_t* w_bytes, const uint64_t* lqc, const uint64_t* term_c, const uint64_t* term_w,
                     const uint8_t* term_k_bytes, const uint8_t* hash, const uint8_t* rngb, size_t rng_len,
                     const uint8_t* tinit, size_t tinit_len, size_t rate, size_t nreq, size_t block_enc,
                     uint8_t* out, size_t out_cap, size_t* out_len, size_t* rng_used) {
  const field* F = orc_field(fid);
  if (!F) return -1;
  elt* W = (elt*)malloc((nw ? nw : 1) * sizeof(elt));
  elt* K = (elt*)malloc((nterms ? nterms : 1) * sizeof(elt));
  size_t* q = (size_t*)malloc((3 * nq + 1) * sizeof(size_t));
  size_t* tc = (size_t*)malloc((nterms ? nterms : 1) * sizeof(size_t));
  size_t* tw = (size_t*)malloc((nterms ? nterms : 1) * sizeof(size_t));
  int rc = 0;
  for (size_t i = 0; i < nw && !rc; ++i) rc = f_of_bytes(F, w_bytes + i * F->kbytes, &W[i]) ? -2 : 0;
  for (size_t i = 0; i < nterms && !rc; ++i) {
    rc = f_of_bytes(F, term_k_bytes + i * F->kbytes, &K[i]) ? -2 : 0;
    tc[i] = (size_t)term_c[i];
    tw[i] = (size_t)term_w[i];
    if (tc[i] >= ncons || tw[i] >= nw) rc = -2;
  }
  for (size_t i = 0; i < 3 * nq && !rc; ++i) {
    q[i] = (size_t)lqc[i];
    if (q[i] >= nw) rc = -2;
  }
  if (!rc) {
    ligero_generic G = {F, nw, nq, ncons, nterms, subfield_boundary, W, q, tc, tw, K, hash};
    bufrng r;
    bufrng_init(&r, rngb, rng_len);
    rc = ligero_prove_generic(&G, &r.base, tinit, tinit_len, rate, nreq, block_enc, out, out_cap, out_len);
    if (rng_used) *rng_used = r.pos;
    if (r.overrun) rc = -20;
  }
  free(W);
  free(K);
  free(q);
  free(tc);
  free(tw);
  return rc;
}
int orc_zk_prove(void* c, const uint8_t* wit, const uint8_t* rngb, size_t rng_len,
                 const uint8_t* tinit, size_t tinit_len, size_t rate, size_t nreq, size_t block_enc,
                 uint8_t* out, size_t out_cap, size_t* out_len, size_t* rng_used, uint8_t* d_witness,
                 size_t d_witness_cap, uint8_t* d_tableau, size_t d_tableau_cap, uint8_t* d_root,
                 uint8_t* d_sumcheck, size_t d_sumcheck_cap) {
  bufrng r;
  bufrng_init(&r, rngb, rng_len);
  zk_dump d = {d_witness, d_witness_cap, d_tableau, d_tableau_cap, d_root, d_sumcheck, d_sumcheck_cap};
  int rc = zk_prove((const circuit*)c, wit, &r.base, tinit, tinit_len, rate, nreq, block_enc, out,
                    out_cap, out_len, &d);
  if (rng_used) *rng_used = r.pos;
  if (r.overrun) return -20;
  return rc;
}
