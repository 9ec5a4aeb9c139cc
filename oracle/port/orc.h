/* oracle/port -- CPU restatement (plain C) of the Longfellow prover hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load liboracle.so; the product
 * (longfellow_zk_b200/) never links, loads or calls it.
 *
 * Parity status: PINNED.  Every function is checked by tests/test_oracle_*.py
 * against (a) the reference's own golden vectors (docs/specs/testvectors.md
 * Merkle + transcript vectors, rust/runtime/ test .bin files) and (b) the
 * unmodified reference compiled in place as oracle/_ref/libref.so, including
 * byte equality of whole serialized ZK proofs.
 *
 * Each function cites the reference file:line (under /root/reference/lib) it
 * restates.  Scalar, single threaded, portable C11 (+ unsigned __int128).
 */
#ifndef ORACLE_PORT_ORC_H_
#define ORACLE_PORT_ORC_H_
#include <stddef.h>
#include <stdint.h>

/* field ids: proto/circuit_io.h:24-36 where the reference has one, >=100 for
 * the benchmark-only fields of algebra/fft_test.cc / reed_solomon_test.cc */
enum {
  ORC_P256 = 1,
  ORC_GF2_128 = 4,
  ORC_SECP256K1 = 10, /* Fp<4> over the secp256k1 prime, random/transcript_test.cc:131-283 */
  ORC_BN254 = 100,    /* Fp<4>, fft_test.cc:33-36 */
  ORC_FP128 = 101,    /* fp_p128.h: 2^128 - 2^108 + 1 */
  ORC_GOLDILOCKS = 102 /* Fp<1>: 2^64 - 2^32 + 1 */
};

/* An element: GF(2^128) uses l[0..1] (bit i of the 128-bit word = coeff of
 * x^i, gf2k/gf2_128.h:64-89); prime fields use w64 little-endian limbs in
 * Montgomery form x*2^(64*w64) mod p (algebra/fp_generic.h:66-67). */
typedef struct { uint64_t l[4]; } elt;

typedef struct field {
  int id;
  int char2;
  int w64;
  size_t kbytes, ksubbytes;
  elt zero, one;
  elt evalpt[6];       /* poly_evaluation_point(i) */
  elt newton[6][6];    /* newton_denominator(k,i) */
  int nevalpt;
  /* prime fields */
  uint64_t m[4];
  uint64_t mprime;
  elt rsq;
  int exact_bits;
  /* GF(2^128): subfield GF(2^16) basis and its row-echelon form */
  elt beta[16];
  uint64_t sub_u[16][2];
  uint64_t sub_linv[16];
  int sub_ldnz[16];
  /* transforms: root of unity for prime fields with 2-power roots */
  elt omega;           /* in F (has_omega==1) */
  elt omega2[2];       /* in Fp2 (re,im) for P-256 (has_omega==2) */
  uint64_t omega_order;
  int has_omega;
} field;

const field* orc_field(int id);

elt f_add(const field* F, elt a, elt b);
elt f_sub(const field* F, elt a, elt b);
elt f_mul(const field* F, elt a, elt b);
elt f_neg(const field* F, elt a);
elt f_inv(const field* F, elt a);
int f_eq(const field* F, elt a, elt b);
int f_is_zero(const field* F, elt a);
elt f_of_scalar(const field* F, uint64_t u);
/* 0 on success, -1 if the bytes are not a canonical element */
int f_of_bytes(const field* F, const uint8_t* b, elt* out);
void f_to_bytes(const field* F, uint8_t* b, elt a);
int f_in_subfield(const field* F, elt a);
void f_to_bytes_subfield(const field* F, uint8_t* b, elt a);
elt f_of_bytes_subfield(const field* F, const uint8_t* b);

/* byte sources (RandomEngine, random/random.h:32-116) */
typedef struct rng {
  void (*bytes)(struct rng*, uint8_t*, size_t);
} rng;
elt rng_elt(rng* r, const field* F);          /* Field::sample */
elt rng_subfield_elt(rng* r, const field* F); /* Field::sample_subfield */
size_t rng_nat(rng* r, size_t n);             /* random.h:57-88 */
void rng_choose(rng* r, size_t* res, size_t n, size_t k); /* random.h:92-105 */

typedef struct {
  rng base;
  const uint8_t* p;
  size_t n, pos;
  int overrun;
} bufrng;
void bufrng_init(bufrng* r, const uint8_t* p, size_t n);

/* SHA-256 (FIPS 180-4), stands in for OpenSSL at util/crypto.h:41-70 */
typedef struct {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len;
} sha256;
void sha256_init(sha256* s);
void sha256_update(sha256* s, const uint8_t* p, size_t n);
void sha256_final(const sha256* s, uint8_t out[32]); /* non-destructive */
/* AES-256 single block encrypt (FIPS 197), util/crypto.h:74-103 */
typedef struct { uint8_t rk[15][16]; } aes256;
void aes256_init(aes256* a, const uint8_t key[32]);
void aes256_encrypt(const aes256* a, const uint8_t in[16], uint8_t out[16]);

/* Transcript (random/transcript.h:70-190) */
typedef struct {
  rng base;
  sha256 sha;
  int have_prf;
  aes256 prf;
  uint64_t nblock;
  size_t rdptr;
  uint8_t saved[16];
} transcript;
void ts_init(transcript* t, const uint8_t* init, size_t n);
void ts_write_bytes(transcript* t, const uint8_t* p, size_t n);
void ts_write0(transcript* t, size_t n);
void ts_write_elt(transcript* t, const field* F, elt e);
void ts_write_array(transcript* t, const field* F, const elt* e, size_t ince, size_t n);

/* LCH14 additive FFT (gf2k/lch14.h) and RS (gf2k/lch14_reed_solomon.h) */
void lch14_fft(size_t l, size_t coset, elt* B);
void lch14_ifft(size_t l, size_t coset, elt* B);
void lch14_bidir(size_t l, size_t k, elt* B);
elt lch14_what(size_t i, size_t j);
/* Field-generic RS extension: y[0..n) -> y[0..m) (reed_solomon.h:93-110,
 * lch14_reed_solomon.h:49-103) */
void rs_interpolate(const field* F, size_t n, size_t m, elt* y);
/* prime-field FFT (algebra/fft.h:185-201); fwd=0: fftb, fwd=1: fftf */
void fp_fft(const field* F, elt* A, size_t n, elt omega, uint64_t order, int fwd);
void fp2_fft(const field* F, elt* A /*2n: re,im*/, size_t n, const elt omega[2],
             uint64_t order, int fwd);

/* Merkle (merkle/merkle_tree.h, merkle_commitment.h) */
void merkle_build(size_t n, uint8_t* nodes /*2n*32, leaves at [n,2n)*/);
size_t merkle_tree_len(size_t n);
size_t merkle_open(size_t n, const uint8_t* nodes, const size_t* pos, size_t np,
                   uint8_t* path /*cap np*len*32*/);

/* Ligero parameters (ligero/ligero_param.h:116-307) */
typedef struct {
  size_t nw, nq, rateinv, nreq;
  size_t block_enc, block, dblock, block_ext, r, w, nwrow, nqtriples, nwqrow,
      nrow, mc_pathlen;
  size_t ildt, idot, iquad, iw, iq;
} ligero_param;
/* block_enc == 0: search powers of two as the deprecated constructor does */
int ligero_param_init(ligero_param* p, const field* F, size_t nw, size_t nq,
                      size_t rateinv, size_t nreq, size_t block_enc);

/* Circuit (sumcheck/circuit.h:29-66, proto/circuit_reader.h:41-258) */
typedef struct {
  size_t nw, logw, nterms;
  uint32_t *g, *h0, *h1, *vi; /* expanded (not delta coded) */
} layer;
typedef struct {
  const field* F;
  size_t nv, logv, nc, logc, nl, ninputs, npub_in, subfield_boundary, nterms;
  size_t nconst;
  elt* consts;
  layer* l;
  uint8_t id[32];
} circuit;
circuit* circuit_parse(const field* F, const uint8_t* b, size_t n);
void circuit_free(circuit* c);
void circuit_id(const circuit* c, uint8_t id[32]); /* sumcheck/circuit_id.h:30-67 */

/* Whole prover (zk/zk_prover.h:72-149 + zk/zk_proof.h:90-184).  Returns 0 on
 * success, <0 on error (-3: witness does not satisfy the circuit). */
typedef struct {
  uint8_t* witness; size_t witness_cap;   /* ligero witness, to_bytes_field */
  uint8_t* tableau; size_t tableau_cap;
  uint8_t* root;                          /* 32 */
  uint8_t* sumcheck; size_t sumcheck_cap; /* serialized sumcheck proof */
} zk_dump;
int zk_prove(const circuit* c, const uint8_t* witness_bytes, rng* r,
             const uint8_t* tinit, size_t tinit_len, size_t rate, size_t nreq,
             size_t block_enc, uint8_t* out, size_t out_cap, size_t* out_len,
             zk_dump* dump);
/* LigeroProver::commit + ::prove on a caller-given statement (ligero/ligero_prover.h:58-146): nw witnesses,
 * nq quadratic constraints W[x]*W[y] = W[z] (lqc: x, y, z per constraint), linear constraint terms
 * (c, w, k): sum over terms of row c of k * W[w] = b[c] (b itself is not needed to prove), the hash of the
 * statement.  out = root (32 bytes) | LigeroProof as ZkProof::write_com_proof serialises it. */
typedef struct {
  const field* F;
  size_t nw, nq, ncons, nterms, subfield_boundary;
  const elt* W;
  const size_t* lqc;                /* 3 * nq */
  const size_t *term_c, *term_w;    /* nterms */
  const elt* term_k;                /* nterms */
  const uint8_t* hash;              /* 32 */
} ligero_generic;
int ligero_prove_generic(const ligero_generic* G, rng* r, const uint8_t* tinit, size_t tinit_len, size_t rate,
                         size_t nreq, size_t block_enc, uint8_t* out, size_t out_cap, size_t* out_len);
#endif
