/* longfellow_b200.h -- C ABI of the B200-native Longfellow prover back end.
 *
 * Drop-in boundary for the data-parallel prover hot path of
 * dwcoen1234/longfellow-zk (lib/): every entry point names the reference
 * interface it replaces (file:line under the reference's lib/).  Plain C
 * linkage, opaque handles, plain pointers and sizes.  Field elements cross
 * this boundary in the reference's WIRE encoding (Field::to_bytes_field:
 * lib/algebra/fp_generic.h:378-380 de-Montgomerised little-endian kBytes;
 * lib/gf2k/gf2_128.h:178-180 16 little-endian bytes).
 *
 * All functions return 0 on success and a negative lf_status otherwise;
 * lf_last_error() gives a message for the calling thread.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with
 * LF_ERR_CUDA.  A context owns one CUDA stream; calls on one context are
 * serialised on it, different contexts are independent (the reference's
 * provers are likewise independent objects without shared mutable state).
 */
#ifndef LONGFELLOW_B200_H_
#define LONGFELLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field ids: the reference's proto FieldID (lib/proto/circuit_io.h:24-36) where
 * it defines one; >= 100 for the benchmark-only fields of
 * lib/algebra/fft_test.cc:33-44 and lib/algebra/reed_solomon_test.cc:337-401. */
enum lf_field {
  LF_FIELD_P256 = 1,        /* Fp256Base, lib/algebra/fp_p256.h */
  LF_FIELD_GF2_128 = 4,     /* GF2_128<>, lib/gf2k/gf2_128.h */
  LF_FIELD_BN254 = 100,     /* Fp<4>, fft_test.cc:33-36 */
  LF_FIELD_FP128 = 101,     /* Fp128, lib/algebra/fp_p128.h */
  LF_FIELD_GOLDILOCKS = 102 /* Fp<1> 2^64-2^32+1 */
};

enum lf_status {
  LF_OK = 0,
  LF_ERR_ARG = -1,          /* null / out-of-range argument */
  LF_ERR_CUDA = -2,         /* CUDA runtime failure (incl. no device) */
  LF_ERR_FORMAT = -3,       /* malformed circuit bytes / non-canonical element */
  LF_ERR_UNSUPPORTED = -4,  /* field or shape not built yet */
  LF_ERR_WITNESS = -5,      /* witness does not satisfy the circuit (ZkProver::prove == false) */
  LF_ERR_RNG = -6,          /* caller-supplied randomness too short (prime fields: including the draws
                               the reference's Field::sample repeats, see lf_zk_prove_batch) */
  LF_ERR_CAPACITY = -7,     /* output buffer too small */
  LF_ERR_VERIFY = -8,       /* the verifier rejected the proof (ZkVerifier::verify == false) */
  LF_ERR_INTERNAL = -100    /* the sumcheck prover's own consistency check failed: never expected */
};

typedef struct lf_ctx lf_ctx;
typedef struct lf_circuit lf_circuit;

/* ---- context ---------------------------------------------------------- */
/* device: CUDA ordinal.  stream: a cudaStream_t the caller owns (e.g. torch's
 * current stream) or NULL to let the context create its own. */
int lf_ctx_create(int device, void* stream, lf_ctx** out);
void lf_ctx_destroy(lf_ctx* ctx);
int lf_ctx_synchronize(lf_ctx* ctx);
const char* lf_last_error(void);
const char* lf_version(void);

/* ---- (a1,a3) field arithmetic, element-wise: out[i] = a[i] * b[i] ------ */
/* replaces Field::mulf (fp_generic.h:187-198, gf2_128.h:233-235); host buffers. */
int lf_elt_mul(lf_ctx* ctx, int field_id, const void* a, const void* b, void* out, size_t n);

/* ---- (a4) FFT ----------------------------------------------------------- */
/* replaces FFT<Field>::fftb (forward == 0: T[j] = sum_k F[k] w^{jk}) and fftf
 * (forward != 0: w^-1), lib/algebra/fft.h:185-201, in place on n = 2^k host
 * elements, unnormalised (fftf o fftb = n * id).  The root of unity is the
 * field's standard one: BN254 (order 2^28, fft_test.cc:38-44), Fp128 and
 * Goldilocks (order 2^32, reed_solomon_test.cc:358-360).  For LF_FIELD_P256 the
 * transform is over Fp2 = Fp[i]/(i^2+1) (fft_test.cc:168-172): elts are n pairs
 * (re, im) and the root is the order-2^31 element of mdoc_zk.cc:83-88.
 * For LF_FIELD_GF2_128 it is the additive FFT LCH14<GF2_128>::FFT(l, 0, B)
 * (forward != 0: novel-basis coefficients -> evaluations on the span of
 * beta_0..beta_{l-1}) resp. ::IFFT (forward == 0), lib/gf2k/lch14.h:106-146,
 * n = 2^l <= 65536. */
int lf_fft(lf_ctx* ctx, int field_id, void* elts, size_t n, int forward);
/* device-resident timing of fftb at size n (CUDA events on the context stream) */
int lf_fft_time(lf_ctx* ctx, int field_id, size_t n, int reps, double* ms_per_fft);
/* the same for nrows independent transforms of size n side by side (one call = all rows) */
int lf_fft_time_rows(lf_ctx* ctx, int field_id, size_t n, size_t nrows, int reps, double* ms_per_call);

/* ---- (a7,a9) Reed-Solomon row extension -------------------------------- */
/* replaces InterpolatorFactory::make(n, m)->interpolate(y) batched over rows
 * (lib/gf2k/lch14_reed_solomon.h:49-123, lib/algebra/reed_solomon.h:93-147):
 * rows is nrows x m elements (host), the first n of each row are the
 * evaluations at the injected points 0..n-1; entries n..m-1 are filled. */
int lf_rs_interpolate(lf_ctx* ctx, int field_id, size_t n, size_t m, void* rows, size_t nrows);
/* same with device-resident rows (row_stride in elements); asynchronous on the
 * context stream */
int lf_rs_interpolate_dev(lf_ctx* ctx, int field_id, size_t n, size_t m, void* d_rows,
                          size_t row_stride, size_t nrows);

/* device-resident timing of one lf_rs_interpolate_dev call on nrows rows (CUDA events on the context
 * stream; row contents are arbitrary limbs, the arithmetic does not depend on them) */
int lf_rs_time(lf_ctx* ctx, int field_id, size_t n, size_t m, size_t nrows, int reps, double* ms_per_call);

/* ---- (a11) Merkle column commitment ------------------------------------ */
/* replaces LigeroProver::commit's MerkleCommitment::commit(updhash, rng)
 * (lib/merkle/merkle_commitment.h:50-64, lib/ligero/ligero_prover.h:73-76):
 * tableau = nrow x block_enc elements (host, wire encoding); leaf j hashes
 * nonce_j || column dblock+j; nonces = block_ext x 32 bytes supplied by the
 * caller in leaf order (so the RandomEngine order stays with the caller).
 * nodes_out (optional) receives the 2*block_ext heap of digests. */
int lf_merkle_commit(lf_ctx* ctx, int field_id, size_t nrow, size_t block_enc, size_t dblock,
                     const void* tableau, const uint8_t* nonces, uint8_t root_out[32],
                     uint8_t* nodes_out);

/* ---- circuits ---------------------------------------------------------- */
/* lfc1: the reference's serialized circuit (lib/proto/circuit_reader.h:41-258).
 * rate/nreq/block_enc: LigeroParam arguments (lib/zk/zk_proof.h:63-75;
 * block_enc == 0 selects the deprecated power-of-two search of
 * lib/ligero/ligero_param.h:152-169). */
int lf_circuit_upload(lf_ctx* ctx, int field_id, const uint8_t* lfc1, size_t len, size_t rate,
                      size_t nreq, size_t block_enc, lf_circuit** out);
void lf_circuit_free(lf_circuit* c);
/* CircuitReader::from_bytes(buf, enforce_circuit_id = true) (lib/proto/circuit_reader.h:55-77):
 * recomputes circuit_id (lib/sumcheck/circuit_id.h:30-67) from the parsed circuit and compares it
 * with the 32 bytes stored in the file; LF_ERR_FORMAT if they differ.  computed_out (optional)
 * receives the recomputed id.  Host-side, one SHA-256 pass over all quad terms. */
int lf_circuit_verify_id(const lf_circuit* c, uint8_t computed_out[32]);

typedef struct lf_circuit_info {
  size_t ninputs, npub_in, nl, nterms, kbytes;
  size_t witness_bytes;   /* ninputs * kbytes: one proof's input wires */
  size_t rng_bytes;       /* random bytes one proof consumes (GF(2^128): exact;
                             prime fields: without rejections -- supply slack) */
  size_t max_proof_bytes; /* upper bound of the serialized proof */
  size_t block_enc, block, dblock, block_ext, nrow, r, w, nwrow, nqtriples, nreq;
  size_t nw;              /* Ligero witnesses = private inputs + pad */
  /* per-proof algorithmic work of the sumcheck kernel:
   * bytes = sum over rounds of T*(8+kB) + T*kB + 4.5*n0*kB + T'*(8+kB);
   * field multiplications = sum over rounds of T + 1.5*n0 + T' (+ bind_g) */
  size_t sumcheck_alg_bytes, sumcheck_mults;
  /* per-proof field multiplications of the whole prover (RS + eval + sumcheck + Ligero) */
  size_t total_mults;
  /* per-proof SHA-256 compressions (Merkle leaves + tree + transcript) */
  size_t sha_compressions;
  /* the same total split by stage (SURVEY.md 8(d) formulas): RS row encodes of commit,
   * eval_circuit, Ligero prove (incl. its RS rows); Merkle commit compressions */
  size_t rs_mults, eval_mults, ligero_mults, merkle_compressions;
  /* bytes of the LFC1 input this circuit occupied; the mdoc circuit file holds the signature
   * circuit and the hash circuit back to back (lib/circuits/mdoc/mdoc_zk.cc:440-456) */
  size_t lfc1_bytes;
  /* Caller randomness of one proof: rng_sample_bytes of field samples (pad,
   * ILDT, IDOT, IQUAD, blinding of the witness and quadratic rows), then block_ext 32-byte Merkle
   * nonces; rng_bytes is their sum.  Over a prime field every sample is one slot of
   * rng_redraw_bytes (= kbytes) and the reference's Field::sample (lib/algebra/fp_generic.h:360-371)
   * draws the slot again while its value is >= p: each redraw moves everything behind it, nonces
   * included, one slot further down the stream, exactly as with the reference's RandomEngine.  So
   * rng_bytes is what a proof consumes when no draw is rejected (2^-32 per sample for P-256) and a
   * lower bound otherwise; up to rng_redraw_cap redraws per proof are followed.  rng_redraw_bytes
   * == 0: the field's samples never fail (GF(2^128)) and rng_bytes is exact. */
  size_t rng_sample_bytes, rng_redraw_bytes, rng_redraw_cap;
  /* elements of the serialized sumcheck proof (ZkProof::write: root[32] | these | y_ldt[block] y_dot[dblock]
   * y_quad_0[r] y_quad_2[dblock-block] | nreq nonces | run-length coded columns | Merkle proof) */
  size_t sumcheck_proof_elts;
  /* Large batches run the big sumcheck rounds as grid-wide kernels (k_sc_eval / k_sc_round / k_sc_bind, one
   * launch each per round): number of such rounds, and per proof the algorithmic bytes and field
   * multiplications of all k_sc_eval launches (per CSR entry two indices, an HQuad value, a wire and one
   * product; per wire pair two wires and two products) and of all k_sc_bind launches. */
  size_t flat_rounds, flat_eval_alg_bytes, flat_eval_mults, flat_bind_alg_bytes, flat_bind_mults;
} lf_circuit_info;
int lf_circuit_get_info(const lf_circuit* c, lf_circuit_info* info);

/* With profiling on (lf_circuit_set_profiling), the most recent batch's device time of one kernel class of
 * the flat sumcheck, summed over its launches (CUDA events around every launch on the context stream):
 * cls 0 k_sc_eval, 1 k_sc_bind, 2 k_sc_round, 3 k_zk_sumcheck over the small rounds. */
int lf_circuit_get_kernel_ms(lf_circuit* c, int cls, float* ms_total, size_t* launches);

/* ---- whole prover, batch of independent proofs -------------------------- */
/* replaces, per proof i (lib/zk/zk_prover.h:72-149, lib/zk/zk_proof.h:90-112):
 *     Transcript tp(tinit, tinit_len);
 *     ZkProof zkp(circuit, rate, nreq[, block_enc]);
 *     ZkProver prover(circuit, F, rs_factory);
 *     prover.commit(zkp, W_i, tp, rng_i);  ok = prover.prove(zkp, W_i, tp);
 *     zkp.write(bytes_i, F);
 * witnesses: nproofs x ninputs elements (wire encoding).
 * rng: nproofs x rng_stride bytes; proof i consumes rng + i*rng_stride exactly
 *      as the reference consumes RandomEngine::bytes,
 *      rejected draws of Field::sample included (lf_circuit_info.rng_redraw_bytes):
 *      rng_stride >= rng_bytes is required, and a proof whose redraws need more
 *      than rng_stride bytes (or more than rng_redraw_cap redraws) is LF_ERR_RNG.
 *      lf_zk_rng_consumed() tells how many bytes a proof took.
 * proofs_out: nproofs x proof_stride bytes; proof_lens[i] = serialized length.
 * status[i]: LF_OK, or the first failure of that proof:
 *      LF_ERR_FORMAT   a witness element is not canonical (>= p),
 *      LF_ERR_RNG      the proof's random stream is too short (see above),
 *      LF_ERR_WITNESS  the witness does not satisfy the circuit (ZkProver::prove == false),
 *      LF_ERR_INTERNAL never expected.  A failed proof has proof_lens[i] == 0.
 * Host pointers; the call copies in, proves on the GPU and copies out. */
int lf_zk_prove_batch(lf_circuit* c, size_t nproofs, const uint8_t* witnesses, const uint8_t* rng,
                      size_t rng_stride, const uint8_t* tinit, size_t tinit_len,
                      uint8_t* proofs_out, size_t proof_stride, size_t* proof_lens, int* status);
/* same with every buffer resident in device memory (proof_lens/status: device
 * uint64_t / int32_t arrays); asynchronous on the context stream. */
int lf_zk_prove_batch_dev(lf_circuit* c, size_t nproofs, const void* d_witnesses, const void* d_rng,
                          size_t rng_stride, const uint8_t* tinit, size_t tinit_len,
                          void* d_proofs_out, size_t proof_stride, void* d_proof_lens,
                          void* d_status);

/* ---- commit and prove as separate calls on a caller-owned transcript ------ */
/* The reference's ZkProver::commit and ::prove take the caller's Transcript
 * (lib/zk/zk_prover.h:72-149), and run_mdoc_prover interleaves two provers on
 * ONE transcript: commit(hash), commit(sig), MAC key from the transcript, patch
 * of public inputs, prove(hash), prove(sig) (lib/circuits/mdoc/mdoc_zk.cc:459-503).
 * lf_transcript is the state that crosses: the SHA-256 state of everything
 * written so far plus the read position of the challenge stream
 * (lib/random/transcript.h:46-62,70-190). */
typedef struct lf_transcript {
  uint32_t h[8];       /* SHA-256 chaining value */
  uint32_t buf[16];    /* buffered message bytes, big-endian words */
  uint64_t len;        /* bytes absorbed */
  uint64_t nblock;     /* FSPRF: next AES block counter */
  uint32_t rdptr, have_prf;
  uint32_t saved[4];   /* current FSPRF block */
} lf_transcript;
/* Transcript(seed, n); write(data, n); bytes(out, n)  -- on the host */
void lf_transcript_init(lf_transcript* ts, const uint8_t* seed, size_t n);
void lf_transcript_write_bytes(lf_transcript* ts, const uint8_t* data, size_t n);
void lf_transcript_challenge_bytes(lf_transcript* ts, uint8_t* out, size_t n);
/* ZkProver::commit for nproofs proofs: ts[i] (in/out) receives the commitment
 * (root i, also returned in roots_out[32*i], may be NULL).  The committed
 * tableaux stay on the device inside `c` until the matching prove call. */
int lf_zk_commit_batch(lf_circuit* c, size_t nproofs, const uint8_t* witnesses, const uint8_t* rng,
                       size_t rng_stride, lf_transcript* ts, uint8_t* roots_out, int* status);
/* ZkProver::prove + ZkProof::write for the batch committed last on `c`.
 * witnesses: the same inputs; PUBLIC inputs may differ from the commit call
 * (they are not committed).  ts[i] in: the transcript to continue from; out:
 * the transcript as the prover left it. */
int lf_zk_prove_committed_batch(lf_circuit* c, size_t nproofs, const uint8_t* witnesses, lf_transcript* ts,
                                uint8_t* proofs_out, size_t proof_stride, size_t* proof_lens, int* status);

/* ---- verifier, batch of independent proofs -------------------------------------- */
/* replaces, per proof i (lib/zk/zk_verifier.h:69-106, lib/zk/zk_proof.h:107-112):
 *     ZkProof zkp(circuit, rate, nreq[, block_enc]);  ok = zkp.read(bytes_i, F);
 *     Transcript tv(tinit, tinit_len);
 *     ZkVerifier verifier(circuit, rs_factory, rate, nreq[, block_enc], F);
 *     verifier.recv_commitment(zkp, tv);  ok = verifier.verify(zkp, pub_i, tv);
 * pub_inputs: nproofs x npub_in elements (wire encoding; may be NULL when npub_in == 0).
 * proofs: nproofs x proof_stride bytes, proof i is proof_lens[i] bytes long.
 * status[i]: LF_OK accepted; LF_ERR_FORMAT the bytes are not a proof of this circuit's shape
 * (ZkProof::read == false) or a public input is not canonical; LF_ERR_VERIFY rejected.
 * why[i] (optional): the first check that failed, in the order of LigeroVerifier::verify
 * (lib/ligero/ligero_verifier.h:92-134): 1 merkle_check, 2 low_degree_check, 3 dot_check,
 * 4 wrong dot product, 5 quadratic_check; 0 otherwise.  Host pointers. */
int lf_zk_verify_batch(lf_circuit* c, size_t nproofs, const uint8_t* pub_inputs, const uint8_t* proofs,
                       size_t proof_stride, const size_t* proof_lens, const uint8_t* tinit, size_t tinit_len,
                       int* status, int* why);

/* ZkVerifier::verify on a caller-owned transcript (lib/zk/zk_verifier.h:69-106).  The reference's
 * recv_commitment and verify take the caller's Transcript so that several proofs can be composed on one:
 * run_mdoc_verifier does recv_commitment(hash), recv_commitment(sig), draws the MAC key, then verify(hash),
 * verify(sig) (lib/circuits/mdoc/mdoc_zk.cc:673-706).  ts[i] in: the transcript that has already received
 * this proof's commitment (recv_commitment is one write of the 32-byte root,
 * lib/ligero/ligero_transcript.h:31-34: lf_transcript_write_bytes(&ts, root, 32)); out: the transcript as
 * verify left it (after the draw of the opened columns).  Everything else as lf_zk_verify_batch. */
int lf_zk_verify_committed_batch(lf_circuit* c, size_t nproofs, const uint8_t* pub_inputs, const uint8_t* proofs,
                                 size_t proof_stride, const size_t* proof_lens, lf_transcript* ts, int* status,
                                 int* why);

/* Test hook, the analogue of BadReedSolomonFactory in lib/ligero/ligero_test.cc:114-160: make exactly one of the
 * verifier's own computations wrong so that each check of LigeroVerifier::verify can be seen to fire on an
 * otherwise valid proof.  fault: 0 none; 1 the interpolation of y_ldt (low_degree_check); 2 of y_dot
 * (dot_check); 3 of y_quad (quadratic_check); 4 of the first row of A (dot_check); 5 the value of b . alphal
 * plus one ("wrong dot product").  Stays set on `c` until reset. */
int lf_zk_verify_set_fault(lf_circuit* c, int fault);

/* bytes of proof `index`'s random stream that the most recent batch on `c` consumed
 * (= rng_bytes + redraws * rng_redraw_bytes) */
int lf_zk_rng_consumed(lf_circuit* c, size_t index, size_t* bytes);

/* ---- stage read-back for parity tests ---------------------------------- */
enum lf_stage {
  LF_STAGE_WITNESS = 1,   /* Ligero witness vector, nw elements */
  LF_STAGE_TABLEAU = 2,   /* nrow x block_enc elements */
  LF_STAGE_ROOT = 3,      /* 32 bytes */
  LF_STAGE_SUMCHECK = 4   /* serialized sumcheck proof */
};
/* copies stage data of proof `index` of the most recent batch (wire encoding) */
int lf_zk_debug_fetch(lf_circuit* c, size_t index, int stage, uint8_t* out, size_t cap,
                      size_t* len);

/* per-stage device times of the most recent batch, measured with CUDA events
 * on the context stream when profiling is enabled (bench.py's roofline leg).
 * stage order: 0 layout 1 rs_encode 2 merkle 3 transcript_init 4 eval_circuit
 *              5 sumcheck 6 ligero_prove ; returns the number of stages */
int lf_circuit_set_profiling(lf_circuit* c, int enable);
int lf_circuit_get_stage_ms(lf_circuit* c, float* ms, size_t cap);

/* number of kernel launches issued on behalf of this context so far */
uint64_t lf_ctx_launch_count(const lf_ctx* ctx);

/* integer-pipe micro-benchmarks (roofline denominators): returns achieved
 * giga-operations per second of `what` on the context's device.
 *   what = 0: IMAD.WIDE  1: LOP3  2: GF(2^128) multiply (Gmul/s)
 *          3: SHA-256 compressions (G/s)  4: P-256 Montgomery multiply (Gmul/s)
 *   what = 100..105: single-thread latency of the transcript primitives, in
 *          CYCLES per call: compression, digest snapshot, AES-256 key
 *          schedule, AES block, 16-byte element write, write + challenge;
 *          107: the first write + challenge of the launch (cold instruction cache) */
int lf_microbench(lf_ctx* ctx, int what, double* gops);

#ifdef __cplusplus
}
#endif
#endif /* LONGFELLOW_B200_H_ */
