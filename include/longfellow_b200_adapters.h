/* longfellow_b200_adapters.h -- header-only C++ adapters that present the C ABI
 * of longfellow_b200.h through the reference's own (duck-typed) interfaces, so
 * that code written against dwcoen1234/longfellow-zk compiles against either
 * back end.  Nothing here needs a change to the reference sources; the header
 * expects the reference's lib/ directory on the include path.
 *
 *   GpuReedSolomonFactory<Field>   the interpolator-factory seam
 *       (lib/algebra/reed_solomon.h:132-147, lib/gf2k/lch14_reed_solomon.h:112-123):
 *       make(n, m)->interpolate(Elt y[m]).  Usable as the second template
 *       argument of the reference's LigeroProver / ZkProver / ZkVerifier:
 *           ZkProver<GF2_128<>, GpuReedSolomonFactory<GF2_128<>>> p(circuit, F, rsf);
 *       Row-at-a-time with host pointers: the parity path, not the fast one.
 *
 *   GpuZkProver<Field>             the whole prover
 *       (ZkProver::commit + ZkProver::prove + ZkProof::write,
 *        lib/zk/zk_prover.h:72-149, lib/zk/zk_proof.h:90-105), batched.
 *
 *   ZkProverGpu<Field, RSFactory>  drop-in for the reference's ZkProver<Field, RSFactory>
 *       (lib/zk/zk_prover.h:52-198): the same constructor (circuit, F, rs_factory) and the same
 *           void commit(ZkProof<Field>&, const Dense<Field>& W, Transcript&, RandomEngine&);
 *           bool prove(ZkProof<Field>&, const Dense<Field>& W, Transcript&);
 *       The reference's Transcript (its SHA-256 state and challenge stream) is carried across the C
 *       ABI in an lf_transcript and written back, and the ZkProof is filled from the returned bytes,
 *       so code such as run_mdoc_prover (lib/circuits/mdoc/mdoc_zk.cc:398-547: two provers over two
 *       fields interleaved on ONE transcript, MAC key drawn between commit and prove) compiles and
 *       runs unchanged with `ZkProver` naming this class: oracle/ref_build/ref_mdoc_gpu.cc does
 *       exactly that (make mdoc_gpu -> oracle/_ref/libref_mdoc_gpu.so exports run_mdoc_prover with the
 *       signature of lib/circuits/mdoc/mdoc_zk.h:157-164).
 *
 *   ZkVerifierGpu<Field, RSFactory>  drop-in for the reference's ZkVerifier<Field, RSFactory>
 *       (lib/zk/zk_verifier.h:41-111), see the class comment; with it run_mdoc_verifier runs on the GPU.
 *
 * oracle/ref_build/ref_common.cc compiles both against the unmodified
 * reference (make gpu -> oracle/_ref/libref_gpu.so); tests/test_gpu_adapters.py
 * checks that the reference's ZkProver running on GpuReedSolomonFactory, and
 * GpuZkProver, both reproduce the reference's proof bytes.
 */
#ifndef LONGFELLOW_B200_ADAPTERS_H_
#define LONGFELLOW_B200_ADAPTERS_H_

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "arrays/dense.h"
#include "longfellow_b200.h"
#include "proto/circuit_io.h"
#include "proto/circuit_writer.h"
#include "random/random.h"
#include "random/transcript.h"
#include "sumcheck/circuit.h"
#include "util/crypto.h"
#include "util/panic.h"
#include "util/readbuffer.h"
#include "zk/zk_proof.h"

namespace longfellow_b200 {

/* invariant failures abort like the reference's check() (lib/util/panic.h:27-36) */
inline void lf_check(int rc) { proofs::check(rc == LF_OK, lf_last_error()); }

template <class Field>
class GpuReedSolomon {
  using Elt = typename Field::Elt;

 public:
  GpuReedSolomon(lf_ctx* ctx, int field_id, const Field& F, size_t n, size_t m)
      : ctx_(ctx), field_id_(field_id), f_(F), n_(n), m_(m) {}

  /* y[0..n) given, y[n..m) produced: ReedSolomon::interpolate
   * (lib/algebra/reed_solomon.h:93-110), LCH14ReedSolomon::interpolate
   * (lib/gf2k/lch14_reed_solomon.h:49-103).  Elements cross the C ABI in the
   * wire encoding (to_bytes_field), whatever the field's in-memory form is. */
  void interpolate(Elt y[/*m*/]) const {
    std::vector<uint8_t> buf(m_ * Field::kBytes);
    for (size_t i = 0; i < n_; ++i) f_.to_bytes_field(&buf[i * Field::kBytes], y[i]);
    lf_check(lf_rs_interpolate(ctx_, field_id_, n_, m_, buf.data(), /*nrows=*/1));
    for (size_t i = n_; i < m_; ++i) {
      auto e = f_.of_bytes_field(&buf[i * Field::kBytes]);
      proofs::check(e.has_value(), "lf_rs_interpolate returned a non-canonical element");
      y[i] = e.value();
    }
  }

 private:
  lf_ctx* ctx_;
  int field_id_;
  const Field& f_;
  size_t n_, m_;
};

template <class Field>
class GpuReedSolomonFactory {
 public:
  GpuReedSolomonFactory(lf_ctx* ctx, int field_id, const Field& F) : ctx_(ctx), field_id_(field_id), f_(F) {}
  std::unique_ptr<GpuReedSolomon<Field>> make(size_t n, size_t m) const {
    return std::make_unique<GpuReedSolomon<Field>>(ctx_, field_id_, f_, n, m);
  }

 private:
  lf_ctx* ctx_;
  int field_id_;
  const Field& f_;
};

template <class Field>
class GpuZkProver {
 public:
  /* the circuit travels as LFC1 bytes (lib/proto/circuit_writer.h:36-86) and stays on the device */
  GpuZkProver(lf_ctx* ctx, const proofs::Circuit<Field>& c, const Field& F, proofs::FieldID fid, size_t rate,
              size_t nreq, size_t block_enc = 0)
      : f_(F) {
    std::vector<uint8_t> lfc1;
    proofs::CircuitWriter<Field>(F, fid).to_bytes(c, lfc1);
    lf_check(lf_circuit_upload(ctx, (int)fid, lfc1.data(), lfc1.size(), rate, nreq, block_enc, &circ_));
    lf_check(lf_circuit_get_info(circ_, &info_));
  }
  ~GpuZkProver() { lf_circuit_free(circ_); }
  GpuZkProver(const GpuZkProver&) = delete;
  GpuZkProver& operator=(const GpuZkProver&) = delete;

  const lf_circuit_info& info() const { return info_; }

  /* Exactly the bytes ZkProver::commit would take from `rng` for one proof, in its order: the field
   * samples -- over a prime field Field::sample draws a kBytes slot again while its value is >= p
   * (lib/algebra/fp_generic.h:360-371), and so does this: every rejected slot is followed by one
   * more drawn slot -- then the Merkle nonces.  A RandomEngine cannot be rewound, so the redraws
   * are resolved here, on the reference's side of the C ABI, and the back end then finds the same
   * rejected slots in the stream it is given. */
  void draw_coins(proofs::RandomEngine& rng, std::vector<uint8_t>& coins) const {
    coins.resize(info_.rng_sample_bytes);
    rng.bytes(coins.data(), coins.size());
    const size_t slot = info_.rng_redraw_bytes;
    if (slot != 0) {
      size_t checked = 0;
      for (;;) {
        size_t redraw = 0;
        for (; checked + slot <= coins.size(); checked += slot)
          if (!f_.of_bytes_field(&coins[checked]).has_value()) ++redraw;
        if (redraw == 0) break;
        const size_t at = coins.size();
        coins.resize(at + redraw * slot);
        rng.bytes(&coins[at], redraw * slot);
      }
    }
    const size_t at = coins.size(), nonce_bytes = info_.rng_bytes - info_.rng_sample_bytes;
    coins.resize(at + nonce_bytes);
    rng.bytes(&coins[at], nonce_bytes);
  }

  /* One proof per witness W[i] (the reference's Dense<Field>(1, ninputs)); every proof
   * starts from Transcript(tinit, tinit_len) and draws its coins from `rng` in the order
   * the reference's ZkProver does.  ok[i] is what ZkProver::prove would return. */
  std::vector<bool> prove_batch(const std::vector<const proofs::Dense<Field>*>& W, const uint8_t* tinit,
                                size_t tinit_len, proofs::RandomEngine& rng,
                                std::vector<std::vector<uint8_t>>& proofs_out) {
    const size_t B = W.size();
    std::vector<uint8_t> wit(B * info_.witness_bytes);
    std::vector<std::vector<uint8_t>> draws(B);
    size_t stride = info_.rng_bytes;
    for (size_t i = 0; i < B; ++i) {
      for (size_t k = 0; k < info_.ninputs; ++k)
        f_.to_bytes_field(&wit[i * info_.witness_bytes + k * Field::kBytes], W[i]->v_[k]);
      draw_coins(rng, draws[i]);
      stride = std::max(stride, draws[i].size());
    }
    std::vector<uint8_t> coins(B * stride);
    for (size_t i = 0; i < B; ++i) std::copy(draws[i].begin(), draws[i].end(), coins.begin() + i * stride);
    std::vector<uint8_t> out(B * info_.max_proof_bytes);
    std::vector<size_t> len(B);
    std::vector<int> st(B);
    lf_check(lf_zk_prove_batch(circ_, B, wit.data(), coins.data(), stride, tinit, tinit_len, out.data(),
                               info_.max_proof_bytes, len.data(), st.data()));
    std::vector<bool> ok(B);
    proofs_out.resize(B);
    for (size_t i = 0; i < B; ++i) {
      ok[i] = st[i] == LF_OK;
      const uint8_t* p = &out[i * info_.max_proof_bytes];
      proofs_out[i].assign(p, p + (ok[i] ? len[i] : 0));
    }
    return ok;
  }

 private:
  const Field& f_;
  lf_circuit* circ_ = nullptr;
  lf_circuit_info info_{};
};

/* ---- the reference's Transcript across the C ABI ------------------------------------------------
 * Transcript keeps its SHA-256 state (OpenSSL's SHA256_CTX inside proofs::SHA256) and its FSPRF
 * private (lib/random/transcript.h:183-186, lib/util/crypto.h:68-69).  The reference is not to be
 * modified, so the members are reached through explicit template instantiation, which may name
 * private members ([temp.explicit]): no #define private, no patched header. */
namespace detail {
template <class Tag>
struct Stash {
  static typename Tag::type ptr;
};
template <class Tag>
typename Tag::type Stash<Tag>::ptr;
template <class Tag, typename Tag::type P>
struct Grab {
  Grab() { Stash<Tag>::ptr = P; }
  static Grab instance;
};
template <class Tag, typename Tag::type P>
Grab<Tag, P> Grab<Tag, P>::instance;
struct TranscriptSha {
  typedef proofs::SHA256 proofs::Transcript::*type;
};
struct TranscriptPrf {
  typedef std::unique_ptr<proofs::FSPRF> proofs::Transcript::*type;
};
struct ShaCtx {
  typedef SHA256_CTX proofs::SHA256::*type;
};
template struct Grab<TranscriptSha, &proofs::Transcript::sha_>;
template struct Grab<TranscriptPrf, &proofs::Transcript::prf_>;
template struct Grab<ShaCtx, &proofs::SHA256::sha_>;
inline SHA256_CTX& sha_ctx(proofs::Transcript& t) {
  return (t.*Stash<TranscriptSha>::ptr).*Stash<ShaCtx>::ptr;
}
}  // namespace detail

/* Everything written so far.  The challenge stream is not exported: ZkProver::commit and ::prove
 * both start with a write, which discards it (transcript.h:169-173). */
inline void transcript_export(proofs::Transcript& t, lf_transcript* o) {
  const SHA256_CTX& c = detail::sha_ctx(t);
  std::memset(o, 0, sizeof(*o));
  for (int i = 0; i < 8; ++i) o->h[i] = c.h[i];
  o->len = ((((uint64_t)c.Nh) << 32) | c.Nl) >> 3;
  const uint8_t* d = reinterpret_cast<const uint8_t*>(c.data);
  for (unsigned k = 0; k < c.num; ++k) o->buf[k >> 2] |= (uint32_t)d[k] << (24 - 8 * (k & 3));
  o->rdptr = 16;
}
/* ... and back, the challenge stream included: the FSPRF is keyed by the hash of what was written,
 * so drawing the bytes the device already drew puts it at the same position. */
inline void transcript_import(proofs::Transcript& t, const lf_transcript& s) {
  SHA256_CTX& c = detail::sha_ctx(t);
  for (int i = 0; i < 8; ++i) c.h[i] = s.h[i];
  const uint64_t bits = s.len << 3;
  c.Nl = (uint32_t)bits;
  c.Nh = (uint32_t)(bits >> 32);
  c.num = (unsigned)(s.len & 63);
  uint8_t* d = reinterpret_cast<uint8_t*>(c.data);
  std::memset(d, 0, 64);
  for (unsigned k = 0; k < c.num; ++k) d[k] = (uint8_t)(s.buf[k >> 2] >> (24 - 8 * (k & 3)));
  (t.*detail::Stash<detail::TranscriptPrf>::ptr).reset();
  if (s.have_prf && s.nblock >= 1) {
    std::vector<uint8_t> drawn((size_t)(s.nblock - 1) * 16 + s.rdptr);
    if (!drawn.empty()) t.bytes(drawn.data(), drawn.size());
  }
}

/* proto FieldID of a reference field type; specialise next to the field's header
 * (GF2_128<> -> GF2_128_ID, Fp256Base -> P256_ID: the two fields of the ZK pipeline) */
template <class Field>
struct LfFieldId;

/* the process-wide context of the drop-in provers: device LF_DEVICE (default 0) */
inline lf_ctx* default_ctx() {
  static lf_ctx* ctx = [] {
    lf_ctx* c = nullptr;
    const char* e = std::getenv("LF_DEVICE");
    lf_check(lf_ctx_create(e ? std::atoi(e) : 0, nullptr, &c));
    return c;
  }();
  return ctx;
}

/* uploaded circuits, by circuit id and Ligero parameters: run_mdoc_prover re-parses its circuits on every
 * call (mdoc_zk.cc:432-456); the device copy and its sumcheck plans are built once */
inline lf_circuit* cached_circuit(const uint8_t id[32], int fid, size_t rate, size_t nreq, size_t block_enc,
                                  const std::vector<uint8_t>& (*make)(void*), void* arg) {
  static std::mutex mu;
  static std::map<std::string, lf_circuit*> cache;
  std::string key(reinterpret_cast<const char*>(id), 32);
  key += "/" + std::to_string(fid) + "/" + std::to_string(rate) + "/" + std::to_string(nreq) + "/" +
         std::to_string(block_enc);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const std::vector<uint8_t>& lfc1 = make(arg);
  lf_circuit* c = nullptr;
  lf_check(lf_circuit_upload(default_ctx(), fid, lfc1.data(), lfc1.size(), rate, nreq, block_enc, &c));
  cache.emplace(key, c);
  return c;
}

template <class Field, class RSFactory>
class ZkProverGpu {
  using Elt = typename Field::Elt;

 public:
  /* the interpolator factory is the reference's CPU one and is not used: Reed-Solomon runs on the device */
  ZkProverGpu(const proofs::Circuit<Field>& c, const Field& F, const RSFactory&) : c_(c), f_(F) {}

  /* ZkProver::commit (lib/zk/zk_prover.h:72-100) */
  void commit(proofs::ZkProof<Field>& zkp, const proofs::Dense<Field>& W, proofs::Transcript& tp,
              proofs::RandomEngine& rng) {
    bind(zkp);
    std::vector<uint8_t> wit, coins;
    witness_bytes(W, wit);
    draw_coins(rng, coins);
    lf_transcript ts;
    transcript_export(tp, &ts);
    uint8_t root[32];
    int st = 0;
    lf_check(lf_zk_commit_batch(circ_, 1, wit.data(), coins.data(), coins.size(), &ts, root, &st));
    proofs::check(st == LF_OK, "lf_zk_commit_batch: the commitment failed");
    transcript_import(tp, ts);
    std::memcpy(zkp.com.root.data, root, 32);
  }

  /* ZkProver::prove (lib/zk/zk_prover.h:102-149): false when the witness does not satisfy the circuit */
  bool prove(proofs::ZkProof<Field>& zkp, const proofs::Dense<Field>& W, proofs::Transcript& tp) {
    proofs::check(circ_ != nullptr, "prove before commit");
    std::vector<uint8_t> wit;
    witness_bytes(W, wit);
    lf_transcript ts;
    transcript_export(tp, &ts);
    std::vector<uint8_t> out(info_.max_proof_bytes);
    size_t len = 0;
    int st = 0;
    lf_check(lf_zk_prove_committed_batch(circ_, 1, wit.data(), &ts, out.data(), out.size(), &len, &st));
    if (st == LF_ERR_WITNESS) return false;
    proofs::check(st == LF_OK, "lf_zk_prove_committed_batch failed");
    transcript_import(tp, ts);
    proofs::ReadBuffer rb(out.data(), len);
    proofs::check(zkp.read(rb, f_), "the device's proof does not parse as a ZkProof of this circuit");
    return true;
  }

 private:
  static const std::vector<uint8_t>& serialize(void* self) {
    auto* p = static_cast<ZkProverGpu*>(self);
    proofs::CircuitWriter<Field>(p->f_, LfFieldId<Field>::value).to_bytes(p->c_, p->lfc1_);
    return p->lfc1_;
  }
  void bind(const proofs::ZkProof<Field>& zkp) {
    circ_ = cached_circuit(c_.id, (int)LfFieldId<Field>::value, zkp.param.rateinv, zkp.param.nreq,
                           zkp.param.block_enc, &ZkProverGpu::serialize, this);
    lfc1_.clear();
    lfc1_.shrink_to_fit();
    lf_check(lf_circuit_get_info(circ_, &info_));
    proofs::check(info_.block_enc == zkp.param.block_enc && info_.nrow == zkp.param.nrow &&
                      info_.block == zkp.param.block,
                  "Ligero parameters of the device circuit differ from the ZkProof's");
  }
  void witness_bytes(const proofs::Dense<Field>& W, std::vector<uint8_t>& wit) const {
    wit.resize(info_.witness_bytes);
    for (size_t k = 0; k < info_.ninputs; ++k) f_.to_bytes_field(&wit[k * Field::kBytes], W.v_[k]);
  }
  /* as GpuZkProver::draw_coins: the bytes ZkProver::commit takes from rng, redraws of Field::sample included */
  void draw_coins(proofs::RandomEngine& rng, std::vector<uint8_t>& coins) const {
    coins.resize(info_.rng_sample_bytes);
    rng.bytes(coins.data(), coins.size());
    const size_t slot = info_.rng_redraw_bytes;
    if (slot != 0) {
      size_t checked = 0;
      for (;;) {
        size_t redraw = 0;
        for (; checked + slot <= coins.size(); checked += slot)
          if (!f_.of_bytes_field(&coins[checked]).has_value()) ++redraw;
        if (redraw == 0) break;
        const size_t at = coins.size();
        coins.resize(at + redraw * slot);
        rng.bytes(&coins[at], redraw * slot);
      }
    }
    const size_t at = coins.size(), nonce_bytes = info_.rng_bytes - info_.rng_sample_bytes;
    coins.resize(at + nonce_bytes);
    rng.bytes(&coins[at], nonce_bytes);
  }

  const proofs::Circuit<Field>& c_;
  const Field& f_;
  lf_circuit* circ_ = nullptr;
  lf_circuit_info info_{};
  std::vector<uint8_t> lfc1_;
};

/* Drop-in for the reference's ZkVerifier<Field, RSFactory> (lib/zk/zk_verifier.h:41-111): the same two
 * constructors and the same
 *     void recv_commitment(const ZkProof<Field>&, Transcript&) const;
 *     bool verify(const ZkProof<Field>&, const Dense<Field>& pub, Transcript&) const;
 * recv_commitment is the reference's one transcript write (lib/ligero/ligero_transcript.h:31-34) on the
 * caller's Transcript; verify serialises the ZkProof with the reference's own ZkProof::write, carries the
 * Transcript across the C ABI and back (lf_zk_verify_committed_batch) and returns the device's verdict.
 * run_mdoc_verifier (lib/circuits/mdoc/mdoc_zk.cc:549-716: two verifiers over two fields on ONE transcript,
 * MAC key drawn between recv_commitment and verify) compiles unchanged with `ZkVerifier` naming this class:
 * oracle/ref_build/ref_mdoc_gpu.cc with -DLF_GPU_VERIFIER (make mdoc_gpu -> libref_mdoc_gpuv.so). */
template <class Field, class RSFactory>
class ZkVerifierGpu {
 public:
  ZkVerifierGpu(const proofs::Circuit<Field>& c, const RSFactory&, size_t rate, size_t nreq, const Field& F)
      : c_(c), f_(F), rate_(rate), nreq_(nreq), block_enc_(0) {}
  ZkVerifierGpu(const proofs::Circuit<Field>& c, const RSFactory&, size_t rate, size_t nreq, size_t block_enc,
                const Field& F)
      : c_(c), f_(F), rate_(rate), nreq_(nreq), block_enc_(block_enc) {}

  void recv_commitment(const proofs::ZkProof<Field>& zk, proofs::Transcript& t) const {
    t.write(zk.com.root.data, proofs::Digest::kLength);
  }

  bool verify(const proofs::ZkProof<Field>& zk, const proofs::Dense<Field>& pub, proofs::Transcript& tv) const {
    /* the ZkProof was constructed with the verifier's parameters (it could not have been read otherwise);
     * its LigeroParam holds the resolved block_enc, which is the key the prover's device copy is cached under */
    proofs::check(zk.param.rateinv == rate_ && zk.param.nreq == nreq_ &&
                      (block_enc_ == 0 || zk.param.block_enc == block_enc_),
                  "ZkProof and ZkVerifier were constructed with different Ligero parameters");
    lf_circuit* circ = cached_circuit(c_.id, (int)LfFieldId<Field>::value, rate_, nreq_, zk.param.block_enc,
                                      &ZkVerifierGpu::serialize, const_cast<ZkVerifierGpu*>(this));
    lfc1_.clear();
    lfc1_.shrink_to_fit();
    lf_circuit_info info;
    lf_check(lf_circuit_get_info(circ, &info));
    proofs::check(info.block_enc == zk.param.block_enc && info.nrow == zk.param.nrow &&
                      info.block == zk.param.block,
                  "Ligero parameters of the device circuit differ from the ZkProof's");
    std::vector<uint8_t> bytes;
    zk.write(bytes, f_);
    std::vector<uint8_t> pubb(std::max<size_t>(info.npub_in * Field::kBytes, 1));
    for (size_t k = 0; k < info.npub_in; ++k) f_.to_bytes_field(&pubb[k * Field::kBytes], pub.v_[k]);
    lf_transcript ts;
    transcript_export(tv, &ts);
    const size_t len = bytes.size();
    int st = 0, why = 0;
    lf_check(lf_zk_verify_committed_batch(circ, 1, info.npub_in ? pubb.data() : nullptr, bytes.data(), len, &len, &ts,
                                          &st, &why));
    transcript_import(tv, ts);
    return st == LF_OK;
  }

 private:
  static const std::vector<uint8_t>& serialize(void* self) {
    auto* p = static_cast<ZkVerifierGpu*>(self);
    proofs::CircuitWriter<Field>(p->f_, LfFieldId<Field>::value).to_bytes(p->c_, p->lfc1_);
    return p->lfc1_;
  }
  const proofs::Circuit<Field>& c_;
  const Field& f_;
  const size_t rate_, nreq_, block_enc_;
  mutable std::vector<uint8_t> lfc1_;
};

}  // namespace longfellow_b200

#endif /* LONGFELLOW_B200_ADAPTERS_H_ */
