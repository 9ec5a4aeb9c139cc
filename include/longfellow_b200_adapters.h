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

#include "arrays/dense.h"
#include "longfellow_b200.h"
#include "proto/circuit_io.h"
#include "proto/circuit_writer.h"
#include "random/random.h"
#include "sumcheck/circuit.h"
#include "util/panic.h"

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

}  // namespace longfellow_b200

#endif /* LONGFELLOW_B200_ADAPTERS_H_ */
