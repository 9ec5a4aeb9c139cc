// oracle/_ref/libref_mdoc.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's mdoc prover (lib/circuits/mdoc/mdoc_zk.cc) with the body of
// run_mdoc_prover (mdoc_zk.cc:398-547) split at the points where the CUDA back
// end takes over, so that tests can drive
//   reference:  fill_witness -> commit(hash) -> commit(sig) -> MAC key -> MACs ->
//               prove(hash) -> prove(sig)
// and the same sequence with lf_zk_commit_batch / lf_zk_prove_committed_batch in
// place of the four ZkProver calls, on the same witness, coins and transcript,
// and compare the proofs byte for byte.  The reference sources are compiled from
// where they lie (this file #includes mdoc_zk.cc to reach its file-local helpers);
// the only change of behaviour is that the commit coins come from a replayed
// byte buffer instead of SecureRandomEngine (the MAC key shares drawn inside
// fill_witness stay random: both flows start from the same filled witness).
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "circuits/mdoc/mdoc_zk.cc"  // NOLINT: reaches fill_witness, compute_macs, update_macs
#include "circuits/mdoc/mdoc_examples.h"
#include "circuits/mdoc/mdoc_test_attributes.h"

namespace proofs {
namespace {

class ReplayEngine : public RandomEngine {
 public:
  ReplayEngine(const uint8_t* p, size_t n) : p_(p), n_(n), pos_(0) {}
  void bytes(uint8_t* buf, size_t n) override {
    check(pos_ + n <= n_, "ReplayEngine exhausted");
    memcpy(buf, p_ + pos_, n);
    pos_ += n;
  }
  size_t consumed() const { return pos_; }

 private:
  const uint8_t* p_;
  size_t n_, pos_;
};

struct MdocCase {
  const ZkSpecStruct* spec;
  std::unique_ptr<Circuit<Fp256Base>> c_sig;
  std::unique_ptr<Circuit<f_128>> c_hash;
  std::unique_ptr<Dense<Fp256Base>> W_sig;
  std::unique_ptr<Dense<f_128>> W_hash;
  ProverState state;
  const MdocTests* test;
  size_t attrs_len;
};

}  // namespace
}  // namespace proofs

using namespace proofs;

extern "C" {

// circuit_raw: the DECOMPRESSED circuit file of kZkSpecs[0] (signature circuit, then hash circuit)
void* ref_mdoc_prepare(const uint8_t* circuit_raw, size_t len) {
  set_log_level(ERROR);
  auto h = std::make_unique<MdocCase>();
  h->spec = &kZkSpecs[0];
  h->test = &mdoc_tests[0];
  h->attrs_len = 1;
  const f_128 Fs;
  ReadBuffer rb(circuit_raw, len);
  CircuitReader<Fp256Base> cr_s(p256_base, P256_ID);
  h->c_sig = cr_s.from_bytes(rb, false);
  if (!h->c_sig) return nullptr;
  CircuitReader<f_128> cr_h(Fs, GF2_128_ID);
  h->c_hash = cr_h.from_bytes(rb, false);
  if (!h->c_hash) return nullptr;
  h->W_sig = std::make_unique<Dense<Fp256Base>>(1, h->c_sig->ninputs);
  h->W_hash = std::make_unique<Dense<f_128>>(1, h->c_hash->ninputs);
  DenseFiller<Fp256Base> sig_filler(*h->W_sig);
  DenseFiller<f_128> hash_filler(*h->W_hash);
  const MdocTests* t = h->test;
  Elt pkX, pkY;
  if (!parsePk(t->pkx.as_pointer, t->pky.as_pointer, pkX, pkY)) return nullptr;
  const RequestedAttribute attrs[1] = {test::age_over_18};
  SecureRandomEngine rng;
  if (fill_witness(sig_filler, hash_filler, t->mdoc, t->mdoc_size, pkX, pkY, t->transcript, t->transcript_size,
                   attrs, 1, (const uint8_t*)t->now, h->state, rng, Fs, h->spec->version) != MDOC_PROVER_SUCCESS)
    return nullptr;
  return h.release();
}
void ref_mdoc_free(void* hv) { delete static_cast<MdocCase*>(hv); }

// out[0..7]: sig ninputs, sig npub, hash ninputs, hash npub, block_enc_sig, block_enc_hash, transcript len, version
void ref_mdoc_info(void* hv, size_t out[8]) {
  auto* h = static_cast<MdocCase*>(hv);
  out[0] = h->c_sig->ninputs;
  out[1] = h->c_sig->npub_in;
  out[2] = h->c_hash->ninputs;
  out[3] = h->c_hash->npub_in;
  out[4] = h->spec->block_enc_sig;
  out[5] = h->spec->block_enc_hash;
  out[6] = h->test->transcript_size;
  out[7] = h->spec->version;
}
void ref_mdoc_transcript(void* hv, uint8_t* out) {
  auto* h = static_cast<MdocCase*>(hv);
  memcpy(out, h->test->transcript, h->test->transcript_size);
}
// the filled witnesses in wire encoding (MAC and a_v inputs still zero)
void ref_mdoc_witness(void* hv, uint8_t* w_sig, uint8_t* w_hash) {
  auto* h = static_cast<MdocCase*>(hv);
  const f_128 Fs;
  for (size_t i = 0; i < h->c_sig->ninputs; ++i)
    p256_base.to_bytes_field(w_sig + i * Fp256Base::kBytes, h->W_sig->v_[i]);
  for (size_t i = 0; i < h->c_hash->ninputs; ++i) Fs.to_bytes_field(w_hash + i * f_128::kBytes, h->W_hash->v_[i]);
}

// mdoc_zk.cc:488-494 on copies of the filled witnesses: MACs of the common inputs under
// the verifier share a_v, patched into the PUBLIC inputs of both circuits
void ref_mdoc_update_macs(void* hv, const uint8_t av_bytes[16], uint8_t* w_sig, uint8_t* w_hash,
                          uint8_t macs_out[96]) {
  auto* h = static_cast<MdocCase*>(hv);
  const f_128 Fs;
  gf2k av = Fs.of_bytes_field(av_bytes).value(), macs[6];
  auto Ws = h->W_sig->clone();
  auto Wh = h->W_hash->clone();
  compute_macs(3, h->state.common, macs, macs_out, h->state.ap, av);
  update_macs(*Ws, *Wh, kSigMacIndex, getHashMacIndex(h->attrs_len, h->spec->version), macs, av, Fs);
  for (size_t i = 0; i < h->c_sig->ninputs; ++i) p256_base.to_bytes_field(w_sig + i * Fp256Base::kBytes, Ws->v_[i]);
  for (size_t i = 0; i < h->c_hash->ninputs; ++i) Fs.to_bytes_field(w_hash + i * f_128::kBytes, Wh->v_[i]);
}

// run_mdoc_prover from "Run prover" on (mdoc_zk.cc:476-535) with replayed commit coins.
// proof_out = [6 MACs][hash proof][sig proof] exactly as run_mdoc_prover serialises it.
int ref_mdoc_prove(void* hv, const uint8_t* coins, size_t coins_len, uint8_t* proof_out, size_t cap,
                   size_t* proof_len, size_t* len_hash_proof, size_t* len_sig_proof, size_t* coins_hash,
                   size_t* coins_total, uint8_t av_out[16]) {
  auto* h = static_cast<MdocCase*>(hv);
  const f_128 Fs;
  const ZkSpecStruct* zk_spec = h->spec;
  auto Ws = h->W_sig->clone();
  auto Wh = h->W_hash->clone();
  ReplayEngine rng(coins, coins_len);
  Transcript tp(h->test->transcript, h->test->transcript_size, zk_spec->version);
  const f2_p256 p256_2(p256_base);
  const Elt2 omega = p256_2.of_string(kRootX, kRootY);
  const FftExtConvolutionFactory fft_b(p256_base, p256_2, omega, 1ull << 31);
  const RSFactory_b rsf_b(fft_b, p256_base);
  const RSFactory the_reed_solomon_factory(Fs);
  size_t r = zk_spec->version < 7 ? kLigeroRate : kLigeroRatev7;
  size_t req = zk_spec->version < 7 ? kLigeroNreq : kLigeroNreqv7;
  ZkProof<f_128> h_zk(*h->c_hash, r, req, zk_spec->block_enc_hash);
  ZkProof<Fp256Base> sig_zk(*h->c_sig, r, req, zk_spec->block_enc_sig);
  ZkProver<f_128, RSFactory> hash_p(*h->c_hash, Fs, the_reed_solomon_factory);
  ZkProver<Fp256Base, RSFactory_b> sig_p(*h->c_sig, p256_base, rsf_b);
  hash_p.commit(h_zk, *Wh, tp, rng);
  *coins_hash = rng.consumed();
  sig_p.commit(sig_zk, *Ws, tp, rng);
  *coins_total = rng.consumed();
  gf2k av = generate_mac_key(tp), macs[6];
  Fs.to_bytes_field(av_out, av);
  uint8_t macs_b[6 * f_128::kBytes];
  compute_macs(3, h->state.common, macs, macs_b, h->state.ap, av);
  update_macs(*Ws, *Wh, kSigMacIndex, getHashMacIndex(h->attrs_len, zk_spec->version), macs, av, Fs);
  if (!hash_p.prove(h_zk, *Wh, tp)) return -3;
  if (!sig_p.prove(sig_zk, *Ws, tp)) return -3;
  std::vector<uint8_t> buf;
  buf.insert(buf.begin(), macs_b, macs_b + 6 * f_128::kBytes);
  h_zk.write(buf, Fs);
  *len_hash_proof = buf.size() - 6 * f_128::kBytes;
  sig_zk.write(buf, p256_base);
  *len_sig_proof = buf.size() - 6 * f_128::kBytes - *len_hash_proof;
  *proof_len = buf.size();
  if (buf.size() > cap) return -4;
  memcpy(proof_out, buf.data(), buf.size());
  return 0;
}
size_t ref_mdoc_ligero_params(size_t* rate, size_t* nreq) {
  *rate = kLigeroRatev7;
  *nreq = kLigeroNreqv7;
  return 0;
}

}  // extern "C"
