// oracle/_ref/libref.so -- TEST INFRASTRUCTURE (see ref_common.cc).
// Workload generator: the reference's own ECDSA P-256 verification circuit and
// witness (circuits/ecdsa/verify_test.cc:349-407 make_circuit / fill_input).
#include "circuits/ecdsa/verify_test.cc"

#include <cstdlib>
#include <cstring>

#include "proto/circuit_io.h"
#include "proto/circuit_writer.h"

namespace proofs {
namespace {
int ecdsa_circuit_impl(size_t nsigs, uint8_t** circ, size_t* circ_len,
                       uint8_t** wit, size_t* wit_len) {
  auto c = make_circuit(nsigs, p256_base);
  Dense<Fp256Base> W(1, c->ninputs);
  fill_input(W, nsigs, p256_base);
  std::vector<uint8_t> bytes;
  CircuitWriter<Fp256Base> wr(p256_base, P256_ID);
  wr.to_bytes(*c, bytes);
  *circ_len = bytes.size();
  *circ = (uint8_t*)malloc(bytes.size());
  memcpy(*circ, bytes.data(), bytes.size());
  *wit_len = c->ninputs * Fp256Base::kBytes;
  *wit = (uint8_t*)malloc(*wit_len);
  for (size_t i = 0; i < c->ninputs; ++i)
    p256_base.to_bytes_field(*wit + i * Fp256Base::kBytes, W.v_[i]);
  return 0;
}
// fill_input (verify_test.cc:378-407) with signature i of the witness taken from
// P256_TEST[(first + i) % ntests] instead of always P256_TEST[0]
int ecdsa_witness_impl(size_t nsigs, size_t first, uint8_t* wit, size_t wit_cap) {
  using Nat = Fp256Base::N;
  using Elt = Fp256Base::Elt;
  using Verw = VerifyWitness3<P256, Fp256Scalar>;
  const size_t nt = sizeof(P256_TEST) / sizeof(P256_TEST[0]);
  auto c = make_circuit(nsigs, p256_base);
  Dense<Fp256Base> W(1, c->ninputs);
  DenseFiller<Fp256Base> filler(W);
  filler.push_back(p256_base.one());
  std::vector<Verw> vws;
  for (size_t i = 0; i < nsigs; ++i) {
    const auto& t = P256_TEST[(first + i) % nt];
    Elt pk_x = p256_base.of_string(t.pk_x), pk_y = p256_base.of_string(t.pk_y);
    Nat e = Nat(t.e), r = Nat(t.r), s = Nat(t.s);
    vws.emplace_back(p256_scalar, p256);
    vws.back().compute_witness(pk_x, pk_y, e, r, s);
    filler.push_back(pk_x);
    filler.push_back(pk_y);
    filler.push_back(p256_base.to_montgomery(e));
  }
  for (size_t i = 0; i < nsigs; ++i) vws[i].fill_witness(filler);
  if (c->ninputs * Fp256Base::kBytes > wit_cap) return -2;
  for (size_t i = 0; i < c->ninputs; ++i) p256_base.to_bytes_field(wit + i * Fp256Base::kBytes, W.v_[i]);
  return (int)c->ninputs;
}
}  // namespace
}  // namespace proofs

extern "C" int ref_ecdsa_ntests() { return (int)(sizeof(proofs::P256_TEST) / sizeof(proofs::P256_TEST[0])); }
extern "C" int ref_ecdsa_witness(size_t nsigs, size_t first, uint8_t* wit, size_t wit_cap) {
  return proofs::ecdsa_witness_impl(nsigs, first, wit, wit_cap);
}
extern "C" int ref_ecdsa_circuit(size_t nsigs, uint8_t** circ, size_t* circ_len,
                                 uint8_t** wit, size_t* wit_len) {
  return proofs::ecdsa_circuit_impl(nsigs, circ, circ_len, wit, wit_len);
}
