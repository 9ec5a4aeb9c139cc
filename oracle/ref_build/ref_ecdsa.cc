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
}  // namespace
}  // namespace proofs

extern "C" int ref_ecdsa_circuit(size_t nsigs, uint8_t** circ, size_t* circ_len,
                                 uint8_t** wit, size_t* wit_len) {
  return proofs::ecdsa_circuit_impl(nsigs, circ, circ_len, wit, wit_len);
}
