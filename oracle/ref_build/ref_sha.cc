// oracle/_ref/libref.so -- TEST INFRASTRUCTURE (see ref_common.cc).
// Workload generator: the reference's own SHA-256 circuit builder and witness
// filler, reached by including its test file unmodified
// (circuits/sha/flatsha256_circuit_test.cc:366-468 make_circuit / fill_input).
#include "circuits/sha/flatsha256_circuit_test.cc"

#include <cstdlib>
#include <cstring>

#include "proto/circuit_io.h"
#include "proto/circuit_writer.h"

extern "C" {
// Builds the nblocks-block flatsha256 circuit over GF(2^128) (plucker 2, as
// BM_ShaZK_fp2_128) and its benchmark witness.  Buffers are malloc'd; free
// with ref_free.
int ref_sha_circuit(size_t nblocks, uint8_t** circ, size_t* circ_len,
                    uint8_t** wit, size_t* wit_len) {
  using namespace proofs;
  using F = GF2_128<>;
  static const F Fs;
  auto c = bench::make_circuit<F, 2>(nblocks, 1, Fs);
  Dense<F> W(1, c->ninputs);
  bench::fill_input<F, 2>(W, nblocks, c->ninputs, 1, Fs);
  std::vector<uint8_t> bytes;
  CircuitWriter<F> wr(Fs, GF2_128_ID);
  wr.to_bytes(*c, bytes);
  *circ_len = bytes.size();
  *circ = (uint8_t*)malloc(bytes.size());
  memcpy(*circ, bytes.data(), bytes.size());
  *wit_len = c->ninputs * F::kBytes;
  *wit = (uint8_t*)malloc(*wit_len);
  for (size_t i = 0; i < c->ninputs; ++i) Fs.to_bytes_field(*wit + i * F::kBytes, W.v_[i]);
  return 0;
}
void ref_free(void* p) { free(p); }
}
