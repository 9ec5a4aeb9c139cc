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
// The witness of the nblocks-block circuit for an arbitrary message (fill_input's steps,
// flatsha256_circuit_test.cc:415-468, with the message and its SHA-256 supplied by the caller instead
// of the benchmark's "aaa..." strings).  msg must pad to exactly nblocks blocks; wit = ninputs x 16 bytes.
int ref_sha_witness(size_t nblocks, const uint8_t* msg, size_t len, const uint8_t hash[32], uint8_t* wit,
                    size_t wit_cap) {
  using namespace proofs;
  using F = GF2_128<>;
  static const F Fs;
  if ((len + 9 + 63) / 64 != nblocks) return -1;
  uint8_t numb;
  std::vector<uint8_t> inb(64 * nblocks);
  std::vector<FlatSHA256Witness::BlockWitness> bwb(nblocks);
  FlatSHA256Witness::transform_and_witness_message(len, msg, nblocks, numb, &inb[0], &bwb[0]);
  std::vector<F::Elt> W;
  auto bit = [&](bool b) { W.push_back(b ? Fs.one() : Fs.zero()); };
  W.push_back(Fs.one());
  for (size_t i = 0; i < 8; ++i) bit((numb >> i) & 1);
  for (size_t j = 0; j < nblocks * 64; ++j)
    for (size_t i = 0; i < 8; ++i) bit((inb[j] >> i) & 1);
  for (size_t j = 0; j < 256; ++j) bit((hash[(255 - j) / 8] >> (j % 8)) & 1);
  BitPluckerEncoder<F, 2> BPENC(Fs);
  auto pk = [&](uint32_t v) {
    auto a = BPENC.mkpacked_v32(v);
    for (auto& e : a) W.push_back(e);
  };
  for (size_t j = 0; j < nblocks; ++j) {
    for (size_t k = 0; k < 48; ++k) pk(bwb[j].outw[k]);
    for (size_t k = 0; k < 64; ++k) {
      pk(bwb[j].oute[k]);
      pk(bwb[j].outa[k]);
    }
    for (size_t k = 0; k < 8; ++k) pk(bwb[j].h1[k]);
  }
  if (W.size() * F::kBytes > wit_cap) return -2;
  for (size_t i = 0; i < W.size(); ++i) Fs.to_bytes_field(wit + i * F::kBytes, W[i]);
  return (int)W.size();
}
void ref_free(void* p) { free(p); }
}
