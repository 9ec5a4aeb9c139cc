// oracle/_ref/libref_mdoc_gpu.so -- the reference's mdoc prover compiled UNCHANGED against the CUDA back end.
//
// lib/circuits/mdoc/mdoc_zk.cc is included where it lies; the one substitution is that the name `ZkProver`
// inside it resolves to longfellow_b200::ZkProverGpu (include/longfellow_b200_adapters.h), a class with
// ZkProver's constructor and commit/prove signatures that runs on the GPU through the C ABI.  Everything
// else of run_mdoc_prover -- CBOR parsing, witness filling, the MAC arithmetic, the Transcript, the
// SecureRandomEngine, proof serialisation through ZkProof::write -- is the reference's own code, and so
// is run_mdoc_verifier (ZkVerifier is not substituted unless LF_GPU_VERIFIER is defined, below), which the
// tests use to check the GPU's proofs.
// Exports run_mdoc_prover / run_mdoc_verifier with the signatures of lib/circuits/mdoc/mdoc_zk.h:157-189,
// plus a small driver over the (claim, mdoc) pairs of lib/circuits/mdoc/mdoc_zk_test.cc:119-170.
#include "zk/zk_prover.h"  // the reference's own ZkProver is defined first, under its own name

#include "ec/p256.h"
#include "gf2k/gf2_128.h"
#include "longfellow_b200_adapters.h"
#include "proto/circuit_io.h"

namespace longfellow_b200 {
template <>
struct LfFieldId<proofs::GF2_128<>> {
  static constexpr proofs::FieldID value = proofs::GF2_128_ID;
};
template <>
struct LfFieldId<proofs::Fp256Base> {
  static constexpr proofs::FieldID value = proofs::P256_ID;
};
}  // namespace longfellow_b200

// -DLF_GPU_VERIFIER (oracle/_ref/libref_mdoc_gpuv.so): `ZkVerifier` resolves to longfellow_b200::ZkVerifierGpu
// as well, so run_mdoc_verifier -- CBOR/transcript handling, MAC checks and the two ZkVerifier objects on one
// transcript, all the reference's own code -- verifies on the GPU too.  The tests load both libraries: proofs
// made by either are checked by the reference's verifier (this file without the flag) and by the GPU's.
#include "zk/zk_verifier.h"  // the reference's own ZkVerifier is defined first, under its own name
#define ZkProver ::longfellow_b200::ZkProverGpu
#ifdef LF_GPU_VERIFIER
#define ZkVerifier ::longfellow_b200::ZkVerifierGpu
#endif
#include "circuits/mdoc/mdoc_zk.cc"  // NOLINT
#undef ZkProver
#ifdef LF_GPU_VERIFIER
#undef ZkVerifier
#endif

#include "circuits/mdoc/mdoc_examples.h"
#include "circuits/mdoc/mdoc_test_attributes.h"

namespace proofs {
namespace {
struct Claim {
  const char* name;
  RequestedAttribute attr;
  size_t mdoc;
};
// lib/circuits/mdoc/mdoc_zk_test.cc:119-170 (TEST_F(MdocZKTest, one_claim))
const Claim kClaims[] = {
    {"+18-mdoc[0]", test::age_over_18, 0},
    {"+18-mdoc[1]", test::age_over_18, 1},
    {"+18-mdoc[2]", test::age_over_18, 2},
    {"+18-mdoc[9]", test::europa_age_over_18, 9},
    {"familyname_mustermann-mdoc[3]", test::familyname_mustermann, 3},
    {"birthdate_1971_09_01-mdoc[3]", test::birthdate_1971_09_01, 3},
    {"height_175-mdoc[3]", test::height_175, 3},
    {"birthdate_1998_09_04-idpass-mdoc[4]", test::birthdate_1998_09_04, 4},
    {"age_over_18-website-mdoc[5]", test::age_over_18, 5},
    {"not_over_18-large-mdoc[6]", test::not_over_18, 6},
    {"age_birth_year-mdoc[8]", test::age_birth_year, 8},
    {"DHS_compliance-mdoc[10]", test::aamva_dhs_compliance, 10},
    {"Sparkasse_Age-mdoc[11]", test::age_over_18, 11},
    {"MT_Prod_Age_Over_18-mdoc[12]", test::age_over_18, 12},
    {"MT_Prod_Age_Over_18-mdoc[14]", test::age_over_18, 14},
    {"AZ_Prod_Age_Over_18-mdoc[13]", test::age_over_18, 13},
    {"EUAV_Age_Over_18-mdoc[15]", test::age_over_18, 15},
    {"EUAV_Age_Over_18-mdoc[16]", test::age_over_18, 16},
    {"EUAV_Age_Over_18-mdoc[17]", test::age_over_18, 17},
    {"EUAV_Age_Over_18-mdoc[18]", test::age_over_18, 18},
    {"EUAV_Age_Over_18-mdoc[19]", test::age_over_18, 19},
    {"EUAV_Age_Over_18-mdoc[20]", test::age_over_18, 20},
    {"EUAV_Age_Over_18-mdoc[21]", test::age_over_18, 21},
    {"EUAV_Age_Over_18-mdoc[22]", test::age_over_18, 22},
    {"EUAV_Age_Over_18-mdoc[23]", test::age_over_18, 23},
    {"EUAV_Age_Over_18-mdoc[24]", test::age_over_18, 24},
    {"Aadhaar_age_above18-mdoc[25]", test::age_above18, 25},
};
}  // namespace
}  // namespace proofs

using namespace proofs;

extern "C" {
size_t ref_mdoc_gpu_nclaims() { return sizeof(kClaims) / sizeof(kClaims[0]); }
const char* ref_mdoc_gpu_claim_name(size_t i) { return kClaims[i].name; }

// MdocZKTest::run_test (mdoc_zk_test.cc:60-94) for claim i: run_mdoc_prover (on the GPU through
// ZkProverGpu), then the reference's run_mdoc_verifier on what it produced.  circuit = the zstd-compressed
// circuit file of kZkSpecs[0].  Returns prover code * 1000 + verifier code (0 = both succeeded);
// *proof_len receives the proof size.  tamper != 0 flips one byte of the proof before verifying.
int ref_mdoc_gpu_run_claim(size_t i, const uint8_t* circuit, size_t circuit_len, size_t* proof_len, int tamper) {
  set_log_level(ERROR);
  const Claim& c = kClaims[i];
  const MdocTests* t = &mdoc_tests[c.mdoc];
  const ZkSpecStruct zk_spec = kZkSpecs[0];
  RequestedAttribute attrs[1] = {c.attr};
  uint8_t* zkproof = nullptr;
  size_t len = 0;
  MdocProverErrorCode pr = run_mdoc_prover(circuit, circuit_len, t->mdoc, t->mdoc_size, t->pkx.as_pointer,
                                           t->pky.as_pointer, t->transcript, t->transcript_size, attrs, 1,
                                           (const char*)t->now, &zkproof, &len, &zk_spec);
  if (proof_len) *proof_len = len;
  if (pr != MDOC_PROVER_SUCCESS) return 1000 * (int)pr;
  if (tamper) zkproof[len / 2] ^= 1;
  MdocVerifierErrorCode vr = run_mdoc_verifier(circuit, circuit_len, t->pkx.as_pointer, t->pky.as_pointer,
                                               t->transcript, t->transcript_size, attrs, 1, (const char*)t->now,
                                               zkproof, len, t->doc_type, &zk_spec);
  free(zkproof);
  return (int)vr;
}

// the two halves separately, so that a proof made by one library can be checked by the other:
// run_mdoc_prover for claim i into out (cap bytes); returns the prover's code
int ref_mdoc_gpu_prove_claim(size_t i, const uint8_t* circuit, size_t circuit_len, uint8_t* out, size_t cap,
                             size_t* proof_len) {
  set_log_level(ERROR);
  const Claim& c = kClaims[i];
  const MdocTests* t = &mdoc_tests[c.mdoc];
  const ZkSpecStruct zk_spec = kZkSpecs[0];
  RequestedAttribute attrs[1] = {c.attr};
  uint8_t* zkproof = nullptr;
  size_t len = 0;
  MdocProverErrorCode pr = run_mdoc_prover(circuit, circuit_len, t->mdoc, t->mdoc_size, t->pkx.as_pointer,
                                           t->pky.as_pointer, t->transcript, t->transcript_size, attrs, 1,
                                           (const char*)t->now, &zkproof, &len, &zk_spec);
  if (pr != MDOC_PROVER_SUCCESS) return (int)pr;
  *proof_len = len;
  if (len > cap) {
    free(zkproof);
    return -1;
  }
  memcpy(out, zkproof, len);
  free(zkproof);
  return 0;
}
// run_mdoc_verifier for claim i on the given proof bytes; returns the verifier's code
int ref_mdoc_gpu_verify_claim(size_t i, const uint8_t* circuit, size_t circuit_len, const uint8_t* proof,
                              size_t proof_len) {
  set_log_level(ERROR);
  const Claim& c = kClaims[i];
  const MdocTests* t = &mdoc_tests[c.mdoc];
  const ZkSpecStruct zk_spec = kZkSpecs[0];
  RequestedAttribute attrs[1] = {c.attr};
  return (int)run_mdoc_verifier(circuit, circuit_len, t->pkx.as_pointer, t->pky.as_pointer, t->transcript,
                                t->transcript_size, attrs, 1, (const char*)t->now, proof, proof_len, t->doc_type,
                                &zk_spec);
}
// 1 when run_mdoc_verifier of this library runs on the GPU
int ref_mdoc_gpu_verifier_is_gpu() {
#ifdef LF_GPU_VERIFIER
  return 1;
#else
  return 0;
#endif
}
}  // extern "C"
