// oracle/_ref/libref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" wrapper around the UNMODIFIED reference implementation,
// compiled from the sources where they lie under /root/reference/lib (nothing
// is copied into this repository).  It exists so that tests/ and bench.py's
// cpu_baseline / --impl reference legs can
//   * pin the plain-C restatement in oracle/ against the real reference, and
//   * run the reference's own ZkProver on the same witness, RNG bytes and
//     transcript seed as the CUDA path, for byte-exact proof comparison.
// Nothing under longfellow_zk_b200/ may link or load this file.
//
// Reference entry points used (file:line under /root/reference/lib):
//   zk/zk_prover.h:72-149        ZkProver::commit / prove
//   zk/zk_proof.h:90-112         ZkProof::write / read
//   zk/zk_verifier.h:39-106      ZkVerifier
//   gf2k/lch14_reed_solomon.h    LCH14ReedSolomon::interpolate
//   gf2k/lch14.h                 LCH14::FFT/IFFT/BidirectionalFFT
//   merkle/merkle_commitment.h   MerkleCommitment::commit/open
//   random/transcript.h          Transcript
//   proto/circuit_reader.h       CircuitReader::from_bytes

#define private public
#define protected public
#include "ligero/ligero_prover.h"
#include "zk/zk_prover.h"
#undef private
#undef protected

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "algebra/convolution.h"
#include "algebra/fft.h"
#include "algebra/fp.h"
#include "algebra/fp2.h"
#include "algebra/fp_p128.h"
#include "algebra/reed_solomon.h"
#include "arrays/dense.h"
#include "ec/p256.h"
#include "gf2k/gf2_128.h"
#include "gf2k/lch14.h"
#include "gf2k/lch14_reed_solomon.h"
#include "ligero/ligero_param.h"
#include "merkle/merkle_commitment.h"
#include "merkle/merkle_tree.h"
#include "proto/circuit_io.h"
#include "proto/circuit_reader.h"
#include "random/random.h"
#include "random/transcript.h"
#include "sumcheck/circuit.h"
#include "util/log.h"
#include "util/readbuffer.h"
#include "zk/zk_proof.h"
#include "zk/zk_verifier.h"

#ifdef LF_WITH_GPU_ADAPTERS
#include "longfellow_b200_adapters.h"
#endif

namespace proofs {

// RandomEngine that replays caller-supplied bytes, so that the reference and
// the CUDA path can be fed the very same coins.
class BufferRandomEngine : public RandomEngine {
 public:
  BufferRandomEngine(const uint8_t* p, size_t n) : p_(p), n_(n), pos_(0) {}
  void bytes(uint8_t* buf, size_t n) override {
    check(pos_ + n <= n_, "BufferRandomEngine exhausted");
    memcpy(buf, p_ + pos_, n);
    pos_ += n;
  }
  size_t consumed() const { return pos_; }

 private:
  const uint8_t* p_;
  size_t n_, pos_;
};

using GF = GF2_128<>;
static const GF& gf() {
  static const GF f;
  return f;
}

static constexpr char kRootX[] =
    "112649224146410281873500457609690258373018840430489408729223714171582664"
    "680802";
static constexpr char kRootY[] =
    "840879943585409076957404614278186605601821689971823787493130182544504602"
    "12908";

using F2P256 = Fp2<Fp256Base>;
using P256Conv = FFTExtConvolutionFactory<Fp256Base, F2P256>;
using P256RS = ReedSolomonFactory<Fp256Base, P256Conv>;
struct P256Ctx {
  F2P256 f2;
  F2P256::Elt omega;
  P256Conv conv;
  P256RS rs;
  P256Ctx()
      : f2(p256_base),
        omega(f2.of_string(kRootX, kRootY)),
        conv(p256_base, f2, omega, 1ull << 31),
        rs(conv, p256_base) {}
};
static const P256Ctx& p256ctx() {
  static const P256Ctx c;
  return c;
}

struct Dump {
  // optional stage dumps; all byte strings use to_bytes_field encoding
  std::vector<uint8_t> witness;  // private inputs || pad (ligero witness)
  std::vector<uint8_t> tableau;  // nrow * block_enc elements
  std::vector<uint8_t> root;     // 32
  std::vector<uint8_t> sumcheck; // serialized sumcheck proof
};

struct CircuitHandle {
  int field_id;
  std::unique_ptr<Circuit<GF>> gf;
  std::unique_ptr<Circuit<Fp256Base>> p256;
};

template <class Field, class RSF>
static int zk_prove_t(const Field& F, const RSF& rsf,
                      const Circuit<Field>* c,
                      const uint8_t* wit, const uint8_t* rng, size_t rng_len,
                      const uint8_t* tinit, size_t tinit_len, size_t rate,
                      size_t nreq, size_t block_enc, uint8_t* out,
                      size_t out_cap, size_t* out_len, size_t* rng_used,
                      Dump* dump) {
  set_log_level(ERROR);
  Dense<Field> W(1, c->ninputs);
  for (size_t i = 0; i < c->ninputs; ++i) {
    auto e = F.of_bytes_field(wit + i * Field::kBytes);
    if (!e.has_value()) return -2;
    W.v_[i] = e.value();
  }
  BufferRandomEngine eng(rng, rng_len);
  Transcript tp(tinit, tinit_len);
  std::unique_ptr<ZkProof<Field>> zkp;
  if (block_enc == 0) {
    zkp = std::make_unique<ZkProof<Field>>(*c, rate, nreq);
  } else {
    zkp = std::make_unique<ZkProof<Field>>(*c, rate, nreq, block_enc);
  }
  ZkProver<Field, RSF> prover(*c, F, rsf);
  prover.commit(*zkp, W, tp, eng);
  if (rng_used) *rng_used = eng.consumed();
  if (dump) {
    dump->witness.resize(prover.witness_.size() * Field::kBytes);
    for (size_t i = 0; i < prover.witness_.size(); ++i)
      F.to_bytes_field(&dump->witness[i * Field::kBytes], prover.witness_[i]);
    auto& tab = prover.lp_->tableau_;
    dump->tableau.resize(tab.size() * Field::kBytes);
    for (size_t i = 0; i < tab.size(); ++i)
      F.to_bytes_field(&dump->tableau[i * Field::kBytes], tab[i]);
    dump->root.assign(zkp->com.root.data, zkp->com.root.data + 32);
  }
  if (!prover.prove(*zkp, W, tp)) return -3;
  std::vector<uint8_t> buf;
  zkp->write(buf, F);
  if (dump) {
    zkp->write_sc_proof(zkp->proof, dump->sumcheck, F);
  }
  *out_len = buf.size();
  if (buf.size() > out_cap) return -4;
  memcpy(out, buf.data(), buf.size());
  return 0;
}

template <class Field, class RSF>
static int zk_verify_t(const Field& F, const RSF& rsf,
                       const Circuit<Field>* c,
                       const uint8_t* pub, const uint8_t* tinit,
                       size_t tinit_len, size_t rate, size_t nreq,
                       size_t block_enc, const uint8_t* proof,
                       size_t proof_len) {
  set_log_level(ERROR);
  Dense<Field> P(1, c->npub_in > 0 ? c->npub_in : 1);
  for (size_t i = 0; i < c->npub_in; ++i) {
    auto e = F.of_bytes_field(pub + i * Field::kBytes);
    if (!e.has_value()) return -2;
    P.v_[i] = e.value();
  }
  std::unique_ptr<ZkProof<Field>> zkp;
  if (block_enc == 0) {
    zkp = std::make_unique<ZkProof<Field>>(*c, rate, nreq);
  } else {
    zkp = std::make_unique<ZkProof<Field>>(*c, rate, nreq, block_enc);
  }
  ReadBuffer pb(proof, proof_len);
  if (!zkp->read(pb, F)) return 1;
  if (pb.remaining() != 0) return 2;
  Transcript tv(tinit, tinit_len);
  if (block_enc == 0) {
    ZkVerifier<Field, RSF> ver(*c, rsf, rate, nreq, F);
    ver.recv_commitment(*zkp, tv);
    return ver.verify(*zkp, P, tv) ? 0 : 3;
  }
  ZkVerifier<Field, RSF> ver(*c, rsf, rate, nreq, block_enc, F);
  ver.recv_commitment(*zkp, tv);
  return ver.verify(*zkp, P, tv) ? 0 : 3;
}

}  // namespace proofs

using namespace proofs;

extern "C" {

// ---------------- GF(2^128) primitives ----------------
// gf2k/gf2_128.h:227-246, gf2k/sysdep.h:51-66
void ref_gf128_mul(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  const GF& F = gf();
  for (size_t i = 0; i < n; ++i) {
    auto x = F.of_bytes_field(a + 16 * i).value();
    auto y = F.of_bytes_field(b + 16 * i).value();
    F.to_bytes_field(out + 16 * i, F.mulf(x, y));
  }
}
void ref_gf128_invert(const uint8_t* a, uint8_t* out, size_t n) {
  const GF& F = gf();
  for (size_t i = 0; i < n; ++i) {
    auto x = F.of_bytes_field(a + 16 * i).value();
    F.to_bytes_field(out + 16 * i, F.invertf(x));
  }
}
// gf2k/gf2_128.h:151-160
void ref_gf128_of_scalar(const uint64_t* u, uint8_t* out, size_t n) {
  const GF& F = gf();
  for (size_t i = 0; i < n; ++i) F.to_bytes_field(out + 16 * i, F.of_scalar(u[i]));
}
// gf2k/gf2_128.h:216-224 ; returns 0xFFFFFFFF for elements outside GF(2^16)
void ref_gf128_subfield_index(const uint8_t* a, uint32_t* out, size_t n) {
  const GF& F = gf();
  for (size_t i = 0; i < n; ++i) {
    auto x = F.of_bytes_field(a + 16 * i).value();
    if (F.in_subfield(x)) {
      uint8_t b[2];
      F.to_bytes_subfield(b, x);
      out[i] = b[0] | (b[1] << 8);
    } else {
      out[i] = 0xFFFFFFFFu;
    }
  }
}
// constants: beta[16], poly_evaluation_point[6], newton_denominator, g, invx
void ref_gf128_constants(uint8_t* beta /*16*16*/, uint8_t* pts /*6*16*/,
                         uint8_t* newton /*6*6*16*/) {
  const GF& F = gf();
  for (size_t i = 0; i < 16; ++i) F.to_bytes_field(beta + 16 * i, F.beta(i));
  for (size_t i = 0; i < 6; ++i)
    F.to_bytes_field(pts + 16 * i, F.poly_evaluation_point(i));
  memset(newton, 0, 6 * 6 * 16);
  for (size_t k = 1; k < 6; ++k)
    for (size_t i = 1; i <= k; ++i)
      F.to_bytes_field(newton + 16 * (k * 6 + i), F.newton_denominator(k, i));
}

// ---------------- LCH14 ----------------
// gf2k/lch14.h:106-149 ; op 0 = FFT, 1 = IFFT, 2 = Bidirectional(k)
void ref_lch14(int op, size_t l, size_t coset_or_k, uint8_t* B) {
  const GF& F = gf();
  static const LCH14<GF> fft(F);
  size_t n = size_t(1) << l;
  std::vector<GF::Elt> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = F.of_bytes_field(B + 16 * i).value();
  if (op == 0) fft.FFT(l, coset_or_k, v.data());
  else if (op == 1) fft.IFFT(l, coset_or_k, v.data());
  else fft.BidirectionalFFT(l, coset_or_k, v.data());
  for (size_t i = 0; i < n; ++i) F.to_bytes_field(B + 16 * i, v[i]);
}
void ref_lch14_what(uint8_t* out /*16*16*16*/) {
  const GF& F = gf();
  static const LCH14<GF> fft(F);
  for (size_t i = 0; i < 16; ++i)
    for (size_t j = 0; j < 16; ++j)
      F.to_bytes_field(out + 16 * (i * 16 + j), fft.WHat_DEBUG(i, j));
}
// gf2k/lch14_reed_solomon.h:49-103 ; rows: nrows x m elements, first n valid
void ref_lch14_interpolate(size_t n, size_t m, uint8_t* rows, size_t nrows) {
  const GF& F = gf();
  LCH14ReedSolomon<GF> rs(n, m, F);
  std::vector<GF::Elt> v(m);
  for (size_t r = 0; r < nrows; ++r) {
    uint8_t* p = rows + r * m * 16;
    for (size_t i = 0; i < n; ++i) v[i] = F.of_bytes_field(p + 16 * i).value();
    rs.interpolate(v.data());
    for (size_t i = 0; i < m; ++i) F.to_bytes_field(p + 16 * i, v[i]);
  }
}

// seconds per LCH14ReedSolomon(n, m)::interpolate of one row on one thread (BM_ReedSolomon_gf128)
double ref_lch14_rs_bench(size_t n, size_t m, int reps) {
  const GF& F = gf();
  LCH14ReedSolomon<GF> rs(n, m, F);
  std::vector<GF::Elt> v(m);
  for (size_t i = 0; i < n; ++i) v[i] = F.of_scalar(i + 1);
  rs.interpolate(v.data());
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) rs.interpolate(v.data());
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}

// ---------------- Merkle ----------------
// merkle/merkle_tree.h:100-151 ; leaves: n x 32 ; nodes_out: 2n x 32 (heap)
void ref_merkle_build(size_t n, const uint8_t* leaves, uint8_t* nodes_out,
                      uint8_t* root_out) {
  MerkleTree mt(n);
  for (size_t i = 0; i < n; ++i) {
    Digest d;
    memcpy(d.data, leaves + 32 * i, 32);
    mt.set_leaf(i, d);
  }
  Digest r = mt.build_tree();
  memcpy(root_out, r.data, 32);
  if (nodes_out) {
    memset(nodes_out, 0, 32);
    for (size_t i = 1; i < 2 * n; ++i) memcpy(nodes_out + 32 * i, mt.layers_[i].data, 32);
  }
}
// merkle/merkle_commitment.h:50-73 with leaf payload = payload[i*len .. ]
// nonces drawn from rng bytes (n x 32 consumed in leaf order).
// Returns proof path length (in digests) for the opening of pos[np].
size_t ref_merkle_commit_open(size_t n, const uint8_t* payload, size_t len,
                              const uint8_t* rng, uint8_t* root_out,
                              const size_t* pos, size_t np,
                              uint8_t* nonce_out /*np*32*/,
                              uint8_t* path_out /*cap np*pathlen*32*/) {
  MerkleCommitment mc(n);
  BufferRandomEngine eng(rng, n * 32);
  auto upd = [&](size_t j, proofs::SHA256& sha) { sha.Update(payload + j * len, len); };
  Digest r = mc.commit(upd, eng);
  memcpy(root_out, r.data, 32);
  if (np == 0) return 0;
  MerkleProof pr(np);
  mc.open(pr, pos, np);
  for (size_t i = 0; i < np; ++i) memcpy(nonce_out + 32 * i, pr.nonce[i].bytes, 32);
  for (size_t i = 0; i < pr.path.size(); ++i) memcpy(path_out + 32 * i, pr.path[i].data, 32);
  return pr.path.size();
}
size_t ref_merkle_tree_len(size_t n) { return merkle_tree_len(n); }

// ---------------- Transcript ----------------
// random/transcript.h:70-190.  A tiny script interpreter so that tests can
// exercise arbitrary write/draw sequences: ops are
//   'B' len32 bytes       write(bytes)
//   'Z' len32             write0(len)
//   'E' 16 bytes          write(GF elt)
//   'A' n32 n*16 bytes    write(array of GF elts)
//   'R' len32             draw len bytes -> appended to out
//   'N' n32               nat(n) -> appended as u32
//   'C' n32 k32           choose(n,k) -> k u32
//   'G' n32               draw n GF elts -> n*16 bytes
size_t ref_transcript_script(const uint8_t* init, size_t init_len,
                             const uint8_t* script, size_t script_len,
                             uint8_t* out, size_t out_cap) {
  const GF& F = gf();
  Transcript ts(init, init_len);
  size_t p = 0, o = 0;
  auto rd32 = [&]() {
    uint32_t v;
    memcpy(&v, script + p, 4);
    p += 4;
    return v;
  };
  while (p < script_len) {
    char op = script[p++];
    if (op == 'B') {
      uint32_t n = rd32();
      ts.write(script + p, n);
      p += n;
    } else if (op == 'Z') {
      ts.write0(rd32());
    } else if (op == 'E') {
      ts.write(F.of_bytes_field(script + p).value(), F);
      p += 16;
    } else if (op == 'A') {
      uint32_t n = rd32();
      std::vector<GF::Elt> v(n);
      for (uint32_t i = 0; i < n; ++i) v[i] = F.of_bytes_field(script + p + 16 * i).value();
      p += 16 * n;
      ts.write(v.data(), 1, n, F);
    } else if (op == 'R') {
      uint32_t n = rd32();
      check(o + n <= out_cap, "out_cap");
      ts.bytes(out + o, n);
      o += n;
    } else if (op == 'N') {
      uint32_t n = rd32();
      uint32_t r = ts.nat(n);
      memcpy(out + o, &r, 4);
      o += 4;
    } else if (op == 'C') {
      uint32_t n = rd32(), k = rd32();
      std::vector<size_t> res(k);
      ts.choose(res.data(), n, k);
      for (uint32_t i = 0; i < k; ++i) {
        uint32_t r = res[i];
        memcpy(out + o, &r, 4);
        o += 4;
      }
    } else if (op == 'G') {
      uint32_t n = rd32();
      for (uint32_t i = 0; i < n; ++i) {
        F.to_bytes_field(out + o, ts.elt(F));
        o += 16;
      }
    } else {
      check(false, "bad script op");
    }
  }
  return o;
}

// ---------------- Ligero parameters ----------------
// ligero/ligero_param.h:116-307 ; out[12]
void ref_ligero_param(int field_id, size_t nw, size_t nq, size_t rate,
                      size_t nreq, size_t block_enc, size_t* out) {
  auto fill = [&](auto& p) {
    out[0] = p.block_enc; out[1] = p.block; out[2] = p.dblock;
    out[3] = p.block_ext; out[4] = p.r; out[5] = p.w; out[6] = p.nwrow;
    out[7] = p.nqtriples; out[8] = p.nwqrow; out[9] = p.nrow;
    out[10] = p.mc_pathlen; out[11] = p.iq;
  };
  if (field_id == GF2_128_ID) {
    if (block_enc) { LigeroParam<GF> p(nw, nq, rate, nreq, block_enc); fill(p); }
    else { LigeroParam<GF> p(nw, nq, rate, nreq); fill(p); }
  } else {
    if (block_enc) { LigeroParam<Fp256Base> p(nw, nq, rate, nreq, block_enc); fill(p); }
    else { LigeroParam<Fp256Base> p(nw, nq, rate, nreq); fill(p); }
  }
}

// ---------------- Circuit info ----------------
// out: nv, logv, nc, logc, nl, ninputs, npub_in, subfield_boundary, nterms,
// then per layer (nw, logw, nterms)   ; returns number of size_t written
size_t ref_circuit_info(int field_id, const uint8_t* circ, size_t circ_len,
                        size_t* out, size_t cap) {
  auto go = [&](const auto& F, FieldID fid) -> size_t {
    using Field = std::decay_t<decltype(F)>;
    CircuitReader<Field> rd(F, fid);
    ReadBuffer rb(circ, circ_len);
    auto c = rd.from_bytes(rb, false);
    if (c == nullptr) return 0;
    size_t k = 0;
    size_t hdr[9] = {c->nv, c->logv, c->nc, c->logc, c->nl, c->ninputs,
                     c->npub_in, c->subfield_boundary, c->nterms()};
    for (size_t v : hdr) if (k < cap) out[k++] = v;
    for (auto& l : c->l) {
      if (k + 3 <= cap) { out[k++] = l.nw; out[k++] = l.logw; out[k++] = l.nterms(); }
    }
    return k;
  };
  if (field_id == GF2_128_ID) return go(gf(), GF2_128_ID);
  return go(p256_base, P256_ID);
}

// ---------------- ZK prover / verifier ----------------
// Dump buffers are optional (may be null); *_cap in bytes.
void* ref_circuit_load(int field_id, const uint8_t* circ, size_t circ_len) {
  auto h = std::make_unique<CircuitHandle>();
  h->field_id = field_id;
  ReadBuffer rb(circ, circ_len);
  if (field_id == GF2_128_ID) {
    CircuitReader<GF> rd(gf(), GF2_128_ID);
    h->gf = rd.from_bytes(rb, true);
    if (!h->gf) return nullptr;
  } else if (field_id == P256_ID) {
    CircuitReader<Fp256Base> rd(p256_base, P256_ID);
    h->p256 = rd.from_bytes(rb, true);
    if (!h->p256) return nullptr;
  } else {
    return nullptr;
  }
  return h.release();
}
void ref_circuit_free(void* h) { delete static_cast<CircuitHandle*>(h); }

int ref_zk_prove(void* handle,
                 const uint8_t* wit, const uint8_t* rng, size_t rng_len,
                 const uint8_t* tinit, size_t tinit_len, size_t rate,
                 size_t nreq, size_t block_enc, uint8_t* out, size_t out_cap,
                 size_t* out_len, size_t* rng_used, uint8_t* d_witness,
                 size_t d_witness_cap, uint8_t* d_tableau, size_t d_tableau_cap,
                 uint8_t* d_root, uint8_t* d_sumcheck, size_t d_sumcheck_cap) {
  Dump dump;
  bool want = d_witness || d_tableau || d_root || d_sumcheck;
  int rc;
  auto* h = static_cast<CircuitHandle*>(handle);
  int field_id = h->field_id;
  if (field_id == GF2_128_ID) {
    LCH14ReedSolomonFactory<GF> rsf(gf());
    rc = zk_prove_t(gf(), rsf, h->gf.get(), wit, rng, rng_len,
                    tinit, tinit_len, rate, nreq, block_enc, out, out_cap,
                    out_len, rng_used, want ? &dump : nullptr);
  } else if (field_id == P256_ID) {
    rc = zk_prove_t(p256_base, p256ctx().rs, h->p256.get(), wit, rng,
                    rng_len, tinit, tinit_len, rate, nreq, block_enc, out,
                    out_cap, out_len, rng_used, want ? &dump : nullptr);
  } else {
    return -100;
  }
  auto cp = [](uint8_t* dst, size_t cap, const std::vector<uint8_t>& src) {
    if (dst && src.size() <= cap) memcpy(dst, src.data(), src.size());
  };
  cp(d_witness, d_witness_cap, dump.witness);
  cp(d_tableau, d_tableau_cap, dump.tableau);
  cp(d_root, 32, dump.root);
  cp(d_sumcheck, d_sumcheck_cap, dump.sumcheck);
  return rc;
}

int ref_zk_verify(void* handle,
                  const uint8_t* pub, const uint8_t* tinit, size_t tinit_len,
                  size_t rate, size_t nreq, size_t block_enc,
                  const uint8_t* proof, size_t proof_len) {
  auto* h = static_cast<CircuitHandle*>(handle);
  int field_id = h->field_id;
  if (field_id == GF2_128_ID) {
    LCH14ReedSolomonFactory<GF> rsf(gf());
    return zk_verify_t(gf(), rsf, h->gf.get(), pub, tinit,
                       tinit_len, rate, nreq, block_enc, proof, proof_len);
  } else if (field_id == P256_ID) {
    return zk_verify_t(p256_base, p256ctx().rs, h->p256.get(), pub,
                       tinit, tinit_len, rate, nreq, block_enc, proof, proof_len);
  }
  return -100;
}

// Throughput/latency of the reference prover on host cores: nthreads
// independent single-threaded provers (the library has no threads of its
// own, docs/content/en/docs/benchmarks.md:7), each producing `per_thread`
// proofs from the same witness / rng bytes.  Returns wall seconds; per-proof
// latencies (ms) of thread 0 are written to lat_ms[per_thread].
double ref_zk_bench(void* handle,
                    const uint8_t* wit, const uint8_t* rng, size_t rng_len,
                    size_t rate, size_t nreq, size_t nthreads,
                    size_t per_thread, double* lat_ms) {
  std::vector<std::thread> th;
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<uint8_t> sink(nthreads);
  auto worker = [&](size_t t) {
    std::vector<uint8_t> out(1 << 20);
    size_t out_len = 0, used = 0;
    ready++;
    while (!go.load()) std::this_thread::yield();
    for (size_t i = 0; i < per_thread; ++i) {
      auto t0 = std::chrono::steady_clock::now();
      int rc = ref_zk_prove(handle, wit, rng, rng_len,
                            (const uint8_t*)"test", 4, rate, nreq, 0,
                            out.data(), out.size(), &out_len, &used, nullptr, 0,
                            nullptr, 0, nullptr, nullptr, 0);
      auto t1 = std::chrono::steady_clock::now();
      if (t == 0 && lat_ms)
        lat_ms[i] = std::chrono::duration<double, std::milli>(t1 - t0).count();
      sink[t] ^= out[out_len / 2] ^ (uint8_t)rc;
    }
  };
  for (size_t t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
  while (ready.load() < (int)nthreads) std::this_thread::yield();
  auto t0 = std::chrono::steady_clock::now();
  go.store(true);
  for (auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Two provers on ONE transcript, the structure of run_mdoc_prover
// (circuits/mdoc/mdoc_zk.cc:459-503): commit(A), commit(B), challenge bytes drawn
// from the transcript (there: the MAC key), prove(A), prove(B).  A is a
// GF(2^128) circuit, B an Fp256 circuit; both draw their coins from one engine.
int ref_zk_prove_pair(void* handle_a, void* handle_b, const uint8_t* wit_a, const uint8_t* wit_b,
                      const uint8_t* rng, size_t rng_len, const uint8_t* tinit, size_t tinit_len, size_t rate,
                      size_t nreq, uint8_t* challenge16, uint8_t* out_a, size_t cap_a, size_t* len_a,
                      uint8_t* out_b, size_t cap_b, size_t* len_b, size_t* rng_used_a, size_t* rng_used_total) {
  set_log_level(ERROR);
  auto* ha = static_cast<CircuitHandle*>(handle_a);
  auto* hb = static_cast<CircuitHandle*>(handle_b);
  if (ha->field_id != GF2_128_ID || hb->field_id != P256_ID) return -100;
  const Circuit<GF>* ca = ha->gf.get();
  const Circuit<Fp256Base>* cb = hb->p256.get();
  Dense<GF> Wa(1, ca->ninputs);
  for (size_t i = 0; i < ca->ninputs; ++i) Wa.v_[i] = gf().of_bytes_field(wit_a + i * GF::kBytes).value();
  Dense<Fp256Base> Wb(1, cb->ninputs);
  for (size_t i = 0; i < cb->ninputs; ++i)
    Wb.v_[i] = p256_base.of_bytes_field(wit_b + i * Fp256Base::kBytes).value();
  BufferRandomEngine eng(rng, rng_len);
  Transcript tp(tinit, tinit_len);
  LCH14ReedSolomonFactory<GF> rsa(gf());
  ZkProof<GF> za(*ca, rate, nreq);
  ZkProof<Fp256Base> zb(*cb, rate, nreq);
  ZkProver<GF, LCH14ReedSolomonFactory<GF>> pa(*ca, gf(), rsa);
  ZkProver<Fp256Base, P256RS> pb(*cb, p256_base, p256ctx().rs);
  pa.commit(za, Wa, tp, eng);
  *rng_used_a = eng.consumed();
  pb.commit(zb, Wb, tp, eng);
  *rng_used_total = eng.consumed();
  tp.bytes(challenge16, 16);
  if (!pa.prove(za, Wa, tp)) return -3;
  if (!pb.prove(zb, Wb, tp)) return -3;
  std::vector<uint8_t> ba, bb;
  za.write(ba, gf());
  zb.write(bb, p256_base);
  *len_a = ba.size();
  *len_b = bb.size();
  if (ba.size() > cap_a || bb.size() > cap_b) return -4;
  memcpy(out_a, ba.data(), ba.size());
  memcpy(out_b, bb.data(), bb.size());
  return 0;
}

#ifdef LF_WITH_GPU_ADAPTERS
// ---- libref_gpu.so only: the reference driving the CUDA back end through
// include/longfellow_b200_adapters.h (tests/test_gpu_adapters.py).  This is the
// reference-side binding of INTEGRATION.md compiled for real.

// (1) the UNMODIFIED reference ZkProver (sumcheck, Ligero, Merkle, transcript on
//     the CPU) with the GPU Reed-Solomon injected at the interpolator-factory seam
int ref_zk_prove_gpu_rs(void* handle, const uint8_t* wit, const uint8_t* rng, size_t rng_len,
                        const uint8_t* tinit, size_t tinit_len, size_t rate, size_t nreq, uint8_t* out,
                        size_t out_cap, size_t* out_len) {
  auto* h = static_cast<CircuitHandle*>(handle);
  lf_ctx* ctx = nullptr;
  if (lf_ctx_create(0, nullptr, &ctx) != LF_OK) return -200;
  int rc;
  size_t used = 0;
  if (h->field_id == GF2_128_ID) {
    longfellow_b200::GpuReedSolomonFactory<GF> rsf(ctx, LF_FIELD_GF2_128, gf());
    rc = zk_prove_t(gf(), rsf, h->gf.get(), wit, rng, rng_len, tinit, tinit_len, rate, nreq, 0, out, out_cap,
                    out_len, &used, nullptr);
  } else if (h->field_id == P256_ID) {
    longfellow_b200::GpuReedSolomonFactory<Fp256Base> rsf(ctx, LF_FIELD_P256, p256_base);
    rc = zk_prove_t(p256_base, rsf, h->p256.get(), wit, rng, rng_len, tinit, tinit_len, rate, nreq, 0, out,
                    out_cap, out_len, &used, nullptr);
  } else {
    rc = -100;
  }
  lf_ctx_destroy(ctx);
  return rc;
}

}  // extern "C"
// (2) the whole prover replaced: GpuZkProver fed from the reference's Circuit,
//     Dense witness and RandomEngine objects
template <class Field>
static int zk_prove_gpu_t(const Field& F, FieldID fid, const Circuit<Field>* c, const uint8_t* wit,
                          const uint8_t* rng, size_t rng_len, const uint8_t* tinit, size_t tinit_len,
                          size_t rate, size_t nreq, size_t copies, uint8_t* out, size_t out_cap,
                          size_t* out_len) {
  lf_ctx* ctx = nullptr;
  if (lf_ctx_create(0, nullptr, &ctx) != LF_OK) return -200;
  int rc = 0;
  {
    longfellow_b200::GpuZkProver<Field> prover(ctx, *c, F, fid, rate, nreq);
    Dense<Field> W(1, c->ninputs);
    for (size_t i = 0; i < c->ninputs; ++i) W.v_[i] = F.of_bytes_field(wit + i * Field::kBytes).value();
    std::vector<const Dense<Field>*> Ws(copies, &W);
    // every copy replays the same coins, so all proofs must be identical
    std::vector<uint8_t> coins;
    for (size_t i = 0; i < copies; ++i) coins.insert(coins.end(), rng, rng + prover.info().rng_bytes);
    (void)rng_len;
    BufferRandomEngine eng(coins.data(), coins.size());
    std::vector<std::vector<uint8_t>> proofs;
    std::vector<bool> ok = prover.prove_batch(Ws, tinit, tinit_len, eng, proofs);
    for (size_t i = 0; i < copies; ++i) {
      if (!ok[i]) rc = -3;
      if (proofs[i] != proofs[0]) rc = -5;
    }
    if (rc == 0) {
      *out_len = proofs[0].size();
      if (proofs[0].size() > out_cap) rc = -4;
      else memcpy(out, proofs[0].data(), proofs[0].size());
    }
  }
  lf_ctx_destroy(ctx);
  return rc;
}
extern "C" {
int ref_zk_prove_gpu(void* handle, const uint8_t* wit, const uint8_t* rng, size_t rng_len, const uint8_t* tinit,
                     size_t tinit_len, size_t rate, size_t nreq, size_t copies, uint8_t* out, size_t out_cap,
                     size_t* out_len) {
  auto* h = static_cast<CircuitHandle*>(handle);
  if (h->field_id == GF2_128_ID)
    return zk_prove_gpu_t(gf(), GF2_128_ID, h->gf.get(), wit, rng, rng_len, tinit, tinit_len, rate, nreq, copies,
                          out, out_cap, out_len);
  if (h->field_id == P256_ID)
    return zk_prove_gpu_t(p256_base, P256_ID, h->p256.get(), wit, rng, rng_len, tinit, tinit_len, rate, nreq,
                          copies, out, out_cap, out_len);
  return -100;
}
#endif  // LF_WITH_GPU_ADAPTERS

}  // extern "C"
