// Batched ZK verifier kernels (one batch = many independent proofs of one circuit).
//
//   reference                                                 here
//   ZkProof::read (zk/zk_proof.h:107-112,218-355)             k_zkv_parse
//   ZkVerifier::recv_commitment + initialize_sumcheck_        k_zk_transcript_init (shared with the prover)
//     fiat_shamir (zk_verifier.h:69-73, zk_common.h:163-180)
//   ZkCommon::verifier_constraints, transcript part           k_zkv_replay        (one thread per proof)
//     (zk_common.h:49-136; transcript_sumcheck.h:49-79)
//   Quad::bind_gh_all (sumcheck/quad.h:188-210)               k_zkv_bind_quad
//   verifier_constraints, the A matrix folded with alphal     k_lig_input_eq + k_lig_avec (shared with the prover)
//   ... its right-hand side b, and <b, alphal>                k_zkv_bvec
//   LigeroVerifier::verify (ligero/ligero_verifier.h:42-268)
//     challenges                                              k_lig_challenges    (shared with the prover)
//     responses into the transcript, idx                      k_zkv_idx           (lig_absorb_and_choose)
//     interpolate_req_columns / layout_Aext + interpolate     k_zkv_rows + the prover's RS row encoder
//     merkle_check, low_degree_check, dot_check,              k_zkv_check
//       "wrong dot product", quadratic_check
//   MerkleTreeVerifier::verify_compressed_proof               inside k_zkv_check
//     (merkle/merkle_tree.h:153-214)
//
// status per proof: LF_OK accepted; LF_ERR_FORMAT the bytes are not a proof (ZkProof::read == false);
// LF_ERR_VERIFY rejected, vwhy = the first check that failed, in the reference's order:
//   1 merkle_check  2 low_degree_check  3 dot_check  4 wrong dot product  5 quadratic_check
#pragma once
#include <stdint.h>

#include "field.cuh"
#include "hash.cuh"
#include "kernels_zk.cuh"
#include "zk_types.cuh"

namespace lf {

// verifier-only per-proof buffers
template <class Elt>
struct ZkVBufs {
  const uint8_t* proofs;  size_t proof_stride;   // serialized proofs (device)
  const uint64_t* lens;
  Elt* req;         // [nrow * nreq] opened columns, row-major as LigeroProof::req_at
  uint8_t* nonce;   // [nreq * 32]
  uint32_t* path;   // [nreq * mc_pathlen * 8] Merkle proof digests, big-endian words
  uint32_t* npath;  // [1]
  Elt* beta;        // [nl]
  Elt* G;           // [40] begin_circuit bindings
  Elt* dots;        // [2] want_dot, proof_dot
  uint8_t* msg;     // message scratch of lig_absorb_and_choose
  size_t msg_stride;
  uint8_t* def;     // [2 * block_ext] "defined" flags of MerkleTreeVerifier::verify_compressed_proof
  int32_t* why;     // [1]
  // small batches (k_zkv_bind_quad_split): eight half EQ tables of half_cap entries each, and the per-CTA
  // partial sums of one layer
  Elt* half;        // [8 * half_cap]
  Elt* part;        // [kVPartMax]
  uint32_t half_cap, pad_;
};
constexpr uint32_t kVPartMax = 1024;

__device__ __forceinline__ uint32_t ld_u32le(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t ld_u32be(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

// ----------------------------------------------------------------------------
// k_zkv_parse: ZkProof::read.  One CTA per proof.  The fixed-size part (root, sumcheck
// proof, responses, nonces) is decoded by all threads; thread 0 walks the run-length
// headers of the opened columns (untrusted sizes: every bound of zk_proof.h:282-327 is
// checked) and records where each element sits; all threads decode them.
// scratch layout as in k_lig_finish: ... | flag bytes[total] | eoff[total]
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_zkv_parse(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const uint8_t* in = v.proofs + p * v.proof_stride;
  const uint64_t len = v.lens[p];
  const uint32_t n = d.block_ext, total = d.nreq * d.nrow;
  uint32_t* sw = b.scratch + p * b.scratch_words;
  uint8_t* mark = reinterpret_cast<uint8_t*>(sw + n);
  uint8_t* flag = mark + 2 * (size_t)((n + 3) & ~3u);
  uint32_t* eoff = reinterpret_cast<uint32_t*>(flag + ((total + 3) & ~3u));
  __shared__ int s_bad;
  __shared__ uint32_t s_path_off, s_npath;
  if (tid == 0) {
    s_bad = 0;
    v.why[p] = 0;
  }
  __syncthreads();
  const uint32_t ny = d.block + d.dblock + d.r + (d.dblock - d.block);
  const uint64_t off_sc = 32, off_y = off_sc + (uint64_t)d.sc_elts * F::kBytes;
  const uint64_t off_nonce = off_y + (uint64_t)ny * F::kBytes, off_req = off_nonce + (uint64_t)d.nreq * 32;
  if (len < off_req) {  // uniform over the CTA
    if (tid == 0) b.status[p] = -3;
    return;
  }
  if (tid < 8) b.nodes[p * (size_t)(2 * d.block_ext * 8) + 8 + tid] = ld_u32be(in + 4 * tid);
  bool ok = true;
  Elt* sc = b.sc + p * d.sc_elts;
  for (uint32_t i = tid; i < d.sc_elts; i += nth) sc[i] = F::from_bytes(in + off_sc + (size_t)i * F::kBytes, &ok);
  {
    // y_ldt | y_dot | y_quad with its witness part cleared (ligero_verifier.h:246-252)
    Elt* y = b.y + p * (size_t)(d.block + 2 * d.dblock);
    const uint32_t n1 = d.block + d.dblock, n2 = n1 + d.r;
    for (uint32_t i = tid; i < ny; i += nth) {
      const Elt e = F::from_bytes(in + off_y + (size_t)i * F::kBytes, &ok);
      y[i < n2 ? i : n1 + d.block + (i - n2)] = e;
    }
    for (uint32_t i = tid; i < d.w; i += nth) y[n1 + d.r + i] = F::zero();
  }
  {
    uint8_t* nz = v.nonce + p * (size_t)d.nreq * 32;
    for (uint32_t i = tid; i < d.nreq * 32; i += nth) nz[i] = in[off_nonce + i];
  }
  if (tid == 0) {
    uint64_t o = off_req;
    uint32_t ci = 0;
    bool sub = false, bad = false;
    while (ci < total) {
      if (o + 4 > len) { bad = true; break; }
      const uint32_t runlen = ld_u32le(in + o);
      o += 4;
      if (runlen >= (1u << 25) || (uint64_t)ci + runlen > total) { bad = true; break; }
      const uint32_t sz = sub ? F::kSubBytes : F::kBytes;
      if (o + (uint64_t)runlen * sz > len) { bad = true; break; }
      for (uint32_t k = ci; k < ci + runlen; ++k) {
        eoff[k] = (uint32_t)o;
        flag[k] = (uint8_t)sub;
        o += sz;
      }
      ci += runlen;
      sub = !sub;
    }
    if (!bad) {
      if (o + 4 > len) {
        bad = true;
      } else {
        const uint32_t sz = ld_u32le(in + o);
        o += 4;
        // zk_proof.h:331-339: a Merkle proof shorter than nreq is not valid; bound the size
        if (sz < d.nreq || sz >= (1u << 25) || o + 32ull * sz > len || sz > d.nreq * d.mc_pathlen) bad = true;
        s_path_off = (uint32_t)o;
        s_npath = sz;
      }
    }
    if (bad) s_bad = 1;
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) b.status[p] = -3;
    return;
  }
  Elt* req = v.req + p * (size_t)total;
  for (uint32_t k = tid; k < total; k += nth) {
    const uint8_t* q = in + eoff[k];
    if (flag[k] && F::kChar2) req[k] = F::of_sub16((uint32_t)q[0] | ((uint32_t)q[1] << 8));
    else req[k] = F::from_bytes(q, &ok);
  }
  {
    uint32_t* path = v.path + p * (size_t)d.nreq * d.mc_pathlen * 8;
    const uint8_t* q = in + s_path_off;
    for (uint32_t i = tid; i < s_npath * 8; i += nth) path[i] = ld_u32be(q + 4 * (size_t)i);
    if (tid == 0) v.npath[p] = s_npath;
  }
  if (!ok) atomicCAS(&b.status[p], 0, -3);  // an element >= p: of_bytes_field fails
}

// public inputs must be canonical field elements (the reference holds them as a Dense<Field>)
template <class F>
__global__ void k_zkv_check_pub(ZkDims d, ZkBufs<typename F::Elt> b) {
  const size_t p = blockIdx.y;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.npub) return;
  bool ok = true;
  (void)F::from_bytes(b.witness_in + p * b.witness_stride + (size_t)i * F::kBytes, &ok);
  if (!ok) atomicCAS(&b.status[p], 0, -3);
}

// ----------------------------------------------------------------------------
// k_zkv_replay: the sumcheck verifier's transcript (zk_common.h:61-118), one thread per
// proof: begin_circuit, per layer alpha / beta, per round the two transmitted evaluations
// in, the challenge out, then the layer's two claims in.  Nothing here depends on the
// bound quads, so all challenges of all layers are known after this one pass.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(32)
k_zkv_replay(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v, const LayerDesc* __restrict__ layers,
             size_t nproofs) {
  typedef typename F::Elt Elt;
  __shared__ AesTables s_aes;
  aes_stage_tables(&s_aes);
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs || b.status[p] != 0) return;
  Transcript* gts = reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript));
  Transcript ts = *gts;
  ts.use_tables(&s_aes);
  ts.have_prf = 0;
  const Elt* sc = b.sc + p * d.sc_elts;
  Elt* hbs = b.hb + p * d.nhb;
  // begin_circuit: Q[40] then G[40] (transcript_sumcheck.h:49-52)
  for (int i = 0; i < 40; ++i) (void)ts_challenge<F>(&ts);
  for (int i = 0; i < 40; ++i) v.G[p * 40 + i] = ts_challenge<F>(&ts);
  for (uint32_t ly = 0; ly < d.nl; ++ly) {
    const LayerDesc L = layers[ly];
    b.alphas[p * d.nl + ly] = ts_challenge<F>(&ts);
    v.beta[p * d.nl + ly] = ts_challenge<F>(&ts);
    for (uint32_t t = 0; t < 2 * L.logw; ++t) {
      const uint32_t hand = t & 1, round = t >> 1;
      ts_write_elt<F>(&ts, sc[L.sc_off + 4 * round + hand]);
      ts_write_elt<F>(&ts, sc[L.sc_off + 4 * round + 2 + hand]);
      hbs[L.hb_off + t] = ts_challenge<F>(&ts);
    }
    ts.begin_array(2);
    ts_array_elt<F>(&ts, sc[L.sc_off + 4 * L.logw]);
    ts_array_elt<F>(&ts, sc[L.sc_off + 4 * L.logw + 1]);
  }
  *gts = ts;
}

// ----------------------------------------------------------------------------
// k_zkv_replay_par: the same transcript without its serial latency.  On the verifier's side every byte
// that enters the transcript is proof data, known before the first challenge is drawn, and a challenge
// never feeds a write.  So the pass splits in three (one CTA per proof):
//   (1) all threads lay the whole byte stream out: the transcript's pending bytes, then per round
//       tag | p(0), tag | p(2), per layer the array header and the two claims;
//   (2) ONE thread runs the SHA-256 chain over the full 64-byte blocks and keeps the chaining value
//       after every block: 0.53 (GF(2^128)) / 1.03 (P-256) compressions per round, nothing else;
//   (3) one thread per KEY POINT -- a stream position after which challenges are drawn: the start
//       (Q, G, alpha_0, beta_0), the end of every round (its challenge), the end of every layer but
//       the last (alpha, beta of the next) -- rebuilds the Transcript as it stood there (chaining value
//       of the blocks before it + the tail bytes), and draws with the ordinary Transcript code: digest
//       snapshot, AES-256 key schedule, blocks, Field::sample with its rejection loop.
// The serial part drops from (absorb + snapshot + key schedule + blocks) per round to the absorb alone.
// rmsg: per proof [rmsg_stride] bytes of stream, then (rmsg_stride / 64 + 1) x 8 words of chaining values.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_zkv_replay_par(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v,
                 const LayerDesc* __restrict__ layers, uint8_t* __restrict__ rmsg, size_t rmsg_stride) {
  typedef typename F::Elt Elt;
  constexpr uint32_t kB = F::kBytes, kEw = 1 + kB;   // bytes of one tagged element write
  __shared__ AesTables s_aes;
  __shared__ uint32_t s_h0[8], s_pos0, s_total, s_nkey;
  __shared__ uint64_t s_len0;
  aes_stage_tables(&s_aes);
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  Transcript* gts = reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript));
  uint8_t* M = rmsg + p * (rmsg_stride + (rmsg_stride / 64 + 1) * 32);
  uint32_t* H = reinterpret_cast<uint32_t*>(M + rmsg_stride);
  const Elt* sc = b.sc + p * d.sc_elts;
  if (tid == 0) {
    const uint32_t pos0 = (uint32_t)(gts->sha.len & 63);
    s_pos0 = pos0;
    s_len0 = gts->sha.len;
    for (int k = 0; k < 8; ++k) s_h0[k] = gts->sha.h[k];
    for (uint32_t k = 0; k < pos0; ++k) M[k] = (uint8_t)(gts->sha.buf[k >> 2] >> (24 - 8 * (k & 3)));
    uint32_t tot = pos0, nkey = 1;
    for (uint32_t ly = 0; ly < d.nl; ++ly) {
      tot += 2 * layers[ly].logw * 2 * kEw + 9 + 2 * kB;
      nkey += 2 * layers[ly].logw + (ly + 1 < d.nl ? 1 : 0);
    }
    s_total = tot;
    s_nkey = nkey;
  }
  __syncthreads();
  const uint32_t pos0 = s_pos0, total = s_total;
  // (1) the byte stream: one thread per element write
  {
    uint32_t base = pos0;
    for (uint32_t ly = 0; ly < d.nl; ++ly) {
      const LayerDesc L = layers[ly];
      const uint32_t nw = 4 * L.logw + 2;   // tagged writes of the rounds, then the two array elements
      for (uint32_t i = tid; i < nw; i += nth) {
        uint32_t off, src;
        bool tagged = true;
        if (i < 4 * L.logw) {
          const uint32_t t = i >> 1, k2 = i & 1, hand = t & 1, round = t >> 1;
          off = base + i * kEw;
          src = L.sc_off + 4 * round + 2 * k2 + hand;
        } else {
          const uint32_t k = i - 4 * L.logw;
          off = base + 4 * L.logw * kEw + 9 + k * kB;
          src = L.sc_off + 4 * L.logw + k;
          tagged = false;
        }
        uint32_t w[F::kWords];
        F::to_wire(w, sc[src]);
        uint8_t* o = M + off;
        if (tagged) *o++ = 1;   // TAG_FIELD_ELEM (transcript.h:136-141)
#pragma unroll
        for (int q = 0; q < F::kWords; ++q) {
          o[4 * q] = (uint8_t)w[q];
          o[4 * q + 1] = (uint8_t)(w[q] >> 8);
          o[4 * q + 2] = (uint8_t)(w[q] >> 16);
          o[4 * q + 3] = (uint8_t)(w[q] >> 24);
        }
      }
      if (tid == 0) {   // array header: TAG_ARRAY, 64-bit little-endian length 2 (transcript.h:144-153,160-171)
        uint8_t* o = M + base + 4 * L.logw * kEw;
        o[0] = 2;
        o[1] = 2;
        for (int k = 2; k < 9; ++k) o[k] = 0;
      }
      base += 4 * L.logw * kEw + 9 + 2 * kB;
    }
  }
  __syncthreads();
  // (2) the chain over the full blocks
  const uint32_t nfull = total / 64;
  if (tid == 0) {
    uint32_t h[8];
    for (int k = 0; k < 8; ++k) h[k] = s_h0[k];
    const uint32_t* m32 = reinterpret_cast<const uint32_t*>(M);
    for (uint32_t j = 0; j < nfull; ++j) {
      uint32_t w[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) w[k] = bswap32(m32[16 * (size_t)j + k]);
      sha256_compress_fn(h, w);
#pragma unroll
      for (int k = 0; k < 8; ++k) H[8 * (size_t)j + k] = h[k];
    }
  }
  __syncthreads();
  // the Transcript as it stands once the first `pos` bytes of the stream are absorbed
  auto at = [&](uint32_t pos, Transcript* ts) {
    const uint32_t nb = pos / 64;
    for (int k = 0; k < 8; ++k) ts->sha.h[k] = nb ? H[8 * (size_t)(nb - 1) + k] : s_h0[k];
    for (int k = 0; k < 16; ++k) ts->sha.buf[k] = 0;
    for (uint32_t k = 64 * nb; k < pos; ++k) ts->sha.buf[(k & 63) >> 2] |= (uint32_t)M[k] << (24 - 8 * (k & 3));
    ts->sha.len = s_len0 - pos0 + pos;
    ts->have_prf = 0;
    ts->nblock = 0;
    ts->rdptr = 16;
    ts->use_tables(&s_aes);
  };
  // (3) the key points
  Elt* hbs = b.hb + p * d.nhb;
  for (uint32_t j = tid; j < s_nkey; j += nth) {
    Transcript ts;
    if (j == 0) {
      // begin_circuit: Q[40] then G[40] (transcript_sumcheck.h:49-52), then alpha, beta of layer 0
      at(pos0, &ts);
      for (int i = 0; i < 40; ++i) (void)ts_challenge<F>(&ts);
      for (int i = 0; i < 40; ++i) v.G[p * 40 + i] = ts_challenge<F>(&ts);
      b.alphas[p * d.nl] = ts_challenge<F>(&ts);
      v.beta[p * d.nl] = ts_challenge<F>(&ts);
      continue;
    }
    uint32_t k = j - 1, base = pos0;
    for (uint32_t ly = 0; ly < d.nl; ++ly) {
      const LayerDesc L = layers[ly];
      const uint32_t nr = 2 * L.logw;
      if (k < nr) {   // the challenge after round k of this layer
        at(base + (k + 1) * 2 * kEw, &ts);
        hbs[L.hb_off + k] = ts_challenge<F>(&ts);
        break;
      }
      k -= nr;
      base += 2 * nr * kEw + 9 + 2 * kB;
      if (k == 0) {   // the end of this layer: alpha, beta of the next
        at(base, &ts);
        b.alphas[p * d.nl + ly + 1] = ts_challenge<F>(&ts);
        v.beta[p * d.nl + ly + 1] = ts_challenge<F>(&ts);
        break;
      }
      k -= 1;
    }
  }
  __syncthreads();
  // the transcript as the pass leaves it (LigeroVerifier::verify continues on it)
  if (tid == 0) {
    Transcript ts = *gts;
    at(total, &ts);
    *gts = ts;
  }
}

// ----------------------------------------------------------------------------
// k_zkv_bind_quad: bq[ly] = Quad::bind_gh_all (quad.h:188-210)
//   = sum over the layer's terms of prep_v(v, EQ2(G0, G1, alpha)[g]) * EQ(H0)[h0] * EQ(H1)[h1],
// G = the bindings of the layer above (begin_circuit's G for layer 0), H = this layer's hand
// challenges.  One CTA per proof and launch per layer; the three tables live in the proof's
// EQ arrays (E0: the G table, E1 and QW: the two H tables).
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_zkv_bind_quad(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v, const uint32_t* __restrict__ arena,
                LayerDesc L, LayerDesc Lprev, uint32_t ly, uint32_t logv, const typename F::Elt* __restrict__ consts) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  Elt* EG = b.eq + p * 3 * (size_t)d.max_eq;
  Elt* EG1 = EG + d.max_eq;   // alpha * EQ(G1), folded into EG below; then EQ(H0)
  Elt* EH1 = EG1 + d.max_eq;
  const Elt* hb = b.hb + p * d.nhb;
  const Elt alpha = b.alphas[p * d.nl + ly], beta = v.beta[p * d.nl + ly];
  auto G = [&](uint32_t hand, uint32_t l) -> Elt {
    return ly == 0 ? v.G[p * 40 + l] : hb[Lprev.hb_off + 2 * l + hand];
  };
  // EQ2(G0, G1, alpha) = EQ(G0) + alpha EQ(G1)  (eqs.h raw_eq2)
  if (tid == 0) {
    EG[0] = F::one();
    EG1[0] = alpha;
  }
  __syncthreads();
  for (uint32_t l = 0; l < logv; ++l) {
    const uint32_t S = 1u << l;
    const Elt g0 = G(0, l), g1 = G(1, l);
    for (uint32_t i = tid; i < 2 * S; i += nth) {
      const uint32_t k = i & (S - 1);
      Elt* E = i < S ? EG : EG1;
      const Elt x = E[k], hi = F::mul(x, i < S ? g0 : g1);
      E[k] = F::sub(x, hi);
      E[k + S] = hi;
    }
    __syncthreads();
  }
  for (uint32_t i = tid; i < (1u << logv); i += nth) EG[i] = F::add(EG[i], EG1[i]);
  __syncthreads();
  // EQ(H0) in EG1's array, EQ(H1) in the QW array
  if (tid == 0) {
    EG1[0] = F::one();
    EH1[0] = F::one();
  }
  __syncthreads();
  for (uint32_t l = 0; l < L.logw; ++l) {
    const uint32_t S = 1u << l;
    const Elt h0 = hb[L.hb_off + 2 * l], h1 = hb[L.hb_off + 2 * l + 1];
    for (uint32_t i = tid; i < 2 * S; i += nth) {
      const uint32_t k = i & (S - 1);
      Elt* E = i < S ? EG1 : EH1;
      const Elt x = E[k], hi = F::mul(x, i < S ? h0 : h1);
      E[k] = F::sub(x, hi);
      E[k + S] = hi;
    }
    __syncthreads();
  }
  Acc acc;
  F::acc_zero(acc);
  // the canonical term list carries g and the constant; its corner (bg_seg) carries the two hands
  const uint32_t *tg = arena + L.bg_g, *tv = arena + L.bg_vi;
  const uint32_t* seg = arena + L.bg_seg;
  const uint32_t *ch0 = arena + L.vq_h0, *ch1 = arena + L.vq_h1;
  for (uint32_t t = tid; t < L.nterms; t += nth) {
    const uint32_t vv = tv[t], c = seg[t];
    const Elt dot = EG[tg[t]];
    const Elt q = (vv & kViOne) ? dot : F::mul((vv & kViZero) ? beta : consts[vv & kViMask], dot);
    F::mac(acc, F::mul(q, EG1[ch0[c]]), EH1[ch1[c]]);
  }
  __shared__ Elt red[8];
  const Elt s = warp_sum<F>(F::reduce(acc));
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    Elt tot = red[0];
    for (uint32_t k = 1; k < nth / 32; ++k) tot = F::add(tot, red[k]);
    b.bq[p * d.nl + ly] = tot;
  }
}

// ----------------------------------------------------------------------------
// The same for small batches.  One CTA per proof leaves the GPU idle, and a layer of millions of terms (the
// mdoc hash circuit: 7.76 M) keeps a single SM busy for ~0.1 s.  An EQ table is a tensor product over its
// variables (eqs.h:46-78), EQ(X)[i] = EQ(X_lo)[i & m] * EQ(X_hi)[i >> s], so the grid-wide form needs only the
// half tables -- at most 2^ceil(log/2) entries, built by one warp each (k_zkv_eq_halves: G0, alpha*G1, H0, H1,
// low and high halves) -- and takes every table value as a product of two: 7 multiplications per term
// instead of 3, spread over the whole GPU (k_zkv_bind_quad_split, grid = chunks x proofs), partial sums added
// by k_zkv_bq_reduce.  Field sums are exact, so the regrouping gives the same element.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_zkv_eq_halves(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v, LayerDesc L, LayerDesc Lprev,
                uint32_t ly, uint32_t logv) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;  // warp w builds table w
  const Elt* hb = b.hb + p * d.nhb;
  const bool isG = w < 4;
  const uint32_t hand = (w >> 1) & 1, hi = w & 1;
  const uint32_t n = isG ? logv : L.logw, s = n / 2;
  const uint32_t first = hi ? s : 0, count = hi ? n - s : s;
  Elt* T = v.half + (p * 8 + w) * (size_t)v.half_cap;
  if (lane == 0) T[0] = (isG && hand == 1 && !hi) ? b.alphas[p * d.nl + ly] : F::one();
  __syncwarp();
  for (uint32_t j = 0; j < count; ++j) {
    const uint32_t l = first + j, S = 1u << j;
    const Elt c = isG ? (ly == 0 ? v.G[p * 40 + l] : hb[Lprev.hb_off + 2 * l + hand]) : hb[L.hb_off + 2 * l + hand];
    for (uint32_t k = lane; k < S; k += 32) {
      const Elt x = T[k], h = F::mul(x, c);
      T[k] = F::sub(x, h);
      T[k + S] = h;
    }
    __syncwarp();
  }
}

template <class F>
__global__ void __launch_bounds__(256)
k_zkv_bind_quad_split(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v,
                      const uint32_t* __restrict__ arena, LayerDesc L, uint32_t ly, uint32_t logv,
                      const typename F::Elt* __restrict__ consts) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const Elt* H = v.half + p * 8 * (size_t)v.half_cap;
  const uint32_t cap = v.half_cap;
  const Elt *G0lo = H, *G0hi = H + cap, *G1lo = H + 2 * cap, *G1hi = H + 3 * cap;
  const Elt *H0lo = H + 4 * cap, *H0hi = H + 5 * cap, *H1lo = H + 6 * cap, *H1hi = H + 7 * cap;
  const uint32_t sG = logv / 2, mG = (1u << sG) - 1, sH = L.logw / 2, mH = (1u << sH) - 1;
  const Elt beta = v.beta[p * d.nl + ly];
  Acc acc;
  F::acc_zero(acc);
  const uint32_t *tg = arena + L.bg_g, *tv = arena + L.bg_vi;
  const uint32_t* seg = arena + L.bg_seg;
  const uint32_t *ch0 = arena + L.vq_h0, *ch1 = arena + L.vq_h1;
  for (uint32_t t = blockIdx.x * nth + tid; t < L.nterms; t += gridDim.x * nth) {
    const uint32_t vv = tv[t], c = seg[t], g = tg[t];
    const uint32_t h0 = ch0[c], h1 = ch1[c];
    const Elt dot = F::add(F::mul(G0lo[g & mG], G0hi[g >> sG]), F::mul(G1lo[g & mG], G1hi[g >> sG]));
    const Elt q = (vv & kViOne) ? dot : F::mul((vv & kViZero) ? beta : consts[vv & kViMask], dot);
    const Elt e0 = F::mul(H0lo[h0 & mH], H0hi[h0 >> sH]), e1 = F::mul(H1lo[h1 & mH], H1hi[h1 >> sH]);
    F::mac(acc, F::mul(q, e0), e1);
  }
  __shared__ Elt red[8];
  const Elt s = warp_sum<F>(F::reduce(acc));
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    Elt tot = red[0];
    for (uint32_t k = 1; k < nth / 32; ++k) tot = F::add(tot, red[k]);
    v.part[p * kVPartMax + blockIdx.x] = tot;
  }
}

template <class F>
__global__ void __launch_bounds__(32)
k_zkv_bq_reduce(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v, uint32_t ly, uint32_t nchunk) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  Elt s = F::zero();
  for (uint32_t k = threadIdx.x; k < nchunk; k += 32) s = F::add(s, v.part[p * kVPartMax + k]);
  s = warp_sum<F>(s);
  if (threadIdx.x == 0) b.bq[p * d.nl + ly] = s;
}

// ----------------------------------------------------------------------------
// k_zkv_bvec: the right-hand sides of verifier_constraints folded with alphal, i.e. the value
// LigeroVerifier::verify compares with the sum of the witness part of y_dot
// (ligero_verifier.h:112-119).  Per layer (zk_common.h:330-399):
//   b[ly] = eqq wc0 wc1 - known,
//   known = P_{-1} (cl0 + alpha cl1) + sum_t ((lag_t[0] - lag_t[1]) p_t(0) + lag_t[2] p_t(2)) P_t
// with P_t the suffix products of lag[1] that k_lig_avec left in the scratch area, (cl0, cl1) the
// transmitted claims of the layer above (zero for layer 0); and the input constraint
// (zk_common.h:406-439): b[nl] = wc0 + alpha_in wc1 - sum_{i < npub} (eq0[i] + alpha_in eq1[i]) pub[i].
// One CTA per proof, after k_lig_input_eq and k_lig_avec.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_zkv_bvec(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v, const LayerDesc* __restrict__ layers) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const Elt* chal = b.chal + p * (size_t)(1 + d.nchal);
  const Elt alpha_in = chal[0];
  const Elt* alphal = chal + 1 + d.nwqrow;
  const Elt* sc = b.sc + p * d.sc_elts;
  const Elt* bq = b.bq + p * d.nl;
  const Elt* alphas = b.alphas + p * d.nl;
  const Elt* lagbuf = reinterpret_cast<const Elt*>(b.scratch + p * b.scratch_words);  // [nhb][4] lag0 lag1 lag2 P
  const Elt* E0 = b.eq + p * 3 * (size_t)d.max_eq;
  const Elt* E1 = E0 + d.max_eq;
  __shared__ Elt red[8];
  Acc acc;
  F::acc_zero(acc);
  for (uint32_t ly = tid; ly < d.nl; ly += nth) {
    const LayerDesc L = layers[ly];
    const Elt wc0 = sc[L.sc_off + 4 * L.logw], wc1 = sc[L.sc_off + 4 * L.logw + 1];
    Elt known = F::zero();
    for (uint32_t t = 0; t < 2 * L.logw; ++t) {
      const Elt* lg = lagbuf + 4 * (size_t)(L.hb_off + t);
      const uint32_t hand = t & 1, round = t >> 1;
      const Elt p0 = sc[L.sc_off + 4 * round + hand], p2 = sc[L.sc_off + 4 * round + 2 + hand];
      const Elt term = F::add(F::mul(F::sub(lg[0], lg[1]), p0), F::mul(lg[2], p2));
      known = F::add(known, F::mul(term, lg[3]));
    }
    if (ly > 0) {
      const LayerDesc Lp = layers[ly - 1];
      const Elt cl0 = sc[Lp.sc_off + 4 * Lp.logw], cl1 = sc[Lp.sc_off + 4 * Lp.logw + 1];
      const Elt* lg = lagbuf + 4 * (size_t)L.hb_off;
      const Elt pm1 = F::mul(lg[3], lg[1]);  // P_{-1} = P_0 lag_0[1]
      known = F::add(known, F::mul(pm1, F::add(cl0, F::mul(alphas[ly], cl1))));
    }
    const Elt rhs = F::sub(F::mul(bq[ly], F::mul(wc0, wc1)), known);
    F::mac(acc, rhs, alphal[ly]);
  }
  // public binding
  Acc pb;
  F::acc_zero(pb);
  for (uint32_t i = tid; i < d.npub; i += nth) {
    bool ok = true;
    const Elt x = F::from_bytes(b.witness_in + p * b.witness_stride + (size_t)i * F::kBytes, &ok);
    F::mac(pb, F::add(E0[i], F::mul(alpha_in, E1[i])), x);
  }
  Elt s = warp_sum<F>(F::reduce(acc)), sp = warp_sum<F>(F::reduce(pb));
  if ((tid & 31) == 0) red[tid >> 5] = F::sub(s, F::mul(sp, alphal[d.nl]));
  __syncthreads();
  if (tid == 0) {
    Elt tot = red[0];
    for (uint32_t k = 1; k < nth / 32; ++k) tot = F::add(tot, red[k]);
    const LayerDesc L = layers[d.nl - 1];
    const Elt got = F::add(sc[L.sc_off + 4 * L.logw], F::mul(alpha_in, sc[L.sc_off + 4 * L.logw + 1]));
    tot = F::add(tot, F::mul(got, alphal[d.nl]));
    v.dots[2 * p] = tot;  // want_dot
  }
  // proof_dot = sum of the witness part of y_dot (ligero_verifier.h:114)
  __syncthreads();
  const Elt* ydot = b.y + p * (size_t)(d.block + 2 * d.dblock) + d.block;
  Elt t = F::zero();
  for (uint32_t j = tid; j < d.w; j += nth) t = F::add(t, ydot[d.r + j]);
  t = warp_sum<F>(t);
  if ((tid & 31) == 0) red[tid >> 5] = t;
  __syncthreads();
  if (tid == 0) {
    Elt tot = red[0];
    for (uint32_t k = 1; k < nth / 32; ++k) tot = F::add(tot, red[k]);
    v.dots[2 * p + 1] = tot;
  }
}

// rows to extend: 0 y_ldt (block), 1 y_dot (dblock), 2 y_quad (dblock), iw + i: [0^r | A_i] (block)
// (ligero_verifier.h:121-131, ligero_param.h:423-430).  One CTA per (row, proof).
template <class F>
__global__ void k_zkv_rows(ZkDims d, ZkBufs<typename F::Elt> b) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t row = blockIdx.x;
  Elt* T = b.tableau + (p * d.nrow + row) * (size_t)d.block_enc;
  const Elt* y = b.y + p * (size_t)(d.block + 2 * d.dblock);
  if (row == 0) {
    for (uint32_t j = threadIdx.x; j < d.block; j += blockDim.x) T[j] = y[j];
  } else if (row < 3) {
    const Elt* src = y + d.block + (row - 1) * d.dblock;
    for (uint32_t j = threadIdx.x; j < d.dblock; j += blockDim.x) T[j] = src[j];
  } else {
    const Elt* A = b.avec + p * (size_t)d.nwqrow * d.w + (size_t)(row - 3) * d.w;
    for (uint32_t j = threadIdx.x; j < d.block; j += blockDim.x) T[j] = j < d.r ? F::zero() : A[j - d.r];
  }
}

// Fault injection for tests, the analogue of BadInterpolator (ligero/ligero_test.cc:114-160): the extension
// part [n, block_enc) of one extended row is rotated left by one, so that exactly one of the verifier's
// interpolations is wrong.  One thread per proof (test-only, tiny).
template <class F>
__global__ void k_zkv_fault(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v, uint32_t row, uint32_t n,
                            int bump_dot, size_t nproofs) {
  typedef typename F::Elt Elt;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  if (bump_dot) {
    v.dots[2 * p] = F::add(v.dots[2 * p], F::one());
    return;
  }
  Elt* T = b.tableau + (p * d.nrow + row) * (size_t)d.block_enc;
  const Elt first = T[n];
  for (uint32_t i = n; i + 1 < d.block_enc; ++i) T[i] = T[i + 1];
  T[d.block_enc - 1] = first;
}

// the responses enter the transcript, the opened columns are drawn (shared with the prover)
template <class F>
__global__ void __launch_bounds__(128, 8)
k_zkv_idx(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v) {
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  uint32_t* sw = b.scratch + p * b.scratch_words;
  uint32_t* perm = sw;
  uint8_t* mark = reinterpret_cast<uint8_t*>(sw + d.block_ext);
  __shared__ AesTables s_aes;
  aes_stage_tables(&s_aes);
  lig_absorb_and_choose<F>(d, b.y + p * (size_t)(d.block + 2 * d.dblock),
                           reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript)), v.msg + p * v.msg_stride, perm,
                           mark, b.idx + p * d.nreq, &s_aes);
}

// ----------------------------------------------------------------------------
// k_zkv_check: merkle_check, low_degree_check, dot_check, the dot value, quadratic_check
// (ligero_verifier.h:92-134,161-268).  One CTA per proof.  T = the extended rows of k_zkv_rows.
// Merkle: leaf j = SHA256(nonce_j || column j) for the nreq opened columns
// (merkle_commitment.h:93-106, ligero_param.h:432-439), then MerkleTreeVerifier::
// verify_compressed_proof (merkle_tree.h:158-207): the nodes the proof must supply are read off
// the marked tree in the reference's order (i = n-1 .. 1), the leaves are set, every inner node
// with both children defined is recomputed level by level, and the root must be defined and equal.
// scratch: perm[n] (unused here) | mark bytes[2n] (tree marks, then "defined") ; nodes = the heap.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_zkv_check(ZkDims d, ZkBufs<typename F::Elt> b, ZkVBufs<typename F::Elt> v) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const uint32_t n = d.block_ext, nreq = d.nreq;
  uint32_t* sw = b.scratch + p * b.scratch_words;
  uint8_t* mark = reinterpret_cast<uint8_t*>(sw + n);      // [2n] tree marks (leaves set by lig_absorb_and_choose)
  uint8_t* def = v.def + p * 2 * (size_t)n;
  uint32_t* nodes = b.nodes + p * (size_t)(2 * n * 8);
  const uint32_t* idx = b.idx + p * nreq;
  const Elt* req = v.req + p * (size_t)nreq * d.nrow;
  const uint8_t* nz = v.nonce + p * (size_t)nreq * 32;
  const uint32_t* path = v.path + p * (size_t)nreq * d.mc_pathlen * 8;
  const uint32_t npath = v.npath[p];
  __shared__ uint32_t s_root[8];
  __shared__ int s_fail[6];
  if (tid < 8) s_root[tid] = nodes[8 + tid];  // the commitment (k_zkv_parse)
  if (tid < 6) s_fail[tid] = 0;
  for (uint32_t i = tid; i < 2 * n; i += nth) def[i] = 0;
  __syncthreads();
  // compressed_merkle_proof_tree (merkle_tree.h:75-98)
  if (n >= 2) {
    int top = 31 - __clz(n - 1);
    for (int lv = top; lv >= 0; --lv) {
      uint32_t lo = 1u << lv, hi = min(2u << lv, n);
      for (uint32_t i = lo + tid; i < hi; i += nth) mark[i] = mark[2 * i] | mark[2 * i + 1];
      __syncthreads();
    }
  }
  if (tid == 0) {
    uint32_t sz = 0;
    bool bad = false;
    for (uint32_t i = n; i-- > 1;) {
      if (mark[i]) {
        uint32_t child = 2 * i;
        if (mark[child]) child = 2 * i + 1;
        if (!mark[child]) {
          if (sz >= npath) {
            bad = true;
            break;
          }
          for (int k = 0; k < 8; ++k) nodes[8 * (size_t)child + k] = path[8 * (size_t)sz + k];
          def[child] = 1;
          ++sz;
        }
      }
    }
    if (bad || sz != npath) s_fail[1] = 1;  // the whole proof must be consumed
  }
  // leaves of the opened columns
  for (uint32_t j = tid; j < nreq; j += nth) {
    uint32_t h[8], w[16];
    sha256_iv(h);
    for (int k = 0; k < 8; ++k) w[k] = ld_u32be(nz + 32 * (size_t)j + 4 * k);
    uint32_t pos = 8;
    const uint64_t totalb = 32 + (uint64_t)d.nrow * F::kBytes;
    for (uint32_t i = 0; i < d.nrow; ++i) {
      uint32_t ww[F::kWords];
      F::to_wire(ww, req[(size_t)i * nreq + j]);
#pragma unroll
      for (int k = 0; k < F::kWords; ++k) {
        w[pos++] = bswap32(ww[k]);
        if (pos == 16) {
          sha256_compress(h, w);
          pos = 0;
        }
      }
    }
    for (uint32_t k = pos; k < 16; ++k) w[k] = 0;
    w[pos] = 0x80000000u;
    if (pos >= 14) {
      sha256_compress(h, w);
      for (int k = 0; k < 16; ++k) w[k] = 0;
    }
    w[14] = (uint32_t)((totalb * 8) >> 32);
    w[15] = (uint32_t)(totalb * 8);
    sha256_compress(h, w);
    const uint32_t l = idx[j] + n;
    for (int k = 0; k < 8; ++k) nodes[8 * (size_t)l + k] = h[k];
    def[l] = 1;
  }
  __syncthreads();
  // recompute every inner node whose children are both defined, deepest level first
  if (n >= 2) {
    int top = 31 - __clz(n - 1);
    for (int lv = top; lv >= 0; --lv) {
      uint32_t lo = 1u << lv, hi = min(2u << lv, n);
      for (uint32_t i = lo + tid; i < hi; i += nth)
        if (def[2 * i] && def[2 * i + 1]) {
          merkle_hash2(nodes, i);
          def[i] = 1;
        }
      __syncthreads();
    }
  }
  if (tid == 0) {
    bool okr = def[1] != 0;
    for (int k = 0; k < 8; ++k) okr = okr && nodes[8 + k] == s_root[k];
    if (!okr) s_fail[1] = 1;
  }
  // the three column checks at the opened positions
  const Elt* T = b.tableau + p * (size_t)d.nrow * d.block_enc;
  const Elt* chal = b.chal + p * (size_t)(1 + d.nchal);
  const Elt* u_ldt = chal + 1;
  const Elt* u_quad = chal + 1 + d.nwqrow + (d.nl + 1) + 3 * d.nq;
  const size_t ld = d.block_enc;
  for (uint32_t j = tid; j < nreq; j += nth) {
    const size_t col = (size_t)d.dblock + idx[j];
    Acc acc;
    F::acc_zero(acc);
    for (uint32_t i = 0; i < d.nwqrow; ++i) F::mac(acc, u_ldt[i], req[(size_t)(i + d.iw) * nreq + j]);
    if (!F::eq(F::add(req[j], F::reduce(acc)), T[col])) s_fail[2] = 1;
    F::acc_zero(acc);
    for (uint32_t i = 0; i < d.nwqrow; ++i) F::mac(acc, T[(size_t)(i + d.iw) * ld + col], req[(size_t)(i + d.iw) * nreq + j]);
    if (!F::eq(F::add(req[(size_t)nreq + j], F::reduce(acc)), T[ld + col])) s_fail[3] = 1;
    F::acc_zero(acc);
    const uint32_t iqx = d.iq, iqy = iqx + d.nqtriples, iqz = iqy + d.nqtriples;
    for (uint32_t i = 0; i < d.nqtriples; ++i) {
      const Elt tmp = F::sub(req[(size_t)(iqz + i) * nreq + j],
                             F::mul(req[(size_t)(iqx + i) * nreq + j], req[(size_t)(iqy + i) * nreq + j]));
      F::mac(acc, u_quad[i], tmp);
    }
    if (!F::eq(F::add(req[2 * (size_t)nreq + j], F::reduce(acc)), T[2 * ld + col])) s_fail[5] = 1;
  }
  __syncthreads();
  if (tid == 0) {
    if (!F::eq(v.dots[2 * p], v.dots[2 * p + 1])) s_fail[4] = 1;
    int why = 0;
    for (int k = 5; k >= 1; --k)
      if (s_fail[k]) why = k;
    v.why[p] = why;
    if (why) b.status[p] = -8;
  }
}

}  // namespace lf
