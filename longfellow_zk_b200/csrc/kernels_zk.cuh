// Batched ZK prover kernels (one batch = many independent proofs of one
// circuit; blockIdx.y or blockIdx.x selects the proof).
//
//   reference                                           here
//   ZkProver::fill_pad (zk/zk_prover.h:152-188)          k_zk_witness
//   LigeroProver::layout_* (ligero_prover.h:171-279)     k_zk_layout
//   initialize_sumcheck_fiat_shamir (zk_common.h:163)    k_zk_transcript_init
//   ProverLayers::eval_circuit (prover_layers.h:52-98)   k_zk_eval_layer
//   ProverLayers::prove/layer (prover_layers.h:114-271)  k_zk_sumcheck
//   ZkCommon::verifier_constraints + LigeroProver::prove k_lig_*
//   ZkProof::write (zk/zk_proof.h:90-184)                k_zk_serialize
#pragma once
#include <cooperative_groups.h>
#include <stdint.h>

#include "field.cuh"
#include "hash.cuh"
#include "zk_types.cuh"

namespace lf {

// transcript writes of an element: tag + wire bytes (transcript.h:136-153)
// (elements are taken by value and their words absorbed with a fully unrolled
// loop, so the wire words stay in registers)
// One out-of-line copy each: they are called from a dozen sites of the serial
// path and every inlined copy carries the word-by-word buffer handling.
template <class F>
__device__ __noinline__ void ts_array_elt(Transcript* ts, typename F::Elt e) {
  uint32_t w[F::kWords];
  F::to_wire(w, e);
  ts->sha.template put_le_words<F::kWords>(w);
}
// tag + element in one pass over the buffer (Sha256::put_stream_words)
template <class F>
__device__ __noinline__ void ts_write_elt(Transcript* ts, typename F::Elt e) {
  uint32_t w[F::kWords];
  F::to_wire(w, e);
  ts->have_prf = 0;  // a write drops the challenge stream (transcript.h:174-178)
  ts->sha.template put_tagged_le_words<F::kWords>(1, w);  // TAG_FIELD_ELEM
}
// Field::sample on the transcript (random.h:37-41), out of line for the same reason
template <class F>
__device__ __noinline__ typename F::Elt ts_challenge(Transcript* ts) {
  return F::ts_elt(ts);
}

// ----------------------------------------------------------------------------
// Caller randomness of a prime field.  Field::sample (algebra/fp_generic.h:360-371)
// draws kBytes at a time from the RandomEngine and draws AGAIN while the value is
// >= p, so the position of sample k in the caller's stream depends on how many
// earlier draws were rejected.  Every caller-random element of the prover comes
// before the Merkle nonces (pad, ILDT, IDOT, IQUAD, witness
// rows, quadratic rows), so the stream up to there is a sequence of kBytes slots
// and the samples are the slots that pass, in order.  k_zk_rng_scan finds the
// rejected slots of each proof once (almost always none: 2^-32 per slot for
// P-256); every reader maps sample index -> slot through that short sorted list.
// rej[0] = number of rejected slots, rej[1 + j] = index of the sample whose draw
// the j-th rejected slot was (= accepted slots before it), non-decreasing in j.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rng_slot(const uint32_t* __restrict__ rej, uint32_t k) {
  if (rej == nullptr) return k;
  const uint32_t n = rej[0];
  if (n == 0) return k;
  uint32_t lo = 0, hi = n;  // number of rejected slots that precede the slot of sample k
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (rej[1 + mid] <= k) lo = mid + 1;
    else hi = mid;
  }
  return k + lo;
}
// pointer to sample k of proof p (prime fields: through the reject list)
template <class F>
__device__ __forceinline__ const uint8_t* rng_sample_ptr(const uint8_t* rng, const uint32_t* rej, uint32_t k) {
  if constexpr (F::kChar2) return rng + (size_t)k * F::kBytes;
  else return rng + (size_t)rng_slot(rej, k) * F::kBytes;
}

// One CTA per proof (prime fields only).  Slots are scanned in stream order, never
// further than the samples still missing, so a slot is tested only if the reference
// would have drawn it as a sample.  The stream must hold the nonces behind the last
// sample: otherwise the proof is LF_ERR_RNG (stream too short).
template <class F>
__global__ void __launch_bounds__(256)
k_zk_rng_scan(ZkDims d, ZkBufs<typename F::Elt> b) {
  const size_t p = blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* rej = b.rej + p * (size_t)(1 + d.rej_cap);
  const uint8_t* rng = b.rng + p * b.rng_stride;
  __shared__ uint32_t s_wcnt[8];
  const size_t max_slots = (b.rng_avail - (size_t)d.block_ext * 32) / F::kBytes;
  uint32_t acc = 0, pos = 0, nrej = 0;
  bool too_short = false;
  while (acc < d.rng_nsamples) {
    const uint32_t nc = min(256u, d.rng_nsamples - acc);
    if ((size_t)pos + nc > max_slots) {
      too_short = true;
      break;
    }
    const bool r = tid < nc && !F::sample_ok(rng + (size_t)(pos + tid) * F::kBytes);
    const unsigned bal = __ballot_sync(0xffffffffu, r);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; ++w) {
      const uint32_t c = s_wcnt[w];
      if (w < warp) before += c;
      total += c;
    }
    if (r) {
      const uint32_t j = nrej + before + __popc(bal & ((1u << lane) - 1u));
      if (j < d.rej_cap) rej[1 + j] = (pos + tid) - j;
    }
    __syncthreads();
    nrej += total;
    acc += nc - total;
    pos += nc;
  }
  if (tid == 0) {
    const bool bad = too_short || nrej > d.rej_cap;
    rej[0] = bad ? 0u : nrej;  // a failed proof reads the un-shifted slots: always inside the buffer
    if (bad) atomicCAS(&b.status[p], 0, -6);
  }
}

// ----------------------------------------------------------------------------
// k_zk_witness: Ligero witness = private inputs || pad (zk_prover.h:78-96,152-188)
// pad of layer i: 4*logw+2 random elements in order (round, hand, k in {0,2}),
// then wc[0], wc[1]; then the product wc[0]*wc[1].
// ----------------------------------------------------------------------------
template <class F>
__global__ void k_zk_witness(ZkDims d, ZkBufs<typename F::Elt> b, const LayerDesc* __restrict__ layers) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nw) return;
  Elt* wit = b.wit + p * d.nw;
  const uint8_t* rng = b.rng + p * b.rng_stride;
  const uint32_t* rej = b.rej ? b.rej + p * (size_t)(1 + d.rej_cap) : nullptr;
  bool ok = true, rok = true;
  if (i < d.n_witness) {
    const uint8_t* w = b.witness_in + p * b.witness_stride + (size_t)(i + d.npub) * F::kBytes;
    wit[i] = F::from_bytes(w, &ok);
    if (!ok) atomicCAS(&b.status[p], 0, -3);  // (the first error of a proof sticks) non-canonical input element
    return;
  }
  uint32_t q = i - d.n_witness;  // index inside the pad block
  if (q >= d.pad_size) {
    wit[i] = F::zero();
    return;
  }
  // locate the layer (nl is small)
  uint32_t ly = 0;
  while (ly + 1 < d.nl && layers[ly + 1].pad_off <= q) ++ly;
  uint32_t j = q - layers[ly].pad_off, cnt = 4 * layers[ly].logw + 2;
  if (j < cnt) {
    wit[i] = F::sample_bytes(rng_sample_ptr<F>(rng, rej, layers[ly].sc_off + j), &rok);
  } else {
    Elt a = F::sample_bytes(rng_sample_ptr<F>(rng, rej, layers[ly].sc_off + cnt - 2), &rok);
    Elt c = F::sample_bytes(rng_sample_ptr<F>(rng, rej, layers[ly].sc_off + cnt - 1), &rok);
    wit[i] = F::mul(a, c);
  }
  // (the first error of a proof sticks) never expected once k_zk_rng_scan accepted the stream
  if (!rok) atomicCAS(&b.status[p], 0, -6);
}

// ----------------------------------------------------------------------------
// k_zk_layout: message part of every tableau row (ligero_prover.h:171-279).
// row_rng[i] = byte offset of row i's randomness; row_sub[i] != 0 marks rows
// whose blinding is sampled from the subfield.  One CTA per (row, proof).
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_zk_layout(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ row_rng,
            const uint8_t* __restrict__ row_sub, const uint32_t* __restrict__ lqc /* [nq][3] */) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  const uint32_t row = blockIdx.x;
  Elt* T = b.tableau + (p * d.nrow + row) * (size_t)d.block_enc;
  const Elt* wit = b.wit + p * d.nw;
  const uint8_t* rng0 = b.rng + p * b.rng_stride;
  const uint8_t* rng = rng0 + row_rng[row];
  const uint32_t* rej = b.rej ? b.rej + p * (size_t)(1 + d.rej_cap) : nullptr;
  const uint32_t k0 = row_rng[row] / F::kBytes;  // prime fields: index of the row's first sample
  // j-th full-size sample of this row
  auto samp = [&](uint32_t j, bool* okp) {
    if constexpr (F::kChar2) return F::sample_bytes(rng + (size_t)j * F::kBytes, okp);
    else return F::sample_bytes(rng_sample_ptr<F>(rng0, rej, k0 + j), okp);
  };
  __shared__ Elt red[8];
  bool rok = true;
  if (row == 0) {  // ILDT: block random elements
    for (uint32_t j = threadIdx.x; j < d.block; j += blockDim.x) T[j] = samp(j, &rok);
  } else if (row == 1) {  // IDOT: dblock random, then T[r] -= sum of the W part
    Elt s = F::zero();
    for (uint32_t j = threadIdx.x; j < d.dblock; j += blockDim.x) {
      Elt e = samp(j, &rok);
      T[j] = e;
      if (j >= d.r && j < d.r + d.w) s = F::add(s, e);
    }
    // block reduction of s
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Elt t;
#pragma unroll
      for (int k = 0; k < F::kWords; ++k) t.w[k] = __shfl_xor_sync(0xffffffffu, s.w[k], o);
      s = F::add(s, t);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      Elt tot = red[0];
      for (uint32_t k = 1; k < blockDim.x / 32; ++k) tot = F::add(tot, red[k]);
      T[d.r] = F::sub(samp(d.r, &rok), tot);
    }
  } else if (row == 2) {  // IQUAD: dblock random with the W part cleared
    for (uint32_t j = threadIdx.x; j < d.dblock; j += blockDim.x) {
      Elt e = samp(j, &rok);
      T[j] = (j >= d.r && j < d.r + d.w) ? F::zero() : e;
    }
  } else if (row < d.iq) {  // witness rows
    uint32_t i = row - d.iw;
    bool sub = row_sub[row] != 0;
    for (uint32_t j = threadIdx.x; j < d.block; j += blockDim.x) {
      Elt e;
      if (j < d.r) {
        if (sub) {
          const uint8_t* q = rng + 2 * (size_t)j;
          e = F::of_sub16((uint32_t)q[0] | ((uint32_t)q[1] << 8));
        } else {
          e = samp(j, &rok);
        }
      } else {
        uint32_t k = i * d.w + (j - d.r);
        e = (k < d.nw) ? wit[k] : F::zero();
      }
      T[j] = e;
    }
  } else {  // quadratic rows: x_i, y_i, z_i
    uint32_t t = row - d.iq;
    uint32_t which = t / d.nqtriples, i = t % d.nqtriples;
    for (uint32_t j = threadIdx.x; j < d.block; j += blockDim.x) {
      Elt e;
      if (j < d.r) {
        e = samp(j, &rok);
      } else {
        uint32_t k = i * d.w + (j - d.r);
        e = (k < d.nq) ? wit[lqc[3 * k + which]] : F::zero();
      }
      T[j] = e;
    }
  }
  if (!rok) atomicCAS(&b.status[p], 0, -6);
}

// ----------------------------------------------------------------------------
// k_zk_transcript_init: Transcript(tinit) ; write(root) ; then
// initialize_sumcheck_fiat_shamir (zk_common.h:163-180).  One thread per proof.
// ----------------------------------------------------------------------------
// ext == nullptr: the proof owns its transcript, Transcript(tinit) and the root are
// written here.  ext != nullptr: the caller's transcript (already holding this and
// possibly other commitments, lf_zk_commit_batch) continues.
template <class F>
__global__ void k_zk_transcript_init(ZkDims d, ZkBufs<typename F::Elt> b, const uint8_t* __restrict__ tinit,
                                     const uint8_t* __restrict__ circuit_id, size_t nproofs,
                                     const TranscriptState* __restrict__ ext) {
  __shared__ uint8_t s_sbox[256];
  aes_stage_sbox(s_sbox);
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  Transcript ts;
  if (ext) {
    ts.sbox = s_sbox;
    ts.tab = nullptr;
    ts.import_state(ext + p);
  } else {
    ts.init(tinit, d.tinit_len);
    ts.sbox = s_sbox;
    // ligero_transcript.h:31-34 write_commitment: root digest is node 1
    const uint32_t* root = b.nodes + p * (size_t)(2 * d.block_ext * 8) + 8;
    uint32_t rw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rw[k] = bswap32(root[k]);
    ts.write_bytes_words(rw, 8);
  }
  ts.write_bytes(circuit_id, 32);
  for (uint32_t i = 0; i < d.npub; ++i) {
    // the wire bytes of the public inputs go into the transcript unchanged
    const uint8_t* q = b.witness_in + p * b.witness_stride + (size_t)i * F::kBytes;
    uint32_t w[F::kWords];
    for (int k = 0; k < F::kWords; ++k)
      w[k] = (uint32_t)q[4 * k] | ((uint32_t)q[4 * k + 1] << 8) | ((uint32_t)q[4 * k + 2] << 16) |
             ((uint32_t)q[4 * k + 3] << 24);
    ts.write_elt_words(w, F::kWords);
  }
  {
    uint32_t w[F::kWords];
    for (int k = 0; k < F::kWords; ++k) w[k] = 0;
    ts.write_elt_words(w, F::kWords);  // F.zero()
  }
  ts.write0(d.nterms);
  *reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript)) = ts;
}

// ZkProver::commit's transcript step on a caller-owned transcript: write the root
// (ligero_transcript.h:31-34) and hand the state back; also returns the roots.
template <class F>
__global__ void k_zk_transcript_commit(ZkDims d, ZkBufs<typename F::Elt> b, TranscriptState* __restrict__ ext,
                                       uint8_t* __restrict__ roots_out, size_t nproofs) {
  __shared__ uint8_t s_sbox[256];
  aes_stage_sbox(s_sbox);
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  Transcript ts;
  ts.sbox = s_sbox;
  ts.tab = nullptr;
  ts.import_state(ext + p);
  const uint32_t* root = b.nodes + p * (size_t)(2 * d.block_ext * 8) + 8;
  uint32_t rw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    rw[k] = bswap32(root[k]);
    reinterpret_cast<uint32_t*>(roots_out)[p * 8 + k] = rw[k];
  }
  ts.write_bytes_words(rw, 8);
  ts.export_state(ext + p);
}
// the transcript as the prover left it (after LigeroProver::prove), for the caller
template <class F>
__global__ void k_zk_transcript_export(ZkBufs<typename F::Elt> b, TranscriptState* __restrict__ ext,
                                       size_t nproofs) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  reinterpret_cast<const Transcript*>(b.ts + p * sizeof(Transcript))->export_state(ext + p);
}

template <class F>
__device__ __forceinline__ typename F::Elt warp_sum(typename F::Elt s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    typename F::Elt t;
#pragma unroll
    for (int k = 0; k < F::kWords; ++k) t.w[k] = __shfl_xor_sync(0xffffffffu, s.w[k], o);
    s = F::add(s, t);
  }
  return s;
}

// ----------------------------------------------------------------------------
// k_zk_eval_layer: V[g] = sum over the quad terms of gate g of v * W[l] * W[r]
// (prover_layers.h:278-305), gathered through the CSR-by-gate plan.  Assert-zero
// terms (v == 0) are checked; a violation marks the proof LF_ERR_WITNESS.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_zk_eval_layer(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena, LayerDesc L,
                const typename F::Elt* __restrict__ consts, int is_output) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  const Elt* W = b.wl + p * d.wl_elts + L.w_off;
  const uint32_t* off = arena + L.ev_off;
  const uint32_t *h0 = arena + L.ev_h0, *h1 = arena + L.ev_h1, *vi = arena + L.ev_vi;
  auto term = [&](typename F::Acc& acc, bool& bad, uint32_t t) {
    uint32_t v = vi[t];
    Elt x = F::mul(W[h1[t]], W[h0[t]]);
    if (v & kViZero) {
      bad |= !F::is_zero(x);
    } else if (v & kViOne) {
      F::acc_add_elt(acc, x);
    } else {
      F::mac(acc, consts[v & kViMask], x);
    }
  };
  auto finish = [&](uint32_t g, const Elt& out, bool bad) {
    if (is_output) {
      bad |= !F::is_zero(out);  // zk_prover.h:117-122: all outputs must be zero
    } else {
      b.wl[p * d.wl_elts + L.out_off + g] = out;
    }
    if (bad) atomicCAS(&b.status[p], 0, -5);
  };
  // Gates are ordered by decreasing term count, so a warp's heaviest gate is its first lane's.  A few gates
  // sum hundreds of terms (the ECDSA circuit has one with 769): a warp whose first gate is that heavy takes its
  // gates one after the other with all 32 lanes on the terms, instead of leaving 31 lanes waiting for one.
  const uint32_t gi0 = gi & ~31u;
  uint32_t heavy = 0;
  if (gi0 < L.nout) {
    const uint32_t g0 = arena[L.ev_perm + gi0];
    heavy = off[g0 + 1] - off[g0];
  }
  if (heavy > 64) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t k = 0; k < 32 && gi0 + k < L.nout; ++k) {
      const uint32_t g = arena[L.ev_perm + gi0 + k];
      typename F::Acc acc;
      F::acc_zero(acc);
      bool bad = false;
      for (uint32_t t = off[g] + lane; t < off[g + 1]; t += 32) term(acc, bad, t);
      const Elt out = warp_sum<F>(F::reduce(acc));
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) finish(g, out, bad);
    }
    return;
  }
  if (gi >= L.nout) return;
  const uint32_t g = arena[L.ev_perm + gi];
  typename F::Acc acc;
  F::acc_zero(acc);
  bool bad = false;
  for (uint32_t t = off[g]; t < off[g + 1]; ++t) term(acc, bad, t);
  finish(g, F::reduce(acc), bad);
}

// ----------------------------------------------------------------------------
// k_zk_sumcheck: the whole layered sumcheck of one proof in one persistent CTA,
// Fiat-Shamir transcript included (thread 0), so that the 2*sum(logw)
// sequential rounds never leave the SM.
//
// All sparse sums are BALANCED SEGMENTED SUMS: the circuit's constant wire 0
// appears in tens of thousands of quad terms, so one QW row (and one bind_g
// corner) can own most of the work.  Every thread therefore takes an equal
// contiguous slice of the CSR-ordered entries, sums runs of equal segment id,
// stores complete segments directly and parks the partial of a segment that
// started in an earlier slice in shared memory; the slice that owns the start
// of the segment adds those partials after one barrier.  Sums are exact field
// sums, so the regrouping cannot change the result.
// ----------------------------------------------------------------------------
constexpr int kScMaxThreads = 1024;
constexpr uint32_t kScSoloWork = 1536;  // cluster mode: steps with n_in + n0 below this run on the leader CTA alone

template <class F>
struct ScShared {
  Transcript ts;
  typename F::Elt G[2][40];   // bindings of the previous layer (Proof::kMaxBindings)
  typename F::Elt red[2][kScMaxThreads / 32];  // per-warp partials of a0, a2
  typename F::Elt r, alpha, beta, sum, wc[2];
  typename F::Elt pc[3];  // monomial coefficients of the current round polynomial
  typename F::Elt* hp;  // [blockDim.x] head partials of the segmented sums
  uint32_t* hr;         // [blockDim.x] their segment ids
  // cluster mode (one proof on several SMs): per-CTA results that the other CTAs
  // of the cluster read through distributed shared memory
  typename F::Elt lead_val;      // sum of the leading run of head partials of this CTA ...
  uint32_t lead_seg;             // ... and its segment id (kNone if the CTA starts a segment)
  typename F::Elt cred[2][16];   // leader only: a0 / a2 partial of every CTA
  AesTables aes;        // AES S-box and round tables staged in shared memory
  int fail;
  long long prof[16];   // LF_PROF: cycles of thread 0 per phase ([7] layer set-up, [8] QW, [9] evaluations, [10] bind, [11] solo rounds)
};

// What the sumcheck prover carries from one kernel to the next when a proof's rounds are split
// over several launches (the flat path, kernels_scflat.cuh): the transcript hash (no live challenge
// stream ever crosses a launch: every launch ends on a write or starts on one), the bindings, and
// the running claim.
template <class F>
struct ScCore {
  Sha256 sha;
  typename F::Elt G[2][40];
  typename F::Elt r, alpha, beta, sum, wc[2];
  int fail;
};
// which part of the layered sumcheck one launch of the per-proof kernel runs
struct ScRange {
  uint32_t ly_begin, ly_end;  // layers [ly_begin, ly_end)
  uint32_t t_begin;           // first round of layer ly_begin (0: the layer starts here, with its EQ tables and bind_g)
  uint32_t first;             // != 0: this launch starts the proof (sc_begin on the transcript k_zk_transcript_init left)
};

template <class F>
__device__ __noinline__ void sc_save(const ScShared<F>* sh, ScCore<F>* g) {
  g->sha = sh->ts.sha;
  for (int i = 0; i < 40; ++i) {
    g->G[0][i] = sh->G[0][i];
    g->G[1][i] = sh->G[1][i];
  }
  g->r = sh->r;
  g->alpha = sh->alpha;
  g->beta = sh->beta;
  g->sum = sh->sum;
  g->wc[0] = sh->wc[0];
  g->wc[1] = sh->wc[1];
  g->fail = sh->fail;
}
template <class F>
__device__ __noinline__ void sc_load(ScShared<F>* sh, const ScCore<F>* g) {
  sh->ts.sha = g->sha;
  sh->ts.have_prf = 0;
  sh->ts.nblock = 0;
  sh->ts.rdptr = 16;
  sh->ts.use_tables(&sh->aes);
  for (int i = 0; i < 40; ++i) {
    sh->G[0][i] = g->G[0][i];
    sh->G[1][i] = g->G[1][i];
  }
  sh->r = g->r;
  sh->alpha = g->alpha;
  sh->beta = g->beta;
  sh->sum = g->sum;
  sh->wc[0] = g->wc[0];
  sh->wc[1] = g->wc[1];
  sh->fail = g->fail;
}

// Work distribution of one proof: CL == false, one CTA (tid/nth); CL == true, a
// thread-block cluster of C CTAs on C SMs, gtid/gnth run over all of them,
// barriers are cluster barriers and the leader CTA (rank 0) owns the transcript.
template <bool CL>
struct ScPar {
  uint32_t tid, nth, rank, ncta, gtid, gnth;
  bool cl;  // barriers and work distribution span the cluster (else: this CTA alone)
  __device__ __forceinline__ ScPar() {
    tid = threadIdx.x;
    nth = blockDim.x;
    if (CL) {
      cooperative_groups::cluster_group cg = cooperative_groups::this_cluster();
      rank = cg.block_rank();
      ncta = cg.num_blocks();
    } else {
      rank = 0;
      ncta = 1;
    }
    cl = CL;
    gtid = rank * nth + tid;
    gnth = ncta * nth;
  }
  // the same CTA working alone (leader CTA of a cluster during the small steps)
  __device__ __forceinline__ ScPar solo() const {
    ScPar s = *this;
    s.rank = 0;
    s.ncta = 1;
    s.gtid = tid;
    s.gnth = nth;
    s.cl = false;
    return s;
  }
  __device__ __forceinline__ void sync() const {
    if (CL && cl) cooperative_groups::this_cluster().sync();  // release/acquire: orders global memory too
    else __syncthreads();
  }
  template <class T>
  __device__ __forceinline__ T* remote(T* p, uint32_t r) const {
    if (CL && cl) return cooperative_groups::this_cluster().map_shared_rank(p, r);
    return p;
  }
};

// out[seg] = sum over the entries e of that segment of term(e); seg[] is
// non-decreasing over e in [0, n).  Segments without entries are not touched.
// Contains barriers; all threads of the proof must call it.
// fetch(e) loads the operands of entry e; accum(acc, ops) adds the entry's term.
// (Issuing the loads one entry ahead into registers, and register-free
// prefetch.global.L1 of the next entry's gathers, were both measured slower.)
template <class F, bool CL, class Fetch, class Accum>
__device__ __forceinline__ void seg_sum(ScShared<F>* sh, const ScPar<CL>& P, uint32_t n,
                                        const uint32_t* __restrict__ seg, typename F::Elt* __restrict__ out,
                                        Fetch fetch, Accum accum) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const uint32_t tid = P.tid, nth = P.nth;
  const uint32_t per = (n + P.gnth - 1) / P.gnth;
  const uint32_t e0 = min(n, P.gtid * per), e1 = min(n, e0 + per);
  const uint32_t kNone = 0xffffffffu;
  sh->hr[tid] = kNone;
  uint32_t own_last = kNone;  // segment whose start this slice owns and which is still open at e1
  if (e0 < e1) {
    uint32_t cur = seg[e0];
    bool head = e0 > 0 && seg[e0 - 1] == cur;
    Acc acc;
    F::acc_zero(acc);
    for (uint32_t e = e0; e < e1; ++e) {
      const uint32_t sg = seg[e];
      const auto cur_ops = fetch(e);
      if (sg != cur) {
        Elt v = F::reduce(acc);
        if (head) {
          sh->hp[tid] = v;
          sh->hr[tid] = cur;
        } else {
          out[cur] = v;
        }
        head = false;
        cur = sg;
        F::acc_zero(acc);
      }
      accum(acc, cur_ops);
    }
    Elt v = F::reduce(acc);
    if (head) {
      sh->hp[tid] = v;
      sh->hr[tid] = cur;
    } else {
      out[cur] = v;
      if (e1 < n && seg[e1] == cur) own_last = cur;
    }
  }
  __syncthreads();
  // Suffix sums of the head partials inside each warp (runs of equal segment id
  // are contiguous in thread order), so that the owner of a long segment adds
  // one value per warp instead of one per thread.
  {
    Elt v = sh->hp[tid];
    const uint32_t sg = sh->hr[tid], lane = tid & 31;
    if (sg == kNone) v = F::zero();
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      Elt v2;
#pragma unroll
      for (int k = 0; k < F::kWords; ++k) v2.w[k] = __shfl_down_sync(0xffffffffu, v.w[k], dlt);
      uint32_t sg2 = __shfl_down_sync(0xffffffffu, sg, dlt);
      if (lane + dlt < 32 && sg != kNone && sg2 == sg) v = F::add(v, v2);
    }
    sh->hp[tid] = v;  // each thread rewrites only its own slot
  }
  __syncthreads();
  if (CL && P.cl) {
    // a segment that began in an earlier CTA: this CTA's share of it is the run of
    // head partials starting at thread 0; its owner (in an earlier CTA) collects it
    if (tid == 0) {
      const uint32_t ls = sh->hr[0];
      sh->lead_seg = ls;
      if (ls != kNone) {
        Elt v = sh->hp[0];
        for (uint32_t t = 32; t < nth && sh->hr[t] == ls; t += 32) v = F::add(v, sh->hp[t]);
        sh->lead_val = v;
      }
    }
    P.sync();
  }
  if (own_last != kNone) {
    Elt v = out[own_last];
    uint32_t t = tid + 1;
    while (t < nth && sh->hr[t] == own_last) {
      v = F::add(v, sh->hp[t]);  // sum of this warp's part of the run starting at t
      t = (t | 31u) + 1;         // first thread of the next warp
    }
    if (CL && P.cl && t >= nth) {
      for (uint32_t c = P.rank + 1; c < P.ncta; ++c) {
        const ScShared<F>* rs = P.remote(sh, c);
        if (rs->lead_seg != own_last) break;
        v = F::add(v, rs->lead_val);
      }
    }
    out[own_last] = v;
  }
  P.sync();
}

// serial part of one round, thread 0 only (prover_layers.h:244-251,320-329,
// transcript_sumcheck.h:63-79, poly.h:59-98)
// (S: ScShared<F> in the per-proof kernel, ScLane<F> -- a thread's own copy -- in k_sc_round)
template <class F, class S>
__device__ __noinline__ void sc_round_serial(S* sh, typename F::Elt a0, typename F::Elt a2,
                                            const typename F::Elt* pad /* hp[hand][round] k=0,2 */,
                                            typename F::Elt* proof0, typename F::Elt* proof2,
                                            typename F::Elt* hb_out) {
  typedef typename F::Elt Elt;
  // coefficients of p(t) = c0 + c1 t + c2 t^2 (eq0 == 1 because logc == 0)
  Elt c0 = a0, c2 = a2;
  Elt c1 = F::sub(F::sub(F::sub(sh->sum, c0), c0), c2);
  // p(0) and p(x2): Horner in the evaluation point (prover_layers.h:395-399)
  long long q0 = clock64();
  Elt ev2 = F::add(F::mul_x2(F::add(F::mul_x2(c2), c1)), c0);
  Elt p0 = F::sub(c0, pad[0]), p2 = F::sub(ev2, pad[1]);
  *proof0 = p0;
  *proof2 = p2;
  long long q1 = clock64();
  ts_write_elt<F>(&sh->ts, p0);
  ts_write_elt<F>(&sh->ts, p2);
  long long q2 = clock64();
  Elt rnd = ts_challenge<F>(&sh->ts);
  long long q3 = clock64();
  sh->prof[4] += q1 - q0;
  sh->prof[5] += q2 - q1;
  sh->prof[6] += q3 - q2;
  *hb_out = rnd;
  sh->r = rnd;
  // the new claim p(rnd) is not needed before the next round's serial part: the
  // caller computes it (sc_new_claim) after releasing the other threads
  sh->pc[0] = c0;
  sh->pc[1] = c1;
  sh->pc[2] = c2;
}
// new claim = p(rnd).  The reference evaluates the Lagrange form through
// Newton differences (poly.h:59-98); it is the same polynomial, so Horner on
// the monomial coefficients gives the same field element with two multiplies.
template <class F, class S>
__device__ __forceinline__ void sc_new_claim(S* sh) {
  const typename F::Elt rnd = sh->r;
  sh->sum = F::add(F::mul(F::add(F::mul(sh->pc[2], rnd), sh->pc[1]), rnd), sh->pc[0]);
}

template <class F>
__device__ __noinline__ void sc_begin(ScShared<F>* sh, const Transcript* src) {
  sh->ts = *src;
  sh->ts.use_tables(&sh->aes);
  sh->ts.have_prf = 0;  // Transcript::clone() carries only the hash (transcript.h:86)
  // begin_circuit: Q[40] then G[40] (transcript_sumcheck.h:49-52)
  for (int i = 0; i < 40; ++i) (void)ts_challenge<F>(&sh->ts);
  for (int i = 0; i < 40; ++i) {
    typename F::Elt g = ts_challenge<F>(&sh->ts);
    sh->G[0][i] = g;
    sh->G[1][i] = g;
  }
  sh->wc[0] = F::zero();
  sh->wc[1] = F::zero();
  sh->fail = 0;
}

template <class F>
__device__ __noinline__ void sc_begin_layer(ScShared<F>* sh, typename F::Elt* alpha_out) {
  sh->alpha = ts_challenge<F>(&sh->ts);
  sh->beta = ts_challenge<F>(&sh->ts);
  *alpha_out = sh->alpha;
  sh->sum = F::add(sh->wc[0], F::mul(sh->alpha, sh->wc[1]));
}

template <class F>
__device__ __noinline__ void sc_end_layer(ScShared<F>* sh, typename F::Elt hquad, typename F::Elt w0,
                                         typename F::Elt w1, const typename F::Elt* padwc,
                                         typename F::Elt* proofwc, typename F::Elt* bq_out) {
  typedef typename F::Elt Elt;
  Elt expect = F::mul(hquad, F::mul(w0, w1));
  if (!F::eq(expect, sh->sum)) sh->fail = 1;
  sh->wc[0] = w0;
  sh->wc[1] = w1;
  *bq_out = hquad;
  Elt t0 = F::sub(w0, padwc[0]), t1 = F::sub(w1, padwc[1]);
  proofwc[0] = t0;
  proofwc[1] = t1;
  sh->ts.begin_array(2);
  ts_array_elt<F>(&sh->ts, t0);
  ts_array_elt<F>(&sh->ts, t1);
}

// The kernel body is shared by the launch configurations below (throughput:
// many small CTAs per SM so that one proof's serial transcript hides under the
// others' parallel work; latency: one large CTA per proof, or -- CL -- a
// thread-block cluster per proof, the parallel phases spread over its SMs and
// the leader CTA's thread 0 running the transcript).
template <class F, bool CL>
__device__ __forceinline__ void sumcheck_body(const ZkDims& d, const ZkBufs<typename F::Elt>& b,
                                              const uint32_t* __restrict__ arena,
                                              const LayerDesc* __restrict__ layers,
                                              const StepDesc* __restrict__ steps,
                                              const typename F::Elt* __restrict__ consts, ScShared<F>& sh,
                                              const ScRange R) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const ScPar<CL> P;
  const ScPar<CL> Psolo = P.solo();
  const size_t p = CL ? blockIdx.x / P.ncta : blockIdx.x;
  const uint32_t tid = P.tid, nth = P.nth;
  uint32_t gtid = P.gtid, gnth = P.gnth;
  const bool leader = P.gtid == 0;  // thread 0 of the (leader) CTA: the transcript thread
  if (b.status[p] != 0) return;   // witness already rejected by eval_circuit (uniform over the proof)
  // the leader CTA's shared state, as seen from this CTA
  const ScShared<F>* lead = P.remote(&sh, 0);

  Elt* wl = b.wl + p * d.wl_elts;
  Elt* whbuf = b.wh + p * 4 * (size_t)d.max_nw;
  Elt* hqbuf = b.hq + p * 2 * (size_t)d.max_hq;
  Elt* E0 = b.eq + p * 3 * (size_t)d.max_eq;
  Elt* E1 = E0 + d.max_eq;
  Elt* QW = E1 + d.max_eq;
  Elt* sc = b.sc + p * d.sc_elts;
  const Elt* wit = b.wit + p * d.nw;
  Elt* hbs = b.hb + p * d.nhb;

  aes_stage_tables(&sh.aes);
  ScCore<F>* core = reinterpret_cast<ScCore<F>*>(b.scst + p * sizeof(ScCore<F>));
  if (leader) {
    for (int i = 0; i < 16; ++i) sh.prof[i] = 0;
    sh.prof[3] = clock64();
    if (R.first) sc_begin<F>(&sh, reinterpret_cast<const Transcript*>(b.ts + p * sizeof(Transcript)));
    else sc_load<F>(&sh, core);
  }
  P.sync();

  uint32_t logv = R.ly_begin == 0 ? d.logv : layers[R.ly_begin - 1].logw;
  for (uint32_t ly = R.ly_begin; ly < R.ly_end; ++ly) {
    const LayerDesc L = layers[ly];
    // a launch may pick a layer up at round t0 > 0: the earlier rounds ran as flat kernels
    const uint32_t t0 = ly == R.ly_begin ? R.t_begin : 0;
    const long long tl0 = clock64();
    if (t0 == 0) {
    if (leader) {
      sc_begin_layer<F>(&sh, &b.alphas[p * d.nl + ly]);
      E0[0] = F::one();
      E1[0] = sh.alpha;
    }
    P.sync();
    // EQ tables: E0[i] = EQ(G0, i), E1[i] = alpha * EQ(G1, i)  (eqs.h:46-78)
    for (uint32_t l = 0; l < logv; ++l) {
      const uint32_t S = 1u << l;
      const Elt g0 = lead->G[0][l], g1 = lead->G[1][l];
      for (uint32_t i = gtid; i < 2 * S; i += gnth) {
        uint32_t kk = i & (S - 1);
        Elt* E = (i < S) ? E0 : E1;
        Elt v = E[kk], hi = F::mul(v, (i < S) ? g0 : g1);
        E[kk] = F::sub(v, hi);
        E[kk + S] = hi;
      }
      P.sync();
    }
    // Quad::bind_g (quad.h:152-185): initial HQuad values, one segment per corner
    {
      const uint32_t *tg = arena + L.bg_g, *tv = arena + L.bg_vi;
      const Elt beta = lead->beta;
      struct BgOps {
        Elt e0, e1, c;
        uint32_t v;
      };
      seg_sum<F, CL>(
          &sh, P, L.nterms, arena + L.bg_seg, hqbuf,
          [&](uint32_t t) {
            uint32_t g = tg[t], v = tv[t];
            BgOps o;
            o.e0 = E0[g];
            o.e1 = E1[g];
            o.v = v;
            o.c = (v & (kViOne | kViZero)) ? beta : consts[v & kViMask];
            return o;
          },
          [&](Acc& acc, const BgOps& o) {
            Elt dot = F::add(o.e0, o.e1);
            if (o.v & kViOne) F::acc_add_elt(acc, dot);
            else F::mac(acc, o.c, dot);
          });
    }
    }  // t0 == 0

    // current array of each hand: hand 0 has been bound ceil(t0 / 2) times, hand 1 floor(t0 / 2) times;
    // bind number k (from 0) of hand h writes whbuf[2h + (k & 1)]; the HQuad buffers alternate every round
    const uint32_t nb0 = (t0 + 1) >> 1, nb1 = t0 >> 1;
    const Elt* wcur0 = nb0 ? whbuf + (size_t)((nb0 - 1) & 1) * d.max_nw : wl + L.w_off;
    const Elt* wcur1 = nb1 ? whbuf + (size_t)(2 + ((nb1 - 1) & 1)) * d.max_nw : wl + L.w_off;
    uint32_t wpar0 = nb0 & 1, wpar1 = nb1 & 1;
    uint32_t hqpar = t0 & 1;
    const Elt* pad = wit + d.n_witness + L.pad_off;

    bool solo = false;
    if (leader) sh.prof[7] += clock64() - tl0;
    for (uint32_t t = t0; t < 2 * L.logw; ++t) {
      const StepDesc S = steps[L.step0 + t];
      const uint32_t hand = t & 1, round = t >> 1;
      const long long tr0 = clock64();
      // Cluster mode: once a layer's steps are small, a cluster-wide phase costs
      // more (barrier + an L2 round trip per phase, ~2 us) than its work; the
      // leader CTA then finishes the layer alone with CTA barriers and the other
      // CTAs wait at the end-of-layer barrier.  (Steps only shrink within a layer.)
      if (CL && !solo && S.n_in + S.n0 <= d.solo_work) solo = true;
      if (CL && solo && P.rank != 0) continue;  // pointer bookkeeping below is only needed by the leader
      const ScPar<CL>& Q = (CL && solo) ? Psolo : P;
      gtid = Q.gtid;
      gnth = Q.gnth;
      const Elt* Wh = hand ? wcur1 : wcur0;
      const Elt* Wo = hand ? wcur0 : wcur1;
      const Elt* HQ = hqbuf + (size_t)hqpar * d.max_hq;
      // QW[l] = sum_r Q[l,r] W[r]  (prover_layers.h:230-243)
      {
        const uint32_t* roff = arena + S.row_off;
        for (uint32_t i = gtid; i < S.n0; i += gnth)
          if (roff[i] == roff[i + 1]) QW[i] = F::zero();
        const uint32_t *rc = arena + S.row_c, *rp = arena + S.row_p1;
        struct QwOps {
          Elt q, w;
        };
        seg_sum<F, CL>(
            &sh, Q, S.n_in, arena + S.row_r, QW,
            [&](uint32_t e) {
              QwOps o;
              o.q = HQ[rc[e]];
              o.w = Wo[rp[e]];
              return o;
            },
            [&](Acc& acc, const QwOps& o) { F::mac(acc, o.q, o.w); });
      }
      // the two dot products of ProverLayers::evaluations (prover_layers.h:357-402)
      const uint32_t npair = (S.n0 + 1) / 2;
      const long long tr1 = clock64();
      {
        Acc a0, a2;
        F::acc_zero(a0);
        F::acc_zero(a2);
        for (uint32_t i = gtid; i < npair; i += gnth) {
          Elt qw0 = QW[2 * i], w0 = Wh[2 * i];
          // odd tail (prover_layers.h:377-384): a2 += qw0*w0 = (0-qw0)*(0-w0)
          const bool two = 2 * i + 1 < S.n0;
          Elt qw1 = two ? QW[2 * i + 1] : F::zero(), w1 = two ? Wh[2 * i + 1] : F::zero();
          F::mac(a0, qw0, w0);
          F::mac(a2, F::sub(qw1, qw0), F::sub(w1, w0));
        }
        Elt s0 = warp_sum<F>(F::reduce(a0)), s2 = warp_sum<F>(F::reduce(a2));
        if ((tid & 31) == 0) {
          sh.red[0][tid >> 5] = s0;
          sh.red[1][tid >> 5] = s2;
        }
        __syncthreads();
        // warp 0 folds the CTA's warps and hands the pair to the leader CTA
        if (tid < 32) {
          const uint32_t nw = nth >> 5;
          Elt r0 = tid < nw ? sh.red[0][tid] : F::zero(), r2 = tid < nw ? sh.red[1][tid] : F::zero();
          r0 = warp_sum<F>(r0);
          r2 = warp_sum<F>(r2);
          if (tid == 0) {
            ScShared<F>* ld = Q.remote(&sh, 0);
            ld->cred[0][Q.rank] = r0;
            ld->cred[1][Q.rank] = r2;
          }
        }
      }
      Q.sync();
      long long tp0 = clock64();
      if (leader) {
        sh.prof[8] += tr1 - tr0;
        sh.prof[9] += tp0 - tr1;
        if (CL && solo) sh.prof[11] += 1;
        Elt s0 = sh.cred[0][0], s2 = sh.cred[1][0];
        for (uint32_t c = 1; c < Q.ncta; ++c) {
          s0 = F::add(s0, sh.cred[0][c]);
          s2 = F::add(s2, sh.cred[1][c]);
        }
        // pad order: (round, hand, k in {0,2}); proof order: (round, k, hand)
        sc_round_serial<F>(&sh, s0, s2, pad + 4 * round + 2 * hand, sc + L.sc_off + 4 * round + hand,
                           sc + L.sc_off + 4 * round + 2 + hand, hbs + L.hb_off + t);
        sh.G[hand][round] = sh.r;
        sh.prof[0] += clock64() - tp0;
        sh.prof[1] += 1;
      }
      Q.sync();
      const long long tb0 = clock64();
      const Elt r = Q.remote(&sh, 0)->r;
      if (leader) sc_new_claim<F>(&sh);  // off the critical path: the others are already binding
      // Dense::bind (dense.h:70-89) and HQuad::bind_h (hquad.h:89-123) through the merge plan, over ONE index
      // space: in the small rounds both lists fit the threads once, and two loops would be two latency
      // chains (load, multiply, store) one after the other
      Elt* Wn = whbuf + (size_t)(2 * hand + (hand ? wpar1 : wpar0)) * d.max_nw;
      Elt* HQn = hqbuf + (size_t)(hqpar ^ 1) * d.max_hq;
      const uint32_t* mg = arena + S.merge;
      for (uint32_t k = gtid; k < npair + S.n_out; k += gnth) {
        Elt f0, f1;
        Elt* dst;
        if (k < npair) {
          // affine_interpolation_nz_z(r, f0) == affine_interpolation(r, f0, 0) (affine.h:25-52)
          f0 = Wh[2 * k];
          f1 = (2 * k + 1 < S.n0) ? Wh[2 * k + 1] : F::zero();
          dst = Wn + k;
        } else {
          // pair: (v0, v1); lone even corner: (v0, 0); lone odd corner: (0, v0) -- the
          // three affine_interpolation variants of hquad.h:99-115 are one formula
          const uint32_t j = k - npair;
          const uint32_t m = mg[j], src = m >> 2, kind = m & 3;
          const Elt v = HQ[src];
          f0 = kind == 2 ? F::zero() : v;
          f1 = kind == 0 ? HQ[src + 1] : (kind == 2 ? v : F::zero());
          dst = HQn + j;
        }
        *dst = affine<F>(r, f0, f1);
      }
      Q.sync();
      if (leader) sh.prof[10] += clock64() - tb0;
      if (hand) {
        wcur1 = Wn;
        wpar1 ^= 1;
      } else {
        wcur0 = Wn;
        wpar0 ^= 1;
      }
      hqpar ^= 1;
    }
    gtid = P.gtid;
    gnth = P.gnth;
    // end of layer (prover_layers.h:263-270,331-344)
    if (leader)
      sc_end_layer<F>(&sh, hqbuf[(size_t)hqpar * d.max_hq], wcur0[0], wcur1[0], pad + 4 * L.logw,
                      sc + L.sc_off + 4 * L.logw, &b.bq[p * d.nl + ly]);
    P.sync();
    logv = L.logw;
  }
  if (leader) {
    if (R.ly_end == d.nl) {
      *reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript)) = sh.ts;
      if (sh.fail) atomicCAS(&b.status[p], 0, -100);  // (the first error of a proof sticks) internal inconsistency: never expected
      sh.prof[2] = clock64() - sh.prof[3];
      long long* dbg = reinterpret_cast<long long*>(hqbuf);  // free after the last layer
      const int nd = (int)min((size_t)16, 2 * (size_t)d.max_hq * sizeof(Elt) / 8);  // tiny circuits: a short buffer
      for (int i = 0; i < nd; ++i) dbg[i] = sh.prof[i];
    } else {
      sc_save<F>(&sh, core);  // the flat kernels of the next layer continue from here
    }
  }
}

// One kernel body, four shapes.  NT threads per CTA, MINB resident CTAs per SM:
//   <128, 8>  throughput (default): 8 proofs per SM in flight
//   < 64,16>  16 proofs per SM (LF_SC_THREADS=64; measured no faster)
//   <1024,1>  latency: one CTA per proof on its own SM
//   cluster   latency, very few proofs: a thread-block cluster of 8 or 16 CTAs
//             (one per SM) per proof; launched with cudaLaunchKernelEx
constexpr int kScTpThreads = 128;
template <class F, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_zk_sumcheck(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena,
              const LayerDesc* __restrict__ layers, const StepDesc* __restrict__ steps,
              const typename F::Elt* __restrict__ consts, const ScRange R) {
  __shared__ ScShared<F> sh;
  __shared__ typename F::Elt hp[NT];
  __shared__ uint32_t hr[NT];
  if (threadIdx.x == 0) {
    sh.hp = hp;
    sh.hr = hr;
  }
  __syncthreads();
  sumcheck_body<F, false>(d, b, arena, layers, steps, consts, sh, R);
}
constexpr int kScClThreads = 512;
template <class F>
__global__ void __launch_bounds__(kScClThreads, 1)
k_zk_sumcheck_cluster(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena,
                      const LayerDesc* __restrict__ layers, const StepDesc* __restrict__ steps,
                      const typename F::Elt* __restrict__ consts, ScRange R) {
  __shared__ ScShared<F> sh;
  __shared__ typename F::Elt hp[kScClThreads];
  __shared__ uint32_t hr[kScClThreads];
  if (threadIdx.x == 0) {
    sh.hp = hp;
    sh.hr = hr;
  }
  __syncthreads();
  sumcheck_body<F, true>(d, b, arena, layers, steps, consts, sh, R);
}

// ----------------------------------------------------------------------------
// k_lig_challenges: thread 0 of one CTA per proof.  alpha of the input
// constraint (zk_common.h:131), write(hash_of_A) (zk_prover.h:143,
// ligero_prover.h:91-95), then u_ldt, alphal, alphaq, u_quad which the
// reference draws back to back from one PRF stream (ligero_prover.h:97-125).
// chal layout: [0] alpha_in | u_ldt[nwqrow] | alphal[nl+1] | alphaq[3nq] | u_quad[nqtriples]
// ----------------------------------------------------------------------------
template <class F>
__global__ void k_lig_challenges(ZkDims d, ZkBufs<typename F::Elt> b, size_t nproofs) {
  __shared__ AesTables s_aes;
  aes_stage_tables(&s_aes);
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs || b.status[p] != 0) return;
  Transcript* gts = reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript));
  Transcript ts = *gts;
  ts.use_tables(&s_aes);
  typename F::Elt* chal = b.chal + p * (size_t)(1 + d.nchal);
  chal[0] = F::ts_elt(&ts);
  uint32_t hashA[8] = {0xefbeaddeu, 0, 0, 0, 0, 0, 0, 0};  // bytes de ad be ef 00 ...
  ts.write_bytes_words(hashA, 8);
  for (uint32_t i = 0; i < d.nchal; ++i) chal[1 + i] = F::ts_elt(&ts);
  *gts = ts;
}

// EQ table over the input wires, bindings = hand challenges of the last layer
// (zk_common.h:412-413): E0 = EQ(hb[0]), E1 = EQ(hb[1]); one CTA per proof.
template <class F>
__global__ void __launch_bounds__(256)
k_lig_input_eq(ZkDims d, ZkBufs<typename F::Elt> b, LayerDesc last) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  Elt* E0 = b.eq + p * 3 * (size_t)d.max_eq;
  Elt* E1 = E0 + d.max_eq;
  const Elt* hb = b.hb + p * d.nhb + last.hb_off;  // [2*round + hand]
  if (threadIdx.x == 0) {
    E0[0] = F::one();
    E1[0] = F::one();
  }
  __syncthreads();
  for (uint32_t l = 0; l < last.logw; ++l) {
    const uint32_t S = 1u << l;
    const Elt g0 = hb[2 * l], g1 = hb[2 * l + 1];
    for (uint32_t i = threadIdx.x; i < 2 * S; i += blockDim.x) {
      uint32_t k = i & (S - 1);
      if (i < S) {
        Elt v = E0[k], hi = F::mul(v, g0);
        E0[k] = F::sub(v, hi);
        E0[k + S] = hi;
      } else {
        Elt v = E1[k], hi = F::mul(v, g1);
        E1[k] = F::sub(v, hi);
        E1[k + S] = hi;
      }
    }
    __syncthreads();
  }
}

// Poly<3>::dot_interpolation::coef (poly.h:139-146): lag[k] = L_k(x)
template <class F>
__device__ __forceinline__ void lagrange3(const typename F::Elt& x, typename F::Elt lag[3]) {
  typedef typename F::Elt Elt;
  Elt d0 = F::sub(x, F::evalpt(0)), d1 = F::sub(x, F::evalpt(1));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Elt e = F::lag_id(k, 2);
    e = F::add(F::mul(e, d1), F::lag_id(k, 1));
    e = F::add(F::mul(e, d0), F::lag_id(k, 0));
    lag[k] = e;
  }
}

// ----------------------------------------------------------------------------
// k_lig_avec: the inner-product vector A of LigeroCommon::inner_product_vector
// (ligero_param.h:382-421) assembled directly from the symbolic sumcheck
// verifier (zk_common.h:49-136,291-439) instead of materialising the sparse
// constraint list:  for layer ly with rounds t = 2*round+hand and Lagrange
// weights lag_t = coef(hb_t), the coefficient of
//    poly pad (t,0):  (lag_t[0] - lag_t[1]) * P_t      poly pad (t,2): lag_t[2] * P_t
//    claim pad of the previous layer: (1, alpha_ly, 0) * P_{-1}
//    claim pad of this layer: (-eqq*wc[1], -eqq*wc[0], -eqq)
// with P_t = prod_{t' > t} lag_t'[1], each multiplied by alphal[ly].
// One CTA per proof; A is zero-filled first.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_lig_avec(ZkDims d, ZkBufs<typename F::Elt> b, const LayerDesc* __restrict__ layers) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  Elt* A = b.avec + p * (size_t)d.nwqrow * d.w;
  const Elt* chal = b.chal + p * (size_t)(1 + d.nchal);
  const Elt alpha_in = chal[0];
  const Elt* alphal = chal + 1 + d.nwqrow;
  const Elt* alphaq = alphal + (d.nl + 1);
  const Elt* hb = b.hb + p * d.nhb;
  const Elt* sc = b.sc + p * d.sc_elts;
  const Elt* bq = b.bq + p * d.nl;
  const Elt* alphas = b.alphas + p * d.nl;
  Elt* lagbuf = reinterpret_cast<Elt*>(b.scratch + p * b.scratch_words);  // [nhb][4]: lag0,lag1,lag2,P
  const Elt* E0 = b.eq + p * 3 * (size_t)d.max_eq;
  const Elt* E1 = E0 + d.max_eq;

  const uint32_t na = d.nwqrow * d.w;
  for (uint32_t i = tid; i < na; i += nth) {
    Elt v = F::zero();
    if (i < d.n_witness) {
      // input constraint (zk_common.h:414-421): b_i = eq0[i] + alpha * eq1[i]
      uint32_t k = i + d.npub;
      v = F::mul(F::add(E0[k], F::mul(alpha_in, E1[k])), alphal[d.nl]);
    }
    A[i] = v;
  }
  // Lagrange weights of every round
  for (uint32_t t = tid; t < d.nhb; t += nth) {
    Elt lag[3];
    lagrange3<F>(hb[t], lag);
    lagbuf[4 * t + 0] = lag[0];
    lagbuf[4 * t + 1] = lag[1];
    lagbuf[4 * t + 2] = lag[2];
  }
  __syncthreads();
  // suffix products per layer (one thread per layer; 2*logw sequential multiplies)
  for (uint32_t ly = tid; ly < d.nl; ly += nth) {
    const LayerDesc L = layers[ly];
    Elt P = F::one();
    for (uint32_t t = 2 * L.logw; t-- > 0;) {
      lagbuf[4 * (L.hb_off + t) + 3] = P;
      P = F::mul(P, lagbuf[4 * (L.hb_off + t) + 1]);
    }
    // P is now P_{-1}.  claim pads: combine this layer's finalize terms with
    // the next layer's first() terms (or the input constraint after the last)
    const Elt eqq = bq[ly];
    const Elt wc0 = sc[L.sc_off + 4 * L.logw], wc1 = sc[L.sc_off + 4 * L.logw + 1];
    const Elt al = alphal[ly];
    uint32_t cp = d.n_witness + L.pad_off + 4 * L.logw;
    // this layer's own terms (zk_common.h:381-384)
    Elt c0 = F::neg(F::mul(F::mul(eqq, wc1), al));
    Elt c1 = F::neg(F::mul(F::mul(eqq, wc0), al));
    Elt c2 = F::neg(F::mul(eqq, al));
    // stash P_{-1} for the previous layer's claim pads
    lagbuf[4 * L.hb_off + 0] = lagbuf[4 * L.hb_off + 0];  // (no-op; keeps layout explicit)
    // quadratic routing (ligero_param.h:408-419): A[lqc.x] -= alphaq[i][0] ...
    if (ly < d.nq) {
      c0 = F::sub(c0, alphaq[3 * ly + 0]);
      c1 = F::sub(c1, alphaq[3 * ly + 1]);
      c2 = F::sub(c2, alphaq[3 * ly + 2]);
    }
    A[cp + 0] = F::add(A[cp + 0], c0);
    A[cp + 1] = F::add(A[cp + 1], c1);
    A[cp + 2] = F::add(A[cp + 2], c2);
    // store P_{-1} * alphal[ly] and alpha_ly for the pass below
    Elt* extra = lagbuf + 4 * (size_t)d.nhb + 2 * ly;
    extra[0] = F::mul(P, al);
    extra[1] = F::mul(F::mul(P, al), alphas[ly]);
  }
  __syncthreads();
  // previous-layer claim pads (cb.first, zk_common.h:330-335) and the input constraint tail
  for (uint32_t ly = tid; ly <= d.nl; ly += nth) {
    if (ly == 0) continue;  // layer 0 does not refer to CLAIM_PAD[-1] (zk_common.h:390-391)
    const LayerDesc Lp = layers[ly - 1];
    uint32_t cp = d.n_witness + Lp.pad_off + 4 * Lp.logw;
    Elt c0, c1;
    if (ly < d.nl) {
      const Elt* extra = lagbuf + 4 * (size_t)d.nhb + 2 * ly;
      c0 = extra[0];
      c1 = extra[1];
    } else {
      // zk_common.h:433-436: (-1, -alpha) * alphal[nl]
      c0 = F::neg(alphal[d.nl]);
      c1 = F::neg(F::mul(alpha_in, alphal[d.nl]));
    }
    A[cp + 0] = F::add(A[cp + 0], c0);
    A[cp + 1] = F::add(A[cp + 1], c1);
  }
  // poly pads
  for (uint32_t t = tid; t < d.nhb; t += nth) {
    // find the layer of round t
    uint32_t ly = 0;
    while (ly + 1 < d.nl && layers[ly + 1].hb_off <= t) ++ly;
    const LayerDesc L = layers[ly];
    uint32_t tt = t - L.hb_off;
    Elt P = F::mul(lagbuf[4 * t + 3], alphal[ly]);
    uint32_t base = d.n_witness + L.pad_off + 2 * tt;
    A[base + 0] = F::mul(F::sub(lagbuf[4 * t + 0], lagbuf[4 * t + 1]), P);
    A[base + 1] = F::mul(lagbuf[4 * t + 2], P);
  }
  // quadratic rows Ax, Ay, Az (ligero_param.h:402-418)
  for (uint32_t i = tid; i < d.nq; i += nth) {
    uint32_t ax = d.nwrow * d.w;
    A[ax + i] = alphaq[3 * i + 0];
    A[ax + d.nqtriples * d.w + i] = alphaq[3 * i + 1];
    A[ax + 2 * d.nqtriples * d.w + i] = alphaq[3 * i + 2];
  }
}

// Aext rows for dot_proof (ligero_param.h:423-430): [0^r | A_i[w]], one CTA per (row, proof)
template <class F>
__global__ void k_lig_aext(ZkDims d, ZkBufs<typename F::Elt> b) {
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t i = blockIdx.x;
  typename F::Elt* X = b.aext + (p * d.nwqrow + i) * (size_t)d.dblock;
  const typename F::Elt* A = b.avec + p * (size_t)d.nwqrow * d.w + (size_t)i * d.w;
  for (uint32_t j = threadIdx.x; j < d.block; j += blockDim.x) X[j] = (j < d.r) ? F::zero() : A[j - d.r];
}

// y_ldt, y_dot, y_quad (ligero_prover.h:281-344): one thread per column.
// y layout: [0,block) ldt | [block, block+dblock) dot | [block+dblock, block+2*dblock) quad
template <class F>
__global__ void __launch_bounds__(128)
k_lig_columns(ZkDims d, ZkBufs<typename F::Elt> b) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d.dblock) return;
  const Elt* T = b.tableau + p * (size_t)d.nrow * d.block_enc;
  const Elt* chal = b.chal + p * (size_t)(1 + d.nchal);
  const Elt* u_ldt = chal + 1;
  const Elt* u_quad = chal + 1 + d.nwqrow + (d.nl + 1) + 3 * d.nq;
  Elt* y = b.y + p * (size_t)(d.block + 2 * d.dblock);
  const size_t ld = d.block_enc;
  typename F::Acc acc;
  if (j < d.block) {
    F::acc_zero(acc);
    for (uint32_t i = 0; i < d.nwqrow; ++i) F::mac(acc, T[(size_t)(i + d.iw) * ld + j], u_ldt[i]);
    y[j] = F::add(T[j], F::reduce(acc));
  }
  {
    const Elt* X = b.aext + p * (size_t)d.nwqrow * d.dblock;
    F::acc_zero(acc);
    for (uint32_t i = 0; i < d.nwqrow; ++i) F::mac(acc, T[(size_t)(i + d.iw) * ld + j], X[(size_t)i * d.dblock + j]);
    y[d.block + j] = F::add(T[ld + j], F::reduce(acc));
  }
  {
    F::acc_zero(acc);
    const uint32_t iqx = d.iq, iqy = iqx + d.nqtriples, iqz = iqy + d.nqtriples;
    for (uint32_t i = 0; i < d.nqtriples; ++i) {
      Elt tmp = F::sub(T[(size_t)(iqz + i) * ld + j], F::mul(T[(size_t)(iqy + i) * ld + j], T[(size_t)(iqx + i) * ld + j]));
      F::mac(acc, tmp, u_quad[i]);
    }
    y[d.block + d.dblock + j] = F::add(T[2 * ld + j], F::reduce(acc));
  }
}

// ----------------------------------------------------------------------------
// k_lig_finish: one CTA per proof.  Thread 0 writes the four response arrays
// into the transcript and draws the opened columns (ligero_prover.h:127-145,
// random.h:92-105); the CTA then marks the Merkle opening, run-length codes the
// opened columns and serializes the whole proof (zk_proof.h:114-184).
// scratch layout (uint32): perm[block_ext] | mark bytes[2*block_ext] | flag bytes[nreq*nrow]
//                          | eoff[nreq*nrow] | path_idx[nreq*mc_pathlen]
// ----------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void put_elt(uint8_t* dst, const typename F::Elt& e) {
  uint32_t w[F::kWords];
  F::to_wire(w, e);
  if ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
    uint32_t* q = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
    for (int k = 0; k < F::kWords; ++k) q[k] = w[k];
  } else {
#pragma unroll
    for (int k = 0; k < F::kWords; ++k) {
      dst[4 * k] = (uint8_t)w[k];
      dst[4 * k + 1] = (uint8_t)(w[k] >> 8);
      dst[4 * k + 2] = (uint8_t)(w[k] >> 16);
      dst[4 * k + 3] = (uint8_t)(w[k] >> 24);
    }
  }
}
__device__ __forceinline__ void put_u32(uint8_t* dst, uint32_t g) {
  dst[0] = (uint8_t)g; dst[1] = (uint8_t)(g >> 8); dst[2] = (uint8_t)(g >> 16); dst[3] = (uint8_t)(g >> 24);
}
__device__ __forceinline__ void put_digest(uint8_t* dst, const uint32_t* be_words) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t x = be_words[k];
    dst[4 * k] = (uint8_t)(x >> 24); dst[4 * k + 1] = (uint8_t)(x >> 16);
    dst[4 * k + 2] = (uint8_t)(x >> 8); dst[4 * k + 3] = (uint8_t)x;
  }
}

// The part of LigeroProver::prove that the prover and the verifier share (ligero_prover.h:127-145,
// ligero_verifier.h:83-90): the four response arrays y_ldt | y_dot | y_quad_0 | y_quad_2 enter the
// transcript and the opened columns are drawn (RandomEngine::choose, random.h:92-105).  Called by
// all threads of a CTA; y in the layout of ZkBufs::y; msg: scratch for the message bytes
// (>= 64 + 36 + (block + 2 dblock + r - block) kBytes ... i.e. what the arrays serialise to);
// perm [n], mark [2n bytes] scratch; on return idx[0..nreq) and mark[n + idx] = 1, *gts updated.
// mode 0: everything; 1: only assemble the message bytes in msg (k_lig_msg); 2: the message has been absorbed
// by k_lig_hash (one thread per proof, for large batches), only the columns are drawn
template <class F>
__device__ __forceinline__ void lig_absorb_and_choose(const ZkDims& d, const typename F::Elt* __restrict__ y,
                                                      Transcript* gts, uint8_t* msg, uint32_t* perm, uint8_t* mark,
                                                      uint32_t* idx, const AesTables* aes, int mode = 0) {
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const uint32_t n = d.block_ext;
  if (mode != 1) {
    for (uint32_t i = tid; i < n; i += nth) perm[i] = i;
    for (uint32_t i = tid; i < 2 * n; i += nth) mark[i] = 0;
  }
  if (mode == 2) {
    __syncthreads();
    if (tid == 0) {
      Transcript ts = *gts;
      ts.use_tables(aes);
      ts.have_prf = 0;
      for (uint32_t i = 0; i < d.nreq; ++i) {
        uint32_t j = i + ts.nat(n - i);
        uint32_t t = perm[i];
        perm[i] = perm[j];
        perm[j] = t;
        idx[i] = perm[i];
        mark[perm[i] + n] = 1;
      }
      *gts = ts;
    }
    __syncthreads();
    return;
  }
  // ---- the four response arrays enter the transcript (ligero_prover.h:84-146) ----
  // Their bytes are known up front, so only the 64 rounds per block have to be serial:
  // the CTA assembles the byte stream (the transcript's pending buffer bytes, the array
  // headers, the wire bytes of the elements -- for prime fields to_wire is a Montgomery
  // multiplication) at the start of the output slot, which is rewritten below, and expands
  // the message schedules of a chunk of blocks into shared memory; thread 0 runs the rounds.
  constexpr uint32_t kChunk = 32;
  __shared__ uint32_t s_w[kChunk][64];
  __shared__ uint32_t s_pos0, s_h[8];
  __shared__ uint64_t s_len0;
  const uint32_t lens[4] = {d.block, d.dblock, d.r, d.dblock - d.block};
  const uint32_t offs[4] = {0, d.block, d.block + d.dblock, d.block + d.dblock + d.block};
  if (tid == 0) {
    const Sha256& sh0 = gts->sha;
    const uint32_t pos0 = (uint32_t)(sh0.len & 63);
    s_pos0 = pos0;
    s_len0 = sh0.len;
    for (int k = 0; k < 8; ++k) s_h[k] = sh0.h[k];
    for (uint32_t k = 0; k < pos0; ++k) msg[k] = (uint8_t)(sh0.buf[k >> 2] >> (24 - 8 * (k & 3)));
  }
  __syncthreads();
  const uint32_t pos0 = s_pos0;
  uint32_t abase[5];
  abase[0] = pos0;
  for (int a = 0; a < 4; ++a) abase[a + 1] = abase[a] + 9 + lens[a] * F::kBytes;
  const uint32_t tbytes = abase[4];  // bytes of the pending block(s) once everything is appended
  if (tid < 4) {
    uint8_t* h = msg + abase[tid];
    h[0] = 2;  // TAG_ARRAY, then the 64-bit little-endian length (transcript.h:144-153,160-171)
    for (int k = 0; k < 8; ++k) h[1 + k] = k < 4 ? (uint8_t)(lens[tid] >> (8 * k)) : 0;
  }
  const uint32_t nelt = lens[0] + lens[1] + lens[2] + lens[3];
  for (uint32_t i = tid; i < nelt; i += nth) {
    uint32_t a = 0, j = i;
    while (j >= lens[a]) j -= lens[a++];
    uint32_t w[F::kWords];
    F::to_wire(w, y[offs[a] + j]);
    uint8_t* o = msg + abase[a] + 9 + (size_t)j * F::kBytes;
#pragma unroll
    for (int q = 0; q < F::kWords; ++q) {
      o[4 * q] = (uint8_t)w[q];
      o[4 * q + 1] = (uint8_t)(w[q] >> 8);
      o[4 * q + 2] = (uint8_t)(w[q] >> 16);
      o[4 * q + 3] = (uint8_t)(w[q] >> 24);
    }
  }
  __syncthreads();
  if (mode == 1) return;
  const uint32_t nblk = tbytes / 64;
  for (uint32_t c0 = 0; c0 < nblk; c0 += kChunk) {
    const uint32_t cnt = min(kChunk, nblk - c0);
    for (uint32_t j = tid; j < cnt; j += nth) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(msg + (size_t)(c0 + j) * 64);
      uint32_t w16[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) w16[k] = bswap32(src[k]);
      sha256_expand(w16, s_w[j]);
    }
    __syncthreads();
    if (tid == 0)
      for (uint32_t j = 0; j < cnt; ++j) sha256_rounds_fn(s_h, s_w[j]);
    __syncthreads();
  }
  if (tid == 0) {
    Transcript ts = *gts;
    ts.use_tables(aes);
    ts.have_prf = 0;  // a write drops the challenge stream (transcript.h:174-178)
    for (int k = 0; k < 8; ++k) ts.sha.h[k] = s_h[k];
    ts.sha.len = s_len0 + (tbytes - pos0);
    for (int k = 0; k < 16; ++k) ts.sha.buf[k] = 0;
    for (uint32_t k = nblk * 64; k < tbytes; ++k)
      ts.sha.buf[(k & 63) >> 2] |= (uint32_t)msg[k] << (24 - 8 * (k & 3));
    // RandomEngine::choose (random.h:92-105)
    for (uint32_t i = 0; i < d.nreq; ++i) {
      uint32_t j = i + ts.nat(n - i);
      uint32_t t = perm[i];
      perm[i] = perm[j];
      perm[j] = t;
      idx[i] = perm[i];
      mark[perm[i] + n] = 1;
    }
    *gts = ts;
  }
  __syncthreads();
}

// Large batches: the responses enter the transcript in two steps.  k_lig_msg (one CTA per proof) assembles the
// byte stream at the start of the output slot; k_lig_hash (one THREAD per proof: a warp hashes 32 proofs in
// lockstep) runs the compressions.  In k_lig_finish one thread of a 128-thread CTA runs them, which is the
// faster way for a handful of proofs and 17x the issue slots for a thousand.
template <class F>
__global__ void __launch_bounds__(128, 8)
k_lig_msg(ZkDims d, ZkBufs<typename F::Elt> b) {
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  lig_absorb_and_choose<F>(d, b.y + p * (size_t)(d.block + 2 * d.dblock),
                           reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript)), b.out + p * b.out_stride, nullptr,
                           nullptr, nullptr, nullptr, 1);
}
// msg: per proof msg_stride bytes as k_lig_msg / lig_absorb_and_choose(mode 1) left them: the transcript's pending
// bytes, then the four arrays with their headers; kbytes = bytes per element
__global__ void __launch_bounds__(32)
k_lig_hash(ZkDims d, uint8_t* ts_base, const uint8_t* msg, size_t msg_stride, const int32_t* status, uint32_t kbytes,
           size_t nproofs) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs || status[p] != 0) return;
  Transcript* gts = reinterpret_cast<Transcript*>(ts_base + p * sizeof(Transcript));
  uint32_t h[8];
  for (int k = 0; k < 8; ++k) h[k] = gts->sha.h[k];
  const uint64_t len0 = gts->sha.len;
  const uint32_t pos0 = (uint32_t)(len0 & 63);
  const uint32_t tbytes = pos0 + 4 * 9 + (2 * d.dblock + d.r) * kbytes;
  const uint32_t nblk = tbytes / 64;
  const uint32_t* m32 = reinterpret_cast<const uint32_t*>(msg + p * msg_stride);
  for (uint32_t j = 0; j < nblk; ++j) {
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = bswap32(m32[16 * (size_t)j + k]);
    sha256_compress(h, w);
  }
  for (int k = 0; k < 8; ++k) gts->sha.h[k] = h[k];
  gts->sha.len = len0 + (tbytes - pos0);
  const uint8_t* mb = msg + p * msg_stride;
  uint32_t buf[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) buf[k] = 0;
  for (uint32_t k = nblk * 64; k < tbytes; ++k) buf[(k & 63) >> 2] |= (uint32_t)mb[k] << (24 - 8 * (k & 3));
#pragma unroll
  for (int k = 0; k < 16; ++k) gts->sha.buf[k] = buf[k];
  gts->have_prf = 0;
}

// 128 threads, 8 CTAs per SM: the kernel is bound by one thread per proof running the
// SHA-256 rounds, so a whole batch of 1024 proofs should be resident at once (at 256
// threads and 124 registers only 2 CTAs fitted an SM: 3.5 waves).
template <class F>
__global__ void __launch_bounds__(128, 8)
k_lig_finish(ZkDims d, ZkBufs<typename F::Elt> b, const LayerDesc* __restrict__ layers, int hashed) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) {
    if (threadIdx.x == 0) b.out_len[p] = 0;
    return;
  }
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  const Elt* T = b.tableau + p * (size_t)d.nrow * d.block_enc;
  const Elt* y = b.y + p * (size_t)(d.block + 2 * d.dblock);
  const Elt* sc = b.sc + p * d.sc_elts;
  const uint32_t* nodes = b.nodes + p * (size_t)(2 * d.block_ext * 8);
  // the nonces follow the last sample: redrawn slots push them back (k_zk_rng_scan)
  const uint8_t* nonces = b.rng + p * b.rng_stride + d.rng_nonce_off +
                          (b.rej ? (size_t)b.rej[p * (size_t)(1 + d.rej_cap)] * F::kBytes : 0);
  uint32_t* idx = b.idx + p * d.nreq;
  uint8_t* out = b.out + p * b.out_stride;
  uint32_t* sw = b.scratch + p * b.scratch_words;
  const uint32_t n = d.block_ext, total = d.nreq * d.nrow;
  uint32_t* perm = sw;
  uint8_t* mark = reinterpret_cast<uint8_t*>(sw + n);
  uint8_t* flag = mark + 2 * (size_t)((n + 3) & ~3u);
  uint32_t* eoff = reinterpret_cast<uint32_t*>(flag + ((total + 3) & ~3u));
  uint32_t* path_idx = eoff + total;
  __shared__ uint32_t s_npath, s_req_end;
  __shared__ AesTables s_aes;
  aes_stage_tables(&s_aes);

  lig_absorb_and_choose<F>(d, y, reinterpret_cast<Transcript*>(b.ts + p * sizeof(Transcript)), out, perm, mark, idx,
                           &s_aes, hashed ? 2 : 0);
  // compressed_merkle_proof_tree (merkle_tree.h:75-98), level by level
  if (n >= 2) {
    int top = 31 - __clz(n - 1);
    for (int lv = top; lv >= 0; --lv) {
      uint32_t lo = 1u << lv, hi = min(2u << lv, n);
      for (uint32_t i = lo + tid; i < hi; i += nth) mark[i] = mark[2 * i] | mark[2 * i + 1];
      __syncthreads();
    }
  }
  // opened columns: subfield flags (zk_proof.h:162-166)
  for (uint32_t k = tid; k < total; k += nth) {
    Elt e = T[(size_t)(k / d.nreq) * d.block_enc + d.dblock + idx[k % d.nreq]];
    uint32_t u = 0;
    bool in_sub = true;  // prime fields: in_subfield() is always true (fp_generic.h:278)
    if (F::kChar2) in_sub = F::solve_sub16(e, &u);
    flag[k] = (uint8_t)in_sub;
  }
  __syncthreads();
  const uint32_t off_sc = 32;
  const uint32_t off_y = off_sc + d.sc_elts * F::kBytes;
  const uint32_t off_nonce = off_y + (d.block + d.dblock + d.r + (d.dblock - d.block)) * F::kBytes;
  const uint32_t off_req = off_nonce + d.nreq * 32;
  if (tid == 0) {
    // run-length layout: runs alternate full / subfield, starting with full
    uint32_t o = off_req, ci = 0;
    uint32_t subrun = 0;
    while (ci < total) {
      uint32_t runlen = 0;
      while (ci + runlen < total && (uint32_t)flag[ci + runlen] == subrun) ++runlen;
      put_u32(out + o, runlen);
      o += 4;
      uint32_t sz = subrun ? F::kSubBytes : F::kBytes;
      for (uint32_t k = ci; k < ci + runlen; ++k) {
        eoff[k] = o;
        o += sz;
      }
      ci += runlen;
      subrun ^= 1;
    }
    s_req_end = o;
    // MerkleTree::generate_compressed_proof (merkle_tree.h:122-143)
    uint32_t sz = 0;
    for (uint32_t i = n; i-- > 1;) {
      if (mark[i]) {
        uint32_t child = 2 * i;
        if (mark[child]) child = 2 * i + 1;
        if (!mark[child]) path_idx[sz++] = child;
      }
    }
    s_npath = sz;
    put_u32(out + o, sz);
    b.out_len[p] = (uint64_t)o + 4 + 32ull * sz;
  }
  __syncthreads();
  // ---- serialize (zk_proof.h:114-184) ----
  if (tid < 8) {
    uint32_t x = nodes[8 + tid];  // root = node 1
    out[4 * tid] = (uint8_t)(x >> 24); out[4 * tid + 1] = (uint8_t)(x >> 16);
    out[4 * tid + 2] = (uint8_t)(x >> 8); out[4 * tid + 3] = (uint8_t)x;
  }
  for (uint32_t i = tid; i < d.sc_elts; i += nth) put_elt<F>(out + off_sc + (size_t)i * F::kBytes, sc[i]);
  {
    // y_ldt | y_dot | y_quad_0 (first r of quad) | y_quad_2 (quad[block..dblock))
    const uint32_t n1 = d.block + d.dblock, n2 = n1 + d.r, n3 = n2 + (d.dblock - d.block);
    for (uint32_t i = tid; i < n3; i += nth) {
      Elt e = i < n1 ? y[i] : (i < n2 ? y[n1 + (i - n1)] : y[n1 + d.block + (i - n2)]);
      put_elt<F>(out + off_y + (size_t)i * F::kBytes, e);
    }
  }
  for (uint32_t i = tid; i < d.nreq * 8; i += nth) {
    uint32_t q = i >> 3, k = i & 7;
    const uint8_t* src = nonces + 32ull * idx[q] + 4 * k;
    uint8_t* dst = out + off_nonce + 32 * q + 4 * k;
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
  }
  for (uint32_t k = tid; k < total; k += nth) {
    Elt e = T[(size_t)(k / d.nreq) * d.block_enc + d.dblock + idx[k % d.nreq]];
    if (flag[k] && F::kChar2) {
      uint32_t u;
      F::solve_sub16(e, &u);
      out[eoff[k]] = (uint8_t)u;
      out[eoff[k] + 1] = (uint8_t)(u >> 8);
    } else {
      put_elt<F>(out + eoff[k], e);
    }
  }
  {
    const uint32_t base = s_req_end + 4;
    for (uint32_t i = tid; i < s_npath; i += nth) put_digest(out + base + 32ull * i, nodes + 8ull * path_idx[i]);
  }
}

// sumcheck proof / witness / tableau read-back helpers use plain memcpy on the host side

}  // namespace lf
