// Shared host/device descriptors of the batched ZK prover pipeline.
//
// A circuit is uploaded once (lf_circuit_upload): the reference's delta-coded
// Quad (lib/sumcheck/quad.h:56-226) is expanded and turned into STATIC PLANS,
// because everything about the sparse structure of the sumcheck -- which quad
// terms merge in Quad::bind_g (quad.h:152-185), which corners pair up in every
// HQuad::bind_h round (hquad.h:89-123), which corners feed which QW entry
// (prover_layers.h:230-243) -- depends only on the circuit's indices, never on
// the witness.  The kernels therefore never sort, scan or compact: they gather
// through precomputed CSR lists, which also makes every sum a deterministic
// per-thread sequence (no atomics; required for prime fields).
#pragma once
#include <stdint.h>

namespace lf {

// flags stored in the top bits of a constant index
constexpr uint32_t kViZero = 0x80000000u;  // constant == 0: assert-zero term
constexpr uint32_t kViOne = 0x40000000u;   // constant == 1: skip the multiply
constexpr uint32_t kViMask = 0x3fffffffu;

struct LayerDesc {
  uint32_t nw, logw, nterms;
  uint32_t nout;     // number of output wires (= nw of the layer above, or nv)
  uint32_t nhq0;     // HQuad corners right after bind_g
  // offsets into the uint32 arena
  uint32_t ev_off;   // [nout+1] CSR by gate: eval_circuit
  uint32_t ev_h0, ev_h1, ev_vi;  // [nterms]
  uint32_t ev_perm;              // [nout] gates sorted by decreasing term count
  uint32_t bg_seg;   // [nterms] initial HQuad corner of each term (canonical order, non-decreasing)
  uint32_t bg_g, bg_vi;          // [nterms]
  uint32_t vq_h0, vq_h1;         // [nhq0] the two hand indices of each initial corner (verifier: bind_gh_all)
  uint32_t step0;    // first StepDesc of this layer (2*logw of them)
  uint32_t w_off;    // element offset of this layer's input wires in the per-proof wire store
  uint32_t out_off;  // element offset of this layer's output wires
  uint32_t pad_off;  // element offset of this layer's pad inside the Ligero witness
  uint32_t sc_off;   // element offset of this layer inside the sumcheck proof (= pad rng index)
  uint32_t hb_off;   // offset of this layer's 2*logw hand challenges
};

struct StepDesc {
  uint32_t n_in;     // corners before the round
  uint32_t n_out;    // corners after bind_h
  uint32_t n0;       // length of the hand's wire array before the round
  uint32_t row_off;  // [n0+1] CSR by p0 = h[hand]
  uint32_t row_r;    // [n_in] p0 of each CSR entry (non-decreasing)
  uint32_t row_c;    // [n_in] corner index
  uint32_t row_p1;   // [n_in] h[other hand]
  uint32_t merge;    // [n_out] (src << 2) | kind ; kind 0 pair, 1 lone even, 2 lone odd
};

// ---- batch-synchronous ("flat") sumcheck rounds ---------------------------------------------
// With hundreds of proofs in a batch the large rounds of a layer are run as ordinary grid-wide
// kernels over (work item, proof) instead of inside one CTA per proof: k_sc_eval (QW gather and the
// two dot products of ProverLayers::evaluations), k_sc_round (the serial Fiat-Shamir step, one
// THREAD per proof, so a warp runs the SHA-256/AES chains of 32 proofs in lockstep) and k_sc_bind
// (Dense::bind and HQuad::bind_h).  The work lists are static per circuit:
//
// k_sc_eval: one lane per PAIR of rows (2i, 2i+1) of the hand's CSR; pairs are sorted by their
// number of entries and stored 32 to a warp in column-major (sliced-ELL) order, so index loads are
// coalesced and the 32 lanes loop about equally often.  Rows with more than kFlatHeavyRow entries
// (the constant wire 0 feeds tens of thousands of corners) are taken out of their pair and cut
// into warp-sized chunks whose partial QW enters the two dot products by linearity.
constexpr uint32_t kFlatHeavyRow = 48;     // entries of one row a pair thread still takes
constexpr uint32_t kFlatHeavyChunk = 256;  // entries per heavy-row warp chunk
constexpr uint32_t kFlatEvalWarps = 4;     // warps per CTA of k_sc_eval
#ifndef LF_EVAL_MINCTA
#define LF_EVAL_MINCTA 5
#endif
constexpr uint32_t kFlatEvalMinCta = LF_EVAL_MINCTA;    // resident CTAs per SM the kernel is compiled for (register cap)
// warps (bins of work) per proof of k_sc_eval, by batch size: a large batch fills the GPU with few warps per
// proof; a handful of proofs of a large circuit (one mdoc hash proof is 72 M multiplications) need thousands
constexpr uint32_t kFlatBinModes = 3;
constexpr uint32_t kFlatBinCap[kFlatBinModes] = {32, 256, 4096};   // mode 0: B >= 128, 1: B >= 8, 2: B < 8
constexpr uint32_t kFlatBinModeMaxB[kFlatBinModes] = {0xffffffffu, 127, 7};
constexpr uint32_t kFlatSumDirect = 64;    // k_sc_round adds up to this many partials itself; above, k_sc_partsum first
constexpr uint32_t kFlatBinCost = 8;       // smallest bin worth a warp, in multiplications per lane
constexpr uint32_t kFlatNone = 0xffffffffu;

struct FlatStepDesc {
  uint32_t nwarp_pair;   // sliced-ELL warps of row pairs
  uint32_t nwarp_heavy;  // heavy-row chunks, one warp each
  // per bin mode: warps per proof, each takes the items bin_item[bin_off[w] .. bin_off[w+1])
  uint32_t nbin[kFlatBinModes];
  uint32_t bin_off[kFlatBinModes];   // [nbin + 1]
  uint32_t bin_item[kFlatBinModes];  // [nwarp_pair + nwarp_heavy] item < nwarp_pair: sliced-ELL warp; else heavy chunk + nwarp_pair
  uint32_t pw_pair;      // [32 * nwarp_pair] pair index of each lane (kFlatNone: padding)
  uint32_t pw_cnt;       // [32 * nwarp_pair] entries of row 2i | entries of row 2i+1 << 16
  uint32_t pw_base;      // [nwarp_pair] entry offset of the warp's column-major block
  uint32_t e_c, e_p;     // entry arrays: corner index, wire index of the other hand
  uint32_t hv_row, hv_off, hv_cnt;  // [nwarp_heavy] row, first entry and entry count inside the step's CSR
};
// Quad::bind_g of a large layer: one thread per initial HQuad corner, same layout; corners with
// more than kFlatHeavyRow terms are cut into warp chunks whose partial sums k_sc_bindg_fix adds.
struct FlatLayerDesc {
  uint32_t nwarp_c, nwarp_heavy, ncta;
  uint32_t cw_corner, cw_cnt, cw_base;  // [32 * nwarp_c], [32 * nwarp_c], [nwarp_c]
  uint32_t t_g, t_v;                    // term arrays in sliced-ELL order: gate index, constant index + flags
  uint32_t hv_off, hv_cnt;              // [nwarp_heavy] first term / term count in canonical order
  uint32_t nheavy;                      // heavy corners
  uint32_t hc_corner, hc_item;          // [nheavy] corner, [nheavy + 1] first chunk of each heavy corner
  uint32_t nflat;                       // rounds 0 .. nflat-1 of this layer run flat
  uint32_t fs0;                         // index of the first FlatStepDesc of this layer
};

struct ZkDims {
  uint32_t nl, ninputs, npub, n_witness, nv, logv, nterms;
  // LigeroParam (lib/ligero/ligero_param.h:116-147)
  uint32_t nw, nq, block_enc, block, dblock, block_ext, r, w, nwrow, nqtriples, nwqrow, nrow, nreq,
      mc_pathlen, iw, iq;
  uint32_t sb;        // subfield boundary rebased to the private inputs (zk_prover.h:84-89)
  uint32_t pad_size;  // sum of 4*logw+3
  uint32_t sc_elts;   // sum of 4*logw+2: elements of the sumcheck proof = random pad elements
  uint32_t nhb;       // sum of 2*logw
  uint32_t max_nw, max_hq, max_eq;
  uint32_t wl_elts;   // per-proof wire store size (elements)
  uint32_t nchal;     // nwqrow + (nl+1) + 3*nq + nqtriples Ligero challenges
  // caller-supplied randomness layout (bytes from the start of one proof's stream)
  uint32_t rng_nonce_off;
  uint32_t rng_total;
  // Prime fields: every caller-random element before the nonces is one kBytes slot of the
  // stream, and a slot whose value is >= p makes the reference draw again
  // (algebra/fp_generic.h:360-371).  rng_nsamples = number of samples (0 for GF(2^128), whose
  // samples never fail); rej_cap = redraws per proof that are followed.
  uint32_t rng_nsamples;
  uint32_t rej_cap;
  uint32_t max_proof_bytes;
  uint32_t tinit_len;
  uint32_t debug_stop;  // LF_DEBUG_STOP: early exit point inside k_lig_finish (bisecting)
  uint32_t solo_work;   // cluster sumcheck: steps with n_in + n0 <= this run on the leader CTA alone
};

// per-proof device buffers: base pointer + stride (in elements of the pointee)
template <class Elt>
struct ZkBufs {
  const uint8_t* witness_in;  size_t witness_stride;   // ninputs * kBytes wire bytes
  const uint8_t* rng;         size_t rng_stride;
  size_t rng_avail;           // valid bytes of each proof's stream (<= rng_stride)
  uint32_t* rej;   // [1 + rej_cap] prime fields: number of rejected slots, then for each (in stream
                   // order) the index of the sample that was drawn again (k_zk_rng_scan); else null
  Elt* wit;        // [nw] Ligero witness
  Elt* tableau;    // [nrow * block_enc]
  uint32_t* nodes; // [2 * block_ext * 8] Merkle heap, big-endian digest words
  Elt* wl;         // [wl_elts] wires of every layer
  Elt* wh;         // [4 * max_nw] hand arrays, ping-pong
  Elt* hq;         // [2 * max_hq] HQuad values, ping-pong
  Elt* eq;         // [3 * max_eq] EQ tables E0, E1 and the QW array
  Elt* sc;         // [sc_elts] sumcheck proof in wire order
  Elt* bq;         // [nl] bound quads (ProofAux)
  Elt* hb;         // [nhb] hand challenges
  Elt* alphas;     // [nl] per-layer alpha
  uint8_t* scst;   // [sizeof(ScCore<F>)] sumcheck prover state between the kernels of the flat path
  Elt* part;       // [part_stride] per-warp partial sums of the two dot products of a round
  size_t part_stride;
  Elt* chal;       // [1 + nchal] alpha_in, u_ldt, alphal, alphaq, u_quad
  Elt* avec;       // [nwqrow * w]
  Elt* aext;       // [nwqrow * dblock]
  Elt* y;          // [block + 2*dblock] y_ldt | y_dot | y_quad
  uint32_t* idx;   // [nreq] opened columns
  uint32_t* scratch;  // [scratch_words]
  uint8_t* ts;     // [sizeof(Transcript)]
  uint8_t* out;    size_t out_stride;
  uint64_t* out_len;
  int32_t* status;
  size_t scratch_words;
};

}  // namespace lf
