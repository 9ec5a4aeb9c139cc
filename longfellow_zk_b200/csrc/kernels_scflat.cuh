// Batch-synchronous ("flat") sumcheck rounds for large batches.
//
// k_zk_sumcheck runs one proof per CTA, rounds and Fiat-Shamir transcript included.  That keeps
// the 2*sum(logw) sequential rounds on one SM, but while thread 0 hashes, the other warps of the
// CTA wait, and the register budget of that one big kernel caps the resident warps (ncu, round 1:
// 30 % of the warp samples of the GF(2^128) instance and 44 % of the P-256 instance at the barrier
// behind the serial round, warps active 34 %).  Every proof of a batch runs the SAME circuit, so
// the large rounds of a layer are instead run for the whole batch at once, as plain grid-wide
// kernels over (work item, proof) with nothing serial in them:
//
//   reference                                              here
//   Eqs / EQ tables (arrays/eqs.h:46-78)                   k_sc_begin (half tables) + k_sc_eq
//   Quad::bind_g (sumcheck/quad.h:152-185)                 k_sc_bindg (+ k_sc_bindg_fix)
//   QW gather + ProverLayers::evaluations                  k_sc_eval
//     (sumcheck/prover_layers.h:230-243,357-402)
//   round polynomial -> transcript -> challenge            k_sc_round   (one thread per PROOF: a warp
//     (prover_layers.h:244-251,320-329,                                  runs the SHA-256 / AES chains
//      transcript_sumcheck.h:63-79)                                      of 32 proofs in lockstep)
//   Dense::bind, HQuad::bind_h (dense.h:70-89,             k_sc_bind
//     hquad.h:89-123)
//
// Once a layer's rounds are small (StepDesc work below the flat threshold) the per-proof kernel
// k_zk_sumcheck picks the layer up at that round (ScRange) and finishes it -- those rounds are
// bound by the latency of the transcript, not by arithmetic -- together with any small layers
// that follow.  All sums are exact field sums, so regrouping them cannot change a result: the
// proof bytes are those of the per-proof kernel and of the reference.
#pragma once
#include <stdint.h>

#include "field.cuh"
#include "hash.cuh"
#include "kernels_zk.cuh"
#include "zk_types.cuh"

namespace lf {

// arrays of the current round, by the same conventions as sumcheck_body (kernels_zk.cuh)
template <class F>
struct FlatPtrs {
  typedef typename F::Elt Elt;
  const Elt *Wh, *Wo, *HQ;
  Elt *Wn, *HQn;
  __device__ __forceinline__ FlatPtrs(const ZkDims& d, const ZkBufs<Elt>& b, const LayerDesc& L, size_t p,
                                      uint32_t t) {
    Elt* wl = b.wl + p * d.wl_elts;
    Elt* whbuf = b.wh + p * 4 * (size_t)d.max_nw;
    Elt* hqbuf = b.hq + p * 2 * (size_t)d.max_hq;
    const uint32_t hand = t & 1, nb0 = (t + 1) >> 1, nb1 = t >> 1;
    const Elt* w0 = nb0 ? whbuf + (size_t)((nb0 - 1) & 1) * d.max_nw : wl + L.w_off;
    const Elt* w1 = nb1 ? whbuf + (size_t)(2 + ((nb1 - 1) & 1)) * d.max_nw : wl + L.w_off;
    Wh = hand ? w1 : w0;
    Wo = hand ? w0 : w1;
    Wn = whbuf + (size_t)(2 * hand + ((hand ? nb1 : nb0) & 1)) * d.max_nw;
    HQ = hqbuf + (size_t)(t & 1) * d.max_hq;
    HQn = hqbuf + (size_t)((t & 1) ^ 1) * d.max_hq;
  }
};

// ----------------------------------------------------------------------------
// k_sc_begin: start of a large layer, one CTA per proof.  Thread 0: the transcript (begin_circuit
// for the first layer of the proof, then alpha and beta: transcript_sumcheck.h:49-61).  All
// threads: the EQ tables of the layer's output bindings as a tensor product,
//   EQ(G, i) = EQ(G_lo, i mod 2^h) * EQ(G_hi, i div 2^h),    h = logv / 2,
// whose two factors (at most 2^8 entries each for logv <= 16) are built here by the doubling
// recurrence of eqs.h:46-78; E1's factor starts from alpha as in sumcheck_body.
// half tables: eq[0 .. 4 * 2^hh) = lo0 | hi0 | lo1 | hi1   (hh = logv - h >= h)
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_sc_begin(ZkDims d, ZkBufs<typename F::Elt> b, LayerDesc L, uint32_t ly, uint32_t logv, uint32_t first) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  __shared__ ScShared<F> sh;
  ScCore<F>* core = reinterpret_cast<ScCore<F>*>(b.scst + p * sizeof(ScCore<F>));
  aes_stage_tables(&sh.aes);
  if (threadIdx.x == 0) {
    if (first) sc_begin<F>(&sh, reinterpret_cast<const Transcript*>(b.ts + p * sizeof(Transcript)));
    else sc_load<F>(&sh, core);
    sc_begin_layer<F>(&sh, &b.alphas[p * d.nl + ly]);
    sc_save<F>(&sh, core);
  }
  __syncthreads();
  const uint32_t h = logv / 2, hh = logv - h, S1 = 1u << hh;
  Elt* half = b.eq + p * 3 * (size_t)d.max_eq + 2 * (size_t)d.max_eq;  // the QW array is free here
  // four independent recurrences, one warp each: lo0, hi0, lo1, hi1
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    const uint32_t which = warp & 1, tab = warp >> 1;  // which: 0 low bits, 1 high bits; tab: E0 / E1
    Elt* T = half + (size_t)warp * S1;
    const uint32_t nb = which ? hh : h, g_off = which ? h : 0;
    if (lane == 0) T[0] = (tab == 1 && which == 1) ? sh.alpha : F::one();
    __syncwarp();
    for (uint32_t l = 0; l < nb; ++l) {
      const uint32_t S = 1u << l;
      const Elt g = sh.G[tab][g_off + l];
      for (uint32_t k = lane; k < S; k += 32) {
        Elt v = T[k], hi = F::mul(v, g);
        T[k] = F::sub(v, hi);
        T[k + S] = hi;
      }
      __syncwarp();
    }
  }
}

// E[i] = lo0[i & m] * hi0[i >> h] + lo1[i & m] * hi1[i >> h]  =  E0[i] + E1[i], i < nout
template <class F>
__global__ void __launch_bounds__(256)
k_sc_eq(ZkDims d, ZkBufs<typename F::Elt> b, uint32_t nout, uint32_t logv) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  const uint32_t h = logv / 2, hh = logv - h, S1 = 1u << hh, m = (1u << h) - 1;
  Elt* E = b.eq + p * 3 * (size_t)d.max_eq;
  const Elt* half = E + 2 * (size_t)d.max_eq;
  const Elt a = F::mul(half[i & m], half[S1 + (i >> h)]);
  const Elt c = F::mul(half[2 * S1 + (i & m)], half[3 * S1 + (i >> h)]);
  E[i] = F::add(a, c);
}

// ----------------------------------------------------------------------------
// k_sc_bindg: initial HQuad values of a large layer (quad.h:152-185),
//   hq[corner] = sum over the corner's terms of v * (E0[g] + E1[g])   (v == 0 stands for beta),
// one thread per corner in sliced-ELL order; heavy corners in warp chunks whose partial sums go to
// the scratch array and are added by k_sc_bindg_fix.
// ----------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void bindg_term(typename F::Acc& acc, const typename F::Elt* __restrict__ E,
                                           const typename F::Elt* __restrict__ consts, const typename F::Elt& beta,
                                           uint32_t g, uint32_t v) {
  const typename F::Elt dot = E[g];
  if (v & kViOne) F::acc_add_elt(acc, dot);
  else F::mac(acc, (v & kViZero) ? beta : consts[v & kViMask], dot);
}

template <class F>
__global__ void __launch_bounds__(256)
k_sc_bindg(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena, LayerDesc L, FlatLayerDesc FL,
           const typename F::Elt* __restrict__ consts) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + warp;
  const Elt* E = b.eq + p * 3 * (size_t)d.max_eq;
  Elt* part = b.eq + p * 3 * (size_t)d.max_eq + d.max_eq;  // E1's array is free in the flat path
  Elt* hq = b.hq + p * 2 * (size_t)d.max_hq;
  const Elt beta = reinterpret_cast<const ScCore<F>*>(b.scst + p * sizeof(ScCore<F>))->beta;
  if (wid < FL.nwarp_c) {
    const uint32_t slot = wid * 32 + lane;
    const uint32_t corner = arena[FL.cw_corner + slot];
    if (corner == kFlatNone) return;
    const uint32_t cnt = arena[FL.cw_cnt + slot], base = arena[FL.cw_base + wid];
    Acc acc;
    F::acc_zero(acc);
    for (uint32_t k = 0; k < cnt; ++k) {
      const uint32_t e = base + k * 32 + lane;
      bindg_term<F>(acc, E, consts, beta, arena[FL.t_g + e], arena[FL.t_v + e]);
    }
    hq[corner] = F::reduce(acc);
  } else if (wid < FL.nwarp_c + FL.nwarp_heavy) {
    const uint32_t it = wid - FL.nwarp_c;
    const uint32_t off = arena[FL.hv_off + it], cnt = arena[FL.hv_cnt + it];
    Acc acc;
    F::acc_zero(acc);
    for (uint32_t e = off + lane; e < off + cnt; e += 32)
      bindg_term<F>(acc, E, consts, beta, arena[L.bg_g + e], arena[L.bg_vi + e]);
    const Elt v = warp_sum<F>(F::reduce(acc));
    if (lane == 0) part[it] = v;
  }
}
template <class F>
__global__ void k_sc_bindg_fix(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena,
                               FlatLayerDesc FL) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= FL.nheavy) return;
  const Elt* part = b.eq + p * 3 * (size_t)d.max_eq + d.max_eq;
  Elt* hq = b.hq + p * 2 * (size_t)d.max_hq;
  Elt v = F::zero();
  for (uint32_t it = arena[FL.hc_item + i]; it < arena[FL.hc_item + i + 1]; ++it) v = F::add(v, part[it]);
  hq[arena[FL.hc_corner + i]] = v;
}

// ----------------------------------------------------------------------------
// k_sc_eval: the two sums of one round (prover_layers.h:357-402),
//   a0 = sum_i QW[2i] W[2i],   a2 = sum_i (QW[2i+1] - QW[2i]) (W[2i+1] - W[2i]),
// with QW[l] = sum_r Q[l, r] W'[r] (prover_layers.h:230-243) computed on the fly per row pair and
// never stored.  grid = (ceil(nbin / warps per CTA), proofs); every warp owns one BIN of work items and leaves
// its partial (a0, a2) in part[2 * bin ..]; k_sc_round adds the nbin partials of a proof.
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(32 * kFlatEvalWarps, kFlatEvalMinCta)
k_sc_eval(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena, LayerDesc L, StepDesc S,
          FlatStepDesc FS, uint32_t t, uint32_t mode) {
  typedef typename F::Elt Elt;
  typedef typename F::Acc Acc;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bin = blockIdx.x * kFlatEvalWarps + warp;
  if (bin >= FS.nbin[mode]) return;
  const FlatPtrs<F> P(d, b, L, p, t);
  const Elt *Wh = P.Wh, *Wo = P.Wo, *HQ = P.HQ;
  Acc c0, c2;
  F::acc_zero(c0);
  F::acc_zero(c2);
  // the bin's items (balanced on the host by their entry counts): sliced-ELL warps of row pairs, then heavy-row chunks
  const uint32_t i0 = arena[FS.bin_off[mode] + bin], i1 = arena[FS.bin_off[mode] + bin + 1];
  for (uint32_t ii = i0; ii < i1; ++ii) {
    const uint32_t item = arena[FS.bin_item[mode] + ii];
    if (item < FS.nwarp_pair) {
      const uint32_t wb = item, slot = wb * 32 + lane;
      const uint32_t pair = arena[FS.pw_pair + slot];
      if (pair == kFlatNone) continue;
      const uint32_t cnt = arena[FS.pw_cnt + slot], n0c = cnt & 0xffffu, tot = n0c + (cnt >> 16);
      const uint32_t base = arena[FS.pw_base + wb];
      Acc q0, q1;
      F::acc_zero(q0);
      F::acc_zero(q1);
      for (uint32_t k = 0; k < tot; ++k) {
        const uint32_t e = base + k * 32 + lane;
        F::mac_sel(q0, q1, k < n0c, HQ[arena[FS.e_c + e]], Wo[arena[FS.e_p + e]]);
      }
      const Elt qw0 = F::reduce(q0), qw1 = F::reduce(q1);
      const Elt w0 = Wh[2 * pair], w1 = (2 * pair + 1 < S.n0) ? Wh[2 * pair + 1] : F::zero();
      F::mac(c0, qw0, w0);
      F::mac(c2, F::sub(qw1, qw0), F::sub(w1, w0));
    } else {
      // heavy row: the chunk's partial QW enters both sums by linearity
      const uint32_t hw = item - FS.nwarp_pair;
      const uint32_t row = arena[FS.hv_row + hw], off = arena[FS.hv_off + hw], cnt = arena[FS.hv_cnt + hw];
      const uint32_t *rc = arena + S.row_c, *rp = arena + S.row_p1;
      Acc q;
      F::acc_zero(q);
      for (uint32_t e = off + lane; e < off + cnt; e += 32) F::mac(q, HQ[rc[e]], Wo[rp[e]]);
      const Elt v = warp_sum<F>(F::reduce(q));
      if (lane == 0) {
        if ((row & 1) == 0) {
          const Elt w0 = Wh[row], w1 = (row + 1 < S.n0) ? Wh[row + 1] : F::zero();
          F::mac(c0, v, w0);
          F::mac(c2, F::neg(v), F::sub(w1, w0));
        } else {
          F::mac(c2, v, F::sub(Wh[row], Wh[row - 1]));
        }
      }
    }
  }
  // no CTA-level reduction (a barrier here made every warp wait for the CTA's slowest): one partial per warp
  const Elt s0 = warp_sum<F>(F::reduce(c0)), s2 = warp_sum<F>(F::reduce(c2));
  if (lane == 0) {
    Elt* part = b.part + p * b.part_stride;
    part[2 * bin] = s0;
    part[2 * bin + 1] = s2;
  }
}

// many bins (a few proofs of a large circuit): the partials of a proof are added by one CTA first, so that
// k_sc_round's single thread per proof reads two values.  part[0], part[1] = the sums.
template <class F>
__global__ void __launch_bounds__(256)
k_sc_partsum(ZkBufs<typename F::Elt> b, uint32_t nbin) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.x;
  if (b.status[p] != 0) return;
  Elt* part = b.part + p * b.part_stride;
  Elt s0 = F::zero(), s2 = F::zero();
  for (uint32_t i = threadIdx.x; i < nbin; i += blockDim.x) {
    s0 = F::add(s0, part[2 * i]);
    s2 = F::add(s2, part[2 * i + 1]);
  }
  __shared__ Elt red[2][8];
  s0 = warp_sum<F>(s0);
  s2 = warp_sum<F>(s2);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = s0;
    red[1][warp] = s2;
  }
  __syncthreads();  // also: every thread has read its partials before part[0..1] is overwritten
  if (threadIdx.x == 0) {
    for (uint32_t w = 1; w < blockDim.x / 32; ++w) {
      s0 = F::add(s0, red[0][w]);
      s2 = F::add(s2, red[1][w]);
    }
    part[0] = s0;
    part[1] = s2;
  }
}

// ----------------------------------------------------------------------------
// k_sc_round: the serial part of one round for 32 proofs per warp, one thread per proof
// (prover_layers.h:244-251,320-329, transcript_sumcheck.h:63-79): add the CTA partials, derive the
// round polynomial, write p(0) and p(2) (minus pad) to the proof and to the transcript, draw the
// challenge, and evaluate the new claim.  The state lives in global memory (ScCore).
// ----------------------------------------------------------------------------
template <class F>
struct ScLane {  // the members sc_round_serial / sc_new_claim touch, in a thread's local memory
  Transcript ts;
  typename F::Elt r, sum, pc[3];
  long long prof[8];
};

template <class F>
__global__ void __launch_bounds__(32)
k_sc_round(ZkDims d, ZkBufs<typename F::Elt> b, LayerDesc L, uint32_t t, uint32_t nbin, size_t nproofs) {
  typedef typename F::Elt Elt;
  __shared__ AesTables s_aes;
  aes_stage_tables(&s_aes);
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs || b.status[p] != 0) return;
  ScCore<F>* core = reinterpret_cast<ScCore<F>*>(b.scst + p * sizeof(ScCore<F>));
  const Elt* part = b.part + p * b.part_stride;
  Elt s0 = part[0], s2 = part[1];
  for (uint32_t c = 1; c < nbin; ++c) {
    s0 = F::add(s0, part[2 * c]);
    s2 = F::add(s2, part[2 * c + 1]);
  }
  ScLane<F> st;
  st.ts.sha = core->sha;
  st.ts.have_prf = 0;
  st.ts.nblock = 0;
  st.ts.rdptr = 16;
  st.ts.use_tables(&s_aes);
  st.sum = core->sum;
  const uint32_t hand = t & 1, round = t >> 1;
  const Elt* pad = b.wit + p * d.nw + d.n_witness + L.pad_off;
  Elt* sc = b.sc + p * d.sc_elts;
  Elt* hbs = b.hb + p * d.nhb;
  sc_round_serial<F>(&st, s0, s2, pad + 4 * round + 2 * hand, sc + L.sc_off + 4 * round + hand,
                     sc + L.sc_off + 4 * round + 2 + hand, hbs + L.hb_off + t);
  sc_new_claim<F>(&st);
  core->sha = st.ts.sha;
  core->G[hand][round] = st.r;
  core->r = st.r;
  core->sum = st.sum;
}

// ----------------------------------------------------------------------------
// k_sc_bind: Dense::bind of the hand's wire array (dense.h:70-89) and HQuad::bind_h through the
// merge plan (hquad.h:89-123) with the round's challenge.  grid = (ceil((npair + n_out) / 256), proofs)
// ----------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256)
k_sc_bind(ZkDims d, ZkBufs<typename F::Elt> b, const uint32_t* __restrict__ arena, LayerDesc L, StepDesc S,
          uint32_t t) {
  typedef typename F::Elt Elt;
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t npair = (S.n0 + 1) / 2;
  if (i >= npair + S.n_out) return;
  const FlatPtrs<F> P(d, b, L, p, t);
  const Elt r = reinterpret_cast<const ScCore<F>*>(b.scst + p * sizeof(ScCore<F>))->r;
  if (i < npair) {
    const Elt f0 = P.Wh[2 * i], f1 = (2 * i + 1 < S.n0) ? P.Wh[2 * i + 1] : F::zero();
    P.Wn[i] = affine<F>(r, f0, f1);
  } else {
    const uint32_t j = i - npair, m = arena[S.merge + j], src = m >> 2, kind = m & 3;
    const Elt v = P.HQ[src];
    const Elt f0 = kind == 2 ? F::zero() : v;
    const Elt f1 = kind == 0 ? P.HQ[src + 1] : (kind == 2 ? v : F::zero());
    P.HQn[j] = affine<F>(r, f0, f1);
  }
}

// ----------------------------------------------------------------------------
// k_sc_bind_tab: k_sc_bind over GF(2^128) with the products by the round's challenge taken from a table.
// Every product of a bind round has the same factor r (per proof), so the CTA builds
//   T[w][b] = r * b(x) * x^(8w)  mod  x^128 + x^7 + x^2 + x + 1,      w < 16, b < 256      (64 KB)
// once -- 128 basis elements r x^k by a shift and the usual fold, the rest as XORs of them -- and then
// r * v = XOR_w T[w][byte w of v]: 16 LDS.128 and 16 XORs instead of 144 IMAD.WIDE + 251 LOP3.  Random
// 16-byte reads conflict in the banks (about 9 clk per warp load), which makes the table form no faster
// than the multiplier on its own -- but it runs on the shared-memory pipe, which the multiplier leaves
// idle, so the elements of a thread alternate between the two forms (every `imad_every`-th goes to the
// multiplier) and the two pipes work side by side.  Same field elements either way.
// grid = (ceil((npair + n_out) / chunk), proofs), 256 threads, 64 KB dynamic shared memory.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint4 gf_tab_mul(const uint4* __restrict__ T, const gf128& v) {
  uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    const uint32_t byte = (v.w[w >> 2] >> (8 * (w & 3))) & 255u;
    const uint4 e = T[w * 256 + byte];
    acc.x ^= e.x;
    acc.y ^= e.y;
    acc.z ^= e.z;
    acc.w ^= e.w;
  }
  return acc;
}

__global__ void __launch_bounds__(256, 3)
k_sc_bind_tab(ZkDims d, ZkBufs<gf128> b, const uint32_t* __restrict__ arena, LayerDesc L, StepDesc S, uint32_t t,
              uint32_t chunk, uint32_t imad_every) {
  typedef FGf128 F;
  extern __shared__ __align__(16) uint4 bt_tab[];
  const size_t p = blockIdx.y;
  if (b.status[p] != 0) return;
  const uint32_t tid = threadIdx.x;
  const gf128 r = reinterpret_cast<const ScCore<F>*>(b.scst + p * sizeof(ScCore<F>))->r;
  // basis: r x^k, k < 128, at T[k / 8][1 << (k % 8)]; T[w][0] = 0
  if (tid < 128) {
    uint32_t wide[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wide[i] = 0;
    const uint32_t ws = tid >> 5, bs = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wide[i + ws] |= r.w[i] << bs;
      if (bs) wide[i + ws + 1] |= r.w[i] >> (32 - bs);
    }
    const gf128 e = gf_reduce(wide);
    bt_tab[(tid >> 3) * 256 + (1u << (tid & 7))] = make_uint4(e.w[0], e.w[1], e.w[2], e.w[3]);
  } else if (tid < 144) {
    bt_tab[(tid - 128) * 256] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  for (uint32_t e = tid; e < 4096; e += blockDim.x) {
    const uint32_t w = e >> 8, j = e & 255;
    if (__popc(j) < 2) continue;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (j & (1u << k)) {
        const uint4 x = bt_tab[w * 256 + (1u << k)];
        acc.x ^= x.x;
        acc.y ^= x.y;
        acc.z ^= x.z;
        acc.w ^= x.w;
      }
    bt_tab[e] = acc;
  }
  __syncthreads();
  const uint32_t npair = (S.n0 + 1) / 2, total = npair + S.n_out;
  const FlatPtrs<F> P(d, b, L, p, t);
  const uint32_t i0 = blockIdx.x * chunk, i1 = min(total, i0 + chunk);
  // operands of the next element are fetched before the current one is multiplied (the table lookups would
  // otherwise wait behind the global loads)
  auto fetch = [&](uint32_t i, gf128& f0, gf128& f1, gf128*& dst) {
    if (i < npair) {
      f0 = P.Wh[2 * i];
      f1 = (2 * i + 1 < S.n0) ? P.Wh[2 * i + 1] : F::zero();
      dst = &P.Wn[i];
    } else {
      const uint32_t j = i - npair, m = arena[S.merge + j], src = m >> 2, kind = m & 3;
      const gf128 v = P.HQ[src];
      f0 = kind == 2 ? F::zero() : v;
      f1 = kind == 0 ? P.HQ[src + 1] : (kind == 2 ? v : F::zero());
      dst = &P.HQn[j];
    }
  };
  uint32_t k = 0;
  uint32_t i = i0 + tid;
  gf128 f0, f1, nf0, nf1;
  gf128 *dst = nullptr, *ndst = nullptr;
  if (i < i1) fetch(i, f0, f1, dst);
  for (; i < i1; i += blockDim.x, ++k) {
    const uint32_t inext = i + blockDim.x;
    if (inext < i1) fetch(inext, nf0, nf1, ndst);
    // affine_interpolation: f0 + r (f1 - f0)
    const gf128 dlt = gf_add(f0, f1);
    gf128 o;
    if (imad_every && (k % imad_every) == imad_every - 1) {
      o = gf_add(f0, F::mul(r, dlt));
    } else {
      const uint4 pr = gf_tab_mul(bt_tab, dlt);
      o.w[0] = f0.w[0] ^ pr.x;
      o.w[1] = f0.w[1] ^ pr.y;
      o.w[2] = f0.w[2] ^ pr.z;
      o.w[3] = f0.w[3] ^ pr.w;
    }
    *dst = o;
    f0 = nf0;
    f1 = nf1;
    dst = ndst;
  }
}

}  // namespace lf
