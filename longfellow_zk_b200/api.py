"""Host-side mirror of the reference's prover interfaces over the C ABI.

Names and argument meaning follow the reference (lib/):
  LCH14ReedSolomonFactory.make(n, m).interpolate(y)   gf2k/lch14_reed_solomon.h:112-123
  MerkleCommitment(n).commit(...)                      merkle/merkle_commitment.h:46-64
  ZkProver(circuit).prove_batch(...)                   zk/zk_prover.h:72-149 + zk_proof.h:90-112
Elements are numpy uint8 arrays in the reference's wire encoding.
"""
import ctypes as C

import numpy as np

from . import _native
from ._native import LongfellowError, check  # noqa: F401

FIELD_P256 = 1
FIELD_GF2_128 = 4
FIELD_BN254 = 100
FIELD_FP128 = 101
FIELD_GOLDILOCKS = 102
KBYTES = {FIELD_P256: 32, FIELD_GF2_128: 16, FIELD_BN254: 32, FIELD_FP128: 16, FIELD_GOLDILOCKS: 8}


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One CUDA device + stream (lf_ctx)."""

    def __init__(self, device=0, stream=None):
        import weakref
        self._h = C.c_void_p()
        self._circuits = weakref.WeakSet()  # circuits uploaded on this context: freed before the context is
        check(_native.lib().lf_ctx_create(int(device), C.c_void_p(stream) if stream else None,
                                          C.byref(self._h)))

    def close(self):
        if self._h:
            for c in list(self._circuits):  # lf_circuit_free needs the live lf_ctx
                c.close()
            _native.lib().lf_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(_native.lib().lf_ctx_synchronize(self._h))

    @property
    def launch_count(self):
        return int(_native.lib().lf_ctx_launch_count(self._h))

    def elt_mul(self, field_id, a, b):
        a, b = _u8(a), _u8(b)
        out = np.empty_like(a)
        check(_native.lib().lf_elt_mul(self._h, field_id, _p(a), _p(b), _p(out),
                                       a.size // KBYTES[field_id]))
        return out

    def fft(self, field_id, elts, n, forward=False):
        """FFT<Field>::fftb / fftf (algebra/fft.h:185-201) on n elements (wire encoding);
        for FIELD_P256 the elements are n (re, im) pairs of Fp2."""
        elts = _u8(elts).copy()
        check(_native.lib().lf_fft(self._h, field_id, _p(elts), n, int(forward)))
        return elts

    def fft_time_ms(self, field_id, n, reps=10):
        ms = C.c_double()
        check(_native.lib().lf_fft_time(self._h, field_id, n, reps, C.byref(ms)))
        return ms.value

    def fft_time_rows_ms(self, field_id, n, nrows, reps=5):
        ms = C.c_double()
        check(_native.lib().lf_fft_time_rows(self._h, field_id, n, nrows, reps, C.byref(ms)))
        return ms.value

    def rs_time_ms(self, field_id, n, m, nrows, reps=5):
        """device-resident time of one ReedSolomon(n, m) extension of nrows rows (lf_rs_time)"""
        ms = C.c_double()
        check(_native.lib().lf_rs_time(self._h, field_id, n, m, nrows, reps, C.byref(ms)))
        return ms.value

    def microbench(self, what):
        g = C.c_double()
        check(_native.lib().lf_microbench(self._h, int(what), C.byref(g)))
        return g.value


def transcripts(n, seed=b"test"):
    """n caller-owned transcripts, each Transcript(seed) (lf_transcript_init)"""
    arr = (_native.Transcript * n)()
    for i in range(n):
        _native.lib().lf_transcript_init(C.byref(arr[i]), seed, len(seed))
    return arr


def transcript_write(ts, data):
    _native.lib().lf_transcript_write_bytes(C.byref(ts), data, len(data))


def transcript_challenge(ts, n):
    out = (C.c_uint8 * n)()
    _native.lib().lf_transcript_challenge_bytes(C.byref(ts), out, n)
    return bytes(out)


class _Interpolator:
    def __init__(self, ctx, field_id, n, m):
        self.ctx, self.field_id, self.n, self.m = ctx, field_id, n, m

    def interpolate(self, y):
        """y: (..., m, kBytes) uint8 with the first n of every row valid; returns
        the rows extended to m evaluations (the reference works in place)."""
        y = _u8(y).copy()
        kb = KBYTES[self.field_id]
        nrows = y.size // (self.m * kb)
        check(_native.lib().lf_rs_interpolate(self.ctx._h, self.field_id, self.n, self.m, _p(y), nrows))
        return y


class LCH14ReedSolomonFactory:
    """gf2k/lch14_reed_solomon.h:112-123 (GF(2^128))."""

    def __init__(self, ctx):
        self.ctx = ctx

    def make(self, n, m):
        return _Interpolator(self.ctx, FIELD_GF2_128, n, m)


class ReedSolomonFactory:
    """algebra/reed_solomon.h:132-147 (prime fields): make(n, m)->interpolate(y)."""

    def __init__(self, ctx, field_id):
        self.ctx, self.field_id = ctx, field_id

    def make(self, n, m):
        return _Interpolator(self.ctx, self.field_id, n, m)


class MerkleCommitment:
    """merkle/merkle_commitment.h:46-64, specialised to Ligero's column hash
    (ligero/ligero_param.h:432-439): leaf j = SHA256(nonce_j || column dblock+j)."""

    def __init__(self, ctx, field_id=FIELD_GF2_128):
        self.ctx, self.field_id = ctx, field_id

    def commit(self, tableau, nrow, block_enc, dblock, nonces, want_nodes=False):
        tableau, nonces = _u8(tableau), _u8(nonces)
        root = np.zeros(32, np.uint8)
        nodes = np.zeros((2 * (block_enc - dblock), 32), np.uint8) if want_nodes else None
        check(_native.lib().lf_merkle_commit(self.ctx._h, self.field_id, nrow, block_enc, dblock,
                                             _p(tableau), _p(nonces), _p(root),
                                             _p(nodes) if want_nodes else None))
        return (root.tobytes(), nodes) if want_nodes else root.tobytes()


class Circuit:
    """A circuit uploaded to the device (lf_circuit): flattened quads, sumcheck
    plans and Ligero parameters, shared by every proof of a batch."""

    def __init__(self, ctx, field_id, lfc1_bytes, rate=7, nreq=132, block_enc=0):
        self.ctx, self.field_id = ctx, field_id
        self._h = C.c_void_p()
        check(_native.lib().lf_circuit_upload(ctx._h, field_id, lfc1_bytes, len(lfc1_bytes), rate, nreq,
                                              block_enc, C.byref(self._h)))
        ctx._circuits.add(self)
        info = _native.CircuitInfo()
        check(_native.lib().lf_circuit_get_info(self._h, C.byref(info)))
        self.info = {n: int(getattr(info, n)) for n, _ in info._fields_}

    def verify_id(self):
        """CircuitReader's enforce_circuit_id: recompute circuit_id and compare with the stored one;
        raises LongfellowError(LF_ERR_FORMAT) on mismatch, returns the id."""
        out = np.zeros(32, np.uint8)
        check(_native.lib().lf_circuit_verify_id(self._h, _p(out)))
        return out.tobytes()

    def close(self):
        if self._h:
            _native.lib().lf_circuit_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZkProver:
    """zk/zk_prover.h:52-198: commit + prove + ZkProof::write for a batch of
    independent proofs of one circuit."""

    def __init__(self, circuit):
        self.c = circuit

    def prove_batch(self, witnesses, rng, tinit=b"test"):
        """witnesses: (B, ninputs*kBytes) uint8; rng: (B, >= rng_bytes) uint8.
        Returns (list of proof bytes, status array)."""
        info = self.c.info
        witnesses, rng = _u8(witnesses), _u8(rng)
        B = witnesses.shape[0]
        assert witnesses.shape[1] == info["witness_bytes"], witnesses.shape
        assert rng.shape[0] == B
        stride = info["max_proof_bytes"]
        out = np.zeros((B, stride), np.uint8)
        lens = np.zeros(B, np.uint64)
        status = np.zeros(B, np.int32)
        check(_native.lib().lf_zk_prove_batch(self.c._h, B, _p(witnesses), _p(rng), rng.shape[1], tinit,
                                              len(tinit), _p(out), stride, _p(lens), _p(status)))
        return [out[i, :int(lens[i])].tobytes() for i in range(B)], status

    def commit_batch(self, witnesses, rng, transcripts):
        """ZkProver::commit on caller-owned transcripts (lf_zk_commit_batch).  transcripts: a
        ctypes array of _native.Transcript, updated in place.  Returns (roots (B, 32), status)."""
        info = self.c.info
        witnesses, rng = _u8(witnesses), _u8(rng)
        B = witnesses.shape[0]
        assert witnesses.shape[1] == info["witness_bytes"] and rng.shape[0] == B and len(transcripts) == B
        roots = np.zeros((B, 32), np.uint8)
        status = np.zeros(B, np.int32)
        check(_native.lib().lf_zk_commit_batch(self.c._h, B, _p(witnesses), _p(rng), rng.shape[1], transcripts,
                                               _p(roots), _p(status)))
        return roots, status

    def prove_committed_batch(self, witnesses, transcripts):
        """ZkProver::prove + ZkProof::write for the batch committed last (lf_zk_prove_committed_batch)."""
        info = self.c.info
        witnesses = _u8(witnesses)
        B = witnesses.shape[0]
        stride = info["max_proof_bytes"]
        out = np.zeros((B, stride), np.uint8)
        lens = np.zeros(B, np.uint64)
        status = np.zeros(B, np.int32)
        check(_native.lib().lf_zk_prove_committed_batch(self.c._h, B, _p(witnesses), transcripts, _p(out), stride,
                                                        _p(lens), _p(status)))
        return [out[i, :int(lens[i])].tobytes() for i in range(B)], status

    STAGES = ["layout", "rs_encode", "merkle", "transcript_init", "eval_circuit", "sumcheck", "ligero_prove"]

    def prove_batch_ptr(self, nproofs, wit_ptr, rng_ptr, rng_stride, out_ptr, out_stride, lens_ptr,
                        status_ptr, tinit=b"test", device=False):
        """Raw-pointer form (pinned host buffers, or device buffers with device=True;
        the device form is asynchronous on the context stream)."""
        fn = _native.lib().lf_zk_prove_batch_dev if device else _native.lib().lf_zk_prove_batch
        check(fn(self.c._h, nproofs, C.c_void_p(wit_ptr), C.c_void_p(rng_ptr), rng_stride, tinit, len(tinit),
                 C.c_void_p(out_ptr), out_stride, C.c_void_p(lens_ptr), C.c_void_p(status_ptr)))

    def set_profiling(self, on=True):
        check(_native.lib().lf_circuit_set_profiling(self.c._h, int(on)))

    def stage_ms(self):
        ms = (C.c_float * 8)()
        n = _native.lib().lf_circuit_get_stage_ms(self.c._h, ms, 8)
        if n < 0:
            check(n)
        return dict(zip(self.STAGES, [float(x) for x in ms[:n]]))

    KERNEL_CLASSES = ["k_sc_eval", "k_sc_bind", "k_sc_round", "k_zk_sumcheck_tail"]

    def kernel_ms(self):
        """per kernel class of the flat sumcheck: (total ms, launches) of the most recent profiled batch"""
        out = {}
        for i, name in enumerate(self.KERNEL_CLASSES):
            ms, n = C.c_float(), C.c_size_t()
            check(_native.lib().lf_circuit_get_kernel_ms(self.c._h, i, C.byref(ms), C.byref(n)))
            out[name] = (float(ms.value), int(n.value))
        return out

    def rng_consumed(self, index):
        """bytes of proof `index`'s random stream the most recent batch consumed (redraws included)"""
        n = C.c_size_t()
        check(_native.lib().lf_zk_rng_consumed(self.c._h, index, C.byref(n)))
        return n.value

    def debug_fetch(self, index, stage, cap=1 << 26):
        buf = np.zeros(cap, np.uint8)
        n = C.c_size_t()
        check(_native.lib().lf_zk_debug_fetch(self.c._h, index, stage, _p(buf), cap, C.byref(n)))
        return buf[:n.value].copy()


class ZkVerifier:
    """zk/zk_verifier.h:39-111 + ZkProof::read (zk/zk_proof.h:107-112): recv_commitment + verify for a batch
    of independent proofs of one circuit (lf_zk_verify_batch)."""

    WHY = ["", "merkle_check", "low_degree_check", "dot_check", "wrong dot product", "quadratic_check"]

    def __init__(self, circuit):
        self.c = circuit

    def set_fault(self, fault):
        """test hook (lf_zk_verify_set_fault): break one of the verifier's interpolations"""
        check(_native.lib().lf_zk_verify_set_fault(self.c._h, int(fault)))

    def verify_batch(self, pub_inputs, proofs, tinit=b"test", transcripts=None):
        """pub_inputs: (B, npub_in*kBytes) uint8 (or None when the circuit has no public inputs); proofs: a list
        of B byte strings.  Returns (status, why) int32 arrays: status 0 accepted, LF_ERR_FORMAT (-3) not a
        proof of this shape, LF_ERR_VERIFY (-8) rejected with why = index into WHY.
        transcripts: an array of B caller-owned transcripts that have already received the commitments
        (recv_commitment); they continue and come back as ZkVerifier::verify leaves them
        (lf_zk_verify_committed_batch); tinit is then unused."""
        info = self.c.info
        B = len(proofs)
        stride = max(16, max(len(p) for p in proofs))
        buf = np.zeros((B, stride), np.uint8)
        lens = np.zeros(B, np.uint64)
        for i, pr in enumerate(proofs):
            buf[i, :len(pr)] = np.frombuffer(pr, np.uint8)
            lens[i] = len(pr)
        pubb = info["npub_in"] * info["kbytes"]
        pub = None
        if pubb:
            pub = _u8(pub_inputs).reshape(B, -1)
            assert pub.shape[1] == pubb, pub.shape
        status = np.zeros(B, np.int32)
        why = np.zeros(B, np.int32)
        if transcripts is not None:
            check(_native.lib().lf_zk_verify_committed_batch(self.c._h, B, _p(pub) if pub is not None else None,
                                                             _p(buf), stride, _p(lens), transcripts, _p(status),
                                                             _p(why)))
            return status, why
        check(_native.lib().lf_zk_verify_batch(self.c._h, B, _p(pub) if pub is not None else None, _p(buf), stride,
                                               _p(lens), tinit, len(tinit), _p(status), _p(why)))
        return status, why
