"""ctypes binding of oracle/_ref/libref.so -- TEST INFRASTRUCTURE ONLY.

libref.so is the UNMODIFIED reference (dwcoen1234/longfellow-zk, lib/) compiled
in place by oracle/ref_build/Makefile behind a small extern "C" wrapper.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (longfellow_zk_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBREF = os.path.join(_HERE, "_ref", "libref.so")
# the same wrapper + include/longfellow_b200_adapters.h, linked against the CUDA back end
LIBREF_GPU = os.path.join(_HERE, "_ref", "libref_gpu.so")

GF2_128_ID = 4  # proto/circuit_io.h:24-36
P256_ID = 1
FID_BN254 = 100  # ref_prime.cc local ids
FID_FP128 = 101
FID_GOLDILOCKS = 102

_lib = None


def available():
    return os.path.exists(LIBREF)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIBREF)
        _lib.ref_circuit_load.restype = C.c_void_p
        _lib.ref_circuit_load.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        _lib.ref_circuit_free.argtypes = [C.c_void_p]
        _lib.ref_zk_bench.restype = C.c_double
        _lib.ref_merkle_commit_open.restype = C.c_size_t
        _lib.ref_merkle_tree_len.restype = C.c_size_t
        _lib.ref_merkle_tree_len.argtypes = [C.c_size_t]
        _lib.ref_transcript_script.restype = C.c_size_t
        _lib.ref_circuit_info.restype = C.c_size_t
    return _lib


_gpu_lib = None


def gpu_adapters_available():
    return os.path.exists(LIBREF_GPU)


def gpu_lib():
    """libref_gpu.so: the reference driving the CUDA back end through the C++ adapters."""
    global _gpu_lib
    if _gpu_lib is None:
        _gpu_lib = C.CDLL(LIBREF_GPU)
        _gpu_lib.ref_circuit_load.restype = C.c_void_p
        _gpu_lib.ref_circuit_load.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        _gpu_lib.ref_circuit_free.argtypes = [C.c_void_p]
    return _gpu_lib


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gf128_mul(a, b):
    a, b = _u8(a).reshape(-1, 16), _u8(b).reshape(-1, 16)
    out = np.empty_like(a)
    lib().ref_gf128_mul(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def gf128_invert(a):
    a = _u8(a).reshape(-1, 16)
    out = np.empty_like(a)
    lib().ref_gf128_invert(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return out


def gf128_of_scalar(u):
    u = np.ascontiguousarray(u, dtype=np.uint64)
    out = np.empty((u.shape[0], 16), np.uint8)
    lib().ref_gf128_of_scalar(_p(u), _p(out), C.c_size_t(u.shape[0]))
    return out


def gf128_subfield_index(a):
    a = _u8(a).reshape(-1, 16)
    out = np.empty(a.shape[0], np.uint32)
    lib().ref_gf128_subfield_index(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return out


def gf128_constants():
    beta = np.empty((16, 16), np.uint8)
    pts = np.empty((6, 16), np.uint8)
    newton = np.empty((6, 6, 16), np.uint8)
    lib().ref_gf128_constants(_p(beta), _p(pts), _p(newton))
    return beta, pts, newton


def lch14(op, l, coset_or_k, B):
    """op: 'fft' | 'ifft' | 'bidir'. B: (2^l,16) uint8; returns a new array."""
    B = _u8(B).reshape(-1, 16).copy()
    assert B.shape[0] == 1 << l
    lib().ref_lch14(C.c_int({"fft": 0, "ifft": 1, "bidir": 2}[op]), C.c_size_t(l),
                    C.c_size_t(coset_or_k), _p(B))
    return B


def lch14_what():
    out = np.empty((16, 16, 16), np.uint8)
    lib().ref_lch14_what(_p(out))
    return out


def lch14_interpolate(n, m, rows):
    """rows: (nrows, m, 16) uint8 with the first n of each row valid."""
    rows = _u8(rows).copy()
    nrows = rows.shape[0]
    assert rows.shape == (nrows, m, 16)
    lib().ref_lch14_interpolate(C.c_size_t(n), C.c_size_t(m), _p(rows), C.c_size_t(nrows))
    return rows


def merkle_build(leaves):
    leaves = _u8(leaves).reshape(-1, 32)
    n = leaves.shape[0]
    nodes = np.zeros((2 * n, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    lib().ref_merkle_build(C.c_size_t(n), _p(leaves), _p(nodes), _p(root))
    return root, nodes


def merkle_commit_open(payload, rng, pos):
    """payload: (n, len) uint8; rng: n*32 bytes; pos: list of leaf indices."""
    payload = _u8(payload)
    n, ln = payload.shape
    rng = _u8(rng)
    assert rng.size >= 32 * n
    pos = np.ascontiguousarray(pos, dtype=np.uint64)
    root = np.zeros(32, np.uint8)
    nonce = np.zeros((max(len(pos), 1), 32), np.uint8)
    cap = max(len(pos), 1) * int(lib().ref_merkle_tree_len(n))
    path = np.zeros((cap, 32), np.uint8)
    k = lib().ref_merkle_commit_open(C.c_size_t(n), _p(payload), C.c_size_t(ln), _p(rng),
                                     _p(root), _p(pos), C.c_size_t(len(pos)), _p(nonce), _p(path))
    return root, nonce[:len(pos)], path[:k]


def transcript_script(init, script, out_cap=1 << 20):
    out = np.zeros(out_cap, np.uint8)
    n = lib().ref_transcript_script(C.c_char_p(init), C.c_size_t(len(init)), C.c_char_p(script),
                                    C.c_size_t(len(script)), _p(out), C.c_size_t(out_cap))
    return out[:n].tobytes()


LIGERO_FIELDS = ["block_enc", "block", "dblock", "block_ext", "r", "w", "nwrow", "nqtriples",
                 "nwqrow", "nrow", "mc_pathlen", "iq"]


def ligero_param(field_id, nw, nq, rate=7, nreq=132, block_enc=0):
    out = (C.c_size_t * 12)()
    lib().ref_ligero_param(C.c_int(field_id), C.c_size_t(nw), C.c_size_t(nq), C.c_size_t(rate),
                           C.c_size_t(nreq), C.c_size_t(block_enc), out)
    return dict(zip(LIGERO_FIELDS, [int(x) for x in out]))


def circuit_info(field_id, circ):
    out = (C.c_size_t * 4096)()
    k = lib().ref_circuit_info(C.c_int(field_id), C.c_char_p(circ), C.c_size_t(len(circ)), out,
                               C.c_size_t(4096))
    v = [int(x) for x in out[:k]]
    hdr = dict(zip(["nv", "logv", "nc", "logc", "nl", "ninputs", "npub_in", "subfield_boundary",
                    "nterms"], v[:9]))
    hdr["layers"] = [dict(nw=v[9 + 3 * i], logw=v[10 + 3 * i], nterms=v[11 + 3 * i])
                     for i in range(hdr["nl"])]
    return hdr


def _take(fn, *args):
    circ, wit = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)()
    cl, wl = C.c_size_t(), C.c_size_t()
    rc = fn(*args, C.byref(circ), C.byref(cl), C.byref(wit), C.byref(wl))
    assert rc == 0
    cb = C.string_at(circ, cl.value)
    wb = C.string_at(wit, wl.value)
    lib().ref_free(circ)
    lib().ref_free(wit)
    return cb, wb


def sha_circuit(nblocks=1):
    """(LFC1 circuit bytes, witness bytes) of BM_ShaZK_fp2_128/nblocks."""
    return _take(lib().ref_sha_circuit, C.c_size_t(nblocks))


def ecdsa_circuit(nsigs=1):
    """(LFC1 circuit bytes, witness bytes) of BM_ECDSAZKProver/nsigs."""
    return _take(lib().ref_ecdsa_circuit, C.c_size_t(nsigs))


def sha_witness(nblocks, msg):
    """witness of the nblocks-block SHA-256 circuit for `msg` (reference witness generator)"""
    import hashlib
    out = np.zeros(1 << 24, np.uint8)
    n = lib().ref_sha_witness(C.c_size_t(nblocks), C.c_char_p(msg), C.c_size_t(len(msg)),
                              C.c_char_p(hashlib.sha256(msg).digest()), _p(out), C.c_size_t(out.size))
    if n < 0:
        raise ValueError(f"ref_sha_witness failed: {n}")
    return out[:n * 16].tobytes()


def ecdsa_ntests():
    return int(lib().ref_ecdsa_ntests())


def ecdsa_witness(nsigs, first):
    """witness of the nsigs-signature ECDSA circuit from the reference's test vectors P256_TEST[first ...]"""
    out = np.zeros(1 << 22, np.uint8)
    n = lib().ref_ecdsa_witness(C.c_size_t(nsigs), C.c_size_t(first), _p(out), C.c_size_t(out.size))
    if n < 0:
        raise ValueError(f"ref_ecdsa_witness failed: {n}")
    return out[:n * 32].tobytes()


class Circuit:
    def __init__(self, field_id, circ_bytes):
        self.field_id = field_id
        self.h = lib().ref_circuit_load(C.c_int(field_id), C.c_char_p(circ_bytes),
                                        C.c_size_t(len(circ_bytes)))
        if not self.h:
            raise ValueError("reference CircuitReader rejected the circuit")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_circuit_free(C.c_void_p(self.h))
            self.h = None

    def prove(self, witness, rng, tinit=b"test", rate=7, nreq=132, block_enc=0, dump=False):
        """Reference ZkProver::commit+prove+ZkProof::write with replayed coins."""
        rng = _u8(np.frombuffer(rng, np.uint8) if isinstance(rng, (bytes, bytearray)) else rng)
        out = np.zeros(1 << 21, np.uint8)
        out_len, used = C.c_size_t(), C.c_size_t()
        dw = np.zeros(1 << 22, np.uint8) if dump else None
        dt = np.zeros(1 << 26, np.uint8) if dump else None
        dr = np.zeros(32, np.uint8) if dump else None
        ds = np.zeros(1 << 20, np.uint8) if dump else None
        n = lambda a: C.c_size_t(a.size if a is not None else 0)
        pp = lambda a: _p(a) if a is not None else None
        rc = lib().ref_zk_prove(C.c_void_p(self.h), C.c_char_p(witness), _p(rng),
                                C.c_size_t(rng.size), C.c_char_p(tinit), C.c_size_t(len(tinit)),
                                C.c_size_t(rate), C.c_size_t(nreq), C.c_size_t(block_enc), _p(out),
                                C.c_size_t(out.size), C.byref(out_len), C.byref(used),
                                pp(dw), n(dw), pp(dt), n(dt), pp(dr), pp(ds), n(ds))
        if rc != 0:
            raise RuntimeError(f"reference prover failed rc={rc}")
        res = dict(proof=out[:out_len.value].tobytes(), rng_used=used.value)
        if dump:
            res.update(witness=dw, tableau=dt, root=dr.tobytes(), sumcheck=ds)
        return res

    def verify(self, pub, proof, tinit=b"test", rate=7, nreq=132, block_enc=0):
        return int(lib().ref_zk_verify(C.c_void_p(self.h), C.c_char_p(pub), C.c_char_p(tinit),
                                       C.c_size_t(len(tinit)), C.c_size_t(rate), C.c_size_t(nreq),
                                       C.c_size_t(block_enc), C.c_char_p(proof),
                                       C.c_size_t(len(proof))))

    def bench(self, witness, rng, nthreads=1, per_thread=4, rate=7, nreq=132):
        rng = _u8(np.frombuffer(rng, np.uint8) if isinstance(rng, (bytes, bytearray)) else rng)
        lat = (C.c_double * per_thread)()
        secs = lib().ref_zk_bench(C.c_void_p(self.h), C.c_char_p(witness), _p(rng),
                                  C.c_size_t(rng.size), C.c_size_t(rate), C.c_size_t(nreq),
                                  C.c_size_t(nthreads), C.c_size_t(per_thread), lat)
        return float(secs), [float(x) for x in lat]


def prove_pair(circ_a, circ_b, wit_a, wit_b, rng, tinit=b"test", rate=7, nreq=132):
    """commit(A), commit(B), 16 challenge bytes, prove(A), prove(B) on ONE reference Transcript
    (the structure of run_mdoc_prover).  circ_a: GF(2^128) Circuit, circ_b: Fp256 Circuit."""
    rng = _u8(np.frombuffer(rng, np.uint8) if isinstance(rng, (bytes, bytearray)) else rng)
    oa, ob = np.zeros(1 << 21, np.uint8), np.zeros(1 << 21, np.uint8)
    la, lb, ua, ut = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    ch = np.zeros(16, np.uint8)
    rc = lib().ref_zk_prove_pair(C.c_void_p(circ_a.h), C.c_void_p(circ_b.h), C.c_char_p(wit_a), C.c_char_p(wit_b),
                                 _p(rng), C.c_size_t(rng.size), C.c_char_p(tinit), C.c_size_t(len(tinit)),
                                 C.c_size_t(rate), C.c_size_t(nreq), _p(ch), _p(oa), C.c_size_t(oa.size),
                                 C.byref(la), _p(ob), C.c_size_t(ob.size), C.byref(lb), C.byref(ua), C.byref(ut))
    if rc != 0:
        raise RuntimeError(f"reference pair prover failed rc={rc}")
    return dict(proof_a=oa[:la.value].tobytes(), proof_b=ob[:lb.value].tobytes(), challenge=ch.tobytes(),
                rng_used_a=ua.value, rng_used_total=ut.value)


# ---- the reference's mdoc prover split at its ZkProver calls (oracle/_ref/libref_mdoc.so) ----
LIBREF_MDOC = os.path.join(_HERE, "_ref", "libref_mdoc.so")
_mdoc_lib = None


def mdoc_available():
    return os.path.exists(LIBREF_MDOC)


def zstd_decompress(data):
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    cap = z.ZSTD_getFrameContentSize(data, len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_decompress(buf, cap, data, len(data))
    assert not z.ZSTD_isError(n)
    return buf.raw[:n]


def zstd_compress(data, level=3):
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    cap = z.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), level)
    assert not z.ZSTD_isError(n)
    return buf.raw[:n]


# ---- run_mdoc_prover compiled unchanged against the CUDA back end (oracle/_ref/libref_mdoc_gpu.so) ----
LIBREF_MDOC_GPU = os.path.join(_HERE, "_ref", "libref_mdoc_gpu.so")
_mdoc_gpu_lib = None


def mdoc_gpu_available():
    return os.path.exists(LIBREF_MDOC_GPU)


def mdoc_gpu_lib():
    global _mdoc_gpu_lib
    if _mdoc_gpu_lib is None:
        _mdoc_gpu_lib = _mdoc_gpu_protos(C.CDLL(LIBREF_MDOC_GPU))
        assert _mdoc_gpu_lib.ref_mdoc_gpu_verifier_is_gpu() == 0
    return _mdoc_gpu_lib


def _mdoc_gpu_protos(L):
    L.ref_mdoc_gpu_nclaims.restype = C.c_size_t
    L.ref_mdoc_gpu_claim_name.restype = C.c_char_p
    L.ref_mdoc_gpu_claim_name.argtypes = [C.c_size_t]
    L.ref_mdoc_gpu_run_claim.argtypes = [C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
    L.ref_mdoc_gpu_prove_claim.argtypes = [C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t)]
    L.ref_mdoc_gpu_verify_claim.argtypes = [C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    return L


# the same with run_mdoc_verifier on the GPU too (ZkVerifier -> ZkVerifierGpu): oracle/_ref/libref_mdoc_gpuv.so
LIBREF_MDOC_GPUV = os.path.join(_HERE, "_ref", "libref_mdoc_gpuv.so")
_mdoc_gpuv_lib = None


def mdoc_gpuv_available():
    return os.path.exists(LIBREF_MDOC_GPUV)


def mdoc_gpuv_lib():
    global _mdoc_gpuv_lib
    if _mdoc_gpuv_lib is None:
        _mdoc_gpuv_lib = _mdoc_gpu_protos(C.CDLL(LIBREF_MDOC_GPUV))
        assert _mdoc_gpuv_lib.ref_mdoc_gpu_verifier_is_gpu() == 1
    return _mdoc_gpuv_lib


def mdoc_prove_claim(L, i, circuit_zstd, cap=1 << 20):
    """run_mdoc_prover of library L (either of the two above) for claim i: (code, proof bytes)"""
    out = C.create_string_buffer(cap)
    n = C.c_size_t()
    rc = L.ref_mdoc_gpu_prove_claim(i, circuit_zstd, len(circuit_zstd), out, cap, C.byref(n))
    return int(rc), out.raw[:n.value] if rc == 0 else b""


def mdoc_verify_claim(L, i, circuit_zstd, proof):
    """run_mdoc_verifier of library L for claim i on the given proof: the MdocVerifierErrorCode (0 = accepted)"""
    return int(L.ref_mdoc_gpu_verify_claim(i, circuit_zstd, len(circuit_zstd), proof, len(proof)))


def mdoc_gpu_run_claim(i, circuit_zstd, tamper=False):
    """MdocZKTest::run_test for claim i of mdoc_zk_test.cc:119-170: the reference's run_mdoc_prover on the GPU
    back end, then the reference's run_mdoc_verifier.  Returns (code, proof_len); code 0 = proved and accepted."""
    n = C.c_size_t()
    rc = mdoc_gpu_lib().ref_mdoc_gpu_run_claim(i, circuit_zstd, len(circuit_zstd), C.byref(n), int(tamper))
    return int(rc), int(n.value)


class MdocCase:
    """mdoc_tests[0] + age_over_18 on kZkSpecs[0] (the reference's benchmark claim,
    circuits/mdoc/mdoc_zk_test.cc:652-656): filled witnesses, and run_mdoc_prover from
    "Run prover" on with replayed commit coins."""

    def __init__(self, circuit_raw):
        global _mdoc_lib
        if _mdoc_lib is None:
            _mdoc_lib = C.CDLL(LIBREF_MDOC)
            _mdoc_lib.ref_mdoc_prepare.restype = C.c_void_p
        self.lib = _mdoc_lib
        self.h = self.lib.ref_mdoc_prepare(C.c_char_p(circuit_raw), C.c_size_t(len(circuit_raw)))
        if not self.h:
            raise RuntimeError("ref_mdoc_prepare failed")
        out = (C.c_size_t * 8)()
        self.lib.ref_mdoc_info(C.c_void_p(self.h), out)
        (self.sig_ninputs, self.sig_npub, self.hash_ninputs, self.hash_npub, self.block_enc_sig,
         self.block_enc_hash, tr_len, self.version) = [int(x) for x in out]
        tr = np.zeros(tr_len, np.uint8)
        self.lib.ref_mdoc_transcript(C.c_void_p(self.h), _p(tr))
        self.transcript = tr.tobytes()
        rate, nreq = C.c_size_t(), C.c_size_t()
        self.lib.ref_mdoc_ligero_params(C.byref(rate), C.byref(nreq))
        self.rate, self.nreq = rate.value, nreq.value

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_mdoc_free(C.c_void_p(self.h))
            self.h = None

    def witnesses(self):
        ws, wh = np.zeros(self.sig_ninputs * 32, np.uint8), np.zeros(self.hash_ninputs * 16, np.uint8)
        self.lib.ref_mdoc_witness(C.c_void_p(self.h), _p(ws), _p(wh))
        return ws, wh

    def update_macs(self, av):
        ws, wh = np.zeros(self.sig_ninputs * 32, np.uint8), np.zeros(self.hash_ninputs * 16, np.uint8)
        macs = np.zeros(96, np.uint8)
        self.lib.ref_mdoc_update_macs(C.c_void_p(self.h), C.c_char_p(av), _p(ws), _p(wh), _p(macs))
        return ws, wh, macs.tobytes()

    def prove(self, coins):
        coins = _u8(coins)
        out = np.zeros(1 << 22, np.uint8)
        n, lh, ls, ch, ct = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        av = np.zeros(16, np.uint8)
        rc = self.lib.ref_mdoc_prove(C.c_void_p(self.h), _p(coins), C.c_size_t(coins.size), _p(out),
                                     C.c_size_t(out.size), C.byref(n), C.byref(lh), C.byref(ls), C.byref(ch),
                                     C.byref(ct), _p(av))
        if rc != 0:
            raise RuntimeError(f"reference mdoc prover failed rc={rc}")
        return dict(proof=out[:n.value].tobytes(), len_hash=lh.value, len_sig=ls.value, coins_hash=ch.value,
                    coins_total=ct.value, av=av.tobytes())


class GpuAdapterCircuit:
    """A reference Circuit object inside libref_gpu.so, proved (1) by the reference's own
    ZkProver with GpuReedSolomonFactory injected, (2) by GpuZkProver."""

    def __init__(self, field_id, circ_bytes):
        self.h = gpu_lib().ref_circuit_load(C.c_int(field_id), C.c_char_p(circ_bytes), C.c_size_t(len(circ_bytes)))
        if not self.h:
            raise ValueError("reference CircuitReader rejected the circuit")

    def __del__(self):
        if getattr(self, "h", None):
            gpu_lib().ref_circuit_free(C.c_void_p(self.h))
            self.h = None

    def _call(self, fn, witness, rng, tinit, rate, nreq, *extra):
        rng = _u8(np.frombuffer(rng, np.uint8) if isinstance(rng, (bytes, bytearray)) else rng)
        out = np.zeros(1 << 21, np.uint8)
        out_len = C.c_size_t()
        rc = fn(C.c_void_p(self.h), C.c_char_p(witness), _p(rng), C.c_size_t(rng.size), C.c_char_p(tinit),
                C.c_size_t(len(tinit)), C.c_size_t(rate), C.c_size_t(nreq), *extra, _p(out), C.c_size_t(out.size),
                C.byref(out_len))
        if rc != 0:
            raise RuntimeError(f"adapter prover failed rc={rc}")
        return out[:out_len.value].tobytes()

    def prove_reference_with_gpu_rs(self, witness, rng, tinit=b"test", rate=7, nreq=132):
        return self._call(gpu_lib().ref_zk_prove_gpu_rs, witness, rng, tinit, rate, nreq)

    def prove_gpu(self, witness, rng, tinit=b"test", rate=7, nreq=132, copies=1):
        return self._call(gpu_lib().ref_zk_prove_gpu, witness, rng, tinit, rate, nreq, C.c_size_t(copies))


def fft(fid, data, n, fwd=False):
    data = _u8(data).copy()
    assert lib().ref_fft(C.c_int(fid), _p(data), C.c_size_t(n), C.c_int(int(fwd))) == 0
    return data


def fft_p256_2(data, n, fwd=False):
    data = _u8(data).copy()
    assert lib().ref_fft_p256_2(_p(data), C.c_size_t(n), C.c_int(int(fwd))) == 0
    return data


def rs(fid, rows, n, m):
    rows = _u8(rows).copy()
    nrows = rows.shape[0]
    assert lib().ref_rs(C.c_int(fid), _p(rows), C.c_size_t(n), C.c_size_t(m), C.c_size_t(nrows)) == 0
    return rows


def fft_bench(fid, n, reps=3):
    """seconds per FFT<Field>::fftb(n) on one host thread"""
    f = lib().ref_fft_bench
    f.restype = C.c_double
    return float(f(C.c_int(fid), C.c_size_t(n), C.c_int(reps)))


def rs_bench(fid, n, m, reps=3):
    """seconds per ReedSolomon(n, m)::interpolate (fid 4: LCH14ReedSolomon) of one row on one host thread"""
    if fid == 4:
        f = lib().ref_lch14_rs_bench
        f.restype = C.c_double
        return float(f(C.c_size_t(n), C.c_size_t(m), C.c_int(reps)))
    f = lib().ref_rs_bench
    f.restype = C.c_double
    return float(f(C.c_int(fid), C.c_size_t(n), C.c_size_t(m), C.c_int(reps)))


def fp_mul(fid, a, b):
    a, b = _u8(a), _u8(b)
    out = np.empty_like(a)
    kb = {FID_BN254: 32, FID_FP128: 16, FID_GOLDILOCKS: 8, P256_ID: 32}[fid]
    assert lib().ref_fp_mul(C.c_int(fid), _p(a), _p(b), _p(out), C.c_size_t(a.size // kb)) == 0
    return out
