"""ctypes binding of oracle/liboracle.so, the plain-C restatement of the
reference's prover hot path (oracle/port/*.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module; the product (longfellow_zk_b200/) never does.  The interface
mirrors oracle/refapi.py so that the same test can run against either.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBORACLE = os.path.join(_HERE, "liboracle.so")

P256_ID = 1
GF2_128_ID = 4
FID_BN254 = 100
FID_FP128 = 101
FID_GOLDILOCKS = 102
FID_SECP256K1 = 10
KBYTES = {P256_ID: 32, GF2_128_ID: 16, 10: 32, FID_BN254: 32, FID_FP128: 16, FID_GOLDILOCKS: 8}

_lib = None


def build(force=False):
    """Compile oracle/port/*.c -> oracle/liboracle.so (gcc, a second or two)."""
    if force or not os.path.exists(LIBORACLE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "port")])
    return LIBORACLE


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIBORACLE)
        _lib.orc_circuit_load.restype = C.c_void_p
        _lib.orc_circuit_load.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        _lib.orc_circuit_free.argtypes = [C.c_void_p]
        _lib.orc_circuit_id.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_merkle_commit_open.restype = C.c_size_t
        _lib.orc_merkle_tree_len.restype = C.c_size_t
        _lib.orc_merkle_tree_len.argtypes = [C.c_size_t]
        _lib.orc_transcript_script.restype = C.c_size_t
    return _lib


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def elt_op(fid, op, a, b=None):
    """op in add|sub|mul|inv on arrays of elements in wire encoding."""
    a = _u8(a)
    out = np.empty_like(a)
    n = a.size // KBYTES[fid]
    bb = _p(_u8(b)) if b is not None else None
    rc = lib().orc_elt_op(C.c_int(fid), C.c_int({"add": 0, "sub": 1, "mul": 2, "inv": 3}[op]), _p(a),
                          bb, _p(out), C.c_size_t(n))
    assert rc == 0, rc
    return out


def gf128_mul(a, b):
    return elt_op(GF2_128_ID, "mul", _u8(a).reshape(-1, 16), _u8(b).reshape(-1, 16))


def gf128_invert(a):
    return elt_op(GF2_128_ID, "inv", _u8(a).reshape(-1, 16))


def gf128_of_scalar(u):
    u = np.ascontiguousarray(u, dtype=np.uint64)
    out = np.empty((u.shape[0], 16), np.uint8)
    lib().orc_gf128_of_scalar(_p(u), _p(out), C.c_size_t(u.shape[0]))
    return out


def gf128_subfield_index(a):
    a = _u8(a).reshape(-1, 16)
    out = np.empty(a.shape[0], np.uint32)
    lib().orc_gf128_subfield_index(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return out


def gf128_constants():
    beta = np.empty((16, 16), np.uint8)
    pts = np.empty((6, 16), np.uint8)
    newton = np.empty((6, 6, 16), np.uint8)
    lib().orc_gf128_constants(_p(beta), _p(pts), _p(newton))
    return beta, pts, newton


def lch14(op, l, coset_or_k, B):
    B = _u8(B).reshape(-1, 16).copy()
    assert B.shape[0] == 1 << l
    lib().orc_lch14(C.c_int({"fft": 0, "ifft": 1, "bidir": 2}[op]), C.c_size_t(l),
                    C.c_size_t(coset_or_k), _p(B))
    return B


def lch14_what():
    out = np.empty((16, 16, 16), np.uint8)
    lib().orc_lch14_what(_p(out))
    return out


def rs_interpolate(fid, n, m, rows):
    rows = _u8(rows).copy()
    nrows = rows.shape[0]
    rc = lib().orc_rs_interpolate(C.c_int(fid), C.c_size_t(n), C.c_size_t(m), _p(rows),
                                  C.c_size_t(nrows))
    assert rc == 0, rc
    return rows


def lch14_interpolate(n, m, rows):
    return rs_interpolate(GF2_128_ID, n, m, rows)


def fft(fid, data, n, fwd=False):
    data = _u8(data).copy()
    assert lib().orc_fft(C.c_int(fid), _p(data), C.c_size_t(n), C.c_int(int(fwd))) == 0
    return data


def sha256(data):
    out = np.zeros(32, np.uint8)
    lib().orc_sha256(C.c_char_p(data), C.c_size_t(len(data)), _p(out))
    return out.tobytes()


def aes256_ecb(key, data):
    out = np.zeros(len(data), np.uint8)
    lib().orc_aes256_ecb(C.c_char_p(key), C.c_char_p(data), _p(out), C.c_size_t(len(data) // 16))
    return out.tobytes()


def merkle_build(leaves):
    leaves = _u8(leaves).reshape(-1, 32)
    n = leaves.shape[0]
    nodes = np.zeros((2 * n, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    lib().orc_merkle_build(C.c_size_t(n), _p(leaves), _p(nodes), _p(root))
    return root, nodes


def merkle_commit_open(payload, rng, pos):
    payload = _u8(payload)
    n, ln = payload.shape
    rng = _u8(rng)
    pos = np.ascontiguousarray(pos, dtype=np.uint64)
    root = np.zeros(32, np.uint8)
    nonce = np.zeros((max(len(pos), 1), 32), np.uint8)
    cap = max(len(pos), 1) * int(lib().orc_merkle_tree_len(n))
    path = np.zeros((cap, 32), np.uint8)
    k = lib().orc_merkle_commit_open(C.c_size_t(n), _p(payload), C.c_size_t(ln), _p(rng), _p(root),
                                     _p(pos), C.c_size_t(len(pos)), _p(nonce), _p(path))
    return root, nonce[:len(pos)], path[:k]


def transcript_script(init, script, fid=GF2_128_ID, out_cap=1 << 20):
    out = np.zeros(out_cap, np.uint8)
    n = lib().orc_transcript_script(C.c_int(fid), C.c_char_p(init), C.c_size_t(len(init)),
                                    C.c_char_p(script), C.c_size_t(len(script)), _p(out),
                                    C.c_size_t(out_cap))
    return out[:n].tobytes()


LIGERO_FIELDS = ["block_enc", "block", "dblock", "block_ext", "r", "w", "nwrow", "nqtriples",
                 "nwqrow", "nrow", "mc_pathlen", "iq"]


def ligero_param(field_id, nw, nq, rate=7, nreq=132, block_enc=0):
    out = (C.c_size_t * 12)()
    rc = lib().orc_ligero_param(C.c_int(field_id), C.c_size_t(nw), C.c_size_t(nq), C.c_size_t(rate),
                                C.c_size_t(nreq), C.c_size_t(block_enc), out)
    assert rc == 0
    return dict(zip(LIGERO_FIELDS, [int(x) for x in out]))


def ligero_prove(field_id, W, lqc, terms_c, terms_w, terms_k, ncons, hash32, rng, subfield_boundary=0,
                 tinit=b"test", rate=7, nreq=132, block_enc=0):
    """LigeroProver::commit + ::prove on a caller-given statement (orc_ligero_prove): W (nw, kB) and terms_k
    (nterms, kB) uint8 in wire encoding, lqc (nq, 3) and terms_c / terms_w index arrays.  Returns
    (root, LigeroProof bytes, rng bytes consumed)."""
    W, K = _u8(W), _u8(terms_k)
    kb = KBYTES[field_id]
    nw, nt = W.size // kb, K.size // kb
    q = np.ascontiguousarray(lqc, dtype=np.uint64).reshape(-1)
    tc = np.ascontiguousarray(terms_c, dtype=np.uint64)
    tw = np.ascontiguousarray(terms_w, dtype=np.uint64)
    rng = _u8(np.frombuffer(rng, np.uint8) if isinstance(rng, (bytes, bytearray)) else rng)
    out = np.zeros(1 << 22, np.uint8)
    out_len, used = C.c_size_t(), C.c_size_t()
    rc = lib().orc_ligero_prove(C.c_int(field_id), C.c_size_t(nw), C.c_size_t(q.size // 3), C.c_size_t(ncons),
                                C.c_size_t(nt), C.c_size_t(subfield_boundary), _p(W), _p(q), _p(tc), _p(tw), _p(K),
                                C.c_char_p(hash32), _p(rng), C.c_size_t(rng.size), C.c_char_p(tinit),
                                C.c_size_t(len(tinit)), C.c_size_t(rate), C.c_size_t(nreq), C.c_size_t(block_enc),
                                _p(out), C.c_size_t(out.size), C.byref(out_len), C.byref(used))
    if rc != 0:
        raise RuntimeError(f"oracle Ligero prover failed rc={rc}")
    b = out[:out_len.value].tobytes()
    return b[:32], b[32:], used.value


class Circuit:
    def __init__(self, field_id, circ_bytes):
        self.field_id = field_id
        self.h = lib().orc_circuit_load(C.c_int(field_id), C.c_char_p(circ_bytes),
                                        C.c_size_t(len(circ_bytes)))
        if not self.h:
            raise ValueError("oracle circuit parser rejected the circuit")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_circuit_free(C.c_void_p(self.h))
            self.h = None

    def id(self):
        out = np.zeros(32, np.uint8)
        lib().orc_circuit_id(C.c_void_p(self.h), _p(out))
        return out.tobytes()

    def prove(self, witness, rng, tinit=b"test", rate=7, nreq=132, block_enc=0, dump=False):
        rng = _u8(np.frombuffer(rng, np.uint8) if isinstance(rng, (bytes, bytearray)) else rng)
        out = np.zeros(1 << 21, np.uint8)
        out_len, used = C.c_size_t(), C.c_size_t()
        dw = np.zeros(1 << 22, np.uint8) if dump else None
        dt = np.zeros(1 << 26, np.uint8) if dump else None
        dr = np.zeros(32, np.uint8) if dump else None
        ds = np.zeros(1 << 20, np.uint8) if dump else None
        n = lambda a: C.c_size_t(a.size if a is not None else 0)
        pp = lambda a: _p(a) if a is not None else None
        rc = lib().orc_zk_prove(C.c_void_p(self.h), C.c_char_p(witness), _p(rng),
                                C.c_size_t(rng.size), C.c_char_p(tinit), C.c_size_t(len(tinit)),
                                C.c_size_t(rate), C.c_size_t(nreq), C.c_size_t(block_enc), _p(out),
                                C.c_size_t(out.size), C.byref(out_len), C.byref(used),
                                pp(dw), n(dw), pp(dt), n(dt), pp(dr), pp(ds), n(ds))
        if rc != 0:
            raise RuntimeError(f"oracle prover failed rc={rc}")
        res = dict(proof=out[:out_len.value].tobytes(), rng_used=used.value)
        if dump:
            res.update(witness=dw, tableau=dt, root=dr.tobytes(), sumcheck=ds)
        return res
