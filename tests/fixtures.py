"""Shared loaders for tests/golden fixtures (no reference needed at run time)."""
import json
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    circ = zlib.decompress(open(os.path.join(GOLDEN, name + ".circuit.z"), "rb").read())
    wit = zlib.decompress(open(os.path.join(GOLDEN, name + ".witness.z"), "rb").read())
    return circ, wit


def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def rng_bytes(seed, n):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def load_witnesses(name):
    """several distinct satisfying witnesses of a 1-instance circuit (make_golden.py distinct_witnesses):
    an (N, witness_bytes) uint8 array"""
    _, wit = load(name)
    raw = zlib.decompress(open(os.path.join(GOLDEN, name + ".witnesses.z"), "rb").read())
    return np.frombuffer(raw, np.uint8).reshape(-1, len(wit)).copy()


def load_size(name):
    """(circuit, witness, golden record) of one of the other published instance sizes (tests/golden/sizes)"""
    import lzma
    d = os.path.join(GOLDEN, "sizes")
    circ = lzma.decompress(open(os.path.join(d, name + ".circuit.xz"), "rb").read())
    wit = zlib.decompress(open(os.path.join(d, name + ".witness.z"), "rb").read())
    return circ, wit, json.load(open(os.path.join(d, "golden_sizes.json")))[name]


SIZE_NAMES = ["sha2_gf128", "sha4_gf128", "sha8_gf128", "sha16_gf128", "sha32_gf128", "sha33_gf128", "ecdsa2_p256",
              "ecdsa3_p256"]


def load_mdoc():
    """The frozen mdoc instance of tests/golden/make_golden_mdoc.py: the two circuits of kZkSpecs[0]
    (signature circuit then hash circuit, LFC1 bytes as the reference's circuit file holds them,
    re-compressed with xz), filled witnesses before / after the MAC patch, commit coins, and what
    run_mdoc_prover produced for them."""
    import json
    import lzma
    import zlib
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mdoc")
    rd = lambda n: np.frombuffer(zlib.decompress(open(os.path.join(d, n + ".bin.z"), "rb").read()), np.uint8).copy()
    return dict(raw=lzma.decompress(open(os.path.join(d, "circuits_v7_1attr.lfc1.xz"), "rb").read()),
                w_sig=rd("w_sig"), w_hash=rd("w_hash"), w_sig_mac=rd("w_sig_mac"), w_hash_mac=rd("w_hash_mac"),
                coins=np.fromfile(os.path.join(d, "coins.bin"), np.uint8),
                expect=json.load(open(os.path.join(d, "expect.json"))))


def load_rfc_vector(oracle):
    """The reference's own known-answer ZK test (tests/golden/make_golden_rfc.py): returns
    (record, circuit bytes, witness bytes, coin stream, expected proof bytes).  `oracle` supplies the
    GF(2^128) helpers (subfield embedding, multiply, LigeroParam) needed to rebuild the inputs the test
    describes; the expected bytes are the reference's."""
    rec = json.load(open(os.path.join(GOLDEN, "rfc_zk_vector1.json")))
    gf = rec["field_id"]
    circ = bytes.fromhex(rec["circuit"])
    one = np.zeros((1, 16), np.uint8)
    one[0, 0] = 1
    x = np.zeros((1, 16), np.uint8)
    x[0, 0] = 2
    n, m = (oracle.gf128_of_scalar([s]) for s in rec["subfield_scalars"])
    w3 = oracle.gf128_mul(oracle.elt_op(gf, "add", n, m), x)
    wit = np.concatenate([one, n, m, w3]).tobytes()
    # header of the circuit (proto/circuit_reader.h:83-120): 3-byte little-endian numbers after the version byte
    num = lambda i: int.from_bytes(circ[1 + 3 * i:4 + 3 * i], "little")
    npub, sb_abs, ninputs, nl = num(3), num(4), num(5), num(6)
    assert nl == 1
    logw = int.from_bytes(circ[1 + 3 * 8 + 16 * num(7):][:3], "little")
    npad = 4 * logw + 2                      # fill_pad: zk_prover.h:152-188
    nw, nq = (ninputs - npub) + npad + 1, nl
    P = oracle.ligero_param(gf, nw, nq, rec["rate"], rec["nreq"], rec["block_enc"])
    sb = max(sb_abs - npub, 0)
    # every RandomEngine::bytes(n) call of the test returns 02 00 ... 00: 16 bytes per field element,
    # 2 per subfield element (random/random.h:36-50), 32 per Merkle nonce (merkle_commitment.h:48-52)
    e16, e2, n32 = b"\x02" + bytes(15), b"\x02\x00", b"\x02" + bytes(31)
    st = e16 * npad + e16 * P["block"] + e16 * P["dblock"] * 2
    for i in range(P["nwrow"]):
        st += (e2 if (i + 1) * P["w"] <= sb else e16) * P["r"]
    st += e16 * 3 * P["nqtriples"] * P["r"]
    st += n32 * (P["block_ext"])
    want = bytes.fromhex(rec["commitment"] + rec["sumcheck_proof"] + rec["ligero_proof"])
    return rec, circ, wit, np.frombuffer(st, np.uint8).copy(), want
