"""The oracle (oracle/port, plain C) against every golden vector the reference
holds for this path:
  * docs/specs/testvectors.md:7-21 + lib/merkle/merkle_tree_test.cc:186-236  (Merkle)
  * docs/specs/testvectors.md:23-98 + lib/random/transcript_test.cc:131-341   (Fiat-Shamir)
  * rust/runtime/merkle/tests/{merkle,commitment}_test_vector.bin (C++-generated;
    layouts per rust/runtime/merkle/tests/merkle.rs:219-299,302-417)
  * rust/runtime/random/tests/transcript_test_vector.bin (C++-generated; transcript.rs:18-67)
  * rust/runtime/ligero/tests/ligero_test_vector.bin (C++-generated; ligero.rs:592-760)
  * tests/golden/golden.json: proofs produced by the unmodified reference
  * tests/golden/rfc_zk_vector1.json: the reference's known-answer test of the whole ZK prover
    (rust/runtime/zk/tests/zk.rs:228-558, bytes produced by the C++ prover)
CPU only."""
import hashlib
import struct

import numpy as np

from fixtures import GOLDEN, golden, load, rng_bytes

H = bytes.fromhex

MERKLE_LEAVES = [
    "4bf5122f344554c53bde2ebb8cd2b7e3d1600ad631c385a5d7cce23c7785459a",
    "dbc1b4c900ffe48d575b5da5c638040125f65db0fe3e24494b76ea986457d986",
    "084fed08b978af4d7d196a7446a86b58009e636b611db16211b65a9aadff29c5",
    "e52d9c508c502347344d8c07ad91cbd6068afc75ff6292f062a09ca381c89e71",
    "e77b9a9ae9e30b0dbdb6f510a264ef9de781501d7b6b92ae89eb059c5ab743db",
]
MERKLE_ROOT = "f22f4501ffd3bdffcecc9e4cd6828a4479aeedd6aa484eb7c1f808ccf71c6e76"


def compressed_proof(nodes, n, pos):
    """merkle_tree.h:75-98,122-143 on a heap of digests (python, tiny cases)"""
    tree = [False] * (2 * n)
    for p in pos:
        tree[p + n] = True
    for i in range(n - 1, 0, -1):
        tree[i] = tree[2 * i] or tree[2 * i + 1]
    out = []
    for i in range(n - 1, 0, -1):
        if tree[i]:
            c = 2 * i
            if tree[c]:
                c = 2 * i + 1
            if not tree[c]:
                out.append(nodes[c].tobytes())
    return out


def test_merkle_spec_vector(oracle):
    leaves = np.frombuffer(b"".join(H(x) for x in MERKLE_LEAVES), np.uint8).reshape(5, 32)
    root, nodes = oracle.merkle_build(leaves)
    assert root.tobytes().hex() == MERKLE_ROOT
    assert [x.hex() for x in compressed_proof(nodes, 5, [0, 1])] == [
        "084fed08b978af4d7d196a7446a86b58009e636b611db16211b65a9aadff29c5",
        "f03808f5b8088c61286d505e8e93aa378991d9889ae2d874433ca06acabcd493"]
    assert [x.hex() for x in compressed_proof(nodes, 5, [1, 3])] == [
        "e77b9a9ae9e30b0dbdb6f510a264ef9de781501d7b6b92ae89eb059c5ab743db",
        "084fed08b978af4d7d196a7446a86b58009e636b611db16211b65a9aadff29c5",
        "4bf5122f344554c53bde2ebb8cd2b7e3d1600ad631c385a5d7cce23c7785459a"]


def test_merkle_open_is_the_compressed_proof(oracle):
    rs = np.random.default_rng(3)
    for n, pos in [(5, [1, 3]), (5, [0, 1]), (7, [6]), (8, [0, 7]), (100, [3, 50, 51, 99]), (1, [0])]:
        payload = rs.integers(0, 256, (n, 9), dtype=np.uint8)
        rng = rs.integers(0, 256, n * 32, dtype=np.uint8)
        root, nonce, path = oracle.merkle_commit_open(payload, rng, pos)
        leaves = np.stack([np.frombuffer(
            hashlib.sha256(rng[32 * i:32 * i + 32].tobytes() + payload[i].tobytes()).digest(), np.uint8)
            for i in range(n)])
        root2, nodes = oracle.merkle_build(leaves)
        assert (root == root2).all()
        assert [p.tobytes() for p in path] == compressed_proof(nodes, n, pos)
        for k, q in enumerate(pos):
            assert (nonce[k] == rng[32 * q:32 * q + 32]).all()


def test_rust_tree_merkle_vector(oracle):
    b = open(f"{GOLDEN}/merkle_test_vector.bin", "rb").read()
    o = 0
    n = struct.unpack_from("<Q", b, o)[0]; o += 8
    leaves = np.frombuffer(b, np.uint8, 32 * n, o).reshape(n, 32); o += 32 * n
    nq = struct.unpack_from("<Q", b, o)[0]; o += 8
    pos = list(struct.unpack_from(f"<{nq}Q", b, o)); o += 8 * nq
    root = b[o:o + 32]; o += 32
    ln = struct.unpack_from("<Q", b, o)[0]; o += 8
    want = [b[o + 32 * i:o + 32 * i + 32] for i in range(ln)]; o += 32 * ln
    assert o == len(b)
    got_root, nodes = oracle.merkle_build(leaves)
    assert got_root.tobytes() == root
    assert compressed_proof(nodes, n, pos) == want


def test_rust_tree_commitment_vector(oracle):
    """MerkleCommitment with the counter RNG starting at 42 and leaf payload
    [3i,5i,7i,11i] mod 256 (merkle.rs:302-417)."""
    b = open(f"{GOLDEN}/commitment_test_vector.bin", "rb").read()
    o = 0
    n = struct.unpack_from("<Q", b, o)[0]; o += 8
    nq = struct.unpack_from("<Q", b, o)[0]; o += 8
    pos = list(struct.unpack_from(f"<{nq}Q", b, o)); o += 8 * nq
    root = b[o:o + 32]; o += 32
    nonces = [b[o + 32 * i:o + 32 * i + 32] for i in range(nq)]; o += 32 * nq
    ln = struct.unpack_from("<Q", b, o)[0]; o += 8
    want = [b[o + 32 * i:o + 32 * i + 32] for i in range(ln)]; o += 32 * ln
    assert o == len(b)
    idx = np.arange(n)
    payload = np.stack([idx * 3, idx * 5, idx * 7, idx * 11], axis=1).astype(np.uint8)
    rng = ((42 + np.arange(32 * n)) % 256).astype(np.uint8)
    got_root, got_nonce, got_path = oracle.merkle_commit_open(payload, rng, pos)
    assert got_root.tobytes() == root
    assert [x.tobytes() for x in got_nonce] == nonces
    assert [x.tobytes() for x in got_path] == want


FS1 = """8b297f0bffd583c6c6b6796385d5fd20a08665733b833970ebdd1054bbbc1b14
0667c08ad7f38efec5f30dc8aa4f20d749cdcf96d63a770f9810ac5c0ca8dcb1
c8037fc12d4da00b5dc7597e3042f33f72a06f970cb71fb6b103ebb5419d8a6b
fbbcfa1eac48728fbfdacc1c21e2f78119457e0846337e46140e38e62856c4c5
5358ae603691cc759faeb572fb6642654ea1c3dbc8f81d00276dd8c4df95aa58
5266158c3c895dede5a23b6ce85a9f564b8059ebfcd1741f54497ec58189873e
3ecea4b2343c007fc32f2aff40dc7320945f101ecae5d52494db21ad326e9739
6462dd575e6b874118607212feec7ce5417ae3bf0f2e86604596f35d48bbaea2
6d56c703c369edea3595db6b958241580ae9b4a76fead961413ed9e9e5852dcd
6d31073cee650212a71b7b13e9f951e00ef3b14a008a79dd95047b26a4a83d06
1b9e2a6666da63c43e52227d91a8a7f0bd5311f63c2e3a18839133375639e6cb
332ea49dd23dd4745631ecbb15696192b1fa127256baf7a0483fd27db6f09a48
43e735927ccbdc4d5ce912675d638d6d3dc8eef3def34504304e938846f157d6
dc4a8868ae75e733a7257a8589230392a98d78594836dfccd01304742b5b3ad5
976353931711c634f2691e507b119fd7f6e653d419a2620676122db08db18765
332729ab436dca654866a9382deaee0add6fb7e90a80261f1488e56598e8bc99""".split()
FS2 = """609db3e9a8f548df038519fa46cef23eb8c6553d3c1f698604e60a51613a738e
1cb69cb31999eb88e83c7586aac53f5e3286b084b0cf9e43619b48df01e0a310
3bf36e3ddc690a1b12b417628c115959b373d056c90c42dc2417baf46f538868
e336594f29dcda52e48896517b5cdb2d062ffd861ab02db5f8ca197aacc635f6""".split()
FS3 = """ae1a921288590205fc24543303ff527476359b8db4a983b2886a133b02f3217e
8c5d52a04b295f9fdb45ab66100fa00ca32c9634aa87cbbdb2bc3e1912459feb
12f82963b5b242156f6e9eb756eddee7652b60c7d6394403f7bd995e0b9bcd9c
880aa50b049b3939055deb7933749d338bb3fb5f64a9adf95019e6cfc232995c""".split()
NAT_ARGS = [1, 1, 1, 2, 2, 2, 7, 7, 7, 7, 32, 32, 32, 32, 256, 256, 256, 256, 1000, 10000, 60000, 65535,
            100000, 100000]
NAT_WANT = [0, 0, 0, 0, 0, 0, 3, 0, 4, 5, 10, 30, 27, 22, 100, 189, 3, 92, 999, 3105, 40886, 51590, 56367,
            10678]
CHOOSE = [(31, [10, 29, 30, 11, 4, 15, 16, 28, 19, 21, 25, 18, 17, 3, 5, 23, 24, 22, 6, 1]),
          (32, [3, 17, 18, 8, 30, 7, 14, 19, 25, 23, 12, 4, 31, 16, 0, 6, 20, 27, 11, 10]),
          (63, [9, 56, 61, 45, 35, 53, 51, 3, 39, 32, 31, 6, 59, 58, 54, 22, 27, 62, 55, 19]),
          (64, [12, 52, 39, 17, 51, 38, 58, 2, 28, 27, 46, 63, 61, 50, 40, 55, 47, 13, 56, 32]),
          (1000, [157, 668, 572, 138, 913, 994, 797, 249, 440, 723, 489, 241, 383, 108, 710, 341, 406, 585, 42,
                  692]),
          (65535, [40745, 48408, 17108, 44500, 53993, 10008, 24910, 52200, 61265, 54989, 41237, 25958, 28697,
                   61187, 34729, 3525, 9005, 38627, 9724, 12169])]


def _elt_le(x):
    return int(x).to_bytes(32, "little")


def test_fiat_shamir_spec_vectors(oracle):
    """Vectors 1-5 of docs/specs/testvectors.md over the secp256k1 prime."""
    u32 = lambda v: struct.pack("<I", v)
    s = b"B" + u32(100) + bytes(range(100)) + b"G" + u32(16)
    s += b"E" + _elt_le(7) + b"G" + u32(16)
    s += b"A" + u32(2) + _elt_le(8) + _elt_le(9) + b"G" + u32(16)
    s += b"B" + u32(4) + b"nats" + b"".join(b"N" + u32(a) for a in NAT_ARGS)
    # the reference test writes "choose" (transcript_test.cc:262); the spec text says "choice"
    s += b"B" + u32(6) + b"choose" + b"".join(b"C" + u32(m) + u32(20) for m, _ in CHOOSE)
    out = oracle.transcript_script(b"test", s, fid=oracle.FID_SECP256K1)
    elts = [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(48)]
    assert [f"{e:064x}" for e in elts[:16]] == FS1
    assert [f"{e:064x}" for e in elts[16:20]] == FS2
    assert [f"{e:064x}" for e in elts[32:36]] == FS3
    rest = np.frombuffer(out[48 * 32:], np.uint32)
    assert list(rest[:len(NAT_ARGS)]) == NAT_WANT
    k = len(NAT_ARGS)
    for m, want in CHOOSE:
        assert list(rest[k:k + 20]) == want, m
        k += 20


def test_rust_tree_transcript_vector(oracle):
    """rust/runtime/random/tests/transcript.rs:18-67 `test_transcript_cpp_compatibility` with the C++-generated
    transcript_test_vector.bin: 100 iterations of write(64 bytes), write0(100 + i), one P-256 element, an array
    of five, then 256 challenge bytes, on Transcript([1..8])."""
    want = open(f"{GOLDEN}/transcript_test_vector.bin", "rb").read()
    assert len(want) == 100 * 256
    u8 = lambda f: bytes((f(j) & 0xFF) for j in range(32))
    script = b""
    for i in range(100):
        script += b"B" + struct.pack("<I", 64) + bytes(((i * 7 + j * 13) & 0xFF) for j in range(64))
        script += b"Z" + struct.pack("<I", 100 + i)
        script += b"E" + u8(lambda j: i * 17 + j * 19)
        script += b"A" + struct.pack("<I", 5) + b"".join(u8(lambda j, k=k: i * 23 + k * 29 + j * 31) for k in range(5))
        script += b"R" + struct.pack("<I", 256)
    got = oracle.transcript_script(bytes(range(1, 9)), script, fid=1)
    assert got == want


def test_rust_tree_ligero_vector(oracle):
    """rust/runtime/ligero/tests/ligero.rs:592-760 `test_cpp_roundtrip_gf2_128` with the C++-generated
    ligero_test_vector.bin: a stand-alone Ligero statement over GF(2^128) -- 1000 witnesses of which the first
    950 lie in the subfield, 50 quadratic constraints, 1000 linear terms in 15 constraints, rate 4, nreq 36,
    block_enc 4096, coins from the test's 64-bit LCG (seed 100) -- with the commitment root and the 55 768
    serialised proof bytes the C++ prover produced.  The oracle's LigeroProver::commit + ::prove
    (ligero_prove_generic: the same code the ZK prover runs behind its sumcheck) reproduces both."""
    d = open(f"{GOLDEN}/ligero_test_vector.bin", "rb").read()
    off = 0

    def u64():
        nonlocal off
        v = struct.unpack_from("<Q", d, off)[0]
        off += 8
        return v
    nw, nq, nreq, nl, sb = [u64() for _ in range(5)]
    W = np.frombuffer(d, np.uint8, nw * 16, off).reshape(nw, 16)
    off += 2 * nw * 16          # W, then the (unused) inner-product vector A
    lqc = np.array([[u64(), u64(), u64()] for _ in range(nq)], np.uint64)
    tc, tw, tk = [], [], []
    for _ in range(u64()):
        tc.append(u64())
        tw.append(u64())
        tk.append(d[off:off + 16])
        off += 16
    off += nl * 16              # b: the verifier's side
    h, root = d[off:off + 32], d[off + 32:off + 64]
    off += 64
    plen = u64()
    proof = d[off:off + plen]
    assert off + plen == len(d) and (nw, nq, nreq, nl, sb) == (1000, 50, 36, 15, 950)
    st, coins = 100, bytearray(1 << 18)
    for i in range(len(coins)):   # SimpleRng of ligero.rs:28-42
        st = (st * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        coins[i] = (st >> 32) & 0xFF
    r, p, used = oracle.ligero_prove(4, W, lqc, tc, tw, np.frombuffer(b"".join(tk), np.uint8), nl, h, bytes(coins),
                                     subfield_boundary=sb, rate=4, nreq=nreq, block_enc=4096)
    assert r == root
    assert p == proof
    assert used == 144360


def test_transcript_key_and_prf_blocks(oracle):
    """lib/random/transcript_test.cc:285-341 Transcript.TestVec."""
    u32 = lambda v: struct.pack("<I", v)
    s = b"B" + u32(100) + bytes(range(100)) + b"K" + b"R" + u32(32) + b"B" + u32(1) + b"\0" + b"K"
    out = oracle.transcript_script(b"test", s)
    assert out[:32].hex() == "60cd1634920f1cf2ae831502bf4bb93a60cd03eeb19f93e2d6d50dbd0984cbd8"
    assert out[32:64].hex().upper() == "141BBCBB5410DDEB7039833B736586A0" "20FDD5856379B6C6C683D5FF0B7F298B"
    assert out[64:96].hex() == "181978380b6ff32185c828d9a007ee930bce2e947f887f85b64f399a94cbe4a8"
    msg = b"\0" + (4).to_bytes(8, "little") + b"test" + b"\0" + (100).to_bytes(8, "little") + bytes(range(100))
    assert hashlib.sha256(msg).digest() == out[:32] == oracle.sha256(msg)
    blocks = (0).to_bytes(16, "little") + (1).to_bytes(16, "little")
    assert oracle.aes256_ecb(out[:32], blocks) == out[32:64]


def test_reference_golden_proofs(oracle):
    """oracle == unmodified reference on whole serialized proofs (SHA-256 over
    GF(2^128) and ECDSA P-256 over Fp256), from tests/golden/golden.json."""
    g = golden()
    for name, nproofs in (("sha1_gf128", 2), ("ecdsa1_p256", 1)):
        circ, wit = load(name)
        assert hashlib.sha256(circ).hexdigest() == g[name]["circuit_sha256"]
        c = oracle.Circuit(g[name]["field_id"], circ)
        assert c.id() == circ[-32:]
        for pr in g[name]["proofs"][:nproofs]:
            r = c.prove(wit, rng_bytes(pr["seed"], 1 << 19), tinit=pr["tinit"].encode(), dump=True)
            assert r["rng_used"] == pr["rng_used"]
            assert r["root"].hex() == pr["root"]
            assert len(r["proof"]) == pr["proof_len"]
            assert hashlib.sha256(r["proof"]).hexdigest() == pr["proof_sha256"]


def test_reference_known_answer_zk_vector(oracle):
    """rust/runtime/zk/tests/zk.rs:228-558 `test_zk_rfc_testvector1`: a 3-term circuit, rate 4, nreq 6,
    block_enc 128, every RandomEngine call returning 02 00 ... 00.  Commitment root, sumcheck proof and
    Ligero proof (subfield run-length coding included) are the bytes that test holds."""
    from fixtures import load_rfc_vector
    rec, circ, wit, coins, want = load_rfc_vector(oracle)
    c = oracle.Circuit(rec["field_id"], circ)
    assert c.id() == circ[-32:]
    r = c.prove(wit, coins, tinit=rec["transcript_seed"].encode(), rate=rec["rate"], nreq=rec["nreq"],
                block_enc=rec["block_enc"], dump=True)
    assert r["rng_used"] == coins.size
    assert r["root"] == bytes.fromhex(rec["commitment"])
    assert r["proof"][32:32 + 160] == bytes.fromhex(rec["sumcheck_proof"])
    assert r["proof"][192:] == bytes.fromhex(rec["ligero_proof"])
    assert r["proof"] == want


def test_lch14_known_answers(oracle):
    ka = golden()["known_answers"]
    rs = np.random.default_rng(99)
    for n, m in [(455, 4096), (909, 4096), (455, 909), (5, 16)]:
        rows = rs.integers(0, 256, (2, m, 16), dtype=np.uint8)
        got = oracle.lch14_interpolate(n, m, rows)
        assert hashlib.sha256(got.tobytes()).hexdigest() == ka[f"lch14_rs_{n}_{m}"]["sha256"]


def test_bad_witness_rejected(oracle):
    circ, wit = load("sha1_gf128")
    bad = bytearray(wit)
    bad[16 * 20] ^= 1
    import pytest
    with pytest.raises(RuntimeError, match="rc=-3"):
        oracle.Circuit(4, circ).prove(bytes(bad), rng_bytes(1, 1 << 18))


def test_mdoc_fixture_is_consistent(oracle):
    """tests/golden/mdoc: the circuit file holds the signature circuit (Fp256) then the hash circuit
    (GF(2^128)); the frozen witnesses have the circuits' input counts and differ only in public inputs"""
    from fixtures import load_mdoc
    f = load_mdoc()
    raw = f["raw"]
    assert raw[0] == 1 and int.from_bytes(raw[1:4], "little") == 1          # version, P256_ID
    nin_sig, npub_sig = int.from_bytes(raw[16:19], "little"), int.from_bytes(raw[10:13], "little")
    assert f["w_sig"].size == nin_sig * 32 and f["w_sig_mac"].size == nin_sig * 32
    d = np.nonzero((f["w_sig"] != f["w_sig_mac"]).reshape(nin_sig, 32).any(axis=1))[0]
    assert d.size > 0 and d.max() < npub_sig                                 # the MAC patch touches public inputs only
    assert oracle.Circuit(oracle.P256_ID, raw) is not None
    e = f["expect"]
    assert e["proof_len"] == 96 + e["len_hash"] + e["len_sig"] and f["coins"].size == e["coins_total"]


import pytest  # noqa: E402

from fixtures import SIZE_NAMES, load_size, load_witnesses  # noqa: E402


@pytest.mark.parametrize("name", SIZE_NAMES)
def test_published_instance_sizes_match_the_reference(oracle, name):
    """BM_ShaZK_fp2_128/2..33 and BM_ECDSAZKProver/2,3 (docs/content/en/docs/benchmarks.md:56-61,74-75):
    the oracle's proof for seed 1 is the reference's, byte for byte (tests/golden/sizes/golden_sizes.json)"""
    circ, wit, g = load_size(name)
    assert hashlib.sha256(circ).hexdigest() == g["circuit_sha256"]
    r = oracle.Circuit(g["field_id"], circ).prove(wit, rng_bytes(1, 1 << 22))
    assert r["rng_used"] == g["proof"]["rng_used"]
    assert len(r["proof"]) == g["proof"]["proof_len"]
    assert hashlib.sha256(r["proof"]).hexdigest() == g["proof"]["proof_sha256"]


@pytest.mark.parametrize("name,fid", [("sha1_gf128", 4), ("ecdsa1_p256", 1)])
def test_distinct_witnesses_are_distinct_and_satisfying(oracle, name, fid):
    circ, wit = load(name)
    W = load_witnesses(name)
    assert W.shape[0] >= 8 and len({w.tobytes() for w in W}) == W.shape[0]
    c = oracle.Circuit(fid, circ)
    for w in W[:3]:
        assert len(c.prove(w.tobytes(), rng_bytes(2, 1 << 19))["proof"]) > 100000
