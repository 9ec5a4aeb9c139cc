"""GPU parity of the whole prover: same witness, same RNG bytes, same transcript
seed => the serialized proof must be byte-identical to the reference's
(tests/golden/golden.json, produced by the unmodified reference) and to the
oracle's, stage by stage."""
import hashlib

import numpy as np
import pytest

from fixtures import golden, load, rng_bytes

pytestmark = pytest.mark.gpu

GF = 4


@pytest.fixture(scope="module")
def sha(ctx):
    import longfellow_zk_b200 as lf
    circ, wit = load("sha1_gf128")
    c = lf.Circuit(ctx, GF, circ)
    return c, circ, wit


def test_sha_circuit_info(sha):
    c, circ, wit = sha
    g = golden()["sha1_gf128"]
    assert c.info["ninputs"] == g["info"]["ninputs"] == 3721
    assert c.info["nterms"] == g["info"]["nterms"]
    assert (c.info["block_enc"], c.info["block"], c.info["dblock"], c.info["nrow"]) == (4096, 455, 909, 20)
    assert c.info["rng_bytes"] == g["proofs"][0]["rng_used"]
    assert c.info["witness_bytes"] == len(wit)


def test_sha_proof_stages_match_oracle(sha, oracle):
    import longfellow_zk_b200 as lf
    c, circ, wit = sha
    rng = rng_bytes(1, c.info["rng_bytes"])
    want = oracle.Circuit(GF, circ).prove(wit, rng, dump=True)
    p = lf.ZkProver(c)
    proofs, status = p.prove_batch(np.frombuffer(wit, np.uint8)[None, :], rng[None, :])
    nw = c.info["nw"]
    got_w = p.debug_fetch(0, 1)
    assert (got_w == want["witness"][:nw * 16]).all(), "Ligero witness (inputs || pad)"
    got_t = p.debug_fetch(0, 2).reshape(20, 4096, 16)
    want_t = want["tableau"][:20 * 4096 * 16].reshape(20, 4096, 16)
    for row in range(20):
        assert (got_t[row, :909] == want_t[row, :909]).all(), f"tableau row {row} message part"
        assert (got_t[row] == want_t[row]).all(), f"tableau row {row} extension"
    assert p.debug_fetch(0, 3).tobytes() == want["root"], "Merkle root"
    got_sc = p.debug_fetch(0, 4)
    want_sc = want["sumcheck"][:got_sc.size]
    if not (got_sc == want_sc).all():
        bad = np.nonzero((got_sc != want_sc).reshape(-1, 16).any(axis=1))[0]
        raise AssertionError(f"sumcheck proof differs first at element {bad[0]} of {got_sc.size // 16}")
    assert status[0] == 0
    assert len(proofs[0]) == len(want["proof"])
    if proofs[0] != want["proof"]:
        a, b = np.frombuffer(proofs[0], np.uint8), np.frombuffer(want["proof"], np.uint8)
        raise AssertionError(f"proof bytes differ first at offset {np.nonzero(a != b)[0][0]} of {a.size}")


def test_sha_proofs_match_reference_golden(sha):
    """three seeded proofs in one batch against the reference's own output"""
    import longfellow_zk_b200 as lf
    c, circ, wit = sha
    g = golden()["sha1_gf128"]
    seeds = [pr["seed"] for pr in g["proofs"]]
    rng = np.stack([rng_bytes(s, 1 << 19)[:c.info["rng_bytes"]] for s in seeds])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], len(seeds), axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    for pr, got, st in zip(g["proofs"], proofs, status):
        assert st == 0
        assert len(got) == pr["proof_len"]
        assert got[:64].hex() == pr["proof_head"]
        assert hashlib.sha256(got).hexdigest() == pr["proof_sha256"]


def test_bad_witness_is_rejected(sha):
    import longfellow_zk_b200 as lf
    c, circ, wit = sha
    bad = bytearray(wit)
    bad[16 * 20] ^= 1  # flip an input bit: the hash no longer matches
    rng = rng_bytes(5, c.info["rng_bytes"])
    W = np.stack([np.frombuffer(bytes(bad), np.uint8), np.frombuffer(wit, np.uint8)])
    proofs, status = lf.ZkProver(c).prove_batch(W, np.stack([rng, rng]))
    assert status[0] == -5 and proofs[0] == b""
    assert status[1] == 0 and len(proofs[1]) > 100000


def test_reference_verifier_accepts_gpu_proofs(sha, ref):
    """the unmodified reference ZkVerifier accepts what the GPU produced"""
    import longfellow_zk_b200 as lf
    c, circ, wit = sha
    B = 8
    rng = np.stack([rng_bytes(100 + i, c.info["rng_bytes"]) for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    rc = ref.Circuit(GF, circ)
    for pr, st in zip(proofs, status):
        assert st == 0
        assert rc.verify(b"", pr) == 0


@pytest.mark.parametrize("rate,nreq,block_enc,tinit", [(4, 189, 0, b"test"), (7, 132, 4151, b"mdoc-style block_enc"),
                                                       (2, 64, 0, b""), (16, 40, 0, b"x" * 100)])
def test_sha_proof_other_ligero_parameters(ctx, oracle, rate, nreq, block_enc, tinit):
    """other (rate, nreq, block_enc) choices of LigeroParam (ligero_param.h:116-307; the mdoc specs use
    block_enc 4151, zk_spec.cc:47-49) and other transcript seeds, against the oracle"""
    import longfellow_zk_b200 as lf
    circ, wit = load("sha1_gf128")
    c = lf.Circuit(ctx, GF, circ, rate=rate, nreq=nreq, block_enc=block_enc)
    rng = rng_bytes(9, c.info["rng_bytes"])
    want = oracle.Circuit(GF, circ).prove(wit, rng, tinit=tinit, rate=rate, nreq=nreq, block_enc=block_enc)
    assert c.info["rng_bytes"] == want["rng_used"]
    proofs, status = lf.ZkProver(c).prove_batch(np.frombuffer(wit, np.uint8)[None, :], rng[None, :], tinit=tinit)
    assert status[0] == 0 and proofs[0] == want["proof"]


@pytest.mark.parametrize("name,fid,B", [("sha1_gf128", 4, 10), ("sha1_gf128", 4, 20), ("sha1_gf128", 4, 40),
                                        ("sha1_gf128", 4, 100), ("sha1_gf128", 4, 160), ("sha1_gf128", 4, 300),
                                        ("sha1_gf128", 4, 600), ("sha1_gf128", 4, 1040), ("ecdsa1_p256", 1, 10),
                                        ("ecdsa1_p256", 1, 20),
                                        ("ecdsa1_p256", 1, 40), ("ecdsa1_p256", 1, 100), ("ecdsa1_p256", 1, 160),
                                        ("ecdsa1_p256", 1, 300), ("ecdsa1_p256", 1, 600)])
def test_every_sumcheck_shape_matches_oracle(ctx, oracle, name, fid, B):
    """The batch size selects the sumcheck launch shape: a cluster of 16 CTAs per proof up to 8 proofs
    (covered by the small-batch tests), of 8 up to 12, of 4 up to 32, of 2 up to 64, one 1024-thread CTA
    per proof below 148, then 512-thread x 2, 256-thread x 4, from 592 proofs on 128-thread x 7 and above
    1036 proofs 128-thread x 8 CTAs per SM.  Every proof of a
    batch that mixes four coin streams must equal the oracle's proof for its stream, wherever it sits."""
    import longfellow_zk_b200 as lf
    circ, wit = load(name)
    c = lf.Circuit(ctx, fid, circ)
    n = c.info["rng_bytes"]
    streams = [rng_bytes(40 + s, 1 << 19)[:n].copy() for s in range(4)]
    want = [oracle.Circuit(fid, circ).prove(wit, s)["proof"] for s in streams]
    rng = np.stack([streams[i % 4] for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert (status == 0).all()
    for i, pr in enumerate(proofs):
        assert pr == want[i % 4], f"proof {i} of {B}"


def test_growing_and_shrinking_batches_on_one_circuit(sha, oracle):
    """staging and per-proof buffers are re-allocated when a later batch is larger"""
    import longfellow_zk_b200 as lf
    c, circ, wit = sha
    p = lf.ZkProver(c)
    want = oracle.Circuit(GF, circ).prove(wit, rng_bytes(1, c.info["rng_bytes"]))["proof"]
    for B in (1, 3, 2, 11, 5, 300, 16, 1, 40):
        rng = np.stack([rng_bytes(1 + i, c.info["rng_bytes"]) for i in range(B)])
        W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
        proofs, status = p.prove_batch(W, rng)
        assert (status == 0).all() and proofs[0] == want, B


def test_reference_known_answer_zk_vector(ctx, oracle):
    """The reference's own known-answer test of the whole ZK prover (rust/runtime/zk/tests/zk.rs:228-558,
    bytes from the C++ prover; tests/golden/rfc_zk_vector1.json): the GPU's proof of that 3-term circuit
    with rate 4, nreq 6, block_enc 128 and the test's constant coins equals the recorded commitment,
    sumcheck proof and Ligero proof, alone and in a batch, and the GPU verifier accepts it."""
    import longfellow_zk_b200 as lf
    from fixtures import load_rfc_vector
    rec, circ, wit, coins, want = load_rfc_vector(oracle)
    tinit = rec["transcript_seed"].encode()
    c = lf.Circuit(ctx, rec["field_id"], circ, rate=rec["rate"], nreq=rec["nreq"], block_enc=rec["block_enc"])
    assert c.info["rng_bytes"] == coins.size and c.info["max_proof_bytes"] >= len(want)
    W = np.frombuffer(wit, np.uint8)
    for B in (1, 70):
        proofs, status = lf.ZkProver(c).prove_batch(np.repeat(W[None, :], B, axis=0),
                                                    np.repeat(coins[None, :], B, axis=0), tinit=tinit)
        assert (status == 0).all()
        assert all(p == want for p in proofs), B
    pub = np.frombuffer(wit[:16 * c.info["npub_in"]], np.uint8)[None, :]
    status, why = lf.ZkVerifier(c).verify_batch(pub, [want], tinit=tinit)
    assert status[0] == 0, why
    bad = bytearray(want)
    bad[40] ^= 1
    status, why = lf.ZkVerifier(c).verify_batch(pub, [bytes(bad)], tinit=tinit)
    assert status[0] != 0
