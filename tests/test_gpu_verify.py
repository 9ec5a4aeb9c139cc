"""GPU verifier (lf_zk_verify_batch) against the unmodified reference's ZkProof::read + ZkVerifier
(zk/zk_verifier.h:69-106, ligero/ligero_verifier.h:42-268, merkle/merkle_tree.h:153-214): the same
accept / reject decision on GPU proofs, on proofs the oracle and the reference produced, and on the
tampered-proof negatives of ligero_test.cc:114-268 / zk_test.cc:147-201 (every region of the
serialized proof corrupted in turn, wrong public inputs, truncations, broken run-length headers)."""
import numpy as np
import pytest

from fixtures import load, rng_bytes

pytestmark = pytest.mark.gpu

CASES = [("sha1_gf128", 4), ("ecdsa1_p256", 1)]


def _regions(info):
    """byte ranges of the fixed-position parts of a serialized proof (zk_proof.h:90-105)"""
    kb = info["kbytes"]
    o_sc = 32
    o_ldt = o_sc + info["sumcheck_proof_elts"] * kb
    o_dot = o_ldt + info["block"] * kb
    o_q0 = o_dot + info["dblock"] * kb
    o_q2 = o_q0 + info["r"] * kb
    o_nonce = o_q2 + (info["dblock"] - info["block"]) * kb
    o_req = o_nonce + info["nreq"] * 32
    return dict(root=(0, 32), sumcheck=(o_sc, o_ldt), y_ldt=(o_ldt, o_dot), y_dot=(o_dot, o_q0), y_quad_0=(o_q0, o_q2),
                y_quad_2=(o_q2, o_nonce), nonce=(o_nonce, o_req)), o_req


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def case(request, ctx, oracle):
    import longfellow_zk_b200 as lf
    name, fid = request.param
    circ, wit = load(name)
    c = lf.Circuit(ctx, fid, circ)
    n = c.info["rng_bytes"]
    B = 6
    rng = np.stack([rng_bytes(700 + i, n + 256) for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert (status == 0).all()
    pub = np.frombuffer(wit, np.uint8)[:c.info["npub_in"] * c.info["kbytes"]].copy()
    return dict(c=c, circ=circ, wit=wit, fid=fid, proofs=proofs, pub=pub, v=lf.ZkVerifier(c), name=name)


def _pubs(case, B):
    return np.repeat(case["pub"][None, :], B, axis=0)


def test_gpu_proofs_are_accepted(case):
    st, why = case["v"].verify_batch(_pubs(case, len(case["proofs"])), case["proofs"])
    assert (st == 0).all() and (why == 0).all(), (st, why)


def test_oracle_and_reference_proofs_are_accepted(case, oracle, ref):
    c = case["c"]
    coins = rng_bytes(31, c.info["rng_bytes"] + 256)
    po = oracle.Circuit(case["fid"], case["circ"]).prove(case["wit"], coins)["proof"]
    pr = ref.Circuit(case["fid"], case["circ"]).prove(case["wit"], coins)["proof"]
    assert po == pr
    st, _ = case["v"].verify_batch(_pubs(case, 2), [po, pr], tinit=b"test")
    assert (st == 0).all(), st
    # a proof made on another transcript seed does not verify on this one, and does on its own
    p2 = oracle.Circuit(case["fid"], case["circ"]).prove(case["wit"], coins, tinit=b"other seed")["proof"]
    st, _ = case["v"].verify_batch(_pubs(case, 2), [p2, p2], tinit=b"test")
    assert (st == -8).all(), st
    st, _ = case["v"].verify_batch(_pubs(case, 1), [p2], tinit=b"other seed")
    assert st[0] == 0


def _ref_verdict(rc):
    # refapi.verify: 0 accepted, 1 ZkProof::read failed, 2 trailing bytes (the test wrapper's own rule), 3 rejected
    return {0: 0, 1: -3, 3: -8}[rc]


def test_tampered_regions_agree_with_the_reference(case, ref):
    """one byte flipped in every fixed region, at its first, a middle and its last element"""
    info = case["c"].info
    regions, o_req = _regions(info)
    base = case["proofs"][0]
    bad, names = [], []
    for name, (lo, hi) in regions.items():
        if hi == lo:
            continue
        for pos in sorted({lo, (lo + hi) // 2, hi - 1}):
            b = bytearray(base)
            b[pos] ^= 0x01
            bad.append(bytes(b))
            names.append(f"{name}@{pos}")
    # the opened columns and the Merkle proof: first run header is 4 bytes, then elements
    for pos in (o_req + 4, o_req + 4 + info["kbytes"] * 7, (o_req + len(base)) // 2, len(base) - 1, len(base) - 33):
        b = bytearray(base)
        b[pos] ^= 0x80 if info["kbytes"] == 16 else 0x01
        bad.append(bytes(b))
        names.append(f"tail@{pos}")
    bad.append(base)
    names.append("untouched")
    st, why = case["v"].verify_batch(_pubs(case, len(bad)), bad)
    rc = ref.Circuit(case["fid"], case["circ"])
    for i, (pr, nm) in enumerate(zip(bad, names)):
        want = _ref_verdict(rc.verify(case["pub"].tobytes(), pr))
        assert st[i] == want, (nm, st[i], why[i], want)
    assert st[-1] == 0
    assert (st[:-1] != 0).all()
    # which check fires first (ligero_verifier.h:92-134)
    by = dict(zip(names, zip(st, why)))
    lo = regions["root"][0]
    assert by[f"root@{lo}"] == (-8, 1)
    lo = regions["nonce"][0]
    assert by[f"nonce@{lo}"] == (-8, 1)
    # a changed response changes the transcript, hence the opened columns: the Merkle check fires first
    lo = regions["y_ldt"][0]
    assert by[f"y_ldt@{lo}"] == (-8, 1)


def test_wrong_public_inputs_and_swapped_proofs(case, ref):
    info = case["c"].info
    if info["npub_in"] == 0:
        pytest.skip("no public inputs")
    kb = info["kbytes"]
    proofs = case["proofs"][:4]
    pubs = _pubs(case, 4).copy()
    pubs[1, kb * (info["npub_in"] - 1)] ^= 1   # last public input changed
    pubs[2, kb * 1] ^= 1                        # an early one
    st, why = case["v"].verify_batch(pubs, proofs)
    rc = ref.Circuit(case["fid"], case["circ"])
    for i in range(4):
        assert st[i] == _ref_verdict(rc.verify(pubs[i].tobytes(), proofs[i])), i
    assert st[0] == 0 and st[3] == 0 and st[1] == -8 and st[2] == -8
    # the public inputs are part of the transcript (zk_common.h:163-180): every challenge changes, and the
    # first check that notices is the Merkle check of the (different) opened columns
    assert why[1] == 1 and why[2] == 1


def test_malformed_proofs_are_format_errors(case, ref):
    info = case["c"].info
    regions, o_req = _regions(info)
    base = case["proofs"][0]
    bad = [base[:10], base[:o_req - 1], base[:o_req + 3], base[:len(base) - 1], base[:len(base) - 32],
           base[:o_req] + (0xFFFFFFFF).to_bytes(4, "little") + base[o_req + 4:],
           base[:o_req] + (1 << 24).to_bytes(4, "little") + base[o_req + 4:],
           b""]
    if info["kbytes"] == 16:
        # (over a prime field every element is a "subfield" element: the first run is empty as written)
        bad.append(base[:o_req] + (0).to_bytes(4, "little") + base[o_req + 4:])
    if info["kbytes"] == 32:
        # an element >= p in the sumcheck proof: of_bytes_field fails
        b = bytearray(base)
        b[32:64] = b"\xff" * 32
        bad.append(bytes(b))
    st, why = case["v"].verify_batch(_pubs(case, len(bad)), bad)
    rc = ref.Circuit(case["fid"], case["circ"])
    for i, pr in enumerate(bad):
        want = rc.verify(case["pub"].tobytes(), pr)
        assert want in (1, 3), (i, want)
        assert st[i] == _ref_verdict(want), (i, st[i], want)
    # trailing bytes behind a complete proof are not looked at (ZkProof::read consumes what it parses)
    st, _ = case["v"].verify_batch(_pubs(case, 1), [base + b"\x00" * 40])
    assert st[0] == 0


def test_each_ligero_check_fires(case):
    """ligero_test.cc:160-268 (test_ligero_verifier_failures): with one of the verifier's interpolations
    broken (BadReedSolomonFactory), or b changed, a valid proof fails at exactly that check"""
    v = case["v"]
    proofs = case["proofs"][:3]
    want = {1: 2, 2: 3, 3: 5, 4: 3, 5: 4}
    try:
        for fault, why_code in want.items():
            v.set_fault(fault)
            st, why = v.verify_batch(_pubs(case, 3), proofs)
            assert (st == -8).all() and (why == why_code).all(), (fault, st, why)
    finally:
        v.set_fault(0)
    st, why = v.verify_batch(_pubs(case, 3), proofs)
    assert (st == 0).all()


def test_prover_still_works_after_a_verify(case, oracle):
    """the verifier reuses the prover's per-proof buffers"""
    import longfellow_zk_b200 as lf
    c = case["c"]
    coins = rng_bytes(5, c.info["rng_bytes"] + 256)
    proofs, status = lf.ZkProver(c).prove_batch(np.frombuffer(case["wit"], np.uint8)[None, :], coins[None, :])
    assert status[0] == 0
    assert proofs[0] == oracle.Circuit(case["fid"], case["circ"]).prove(case["wit"], coins)["proof"]


@pytest.mark.parametrize("name,fid", CASES)
def test_full_batch_verifies(ctx, name, fid):
    """1024 distinct proofs, every 64th corrupted somewhere: exactly those are rejected"""
    import longfellow_zk_b200 as lf
    circ, wit = load(name)
    c = lf.Circuit(ctx, fid, circ)
    B, n = 1024, c.info["rng_bytes"]
    rng = np.random.default_rng(77).integers(0, 256, (B, n + 256), dtype=np.uint8)
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert (status == 0).all()
    rs = np.random.default_rng(3)
    badset = set(range(5, B, 64))
    for i in badset:
        b = bytearray(proofs[i])
        b[int(rs.integers(0, len(b) - 40))] ^= 0x04
        proofs[i] = bytes(b)
    pub = np.frombuffer(wit, np.uint8)[:c.info["npub_in"] * c.info["kbytes"]]
    st, why = lf.ZkVerifier(c).verify_batch(np.repeat(pub[None, :], B, axis=0), proofs)
    for i in range(B):
        assert (st[i] != 0) == (i in badset), (i, st[i], why[i])


@pytest.mark.parametrize("rate,nreq,block_enc,tinit", [(4, 189, 0, b"test"), (7, 132, 4151, b"mdoc-style block_enc"),
                                                       (2, 64, 0, b""), (16, 40, 0, b"x" * 100)])
def test_verifier_other_ligero_parameters(ctx, oracle, ref, rate, nreq, block_enc, tinit):
    """other (rate, nreq, block_enc) choices and transcript seeds: proofs of the oracle and of the GPU are
    accepted under the same parameters, refused under the default ones, and the reference agrees"""
    import longfellow_zk_b200 as lf
    circ, wit = load("sha1_gf128")
    c = lf.Circuit(ctx, 4, circ, rate=rate, nreq=nreq, block_enc=block_enc)
    coins = rng_bytes(13, c.info["rng_bytes"])
    po = oracle.Circuit(4, circ).prove(wit, coins, tinit=tinit, rate=rate, nreq=nreq, block_enc=block_enc)["proof"]
    pg, st = lf.ZkProver(c).prove_batch(np.frombuffer(wit, np.uint8)[None, :], coins[None, :], tinit=tinit)
    assert st[0] == 0 and pg[0] == po
    bad = bytearray(po)
    bad[len(bad) // 3] ^= 2
    stv, _ = lf.ZkVerifier(c).verify_batch(None, [po, bytes(bad)], tinit=tinit)
    assert stv[0] == 0 and stv[1] != 0
    rc = ref.Circuit(4, circ)
    assert rc.verify(b"", po, tinit=tinit, rate=rate, nreq=nreq, block_enc=block_enc) == 0
    assert rc.verify(b"", bytes(bad), tinit=tinit, rate=rate, nreq=nreq, block_enc=block_enc) != 0
    if (rate, nreq, block_enc) != (7, 132, 0):
        c0 = lf.Circuit(ctx, 4, circ)
        stv, _ = lf.ZkVerifier(c0).verify_batch(None, [po], tinit=tinit)
        assert stv[0] != 0


@pytest.mark.parametrize("B", [3, 40])
def test_verify_on_a_caller_owned_transcript(case, B):
    """lf_zk_verify_committed_batch (ZkVerifier::recv_commitment / ::verify on the caller's Transcript,
    zk_verifier.h:69-106): the transcript that received the commitment continues on the device and comes back
    as verify left it -- which is, by Fiat-Shamir, the state the prover's transcript is in after prove
    (lf_zk_prove_committed_batch), so both must yield the same next challenge bytes.  A transcript that did
    not receive the commitment (or received another) makes the same proof fail.  B = 3: head of the
    transcript on the host; 40: on the device."""
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    c = case["c"]
    n = c.info["rng_bytes"]
    seed = b"caller-owned"
    rng = np.stack([rng_bytes(900 + i, n + 256) for i in range(B)])
    W = np.repeat(np.frombuffer(case["wit"], np.uint8)[None, :], B, axis=0)
    p = lf.ZkProver(c)
    tp = api.transcripts(B, seed)
    roots, st = p.commit_batch(W, rng, tp)
    assert (st == 0).all()
    proofs, st = p.prove_committed_batch(W, tp)
    assert (st == 0).all()
    tv = api.transcripts(B, seed)
    for i in range(B):
        assert proofs[i][:32] == bytes(roots[i])
        api.transcript_write(tv[i], proofs[i][:32])      # recv_commitment (ligero_transcript.h:31-34)
    status, why = case["v"].verify_batch(_pubs(case, B), proofs, transcripts=tv)
    assert (status == 0).all(), (status, why)
    for i in (0, B - 1):
        assert api.transcript_challenge(tv[i], 48) == api.transcript_challenge(tp[i], 48)
    # the same call equals the self-contained one
    status, _ = case["v"].verify_batch(_pubs(case, B), proofs, tinit=seed)
    assert (status == 0).all()
    # no commitment received / the neighbour's commitment received: rejected
    tv = api.transcripts(B, seed)
    for i in range(1, B):
        api.transcript_write(tv[i], proofs[i - 1][:32])
    status, _ = case["v"].verify_batch(_pubs(case, B), proofs, transcripts=tv)
    assert (status != 0).all()
