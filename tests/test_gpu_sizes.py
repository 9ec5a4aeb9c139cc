"""GPU parity on the other instance sizes the reference publishes numbers for (BM_ShaZK_fp2_128/2,4,8,16,32,33
and BM_ECDSAZKProver/2,3: docs/content/en/docs/benchmarks.md:56-61,74-75) and on batches whose proofs
have DIFFERENT witnesses (other SHA-256 messages, other signatures)."""
import hashlib

import numpy as np
import pytest

from fixtures import SIZE_NAMES, load, load_size, load_witnesses, rng_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", SIZE_NAMES)
def test_published_sizes_match_reference_golden(ctx, name):
    """proof for seed 1 == the reference's (golden hash); all proofs of the batch pass the GPU verifier"""
    import longfellow_zk_b200 as lf
    circ, wit, g = load_size(name)
    c = lf.Circuit(ctx, g["field_id"], circ)
    assert c.info["nterms"] == g["info"]["nterms"] and c.info["ninputs"] == g["info"]["ninputs"]
    assert c.info["rng_bytes"] == g["proof"]["rng_used"]
    assert c.verify_id() == circ[-32:]
    B, n = 3, c.info["rng_bytes"]
    rng = np.stack([rng_bytes(1 + i, 1 << 22)[:n + 256] for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert (status == 0).all()
    assert len(proofs[0]) == g["proof"]["proof_len"]
    assert hashlib.sha256(proofs[0]).hexdigest() == g["proof"]["proof_sha256"]
    pub = np.frombuffer(wit, np.uint8)[:c.info["npub_in"] * c.info["kbytes"]]
    st, why = lf.ZkVerifier(c).verify_batch(np.repeat(pub[None, :], B, axis=0), proofs)
    assert (st == 0).all(), (st, why)


@pytest.mark.parametrize("name,B", [("sha2_gf128", 256), ("ecdsa2_p256", 256), ("sha4_gf128", 40)])
def test_larger_instances_in_every_batch_shape(ctx, oracle, ref, name, B):
    """the batch-synchronous (flat) sumcheck path and the cluster path on a larger circuit, against the oracle"""
    import longfellow_zk_b200 as lf
    circ, wit, g = load_size(name)
    fid = g["field_id"]
    c = lf.Circuit(ctx, fid, circ)
    n = c.info["rng_bytes"]
    streams = [rng_bytes(60 + s, 1 << 22)[:n + 256].copy() for s in range(3)]
    want = [oracle.Circuit(fid, circ).prove(wit, s)["proof"] for s in streams]
    rng = np.stack([streams[i % 3] for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert (status == 0).all()
    for i, pr in enumerate(proofs):
        assert pr == want[i % 3], f"proof {i} of {B}"
    pub = wit[:c.info["npub_in"] * c.info["kbytes"]]
    assert ref.Circuit(fid, circ).verify(pub, proofs[B - 1]) == 0


@pytest.mark.parametrize("name,fid", [("sha1_gf128", 4), ("ecdsa1_p256", 1)])
def test_batch_of_distinct_witnesses(ctx, oracle, ref, name, fid):
    """1024 proofs whose witnesses cycle through all distinct fixtures (other messages / signatures), own coins
    each: sampled proofs equal the oracle's, every proof is accepted by the GPU verifier with ITS public inputs,
    and by the unmodified reference verifier"""
    import longfellow_zk_b200 as lf
    circ, _ = load(name)
    Ws = load_witnesses(name)
    c = lf.Circuit(ctx, fid, circ)
    B, n, kb, npub = 1024, c.info["rng_bytes"], c.info["kbytes"], c.info["npub_in"]
    rng = np.random.default_rng(11).integers(0, 256, (B, n + 256), dtype=np.uint8)
    W = Ws[np.arange(B) % Ws.shape[0]]
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert (status == 0).all()
    assert len({hashlib.sha256(p).digest() for p in proofs}) == B
    oc = oracle.Circuit(fid, circ)
    for i in (0, 1, Ws.shape[0] + 2, B - 1):
        assert proofs[i] == oc.prove(W[i].tobytes(), rng[i])["proof"], i
    pubs = np.ascontiguousarray(W[:, :npub * kb])
    st, why = lf.ZkVerifier(c).verify_batch(pubs if npub else None, proofs)
    assert (st == 0).all(), np.nonzero(st)[0][:5]
    rc = ref.Circuit(fid, circ)
    for i in range(B):
        assert rc.verify(pubs[i].tobytes(), proofs[i]) == 0, i
    if npub:
        # a proof checked against ANOTHER witness's public inputs is rejected
        st, _ = lf.ZkVerifier(c).verify_batch(np.roll(pubs[:Ws.shape[0]], 1, axis=0), proofs[:Ws.shape[0]])
        assert (st == -8).all()
