"""ZkProver::commit and ::prove as separate calls on caller-owned transcripts
(lf_zk_commit_batch / lf_zk_prove_committed_batch / lf_transcript_*), the shape run_mdoc_prover needs:
two provers over two fields sharing ONE transcript (lib/circuits/mdoc/mdoc_zk.cc:459-503)."""
import numpy as np
import pytest

from fixtures import load, rng_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,fid,B", [("sha1_gf128", 4, 3), ("ecdsa1_p256", 1, 2), ("sha1_gf128", 4, 160)])
def test_commit_then_prove_equals_one_call(ctx, name, fid, B):
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    circ, wit = load(name)
    c = lf.Circuit(ctx, fid, circ)
    n = c.info["rng_bytes"]
    rng = np.stack([rng_bytes(70 + (i % 3), 1 << 19)[:n].copy() for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    p = lf.ZkProver(c)
    want, st = p.prove_batch(W, rng, tinit=b"split")
    assert (st == 0).all()
    ts = api.transcripts(B, b"split")
    roots, st = p.commit_batch(W, rng, ts)
    assert (st == 0).all()
    assert roots[0].tobytes() == p.debug_fetch(0, 3).tobytes()
    got, st = p.prove_committed_batch(W, ts)
    assert (st == 0).all()
    for i in range(B):
        assert got[i] == want[i], i
    # a second prove without a new commit is refused
    with pytest.raises(lf.LongfellowError):
        p.prove_committed_batch(W, ts)


def test_two_provers_on_one_transcript_match_reference(ctx, ref):
    """commit(SHA/GF), commit(ECDSA/Fp256), 16 challenge bytes, prove(SHA), prove(ECDSA) on one
    transcript, coins from one stream: both proofs and the challenge equal the reference's"""
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    circ_a, wit_a = load("sha1_gf128")
    circ_b, wit_b = load("ecdsa1_p256")
    ca, cb = lf.Circuit(ctx, 4, circ_a), lf.Circuit(ctx, 1, circ_b)
    na, nb = ca.info["rng_bytes"], cb.info["rng_bytes"]
    coins = rng_bytes(321, 1 << 20)[:na + nb + 4096].copy()
    want = ref.prove_pair(ref.Circuit(4, circ_a), ref.Circuit(1, circ_b), wit_a, wit_b, coins, tinit=b"pair")
    assert want["rng_used_a"] == na and want["rng_used_total"] == na + nb
    pa, pb = lf.ZkProver(ca), lf.ZkProver(cb)
    Wa, Wb = np.frombuffer(wit_a, np.uint8)[None, :], np.frombuffer(wit_b, np.uint8)[None, :]
    ts = api.transcripts(1, b"pair")
    _, st = pa.commit_batch(Wa, coins[None, :na], ts)
    assert st[0] == 0
    _, st = pb.commit_batch(Wb, coins[None, na:na + nb], ts)
    assert st[0] == 0
    assert api.transcript_challenge(ts[0], 16) == want["challenge"]
    proof_a, st = pa.prove_committed_batch(Wa, ts)
    assert st[0] == 0 and proof_a[0] == want["proof_a"]
    proof_b, st = pb.prove_committed_batch(Wb, ts)
    assert st[0] == 0 and proof_b[0] == want["proof_b"]
