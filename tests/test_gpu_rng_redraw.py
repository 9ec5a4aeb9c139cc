"""Caller randomness over a prime field: Field::sample (lib/algebra/fp_generic.h:360-371) draws 32 bytes
from the RandomEngine and draws AGAIN while the value is >= p, so one rejected draw moves every later
sample and the Merkle nonces 32 bytes down the stream.  Same witness + same coin stream must give the
reference's bytes whatever the coins are -- including coins that force redraws (2^-32 per sample for
honest coins, so the cases are constructed)."""
import numpy as np
import pytest

from fixtures import load, rng_bytes

pytestmark = pytest.mark.gpu

P256 = 1


def _coins_with_redraws(info, seed, nslack=16):
    """random coins with slots >= p planted: the first pad sample, two back to back in the pad, one in
    IDOT, one in the blinding of a quadratic row, and the slot where the LAST sample before the nonces
    would have been read"""
    n = info["rng_bytes"]
    ns = info["rng_sample_bytes"] // 32          # samples of one proof
    coins = rng_bytes(seed, 1 << 20)[:n + 32 * nslack].copy()
    quad_first = ns - 3 * info["nqtriples"] * info["r"]   # sample index of the first x-row blinding element
    planted = [0, 5, 6]                                   # stream slots (the pad is sampled first)
    pad = (info["rng_sample_bytes"] // 32) - info["block"] - 2 * info["dblock"] - \
        (info["nrow"] - 3) * info["r"]                   # pad samples = sumcheck proof elements
    planted.append(pad + info["block"] + 10 + len(planted))              # IDOT sample 10
    planted.append(quad_first + 7 + len(planted))                        # a quadratic-row sample
    planted.append(ns - 1 + len(planted))                                # the last sample's first draw
    for s in planted:
        coins[32 * s:32 * s + 32] = 0xFF                                 # 2^256 - 1 >= p
    return coins, planted


def test_redrawn_samples_match_oracle_and_reference(ctx, oracle, ref):
    import longfellow_zk_b200 as lf
    circ, wit = load("ecdsa1_p256")
    c = lf.Circuit(ctx, P256, circ)
    info = c.info
    assert info["rng_redraw_bytes"] == 32 and info["rng_sample_bytes"] + 32 * info["block_ext"] == info["rng_bytes"]
    coins, planted = _coins_with_redraws(info, 901)
    plain = rng_bytes(902, 1 << 20)[:coins.size].copy()
    want = oracle.Circuit(P256, circ).prove(wit, coins)
    assert want["rng_used"] == info["rng_bytes"] + 32 * len(planted)
    want_ref = ref.Circuit(P256, circ).prove(wit, coins)
    assert want_ref["proof"] == want["proof"] and want_ref["rng_used"] == want["rng_used"]
    want_plain = oracle.Circuit(P256, circ).prove(wit, plain)
    assert want_plain["rng_used"] == info["rng_bytes"]
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], 3, axis=0)
    p = lf.ZkProver(c)
    proofs, status = p.prove_batch(W, np.stack([plain, coins, plain]))
    assert (status == 0).all()
    assert proofs[1] == want["proof"], "proof with redrawn samples differs from the reference"
    assert proofs[0] == want_plain["proof"] and proofs[2] == want_plain["proof"]
    assert p.rng_consumed(1) == want["rng_used"] and p.rng_consumed(0) == info["rng_bytes"]
    # the same stream cut one byte short of what the redraws need: stream too short, the others unaffected
    short = want["rng_used"] - 1
    proofs, status = p.prove_batch(W, np.stack([plain[:short], coins[:short], plain[:short]]))
    assert list(status) == [0, -6, 0] and proofs[1] == b"" and proofs[0] == want_plain["proof"]
    # exactly enough
    proofs, status = p.prove_batch(W[:1], coins[None, :want["rng_used"]])
    assert status[0] == 0 and proofs[0] == want["proof"]


def test_redraws_in_commit_then_prove(ctx, oracle):
    """the reject list found at commit time also places the nonces that the opening reads at prove time"""
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    circ, wit = load("ecdsa1_p256")
    c = lf.Circuit(ctx, P256, circ)
    coins, planted = _coins_with_redraws(c.info, 903)
    want = oracle.Circuit(P256, circ).prove(wit, coins, tinit=b"split")["proof"]
    W = np.frombuffer(wit, np.uint8)[None, :]
    p = lf.ZkProver(c)
    ts = api.transcripts(1, b"split")
    _, st = p.commit_batch(W, coins[None, :], ts)
    assert st[0] == 0
    got, st = p.prove_committed_batch(W, ts)
    assert st[0] == 0 and got[0] == want


def test_a_stream_of_rejected_draws_is_too_short(ctx):
    """coins that never yield a sample: the reference would loop until its RandomEngine runs dry"""
    import longfellow_zk_b200 as lf
    circ, wit = load("ecdsa1_p256")
    c = lf.Circuit(ctx, P256, circ)
    coins = np.full(c.info["rng_bytes"] + 4096, 0xFF, np.uint8)
    proofs, status = lf.ZkProver(c).prove_batch(np.frombuffer(wit, np.uint8)[None, :], coins[None, :])
    assert status[0] == -6 and proofs[0] == b""


def test_stale_commit_is_refused(ctx):
    """commit(A), then an unrelated prove_batch on the same circuit object overwrites the committed
    tableaux and coins: the prove of A must be refused, not emit proofs from the other batch's state"""
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    circ, wit = load("sha1_gf128")
    c = lf.Circuit(ctx, 4, circ)
    n = c.info["rng_bytes"]
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], 2, axis=0)
    p = lf.ZkProver(c)
    ts = api.transcripts(2, b"stale")
    _, st = p.commit_batch(W, np.stack([rng_bytes(1, n), rng_bytes(2, n)]), ts)
    assert (st == 0).all()
    p.prove_batch(W, np.stack([rng_bytes(3, n), rng_bytes(4, n)]))
    with pytest.raises(lf.LongfellowError) as e:
        p.prove_committed_batch(W, ts)
    assert e.value.code == -1
