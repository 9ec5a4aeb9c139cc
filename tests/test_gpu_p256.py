"""GPU parity over the P-256 base field (Fp256): Montgomery multiply, RS row
extension, Merkle commit and the whole ECDSA ZK proof (BM_ECDSAZKProver/1),
against the oracle and the reference's golden proofs."""
import hashlib

import numpy as np
import pytest

from fixtures import golden, load, rng_bytes

pytestmark = pytest.mark.gpu

P256 = 1
P = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff


def rand_elts(rs, n):
    v = [int.from_bytes(rs.bytes(32), "little") % P for _ in range(n)]
    return np.frombuffer(b"".join(x.to_bytes(32, "little") for x in v), np.uint8).reshape(n, 32).copy()


def test_fp256_mul_matches_oracle(ctx, oracle):
    rs = np.random.default_rng(21)
    a, b = rand_elts(rs, 5000), rand_elts(rs, 5000)
    a[0] = 0
    a[1] = np.frombuffer((P - 1).to_bytes(32, "little"), np.uint8)
    b[1] = a[1]
    a[2] = 0
    a[2, 0] = 1
    assert (ctx.elt_mul(P256, a, b) == oracle.elt_op(P256, "mul", a, b)).all()


def test_fp256_rejects_non_canonical(ctx):
    import longfellow_zk_b200 as lf
    a = np.full((4, 32), 255, np.uint8)  # 2^256-1 >= p
    with pytest.raises(lf.LongfellowError) as e:
        ctx.elt_mul(P256, a, a)
    assert e.value.code == -3


@pytest.mark.parametrize("n,m", [(455, 4096), (909, 4096), (455, 909), (5, 16), (1, 3), (2, 2), (64, 65)])
def test_fp256_rs_matches_oracle(ctx, oracle, n, m):
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(n + m)
    rows = rand_elts(rs, 2 * m).reshape(2, m, 32)
    got = lf.ReedSolomonFactory(ctx, P256).make(n, m).interpolate(rows)
    assert (got == oracle.rs_interpolate(P256, n, m, rows)).all()


def test_fp256_merkle_matches_oracle(ctx, oracle):
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(8)
    nrow, block_enc, dblock = 11, 4096, 909
    ext = block_enc - dblock
    tab = rand_elts(rs, nrow * block_enc).reshape(nrow, block_enc, 32)
    nonces = rs.integers(0, 256, (ext, 32), dtype=np.uint8)
    root = lf.MerkleCommitment(ctx, P256).commit(tab, nrow, block_enc, dblock, nonces)
    payload = np.ascontiguousarray(tab[:, dblock:, :].transpose(1, 0, 2)).reshape(ext, nrow * 32)
    want_root, _, _ = oracle.merkle_commit_open(payload, nonces.reshape(-1), [])
    assert root == want_root.tobytes()


@pytest.fixture(scope="module")
def ecdsa(ctx):
    import longfellow_zk_b200 as lf
    circ, wit = load("ecdsa1_p256")
    return lf.Circuit(ctx, P256, circ), circ, wit


def test_ecdsa_proof_stages_match_oracle(ecdsa, oracle):
    import longfellow_zk_b200 as lf
    c, circ, wit = ecdsa
    g = golden()["ecdsa1_p256"]
    assert c.info["rng_bytes"] == g["proofs"][0]["rng_used"]
    rng = rng_bytes(1, 1 << 19)[:c.info["rng_bytes"]]
    want = oracle.Circuit(P256, circ).prove(wit, rng, dump=True)
    p = lf.ZkProver(c)
    proofs, status = p.prove_batch(np.frombuffer(wit, np.uint8)[None, :], rng[None, :])
    nw, nrow, be = c.info["nw"], c.info["nrow"], c.info["block_enc"]
    assert (p.debug_fetch(0, 1) == want["witness"][:nw * 32]).all(), "Ligero witness"
    got_t = p.debug_fetch(0, 2).reshape(nrow, be, 32)
    want_t = want["tableau"][:nrow * be * 32].reshape(nrow, be, 32)
    for row in range(nrow):
        assert (got_t[row, :909] == want_t[row, :909]).all(), f"tableau row {row} message part"
        assert (got_t[row] == want_t[row]).all(), f"tableau row {row} extension"
    assert p.debug_fetch(0, 3).tobytes() == want["root"], "Merkle root"
    got_sc = p.debug_fetch(0, 4)
    want_sc = want["sumcheck"][:got_sc.size]
    if not (got_sc == want_sc).all():
        bad = np.nonzero((got_sc != want_sc).reshape(-1, 32).any(axis=1))[0]
        raise AssertionError(f"sumcheck proof differs first at element {bad[0]} of {got_sc.size // 32}")
    assert status[0] == 0
    if proofs[0] != want["proof"]:
        a, b = np.frombuffer(proofs[0], np.uint8), np.frombuffer(want["proof"], np.uint8)
        n = min(a.size, b.size)
        raise AssertionError(f"proof bytes differ first at offset {np.nonzero(a[:n] != b[:n])[0][:1]} "
                             f"(lengths {a.size} vs {b.size})")


def test_ecdsa_proofs_match_reference_golden(ecdsa):
    import longfellow_zk_b200 as lf
    c, circ, wit = ecdsa
    g = golden()["ecdsa1_p256"]
    seeds = [pr["seed"] for pr in g["proofs"]]
    rng = np.stack([rng_bytes(s, 1 << 19)[:c.info["rng_bytes"]] for s in seeds])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], len(seeds), axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    for pr, got, st in zip(g["proofs"], proofs, status):
        assert st == 0
        assert len(got) == pr["proof_len"]
        assert hashlib.sha256(got).hexdigest() == pr["proof_sha256"]


def test_reference_verifier_accepts_gpu_ecdsa_proofs(ecdsa, ref):
    import longfellow_zk_b200 as lf
    c, circ, wit = ecdsa
    B = 4
    rng = np.stack([rng_bytes(300 + i, c.info["rng_bytes"]) for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    rc = ref.Circuit(P256, circ)
    npub = c.info["npub_in"]
    for pr, st in zip(proofs, status):
        assert st == 0
        assert rc.verify(wit[:npub * 32], pr) == 0
