"""Error behaviour of the C ABI: the reference aborts on invariant violations (lib/util/panic.h:27-36)
and returns false for an unsatisfied witness; here every such case is a status code, never a crash and
never a silently wrong proof."""
import numpy as np
import pytest

from fixtures import load, rng_bytes

pytestmark = pytest.mark.gpu


def test_malformed_circuits_are_rejected(ctx):
    import longfellow_zk_b200 as lf
    circ, _ = load("sha1_gf128")
    for bad in (b"", circ[:100], b"\x02" + circ[1:], circ[:-1], bytes(len(circ))):
        with pytest.raises(lf.LongfellowError) as e:
            lf.Circuit(ctx, 4, bad)
        assert e.value.code == -3, (len(bad), e.value)            # LF_ERR_FORMAT
    # like CircuitReader::from_bytes (proto/circuit_reader.h:83-140) the parser consumes one circuit and
    # leaves what follows alone (the mdoc file holds two circuits back to back)
    assert lf.Circuit(ctx, 4, circ + b"\x00" * 7).info["nterms"] == 155197
    with pytest.raises(lf.LongfellowError) as e:
        lf.Circuit(ctx, 1, circ)                                  # field id in the file is 4, not 1
    assert e.value.code == -3
    with pytest.raises(lf.LongfellowError) as e:
        lf.Circuit(ctx, 101, circ)                                # no ZK path over Fp128
    assert e.value.code == -4                                     # LF_ERR_UNSUPPORTED


def test_short_buffers_are_rejected(ctx):
    import ctypes as C
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import _native
    circ, wit = load("sha1_gf128")
    c = lf.Circuit(ctx, 4, circ)
    info = c.info
    W = np.frombuffer(wit, np.uint8).copy()
    out = np.zeros(info["max_proof_bytes"], np.uint8)
    lens, st = np.zeros(1, np.uint64), np.zeros(1, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = rng_bytes(3, info["rng_bytes"])
    call = lambda rstride, ostride: _native.lib().lf_zk_prove_batch(c._h, 1, p(W), p(rng), rstride, b"test", 4, p(out),
                                                                    ostride, p(lens), p(st))
    assert call(info["rng_bytes"] - 1, info["max_proof_bytes"]) == -6      # LF_ERR_RNG
    assert call(info["rng_bytes"], info["max_proof_bytes"] - 16) == -7     # LF_ERR_CAPACITY
    assert call(info["rng_bytes"], info["max_proof_bytes"]) == 0 and st[0] == 0 and lens[0] > 100000


def test_non_canonical_witness_and_coins_are_flagged_per_proof(ctx, oracle):
    """Fp256: a witness element >= p is LF_ERR_FORMAT for that proof (the reference's of_bytes_field
    fails); a caller-random slot >= p is drawn again like the reference's Field::sample does, which this
    proof's stream of exactly rng_bytes has no room for: LF_ERR_RNG (stream too short); the other proofs
    of the batch are unaffected and still equal the oracle's."""
    import longfellow_zk_b200 as lf
    circ, wit = load("ecdsa1_p256")
    c = lf.Circuit(ctx, 1, circ)
    n = c.info["rng_bytes"]
    good = rng_bytes(50, 1 << 19)[:n].copy()
    want = oracle.Circuit(1, circ).prove(wit, good)["proof"]
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], 3, axis=0).copy()
    W[1, 32 * 10:32 * 11] = 0xFF                         # input 10 of proof 1 = 2^256 - 1 >= p
    rng = np.stack([good, good, good]).copy()
    rng[2, 0:32] = 0xFF                                   # first sampled element of proof 2 >= p
    proofs, status = lf.ZkProver(c).prove_batch(W, rng)
    assert status[0] == 0 and proofs[0] == want
    assert status[1] == -3 and proofs[1] == b""
    assert status[2] == -6 and proofs[2] == b""


def test_circuit_id_is_recomputed(ctx):
    """CircuitReader::from_bytes(enforce_circuit_id = true): the id the reference's compiler stored in
    the fixture equals circuit_id (lib/sumcheck/circuit_id.h:30-67) recomputed from the parsed terms,
    for both fields; a file with another trailing id still parses (the reference's default) but
    verify_id refuses it"""
    import longfellow_zk_b200 as lf
    for name, fid in (("sha1_gf128", 4), ("ecdsa1_p256", 1)):
        circ, _ = load(name)
        c = lf.Circuit(ctx, fid, circ)
        assert c.verify_id() == circ[c.info["lfc1_bytes"] - 32:c.info["lfc1_bytes"]]
        bad = bytearray(circ)
        bad[c.info["lfc1_bytes"] - 1] ^= 1
        cb = lf.Circuit(ctx, fid, bytes(bad))
        with pytest.raises(lf.LongfellowError) as e:
            cb.verify_id()
        assert e.value.code == -3


def test_two_devices_in_one_process(oracle):
    """lf_ctx_create(device, ...) on two GPUs of one process: kernel attributes (dynamic shared memory,
    cluster size) are per device, so every launch shape must work on the second device too (batch of one
    = 16-CTA clusters, RS rows above 48 KB of shared memory, the 64 KB FFT tiles)"""
    import numpy as np
    import torch
    import longfellow_zk_b200 as lf
    from fixtures import load, rng_bytes
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in the box")
    circ, wit = load("sha1_gf128")
    want = None
    for dev in (0, 1, 0):
        cx = lf.Context(dev)
        c = lf.Circuit(cx, lf.FIELD_GF2_128, circ)
        coins = rng_bytes(3, c.info["rng_bytes"])
        if want is None:
            want = oracle.Circuit(4, circ).prove(wit, coins)["proof"]
        for B in (1, 40, 300):
            proofs, status = lf.ZkProver(c).prove_batch(np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0),
                                                        np.repeat(coins[None, :], B, axis=0))
            assert (status == 0).all() and proofs[0] == want and proofs[-1] == want, (dev, B)
        rs = np.random.default_rng(0)
        rows = rs.integers(0, 256, (2, 4096, 16), dtype=np.uint8)
        got = lf.LCH14ReedSolomonFactory(cx).make(2000, 4096).interpolate(rows)
        assert (got == oracle.lch14_interpolate(2000, 4096, rows)).all()
        x = rs.integers(0, 256, (65536, 8), dtype=np.uint8)
        x[:, 7] &= 0x7f
        assert (cx.fft(lf.FIELD_GOLDILOCKS, cx.fft(lf.FIELD_GOLDILOCKS, x, 65536), 65536, forward=True).shape == x.shape)
        c.close()
        cx.close()
