"""GPU parity: GF(2^128) multiply, LCH14 Reed-Solomon row extension and the
SHA-256 Merkle column commitment, through the C ABI, against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GF = 4


def test_gf128_mul_matches_oracle(ctx, oracle):
    rs = np.random.default_rng(11)
    n = 1 << 16
    a = rs.integers(0, 256, (n, 16), dtype=np.uint8)
    b = rs.integers(0, 256, (n, 16), dtype=np.uint8)
    a[0] = 0
    b[1] = 0
    a[2] = 255
    b[2] = 255
    a[3] = 0
    a[3, 0] = 1  # one
    got = ctx.elt_mul(GF, a, b)
    assert (got == oracle.gf128_mul(a, b)).all()


@pytest.mark.parametrize("n,m", [(455, 4096), (909, 4096), (455, 909), (461, 4151), (921, 4151),
                                 (1, 5), (2, 3), (7, 8), (8, 8), (300, 301), (512, 2048), (513, 1025),
                                 (4096, 16384)])
def test_rs_interpolate_matches_oracle(ctx, oracle, n, m):
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(n * 7 + m)
    rows = rs.integers(0, 256, (3, m, 16), dtype=np.uint8)
    got = lf.LCH14ReedSolomonFactory(ctx).make(n, m).interpolate(rows)
    want = oracle.lch14_interpolate(n, m, rows)
    assert (got[:, :n] == rows[:, :n]).all()
    assert (got == want).all()


@pytest.mark.parametrize("n,m", [(16384, 65536), (5000, 20000), (8192, 8200)])
def test_rs_large_rows_match_oracle(ctx, oracle, n, m):
    """rows whose first coset does not fit in shared memory take the one-launch-per-step
    path (lch14_reed_solomon_test.cc:72-107 benchmarks 16384 -> 65536)"""
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(n + m)
    rows = rs.integers(0, 256, (2, m, 16), dtype=np.uint8)
    got = lf.LCH14ReedSolomonFactory(ctx).make(n, m).interpolate(rows)
    assert (got == oracle.lch14_interpolate(n, m, rows)).all()


@pytest.mark.parametrize("l", [0, 1, 5, 10, 16])
def test_lch14_fft_matches_oracle(ctx, oracle, l):
    """LCH14::FFT(l, 0, B) and ::IFFT through lf_fft (lch14.h:106-146; lch14_test.cc:199-243 runs l = 16)"""
    rs = np.random.default_rng(100 + l)
    n = 1 << l
    B = rs.integers(0, 256, (n, 16), dtype=np.uint8)
    ev = ctx.fft(GF, B, n, True)
    assert (ev == oracle.lch14("fft", l, 0, B)).all()
    assert (ctx.fft(GF, B, n, False) == oracle.lch14("ifft", l, 0, B)).all()
    assert (ctx.fft(GF, ev, n, False) == B).all()  # IFFT o FFT = id


def test_rs_is_linear_and_idempotent_at_scale(ctx):
    """size-independent properties on a production-sized batch (1024 rows)."""
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(5)
    n, m, R = 455, 4096, 1024
    f = lf.LCH14ReedSolomonFactory(ctx).make(n, m)
    x = np.zeros((R, m, 16), np.uint8)
    y = np.zeros((R, m, 16), np.uint8)
    x[:, :n] = rs.integers(0, 256, (R, n, 16), dtype=np.uint8)
    y[:, :n] = rs.integers(0, 256, (R, n, 16), dtype=np.uint8)
    ex, ey, exy = f.interpolate(x), f.interpolate(y), f.interpolate(x ^ y)
    assert (exy == (ex ^ ey)).all()          # GF(2)-linear
    assert (f.interpolate(ex) == ex).all()   # re-encoding a codeword's message is a no-op
    # a codeword restricted to ANY n positions determines it: re-derive from a shifted window
    g = lf.LCH14ReedSolomonFactory(ctx).make(512, m)
    assert (g.interpolate(ex[:8]) == ex[:8]).all()  # degree < 455 < 512


@pytest.mark.parametrize("nrow,block_enc,dblock", [(20, 4096, 909), (11, 4096, 909), (1, 40, 9),
                                                    (3, 12, 9), (266, 4151, 921), (2, 11, 10)])
def test_merkle_commit_matches_oracle(ctx, oracle, nrow, block_enc, dblock):
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(nrow + block_enc)
    ext = block_enc - dblock
    tab = rs.integers(0, 256, (nrow, block_enc, 16), dtype=np.uint8)
    nonces = rs.integers(0, 256, (ext, 32), dtype=np.uint8)
    root, nodes = lf.MerkleCommitment(ctx).commit(tab, nrow, block_enc, dblock, nonces, want_nodes=True)
    payload = np.ascontiguousarray(tab[:, dblock:, :].transpose(1, 0, 2)).reshape(ext, nrow * 16)
    want_root, _, _ = oracle.merkle_commit_open(payload, nonces.reshape(-1), [])
    # leaves
    import hashlib
    for j in (0, ext - 1, ext // 2):
        leaf = hashlib.sha256(nonces[j].tobytes() + payload[j].tobytes()).digest()
        assert nodes[ext + j].tobytes() == leaf
    assert root == want_root.tobytes()
    if ext > 1:
        _, want_nodes = oracle.merkle_build(nodes[ext:])
        assert (nodes[1:] == want_nodes[1:]).all()


def test_microbench_reports_integer_peaks(ctx):
    import json
    import os
    res = {name: ctx.microbench(i) for i, name in enumerate(
        ["imad_wide_gops", "lop3_gops", "gf128_mul_gops", "sha256_compress_gops"])}
    print("MICROBENCH", json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/microbench.json", "w") as f:
        json.dump(res, f)
    assert all(v > 0 for v in res.values())
