"""GPU parity of the stand-alone prime-field transforms (BASELINE config 1):
FFT<Field>::fftb / fftf and ReedSolomon::interpolate over the reference's
benchmark fields, against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MOD = {
    100: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    101: 2**128 - 2**108 + 1,
    102: 2**64 - 2**32 + 1,
    1: 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff,
}
KB = {100: 32, 101: 16, 102: 8, 1: 32}


def rand_elts(rs, fid, n):
    p, kb = MOD[fid], KB[fid]
    v = [int.from_bytes(rs.bytes(kb + 8), "little") % p for _ in range(n)]
    return np.frombuffer(b"".join(x.to_bytes(kb, "little") for x in v), np.uint8).reshape(n, kb).copy()


@pytest.mark.parametrize("fid", [100, 101, 102, 1])
def test_field_mul_matches_oracle(ctx, oracle, fid):
    rs = np.random.default_rng(fid)
    a, b = rand_elts(rs, fid, 3000), rand_elts(rs, fid, 3000)
    a[0] = 0
    a[1] = np.frombuffer((MOD[fid] - 1).to_bytes(KB[fid], "little"), np.uint8)
    b[1] = a[1]
    assert (ctx.elt_mul(fid, a, b) == oracle.elt_op(fid, "mul", a, b)).all()


@pytest.mark.parametrize("fid", [100, 101, 102])
@pytest.mark.parametrize("n", [1, 2, 8, 1024, 4096, 65536])
def test_fft_matches_oracle(ctx, oracle, fid, n):
    rs = np.random.default_rng(fid * 7 + n)
    a = rand_elts(rs, fid, n)
    for fwd in (False, True):
        assert (ctx.fft(fid, a, n, fwd) == oracle.fft(fid, a, n, fwd)).all(), (fid, n, fwd)


@pytest.mark.parametrize("n", [2, 64, 2048, 65536])
def test_fft_fp2_p256_matches_oracle(ctx, oracle, n):
    rs = np.random.default_rng(n)
    a = rand_elts(rs, 1, 2 * n)
    for fwd in (False, True):
        assert (ctx.fft(1, a, n, fwd) == oracle.fft(1, a, n, fwd)).all(), (n, fwd)


@pytest.mark.parametrize("fid", [100, 101, 102])
def test_fft_roundtrip_and_linearity_at_scale(ctx, fid):
    """fftf(fftb(x)) = n x and linearity at n = 2^18 (no oracle needed)"""
    n = 1 << 18
    rs = np.random.default_rng(fid)
    p, kb = MOD[fid], KB[fid]
    x = rand_elts(rs, fid, n)
    y = ctx.fft(fid, ctx.fft(fid, x, n, False), n, True)
    # n * x mod p on a sample of positions
    for i in list(range(0, n, n // 16)) + [n - 1]:
        xi = int.from_bytes(x[i].tobytes(), "little")
        assert int.from_bytes(y[i].tobytes(), "little") == (xi * n) % p


@pytest.mark.parametrize("fid", [100, 101, 102])
@pytest.mark.parametrize("n,m", [(455, 4096), (5, 16), (1, 3), (300, 301), (4096, 16384)])
def test_rs_matches_oracle(ctx, oracle, fid, n, m):
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(fid + n + m)
    rows = rand_elts(rs, fid, 2 * m).reshape(2, m, KB[fid])
    got = lf.ReedSolomonFactory(ctx, fid).make(n, m).interpolate(rows)
    assert (got == oracle.rs_interpolate(fid, n, m, rows)).all()


@pytest.mark.parametrize("n,m", [(3300, 8000), (5000, 20000)])
def test_rs_p256_large_rows_match_oracle(ctx, oracle, n, m):
    """P-256 rows beyond the shared-memory real-FFT kernel (m > 6400): Fp2 convolution in global memory"""
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(n + m)
    rows = rand_elts(rs, 1, m).reshape(1, m, 32)
    got = lf.ReedSolomonFactory(ctx, 1).make(n, m).interpolate(rows)
    assert (got == oracle.rs_interpolate(1, n, m, rows)).all()


def test_rs_p256_benchmark_shape_is_consistent(ctx):
    """ReedSolomon(65536, 262144) over Fp256 (reed_solomon_test.cc:337-401) is too slow for the
    scalar oracle; check it through the code's defining property instead: the extension of a
    polynomial of degree < n' <= n computed with (n, m) equals the one computed with (n', m)."""
    import longfellow_zk_b200 as lf
    rs = np.random.default_rng(77)
    n, m, small = 65536, 262144, 455
    base = rand_elts(rs, 1, 4096).reshape(1, 4096, 32)
    cw = lf.ReedSolomonFactory(ctx, 1).make(small, 4096).interpolate(base)   # degree < 455 on 0..4095
    cw = lf.ReedSolomonFactory(ctx, 1).make(4096, m).interpolate(
        np.concatenate([cw, np.zeros((1, m - 4096, 32), np.uint8)], axis=1))  # the same polynomial on 0..m-1
    again = cw.copy()
    again[:, n:] = 0
    assert (lf.ReedSolomonFactory(ctx, 1).make(n, m).interpolate(again) == cw).all()


@pytest.mark.parametrize("fid", [100, 101, 102, 1])
def test_rs_config1_batch_of_1024_rows(ctx, oracle, fid):
    """BASELINE config 1's batched shape, 1024 rows of ReedSolomon(455, 4096), over every prime field:
    sampled rows against the oracle, and all rows through linearity (row 2k+1 = row 2k + row 0 on the
    message part => the same on the extension)."""
    import longfellow_zk_b200 as lf
    p, kb = MOD[fid], KB[fid]
    rs = np.random.default_rng(fid)
    n, m, R = 455, 4096, 1024
    rows = np.zeros((R, m, kb), np.uint8)
    half = rand_elts(rs, fid, (R // 2) * n).reshape(R // 2, n, kb)
    rows[0::2, :n] = half
    to_int = lambda a: [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
    r0 = to_int(half[0])
    for k in (0, 1, 100, R // 2 - 1):   # a few odd rows = even row + row 0 (mod p)
        rk = to_int(half[k])
        rows[2 * k + 1, :n] = np.frombuffer(b"".join(((a + b) % p).to_bytes(kb, "little") for a, b in zip(rk, r0)),
                                            np.uint8).reshape(n, kb)
    got = lf.ReedSolomonFactory(ctx, fid).make(n, m).interpolate(rows)
    for i in (0, 1, 2, 201, R - 2):
        want = oracle.rs_interpolate(fid, n, m, rows[i:i + 1].copy())
        assert (got[i] == want[0]).all(), i
    e0 = to_int(got[0])
    for k in (0, 1, 100, R // 2 - 1):
        ek, eo = to_int(got[2 * k]), to_int(got[2 * k + 1])
        assert all((a + b) % p == c for a, b, c in zip(ek, e0, eo)), k


def test_fft_timing_reports(ctx):
    import json
    import os
    res = {}
    for name, fid in [("bn254_fp4", 100), ("fp128", 101), ("goldilocks", 102), ("fp2_p256", 1)]:
        res[f"fft_{name}_65536_ms"] = ctx.fft_time_ms(fid, 65536, reps=20)
    print("FFT_TIMING", json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/fft_timing.json", "w"))
    assert all(v > 0 for v in res.values())
