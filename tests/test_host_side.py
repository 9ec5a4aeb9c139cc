"""CPU-only checks of the product's host side: the C-ABI library loads and
exports every symbol include/longfellow_b200.h declares, fails loudly without a
GPU, and the header-only device primitives (compiled here as plain C++) agree
with the oracle."""
import ctypes as C
import hashlib
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from longfellow_zk_b200 import _native, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "longfellow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(lf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = C.CDLL(_native.LIB)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_native.EXPORTS) == declared
    assert b"sm_100a" in _native.lib().lf_version()


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import longfellow_zk_b200 as lf
    with pytest.raises(lf.LongfellowError) as e:
        lf.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


@pytest.fixture(scope="module")
def prim():
    exe = "/tmp/lf_prim_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/prim_check.cc")])
    return lambda mode, data: subprocess.run([exe, mode], input=data, capture_output=True, check=True).stdout


def test_device_gf128_mul_source_on_host(prim, oracle):
    rs = np.random.default_rng(5)
    a = rs.integers(0, 256, (4000, 16), dtype=np.uint8)
    b = rs.integers(0, 256, (4000, 16), dtype=np.uint8)
    a[0] = 0
    a[1] = 255
    b[1] = 255
    assert prim("gfmul", np.concatenate([a, b], axis=1).tobytes()) == oracle.gf128_mul(a, b).tobytes()
    assert prim("gfinv", a[2:40].tobytes()) == oracle.gf128_invert(a[2:40]).tobytes()


def test_device_sha_aes_transcript_source_on_host(prim, oracle):
    rs = np.random.default_rng(6)
    for n in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 128, 1000]:
        d = rs.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert prim("sha", d) == hashlib.sha256(d).digest(), n
    key = rs.integers(0, 256, 32, dtype=np.uint8).tobytes()
    blocks = rs.integers(0, 256, 160, dtype=np.uint8).tobytes()
    assert prim("aes", key + blocks) == oracle.aes256_ecb(key, blocks)
    u32 = lambda v: struct.pack("<I", v)
    s = b"B" + u32(5) + b"hello" + b"R" + u32(40) + b"E" + bytes(range(16)) + b"G" + u32(3)
    s += b"Z" + u32(1000) + b"N" + u32(3187) + b"N" + u32(256) + b"N" + u32(1)
    s += b"A" + u32(2) + bytes(range(32)) + b"Z" + u32(155197) + b"R" + u32(17) + b"Z" + u32(3) + b"G" + u32(1)
    assert prim("transcript", s) == oracle.transcript_script(b"test", s)
    # the single-pass element writes (Sha256::put_stream_words) at every byte alignment and across block
    # boundaries: runs of tagged elements and arrays separated by byte strings of every length mod 64
    for kb, mode, fid in ((16, "transcript", 4), (32, "transcript8", 1)):
        s = b""
        for i in range(140):
            e = lambda: rs.integers(0, 256, kb - 1, dtype=np.uint8).tobytes() + b"\x00"   # canonical for P-256 too
            s += b"B" + u32(i % 67) + rs.integers(0, 256, i % 67, dtype=np.uint8).tobytes()
            s += b"E" + e() + b"E" + e()
            if i % 3 == 0:
                s += b"A" + u32(2) + e() + e()
            if i % 5 == 0:
                s += b"E" + e() + b"E" + e() + b"E" + e()
            s += b"R" + u32(16 if i % 4 else 35)
        assert prim(mode, s) == oracle.transcript_script(b"test", s, fid=fid), mode


def test_oracle_matches_unmodified_reference(oracle, ref):
    """skipped on boxes where oracle/_ref was not built"""
    rs = np.random.default_rng(7)
    a = rs.integers(0, 256, (500, 16), dtype=np.uint8)
    b = rs.integers(0, 256, (500, 16), dtype=np.uint8)
    assert (ref.gf128_mul(a, b) == oracle.gf128_mul(a, b)).all()
    assert (ref.lch14_what() == oracle.lch14_what()).all()
    for x, y in zip(ref.gf128_constants(), oracle.gf128_constants()):
        assert (x == y).all()
    for l in (1, 4, 9):
        B = rs.integers(0, 256, (1 << l, 16), dtype=np.uint8)
        for op, ck in [("fft", 0), ("fft", 5 << l), ("ifft", 3 << l), ("bidir", 1), ("bidir", (1 << l) - 1)]:
            assert (ref.lch14(op, l, ck, B) == oracle.lch14(op, l, ck, B)).all()
    for n, m in [(455, 4096), (455, 909), (2, 3)]:
        rows = rs.integers(0, 256, (2, m, 16), dtype=np.uint8)
        assert (ref.lch14_interpolate(n, m, rows) == oracle.lch14_interpolate(n, m, rows)).all()
    for nw, nq in [(4368, 13), (1495, 11), (87000, 17), (300000, 30000)]:
        for fid in (4, 1):
            assert ref.ligero_param(fid, nw, nq) == oracle.ligero_param(fid, nw, nq)
    for fid in (oracle.FID_BN254, oracle.FID_FP128, oracle.FID_GOLDILOCKS, oracle.P256_ID):
        kb = oracle.KBYTES[fid]
        x = rs.integers(0, 256, (300, kb), dtype=np.uint8)
        y = rs.integers(0, 256, (300, kb), dtype=np.uint8)
        x[:, -1] &= 0x1f
        y[:, -1] &= 0x1f
        assert (ref.fp_mul(fid, x, y) == oracle.elt_op(fid, "mul", x, y)).all()
        rows = rs.integers(0, 256, (2, 64, kb), dtype=np.uint8)
        rows[:, :, -1] &= 0x1f
        assert (ref.rs(fid, rows, 20, 64) == oracle.rs_interpolate(fid, 20, 64, rows)).all()
    for fid in (oracle.FID_BN254, oracle.FID_FP128, oracle.FID_GOLDILOCKS):
        kb = oracle.KBYTES[fid]
        x = rs.integers(0, 256, (256, kb), dtype=np.uint8)
        x[:, -1] &= 0x1f
        for fwd in (False, True):
            assert (ref.fft(fid, x, 256, fwd) == oracle.fft(fid, x, 256, fwd)).all()


def test_host_transcript_helpers_match_oracle(oracle):
    """lf_transcript_init / write_bytes / challenge_bytes (host side of the commit/prove split):
    the state crosses the ABI between every operation, including in the middle of a challenge block"""
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    u32 = lambda v: struct.pack("<I", v)
    ts = api.transcripts(1, b"test")[0]
    got = b""
    api.transcript_write(ts, b"hello")
    got += api.transcript_challenge(ts, 40)
    got += api.transcript_challenge(ts, 5)     # continues inside the third AES block
    got += api.transcript_challenge(ts, 20)
    api.transcript_write(ts, bytes(range(100)))
    got += api.transcript_challenge(ts, 16)
    api.transcript_write(ts, b"")
    got += api.transcript_challenge(ts, 1)
    s = (b"B" + u32(5) + b"hello" + b"R" + u32(40) + b"R" + u32(5) + b"R" + u32(20) + b"B" + u32(100) +
         bytes(range(100)) + b"R" + u32(16) + b"B" + u32(0) + b"R" + u32(1))
    assert got == oracle.transcript_script(b"test", s)


def test_reference_arm_prints_the_contract_line(ref):
    """`bench.py --impl reference` (the unmodified reference on the host cores, no GPU involved) prints ONE JSON
    line on stdout with the keys of the bench contract: metric / unit / value of the headline workload,
    `impl: reference`, a `cpu_baseline` describing the run and an `e2e` object with zero copy bytes."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, check=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sha256_zk_prover_throughput" and d["unit"] == "proofs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "BM_ShaZK_fp2_128/1" in d["config"]["workload"]
