"""Shared loaders for tests/golden fixtures (no reference needed at run time)."""
import json
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    circ = zlib.decompress(open(os.path.join(GOLDEN, name + ".circuit.z"), "rb").read())
    wit = zlib.decompress(open(os.path.join(GOLDEN, name + ".witness.z"), "rb").read())
    return circ, wit


def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def rng_bytes(seed, n):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def zstd_decompress(data):
    """libzstd.so.1 is in the image (no Python binding, no header): a ctypes call suffices"""
    import ctypes as C
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    cap = z.ZSTD_getFrameContentSize(data, len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_decompress(buf, cap, data, len(data))
    assert not z.ZSTD_isError(n)
    return buf.raw[:n]


def load_mdoc():
    """The frozen mdoc instance of tests/golden/make_golden_mdoc.py: the reference's circuit file of
    kZkSpecs[0] (signature circuit then hash circuit, LFC1), filled witnesses before / after the MAC
    patch, commit coins, and what run_mdoc_prover produced for them."""
    import json
    import zlib
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mdoc")
    rd = lambda n: np.frombuffer(zlib.decompress(open(os.path.join(d, n + ".bin.z"), "rb").read()), np.uint8).copy()
    return dict(raw=zstd_decompress(open(os.path.join(d, "circuit_v7_1attr.zst"), "rb").read()),
                w_sig=rd("w_sig"), w_hash=rd("w_hash"), w_sig_mac=rd("w_sig_mac"), w_hash_mac=rd("w_hash_mac"),
                coins=np.fromfile(os.path.join(d, "coins.bin"), np.uint8),
                expect=json.load(open(os.path.join(d, "expect.json"))))
