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


def load_witnesses(name):
    """several distinct satisfying witnesses of a 1-instance circuit (make_golden.py distinct_witnesses):
    an (N, witness_bytes) uint8 array"""
    _, wit = load(name)
    raw = zlib.decompress(open(os.path.join(GOLDEN, name + ".witnesses.z"), "rb").read())
    return np.frombuffer(raw, np.uint8).reshape(-1, len(wit)).copy()


def load_size(name):
    """(circuit, witness, golden record) of one of the other published instance sizes (tests/golden/sizes)"""
    import lzma
    d = os.path.join(GOLDEN, "sizes")
    circ = lzma.decompress(open(os.path.join(d, name + ".circuit.xz"), "rb").read())
    wit = zlib.decompress(open(os.path.join(d, name + ".witness.z"), "rb").read())
    return circ, wit, json.load(open(os.path.join(d, "golden_sizes.json")))[name]


SIZE_NAMES = ["sha2_gf128", "sha4_gf128", "sha8_gf128", "sha16_gf128", "sha32_gf128", "sha33_gf128", "ecdsa2_p256",
              "ecdsa3_p256"]


def load_mdoc():
    """The frozen mdoc instance of tests/golden/make_golden_mdoc.py: the two circuits of kZkSpecs[0]
    (signature circuit then hash circuit, LFC1 bytes as the reference's circuit file holds them,
    re-compressed with xz), filled witnesses before / after the MAC patch, commit coins, and what
    run_mdoc_prover produced for them."""
    import json
    import lzma
    import zlib
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mdoc")
    rd = lambda n: np.frombuffer(zlib.decompress(open(os.path.join(d, n + ".bin.z"), "rb").read()), np.uint8).copy()
    return dict(raw=lzma.decompress(open(os.path.join(d, "circuits_v7_1attr.lfc1.xz"), "rb").read()),
                w_sig=rd("w_sig"), w_hash=rd("w_hash"), w_sig_mac=rd("w_sig_mac"), w_hash_mac=rd("w_hash_mac"),
                coins=np.fromfile(os.path.join(d, "coins.bin"), np.uint8),
                expect=json.load(open(os.path.join(d, "expect.json"))))
