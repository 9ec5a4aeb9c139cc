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
