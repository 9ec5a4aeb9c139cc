"""Freeze one instance of the reference's mdoc benchmark claim (mdoc_tests[0], age_over_18, kZkSpecs[0])
as fixtures: the filled witnesses before and after the MAC patch, the commit coins and the hash of
run_mdoc_prover's output for them.  Run in the container that has /root/reference (needs
oracle/_ref/libref_mdoc.so):   python tests/golden/make_golden_mdoc.py"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refapi as ref  # noqa: E402

D = os.path.join(HERE, "mdoc")
# the reference's circuit file of kZkSpecs[0] (zstd) -> LFC1 bytes -> xz fixture
import lzma  # noqa: E402
SRC = "/root/reference/lib/circuits/mdoc/circuits/8d079211715200ff06c5109639245502bfe94aa869908d31176aae4016182121"
raw = ref.zstd_decompress(open(SRC, "rb").read())
open(os.path.join(D, "circuits_v7_1attr.lfc1.xz"), "wb").write(lzma.compress(raw, preset=6))
m = ref.MdocCase(raw)
coins = np.random.default_rng(20261018).integers(0, 256, 1 << 20, dtype=np.uint8)
want = m.prove(coins)
coins = coins[:want["coins_total"]]
ws, wh = m.witnesses()
ws2, wh2, macs = m.update_macs(want["av"])
for name, arr in (("w_sig", ws), ("w_hash", wh), ("w_sig_mac", ws2), ("w_hash_mac", wh2)):
    open(os.path.join(D, name + ".bin.z"), "wb").write(zlib.compress(arr.tobytes(), 9))
open(os.path.join(D, "coins.bin"), "wb").write(coins.tobytes())
p = want["proof"]
json.dump(dict(transcript=m.transcript.hex(), version=m.version, rate=m.rate, nreq=m.nreq,
               block_enc_sig=m.block_enc_sig, block_enc_hash=m.block_enc_hash, av=want["av"].hex(),
               macs=macs.hex(), coins_hash=want["coins_hash"], coins_total=want["coins_total"],
               len_hash=want["len_hash"], len_sig=want["len_sig"], proof_len=len(p),
               proof_sha256=hashlib.sha256(p).hexdigest(),
               hash_proof_sha256=hashlib.sha256(p[96:96 + want["len_hash"]]).hexdigest(),
               sig_proof_sha256=hashlib.sha256(p[96 + want["len_hash"]:]).hexdigest()),
          open(os.path.join(D, "expect.json"), "w"), indent=1)
print("wrote", os.listdir(D))
