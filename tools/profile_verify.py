"""One verification of a batch of B proofs (lf_zk_verify_batch) between cudaProfilerStart/Stop, after proving
un-profiled:  ncu --profile-from-start off ... python tools/profile_verify.py [B] [circuit]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf  # noqa: E402
from fixtures import load, load_witnesses  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
which = sys.argv[2] if len(sys.argv) > 2 else "sha1_gf128"
fid = lf.FIELD_GF2_128 if "gf128" in which else lf.FIELD_P256
circ, _ = load(which)
Ws = load_witnesses(which)
ctx = lf.Context(0)
c = lf.Circuit(ctx, fid, circ)
W = Ws[np.arange(B) % Ws.shape[0]]
rng = np.random.default_rng(3).integers(0, 256, (B, c.info["rng_bytes"] + 256), dtype=np.uint8)
proofs, st = lf.ZkProver(c).prove_batch(W, rng)
assert (st == 0).all()
npub = c.info["npub_in"] * c.info["kbytes"]
pubs = np.ascontiguousarray(W[:, :npub]) if npub else None
v = lf.ZkVerifier(c)
sv, _ = v.verify_batch(pubs, proofs)
assert (sv == 0).all()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sv, _ = v.verify_batch(pubs, proofs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
assert (sv == 0).all()
print("ok", which, B)
