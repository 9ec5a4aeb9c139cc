import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import longfellow_zk_b200 as lf
from fixtures import load, rng_bytes
circ, wit = load("sha1_gf128")
ctx = lf.Context(0)
c = lf.Circuit(ctx, 4, circ)
p = lf.ZkProver(c)
for B in (1, 3, 2, 8):
    rng = np.stack([rng_bytes(1 + i, c.info["rng_bytes"]) for i in range(B)])
    W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    try:
        proofs, status = p.prove_batch(W, rng)
        print(B, "ok", [len(x) for x in proofs], status)
    except Exception as e:
        print(B, "ERR", e)
