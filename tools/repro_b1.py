import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf
from fixtures import load, rng_bytes
which = sys.argv[1] if len(sys.argv) > 1 else "sha1_gf128"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fid = 4 if "gf128" in which else 1
circ, wit = load(which)
ctx = lf.Context(0)
c = lf.Circuit(ctx, fid, circ)
rng = np.stack([rng_bytes(1 + i, c.info["rng_bytes"]) for i in range(B)])
W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
pr0 = lf.ZkProver(c)
proofs, status = pr0.prove_batch(W, rng)
dbg = pr0.debug_fetch(0, 99).view(np.int64)
if len(dbg) >= 12:
    print('per round: QW', dbg[8] / dbg[1], 'evaluations+barrier', dbg[9] / dbg[1], 'bind+barrier', dbg[10] / dbg[1],
          'layer set-up total', dbg[7], 'solo rounds', dbg[11])
print('serial cycles', dbg[0], 'rounds', dbg[1], 'cycles/round', dbg[0] / max(dbg[1], 1), 'kernel cycles', dbg[2], 'per round: evals', dbg[4] / dbg[1], 'absorb', dbg[5] / dbg[1], 'challenge', dbg[6] / dbg[1])
print("ok", which, B, [len(p) for p in proofs][:4], status[:4])
pr = lf.ZkProver(c)

