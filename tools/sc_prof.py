"""Sumcheck stage time and the per-proof kernel's cycle counters (LF_PROF words parked in the HQuad
buffer by the last launch: [0] cycles in serial rounds, [1] rounds, [2] kernel cycles, [4..6] round
polynomial / element writes / challenge) for one configuration of the LF_SC_* knobs (environment)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf  # noqa: E402
from fixtures import load  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
which = sys.argv[2] if len(sys.argv) > 2 else "sha1_gf128"
circ, wit = load(which)
ctx = lf.Context(0)
c = lf.Circuit(ctx, lf.FIELD_GF2_128 if "gf128" in which else lf.FIELD_P256, circ)
p = lf.ZkProver(c)
info = c.info
rstride = (info["rng_bytes"] + 8 * info["rng_redraw_bytes"] + 15) & ~15
d_wit = torch.from_numpy(np.frombuffer(wit, np.uint8).copy()).repeat(B, 1).cuda()
d_rng = torch.randint(0, 256, (B, rstride), dtype=torch.uint8).cuda()
d_out = torch.empty((B, info["max_proof_bytes"]), dtype=torch.uint8, device="cuda")
d_len = torch.zeros(B, dtype=torch.int64, device="cuda")
d_st = torch.zeros(B, dtype=torch.int32, device="cuda")


def step():
    p.prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), info["max_proof_bytes"],
                      d_len.data_ptr(), d_st.data_ptr(), device=True)


for _ in range(2):
    step()
ctx.synchronize()
p.set_profiling(True)
acc = {}
for _ in range(3):
    step()
    for k, v in p.stage_ms().items():
        acc[k] = acc.get(k, 0.0) + v / 3
assert int(d_st.abs().sum().item()) == 0
prof = np.frombuffer(p.debug_fetch(0, 99).tobytes(), np.int64)
env = {k: v for k, v in os.environ.items() if k.startswith("LF_")}
l0 = ctx.launch_count
step()
ctx.synchronize()
print(which, B, env, "sumcheck %.2f ms  total %.2f ms  launches/step %d" % (acc["sumcheck"], sum(acc.values()),
                                                                        ctx.launch_count - l0),
      "prof", prof.tolist())
