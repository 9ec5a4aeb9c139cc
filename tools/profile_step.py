"""One prover step (B proofs of the SHA-256 circuit, device resident) and exit:
the command that is run under ncu (profiles/README.md)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf  # noqa: E402
from fixtures import load  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
which = sys.argv[3] if len(sys.argv) > 3 else "sha1_gf128"
circ, wit = load(which)
ctx = lf.Context(0)
c = lf.Circuit(ctx, lf.FIELD_GF2_128 if "gf128" in which else lf.FIELD_P256, circ)
p = lf.ZkProver(c)
info = c.info
rstride = (info["rng_bytes"] + 15) & ~15
d_wit = torch.from_numpy(np.frombuffer(wit, np.uint8).copy()).repeat(B, 1).cuda()
d_rng = torch.randint(0, 256, (B, rstride), dtype=torch.uint8).cuda()
d_out = torch.empty((B, info["max_proof_bytes"]), dtype=torch.uint8, device="cuda")
d_len = torch.zeros(B, dtype=torch.int64, device="cuda")
d_st = torch.zeros(B, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
if os.environ.get("LF_STAGES"):
    p.set_profiling(True)
import time
t0 = time.time()
for _ in range(steps):
    p.prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), info["max_proof_bytes"],
                      d_len.data_ptr(), d_st.data_ptr(), device=True)
ctx.synchronize()
dt = time.time() - t0
assert int(d_st.abs().sum().item()) == 0, d_st[:8]
print("ok", which, B, int(d_len[0].item()), "wall ms/step %.2f  proofs/s %.0f" % (1e3 * dt / steps, B * steps / dt))
if os.environ.get("LF_STAGES"):
    print({k: round(v, 3) for k, v in p.stage_ms().items()})
