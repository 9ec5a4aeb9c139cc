"""Device-resident prover latency for small batches (CUDA events), both circuits; prints one JSON line.
Knobs through the environment (LF_LIG_HOST_MAX, LF_HOST_INIT_MAX, LF_SC_CLUSTER, ...)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf  # noqa: E402
from fixtures import load  # noqa: E402

out = dict(env={k: v for k, v in os.environ.items() if k.startswith("LF_")})
# The context runs on a torch stream and the events are recorded on THAT stream: a context created without
# a stream owns a non-blocking one, which events recorded on torch's current stream do not order against
# (such an interval ends before the last step does and reads ~1.7 ms too low).
# LATB_WARM=n: one batch of n proofs first, so that the per-proof buffers are allocated for n (as in bench.py)
xs = torch.cuda.Stream()
ctx = lf.Context(0, stream=xs.cuda_stream)
warm = int(os.environ.get("LATB_WARM", "0"))
out["warm"] = warm
for which in ("sha1_gf128", "ecdsa1_p256"):
    circ, wit = load(which)
    c = lf.Circuit(ctx, lf.FIELD_GF2_128 if "gf128" in which else lf.FIELD_P256, circ)
    p = lf.ZkProver(c)
    info = c.info
    rstride = (info["rng_bytes"] + 8 * info["rng_redraw_bytes"] + 15) & ~15
    Bm = max(8, warm)
    d_wit = torch.from_numpy(np.frombuffer(wit, np.uint8).copy()).repeat(Bm, 1).cuda()
    d_rng = torch.randint(0, 256, (Bm, rstride), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    d_out = torch.empty((Bm, info["max_proof_bytes"]), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(Bm, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(Bm, dtype=torch.int32, device="cuda")
    if warm:
        p.prove_batch_ptr(warm, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(),
                          info["max_proof_bytes"], d_len.data_ptr(), d_st.data_ptr(), device=True)
        torch.cuda.synchronize()
    for B in (1, 4, 8):
        def step():
            p.prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(),
                              info["max_proof_bytes"], d_len.data_ptr(), d_st.data_ptr(), device=True)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(xs)
        for _ in range(10):
            step()
        e1.record(xs)
        torch.cuda.synchronize()
        assert int(d_st[:B].abs().sum().item()) == 0
        out[f"{which}_B{B}_ms"] = e0.elapsed_time(e1) / 10
print(json.dumps(out))
