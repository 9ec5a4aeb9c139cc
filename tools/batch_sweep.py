"""Whole-prover time per batch and per proof over the batch size (device resident, one stream)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf
from fixtures import load
res = {}
for which, fid in (("sha1_gf128", 4), ("ecdsa1_p256", 1)):
    circ, wit = load(which)
    stream = torch.cuda.Stream()
    ctx = lf.Context(0, stream=stream.cuda_stream)
    p = lf.ZkProver(lf.Circuit(ctx, fid, circ))
    info = p.c.info
    rstride = (info["rng_bytes"] + 15) & ~15
    BMAX = 2048
    d_wit = torch.from_numpy(np.frombuffer(wit, np.uint8).copy()).repeat(BMAX, 1).cuda()
    d_rng = torch.randint(0, 256, (BMAX, rstride), dtype=torch.uint8)
    d_rng = d_rng.cuda()
    d_out = torch.empty((BMAX, info["max_proof_bytes"]), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(BMAX, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(BMAX, dtype=torch.int32, device="cuda")
    rows = []
    BS = [int(x) for x in os.environ["LF_SWEEP"].split(",")] if os.environ.get("LF_SWEEP") else \
        [1, 2, 4, 9, 10, 16, 32, 64, 128, 147, 148, 192, 256, 384, 512, 768, 1024, 1536, 2048]
    for B in [2048] + BS:
        def step():
            p.prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), info["max_proof_bytes"],
                              d_len.data_ptr(), d_st.data_ptr(), device=True)
        step(); step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record(stream)
        for _ in range(n):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        assert int(d_st[:B].abs().sum().item()) == 0
        rows.append(dict(batch=B, ms_per_batch=round(ms, 3), ms_per_proof=round(ms / B, 4), proofs_per_s=round(B / ms * 1e3)))
        print(which, rows[-1], flush=True)
    res[which] = rows[1:]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "batch_sweep.json"), "w"), indent=1)
