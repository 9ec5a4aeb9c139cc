"""K steps of 1024 SHA-256 proofs issued round-robin on S contexts/streams: proofs/s for S = 1..4."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf
from fixtures import load
which = sys.argv[1] if len(sys.argv) > 1 else "sha1_gf128"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
circ, wit = load(which)
fid = 4 if "gf128" in which else 1
S_MAX = 6
streams = [torch.cuda.Stream() for _ in range(S_MAX)]
ctxs = [lf.Context(0, stream=s.cuda_stream) for s in streams]
provers = [lf.ZkProver(lf.Circuit(c, fid, circ)) for c in ctxs]
info = provers[0].c.info
rstride = (info["rng_bytes"] + 8 * info["rng_redraw_bytes"] + 15) & ~15
d_wit = torch.from_numpy(np.frombuffer(wit, np.uint8).copy()).repeat(B, 1).cuda()
d_rng = torch.randint(0, 256, (B, rstride), dtype=torch.uint8)
d_rng = d_rng.cuda()
outs = [(torch.empty((B, info["max_proof_bytes"]), dtype=torch.uint8, device="cuda"),
         torch.zeros(B, dtype=torch.int64, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"))
        for _ in range(S_MAX)]
def step(i):
    o = outs[i]
    provers[i].prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, o[0].data_ptr(), info["max_proof_bytes"],
                               o[1].data_ptr(), o[2].data_ptr(), device=True)
for i in range(S_MAX):
    step(i); step(i)
torch.cuda.synchronize()
for S in ([int(x) for x in os.environ["LF_STREAMS"].split(",")] if os.environ.get("LF_STREAMS") else (1, 2, 3, 4, 5, 6)):
    K = 12
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(K):
        step(k % S)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(which, {k: v for k, v in os.environ.items() if k.startswith("LF_")}, "streams", S,
          "proofs/s %.0f" % (K * B / dt), "ms/step %.2f" % (1e3 * dt / K))
for o in outs:
    assert int(o[2].abs().sum().item()) == 0
