"""One mdoc verification (hash + signature circuit, batch of B, one transcript per proof) between
cudaProfilerStart/Stop, after proving un-profiled:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/x.csv python tools/profile_mdoc_verify.py [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf  # noqa: E402
from longfellow_zk_b200 import api  # noqa: E402
from fixtures import load_mdoc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
f = load_mdoc()
e = f["expect"]
ctx = lf.Context(0)
sig = lf.Circuit(ctx, lf.FIELD_P256, f["raw"], rate=e["rate"], nreq=e["nreq"], block_enc=e["block_enc_sig"])
hsh = lf.Circuit(ctx, lf.FIELD_GF2_128, f["raw"][sig.info["lfc1_bytes"]:], rate=e["rate"], nreq=e["nreq"],
                 block_enc=e["block_enc_hash"])
nh, ns = hsh.info["rng_bytes"], sig.info["rng_bytes"]
rep = lambda a: np.repeat(a[None, :], B, axis=0)
ph, ps = lf.ZkProver(hsh), lf.ZkProver(sig)
seed = bytes.fromhex(e["transcript"])
ts = api.transcripts(B, seed)
ph.commit_batch(rep(f["w_hash"]), rep(f["coins"][:nh]), ts)
ps.commit_batch(rep(f["w_sig"]), rep(f["coins"][nh:nh + ns]), ts)
for i in range(B):
    api.transcript_challenge(ts[i], 16)
a, st1 = ph.prove_committed_batch(rep(f["w_hash_mac"]), ts)
b, st2 = ps.prove_committed_batch(rep(f["w_sig_mac"]), ts)
assert (st1 == 0).all() and (st2 == 0).all()
pub_h = np.ascontiguousarray(rep(f["w_hash_mac"])[:, :hsh.info["npub_in"] * hsh.info["kbytes"]])
pub_s = np.ascontiguousarray(rep(f["w_sig_mac"])[:, :sig.info["npub_in"] * sig.info["kbytes"]])
vh, vs = lf.ZkVerifier(hsh), lf.ZkVerifier(sig)


def verify():
    tv = api.transcripts(B, seed)
    for i in range(B):
        api.transcript_write(tv[i], a[i][:32])
        api.transcript_write(tv[i], b[i][:32])
        api.transcript_challenge(tv[i], 16)
    s1, _ = vh.verify_batch(pub_h, a, transcripts=tv)
    s2, _ = vs.verify_batch(pub_s, b, transcripts=tv)
    assert (s1 == 0).all() and (s2 == 0).all()


verify()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
verify()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
