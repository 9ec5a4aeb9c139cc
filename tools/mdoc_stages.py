"""Stage times of the two mdoc circuits proved stand-alone (same witnesses as the mdoc flow)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf
from fixtures import load_mdoc
f = load_mdoc(); e = f["expect"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = lf.Context(0)
sig = lf.Circuit(ctx, 1, f["raw"], rate=e["rate"], nreq=e["nreq"], block_enc=e["block_enc_sig"])
hsh = lf.Circuit(ctx, 4, f["raw"][sig.info["lfc1_bytes"]:], rate=e["rate"], nreq=e["nreq"], block_enc=e["block_enc_hash"])
nh, ns = hsh.info["rng_bytes"], sig.info["rng_bytes"]
for name, c, w, coins in (("hash", hsh, f["w_hash_mac"], f["coins"][:nh]), ("sig", sig, f["w_sig_mac"], f["coins"][nh:nh + ns])):
    p = lf.ZkProver(c)
    W = np.repeat(w[None, :], B, axis=0); R = np.repeat(coins[None, :], B, axis=0)
    p.prove_batch(W, R)
    p.set_profiling(True)
    proofs, st = p.prove_batch(W, R)
    assert (st == 0).all()
    ms = p.stage_ms()
    info = c.info
    print(name, "B", B, {k: round(v, 1) for k, v in ms.items()}, "sum %.1f" % sum(ms.values()),
          "mults/proof: rs %d eval %d sumcheck %d ligero %d" % (info["rs_mults"], info["eval_mults"], info["sumcheck_mults"], info["ligero_mults"]))
