"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): prove + verify a few proofs of both
circuits in the cluster, per-proof and flat shapes, RS over every field."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf
from fixtures import load, load_witnesses
ctx = lf.Context(0)
batches = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 9, 210]
for name, fid in (("sha1_gf128", 4), ("ecdsa1_p256", 1)):
    circ, wit = load(name)
    W = load_witnesses(name)
    c = lf.Circuit(ctx, fid, circ)
    for B in batches:
        rng = np.random.default_rng(B).integers(0, 256, (B, c.info["rng_bytes"] + 256), dtype=np.uint8)
        Wb = W[np.arange(B) % W.shape[0]]
        proofs, st = lf.ZkProver(c).prove_batch(Wb, rng)
        assert (st == 0).all()
        npub = c.info["npub_in"] * c.info["kbytes"]
        sv, _ = lf.ZkVerifier(c).verify_batch(np.ascontiguousarray(Wb[:, :npub]) if npub else None, proofs)
        assert (sv == 0).all()
        print(name, B, "ok", flush=True)
# the reference's tiny known-answer circuit (3 terms, block_enc 128): the smallest buffers of every kind
from fixtures import load_rfc_vector
from oracle import portapi as O
rec, circ, wit, coins, want = load_rfc_vector(O)
c = lf.Circuit(ctx, rec["field_id"], circ, rate=rec["rate"], nreq=rec["nreq"], block_enc=rec["block_enc"])
for B in (1, 70):
    Wb = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
    proofs, st = lf.ZkProver(c).prove_batch(Wb, np.repeat(coins[None, :], B, axis=0), tinit=b"test")
    assert (st == 0).all() and proofs[B - 1] == want
    npub = c.info["npub_in"] * c.info["kbytes"]
    sv, _ = lf.ZkVerifier(c).verify_batch(np.ascontiguousarray(Wb[:, :npub]), proofs, tinit=b"test")
    assert (sv == 0).all()
    print("rfc vector", B, "ok", flush=True)
rs = np.random.default_rng(0)
for fid, kb in ((4, 16), (1, 32), (100, 32), (101, 16), (102, 8)):
    rows = rs.integers(0, 256, (3, 4096, kb), dtype=np.uint8)
    rows[:, :, kb - 1] &= 0x0f
    f = lf.LCH14ReedSolomonFactory(ctx) if fid == 4 else lf.ReedSolomonFactory(ctx, fid)
    f.make(455, 4096).interpolate(rows)
    print("rs", fid, "ok", flush=True)
