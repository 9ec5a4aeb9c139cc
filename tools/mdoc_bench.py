"""mdoc proof (BASELINE config 4) on the GPU: commit(hash) + commit(sig) + prove(hash) + prove(sig) through
the host-pointer C ABI on the frozen instance of tests/golden/mdoc, for a batch of B identical instances."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf
from longfellow_zk_b200 import api
from fixtures import load_mdoc


def measure(batches=(1, 8, 32), reps=3, device=0, verify_max_batch=8):
    f = load_mdoc(); e = f["expect"]
    ctx = lf.Context(device)
    t0 = time.perf_counter()
    sig = lf.Circuit(ctx, lf.FIELD_P256, f["raw"], rate=e["rate"], nreq=e["nreq"], block_enc=e["block_enc_sig"])
    hsh = lf.Circuit(ctx, lf.FIELD_GF2_128, f["raw"][sig.info["lfc1_bytes"]:], rate=e["rate"], nreq=e["nreq"],
                     block_enc=e["block_enc_hash"])
    upload_s = time.perf_counter() - t0
    nh, ns = hsh.info["rng_bytes"], sig.info["rng_bytes"]
    ph, ps = lf.ZkProver(hsh), lf.ZkProver(sig)
    res = dict(upload_s=upload_s, hash_terms=hsh.info["nterms"], sig_terms=sig.info["nterms"],
               hash_tableau=[hsh.info["nrow"], hsh.info["block_enc"]], sig_tableau=[sig.info["nrow"], sig.info["block_enc"]],
               hash_total_mults=hsh.info["total_mults"], sig_total_mults=sig.info["total_mults"], batches=[])
    for B in batches:
        rep = lambda a: np.repeat(a[None, :], B, axis=0)
        wh, ws, whm, wsm = rep(f["w_hash"]), rep(f["w_sig"]), rep(f["w_hash_mac"]), rep(f["w_sig_mac"])
        ch, cs = rep(f["coins"][:nh]), rep(f["coins"][nh:nh + ns])
        best = None
        for r in range(reps + 1):
            ts = api.transcripts(B, bytes.fromhex(e["transcript"]))
            t = [time.perf_counter()]
            ph.commit_batch(wh, ch, ts); t.append(time.perf_counter())
            ps.commit_batch(ws, cs, ts); t.append(time.perf_counter())
            for i in range(B):
                api.transcript_challenge(ts[i], 16)
            a, st1 = ph.prove_committed_batch(whm, ts); t.append(time.perf_counter())
            b, st2 = ps.prove_committed_batch(wsm, ts); t.append(time.perf_counter())
            assert (st1 == 0).all() and (st2 == 0).all()
            d = [1e3 * (t[i + 1] - t[i]) for i in range(4)]
            if r > 0 and (best is None or sum(d) < sum(best)):
                best = d
        rec = dict(batch=B, ms_commit_hash=best[0], ms_commit_sig=best[1], ms_prove_hash=best[2],
                   ms_prove_sig=best[3], ms_total=sum(best), ms_per_proof=sum(best) / B,
                   proofs_per_s=B / sum(best) * 1e3)
        # the verifier's side of the same flow (run_mdoc_verifier, mdoc_zk.cc:673-706): recv_commitment(hash),
        # recv_commitment(sig), the MAC key, verify(hash), verify(sig) on one transcript per proof
        # (lf_zk_verify_committed_batch), host buffers in / status out
        if B <= verify_max_batch:
            kbh, kbs = hsh.info["kbytes"], sig.info["kbytes"]
            pub_h = np.ascontiguousarray(whm[:, :hsh.info["npub_in"] * kbh])
            pub_s = np.ascontiguousarray(wsm[:, :sig.info["npub_in"] * kbs])
            vh, vs = lf.ZkVerifier(hsh), lf.ZkVerifier(sig)
            bestv = None
            for r in range(reps + 1):
                tv = api.transcripts(B, bytes.fromhex(e["transcript"]))
                t = [time.perf_counter()]
                for i in range(B):
                    api.transcript_write(tv[i], a[i][:32])
                    api.transcript_write(tv[i], b[i][:32])
                    api.transcript_challenge(tv[i], 16)
                st1, _ = vh.verify_batch(pub_h, a, transcripts=tv); t.append(time.perf_counter())
                st2, _ = vs.verify_batch(pub_s, b, transcripts=tv); t.append(time.perf_counter())
                assert (st1 == 0).all() and (st2 == 0).all(), (st1, st2)
                dv = [1e3 * (t[i + 1] - t[i]) for i in range(2)]
                if r > 0 and (bestv is None or sum(dv) < sum(bestv)):
                    bestv = dv
            rec.update(ms_verify_hash=bestv[0], ms_verify_sig=bestv[1], ms_verify_total=sum(bestv))
        res["batches"].append(rec)
    return res


def measure_two_in_flight(B=128, rounds=2, device=0, rendezvous=None):
    """two host threads, each with its own context and circuit objects, each proving `rounds` batches of B:
    one batch's serial zero-block hashing (118 ms for the hash circuit) runs under the other's sumcheck"""
    import threading
    f = load_mdoc(); e = f["expect"]
    lanes = []
    for _ in range(2):
        ctx = lf.Context(device)
        sig = lf.Circuit(ctx, lf.FIELD_P256, f["raw"], rate=e["rate"], nreq=e["nreq"], block_enc=e["block_enc_sig"])
        hsh = lf.Circuit(ctx, lf.FIELD_GF2_128, f["raw"][sig.info["lfc1_bytes"]:], rate=e["rate"], nreq=e["nreq"],
                         block_enc=e["block_enc_hash"])
        lanes.append((ctx, lf.ZkProver(hsh), lf.ZkProver(sig), hsh.info["rng_bytes"], sig.info["rng_bytes"]))
    rep = lambda a: np.repeat(a[None, :], B, axis=0)

    def one_batch(lane):
        ctx, ph, ps, nh, ns = lane
        ts = api.transcripts(B, bytes.fromhex(e["transcript"]))
        ph.commit_batch(rep(f["w_hash"]), rep(f["coins"][:nh]), ts)
        ps.commit_batch(rep(f["w_sig"]), rep(f["coins"][nh:nh + ns]), ts)
        for i in range(B):
            api.transcript_challenge(ts[i], 16)
        _, st1 = ph.prove_committed_batch(rep(f["w_hash_mac"]), ts)
        _, st2 = ps.prove_committed_batch(rep(f["w_sig_mac"]), ts)
        assert (st1 == 0).all() and (st2 == 0).all()
    for lane in lanes:
        one_batch(lane)  # warm-up
    ths = [threading.Thread(target=lambda l=l: [one_batch(l) for _ in range(rounds)]) for l in lanes]
    if rendezvous:
        rendezvous()
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return dict(batch=B, batches=2 * rounds, seconds=dt, proofs_per_s=2 * rounds * B / dt, ms_per_proof=1e3 * dt / (2 * rounds * B))


if __name__ == "__main__":
    r = measure()
    r["two_batches_in_flight"] = measure_two_in_flight()
    print(json.dumps(r, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(r, open(os.path.join(ROOT, "gpurun_out", "mdoc_bench.json"), "w"), indent=1)
