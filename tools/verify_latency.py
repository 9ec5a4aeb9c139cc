"""Verifier latency for small batches (lf_zk_verify_batch, host buffers in / status out) on the SHA-256 and
ECDSA circuits, and the reference's run_mdoc_verifier end to end with its ZkVerifier objects on the host
(libref_mdoc_gpu.so) and on the GPU (libref_mdoc_gpuv.so).  LF_VERIFY_SPLIT=0 turns the grid-wide
bind_gh_all of small batches off (one CTA per proof, as for large batches)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import longfellow_zk_b200 as lf  # noqa: E402
from fixtures import load, load_mdoc, rng_bytes  # noqa: E402

ctx = lf.Context(0)
out = dict(env={k: v for k, v in os.environ.items() if k.startswith("LF_")})
for name, fid in (("sha1_gf128", 4), ("ecdsa1_p256", 1)):
    circ, wit = load(name)
    c = lf.Circuit(ctx, fid, circ)
    n = c.info["rng_bytes"]
    for B in (1, 8, 32):
        rng = np.stack([rng_bytes(5 + i, n + 256) for i in range(B)])
        W = np.repeat(np.frombuffer(wit, np.uint8)[None, :], B, axis=0)
        proofs, st = lf.ZkProver(c).prove_batch(W, rng)
        assert (st == 0).all()
        npubb = c.info["npub_in"] * c.info["kbytes"]
        pubs = np.ascontiguousarray(W[:, :npubb]) if npubb else None
        v = lf.ZkVerifier(c)
        st, why = v.verify_batch(pubs, proofs)
        assert (st == 0).all(), (st, why)
        bad = [bytes(p[:200]) + bytes([p[200] ^ 1]) + bytes(p[201:]) for p in proofs]
        st, _ = v.verify_batch(pubs, bad)
        assert (st != 0).all()
        t0 = time.perf_counter()
        for _ in range(5):
            v.verify_batch(pubs, proofs)
        out[f"{name}_B{B}_ms_per_batch"] = 1e3 * (time.perf_counter() - t0) / 5
try:
    from oracle import refapi
    if refapi.mdoc_gpu_available() and refapi.mdoc_gpuv_available():
        circuit = refapi.zstd_compress(load_mdoc()["raw"])
        A, V = refapi.mdoc_gpu_lib(), refapi.mdoc_gpuv_lib()
        code, proof = refapi.mdoc_prove_claim(V, 0, circuit)
        for nm, L in (("mdoc_run_verifier_reference_ms", A), ("mdoc_run_verifier_gpu_ms", V)):
            rc = refapi.mdoc_verify_claim(L, 0, circuit, proof)
            t0 = time.perf_counter()
            for _ in range(3):
                rc |= refapi.mdoc_verify_claim(L, 0, circuit, proof)
            out[nm] = 1e3 * (time.perf_counter() - t0) / 3
            out[nm + "_accepted"] = rc == 0 and code == 0
        bad = bytearray(proof)
        bad[len(bad) // 2] ^= 4
        out["mdoc_tampered_codes_ref_gpu"] = [refapi.mdoc_verify_claim(A, 0, circuit, bytes(bad)),
                                              refapi.mdoc_verify_claim(V, 0, circuit, bytes(bad))]
except Exception as ex:  # noqa: BLE001
    out["mdoc_error"] = repr(ex)
print(json.dumps(out))
