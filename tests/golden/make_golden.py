"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref/libref.so,
built from /root/reference by oracle/ref_build/Makefile).  Run in the build
container only:  python tests/golden/make_golden.py

Fixtures:
  sha1_gf128.circuit.z / .witness.z    BM_ShaZK_fp2_128/1 circuit (LFC1) + witness
                                       (circuits/sha/flatsha256_circuit_test.cc:366-468)
  ecdsa1_p256.circuit.z / .witness.z   BM_ECDSAZKProver/1 circuit + witness
                                       (circuits/ecdsa/verify_test.cc:349-407)
  golden.json                          reference proofs (sha256, length, Merkle root,
                                       sumcheck-proof sha256) for seeded RNG streams,
                                       plus small RS / Merkle known answers.
RNG streams are numpy PCG64: default_rng(seed).integers(0,256,n,dtype=uint8).
"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refapi as R  # noqa: E402


def rng_bytes(seed, n):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def main():
    out = {}
    for name, fid, gen in [("sha1_gf128", R.GF2_128_ID, lambda: R.sha_circuit(1)),
                           ("ecdsa1_p256", R.P256_ID, lambda: R.ecdsa_circuit(1))]:
        circ, wit = gen()
        open(os.path.join(HERE, name + ".circuit.z"), "wb").write(zlib.compress(circ, 9))
        open(os.path.join(HERE, name + ".witness.z"), "wb").write(zlib.compress(wit, 9))
        c = R.Circuit(fid, circ)
        info = R.circuit_info(fid, circ)
        proofs = []
        for seed in (1, 2, 3):
            rng = rng_bytes(seed, 1 << 19)
            r = c.prove(wit, rng, tinit=b"test", dump=True)
            kb = 16 if fid == R.GF2_128_ID else 32
            assert c.verify(wit[:info["npub_in"] * kb], r["proof"]) == 0
            sc_len = sum(4 * l["logw"] + 2 for l in info["layers"]) * kb
            proofs.append(dict(seed=seed, tinit="test", rng_used=r["rng_used"], proof_len=len(r["proof"]),
                               proof_sha256=hashlib.sha256(r["proof"]).hexdigest(), root=r["root"].hex(),
                               sumcheck_sha256=hashlib.sha256(r["sumcheck"][:sc_len].tobytes()).hexdigest(),
                               proof_head=r["proof"][:64].hex()))
        out[name] = dict(field_id=fid, circuit_sha256=hashlib.sha256(circ).hexdigest(),
                         witness_sha256=hashlib.sha256(wit).hexdigest(),
                         info={k: v for k, v in info.items()}, proofs=proofs)
    # small known answers of the transforms (reference outputs on seeded inputs)
    ka = {}
    rs = np.random.default_rng(99)
    for n, m in [(455, 4096), (909, 4096), (455, 909), (5, 16)]:
        rows = rs.integers(0, 256, (2, m, 16), dtype=np.uint8)
        ka[f"lch14_rs_{n}_{m}"] = dict(seed=99, sha256=hashlib.sha256(R.lch14_interpolate(n, m, rows).tobytes()).hexdigest())
    out["known_answers"] = ka
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(HERE)))


SHA_MESSAGES = [b"longfellow-zk b200 message %02d" % i + b"." * (i % 20) for i in range(16)]


def distinct_witnesses():
    """witnesses.z: several DIFFERENT satisfying witnesses of the two 1-instance circuits (other SHA-256
    messages; the other ECDSA test vectors of circuits/ecdsa/verify_test.cc:52-169), made by the reference's
    witness generators (oracle/ref_build/ref_sha.cc ref_sha_witness, ref_ecdsa.cc ref_ecdsa_witness), each
    checked by proving and verifying with the reference.  Layout: N x witness_bytes, concatenated."""
    rng = rng_bytes(5, 1 << 19)
    circ, _ = R.sha_circuit(1)
    c = R.Circuit(R.GF2_128_ID, circ)
    ws = [R.sha_witness(1, m) for m in SHA_MESSAGES]
    for w in ws:
        assert c.verify(b"", c.prove(w, rng)["proof"]) == 0
    open(os.path.join(HERE, "sha1_gf128.witnesses.z"), "wb").write(zlib.compress(b"".join(ws), 9))
    circ, _ = R.ecdsa_circuit(1)
    c = R.Circuit(R.P256_ID, circ)
    ws = [R.ecdsa_witness(1, i) for i in range(R.ecdsa_ntests())]
    npub = R.circuit_info(R.P256_ID, circ)["npub_in"]
    for w in ws:
        assert c.verify(w[:npub * 32], c.prove(w, rng)["proof"]) == 0
    open(os.path.join(HERE, "ecdsa1_p256.witnesses.z"), "wb").write(zlib.compress(b"".join(ws), 9))


def published_sizes():
    """sizes/: the other instances docs/content/en/docs/benchmarks.md publishes (BM_ShaZK_fp2_128/2..33,
    BM_ECDSAZKProver/2,3): circuit (LFC1, xz), benchmark witness, and the reference's proof for seed 1."""
    import lzma
    d = os.path.join(HERE, "sizes")
    os.makedirs(d, exist_ok=True)
    out = {}
    todo = [("sha%d_gf128" % n, R.GF2_128_ID, (lambda n=n: R.sha_circuit(n))) for n in (2, 4, 8, 16, 32, 33)]
    todo += [("ecdsa%d_p256" % n, R.P256_ID, (lambda n=n: R.ecdsa_circuit(n))) for n in (2, 3)]
    for name, fid, gen in todo:
        circ, wit = gen()
        open(os.path.join(d, name + ".circuit.xz"), "wb").write(lzma.compress(circ, preset=9))
        open(os.path.join(d, name + ".witness.z"), "wb").write(zlib.compress(wit, 9))
        c = R.Circuit(fid, circ)
        info = R.circuit_info(fid, circ)
        kb = 16 if fid == R.GF2_128_ID else 32
        rng = rng_bytes(1, 1 << 22)
        r = c.prove(wit, rng, tinit=b"test")
        assert c.verify(wit[:info["npub_in"] * kb], r["proof"]) == 0
        out[name] = dict(field_id=fid, circuit_sha256=hashlib.sha256(circ).hexdigest(),
                         witness_sha256=hashlib.sha256(wit).hexdigest(),
                         info={k: v for k, v in info.items() if k != "layers"}, nl=len(info["layers"]),
                         proof=dict(seed=1, tinit="test", rng_used=r["rng_used"], proof_len=len(r["proof"]),
                                    proof_sha256=hashlib.sha256(r["proof"]).hexdigest()))
        print(name, out[name]["proof"])
    json.dump(out, open(os.path.join(d, "golden_sizes.json"), "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        distinct_witnesses()
        published_sizes()
    else:
        main()
