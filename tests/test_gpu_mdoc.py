"""BASELINE config 4: the ISO mdoc proof -- the signature circuit over Fp256 (482 k quad terms) and the
hash circuit over GF(2^128) (7.76 M quad terms, 266 x 4151 tableau) proved on ONE transcript with the
MAC patch of the public inputs between commit and prove (lib/circuits/mdoc/mdoc_zk.cc:398-547).

oracle/_ref/libref_mdoc.so is the reference's mdoc prover split at its four ZkProver calls
(oracle/ref_build/ref_mdoc.cc); the witness comes from the reference's own fill_witness on its benchmark
claim (mdoc_tests[0], age_over_18, kZkSpecs[0]).  The CUDA back end replaces the four calls; the result
must be run_mdoc_prover's bytes: [6 MACs][hash proof][signature proof]."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mdoc_proof_matches_frozen_reference_output(ctx):
    """the same sequence on the committed fixtures (tests/golden/make_golden_mdoc.py): needs nothing
    of the reference at run time"""
    import hashlib
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    from fixtures import load_mdoc
    f = load_mdoc()
    e = f["expect"]
    sig = lf.Circuit(ctx, lf.FIELD_P256, f["raw"], rate=e["rate"], nreq=e["nreq"], block_enc=e["block_enc_sig"])
    hsh = lf.Circuit(ctx, lf.FIELD_GF2_128, f["raw"][sig.info["lfc1_bytes"]:], rate=e["rate"], nreq=e["nreq"],
                     block_enc=e["block_enc_hash"])
    nh, ns = hsh.info["rng_bytes"], sig.info["rng_bytes"]
    assert nh == e["coins_hash"] and nh + ns == e["coins_total"]
    assert (hsh.info["nrow"], hsh.info["block_enc"], hsh.info["nterms"]) == (266, 4151, 7757579)
    ph, ps = lf.ZkProver(hsh), lf.ZkProver(sig)
    ts = api.transcripts(1, bytes.fromhex(e["transcript"]))
    coins = f["coins"]
    _, st = ph.commit_batch(f["w_hash"][None, :], coins[None, :nh], ts)
    assert st[0] == 0
    _, st = ps.commit_batch(f["w_sig"][None, :], coins[None, nh:nh + ns], ts)
    assert st[0] == 0
    assert api.transcript_challenge(ts[0], 16).hex() == e["av"]
    proof_h, st = ph.prove_committed_batch(f["w_hash_mac"][None, :], ts)
    assert st[0] == 0 and hashlib.sha256(proof_h[0]).hexdigest() == e["hash_proof_sha256"]
    proof_s, st = ps.prove_committed_batch(f["w_sig_mac"][None, :], ts)
    assert st[0] == 0 and hashlib.sha256(proof_s[0]).hexdigest() == e["sig_proof_sha256"]
    whole = bytes.fromhex(e["macs"]) + proof_h[0] + proof_s[0]
    assert len(whole) == e["proof_len"] and hashlib.sha256(whole).hexdigest() == e["proof_sha256"]
    # proving without the MAC patch must fail: the patched inputs are checked by the circuits
    ts = api.transcripts(1, bytes.fromhex(e["transcript"]))
    ph.commit_batch(f["w_hash"][None, :], coins[None, :nh], ts)
    _, st = ph.prove_committed_batch(f["w_hash"][None, :], ts)
    assert st[0] == -5


def test_mdoc_proof_matches_reference(ctx):
    import longfellow_zk_b200 as lf
    from longfellow_zk_b200 import api
    from oracle import refapi as ref
    if not ref.mdoc_available():
        pytest.skip("oracle/_ref/libref_mdoc.so not built (the fixture-based test above covers the same bytes)")
    from fixtures import load_mdoc
    raw = load_mdoc()["raw"]
    m = ref.MdocCase(raw)
    coins = np.random.default_rng(2027).integers(0, 256, 4 << 20, dtype=np.uint8)
    want = m.prove(coins)

    sig = lf.Circuit(ctx, lf.FIELD_P256, raw, rate=m.rate, nreq=m.nreq, block_enc=m.block_enc_sig)
    hsh = lf.Circuit(ctx, lf.FIELD_GF2_128, raw[sig.info["lfc1_bytes"]:], rate=m.rate, nreq=m.nreq,
                     block_enc=m.block_enc_hash)
    assert (sig.info["ninputs"], sig.info["npub_in"]) == (m.sig_ninputs, m.sig_npub)
    assert (hsh.info["ninputs"], hsh.info["npub_in"]) == (m.hash_ninputs, m.hash_npub)
    assert hsh.info["nterms"] > 7_000_000 and hsh.info["block_enc"] == 4151
    nh, ns = hsh.info["rng_bytes"], sig.info["rng_bytes"]
    assert nh == want["coins_hash"] and nh + ns == want["coins_total"]

    ws, wh = m.witnesses()
    ph, ps = lf.ZkProver(hsh), lf.ZkProver(sig)
    ts = api.transcripts(1, m.transcript)
    _, st = ph.commit_batch(wh[None, :], coins[None, :nh], ts)
    assert st[0] == 0
    _, st = ps.commit_batch(ws[None, :], coins[None, nh:nh + ns], ts)
    assert st[0] == 0
    av = api.transcript_challenge(ts[0], 16)      # generate_mac_key(tp)
    assert av == want["av"]
    ws2, wh2, macs = m.update_macs(av)            # compute_macs + update_macs (public inputs only)
    proof_h, st = ph.prove_committed_batch(wh2[None, :], ts)
    assert st[0] == 0 and len(proof_h[0]) == want["len_hash"]
    proof_s, st = ps.prove_committed_batch(ws2[None, :], ts)
    assert st[0] == 0 and len(proof_s[0]) == want["len_sig"]
    got = macs + proof_h[0] + proof_s[0]
    assert got == want["proof"]


def test_run_mdoc_prover_drop_in_all_claims(ctx):
    """N1: the reference's run_mdoc_prover, compiled UNCHANGED with ZkProver resolved to ZkProverGpu
    (include/longfellow_b200_adapters.h; oracle/ref_build/ref_mdoc_gpu.cc), on every (claim, mdoc) pair of
    lib/circuits/mdoc/mdoc_zk_test.cc:119-170 -- 27 different documents, issuers, attributes and session
    transcripts, fresh SecureRandomEngine coins each -- and the reference's own run_mdoc_verifier accepts
    every proof; a proof with one byte flipped is rejected."""
    from oracle import refapi as ref
    if not ref.mdoc_gpu_available():
        pytest.skip("oracle/_ref/libref_mdoc_gpu.so not built")
    from fixtures import load_mdoc
    circuit = ref.zstd_compress(load_mdoc()["raw"])
    lib = ref.mdoc_gpu_lib()
    n = lib.ref_mdoc_gpu_nclaims()
    assert n == 27
    lens = set()
    for i in range(n):
        code, plen = ref.mdoc_gpu_run_claim(i, circuit)
        assert code == 0, (i, lib.ref_mdoc_gpu_claim_name(i), code)
        assert 300000 < plen < 400000
        lens.add(plen)
    code, _ = ref.mdoc_gpu_run_claim(0, circuit, tamper=True)
    assert code != 0 and code < 1000   # the prover succeeded, the verifier refused


def test_run_mdoc_verifier_drop_in(ctx):
    """N2 from the reference's side: run_mdoc_verifier compiled UNCHANGED with ZkVerifier resolved to
    ZkVerifierGpu (include/longfellow_b200_adapters.h; ref_mdoc_gpu.cc -DLF_GPU_VERIFIER ->
    libref_mdoc_gpuv.so): two verifiers over two fields on ONE caller transcript, the MAC key drawn between
    recv_commitment and verify (mdoc_zk.cc:673-706), through lf_zk_verify_committed_batch.  For every
    (claim, mdoc) pair of mdoc_zk_test.cc:119-170 the proof of run_mdoc_prover is accepted by the GPU
    verifier; for a sample of them the reference's own verifier agrees, and a flipped byte anywhere in the
    MACs, the hash proof or the signature proof is refused by both with the same code."""
    from oracle import refapi as ref
    if not (ref.mdoc_gpu_available() and ref.mdoc_gpuv_available()):
        pytest.skip("oracle/_ref/libref_mdoc_gpu.so / libref_mdoc_gpuv.so not built")
    from fixtures import load_mdoc
    circuit = ref.zstd_compress(load_mdoc()["raw"])
    A, V = ref.mdoc_gpu_lib(), ref.mdoc_gpuv_lib()
    n = A.ref_mdoc_gpu_nclaims()
    for i in range(n):
        code, proof = ref.mdoc_prove_claim(A, i, circuit)
        assert code == 0 and len(proof) > 300000, (i, code)
        assert ref.mdoc_verify_claim(V, i, circuit, proof) == 0, (i, A.ref_mdoc_gpu_claim_name(i))
        if i % 9 == 0:
            assert ref.mdoc_verify_claim(A, i, circuit, proof) == 0
            for pos in (3, 96 + 40, 96 + 5000, len(proof) // 2, len(proof) - 7):
                bad = bytearray(proof)
                bad[pos] ^= 0x10
                want = ref.mdoc_verify_claim(A, i, circuit, bytes(bad))
                got = ref.mdoc_verify_claim(V, i, circuit, bytes(bad))
                assert want != 0 and got == want, (i, pos, want, got)
    # a proof for another claim's session does not verify
    _, p0 = ref.mdoc_prove_claim(V, 0, circuit)   # both halves on the GPU, one library
    assert ref.mdoc_verify_claim(V, 0, circuit, p0) == 0
    assert ref.mdoc_verify_claim(V, 1, circuit, p0) != 0
