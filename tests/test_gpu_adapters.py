"""The drop-in boundary exercised from the reference's side: oracle/_ref/libref_gpu.so is the
UNMODIFIED reference compiled together with include/longfellow_b200_adapters.h.

(1) The reference's own ZkProver<Field, GpuReedSolomonFactory<Field>> -- sumcheck, Ligero, Merkle and
    transcript on the CPU, every Reed-Solomon row encoded on the GPU through the interpolator-factory
    seam (lib/zk/zk_prover.h:52-53) -- must emit the reference's proof bytes.
(2) GpuZkProver<Field>, fed the reference's Circuit / Dense / RandomEngine objects, must emit them too."""
import hashlib

import pytest

from fixtures import golden, load, rng_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adapters():
    from oracle import refapi
    if not refapi.gpu_adapters_available():
        pytest.fail("oracle/_ref/libref_gpu.so is missing: run __graft_entry__.build() where /root/reference exists")
    return refapi


@pytest.mark.parametrize("name,fid", [("sha1_gf128", 4), ("ecdsa1_p256", 1)])
def test_reference_zkprover_on_gpu_reed_solomon(adapters, name, fid):
    circ, wit = load(name)
    g = golden()[name]["proofs"][0]
    coins = rng_bytes(g["seed"], 1 << 19)
    proof = adapters.GpuAdapterCircuit(fid, circ).prove_reference_with_gpu_rs(wit, coins)
    assert len(proof) == g["proof_len"]
    assert hashlib.sha256(proof).hexdigest() == g["proof_sha256"]


@pytest.mark.parametrize("name,fid", [("sha1_gf128", 4), ("ecdsa1_p256", 1)])
def test_gpu_zkprover_adapter_matches_reference(adapters, name, fid):
    circ, wit = load(name)
    g = golden()[name]["proofs"][0]
    coins = rng_bytes(g["seed"], 1 << 19)
    proof = adapters.GpuAdapterCircuit(fid, circ).prove_gpu(wit, coins, copies=3)
    assert len(proof) == g["proof_len"]
    assert hashlib.sha256(proof).hexdigest() == g["proof_sha256"]
