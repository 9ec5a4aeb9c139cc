"""world_size-2 gloo test (CPU) of the throughput-mode host logic: proofs of a
job are sharded across ranks with no data-path collective, per-proof seeds do
not depend on the rank layout, and the timing reduction is a max over ranks.
The oracle stands in for the GPU prover here (no GPU in this container)."""
import hashlib
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, njobs, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from fixtures import load, rng_bytes
    from longfellow_zk_b200 import dist as lfd
    from oracle import portapi as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    circ, wit = load("sha1_gf128")
    c = O.Circuit(O.GF2_128_ID, circ)
    lo, hi = lfd.shard_range(njobs, rank, world)
    proofs = [c.prove(wit, rng_bytes(lfd.proof_seed(500, i), 1 << 18))["proof"] for i in range(lo, hi)]
    digests = lfd.gather_digests(proofs, dist)
    tmax = lfd.max_over_ranks(10.0 + rank, dist)
    if rank == 0:
        q.put((digests, tmax, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from longfellow_zk_b200.dist import shard_range
    for n in (0, 1, 5, 8, 1024, 1027):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_job_equals_single_rank_job():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fixtures import load, rng_bytes
    from oracle import portapi as O
    njobs = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, njobs, q)) for r in range(2)]
    for p in procs:
        p.start()
    digests, tmax, shard0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 11.0              # max over ranks, not rank 0's value
    assert shard0 == (0, 2)
    circ, wit = load("sha1_gf128")
    c = O.Circuit(O.GF2_128_ID, circ)
    want = [hashlib.sha256(c.prove(wit, rng_bytes(500 + i, 1 << 18))["proof"]).hexdigest() for i in range(njobs)]
    assert digests == want           # same job, same proofs, any world size
