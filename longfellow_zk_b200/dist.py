"""Multi-GPU plumbing for throughput mode: independent proofs are sharded
across ranks (one process per GPU) with NO data-path collective; the only
communication is the max-over-ranks reduction of the timing and, optionally,
the gather of per-proof digests.  Works with torch.distributed backends "nccl"
(GPU box) and "gloo" (CPU tests)."""
import hashlib


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` of `world`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def proof_seed(base_seed, index):
    """RNG seed of proof `index` of a job: depends on the proof, not on the rank
    layout, so a job gives the same proofs at any world size."""
    return base_seed + index


def max_over_ranks(value, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_digests(local_proofs, dist=None):
    """sha256 of every local proof, gathered on all ranks in rank order."""
    local = [hashlib.sha256(p).hexdigest() for p in local_proofs]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return [d for part in out for d in part]
