"""
Multi-GPU layer: one process per GPU (`torch.distributed`, NCCL over NVLink), pairs sharded statically.

Pairs are independent, so there is no data-path collective: every rank traces its contiguous
block of vertices against all antennas.  The only exchange is an optional gather of the compact per-pair records to
rank 0 (`gather_compact`), which the reference's own bookkeeping would need to write one output file.
"""
import numpy as np


def shard_bounds(n_items, world_size, rank):
    """contiguous block [lo, hi) of `n_items` for `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_vertices(n_vertices, world_size, rank, permutation_seed=None):
    """indices of the vertices of `rank`.  With `permutation_seed` the vertices are shuffled first (same permutation on
    every rank) so that geometric clustering of cheap shadow-zone pairs cannot unbalance the ranks."""
    lo, hi = shard_bounds(n_vertices, world_size, rank)
    if permutation_seed is None:
        return np.arange(lo, hi)
    perm = np.random.default_rng(permutation_seed).permutation(n_vertices)
    return np.sort(perm[lo:hi])


def gather_compact(local, group=None, dst=0):
    """
    Gather a dict of per-pair torch tensors (leading dimension = local pairs, unequal across ranks) to rank `dst`.
    Uses all_gather on padded tensors (NCCL has no gatherv); returns the concatenated dict on every rank (cheap for the
    compact records: <= ~200 B per pair) -- callers on ranks != dst may drop it.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    keys = sorted(local.keys())
    n_local = torch.tensor([local[keys[0]].shape[0]], dtype=torch.int64, device=local[keys[0]].device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    out = {}
    for k in keys:
        t = local[k]
        pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out
