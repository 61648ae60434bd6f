"""Multi-GPU plumbing: one process per GPU, reads sharded, genome replicated, count tables summed.

Nothing here touches the data path: each rank tallies its own shard through the C ABI; the only exchange is the
final all-reduce of the small count tables (NCCL over NVLink on GPUs; the same code runs over gloo on CPUs for
tests).  Integer sums are order independent, so any shard count gives identical tables.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous split of n_items over `world` ranks (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sam_bytes(sam, rank: int, world: int):
    """Byte range [lo, hi) of `sam` (bytes / np.uint8) for this rank, cut after newlines so that every record
    belongs to exactly one rank (records are independent: pss-bam.c:764-783 handles one line at a time)."""
    a = sam if isinstance(sam, np.ndarray) else np.frombuffer(sam, dtype=np.uint8)
    n = a.size

    def cut(pos):
        if pos <= 0:
            return 0
        if pos >= n:
            return n
        nl = np.flatnonzero(a[pos - 1:min(n, pos - 1 + (1 << 20))] == 10)     # first newline at or after pos-1
        while nl.size == 0:
            pos += 1 << 20
            if pos >= n:
                return n
            nl = np.flatnonzero(a[pos - 1:min(n, pos - 1 + (1 << 20))] == 10)
        return pos - 1 + int(nl[0]) + 1

    return cut(n * rank // world), cut(n * (rank + 1) // world)


def allreduce_tables(t):
    """Sum a torch int64 tensor of counters over all ranks in place (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
