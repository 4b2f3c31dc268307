"""
Index sharding of a batch across the GPUs of one box: ciphertexts and
biprime candidates are independent, so rank r of `world` owns one contiguous index range, is fed
by its own host->device copy and writes into a disjoint slice of the output ("host gather").
No data-path collective exists; torch.distributed is used by bench.py only for a barrier and a
max-reduce of the elapsed time.
"""
from __future__ import annotations


def shard_bounds(count: int, world: int, rank: int, granule: int = 1) -> tuple[int, int]:
    """Contiguous range [begin, end) of rank `rank`; boundaries fall on multiples of `granule`
    (e.g. the 40 bases of one biprime candidate stay on one GPU)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    units = (count + granule - 1) // granule
    lo = units * rank // world * granule
    hi = units * (rank + 1) // world * granule
    return min(lo, count), min(hi, count)


def all_shards(count: int, world: int, granule: int = 1) -> list[tuple[int, int]]:
    return [shard_bounds(count, world, r, granule) for r in range(world)]
