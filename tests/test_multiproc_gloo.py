"""The N>1 host path on CPU: two processes over gloo shard one batch by index, each fills its
disjoint slice of the result (the "host gather"), and the bench's barrier + max-over-ranks timing
reduction works.  The per-shard compute is stood in by a cheap deterministic function: the GPU
kernels themselves are covered by the -m gpu tests."""
from __future__ import annotations

import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, count: int, out_path: str) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from protocols.distributed_keygen_b200.sharding import shard_bounds

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(count, world, rank)
    result = np.lib.format.open_memmap(out_path, mode="r+")
    idx = np.arange(lo, hi, dtype=np.uint64)
    result[lo:hi] = (idx * idx + 7) % 1000003  # disjoint slice of one host array
    result.flush()
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo], dtype=torch.int64))
    assert sum(int(s.item()) for s in sizes) == count
    dist.destroy_process_group()


def test_two_process_index_sharding(tmp_path):
    import numpy as np
    import torch.multiprocessing as mp

    count, world = 100003, 2
    out_path = str(tmp_path / "gathered.npy")
    arr = np.lib.format.open_memmap(out_path, mode="w+", dtype=np.uint64, shape=(count,))
    arr[:] = np.uint64(2**63)
    arr.flush()
    del arr
    port = _free_port()
    mp.spawn(_worker, args=(world, port, count, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    idx = np.arange(count, dtype=np.uint64)
    assert (got == (idx * idx + 7) % 1000003).all()
