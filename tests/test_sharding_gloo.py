"""world_size-2 gloo test of the N>1 host logic: shard plan, per-rank work, gather, MAX timing.

The per-rank compute here is the CPU oracle standing in for the kernel (tests may
use it as the checker); what is under test is that contiguous tile-aligned row
shards hashed independently and reassembled equal the single-process result --
the path has no exchange step, so that is the whole N>1 contract.
"""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, result_path: str) -> None:
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from lshrs_b200.sharding import gather_rows, max_over_ranks, shard_range
    from oracle import lshrs_oracle as oracle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        projs = oracle.make_projections(4, 6, 24, seed=42)  # replicated planes, same seed on every rank
        X = np.random.default_rng(0).standard_normal((n, 24)).astype(np.float32)
        lo, hi = shard_range(n, rank, world)
        local = oracle.hash_batch_vectorized(projs, X[lo:hi])
        full = gather_rows(local, n, dst=0)
        slowest = max_over_ranks(1.0 + rank)
        assert slowest == float(world)
        if rank == 0:
            want = oracle.hash_batch_vectorized(projs, X)
            np.save(result_path, np.array([int(np.array_equal(full, want)), full.shape[0]]))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 130, 5])
def test_two_rank_sharded_hash_equals_single(tmp_path, n):
    import torch.multiprocessing as mp

    port = _free_port()
    result = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, port, n, result), nprocs=2, join=True)
    ok, rows = np.load(result)
    assert ok == 1 and rows == n
