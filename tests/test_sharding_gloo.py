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


def _weighted_worker(rank: int, world: int, port: int, n: int, result_path: str) -> None:
    """The e2e split of bench.py at N > 1: every rank reports its measured link rate (all_gather_object over
    gloo), all ranks derive the same link-weighted row counts, hash their block, and the blocks reassemble."""
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from lshrs_b200 import fabric
    from oracle import lshrs_oracle as oracle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rates = [None] * world
        dist.all_gather_object(rates, [23.5, 36.0][rank])          # a slow and a fast host link
        rows = fabric.weighted_rows(n, rates, align=128)
        d2h = [None] * world
        dist.all_gather_object(d2h, [12.0, 18.8][rank])            # D2H with both ranks copying
        plan = fabric.plan_relay(d2h, [0.0, 44.0], kernel_gbs=29.0)
        assert sum(rows) == n and rows[1] > rows[0]
        assert plan["policy"] == "relay" and plan["pairs"] == {0: 1}       # the slow rank hands over to the fast one
        lo = sum(rows[:rank])
        projs = oracle.make_projections(4, 6, 24, seed=42)
        X = np.random.default_rng(0).standard_normal((n, 24)).astype(np.float32)
        local = oracle.hash_batch_vectorized(projs, X[lo:lo + rows[rank]])
        blocks = [None] * world
        dist.all_gather_object(blocks, local)
        if rank == 0:
            want = oracle.hash_batch_vectorized(projs, X)
            np.save(result_path, np.array([int(np.array_equal(np.concatenate(blocks), want)), rows[0], rows[1]]))
    finally:
        dist.destroy_process_group()


def test_two_rank_link_weighted_split(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    result = str(tmp_path / "ok.npy")
    mp.spawn(_weighted_worker, args=(2, port, 4000, result), nprocs=2, join=True)
    ok, r0, r1 = np.load(result)
    assert ok == 1 and r0 + r1 == 4000 and r0 % 128 == 0 and abs(r1 / r0 - 36.0 / 23.5) < 0.1
