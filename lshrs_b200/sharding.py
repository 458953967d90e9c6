"""Multi-GPU sharding of the hash path: independent row blocks, no collective.

Each vector's signature depends only on that vector and the (replicated)
projection matrix (reference lshrs/hash/lsh.py:200-211), so a batch splits into
contiguous row blocks, one per GPU, and the results land in disjoint slices of
one host array.  Two ways to drive it:

* one process per GPU under ``torchrun`` (what ``bench.py`` does):
  :func:`shard_bounds` gives rank ``r`` its row range, :func:`max_over_ranks`
  reduces a device-measured duration, :func:`gather_rows` reassembles per-rank
  results on rank 0 (only used by tests / small jobs -- at scale every rank
  writes its own pinned slice and nothing is exchanged);
* one process, one thread per GPU: :class:`ShardedHasher` (ctypes drops the
  GIL during the C call).

The host links of a multi-GPU box are not equal (``lshrs_b200/fabric.py``):
``ShardedHasher(balance="links")`` measures every GPU's H2D rate once and sizes
the row blocks to it instead of splitting evenly.
"""

from __future__ import annotations

import threading
from collections.abc import Sequence

import numpy as np

__all__ = ["shard_bounds", "weighted_bounds", "shard_range", "max_over_ranks", "gather_rows", "ShardedHasher"]

_TILE = 128  # rows per CTA tile; shard edges are tile-aligned so no tile straddles two GPUs


def shard_bounds(n: int, world_size: int, align: int = _TILE) -> list[tuple[int, int]]:
    """Contiguous, balanced, ``align``-row-aligned ``[start, stop)`` ranges covering ``range(n)``."""
    if world_size <= 0:
        raise ValueError("world_size must be > 0")
    if n < 0:
        raise ValueError("n must be >= 0")
    tiles = (n + align - 1) // align
    base, extra = divmod(tiles, world_size)
    bounds = []
    start_tile = 0
    for r in range(world_size):
        t = base + (1 if r < extra else 0)
        lo = min(start_tile * align, n)
        hi = min((start_tile + t) * align, n)
        bounds.append((lo, hi))
        start_tile += t
    return bounds


def weighted_bounds(n: int, weights: Sequence[float], align: int = _TILE) -> list[tuple[int, int]]:
    """Contiguous ``align``-row-aligned ranges covering ``range(n)`` with sizes proportional to ``weights``."""
    if not weights or min(weights) <= 0:
        raise ValueError("weights must be positive")
    if n < 0:
        raise ValueError("n must be >= 0")
    total = float(sum(weights))
    bounds, start, acc = [], 0, 0.0
    for i, w in enumerate(weights):
        acc += w
        stop = n if i == len(weights) - 1 else min(n, int(round(n * acc / total / align)) * align)
        stop = max(stop, start)
        bounds.append((start, stop))
        start = stop
    return bounds


def shard_range(n: int, rank: int, world_size: int, align: int = _TILE) -> tuple[int, int]:
    return shard_bounds(n, world_size, align)[rank]


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a per-rank scalar (a device-measured duration) over the default process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local: np.ndarray, n_total: int, dst: int = 0):
    """Reassemble per-rank row blocks (split by :func:`shard_bounds`) on rank ``dst``; None elsewhere."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(n_total, world)
    assert local.shape[0] == bounds[rank][1] - bounds[rank][0], "local block does not match its shard"
    trailing = local.shape[1:]
    max_rows = max(hi - lo for lo, hi in bounds)
    pad = np.zeros((max_rows, *trailing), dtype=local.dtype)
    pad[: local.shape[0]] = local
    mine = torch.from_numpy(pad)
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    out = np.empty((n_total, *trailing), dtype=local.dtype)
    for (lo, hi), b in zip(bounds, bufs):
        out[lo:hi] = b.numpy()[: hi - lo]
    return out


class ShardedHasher:
    """One ``LSHHasher`` per GPU with identical planes; ``hash_batch_packed`` splits rows across them."""

    def __init__(self, num_bands: int, rows_per_band: int, dim: int, seed: int = 42,
                 devices: Sequence[int] | None = None, balance: str = "equal") -> None:
        from lshrs_b200 import _native
        from lshrs_b200.hash.lsh import LSHHasher

        if devices is None:
            devices = list(range(max(1, _native.device_count())))
        self.devices = list(devices)
        if balance not in ("equal", "links"):
            raise ValueError("balance must be 'equal' or 'links'")
        self.weights: list[float] | None = None
        if balance == "links" and len(self.devices) > 1:
            from lshrs_b200 import fabric

            # host batches cross PCIe host -> device: shards sized to each GPU's measured H2D rate
            self.weights = [max(r, 1e-3) for r in fabric.probe_links(self.devices)["h2d_gbs"]]
        self.hashers = [LSHHasher(num_bands, rows_per_band, dim, seed, device=d) for d in self.devices]
        self.num_bands, self.rows_per_band, self.dim = num_bands, rows_per_band, dim
        self.bytes_per_band = self.hashers[0].bytes_per_band
        self.signature_bytes = self.hashers[0].signature_bytes

    @property
    def projections(self):
        return self.hashers[0].projections

    @projections.setter
    def projections(self, value) -> None:
        for h in self.hashers:
            h.projections = value

    def hash_batch_packed(self, vectors: np.ndarray) -> np.ndarray:
        arr = self.hashers[0]._validate_batch(vectors)
        arr = np.ascontiguousarray(arr, dtype=np.float32)   # (hash_into below takes float32 rows)
        n = arr.shape[0]
        out = np.empty((n, self.signature_bytes), dtype=np.uint8)
        bounds = weighted_bounds(n, self.weights) if self.weights else shard_bounds(n, len(self.hashers))
        errors: list[BaseException] = []

        def work(h, lo, hi):
            try:
                if hi > lo:
                    h.hash_into(arr[lo:hi], hi - lo, out[lo:hi], x_on_device=False, out_on_device=False)
            except BaseException as exc:  # noqa: BLE001
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(h, lo, hi)) for h, (lo, hi) in zip(self.hashers, bounds)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out.reshape(n, self.num_bands, self.bytes_per_band)

    def close(self) -> None:
        for h in self.hashers:
            h.close()
