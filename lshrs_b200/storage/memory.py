"""Bucket-storage protocol and an in-process implementation.

The reference keeps buckets in Redis (``RedisStorage``, reference
lshrs/storage/redis.py) and that component stays on the host UNCHANGED: pass a
reference ``RedisStorage`` instance as ``LSHRS(storage=...)`` and it is used as
is.  This module only states the five-method protocol ``LSHRS`` relies on and
provides a dict-backed implementation for tests, benches and single-process
use.  ``bucket_key`` reproduces the reference's key format byte for byte
(redis.py:225) -- it is the only consumer of the packed band bytes.
"""

from __future__ import annotations

import threading
from collections.abc import Iterable
from typing import Protocol, runtime_checkable

BucketOperation = tuple[int, bytes, int]  # (band_id, band bytes, vector index), as in the reference

__all__ = ["BucketOperation", "BucketStorage", "InMemoryStorage", "bucket_key"]


def bucket_key(prefix: str, band_id: int, hash_val: bytes) -> str:
    """``"{prefix}:{band_id}:bucket:{hex}"`` -- lower-case hex, byte 0 first."""
    return f"{prefix}:{band_id}:bucket:{hash_val.hex()}"


@runtime_checkable
class BucketStorage(Protocol):
    """What ``LSHRS`` needs from a bucket store (the public surface of RedisStorage it calls)."""

    def batch_add(self, operations: list[BucketOperation]) -> None: ...

    def get_bucket(self, band_id: int, hash_val: bytes) -> set[int]: ...

    def remove_indices(self, indices: list[int]) -> None: ...

    def clear(self) -> None: ...

    def close(self) -> None: ...


class InMemoryStorage:
    """Thread-safe dict of sets per band, one set per band key -- the buckets a Redis key of :func:`bucket_key`
    names (the string itself is only formatted on request: at hundreds of thousands of operations per second the
    f-string per operation was most of this double's time, and so was a ``(band, key)`` tuple per lookup -- hence
    one dict per band, keyed by the band bytes alone)."""

    def __init__(self, prefix: str = "lsh") -> None:
        self.prefix = prefix
        self._bands: dict[int, dict[bytes, set[int]]] = {}
        self._lock = threading.Lock()

    def bucket_key(self, band_id: int, hash_val: bytes) -> str:
        return bucket_key(self.prefix, band_id, hash_val)

    def batch_add(self, operations: Iterable[BucketOperation]) -> None:
        bands = self._bands
        with self._lock:
            last_band, table = None, None
            for band_id, hash_val, index in operations:
                if band_id != last_band:
                    table = bands.get(band_id)
                    if table is None:
                        table = bands[band_id] = {}
                    last_band = band_id
                members = table.get(hash_val)
                if members is None:
                    table[hash_val] = {index}
                else:
                    members.add(index)

    def add_to_bucket(self, band_id: int, hash_val: bytes, index: int) -> None:
        self.batch_add([(band_id, hash_val, int(index))])

    def get_bucket(self, band_id: int, hash_val: bytes) -> set[int]:
        with self._lock:
            return set(self._bands.get(band_id, {}).get(bytes(hash_val), ()))

    def get_buckets(self, keys: Iterable[tuple[int, bytes]]) -> list[set[int]]:
        """Many buckets in one call (the batched query path asks for nq * num_bands at once)."""
        bands, empty = self._bands, {}
        with self._lock:
            return [set(bands.get(b, empty).get(h, ())) for b, h in keys]

    def remove_indices(self, indices: Iterable[int]) -> None:
        drop = {int(i) for i in indices}
        with self._lock:
            for table in self._bands.values():
                for members in table.values():
                    members -= drop

    def clear(self) -> None:
        with self._lock:
            self._bands.clear()

    def close(self) -> None:
        pass

    def keys(self) -> list[str]:
        """The Redis-style key strings of the non-empty buckets."""
        with self._lock:
            return [bucket_key(self.prefix, b, h) for b, table in self._bands.items() for h, m in table.items() if m]

    def __len__(self) -> int:
        with self._lock:
            return sum(1 for table in self._bands.values() for m in table.values() if m)

    def __getstate__(self) -> dict:
        with self._lock:
            return {"prefix": self.prefix,
                    "buckets": {(b, h): set(m) for b, table in self._bands.items() for h, m in table.items()}}

    def __setstate__(self, state: dict) -> None:
        self.prefix = state["prefix"]
        self._bands = {}
        for (b, h), members in state["buckets"].items():
            self._bands.setdefault(b, {})[h] = set(members)
        self._lock = threading.Lock()
