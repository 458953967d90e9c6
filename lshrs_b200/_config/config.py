"""Value type returned by the hasher (mirrors reference lshrs/_config/config.py:12-71)."""

from __future__ import annotations

from collections.abc import Iterator
from dataclasses import dataclass, field


@dataclass(frozen=True, eq=True)
class HashSignatures:
    """Band signatures of ONE vector: ``bands[b]`` is the packed key of band ``b``.

    Same contract as the reference type: immutable, hashable, iterable in band
    order, every element coerced to ``bytes`` on construction (so bytearrays /
    memoryviews / numpy rows are accepted).
    """

    bands: tuple[bytes, ...] = field(default=())

    def __post_init__(self) -> None:
        coerced = tuple(b if type(b) is bytes else bytes(b) for b in self.bands)
        object.__setattr__(self, "bands", coerced)

    @classmethod
    def from_packed(cls, row, bytes_per_band: int) -> "HashSignatures":
        """Build from one row of the packed C-ABI layout uint8[num_bands * bytes_per_band]."""
        raw = row.tobytes() if hasattr(row, "tobytes") else bytes(row)
        return cls(tuple(raw[i : i + bytes_per_band] for i in range(0, len(raw), bytes_per_band)))

    def __iter__(self) -> Iterator[bytes]:
        return iter(self.bands)

    def __len__(self) -> int:
        return len(self.bands)

    def __getitem__(self, item: int) -> bytes:
        return self.bands[item]

    def as_tuple(self) -> tuple[bytes, ...]:
        return self.bands

    def hex(self) -> tuple[str, ...]:
        """Lower-case hex of each band -- the variable part of RedisStorage.bucket_key."""
        return tuple(b.hex() for b in self.bands)


def signatures_from_packed(packed, bytes_per_band: int) -> list[HashSignatures]:
    """uint8[n, num_bands * bytes_per_band] (or [n, num_bands, bpb]) -> list of HashSignatures.

    The band ``bytes`` objects are created in C by viewing the rows as fixed-width void items
    (``tolist()`` of a ``V<bpb>`` array yields exact-length ``bytes``, trailing zeros included), and
    the already-normalised tuples bypass ``__post_init__``: about 2x the per-object Python path.
    """
    import numpy as np

    arr = np.ascontiguousarray(packed).reshape(packed.shape[0], -1)
    rows = arr.view(f"V{bytes_per_band}").tolist()
    new, set_field = object.__new__, object.__setattr__
    out = []
    append = out.append
    for row in rows:
        sig = new(HashSignatures)
        set_field(sig, "bands", tuple(row))
        append(sig)
    return out


__all__ = ["HashSignatures", "signatures_from_packed"]
