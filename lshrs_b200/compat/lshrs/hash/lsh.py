"""``lshrs.hash.lsh`` -> the B200 hasher (replaces reference lshrs/hash/lsh.py:18-247)."""

from lshrs_b200.hash.lsh import LSHHasher

__all__ = ["LSHHasher"]
