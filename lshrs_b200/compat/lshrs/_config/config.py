"""``lshrs._config.config`` -> the signature container (replaces reference lshrs/_config/config.py:12-71)."""

from lshrs_b200._config.config import HashSignatures

__all__ = ["HashSignatures"]
