"""``lshrs`` import paths on top of ``lshrs_b200`` -- the drop-in form of the package.

The reference's users and its own test-suite import ``lshrs.LSHRS``, ``lshrs.hash.lsh.LSHHasher``,
``lshrs._config.config.HashSignatures``, ``lshrs.utils.similarity.{cosine_similarity, top_k_cosine}``,
``lshrs.utils.norm.l2_norm`` and ``lshrs.utils.br.*`` (reference tests/test_lshrs.py:6-15,
tests/conftest.py:11-12).  Putting :func:`path` FIRST on ``sys.path`` makes those names resolve to the
B200 implementation; everything this repository leaves to the host unchanged -- ``lshrs.storage.redis``
(Redis buckets) and ``lshrs.io`` (PostgreSQL / Parquet loaders) -- keeps resolving to the reference package
installed behind it, through the shim package's ``__path__``.

    import sys, lshrs_b200.compat
    lshrs_b200.compat.install()          # or: sys.path.insert(0, lshrs_b200.compat.path())
    from lshrs import LSHRS              # lshrs_b200.LSHRS; LSHRS(...).ingest / get_top_k / get_above_p

``tests/test_reference_suite.py`` runs the reference's own 71 tests through exactly this arrangement.
"""

from __future__ import annotations

import sys
from pathlib import Path

__all__ = ["path", "install"]


def path() -> str:
    """Directory to put first on ``sys.path`` so that ``import lshrs`` is the B200 drop-in."""
    return str(Path(__file__).resolve().parent)


def install() -> None:
    """Put :func:`path` first on ``sys.path`` (no-op when ``lshrs`` is already imported from it)."""
    p = path()
    loaded = sys.modules.get("lshrs")
    if loaded is not None and not str(getattr(loaded, "__file__", "")).startswith(p):
        raise RuntimeError("another 'lshrs' package is already imported; call lshrs_b200.compat.install() first")
    if p in sys.path:
        sys.path.remove(p)
    sys.path.insert(0, p)
