#!/usr/bin/env python
"""Stage the UNMODIFIED reference (mxngjxa/lshrs) under the git-ignored ``oracle/_ref/``.

    python tools/make_ref.py [--reference /root/reference]

Run in the build container (``__graft_entry__.build()`` calls it whenever ``/root/reference`` exists); the
result travels to the GPU box with the snapshot (``oracle/_ref/`` is git-ignored, not gpurun-ignored), where
``/root/reference`` does not exist.  Nothing from the reference enters the git history.

    oracle/_ref/reference/lshrs/    byte-for-byte copy of the reference package: the CPU arm of
                                    ``bench.py --impl reference`` and ``cpu_baseline`` (kind "reference")
    oracle/_ref/two_import/lshrs/   the same copy with INTEGRATION.md's two-import patch applied to
                                    core/main.py (the reference's own LSHRS on the B200 hasher / rerank)
    oracle/_ref/suite/tests/        the reference's test-suite (+ an empty __init__.py: its files import
                                    ``tests.conftest``, reference tests/test_core.py:9)
    oracle/_ref/stubs/redis/        a stand-in ``redis`` module: the reference imports redis-py at package
                                    import (reference lshrs/storage/redis.py:33), the image has none, and
                                    nothing on the hash / rerank path touches it
    oracle/_ref/MANIFEST.json       sha256 of every copied file, for tests/test_reference_suite.py
"""

from __future__ import annotations

import argparse
import hashlib
import json
import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
OUT = REPO / "oracle" / "_ref"

REDIS_STUB = '''"""Stand-in for redis-py (not installed in this image): just what lshrs.storage.redis touches at import
and construction time.  No server is ever contacted; bucket storage in the tests is an in-memory double."""


class ConnectionPool:
    def __init__(self, **kwargs):
        self.connection_kwargs = kwargs

    def disconnect(self):
        pass


class Redis:
    def __init__(self, connection_pool=None, **kwargs):
        self.connection_pool = connection_pool

    def pipeline(self, *args, **kwargs):
        raise ConnectionError("redis stub: no server in this environment")
'''

TWO_IMPORT_PATCH = (
    ("from lshrs.hash.lsh import LSHHasher\n", "from lshrs_b200.hash.lsh import LSHHasher\n"),
    ("from lshrs.utils.similarity import top_k_cosine\n", "from lshrs_b200.utils.similarity import top_k_cosine\n"),
)


def _copy_tree(src: Path, dst: Path) -> None:
    if dst.exists():
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".DS_Store"))
    for p in dst.rglob("*"):      # the reference tree is read-only; the copy must be replaceable
        p.chmod(0o755 if p.is_dir() else 0o644)
    dst.chmod(0o755)


def stage(reference: Path) -> Path:
    if not (reference / "lshrs" / "core" / "main.py").exists():
        raise FileNotFoundError(f"{reference} does not look like the lshrs reference")
    OUT.mkdir(parents=True, exist_ok=True)
    _copy_tree(reference / "lshrs", OUT / "reference" / "lshrs")
    _copy_tree(reference / "lshrs", OUT / "two_import" / "lshrs")
    main = OUT / "two_import" / "lshrs" / "core" / "main.py"
    text = main.read_text()
    for old, new in TWO_IMPORT_PATCH:
        if text.count(old) != 1:
            raise RuntimeError(f"two-import patch: expected exactly one {old!r} in {main}")
        text = text.replace(old, new)
    main.write_text(text)
    _copy_tree(reference / "tests", OUT / "suite" / "tests")
    (OUT / "suite" / "tests" / "__init__.py").write_text("")
    stub = OUT / "stubs" / "redis"
    stub.mkdir(parents=True, exist_ok=True)
    (stub / "__init__.py").write_text(REDIS_STUB)
    manifest = {"reference": str(reference), "files": {}}
    for p in sorted((OUT / "reference").rglob("*.py")) + sorted((OUT / "suite").rglob("*.py")):
        manifest["files"][str(p.relative_to(OUT))] = hashlib.sha256(p.read_bytes()).hexdigest()
    manifest["tests_collected_expected"] = 71
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    out = stage(Path(args.reference))
    print(f"staged the reference under {out}", file=sys.stderr)
