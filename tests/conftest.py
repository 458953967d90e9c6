"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
GOLDEN = REPO / "tests" / "golden"
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_manifest() -> dict:
    return json.loads((GOLDEN / "manifest.json").read_text())


def load_golden(name: str):
    return np.load(GOLDEN / f"{name}.npz")


def hash_case_names() -> list[str]:
    return sorted(p.stem for p in GOLDEN.glob("hash_*.npz"))


def rerank_case_names() -> list[str]:
    return sorted(p.stem for p in GOLDEN.glob("rerank_*.npz"))


def projections_for(case) -> list[np.ndarray]:
    """Rebuild the case's projection matrices with the oracle and check them against the reference's."""
    import hashlib

    from oracle import lshrs_oracle as oracle

    nb, r, dim, seed = (int(case[k]) for k in ("num_bands", "rows_per_band", "dim", "seed"))
    projs = oracle.make_projections(nb, r, dim, seed)
    R = np.concatenate(projs, axis=0)
    assert hashlib.sha256(R.tobytes()).digest() == case["R_sha256"].tobytes(), "projection stream drifted"
    return projs


@pytest.fixture
def rng() -> np.random.Generator:
    return np.random.default_rng(12345)
