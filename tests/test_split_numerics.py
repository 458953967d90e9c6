"""CPU model of the tcgen05 kernel's operand splits (tools/sim_splits.py): the error of every arm stays
orders of magnitude inside the 1e-5 * |x||r| parity margin, and the scaled FP16x3 arm flags exactly the
vectors that leave FP16's range (the kernel recomputes those in FP32)."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import sim_splits  # noqa: E402


@pytest.mark.parametrize("dim, nperm, n, kind", [(768, 256, 1500, "gauss"), (128, 64, 6000, "sift"),
                                                 (1536, 512, 400, "gauss"), (768, 256, 1000, "wide")])
def test_split_errors_are_far_inside_the_margin(dim, nperm, n, kind):
    R = np.random.default_rng(42).standard_normal((nperm, dim)).astype(np.float32)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    if kind == "sift":
        X = np.minimum(255, np.floor(np.abs(X) * 40)).astype(np.float32)
    if kind == "wide":
        X = (X * np.exp2(rng.integers(-12, 6, size=X.shape))).astype(np.float32)
    rep = sim_splits.report(X, R)
    for arm in ("3xtf32", "tf32+bf16", "fp16x3"):
        assert rep[arm]["flips_outside_margin"] == 0, (arm, rep[arm])
        assert rep[arm]["max_rel_err"] < 1e-6, (arm, rep[arm])       # the margin is 1e-5
    assert rep["fp16x3"]["max_rel_err"] < 2e-7 and rep["3xtf32"]["max_rel_err"] < 2e-7
    assert rep["fp16x3"]["recomputed_rows"] == 0


def test_fp16x3_flags_vectors_that_leave_the_range():
    rng = np.random.default_rng(9)
    R = rng.standard_normal((64, 256)).astype(np.float32)
    X = rng.standard_normal((200, 256)).astype(np.float32)
    X[::2, :32] *= np.float32(1e-9)        # scale fixed by a tiny first chunk: the rest overflows
    X[1, 200] = 1e9                        # one huge late element
    _, redo = sim_splits.dot_fp16x3(X, R)
    want = np.zeros(200, dtype=bool)
    want[::2] = True
    want[1] = True
    np.testing.assert_array_equal(redo, want)
    # within the range the scale never changes a sign
    d, redo = sim_splits.dot_fp16x3(X[3::2], R)
    truth = X[3::2].astype(np.float64) @ R.astype(np.float64).T
    assert not redo.any()
    scale = np.linalg.norm(X[3::2].astype(np.float64), axis=1)[:, None] * np.linalg.norm(R.astype(np.float64), axis=1)
    assert (np.abs(d - truth) / scale).max() < 2e-7
