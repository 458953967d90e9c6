"""BASELINE config 1 in full: 100 000 x 768 -> 16 x 16 bits against the faithful reference loop."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import lshrs_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["tcgen05", "tcgen05_3xtf32", "tcgen05_tf32bf16", "ffma"])
def test_config1_all_band_keys(kernel):
    from lshrs_b200 import LSHHasher

    # R = LSHHasher(16, 16, 768, seed=42); X = default_rng(0) standard normal
    X = np.random.default_rng(0).standard_normal((100_000, 768)).astype(np.float32)
    h = LSHHasher(16, 16, 768, seed=42)
    h._ensure_handle()
    h.set_kernel(kernel)
    got = h.hash_batch_packed(X)
    assert h.last_kernel == kernel
    # the golden fixture holds the reference's own bytes for the first 64 of these rows
    from conftest import load_golden

    case = load_golden("hash_cfg_768_16x16")
    np.testing.assert_array_equal(case["X"], X[:64])
    margins64 = oracle.projection_margins(h.projections, X[:64])
    rep64 = oracle.compare_packed(got[:64], case["signatures"], margins64, 1e-5)
    assert rep64["flips_outside_margin"] == 0, rep64
    # all 1.6 M band keys: the one-sgemm oracle everywhere, the per-vector reference loop on 8192 rows
    margins = oracle.projection_margins(h.projections, X)
    rep = oracle.compare_packed(got, oracle.hash_batch_vectorized(h.projections, X), margins, 1e-5)
    assert rep["band_keys"] == 1_600_000
    assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0, rep
    sl = slice(40_000, 48_192)
    rep_ref = oracle.compare_packed(got[sl], oracle.hash_batch_packed(h.projections, X[sl]), margins[sl], 1e-5)
    assert rep_ref["flips_outside_margin"] == 0, rep_ref
    # near-zero flips are counted and must stay a small fraction of the exempt bits
    assert rep["flips_inside_margin"] <= 0.05 * rep["bits_inside_margin"], rep
    print(f"[config 1 {kernel}] {rep}")
