"""The numpy oracle must reproduce every vector the unmodified reference produced.

Fixtures: tests/golden/*.npz, written by tools/make_golden.py from the reference
itself (LSHHasher / cosine_similarity / top_k_cosine / bucket_key /
get_optimal_config).  Plus the reference's own known-answer tests
(reference tests/test_lshrs.py:115-153).
"""

from __future__ import annotations

import numpy as np
import pytest
from conftest import hash_case_names, load_golden, projections_for, rerank_case_names

from oracle import lshrs_oracle as oracle


@pytest.mark.parametrize("name", hash_case_names())
def test_oracle_signatures_match_reference(name):
    case = load_golden(name)
    projs = projections_for(case)  # also checks the PCG64 projection stream (sha256 of R)
    R = np.concatenate(projs, axis=0)
    np.testing.assert_array_equal(np.array([R[0, 0], R[0, -1], R[-1, 0], R[-1, -1]]), case["R_corner"])
    if case["R_small"].size:
        np.testing.assert_array_equal(R, case["R_small"])
    got = oracle.hash_batch_packed(projs, case["X"])
    np.testing.assert_array_equal(got, case["signatures"])
    # per-vector path and container types
    first = oracle.hash_vector(projs, case["X"][0])
    assert all(isinstance(b, bytes) and len(b) == oracle.band_bytes(int(case["rows_per_band"])) for b in first)
    assert b"".join(first) == case["signatures"][0].tobytes()


def test_survey_known_answers(golden_manifest):
    # KAT1 / KAT3 / KAT5 of SURVEY.md section 8c, regenerated from the reference by make_golden.py
    assert golden_manifest["cases"]["hash_kat1_3x5x4"]["first_hex"] == ["0b", "13", "11"]
    assert golden_manifest["cases"]["hash_cfg_768_16x16"]["first_hex"][:4] == ["741b", "cae0", "4a6f", "d3fd"]
    assert golden_manifest["cases"]["hash_cfg_128_16x4_gauss"]["first_hex"][:4] == ["02", "0f", "09", "0d"]
    projs = oracle.make_projections(16, 16, 768, 42)
    assert np.float32(projs[0][0, 0]) == np.float32(0.3047171)
    assert np.float32(projs[15][15, -1]) == np.float32(-0.40886137)


def test_bit_order_and_padding():
    # packbits([1,0,0,0,0,0,0,0,1,1], "little") -> 01 03
    proj = np.eye(10, dtype=np.float32)
    v = np.array([1, -1, -1, -1, -1, -1, -1, -1, 1, 1], dtype=np.float32)
    assert oracle.project_and_pack(proj, v) == bytes([0x01, 0x03])


def test_vectorized_variant_agrees_outside_margin():
    projs = oracle.make_projections(16, 16, 768, 42)
    X = np.random.default_rng(3).standard_normal((512, 768)).astype(np.float32)
    ref = oracle.hash_batch_packed(projs, X)
    fast = oracle.hash_batch_vectorized(projs, X)
    rep = oracle.compare_packed(fast, ref, oracle.projection_margins(projs, X))
    assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0


def test_bucket_keys(golden_manifest):
    assert oracle.bucket_key("lsh", 5, b"\xab\xcd") == golden_manifest["bucket_key_example"] == "lsh:5:bucket:abcd"
    projs = oracle.make_projections(16, 16, 768, 42)
    X = np.random.default_rng(0).standard_normal((2, 768)).astype(np.float32)
    for row, want in zip(X, golden_manifest["bucket_keys_768"]):
        got = [oracle.bucket_key("lsh", b, hv) for b, hv in enumerate(oracle.hash_vector(projs, row))]
        assert got == want


def test_hash_validation_matches_reference():
    projs = oracle.make_projections(2, 3, 4)
    with pytest.raises(ValueError):
        oracle.hash_vector(projs, np.arange(5, dtype=np.float32))
    with pytest.raises(ValueError):
        oracle.hash_batch(projs, np.arange(3, dtype=np.float32))
    with pytest.raises(ValueError):
        oracle.hash_batch(projs, np.ones((2, 5), dtype=np.float32))
    for bad in ((0, 1, 1), (1, 0, 1), (1, 1, 0)):
        with pytest.raises(ValueError):
            oracle.make_projections(*bad)


@pytest.mark.parametrize("name", rerank_case_names())
def test_oracle_rerank_matches_reference(name):
    case = load_golden(name)
    q, C = case["query"], case["candidates"]
    sims = oracle.cosine_similarity(q, C)
    np.testing.assert_array_equal(sims, case["similarities"])
    np.testing.assert_array_equal(oracle.l2_norm(q), case["normalized_query"])
    for k in case["ks"]:
        res = oracle.top_k_cosine(q, C, k=int(k))
        np.testing.assert_array_equal(np.array([i for i, _ in res]), case[f"top{k}_idx"])
        np.testing.assert_array_equal(np.array([s for _, s in res]), case[f"top{k}_score"])


def test_reference_known_answer_cosine():
    # reference tests/test_lshrs.py:115-132
    sims = oracle.cosine_similarity(np.array([1.0, 0.0, 0.0]), [[1, 0, 0], [0, 1, 0], [-1, 0, 0], [1, 1, 0]])
    np.testing.assert_allclose(sims, [1.0, 0.0, -1.0, 0.70710677], atol=1e-6)
    # reference tests/test_lshrs.py:135-161
    res = oracle.top_k_cosine(np.array([1.0, 0.0]), [[1.0, 0.1], [0.0, 1.0], [1.0, 0.0], [-1.0, 0.0], [0.9, 0.2]], k=3)
    assert [i for i, _ in res] == [2, 0, 4]
    with pytest.raises(ValueError):
        oracle.top_k_cosine(np.ones(2), [np.ones(2)], k=0)
    with pytest.raises(ValueError):
        oracle.l2_norm(np.zeros(3))


def test_top_p_limit_and_zero_test():
    assert oracle.top_p_limit(2000, 0.2) == 400
    assert oracle.top_p_limit(3, 0.01) == 1
    assert oracle.top_p_limit(2000, 0.2, top_k=10) == 10
    assert oracle.is_zero_vector(np.zeros(4)) and oracle.is_zero_vector(np.full(4, 1e-9))
    assert not oracle.is_zero_vector(np.array([0, 0, 2e-8, 0])) and not oracle.is_zero_vector(np.array([np.nan, 0]))


def test_auto_config_matches_reference(golden_manifest):
    """LSHRS(dim, num_perm=...) without num_bands / rows_per_band: lshrs_b200.utils.br must select what the
    reference's get_optimal_config selects (reference lshrs/utils/br.py:325-395) on the whole recorded grid."""
    from lshrs_b200.core.main import _auto_config
    from lshrs_b200.utils.br import get_optimal_config

    for n, want in golden_manifest["optimal_config"].items():
        assert list(_auto_config(int(n), 0.5)) == want
    assert list(get_optimal_config(4096, 0.9)) == golden_manifest["optimal_config_4096_0.9"]
    grid = golden_manifest["optimal_config_grid"]
    assert len(grid) >= 200
    for key, want in grid.items():
        n, t = key.split("@")
        assert list(map(int, get_optimal_config(int(n), float(t)))) == want, key
