"""GPU parity tests of the rerank path (through the C ABI) against the oracle and golden vectors.

Bar (BASELINE.json north_star): cosine scores within 1e-5 of the reference,
identical top-k sets modulo exact ties.
"""

from __future__ import annotations

import math

import numpy as np
import pytest
from conftest import load_golden, rerank_case_names

from oracle import lshrs_oracle as oracle

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5  # north_star tolerance on cosine scores


def _check_topk(res, sims_ref, k):
    """res = [(pos, score)] best first; same set as the reference modulo ties at the cut."""
    n = len(sims_ref)
    kk = min(k, n)
    assert len(res) == kk
    pos = np.array([p for p, _ in res], dtype=np.int64)
    sc = np.array([s for _, s in res], dtype=np.float64)
    assert len(set(pos.tolist())) == kk and pos.min() >= 0 and pos.max() < n
    np.testing.assert_allclose(sc, sims_ref[pos], atol=SCORE_TOL, rtol=0)
    assert np.all(np.diff(sc) <= 0), "scores must be descending"
    order = np.sort(sims_ref)[::-1]
    cut = order[kk - 1]
    # everything strictly above the cut (by more than the tolerance) must be present, and nothing
    # clearly below it may be
    must = set(np.nonzero(sims_ref > cut + 2 * SCORE_TOL)[0].tolist())
    assert must <= set(pos.tolist())
    assert np.all(sims_ref[pos] >= cut - 2 * SCORE_TOL)


@pytest.mark.parametrize("name", rerank_case_names())
def test_golden_rerank(name):
    from lshrs_b200 import cosine_similarity, top_k_cosine

    case = load_golden(name)
    q, C = case["query"], case["candidates"]
    sims = cosine_similarity(q, C)
    assert sims.dtype == np.float32 and sims.shape == (C.shape[0],)
    np.testing.assert_allclose(sims, case["similarities"], atol=SCORE_TOL, rtol=0)
    for k in case["ks"]:
        res = top_k_cosine(q, C, k=int(k))
        assert all(isinstance(p, int) and isinstance(s, float) for p, s in res)
        _check_topk(res, case["similarities"].astype(np.float64), int(k))
    # candidates as a LIST of 1-D arrays, as the reference's tests pass them
    res_list = top_k_cosine(q, [row for row in C], k=int(case["ks"][0]))
    assert res_list == top_k_cosine(q, C, k=int(case["ks"][0]))


def test_reference_known_answers():
    from lshrs_b200 import cosine_similarity, top_k_cosine

    # reference tests/test_lshrs.py:115-132
    sims = cosine_similarity(np.array([1.0, 0.0, 0.0]), [[1, 0, 0], [0, 1, 0], [-1, 0, 0], [1, 1, 0]])
    np.testing.assert_allclose(sims, [1.0, 0.0, -1.0, 0.70710677], atol=1e-6)
    # reference tests/test_lshrs.py:135-153
    cands = [[1.0, 0.1], [0.0, 1.0], [1.0, 0.0], [-1.0, 0.0], [0.9, 0.2]]
    res = top_k_cosine(np.array([1.0, 0.0]), cands, k=3)
    assert [i for i, _ in res] == [2, 0, 4]
    assert res[0][1] == pytest.approx(1.0, abs=1e-6)
    assert len(top_k_cosine(np.array([1.0, 0.0]), cands, k=10)) == 5
    # reference tests/test_lshrs.py:156-161
    with pytest.raises(ValueError):
        top_k_cosine(np.ones(2), [np.ones(2)], k=0)


def test_error_behaviour():
    from lshrs_b200 import cosine_similarity, top_k_cosine

    with pytest.raises(ValueError, match="zero vector"):
        cosine_similarity(np.zeros(3), [np.ones(3)])
    with pytest.raises(ValueError, match="zero vector"):
        top_k_cosine(np.ones(3), [np.ones(3), np.zeros(3)], k=1)
    with pytest.raises(ValueError, match="at least one array"):
        top_k_cosine(np.ones(3), [], k=1)  # the reference reaches np.stack([]) (similarity.py:85)


def test_ties_break_by_position():
    from lshrs_b200 import top_k_cosine

    c = np.array([[1, 0], [2, 0], [0, 1], [3, 0], [0, 2]], dtype=np.float32)
    res = top_k_cosine(np.array([1.0, 0.0]), c, k=5)
    assert [p for p, _ in res] == [0, 1, 3, 2, 4]


@pytest.mark.parametrize("dim", [768, 128, 33, 1])
@pytest.mark.parametrize("n", [1, 2, 31, 1000, 2000])
def test_scores_and_topk_vs_oracle(dim, n):
    from lshrs_b200 import cosine_similarity, top_k_cosine

    rng = np.random.default_rng(n * 1000 + dim)
    q = rng.standard_normal(dim).astype(np.float32)
    C = rng.standard_normal((n, dim)).astype(np.float32)
    ref = oracle.cosine_similarity(q, C).astype(np.float64)
    np.testing.assert_allclose(cosine_similarity(q, C), ref, atol=SCORE_TOL, rtol=0)
    for k in (1, 10, max(1, math.ceil(n * 0.2)), n, n + 5):
        if dim == 1:  # scores are all +-1: only check the sizes and values
            res = top_k_cosine(q, C, k=k)
            assert len(res) == min(k, n)
            continue
        _check_topk(top_k_cosine(q, C, k=k), ref, k)


def test_batched_gather_from_device_corpus():
    """BASELINE config 4 in miniature: corpus in HBM, CSR candidate ids, k = 10 and p = 0.2."""
    import torch

    from lshrs_b200 import top_k_cosine_batch

    rng = np.random.default_rng(1)
    N, dim, nq, nc = 20_000, 768, 64, 2000
    corpus = rng.standard_normal((N, dim)).astype(np.float32)
    Q = np.random.default_rng(2).standard_normal((nq, dim)).astype(np.float32)
    ids = np.stack([np.random.default_rng(3 + i).choice(N, nc, replace=False) for i in range(nq)]).astype(np.int64)
    offsets = np.arange(nq + 1, dtype=np.int64) * nc
    d_corpus = torch.from_numpy(corpus).cuda()
    for kw, limit in (({"k": 10}, 10), ({"p": 0.2}, oracle.top_p_limit(nc, 0.2)), ({"k": 7, "p": 0.2}, 7)):
        pos, score, count = top_k_cosine_batch(Q, d_corpus, offsets, ids.reshape(-1), vectors_on_device=True, **kw)
        assert (count == limit).all() and pos.shape == (nq, limit)
        pos_h, score_h, count_h = top_k_cosine_batch(Q, corpus, offsets, ids.reshape(-1), **kw)
        np.testing.assert_array_equal(pos, pos_h)
        np.testing.assert_array_equal(score, score_h)
        for i in range(0, nq, 7):
            ref = oracle.cosine_similarity(Q[i], corpus[ids[i]]).astype(np.float64)
            _check_topk([(int(p), float(s)) for p, s in zip(pos[i], score[i])], ref, limit)
    assert limit == 7 and oracle.top_p_limit(nc, 0.2) == 400


def test_ragged_candidate_lists_and_empty_queries():
    from lshrs_b200 import top_k_cosine_batch

    rng = np.random.default_rng(9)
    dim = 64
    lens = np.array([5, 0, 1, 300, 17, 0, 2048, 3], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    V = rng.standard_normal((int(offsets[-1]), dim)).astype(np.float32)
    Q = rng.standard_normal((len(lens), dim)).astype(np.float32)
    pos, score, count = top_k_cosine_batch(Q, V, offsets, None, k=10)
    np.testing.assert_array_equal(count, np.minimum(lens, 10))
    for i, n in enumerate(lens):
        if n == 0:
            continue
        ref = oracle.cosine_similarity(Q[i], V[offsets[i]:offsets[i + 1]]).astype(np.float64)
        _check_topk([(int(p), float(s)) for p, s in zip(pos[i, :count[i]], score[i, :count[i]])], ref, 10)


def test_more_candidates_than_one_sort_buffer():
    """n > 16384 candidates for one query: chunked running top-k inside the kernel, and -- when more than
    8192 results are wanted as well -- the global-memory sort (the reference has no size limit,
    similarity.py:174-179)."""
    from lshrs_b200 import top_k_cosine

    rng = np.random.default_rng(4)
    n, dim = 40_000, 32
    q = rng.standard_normal(dim).astype(np.float32)
    C = rng.standard_normal((n, dim)).astype(np.float32)
    ref = oracle.cosine_similarity(q, C).astype(np.float64)
    _check_topk(top_k_cosine(q, C, k=100), ref, 100)
    _check_topk(top_k_cosine(q, C, k=8192), ref, 8192)
    _check_topk(top_k_cosine(q, C, k=8193), ref, 8193)
    _check_topk(top_k_cosine(q, C, k=n), ref, n)          # every candidate, fully sorted
    _check_topk(top_k_cosine(q, C, k=10 * n), ref, n)


def test_oversized_selection_batch_against_host_sort():
    """Several queries of very different sizes in one oversized launch (one > 16384 candidates with p = 1.0):
    positions must be exactly the stable descending sort of the kernel's own scores."""
    from lshrs_b200.utils.similarity import _get_reranker

    rng = np.random.default_rng(41)
    dim = 24
    sizes = [5, 70_000, 0, 16_385, 1, 20_000]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    V = rng.standard_normal((int(offsets[-1]), dim)).astype(np.float32)
    V[offsets[1] + 7] = V[offsets[1] + 3]      # an exact tie: lower position first
    Q = rng.standard_normal((len(sizes), dim)).astype(np.float32)
    rer = _get_reranker(dim)
    scores, zero = rer.scores(Q, V, offsets)
    for p, k in ((1.0, 0), (0.6, 0), (1.0, 30_000)):
        pos, score, count, zero2 = rer.topk(Q, V, offsets, k=k, p=p)
        for i, n in enumerate(sizes):
            want_n = 0 if n == 0 else min(max(1, math.ceil(n * p)), k or n, n)
            assert count[i] == want_n, (i, p, k)
            s = scores[offsets[i]:offsets[i + 1]]
            order = np.lexsort((np.arange(n), -s.astype(np.float64)))[:want_n]
            np.testing.assert_array_equal(pos[i, :want_n], order)
            np.testing.assert_array_equal(score[i, :want_n], s[order])
        np.testing.assert_array_equal(zero, zero2)


def test_get_above_p_with_more_than_16384_candidates():
    """ADVICE r1 (rerank.cu:278): LSHRS.get_above_p(p=0.95) on buckets that hold most of the corpus."""
    from lshrs_b200 import LSHRS, InMemoryStorage

    rng = np.random.default_rng(5)
    n, dim = 20_000, 16
    base = rng.standard_normal(dim).astype(np.float32)
    X = (base[None, :] + 0.01 * rng.standard_normal((n, dim))).astype(np.float32)   # one tight cluster
    lsh = LSHRS(dim=dim, num_perm=4, num_bands=2, rows_per_band=2, storage=InMemoryStorage(),
                vector_fetch_fn=lambda ids: X[np.asarray(ids, dtype=np.int64)])
    lsh.index(list(range(n)), X)
    res = lsh.get_above_p(X[0], p=0.95)
    cands = sorted(lsh.query(X[0], top_k=None))
    assert len(cands) > 16_384
    assert len(res) == max(1, math.ceil(len(cands) * 0.95))
    ref = oracle.cosine_similarity(X[0], X[cands]).astype(np.float64)
    by_id = dict(zip(cands, ref))
    sc = np.array([s for _, s in res])
    assert np.all(np.diff(sc) <= 0)
    np.testing.assert_allclose(sc, [by_id[i] for i, _ in res], atol=SCORE_TOL, rtol=0)
    assert res[0][0] == 0 or abs(res[0][1] - 1.0) < 1e-6


def test_l2_norm_on_gpu():
    from lshrs_b200 import l2_norm
    from lshrs_b200.utils.norm import l2_norm_batch

    # reference tests/test_lshrs.py:100-112
    out = l2_norm([3.0, 4.0])
    np.testing.assert_allclose(out, [0.6, 0.8], atol=1e-7)
    assert out.dtype == np.float32 and out.shape == (2,)
    with pytest.raises(ValueError, match="Cannot normalize zero vector"):
        l2_norm(np.zeros(3))
    rng = np.random.default_rng(0)
    for dim in (1, 33, 768):
        X = (rng.standard_normal((257, dim)) * np.logspace(-3, 3, 257)[:, None]).astype(np.float32)
        got = l2_norm_batch(X)
        want = np.stack([oracle.l2_norm(x) for x in X])
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
    case = load_golden("rerank_gauss_300x768")
    np.testing.assert_allclose(l2_norm(case["query"]), case["normalized_query"], rtol=2e-6, atol=1e-7)
