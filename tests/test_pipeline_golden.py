"""The LSHRS pipeline against outputs of the REFERENCE's own ``LSHRS``.

``tests/golden/pipeline_lshrs.npz`` is written by tools/make_golden_pipeline.py, which runs the unmodified reference
end to end (index -> get_top_k / query / get_above_p / delete) on a seeded data set with its own kind of dict-of-sets
storage double.  The data set is filtered so that no projection is within 10x the parity margin of zero and no two
candidate scores of a query are closer than 1e-4: the ids must then match EXACTLY whatever the float32 summation
order, on the CPU double (oracle arithmetic) and on the B200 alike, with the dict store and with the store in HBM.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

from lshrs_b200 import LSHRS, DeviceBucketStorage, InMemoryStorage

sys.path.insert(0, str(Path(__file__).resolve().parent))
GOLDEN = Path(__file__).resolve().parent / "golden" / "pipeline_lshrs.npz"


@pytest.fixture(params=["double", pytest.param("b200", marks=pytest.mark.gpu)])
def lib(request):
    from lshrs_b200 import _native
    from lshrs_b200.utils import similarity

    made: list = []
    if request.param == "b200":
        yield made
        return
    import fake_lshx

    saved = _native._lib
    fake_lshx.install()
    yield made
    for lsh in made:
        lsh._hasher.close()
        if lsh._dindex is not None:
            lsh._dindex.close()
    for r in list(similarity._rerankers.values()):
        r.close()
    similarity._rerankers.clear()
    _native._lib = saved


def _lists(g, name, dtype=None):
    offs, flat = g[f"{name}_offs"], g[f"{name}_ids" if dtype is None else f"{name}_{dtype}"]
    return [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


@pytest.mark.parametrize("store", ["dict", "hbm", "dict+mirror"])
def test_pipeline_reproduces_the_reference(lib, store):
    g = np.load(GOLDEN)
    X, Q = g["X"], g["Q"]
    kw = dict(dim=int(g["dim"]), num_perm=int(g["num_perm"]), seed=int(g["seed"]),
              vector_fetch_fn=lambda ids: X[np.asarray(ids, dtype=np.int64)])
    if store == "hbm":
        lsh = LSHRS(storage=DeviceBucketStorage(), **kw)
    else:
        lsh = LSHRS(storage=InMemoryStorage(), device_index=(store == "dict+mirror"), **kw)
    lib.append(lsh)
    # the band / row split the reference's auto-configuration chose
    assert (lsh.stats()["num_bands"], lsh.stats()["rows_per_band"]) == (int(g["num_bands"]), int(g["rows_per_band"]))
    lsh.index(list(range(X.shape[0])), X)

    want_topk, want_all = _lists(g, "topk"), _lists(g, "all")
    want_above, want_above_s = _lists(g, "above"), _lists(g, "above", "scores")
    want_both, want_both_s = _lists(g, "both"), _lists(g, "both", "scores")
    for i, q in enumerate(Q):
        assert lsh.get_top_k(q, topk=10) == want_topk[i], i
        assert lsh.query(q, top_k=None) == want_all[i], i
        got = lsh.get_above_p(q, p=0.3)
        assert [j for j, _ in got] == want_above[i], i
        np.testing.assert_allclose([s for _, s in got], want_above_s[i], atol=1e-5, rtol=0)
        got = lsh.query(q, top_k=5, top_p=0.5)
        assert [j for j, _ in got] == want_both[i], i
        np.testing.assert_allclose([s for _, s in got], want_both_s[i], atol=1e-5, rtol=0)
    # the batched calls return what the reference's per-query calls return
    batched_mirror = {} if store == "dict" else {"device_index": True}
    assert lsh.query_batch(Q, top_k=10, **batched_mirror) == want_topk
    assert lsh.query_batch(Q, top_k=None, **batched_mirror) == want_all
    got = lsh.query_batch(Q, top_k=None, top_p=0.3, **batched_mirror)
    assert [[j for j, _ in r] for r in got] == want_above
    # delete, then the same queries
    lsh.delete(g["deleted"].tolist())
    want_after = _lists(g, "after_delete")
    for i, q in enumerate(Q):
        assert lsh.get_top_k(q, topk=10) == want_after[i], i
    assert lsh.query_batch(Q, top_k=10, **batched_mirror) == want_after
