"""``DeviceBucketStorage`` -- the bucket store itself in HBM -- against the reference's data model.

The storage protocol is the reference's (lshrs/storage/redis.py: batch_add / get_bucket / remove_indices / clear;
tests/conftest.py:15-78 MockStorage is its dict-of-sets model), so the bar is exact equality with
``InMemoryStorage`` fed the same calls: bucket members, candidate lists, ``LSHRS`` results, what stays buffered
when ``index()`` meets an invalid row, and the order of operations.  Every scenario runs twice: in the CPU tier on
the oracle-backed double of the C ABI (host logic), and with ``-m gpu`` on the real library (kernels).
"""

from __future__ import annotations

import pickle
import sys
from pathlib import Path

import numpy as np
import pytest

from lshrs_b200 import LSHRS, DeviceBucketStorage, InMemoryStorage

sys.path.insert(0, str(Path(__file__).resolve().parent))


@pytest.fixture(params=["double", pytest.param("b200", marks=pytest.mark.gpu)])
def lib(request):
    """Either the CPU double of liblshx (installed for the test) or the real library on the GPU box."""
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
    for obj in made:                        # handles of the double must never reach the real library
        if isinstance(obj, LSHRS):
            obj._hasher.close()
            if obj._dindex is not None:
                obj._dindex.close()
        elif isinstance(obj, DeviceBucketStorage) and obj.index is not None:
            obj.index.close()
    for r in list(similarity._rerankers.values()):
        r.close()
    similarity._rerankers.clear()
    _native._lib = saved


def _clustered(n, dim, seed=0):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((max(1, n // 4), dim)).astype(np.float32)
    return (np.repeat(centers, 4, axis=0)[:n] + 0.05 * rng.standard_normal((n, dim))).astype(np.float32)


def _pair(made, X, **kw):
    """The same LSHRS twice: dict-of-sets store vs the store in HBM."""
    dim = X.shape[1]
    args = dict(dim=dim, num_perm=16, num_bands=8, rows_per_band=2, vector_fetch_fn=lambda ids: X[np.asarray(ids, dtype=np.int64)])
    args.update(kw)
    a = LSHRS(storage=InMemoryStorage(), **args)
    b = LSHRS(storage=DeviceBucketStorage(), **args)
    made += [a, b]
    return a, b


def test_storage_protocol_equals_dict_of_sets(lib):
    rng = np.random.default_rng(3)
    nb, bpb = 6, 3
    mem, dev = InMemoryStorage(), DeviceBucketStorage(nb, bpb)
    lib.append(dev)
    pool = [bytes(rng.integers(0, 256, bpb, dtype=np.uint8)) for _ in range(12)]

    def same():
        keys = [(b, k) for b in range(nb) for k in pool] + [(0, b"\x00" * bpb)]
        assert dev.get_buckets(keys) == mem.get_buckets(keys)
        assert dev.get_bucket(2, pool[0]) == mem.get_bucket(2, pool[0])

    same()                                                     # empty store
    # whole-vector runs (what LSHRS enqueues) -> the packed path
    ops = [(b, pool[int(rng.integers(12))], i) for i in range(40) for b in range(nb)]
    mem.batch_add(ops), dev.batch_add(ops)
    same()
    # stray operations, repeats, one id under several keys of a band -> single entries
    ops = [(int(rng.integers(nb)), pool[int(rng.integers(12))], int(rng.integers(60))) for _ in range(101)]
    mem.batch_add(ops), dev.batch_add(ops + ops[:7])
    mem.add_to_bucket(1, pool[3], 2 ** 40 + 5), dev.add_to_bucket(1, pool[3], 2 ** 40 + 5)
    same()
    mem.remove_indices([3, 4, 59, 1000]), dev.remove_indices([3, 4, 59, 1000])
    same()
    mem.batch_add([(0, pool[0], 3)]), dev.batch_add([(0, pool[0], 3)])     # a removed id comes back
    same()
    assert dev.get_bucket(0, b"\x01") == set()                  # a key of another length names no bucket
    with pytest.raises(ValueError):
        dev.batch_add([(nb, pool[0], 1)])
    with pytest.raises(ValueError):
        dev.batch_add([(0, pool[0], -2)])
    # persistence: npz round trip and pickle carry exactly the live members
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        dev.save(Path(tmp) / "store.npz")
        back = DeviceBucketStorage.load(Path(tmp) / "store.npz")
    lib.append(back)
    clone = pickle.loads(pickle.dumps(dev))
    lib.append(clone)
    keys = [(b, k) for b in range(nb) for k in pool]
    assert back.get_buckets(keys) == mem.get_buckets(keys) == clone.get_buckets(keys)
    mem.clear(), dev.clear()
    same()
    with pytest.raises(RuntimeError, match="not bound"):
        DeviceBucketStorage().get_bucket(0, b"ab")


def test_lshrs_on_the_device_store_equals_the_dict_store(lib):
    X = _clustered(600, 16)
    a, b = _pair(lib, X, buffer_size=1000)
    assert b._store_on_device and b._dindex is b._storage.index
    rng = np.random.default_rng(5)
    Q = X[rng.integers(0, 500, 48)] + 0.01 * rng.standard_normal((48, 16)).astype(np.float32)

    def same():
        for q in Q[:12]:
            assert a.get_top_k(q, topk=7) == b.get_top_k(q, topk=7)
            assert a.query(q, top_k=None) == b.query(q, top_k=None)
            ra, rb = a.get_above_p(q, p=0.5), b.get_above_p(q, p=0.5)
            assert [i for i, _ in ra] == [i for i, _ in rb]
            np.testing.assert_allclose([s for _, s in ra], [s for _, s in rb], atol=1e-6)
            assert a._candidate_counts(a._prepare_vector(q)) == b._candidate_counts(b._prepare_vector(q))
        for kw in (dict(top_k=None), dict(top_k=5), dict(top_k=3, top_p=0.5)):
            ra, rb = a.query_batch(Q, **kw), b.query_batch(Q, **kw)          # b: the device join by default
            if "top_p" in kw:
                assert [[i for i, _ in r] for r in ra] == [[i for i, _ in r] for r in rb]
            else:
                assert ra == rb
        assert b.query_batch(Q, top_k=5, device_index=False) == a.query_batch(Q, top_k=5)   # through get_buckets

    for lsh in (a, b):
        lsh.index(list(range(400)), X[:400])
        for i in range(400, 430):
            lsh.ingest(i, X[i])                     # 30 x 8 = 240 buffered operations, invisible to queries
    assert len(a._buffer) == len(b._buffer) == 240
    same()
    for lsh in (a, b):
        lsh.index(np.arange(430, 500), X[430:500])   # ndarray ids; the buffered ingests are flushed first
    assert not a._buffer and not b._buffer
    same()
    for lsh in (a, b):
        lsh.delete([5, 6, 433])
        lsh.delete(7)
    same()
    for lsh in (a, b):
        lsh.index([5, 599], X[[5, 501]])
    same()
    restored = pickle.loads(pickle.dumps(b))
    lib.append(restored)
    restored._vector_fetch_fn = b._vector_fetch_fn
    assert restored.query_batch(Q, top_k=None) == a.query_batch(Q, top_k=None)
    for lsh in (a, b):
        lsh.clear()
    same()


@pytest.mark.parametrize("buffer_size, pre, bad_row, bad_kind", [
    (40, 0, 13, "zero"), (40, 3, 13, "negative"), (40, 0, 4, "zero"), (40, 0, 5, "zero"), (40, 4, 9, "zero"),
    (8, 0, 7, "zero"), (1000, 2, 19, "negative"), (40, 0, 0, "zero"), (17, 1, 11, "zero"),
])
def test_invalid_row_leaves_the_same_buffer_as_the_per_row_loop(lib, buffer_size, pre, bad_row, bad_kind):
    """index() on the packed path = the reference's ingest loop (main.py:504-518): flushed rows, buffered rows and
    their order when row ``bad_row`` is rejected, for buffers that flush before, at and after it."""
    X = _clustered(20, 16, seed=2)
    a, b = _pair(lib, X, buffer_size=buffer_size)
    ids = list(range(100, 120))
    bad = X.copy()
    if bad_kind == "zero":
        bad[bad_row] = 0
    else:
        ids[bad_row] = -1
    for lsh in (a, b):
        for i in range(pre):
            lsh.ingest(50 + i, X[i])
        with pytest.raises(ValueError, match="zero vector" if bad_kind == "zero" else "non-negative"):
            lsh.index(ids, bad)
    assert a._buffer == b._buffer
    probe = [(band, bytes(sig)) for row in a._hasher.hash_batch_packed(X) for band, sig in enumerate(row)]
    assert a._storage.get_buckets(probe) == b._storage.get_buckets(probe)       # what reached the stores
    for lsh in (a, b):
        lsh.flush()
    assert a._storage.get_buckets(probe) == b._storage.get_buckets(probe)


def test_failed_packed_add_keeps_the_operations(lib):
    X = _clustered(12, 16, seed=4)

    class Flaky(DeviceBucketStorage):
        fail = True

        def add_packed(self, signatures, ids):
            if self.fail:
                raise ConnectionError("down")
            super().add_packed(signatures, ids)

    st = Flaky()
    lsh = LSHRS(dim=16, num_perm=16, num_bands=8, rows_per_band=2, storage=st)
    lib.append(lsh)
    lsh.ingest(99, X[0])
    with pytest.raises(ConnectionError):
        lsh.index(list(range(12)), X)
    # the earlier ingest went out with the flush that precedes the packed add; the batch is buffered again
    assert len(st) == 1 and len(lsh._buffer) == 12 * 8
    st.fail = False
    lsh.flush()
    assert len(lsh._buffer) == 0
    got = lsh.query_batch(X, top_k=None)
    assert all(i in got[i] for i in range(12)) and 99 in got[0]


@pytest.mark.gpu
def test_index_of_resident_vectors_equals_the_host_path():
    """index(ids, <CUDA tensor>) on the device store: hash + append in HBM; same store contents as from host arrays,
    same per-row semantics when a row is invalid."""
    import torch

    X = _clustered(5000, 64, seed=9)
    made: list = []
    a, b = _pair(made, X, dim=64, num_perm=32, num_bands=8, rows_per_band=4, buffer_size=100)
    b.ingest(7000, X[0])
    a.ingest(7000, X[0])
    a.index(list(range(5000)), X)
    xd = torch.from_numpy(X).cuda()
    b.index(torch.arange(3000, dtype=torch.int64), xd[:3000])          # tensor ids
    b.index(list(range(3000, 5000)), xd[3000:])                        # sequence ids
    Q = X[::50] + 0.01
    assert a.query_batch(Q, top_k=None) == b.query_batch(Q, top_k=None)
    assert a.query_batch(Q, top_k=7) == b.query_batch(Q, top_k=7, corpus=None)
    bad = xd[:40].clone()
    bad[17] = 0
    for lsh, v in ((a, bad.cpu().numpy()), (b, bad)):
        with pytest.raises(ValueError, match="zero vector"):
            lsh.index(list(range(9000, 9040)), v)
    assert a._buffer == b._buffer
    for lsh in (a, b):
        lsh.flush()
    assert a.query_batch(Q, top_k=None) == b.query_batch(Q, top_k=None)
    with pytest.raises(ValueError, match="shape"):
        b.index([1, 2], xd[:2, :10])
    with pytest.raises(ValueError, match="does not match"):
        b.index([1, 2, 3], xd[:2])


@pytest.mark.gpu
def test_single_queries_rerank_against_a_resident_corpus():
    """LSHRS(corpus=<CUDA tensor>): get_above_p / query(top_p=) gather candidate vectors from HBM by id -- same ids,
    scores within 1e-6 of the vector_fetch_fn path; works with either store; query_batch picks the corpus up."""
    import torch

    X = _clustered(4000, 96, seed=11)
    xd = torch.from_numpy(X).cuda()
    fetch = lambda ids: X[np.asarray(ids, dtype=np.int64)]   # noqa: E731
    kw = dict(dim=96, num_perm=128, num_bands=16, rows_per_band=8)
    host = LSHRS(storage=InMemoryStorage(), vector_fetch_fn=fetch, **kw)
    res = LSHRS(storage=DeviceBucketStorage(), corpus=xd, **kw)
    mem = LSHRS(storage=InMemoryStorage(), corpus=xd, **kw)
    for lsh in (host, res, mem):
        lsh.index(list(range(4000)), X)
    Q = X[::97] + 0.02
    for q in Q:
        for call in (lambda l: l.get_above_p(q, p=0.3), lambda l: l.query(q, top_k=5, top_p=0.9),
                     lambda l: l.query(q, top_k=None, top_p=1.0)):
            want = call(host)
            for lsh in (res, mem):
                got = call(lsh)
                assert [i for i, _ in got] == [i for i, _ in want]
                np.testing.assert_allclose([s for _, s in got], [s for _, s in want], atol=1e-6)
    from lshrs_b200 import _native

    before = _native.launch_count()
    assert res.get_above_p(Q[3], p=0.3)
    assert _native.launch_count() - before == 3          # hash + lookup/join, rerank, id gather: the fused path
    want = host.query_batch(Q, top_k=4, top_p=0.5)
    for lsh in (res, mem):
        got = lsh.query_batch(Q, top_k=4, top_p=0.5)                 # corpus= defaults to the instance's
        assert [[i for i, _ in r] for r in got] == [[i for i, _ in r] for r in want]
    res.index([4500], X[:1])                                          # an id beyond the corpus
    with pytest.raises(ValueError, match="not a row of the corpus"):
        res.get_above_p(X[0], p=1.0)
    with pytest.raises(ValueError, match="corpus must be"):
        LSHRS(storage=InMemoryStorage(), corpus=torch.zeros(4, 95, device="cuda"), **kw)
    with pytest.raises(ValueError, match="corpus must be"):
        LSHRS(storage=InMemoryStorage(), corpus=X, **kw)
    res.set_corpus(None)
    with pytest.raises(RuntimeError, match="vector_fetch_fn"):
        res.get_above_p(X[1], p=0.5)


def test_packed_index_matches_the_per_row_loop_for_random_schedules():
    """Property test (CPU double): for random buffer sizes, pre-buffered ingests, batch lengths and an optional
    invalid row, index() on the packed path leaves the same buffer, the same store and raises the same error as
    the tuple path that walks the rows like the reference's loop (main.py:504-518)."""
    import fake_lshx
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from lshrs_b200 import _native
    from lshrs_b200.utils import similarity

    X = _clustered(64, 16, seed=8)
    probe_rows = None

    @settings(max_examples=60, deadline=None)
    @given(buffer_size=st.integers(1, 150), pre=st.integers(0, 12), n=st.integers(1, 40),
           bad=st.one_of(st.none(), st.tuples(st.integers(0, 39), st.sampled_from(["zero", "negative"]))))
    def run(buffer_size, pre, n, bad):
        class Recording(InMemoryStorage):
            def __init__(self):
                super().__init__()
                self.batches = []

            def batch_add(self, operations):
                self.batches.append(list(operations))
                super().batch_add(operations)

        made: list = []
        a, b = _pair(made, X, buffer_size=buffer_size)
        a._storage = Recording()
        # c: the literal row-by-row loop (LSHRS._index_rows), the yardstick for both batched paths
        c = LSHRS(storage=Recording(), dim=16, num_perm=16, num_bands=8, rows_per_band=2, buffer_size=buffer_size)
        made.append(c)
        ids = list(range(100, 100 + n))
        batch = X[:n].copy()
        expect = None
        if bad is not None and bad[0] < n:
            if bad[1] == "zero":
                batch[bad[0]] = 0
                expect = "zero vector"
            else:
                ids[bad[0]] = -1
                expect = "non-negative"
        for lsh in (a, b):
            for i in range(pre):
                lsh.ingest(500 + i, X[40 + i])
            if expect:
                with pytest.raises(ValueError, match=expect):
                    lsh.index(ids, batch)
            else:
                lsh.index(ids, batch)
        for i in range(pre):
            c.ingest(500 + i, X[40 + i])
        packed, flag = c._hasher.hash_batch_packed(batch, return_zero_flag=True)
        if expect:
            with pytest.raises(ValueError, match=expect):
                c._index_rows(ids, packed, flag)
        else:
            c._index_rows(ids, packed, flag)
        assert a._buffer == b._buffer == c._buffer
        assert a._storage.batches == c._storage.batches          # the same batch_add calls, batch for batch
        probe = [(band, bytes(sig)) for row in a._hasher.hash_batch_packed(X) for band, sig in enumerate(row)]
        assert a._storage.get_buckets(probe) == b._storage.get_buckets(probe) == c._storage.get_buckets(probe)
        for lsh in (a, b, c):
            lsh.flush()
        assert a._storage.get_buckets(probe) == b._storage.get_buckets(probe) == c._storage.get_buckets(probe)
        for lsh in made:
            lsh._hasher.close()
            if lsh._dindex is not None:
                lsh._dindex.close()

    saved = _native._lib
    fake_lshx.install()
    try:
        run()
    finally:
        for r in list(similarity._rerankers.values()):
            r.close()
        similarity._rerankers.clear()
        _native._lib = saved


def test_single_queries_beyond_the_latency_path_fall_back_to_the_batched_join(lib):
    """A query that matches more bucket entries than the one-launch path sorts in shared memory (4096) is reported
    as -1 by lshx_index_query_vectors and answered by the batched join instead -- same results as the dict store."""
    rng = np.random.default_rng(21)
    base = rng.standard_normal((1, 16)).astype(np.float32)
    X = np.concatenate([np.repeat(base, 700, axis=0) + 1e-4 * rng.standard_normal((700, 16)).astype(np.float32),
                        _clustered(100, 16, seed=22)])
    a, b = _pair(lib, X)
    for lsh in (a, b):
        lsh.index(list(range(800)), X)
    ids, coll, counts, _ = b._dindex.query_vectors(b._hasher, X[:1], 16)
    assert counts[0] == -1                                   # 700 near-duplicates x 8 bands = 5600 bucket entries
    assert a.get_top_k(X[0], topk=20) == b.get_top_k(X[0], topk=20)
    assert a.query(X[0], top_k=None) == b.query(X[0], top_k=None) and len(b.query(X[0], top_k=None)) >= 700
    ra, rb = a.get_above_p(X[0], p=0.01), b.get_above_p(X[0], p=0.01)
    assert [i for i, _ in ra] == [i for i, _ in rb]
    assert a.get_top_k(X[750], topk=5) == b.get_top_k(X[750], topk=5)      # an ordinary query next to it
