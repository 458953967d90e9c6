"""End-to-end behaviour of ``LSHRS`` on the GPU kernels, mirroring the reference's own suite
(reference tests/test_core.py, test_buffer_semantics.py, test_concurrency.py) with an
in-memory bucket store in place of Redis, exactly as the reference's MockStorage does.
"""

from __future__ import annotations

import threading

import numpy as np
import pytest

from oracle import lshrs_oracle as oracle

pytestmark = pytest.mark.gpu


class RecordingStorage:
    """In-memory bucket store that records every batch (the role of the reference's MockStorage)."""

    def __init__(self, fail_on_flush: bool = False):
        from lshrs_b200 import InMemoryStorage

        self.inner = InMemoryStorage()
        self.batches: list[list] = []
        self.fail_on_flush = fail_on_flush
        self.closed = False
        self._lock = threading.Lock()

    def batch_add(self, operations):
        if self.fail_on_flush:
            raise ConnectionError("Simulated Redis failure")
        with self._lock:
            self.batches.append(list(operations))
        self.inner.batch_add(operations)

    def get_bucket(self, band_id, hash_val):
        return self.inner.get_bucket(band_id, hash_val)

    def remove_indices(self, indices):
        self.inner.remove_indices(indices)

    def clear(self):
        self.inner.clear()

    def close(self):
        self.closed = True

    @property
    def all_operations(self):
        with self._lock:
            return [op for b in self.batches for op in b]


def make_lsh(storage=None, **kw):
    from lshrs_b200 import LSHRS

    args = dict(dim=32, num_bands=4, rows_per_band=4, num_perm=16, buffer_size=10_000, seed=42,
                storage=storage or RecordingStorage())
    args.update(kw)
    return LSHRS(**args)


def test_index_produces_the_reference_operations(rng):
    """index() (one batched kernel call) == the reference's per-row ingest loop, op for op."""
    st = RecordingStorage()
    lsh = make_lsh(st, buffer_size=64)
    X = rng.standard_normal((50, 32)).astype(np.float32)
    lsh.index(list(range(100, 150)), X)
    projs = oracle.make_projections(4, 4, 32, 42)
    want_ops = [(b, hv, 100 + i) for i, x in enumerate(X) for b, hv in enumerate(oracle.hash_vector(projs, x))]
    assert st.all_operations == want_ops
    # flush boundaries: 16 vectors x 4 bands = 64 ops per auto-flush, then the final flush
    assert [len(b) for b in st.batches] == [64, 64, 64, 8]


def test_ingest_and_query_roundtrip(rng):
    # reference tests/test_core.py:111-150
    data = rng.standard_normal((200, 32)).astype(np.float32)
    lsh = make_lsh(vector_fetch_fn=lambda ids: data[np.asarray(ids)])
    lsh.index(list(range(200)), data)
    for probe in (0, 17, 199):
        assert probe in lsh.get_top_k(data[probe], topk=5)
    noisy = data[42] + 0.01 * rng.standard_normal(32).astype(np.float32)
    assert 42 in lsh.get_top_k(noisy, topk=5)
    res = lsh.get_above_p(data[42], p=0.5)
    assert res and all(isinstance(i, int) and isinstance(s, float) for i, s in res)
    assert res[0][0] == 42 and res[0][1] == pytest.approx(1.0, abs=1e-5)
    assert [s for _, s in res] == sorted([s for _, s in res], reverse=True)


def test_query_matches_oracle_pipeline(rng):
    """query(top_p) == collision ordering + reference rerank + rank-fraction cut, id for id."""
    data = rng.standard_normal((500, 32)).astype(np.float32)
    data[250:] = data[:250] + 0.05 * rng.standard_normal((250, 32)).astype(np.float32)  # near duplicates
    st = RecordingStorage()
    lsh = make_lsh(st, vector_fetch_fn=lambda ids: data[np.asarray(ids)])
    lsh.index(list(range(500)), data)
    projs = oracle.make_projections(4, 4, 32, 42)
    for probe in (3, 77, 260):
        counts: dict[int, int] = {}
        for b, hv in enumerate(oracle.hash_vector(projs, data[probe])):
            for c in st.get_bucket(b, hv):
                counts[c] = counts.get(c, 0) + 1
        ordered = [i for i, _ in sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))]
        assert lsh.get_top_k(data[probe], topk=7) == ordered[:7]
        ref = oracle.top_k_cosine(data[probe], data[ordered], k=len(ordered))
        limit = oracle.top_p_limit(len(ordered), 0.3)
        got = lsh.get_above_p(data[probe], p=0.3)
        assert len(got) == limit
        np.testing.assert_allclose([s for _, s in got], [s for _, s in ref[:limit]], atol=1e-5)
        assert {i for i, _ in got} == {ordered[p] for p, _ in ref[:limit]}
        both = lsh.query(data[probe], top_k=2, top_p=0.3)
        assert both == got[:2]
    # batched form == per-row form
    probes = data[[3, 77, 260]]
    assert lsh.query_batch(probes, top_k=7) == [lsh.get_top_k(p, topk=7) for p in probes]
    batch = lsh.query_batch(probes, top_k=None, top_p=0.3)
    for row, p in zip(batch, probes):
        single = lsh.get_above_p(p, p=0.3)
        assert [i for i, _ in row] == [i for i, _ in single]
        np.testing.assert_allclose([s for _, s in row], [s for _, s in single], atol=1e-6)


def test_query_validation(rng):
    # reference tests/test_core.py:158-192
    data = rng.standard_normal((10, 32)).astype(np.float32)
    lsh = make_lsh()
    lsh.index(list(range(10)), data)
    with pytest.raises(ValueError, match="top_k"):
        lsh.query(data[0], top_k=0)
    with pytest.raises(ValueError, match="top_p"):
        lsh.query(data[0], top_p=1.5)
    with pytest.raises(RuntimeError, match="vector_fetch_fn"):
        lsh.query(data[0], top_p=0.5)
    assert make_lsh().get_top_k(data[0]) == []


def test_index_stops_at_first_zero_vector_like_the_reference(rng):
    st = RecordingStorage()
    lsh = make_lsh(st)
    X = rng.standard_normal((6, 32)).astype(np.float32)
    X[4] = 0.0
    with pytest.raises(ValueError, match="Cannot index zero vector"):
        lsh.index([0, 1, 2, 3, 4, 5], X)
    lsh.flush()
    assert {idx for _, _, idx in st.all_operations} == {0, 1, 2, 3}


def test_buffer_semantics(rng):
    # reference tests/test_buffer_semantics.py
    st = RecordingStorage()
    lsh = make_lsh(st, buffer_size=10_000)
    v = rng.standard_normal(32).astype(np.float32)
    lsh.ingest(1, v)
    assert st.batches == [] and lsh.get_top_k(v) == []      # buffered, not queryable yet
    lsh.flush()
    assert lsh.get_top_k(v) == [1]
    st2 = RecordingStorage()
    lsh2 = make_lsh(st2, buffer_size=8)                       # 2 vectors x 4 bands
    lsh2.ingest(1, v)
    assert st2.batches == []
    lsh2.ingest(2, v * 2)
    assert [len(b) for b in st2.batches] == [8]
    lsh2.ingest(3, v * 3)
    lsh2.close()
    assert [len(b) for b in st2.batches] == [8, 4] and st2.closed


def test_flush_failure_restores_buffer(rng):
    # reference tests/test_core.py:337-357
    st = RecordingStorage(fail_on_flush=True)
    lsh = make_lsh(st)
    lsh.ingest(1, rng.standard_normal(32).astype(np.float32))
    with pytest.raises(ConnectionError):
        lsh.flush()
    assert len(lsh._buffer) == 4
    st.fail_on_flush = False
    lsh.flush()
    assert len(st.all_operations) == 4 and not lsh._buffer


def test_delete_and_clear(rng):
    # reference tests/test_core.py:278-329
    lsh = make_lsh()
    v = rng.standard_normal(32).astype(np.float32)
    lsh.index([0], v[None, :])
    assert lsh.get_top_k(v) == [0]
    lsh.delete(0)
    assert lsh.get_top_k(v) == []
    lsh.index([5, 6], np.stack([v, v]))
    lsh.clear()
    assert lsh.get_top_k(v) == []


def test_concurrent_ingest(rng):
    # reference tests/test_concurrency.py:13-48
    st = RecordingStorage()
    lsh = make_lsh(st, buffer_size=10)
    vectors = rng.standard_normal((100, 32)).astype(np.float32)

    def work(t):
        for i in range(t * 10, t * 10 + 10):
            lsh.ingest(i, vectors[i])

    threads = [threading.Thread(target=work, args=(t,)) for t in range(10)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    lsh.flush()
    ops = st.all_operations
    assert len(ops) == 100 * 4
    assert {idx for _, _, idx in ops} == set(range(100))


def test_full_size_config_autoconfig_and_bucket_keys():
    from lshrs_b200 import LSHRS, InMemoryStorage

    st = InMemoryStorage()
    lsh = LSHRS(dim=768, num_perm=256, storage=st)  # auto-config -> 16 x 16 like the reference
    X = np.random.default_rng(0).standard_normal((2, 768)).astype(np.float32)
    lsh.index([10, 11], X)
    import json
    from pathlib import Path

    manifest = json.loads((Path(__file__).parent / "golden" / "manifest.json").read_text())
    by_key = {st.bucket_key(b, h): members for b, table in st._bands.items() for h, members in table.items()}
    assert sorted(by_key) == sorted(st.keys())
    for row, keys in zip((10, 11), manifest["bucket_keys_768"]):
        for key in keys:                      # the reference's own Redis key strings for these two vectors
            assert row in by_key[key]


def test_create_signatures_from_parquet(tmp_path, rng):
    pa = pytest.importorskip("pyarrow")
    import pyarrow.parquet as pq

    data = rng.standard_normal((300, 32)).astype(np.float32)
    path = tmp_path / "vectors.parquet"
    pq.write_table(pa.table({"index": pa.array(np.arange(300), pa.int64()),
                             "vector": pa.array(data.tolist(), pa.list_(pa.float32()))}), path)
    st = RecordingStorage()
    lsh = make_lsh(st)
    lsh.create_signatures(format="parquet", source=path, batch_size=128)
    assert len(st.all_operations) == 300 * 4
    assert [len(b) for b in st.batches] == [128 * 4, 128 * 4, 44 * 4]   # one flush per loader batch
    for probe in (0, 150, 299):
        assert probe in lsh.get_top_k(data[probe], topk=5)


def test_query_batch_with_device_resident_corpus(rng):
    """vector_fetch_fn replaced by an id gather from a corpus held in HBM: same results."""
    import torch

    from lshrs_b200 import InMemoryStorage

    data = rng.standard_normal((400, 32)).astype(np.float32)
    data[200:] = data[:200] + 0.05 * rng.standard_normal((200, 32)).astype(np.float32)
    lsh = make_lsh(InMemoryStorage(), vector_fetch_fn=lambda ids: data[np.asarray(ids)])
    lsh.index(list(range(400)), data)
    probes = data[[1, 50, 333, 399]]
    host = lsh.query_batch(probes, top_k=None, top_p=0.5)
    dev = lsh.query_batch(probes, top_k=None, top_p=0.5, corpus=torch.from_numpy(data).cuda())
    assert [[i for i, _ in row] for row in dev] == [[i for i, _ in row] for row in host]
    for a, b in zip(dev, host):
        np.testing.assert_allclose([s for _, s in a], [s for _, s in b], atol=1e-6)
    assert lsh.query_batch(probes, top_k=3, top_p=0.5, corpus=torch.from_numpy(data).cuda()) == [row[:3] for row in dev]
    empty = make_lsh(InMemoryStorage())
    assert empty.query_batch(probes, top_k=5) == [[], [], [], []]
