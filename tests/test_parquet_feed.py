"""The Arrow-buffer Parquet feed yields what the reference loader yields (reference lshrs/io/parquet.py)."""

from __future__ import annotations

import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")
import pyarrow.parquet as pq  # noqa: E402

from lshrs_b200.io.parquet import iter_parquet_vectors  # noqa: E402


def _write(path, ids, vectors, typ):
    table = pa.table({"index": pa.array(ids, pa.int64()), "vector": pa.array(vectors, typ)})
    pq.write_table(table, path)


@pytest.mark.parametrize("typ", [pa.list_(pa.float32()), pa.list_(pa.float64()), pa.large_list(pa.float32()),
                                 pa.list_(pa.float32(), 8)])
def test_batches_match_rows(tmp_path, typ):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((25, 8)).astype(np.float32)
    ids = list(range(100, 125))
    path = tmp_path / "v.parquet"
    _write(path, ids, X.tolist(), typ)
    got_ids, got = [], []
    for idx, vecs in iter_parquet_vectors(path, batch_size=10):
        assert isinstance(idx, list) and all(isinstance(i, int) for i in idx)
        assert vecs.dtype == np.float32 and vecs.ndim == 2 and vecs.flags.c_contiguous
        got_ids += idx
        got.append(vecs)
    assert got_ids == ids
    np.testing.assert_array_equal(np.concatenate(got), X)


def test_errors_match_the_reference(tmp_path):
    with pytest.raises(FileNotFoundError):
        next(iter_parquet_vectors(tmp_path / "missing.parquet"))
    path = tmp_path / "v.parquet"
    _write(path, [1, 2], [[1.0, 2.0], [3.0, 4.0]], pa.list_(pa.float32()))
    with pytest.raises(ValueError, match="batch_size"):
        next(iter_parquet_vectors(path, batch_size=0))
    with pytest.raises(ValueError, match="not found"):
        next(iter_parquet_vectors(path, vector_column="nope"))
    ragged = tmp_path / "r.parquet"
    _write(ragged, [1, 2], [[1.0, 2.0], [3.0]], pa.list_(pa.float32()))
    with pytest.raises(ValueError, match="same dimensionality"):
        next(iter_parquet_vectors(ragged))
    empty = tmp_path / "e.parquet"
    _write(empty, [1, 2], [[1.0], []], pa.list_(pa.float32()))
    with pytest.raises(ValueError, match="empty vector"):
        next(iter_parquet_vectors(empty))
    nulls = tmp_path / "n.parquet"
    _write(nulls, [1, 2], [[1.0, None], [3.0, 4.0]], pa.list_(pa.float32()))
    out = next(iter_parquet_vectors(nulls))[1]  # row-wise path: None -> nan, like np.asarray
    assert np.isnan(out[0, 1]) and out[1, 1] == 4.0


def test_prefetched_keeps_order_and_propagates_errors(tmp_path):
    from lshrs_b200.io.parquet import prefetched

    assert list(prefetched(iter(range(50)), depth=1)) == list(range(50))
    assert list(prefetched(iter(()))) == []

    def failing():
        yield 1
        yield 2
        raise ValueError("boom")

    it = prefetched(failing())
    assert next(it) == 1 and next(it) == 2
    with pytest.raises(ValueError, match="boom"):
        next(it)
    # abandoning the iterator early must not leave the worker blocked on a full queue
    it = prefetched(iter(range(10_000)), depth=1)
    assert next(it) == 0
    it.close()
    # the Parquet loader behind it: same batches as without
    rng = np.random.default_rng(1)
    X = rng.standard_normal((35, 4)).astype(np.float32)
    path = tmp_path / "p.parquet"
    _write(path, list(range(35)), X.tolist(), pa.list_(pa.float32()))
    plain = list(iter_parquet_vectors(path, batch_size=8))
    ahead = list(prefetched(iter_parquet_vectors(path, batch_size=8)))
    assert [i for i, _ in plain] == [i for i, _ in ahead]
    for (_, a), (_, b) in zip(plain, ahead):
        np.testing.assert_array_equal(a, b)


def test_loader_resolution():
    from lshrs_b200 import LSHRS

    assert LSHRS._resolve_loader("parquet") is iter_parquet_vectors
    assert LSHRS._resolve_loader("PQ") is iter_parquet_vectors
    with pytest.raises(ValueError, match="Unsupported"):
        LSHRS._resolve_loader("csv")
