"""Parquet -> (indices, float32 matrix) batches without a Python object per element.

Same call contract as the reference loader (reference lshrs/io/parquet.py:47-227:
``iter_parquet_vectors(source, *, index_column, vector_column, batch_size)`` yielding
``(list[int], ndarray[n, dim] float32)``, the same errors for a missing file / column, a
non-positive batch size, empty or ragged vectors).  The reference goes Arrow -> ``to_pylist()`` ->
per-row ``np.asarray`` -> ``np.stack`` (``parquet.py:219-223, 294-320``), which is Python-object
bound at well under 1 M vectors/s -- below what one B200 hashes by three orders of magnitude.  Here a
``List`` / ``LargeList`` / ``FixedSizeList`` column of float32 (or float64) without nulls is viewed
through its Arrow value buffer, so a batch costs one reshape (plus one cast for float64), and the
matrix can go straight to ``LSHRS.index`` / ``lshx_hash_batch``.  This is host-side feeding code
(SURVEY section 8f rank 3); it contains no arithmetic of the hot path.
"""

from __future__ import annotations

from collections.abc import Iterator
from pathlib import Path

import numpy as np

try:  # optional dependency, as in the reference
    import pyarrow as pa
    import pyarrow.parquet as pq
except ImportError:  # pragma: no cover
    pa = None
    pq = None

DEFAULT_PARQUET_BATCH_SIZE = 262_144  # rows per yielded batch; the GPU wants large batches

__all__ = ["iter_parquet_vectors", "prefetched", "DEFAULT_PARQUET_BATCH_SIZE"]


def prefetched(batches: Iterator, depth: int = 2) -> Iterator:
    """Run a batch iterator one step ahead in a worker thread.

    ``LSHRS.create_signatures`` wraps the Parquet loader in this, so that decoding row group ``i + 1``
    (pyarrow releases the GIL) overlaps hashing and bucket bookkeeping of batch ``i``.  Order is kept;
    an exception in the loader is re-raised at the point of iteration; abandoning the iterator stops the
    worker at its next hand-over.
    """
    import queue
    import threading

    q: queue.Queue = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()
    done = object()

    def put(item) -> bool:
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def work() -> None:
        try:
            for item in batches:
                if not put(("item", item)):
                    return
            put((done, None))
        except BaseException as exc:  # noqa: BLE001 -- handed to the consumer
            put(("error", exc))

    worker = threading.Thread(target=work, name="lshrs-parquet-prefetch", daemon=True)
    worker.start()
    try:
        while True:
            kind, payload = q.get()
            if kind is done:
                return
            if kind == "error":
                raise payload
            yield payload
    finally:
        stop.set()


def _matrix_from_list_array(arr) -> np.ndarray:
    """Arrow list-like array of numbers -> contiguous float32 (n, dim), zero-copy where the types allow."""
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    n = len(arr)
    typ = arr.type
    is_list = pa.types.is_list(typ) or pa.types.is_large_list(typ) or pa.types.is_fixed_size_list(typ)
    if not is_list or arr.null_count or arr.values.null_count:
        return _matrix_from_rows(arr.to_pylist())
    if pa.types.is_fixed_size_list(typ):
        dim = typ.list_size
        flat = arr.flatten()
    else:
        offsets = arr.offsets.to_numpy()
        lengths = np.diff(offsets)
        if n and (lengths == 0).any():
            raise ValueError("Encountered empty vector while reading Parquet data")
        dim = int(lengths[0]) if n else 0
        if n and (lengths != dim).any():
            bad = int(lengths[lengths != dim][0])
            raise ValueError(f"All vectors must share the same dimensionality; expected {dim}, received {bad}")
        flat = arr.flatten()
    if dim == 0 and n:
        raise ValueError("Encountered empty vector while reading Parquet data")
    values = flat.to_numpy(zero_copy_only=False)
    return np.ascontiguousarray(values, dtype=np.float32).reshape(n, dim)


def _matrix_from_rows(rows) -> np.ndarray:
    """Row-by-row path for exotic column types (nulls, nested objects); same checks as the reference."""
    out = []
    dim = None
    for row in rows:
        vec = np.asarray(row, dtype=np.float32).reshape(-1)
        if vec.size == 0:
            raise ValueError("Encountered empty vector while reading Parquet data")
        if dim is None:
            dim = vec.shape[0]
        elif vec.shape[0] != dim:
            raise ValueError(f"All vectors must share the same dimensionality; expected {dim}, received {vec.shape[0]}")
        out.append(vec)
    return np.stack(out, axis=0)


def iter_parquet_vectors(source, *, index_column: str = "index", vector_column: str = "vector",
                         batch_size: int = DEFAULT_PARQUET_BATCH_SIZE) -> Iterator[tuple[list[int], np.ndarray]]:
    """Stream ``(indices, vectors)`` batches from a Parquet file."""
    if pq is None:
        raise ImportError("pyarrow is required to stream vectors from Parquet files. Install it via `pip install pyarrow`.")
    path = Path(source).expanduser()
    if not path.exists():
        raise FileNotFoundError(f"Parquet source '{path}' does not exist")
    if batch_size <= 0:
        raise ValueError("batch_size must be greater than zero")
    pf = pq.ParquetFile(path)
    schema = pf.schema_arrow
    for column in (index_column, vector_column):
        if schema.get_field_index(column) == -1:
            raise ValueError(f"Column '{column}' was not found in Parquet schema {schema.names}")
    for batch in pf.iter_batches(batch_size=batch_size, columns=[index_column, vector_column]):
        if batch.num_rows == 0:
            continue
        indices = batch.column(batch.schema.get_field_index(index_column)).to_numpy(zero_copy_only=False)
        vectors = _matrix_from_list_array(batch.column(batch.schema.get_field_index(vector_column)))
        yield [int(i) for i in indices.tolist()], vectors
