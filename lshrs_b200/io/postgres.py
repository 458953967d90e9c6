"""PostgreSQL -> (indices, float32 matrix) batches without a Python object per element.

Same call contract as the reference loader (reference lshrs/io/postgres.py:33-141:
``iter_postgres_vectors(*, dsn | connection_factory, table, index_column, vector_column, batch_size, limit,
where_clause, order_by, params, fetch_query)`` yielding ``(list[int], ndarray[n, dim] float32)``, the same
argument errors, the same "inconsistent dimensionality" / "empty vector" errors).  The reference reads a
server-side cursor with ``fetchmany`` and turns every row into its own ndarray (``_coerce_vector``,
postgres.py:174-208) before ``np.stack`` -- one Python object per element for ``real[]`` columns, which tops
out far below what a B200 hashes.  Two transports here:

* ``COPY (<query>) TO STDOUT (FORMAT binary)`` when the driver offers ``cursor.copy`` (psycopg 3).  The
  PGCOPY stream is parsed with numpy: rows of one query have one layout (same dimensionality is required
  anyway), so a block of tuples is a fixed-stride record array -- ids and big-endian floats are sliced out
  with strides and byte-swapped in one vectorised cast.  ``real[]`` / ``double precision[]`` arrays, pgvector's
  ``vector`` and ``bytea`` holding native float32 (what ``_coerce_vector`` assumes for bytes) are understood.
* otherwise ``fetchmany`` like the reference, but a batch is coerced in bulk (one ``b"".join`` + ``frombuffer``
  for bytes-like cells, one ``np.asarray`` for sequences; text cells are parsed per row as in the reference).

Either way the batch arrives as one contiguous float32 matrix, ready for ``LSHRS.index`` /
``lshx_hash_batch``.  Host-side feeding code: no arithmetic of the hot path.
"""

from __future__ import annotations

import struct
from collections.abc import Callable, Iterator, Sequence
from typing import Any, Optional

import numpy as np

DEFAULT_POSTGRES_BATCH_SIZE = 262_144  # rows per yielded batch; the GPU wants large batches (reference: 10 000)

try:  # optional dependency, as in the reference
    import psycopg  # type: ignore[import-not-found]
except ImportError:  # pragma: no cover
    psycopg = None  # type: ignore[assignment]

__all__ = ["iter_postgres_vectors", "parse_pgcopy_binary", "DEFAULT_POSTGRES_BATCH_SIZE"]

_PGCOPY_MAGIC = b"PGCOPY\n\xff\r\n\x00"
_INCONSISTENT = ("Inconsistent vector dimensionality detected while streaming from PostgreSQL: "
                 "expected {}, received {}")


def iter_postgres_vectors(
    *,
    dsn: Optional[str] = None,
    connection_factory: Optional[Callable[[], Any]] = None,
    table: str = "vectors",
    index_column: str = "id",
    vector_column: str = "embedding",
    batch_size: int = DEFAULT_POSTGRES_BATCH_SIZE,
    limit: Optional[int] = None,
    where_clause: Optional[str] = None,
    order_by: Optional[str] = None,
    params: Optional[Sequence[Any]] = None,
    fetch_query: Optional[str] = None,
    transport: str = "auto",
) -> Iterator[tuple[list[int], np.ndarray]]:
    """Stream ``(indices, vectors)`` batches from a table or query.

    ``transport``: ``"copy"`` (binary COPY), ``"cursor"`` (``fetchmany``) or ``"auto"`` -- COPY when the cursor
    has a ``copy`` method and the query takes no parameters (COPY cannot bind any), else the cursor.
    """
    if connection_factory is None and dsn is None:
        raise ValueError("Either `dsn` or `connection_factory` must be provided")
    if fetch_query is None and params is not None:
        raise ValueError("`params` can only be used when `fetch_query` is supplied")
    if batch_size <= 0:
        raise ValueError("batch_size must be greater than zero")
    if transport not in ("auto", "copy", "cursor"):
        raise ValueError("transport must be 'auto', 'copy' or 'cursor'")
    owned = False
    if connection_factory is not None:
        connection = connection_factory()
    else:
        if psycopg is None:
            raise ImportError(
                "psycopg is required to stream data from PostgreSQL. Install it via `pip install psycopg[binary]`."
            )
        connection = psycopg.connect(dsn)
        connection.autocommit = True
        owned = True
    try:
        query, query_params = _build_query(fetch_query, table, index_column, vector_column, limit, where_clause,
                                           order_by, params)
        probe = connection.cursor()
        can_copy = hasattr(probe, "copy") and not query_params
        probe.close() if hasattr(probe, "close") else None
        if transport == "copy" and not can_copy:
            raise ValueError("transport='copy' needs a driver with cursor.copy() and a query without parameters")
        if transport == "copy" or (transport == "auto" and can_copy):
            yield from _iter_copy(connection, query, batch_size)
        else:
            yield from _iter_cursor(connection, query, query_params, batch_size)
    finally:
        if owned:
            connection.close()


def _build_query(fetch_query, table, index_column, vector_column, limit, where_clause, order_by, batch_params):
    """The reference's query (postgres.py:144-171); identifiers are quoted by psycopg when it is there."""
    if fetch_query is not None:
        return fetch_query, tuple(batch_params or ())
    if psycopg is not None:
        from psycopg import sql

        ident = lambda name: sql.Identifier(name)  # noqa: E731
        query = sql.SQL("SELECT {}, {} FROM {}").format(ident(index_column), ident(vector_column), ident(table))
        if where_clause:
            query += sql.SQL(" WHERE ") + sql.SQL(where_clause)
        if order_by:
            query += sql.SQL(" ORDER BY ") + sql.SQL(order_by)
        if limit is not None:
            query += sql.SQL(" LIMIT {}").format(sql.Literal(int(limit)))
        return query, ()
    quote = lambda name: '"' + str(name).replace('"', '""') + '"'  # noqa: E731
    text = f"SELECT {quote(index_column)}, {quote(vector_column)} FROM {quote(table)}"
    if where_clause:
        text += f" WHERE {where_clause}"
    if order_by:
        text += f" ORDER BY {order_by}"
    if limit is not None:
        text += f" LIMIT {int(limit)}"
    return text, ()


# ----------------------------------------------------------------------------------------- cursor transport
def _coerce_text(raw: str) -> np.ndarray:
    stripped = raw.strip("{}[]() ")
    if not stripped:
        raise ValueError("Encountered empty vector representation in PostgreSQL row")
    return np.fromiter((float(part) for part in stripped.split(",")), dtype=np.float32)


def _coerce_rows(cells: list) -> np.ndarray:
    """A batch of vector cells -> one float32 matrix; same per-cell meaning as the reference's _coerce_vector."""
    first = cells[0]
    if isinstance(first, (bytes, bytearray, memoryview)) and all(
            isinstance(c, (bytes, bytearray, memoryview)) for c in cells):
        sizes = {len(c) if not isinstance(c, memoryview) else c.nbytes for c in cells}
        if len(sizes) == 1:
            nbytes = sizes.pop()
            if nbytes == 0 or nbytes % 4:
                raise ValueError("Encountered empty vector while decoding PostgreSQL row")
            return np.frombuffer(b"".join(bytes(c) for c in cells), dtype=np.float32).reshape(len(cells), nbytes // 4)
    if not isinstance(first, (str, bytes, bytearray, memoryview)):
        try:
            arr = np.asarray(cells, dtype=np.float32)        # sequences of equal length: one C loop
            if arr.ndim >= 2 and arr.shape[0] == len(cells):
                arr = arr.reshape(len(cells), -1)
                if arr.shape[1] == 0:
                    raise ValueError("Encountered empty vector while decoding PostgreSQL row")
                return arr
        except (ValueError, TypeError) as exc:
            if "empty vector" in str(exc):
                raise
    rows = []
    expected = None
    for cell in cells:                                        # mixed / ragged / text cells: row by row
        if isinstance(cell, memoryview):
            vec = np.frombuffer(cell.tobytes(), dtype=np.float32)
        elif isinstance(cell, (bytes, bytearray)):
            vec = np.frombuffer(cell, dtype=np.float32)
        elif isinstance(cell, str):
            vec = _coerce_text(cell)
        else:
            vec = np.asarray(cell, dtype=np.float32).reshape(-1)
        if vec.size == 0:
            raise ValueError("Encountered empty vector while decoding PostgreSQL row")
        if expected is None:
            expected = vec.shape[0]
        elif vec.shape[0] != expected:
            raise ValueError(_INCONSISTENT.format(expected, vec.shape[0]))
        rows.append(vec)
    return np.stack(rows, axis=0).astype(np.float32, copy=False)


def _iter_cursor(connection, query, query_params, batch_size):
    with connection.cursor(name="lshrs_stream") as cursor:
        cursor.itersize = batch_size
        cursor.execute(query, query_params)
        expected_dim = None
        while True:
            rows = cursor.fetchmany(batch_size)
            if not rows:
                break
            indices = [int(row[0]) for row in rows]
            matrix = np.ascontiguousarray(_coerce_rows([row[1] for row in rows]), dtype=np.float32)
            if expected_dim is None:
                expected_dim = matrix.shape[1]
            elif matrix.shape[1] != expected_dim:
                raise ValueError(_INCONSISTENT.format(expected_dim, matrix.shape[1]))
            yield indices, matrix


# ------------------------------------------------------------------------------------------- COPY transport
class _Layout:
    """Byte layout of one tuple of the stream, learnt from the first tuple."""

    def __init__(self, buf: memoryview, pos: int) -> None:
        (nfields,) = struct.unpack_from(">h", buf, pos)
        if nfields != 2:
            raise ValueError(f"expected two columns (index, vector) in the COPY stream, found {nfields}")
        (id_len,) = struct.unpack_from(">i", buf, pos + 2)
        if id_len not in (2, 4, 8):
            raise ValueError(f"index column must be smallint / integer / bigint (binary length {id_len})")
        self.id_off, self.id_len = 6, id_len
        vpos = pos + 6 + id_len
        (vec_len,) = struct.unpack_from(">i", buf, vpos)
        if vec_len < 0:
            raise ValueError("Encountered empty vector while decoding PostgreSQL row")
        self.vec_len = vec_len
        self.payload_off = 6 + id_len + 4
        self.tuple_bytes = self.payload_off + vec_len
        if len(buf) < pos + self.tuple_bytes:
            raise struct.error("tuple not complete yet")
        payload = bytes(buf[vpos + 4: vpos + 4 + min(vec_len, 20)])
        self.kind, self.dim, self.elem, self.first, self.stride = self._classify(payload, vec_len)
        if self.dim == 0:
            raise ValueError("Encountered empty vector while decoding PostgreSQL row")
        # the bytes every tuple must share: field count, both length words (and the array / vector header)
        self.header_words = [(0, struct.pack(">h", 2)), (2, struct.pack(">i", id_len)),
                             (6 + id_len, struct.pack(">i", vec_len))]
        if self.kind in ("array", "pgvector"):
            self.header_words.append((self.payload_off, payload[: self.first]))

    @staticmethod
    def _classify(payload: bytes, vec_len: int):
        if len(payload) >= 12:
            ndim, hasnull, oid = struct.unpack_from(">iii", payload, 0)
            if ndim == 1 and hasnull in (0, 1) and oid in (700, 701) and len(payload) >= 20:
                (n,) = struct.unpack_from(">i", payload, 12)
                elem = 4 if oid == 700 else 8
                if hasnull == 0 and vec_len == 20 + n * (4 + elem):
                    return "array", n, elem, 20, 4 + elem      # per element: int32 length + big-endian value
                if hasnull:
                    raise ValueError("vector arrays with NULL elements are not supported")
            if ndim == 0 and oid in (700, 701) and vec_len == 12:
                return "array", 0, 4, 12, 8
        if len(payload) >= 4:
            n, unused = struct.unpack_from(">hh", payload, 0)
            if unused == 0 and n > 0 and vec_len == 4 + 4 * n:
                return "pgvector", n, 4, 4, 4                  # int16 dim, int16 0, dim x big-endian float4
        if vec_len % 4 == 0:
            return "bytea", vec_len // 4, 4, 0, 4              # native float32 bytes, as _coerce_vector reads them
        raise ValueError("unrecognised binary layout of the vector column")


def parse_pgcopy_binary(chunks, batch_size: int) -> Iterator[tuple[list[int], np.ndarray]]:
    """``(indices, matrix)`` batches from an iterable of byte blocks of a ``COPY ... (FORMAT binary)`` stream."""
    pending = bytearray()
    header_done = False
    layout: Optional[_Layout] = None
    ids_acc: list[np.ndarray] = []
    vec_acc: list[np.ndarray] = []
    have = 0
    finished = False

    def flush(final: bool):
        nonlocal ids_acc, vec_acc, have
        while have >= batch_size or (final and have > 0):
            ids = np.concatenate(ids_acc) if len(ids_acc) > 1 else ids_acc[0]
            vec = np.concatenate(vec_acc) if len(vec_acc) > 1 else vec_acc[0]
            take = min(batch_size, have)
            yield ids[:take].tolist(), np.ascontiguousarray(vec[:take])
            ids_acc, vec_acc = ([ids[take:]], [vec[take:]]) if take < have else ([], [])
            have -= take

    for block in chunks:
        if finished:
            break
        pending += bytes(block)
        if not header_done:
            if len(pending) < 19:
                continue
            if bytes(pending[:11]) != _PGCOPY_MAGIC:
                raise ValueError("not a PostgreSQL binary COPY stream")
            (ext,) = struct.unpack_from(">i", pending, 15)
            if len(pending) < 19 + ext:
                continue
            del pending[: 19 + ext]
            header_done = True
        data = bytes(pending)          # immutable snapshot: numpy views of it never pin the growing buffer
        pos = 0
        if layout is None:
            if len(data) >= 2 and struct.unpack_from(">h", data, 0)[0] == -1:
                finished = True        # empty result set: header followed by the trailer
                break
            try:
                layout = _Layout(memoryview(data), 0)
            except struct.error:
                continue               # the first tuple is not complete yet
        tb = layout.tuple_bytes
        n = len(data) // tb
        if n:
            rec = np.frombuffer(data, dtype=np.uint8, count=n * tb).reshape(n, tb)
            ok = np.ones(n, dtype=bool)
            for off, want in layout.header_words:
                ok &= (rec[:, off: off + len(want)] == np.frombuffer(want, np.uint8)).all(axis=1)
            good = n if ok.all() else int(np.argmin(ok))
            if good:
                r = rec[:good]
                ids = np.ascontiguousarray(r[:, layout.id_off: layout.id_off + layout.id_len]).view(
                    f">i{layout.id_len}").reshape(good).astype(np.int64)
                body = r[:, layout.payload_off + layout.first:]
                if layout.kind == "bytea":
                    vec = np.ascontiguousarray(body).view("<f4").reshape(good, layout.dim)
                else:
                    cells = body.reshape(good, layout.dim, layout.stride)[:, :, layout.stride - layout.elem:]
                    vec = np.ascontiguousarray(cells).view(f">f{layout.elem}").reshape(good, layout.dim)
                ids_acc.append(ids)
                vec_acc.append(np.ascontiguousarray(vec, dtype=np.float32))   # byte swap + cast in one pass
                have += good
                pos = good * tb
            if good < n:               # a tuple that does not share the first tuple's layout
                if struct.unpack_from(">h", data, pos)[0] == -1:
                    finished = True
                else:
                    try:
                        other_dim = _Layout(memoryview(data), pos).dim
                    except (struct.error, ValueError):
                        other_dim = "another layout"
                    raise ValueError(_INCONSISTENT.format(layout.dim, other_dim))
        if not finished and len(data) - pos >= 2 and struct.unpack_from(">h", data, pos)[0] == -1:
            finished = True
        del pending[:pos]
        yield from flush(False)
    yield from flush(True)


def _iter_copy(connection, query, batch_size):
    cursor = connection.cursor()
    try:
        if psycopg is not None and not isinstance(query, str):
            from psycopg import sql

            statement = sql.SQL("COPY ({}) TO STDOUT (FORMAT binary)").format(query)
        else:
            statement = f"COPY ({query}) TO STDOUT (FORMAT binary)"
        with cursor.copy(statement) as copy:
            def blocks():
                while True:
                    data = copy.read()
                    if not data:
                        return
                    yield data

            yield from parse_pgcopy_binary(blocks(), batch_size)
    finally:
        if hasattr(cursor, "close"):
            cursor.close()
