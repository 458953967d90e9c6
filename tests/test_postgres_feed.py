"""PostgreSQL feed (second half): binary COPY and bulk-coerced cursor batches.

No server exists in the build container or on the GPU box, so the transports are driven by fake DB-API
connections: one that serves ``cursor.copy()`` blocks of a PGCOPY binary stream (encoded here, following the
PostgreSQL documentation of the format), one that only has a named cursor with ``fetchmany``.  Where the staged
reference is present, the cursor transport is compared batch for batch with the reference's own loader
(reference lshrs/io/postgres.py:33-141) on the same fake rows.
"""

from __future__ import annotations

import struct
import sys
import types
from pathlib import Path

import numpy as np
import pytest

from lshrs_b200.io.postgres import iter_postgres_vectors, parse_pgcopy_binary

REPO = Path(__file__).resolve().parents[1]


# ---------------------------------------------------------------------------------------------- stream encoders
def _field_array(vec, oid=700):
    fmt, size = (">f", 4) if oid == 700 else (">d", 8)
    out = struct.pack(">iiiii", 1, 0, oid, len(vec), 1)
    for v in vec:
        out += struct.pack(">i", size) + struct.pack(fmt, float(v))
    return out


def _field_pgvector(vec):
    return struct.pack(">hh", len(vec), 0) + b"".join(struct.pack(">f", float(v)) for v in vec)


def _field_bytea(vec):
    return np.asarray(vec, dtype="<f4").tobytes()


def pgcopy_stream(ids, X, *, field=_field_array, id_fmt=">i", header_ext=b"", trailer=True) -> bytes:
    out = b"PGCOPY\n\xff\r\n\x00" + struct.pack(">ii", 0, len(header_ext)) + header_ext
    id_size = struct.calcsize(id_fmt)
    for i, row in zip(ids, X):
        payload = field(row)
        out += struct.pack(">h", 2) + struct.pack(">i", id_size) + struct.pack(id_fmt, int(i))
        out += struct.pack(">i", len(payload)) + payload
    if trailer:
        out += struct.pack(">h", -1)
    return out


def _chunks(data: bytes, size: int):
    return [data[i:i + size] for i in range(0, len(data), size)]


# ---------------------------------------------------------------------------------------------- fake connections
class _Copy:
    def __init__(self, blocks):
        self.blocks = list(blocks)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def read(self):
        return self.blocks.pop(0) if self.blocks else b""


class _CopyCursor:
    def __init__(self, conn):
        self.conn = conn

    def copy(self, statement):
        self.conn.statements.append(str(statement))
        return _Copy(self.conn.blocks)

    def close(self):
        pass


class CopyConnection:
    def __init__(self, blocks):
        self.blocks, self.statements, self.closed = blocks, [], False

    def cursor(self, name=None):
        return _CopyCursor(self)

    def close(self):
        self.closed = True


class _RowCursor:
    def __init__(self, conn):
        self.conn, self.itersize, self.pos = conn, None, 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def execute(self, query, params=()):
        self.conn.executed.append((str(query), tuple(params)))

    def fetchmany(self, n):
        rows = self.conn.rows[self.pos:self.pos + n]
        self.pos += len(rows)
        return rows


class RowConnection:
    """Only a (named) cursor with fetchmany -- the surface the reference's loader uses."""

    def __init__(self, rows):
        self.rows, self.executed = rows, []

    def cursor(self, name=None):
        return _RowCursor(self)


# ---------------------------------------------------------------------------------------------- COPY transport
@pytest.mark.parametrize("field, oid", [(_field_array, 700), (_field_array, 701), (_field_pgvector, None),
                                         (_field_bytea, None)])
@pytest.mark.parametrize("block", [7, 4096, 1 << 20])
def test_binary_copy_stream_to_matrix(field, oid, block):
    rng = np.random.default_rng(0)
    n, dim = 257, 48
    X = rng.standard_normal((n, dim)).astype(np.float32)
    X[3, 5], X[9, 0] = np.float32(-0.0), np.float32(1e-42)     # signed zero and a denormal survive the byte swap
    ids = (np.arange(n) * 3 + 11).tolist()
    enc = (lambda v: _field_array(v, oid)) if oid else field
    stream = pgcopy_stream(ids, X, field=enc, id_fmt=">q" if block == 7 else ">i", header_ext=b"xyz" if block == 7 else b"")
    conn = CopyConnection(_chunks(stream, block))
    got = list(iter_postgres_vectors(connection_factory=lambda: conn, fetch_query="SELECT id, embedding FROM t",
                                     batch_size=100))
    assert [len(i) for i, _ in got] == [100, 100, 57]
    assert conn.statements == ["COPY (SELECT id, embedding FROM t) TO STDOUT (FORMAT binary)"]
    assert sum((i for i, _ in got), []) == ids
    M = np.concatenate([m for _, m in got])
    assert M.dtype == np.float32 and all(m.flags.c_contiguous for _, m in got)
    np.testing.assert_array_equal(M.view(np.uint32), X.view(np.uint32))     # bit for bit (float64 cells round-trip)


def test_binary_copy_edge_cases():
    X = np.arange(12, dtype=np.float32).reshape(3, 4)
    # empty result set
    assert list(parse_pgcopy_binary([pgcopy_stream([], X[:0])], 10)) == []
    # one row, whole stream in one block, batch larger than the stream
    (ids, M), = list(parse_pgcopy_binary([pgcopy_stream([5], X[:1])], 10))
    assert ids == [5] and M.tolist() == X[:1].tolist()
    # ragged dimensionality -> the reference's error (postgres.py:131-135)
    bad = pgcopy_stream([1, 2], X[:2], trailer=False) + pgcopy_stream([3], np.ones((1, 5), np.float32))[19:]
    with pytest.raises(ValueError, match="Inconsistent vector dimensionality.*expected 4, received 5"):
        list(parse_pgcopy_binary(_chunks(bad, 50), 10))
    # NULL vector / empty array / not a COPY stream
    null_vec = b"PGCOPY\n\xff\r\n\x00" + struct.pack(">ii", 0, 0) + struct.pack(">hiii", 2, 4, 1, -1)
    with pytest.raises(ValueError, match="empty vector"):
        list(parse_pgcopy_binary([null_vec + b"\x00" * 32], 10))
    empty = pgcopy_stream([1], [[]], field=lambda v: struct.pack(">iii", 0, 0, 700))
    with pytest.raises(ValueError, match="empty vector"):
        list(parse_pgcopy_binary([empty], 10))
    with pytest.raises(ValueError, match="not a PostgreSQL binary COPY"):
        list(parse_pgcopy_binary([b"x" * 64], 10))


def test_argument_errors_match_the_reference():
    with pytest.raises(ValueError, match="Either `dsn` or `connection_factory`"):
        list(iter_postgres_vectors())
    with pytest.raises(ValueError, match="`params` can only be used"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection([]), params=[1]))
    with pytest.raises(ValueError, match="batch_size must be greater than zero"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection([]), batch_size=0))
    with pytest.raises(ValueError, match="transport='copy'"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection([]), transport="copy"))


# ---------------------------------------------------------------------------------------------- cursor transport
def _fake_rows(kind, n=53, dim=6, seed=1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    cell = {
        "list": lambda v: [float(x) for x in v],
        "bytes": lambda v: v.tobytes(),
        "memoryview": lambda v: memoryview(v.tobytes()),
        "text": lambda v: "{" + ",".join(repr(float(x)) for x in v) + "}",
        "pgvector_text": lambda v: "[" + ",".join(repr(float(x)) for x in v) + "]",
        "ndarray": lambda v: v.astype(np.float64),
    }[kind]
    return [(100 + i, cell(X[i])) for i in range(n)], X


@pytest.mark.parametrize("kind", ["list", "bytes", "memoryview", "text", "pgvector_text", "ndarray"])
def test_cursor_transport_equals_reference_loader(kind):
    rows, X = _fake_rows(kind)
    conn = RowConnection(rows)
    got = list(iter_postgres_vectors(connection_factory=lambda: conn, table="t", batch_size=20))
    assert [len(i) for i, _ in got] == [20, 20, 13]
    assert sum((i for i, _ in got), []) == [r[0] for r in rows]
    np.testing.assert_array_equal(np.concatenate([m for _, m in got]), X)
    assert all(m.dtype == np.float32 and m.flags.c_contiguous for _, m in got)
    assert 'SELECT "id", "embedding" FROM "t"' in conn.executed[0][0]
    ref_root = REPO / "oracle" / "_ref" / "reference"
    if not (ref_root / "lshrs" / "io" / "postgres.py").exists():
        return
    # the reference's own loader on the same rows (stub psycopg: it only checks that the module exists and
    # builds the default query with psycopg.sql, so a fetch_query is used)
    import importlib.util

    saved = sys.modules.get("psycopg")
    sys.modules["psycopg"] = types.ModuleType("psycopg")
    try:
        spec = importlib.util.spec_from_file_location("ref_postgres_loader", ref_root / "lshrs" / "io" / "postgres.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        want = list(ref.iter_postgres_vectors(connection_factory=lambda: RowConnection(rows), fetch_query="q",
                                              batch_size=20))
    finally:
        if saved is None:
            sys.modules.pop("psycopg", None)
        else:
            sys.modules["psycopg"] = saved
    assert len(want) == len(got)
    for (gi, gm), (wi, wm) in zip(got, want):
        assert gi == wi
        np.testing.assert_array_equal(gm, wm)


def test_cursor_transport_errors():
    rows, _ = _fake_rows("list")
    rows[30] = (rows[30][0], [1.0, 2.0])
    with pytest.raises(ValueError, match="Inconsistent vector dimensionality"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection(rows), batch_size=40))
    with pytest.raises(ValueError, match="empty vector"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection([(1, "{}")]), batch_size=4))
    with pytest.raises(ValueError, match="empty vector"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection([(1, b"")]), batch_size=4))
    # dimensionality changing BETWEEN batches is caught as well
    rows2, _ = _fake_rows("bytes", n=8)
    rows2[5] = (rows2[5][0], np.ones(3, np.float32).tobytes())
    rows2[6] = (rows2[6][0], np.ones(3, np.float32).tobytes())
    rows2[7] = (rows2[7][0], np.ones(3, np.float32).tobytes())
    with pytest.raises(ValueError, match="Inconsistent vector dimensionality"):
        list(iter_postgres_vectors(connection_factory=lambda: RowConnection(rows2), batch_size=5))


def test_create_signatures_takes_the_postgres_feed():
    """LSHRS.create_signatures(format="postgres") -> iter_postgres_vectors -> index(): ops as from a plain index()."""
    sys.path.insert(0, str(REPO / "tests"))
    import fake_lshx

    from lshrs_b200 import LSHRS, InMemoryStorage, _native

    saved = _native._lib
    fake_lshx.install()
    try:
        rng = np.random.default_rng(2)
        X = rng.standard_normal((300, 32)).astype(np.float32)
        ids = list(range(1000, 1300))
        a = LSHRS(dim=32, num_perm=16, num_bands=4, rows_per_band=4, storage=InMemoryStorage())
        a.create_signatures(format="postgres", connection_factory=lambda: CopyConnection(
            _chunks(pgcopy_stream(ids, X), 3000)), fetch_query="SELECT id, embedding FROM t", batch_size=128)
        b = LSHRS(dim=32, num_perm=16, num_bands=4, rows_per_band=4, storage=InMemoryStorage())
        b.index(ids, X)
        assert a._storage._bands == b._storage._bands and len(a._storage) > 0
        a._hasher.close()      # handles of the double must not reach the real library's destroy
        b._hasher.close()
    finally:
        _native._lib = saved
