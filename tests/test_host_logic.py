"""Host-side logic that needs no GPU: value types, storage, validation order, sharding plan."""

from __future__ import annotations

import numpy as np
import pytest

from lshrs_b200 import LSHRS, HashSignatures, InMemoryStorage, LSHHasher, bucket_key
from lshrs_b200._config.config import signatures_from_packed
from lshrs_b200.sharding import shard_bounds, weighted_bounds


def test_hash_signatures_normalizes_iterables():
    # reference tests/test_lshrs.py:92-97
    sig = HashSignatures((b"\x00\x01", bytearray(b"\x02\x03")))  # type: ignore[arg-type]
    assert sig.as_tuple() == (b"\x00\x01", b"\x02\x03")
    assert list(sig) == [b"\x00\x01", b"\x02\x03"]
    assert len(sig) == 2 and sig[1] == b"\x02\x03"
    assert hash(sig) == hash(HashSignatures((b"\x00\x01", b"\x02\x03")))
    with pytest.raises(Exception):
        sig.bands = ()  # frozen


def test_signatures_from_packed_layout():
    packed = np.arange(2 * 3 * 2, dtype=np.uint8).reshape(2, 3, 2)
    sigs = signatures_from_packed(packed, 2)
    assert sigs[0].as_tuple() == (b"\x00\x01", b"\x02\x03", b"\x04\x05")
    assert sigs[1].as_tuple() == (b"\x06\x07", b"\x08\x09", b"\x0a\x0b")
    assert HashSignatures.from_packed(packed[1].reshape(-1), 2) == sigs[1]


@pytest.mark.parametrize("nb, r, dim", [(0, 1, 1), (1, 0, 1), (1, 1, 0)])
def test_lsh_hasher_invalid_init_parameters(nb, r, dim):
    # reference tests/test_lshrs.py:18-28
    with pytest.raises(ValueError):
        LSHHasher(num_bands=nb, rows_per_band=r, dim=dim)


def test_hasher_projections_are_the_reference_stream():
    from oracle import lshrs_oracle as oracle

    h = LSHHasher(16, 16, 768, seed=42)
    want = oracle.make_projections(16, 16, 768, 42)
    assert len(h.projections) == 16
    for a, b in zip(h.projections, want):
        assert a.dtype == np.float32 and a.shape == (16, 768)
        np.testing.assert_array_equal(a, b)
    assert h.bytes_per_band == 2 and h.signature_bytes == 32
    assert LSHHasher(16, 4, 128).signature_bytes == 16 and LSHHasher(3, 20, 5).bytes_per_band == 3


def test_hasher_validation_runs_before_any_device_work():
    h = LSHHasher(2, 3, 4)
    with pytest.raises(ValueError):
        h.hash_vector(np.arange(5, dtype=np.float32))
    with pytest.raises(ValueError, match="2D"):
        h.hash_batch(np.arange(3, dtype=np.float32))
    with pytest.raises(ValueError):
        h.hash_batch(np.ones((2, 5), dtype=np.float32))
    assert h.hash_batch(np.empty((0, 4), dtype=np.float32)) == []
    assert h.hash_batch_packed(np.empty((0, 4), dtype=np.float32)).shape == (0, 2, 1)


def test_in_memory_storage_and_bucket_key():
    st = InMemoryStorage(prefix="lsh")
    assert bucket_key("lsh", 5, b"\xab\xcd") == "lsh:5:bucket:abcd" == st.bucket_key(5, b"\xab\xcd")
    st.batch_add([(0, b"\x01", 7), (0, b"\x01", 8), (1, b"\x01", 7)])
    assert st.get_bucket(0, b"\x01") == {7, 8} and st.get_bucket(1, b"\x01") == {7}
    assert st.get_bucket(2, b"\x01") == set()
    st.remove_indices([7])
    assert st.get_bucket(0, b"\x01") == {8} and st.get_bucket(1, b"\x01") == set()
    st.clear()
    assert st.get_bucket(0, b"\x01") == set()


def _lsh(**kw):
    args = dict(dim=32, num_perm=16, num_bands=4, rows_per_band=4, storage=InMemoryStorage())
    args.update(kw)
    return LSHRS(**args)


def test_lshrs_constructor_validation():
    # reference tests/test_core.py:16-35
    with pytest.raises(ValueError, match="dimensionality"):
        _lsh(dim=0)
    with pytest.raises(ValueError, match="num_perm"):
        _lsh(num_perm=0)
    with pytest.raises(ValueError, match="buffer_size"):
        _lsh(buffer_size=0)
    with pytest.raises(ValueError, match="must equal num_perm"):
        _lsh(num_bands=3, rows_per_band=4, num_perm=16)
    auto = LSHRS(dim=768, num_perm=256, storage=InMemoryStorage())
    assert (auto.stats()["num_bands"], auto.stats()["rows_per_band"]) == (16, 16)


def test_lshrs_rejects_bad_input_before_hashing():
    lsh = _lsh()
    with pytest.raises(ValueError, match="non-negative"):
        lsh.ingest(-1, np.ones(32, dtype=np.float32))
    with pytest.raises(ValueError, match="Cannot index zero vector"):
        lsh.ingest(0, np.zeros(32, dtype=np.float32))
    with pytest.raises(ValueError, match="dimension"):
        lsh.ingest(0, np.ones(31, dtype=np.float32))
    with pytest.raises(ValueError, match="zero vector"):
        lsh.get_top_k(np.zeros(32, dtype=np.float32))
    with pytest.raises(ValueError, match="shape"):
        lsh.index([0, 1], np.ones((2, 31), dtype=np.float32))
    with pytest.raises(ValueError, match="does not match"):
        lsh.index([0, 1, 2], np.ones((2, 32), dtype=np.float32))
    with pytest.raises(RuntimeError, match="vector_fetch_fn"):
        lsh.index([0, 1])
    lsh.index([])  # no-op
    assert lsh.stats()["dimension"] == 32


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 127, 128, 129, 1000, 100_000, 12_500_001):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and lo <= hi
            assert all(lo % 128 == 0 for lo, _ in b if lo < n)
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 128 + (128 - 1)


def test_save_load_and_pickle_round_trip(tmp_path):
    # reference tests/test_persistence_security.py:15-70, 114-132 (formats are the reference's)
    import json
    import pickle

    lsh = _lsh(redis_password="secret", seed=7)
    lsh._hasher.projections = [m + 1.0 for m in lsh._hasher.projections]  # state that only a save carries
    lsh.save_to_disk(tmp_path / "idx")
    meta = json.loads((tmp_path / "idx" / "metadata.json").read_text())
    assert meta["redis_config"]["password"] == "<REDACTED>" and meta["config"]["seed"] == 7
    with np.load(tmp_path / "idx" / "projections.npz") as data:
        assert sorted(data.files) == ["arr_0", "arr_1", "arr_2", "arr_3"]
    back = LSHRS.load_from_disk(tmp_path / "idx", storage=InMemoryStorage())
    assert back.stats() == lsh.stats()
    for a, b in zip(back._hasher.projections, lsh._hasher.projections):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(FileNotFoundError):
        LSHRS.load_from_disk(tmp_path / "missing")
    clone = pickle.loads(pickle.dumps(lsh))
    assert clone.stats() == lsh.stats()
    for a, b in zip(clone._hasher.projections, lsh._hasher.projections):
        np.testing.assert_array_equal(a, b)


def test_fetch_buckets_strategies():
    """query_batch asks for nq * num_bands buckets: one call / one pipelined round trip / per-key fallback."""
    keys = [(0, b"\x01"), (1, b"\x02"), (0, b"\xff")]
    mem = InMemoryStorage()
    mem.batch_add([(0, b"\x01", 5), (1, b"\x02", 6), (1, b"\x02", 7)])
    assert _lsh(storage=mem)._fetch_buckets(keys) == [{5}, {6, 7}, set()]

    class FakePipe:
        def __init__(self, data, log):
            self.data, self.log, self.queued = data, log, []

        def smembers(self, key):
            self.queued.append(key)

        def execute(self):
            self.log.append(len(self.queued))
            return [self.data.get(k, set()) for k in self.queued]

        def reset(self):
            self.queued = []

    class FakeRedisStorage:  # the surface of the reference's RedisStorage that _fetch_buckets touches
        prefix = "lsh"

        def __init__(self):
            self.round_trips = []
            data = {"lsh:0:bucket:01": {b"5"}, "lsh:1:bucket:02": {b"6", b"7"}}
            outer = self

            class Client:
                def pipeline(self):
                    return FakePipe(data, outer.round_trips)

            self._client = Client()

        def bucket_key(self, band_id, hash_val):
            return f"{self.prefix}:{band_id}:bucket:{hash_val.hex()}"

        def batch_add(self, ops): ...
        def get_bucket(self, b, h): raise AssertionError("per-key path must not be used")
        def remove_indices(self, i): ...
        def clear(self): ...
        def close(self): ...

    fake = FakeRedisStorage()
    assert _lsh(storage=fake)._fetch_buckets(keys) == [{5}, {6, 7}, set()]
    assert fake.round_trips == [3]  # one pipelined round trip

    class PerKey:
        def __init__(self):
            self.calls = 0

        def get_bucket(self, b, h):
            self.calls += 1
            return {b}

        def batch_add(self, ops): ...
        def remove_indices(self, i): ...
        def clear(self): ...
        def close(self): ...

    pk = PerKey()
    assert _lsh(storage=pk)._fetch_buckets(keys) == [{0}, {1}, {0}] and pk.calls == 3


# ---------------------------------------------------------------------------------------------------------------
# device mirror of the bucket store (host logic; the C ABI is the oracle-backed double here, the kernels are
# checked in tests/test_index_gpu.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def cpu_double():
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import fake_lshx

    from lshrs_b200 import _native

    saved = _native._lib
    fake = fake_lshx.install()
    made: list = []
    yield fake, made
    for lsh in made:                       # handles of the double must never reach the real library
        lsh._hasher.close()
        if lsh._dindex is not None:
            lsh._dindex.close()
    from lshrs_b200.utils import similarity

    for r in list(similarity._rerankers.values()):
        r.close()
    similarity._rerankers.clear()
    _native._lib = saved


def test_device_mirror_follows_the_bucket_store(cpu_double):
    """What query_batch(device_index=True) sees is what the storage path sees: buffered (unflushed) operations are
    in neither, flushed ones in both; delete / clear reach both; results are identical."""
    fake, made = cpu_double
    rng = np.random.default_rng(0)
    n, dim = 400, 16
    centers = rng.standard_normal((n // 4, dim)).astype(np.float32)
    X = (np.repeat(centers, 4, axis=0) + 0.05 * rng.standard_normal((n, dim))).astype(np.float32)
    lsh = LSHRS(dim=dim, num_perm=16, num_bands=8, rows_per_band=2, storage=InMemoryStorage(), buffer_size=1000,
                vector_fetch_fn=lambda ids: X[np.asarray(ids, dtype=np.int64)], device_index=True)
    made.append(lsh)
    lsh.index(list(range(300)), X[:300])
    for i in range(300, 320):
        lsh.ingest(i, X[i])                # 20 x 8 = 160 operations stay in the buffer
    Q = X[rng.integers(0, 320, 64)] + 0.01 * rng.standard_normal((64, dim)).astype(np.float32)

    def same(**kw):
        a = lsh.query_batch(Q, **kw)
        b = lsh.query_batch(Q, device_index=True, **kw)
        assert a == b
        return a

    before = same(top_k=None)
    assert all(i < 300 for ids in before for i in ids)          # the buffered 20 are invisible to both paths
    lsh.flush()
    after = same(top_k=None)
    assert any(i >= 300 for ids in after for i in ids)
    same(top_k=5)
    same(top_k=3, top_p=0.5)
    same(top_k=None, top_p=1.0)
    ids, counts = lsh.query_batch(Q, top_k=5, device_index=True, as_arrays=True)
    assert ids.shape == (64, 5) and [row[:c].tolist() for row, c in zip(ids, counts)] == same(top_k=5)
    lsh.delete([int(after[0][0]), 7])
    gone = same(top_k=None)
    assert int(after[0][0]) not in gone[0] and all(7 not in ids for ids in gone)
    lsh.index([7], X[7:8])                                       # a removed id comes back
    assert any(7 in ids for ids in same(top_k=None))
    lsh.clear()
    assert same(top_k=None) == [[] for _ in range(64)]
    # rows before an invalid one are buffered, flushed by the next flush, and reach the mirror with it
    bad = X[:10].copy()
    bad[6] = 0
    with pytest.raises(ValueError, match="zero vector"):
        lsh.index(list(range(10)), bad)
    lsh.flush()
    assert sorted({i for ids in same(top_k=None) for i in ids}) == [i for i in range(6)
                                                                    if any(i in ids for ids in same(top_k=None))]
    with pytest.raises(RuntimeError, match="device_index=True needs"):
        LSHRS(dim=dim, num_perm=16, num_bands=8, rows_per_band=2, storage=InMemoryStorage()).query_batch(
            Q, device_index=True)


def test_device_mirror_flush_failure_keeps_both_sides_unflushed(cpu_double):
    fake, made = cpu_double

    class Flaky(InMemoryStorage):
        fail = True

        def batch_add(self, operations):
            if self.fail:
                raise ConnectionError("down")
            super().batch_add(operations)

    st = Flaky()
    lsh = LSHRS(dim=8, num_perm=8, num_bands=4, rows_per_band=2, storage=st, device_index=True)
    made.append(lsh)
    X = np.random.default_rng(1).standard_normal((5, 8)).astype(np.float32)
    with pytest.raises(ConnectionError):
        lsh.index(list(range(5)), X)
    assert len(lsh._dindex) == 0 and len(lsh._buffer) == 20 and len(lsh._mirror_pending) == 1
    st.fail = False
    lsh.flush()
    assert len(lsh._dindex) == 5 and not lsh._mirror_pending
    got = lsh.query_batch(X, top_k=None, device_index=True)
    assert all(i in got[i] for i in range(5)) and got == lsh.query_batch(X, top_k=None)


def test_weighted_bounds_cover_and_align():
    b = weighted_bounds(1_000_003, [23.5] * 4 + [36.0] * 4)
    assert b[0][0] == 0 and b[-1][1] == 1_000_003 and all(b[i][1] == b[i + 1][0] for i in range(7))
    assert all((hi - lo) % 128 == 0 for lo, hi in b[:-1])
    sizes = [hi - lo for lo, hi in b]
    assert abs(sizes[4] / sizes[0] - 36.0 / 23.5) < 0.01
    assert weighted_bounds(0, [1, 2]) == [(0, 0), (0, 0)]
    assert sum(hi - lo for lo, hi in weighted_bounds(100, [1, 1, 1])) == 100
    with pytest.raises(ValueError):
        weighted_bounds(10, [1, 0])
