"""``DeviceIndex`` -- a mirror of the bucket store in HBM for batched candidate generation.

The reference answers a query with one ``SMEMBERS`` per band and counts collisions in a Python dict
(``LSHRS._candidate_counts``, reference lshrs/core/main.py:1088-1111; ordering ``(-collisions, id)``, main.py:614).
For a BATCH of queries that walk is the bottleneck long before the kernels are, so ``LSHRS(device_index=True)``
keeps the ``(band, key, id)`` triples it sends to the bucket store in a sorted device index as well
(``lshx_index_*``, csrc/index_join.cu) and ``query_batch(..., device_index=True)`` joins the whole batch there and
hands the candidate lists to the rerank kernel without leaving the GPU.  The bucket store (Redis) stays the system
of record: every operation still goes to it unchanged; the mirror is only as current as the ``LSHRS`` object that
feeds it.  Integer work -- lists, order and collision counts equal the storage path's exactly (tested against it).
"""

from __future__ import annotations

import ctypes
import threading

import numpy as np

from lshrs_b200 import _native

__all__ = ["DeviceIndex", "DeviceBucketStorage"]


class DeviceIndex:
    def __init__(self, num_bands: int, bytes_per_band: int, device: int | None = None) -> None:
        self.num_bands, self.bytes_per_band = int(num_bands), int(bytes_per_band)
        self.device = _native.default_device() if device is None else int(device)
        handle = ctypes.c_void_p()
        _native.check(_native.lib().lshx_index_create(self.device, self.num_bands, self.bytes_per_band,
                                                      ctypes.byref(handle)))
        self._handle = handle
        self._nq = 0
        self._total = 0
        # a query's result stays in the handle until fetch / topk / rerank consume it: callers that share an
        # index between threads hold this around the pair
        self.lock = threading.RLock()

    # ---- mirror of the storage operations
    def add(self, signatures: np.ndarray, ids) -> None:
        """``signatures``: uint8 ``(n, num_bands, bytes_per_band)`` as ``hash_batch_packed`` returns them."""
        sig = np.ascontiguousarray(signatures, dtype=np.uint8)
        idx = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        n = idx.shape[0]
        if sig.size != n * self.num_bands * self.bytes_per_band:
            raise ValueError(f"signatures of {sig.shape} do not match {n} ids x {self.num_bands} x {self.bytes_per_band}")
        if n:
            with self.lock:
                _native.check(_native.lib().lshx_index_add(self._handle, sig.ctypes.data, idx.ctypes.data, n, 0, None))

    def add_device(self, signatures, ids, stream=None) -> None:
        """Same with CUDA tensors (uint8 signatures / int64 ids resident in HBM): nothing crosses PCIe."""
        n = int(ids.numel())
        if n:
            with self.lock:
                _native.check(_native.lib().lshx_index_add(
                    self._handle, int(signatures.data_ptr()), int(ids.data_ptr()), n, 1,
                    ctypes.c_void_p(stream) if stream else None))

    def add_entries(self, signatures: np.ndarray, ids_per_band: np.ndarray) -> None:
        """Single ``(band, key, id)`` entries: ``ids_per_band`` int64 ``(n, num_bands)``, -1 = no entry in that band."""
        sig = np.ascontiguousarray(signatures, dtype=np.uint8)
        idx = np.ascontiguousarray(ids_per_band, dtype=np.int64).reshape(-1, self.num_bands)
        n = idx.shape[0]
        if sig.size != n * self.num_bands * self.bytes_per_band:
            raise ValueError(f"signatures of {sig.shape} do not match {n} rows x {self.num_bands} x {self.bytes_per_band}")
        if n:
            with self.lock:
                _native.check(_native.lib().lshx_index_add_entries(self._handle, sig.ctypes.data, idx.ctypes.data, n))

    def get_buckets(self, band_ids, keys: np.ndarray):
        """Members of ``m`` buckets: ``(offsets int64[m + 1], ids int64[total])``, bucket t = ``ids[offsets[t]:offsets[t+1]]``
        (live ids, each once, ascending).  ``keys``: uint8 ``(m, bytes_per_band)``."""
        bands = np.ascontiguousarray(band_ids, dtype=np.int32).reshape(-1)
        k = np.ascontiguousarray(keys, dtype=np.uint8).reshape(-1, self.bytes_per_band)
        m = bands.shape[0]
        if k.shape[0] != m:
            raise ValueError(f"{k.shape[0]} keys for {m} band ids")
        offs = np.zeros(m + 1, dtype=np.int64)
        if m == 0:
            return offs, np.empty(0, dtype=np.int64)
        lib, need = _native.lib(), ctypes.c_int64(0)
        with self.lock:     # sizes, then members: nothing may be added in between
            _native.check(lib.lshx_index_get_buckets(self._handle, bands.ctypes.data, k.ctypes.data, m, offs.ctypes.data,
                                                     None, 0, ctypes.byref(need)))
            ids = np.empty(max(1, int(need.value)), dtype=np.int64)
            if need.value:
                _native.check(lib.lshx_index_get_buckets(self._handle, bands.ctypes.data, k.ctypes.data, m,
                                                         offs.ctypes.data, ids.ctypes.data, ids.shape[0], ctypes.byref(need)))
            else:
                offs[:] = 0
        return offs, ids[: int(offs[-1])]

    def export(self):
        """``(keys uint8[num_bands, n, bytes_per_band], ids int64[num_bands, n])`` of everything held (-1 = removed)."""
        lib, n = _native.lib(), ctypes.c_int64(0)
        with self.lock:
            _native.check(lib.lshx_index_export(self._handle, None, None, 0, ctypes.byref(n)))
            keys = np.zeros((self.num_bands, int(n.value), self.bytes_per_band), dtype=np.uint8)
            ids = np.full((self.num_bands, int(n.value)), -1, dtype=np.int64)
            if n.value:
                _native.check(lib.lshx_index_export(self._handle, keys.ctypes.data, ids.ctypes.data, int(n.value),
                                                    ctypes.byref(n)))
        return keys, ids

    def remove(self, ids) -> None:
        idx = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        if idx.shape[0]:
            with self.lock:
                _native.check(_native.lib().lshx_index_remove(self._handle, idx.ctypes.data, idx.shape[0]))

    def clear(self) -> None:
        with self.lock:
            _native.check(_native.lib().lshx_index_clear(self._handle))

    def __len__(self) -> int:
        return int(_native.lib().lshx_index_size(self._handle))

    # ---- batched queries
    def query(self, signatures: np.ndarray) -> tuple[int, int]:
        """Join ``nq`` query signatures against the index; returns ``(candidate slots, largest list bound)``."""
        sig = np.ascontiguousarray(signatures, dtype=np.uint8)
        nq = sig.size // (self.num_bands * self.bytes_per_band)
        total, maxc = ctypes.c_int64(0), ctypes.c_int64(0)
        _native.check(_native.lib().lshx_index_query(self._handle, sig.ctypes.data, nq, 0, None,
                                                     ctypes.byref(total), ctypes.byref(maxc)))
        self._nq, self._total = nq, int(total.value)
        return self._total, int(maxc.value)

    SMALL_MAX_QUERIES, SMALL_MAX_CAPACITY = 32, 4096

    def query_vectors(self, hasher, vectors: np.ndarray, capacity: int):
        """Latency path (``lshx_index_query_vectors``): hash up to 32 host vectors with ``hasher`` and join them in
        two launches and one synchronisation.  Returns ``(ids int64[nq, capacity], collisions int32[nq, capacity],
        counts int32[nq], zero uint8[nq])``; ``counts[i]`` is the FULL list length (the arrays hold its first
        ``capacity`` entries) or -1 when query i matches too many bucket entries for this path."""
        x = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, hasher.dim)
        nq, cap = x.shape[0], int(capacity)
        ids = np.empty((nq, cap), dtype=np.int64)
        coll = np.empty((nq, cap), dtype=np.int32)
        counts = np.zeros(nq, dtype=np.int32)
        zero = np.zeros(nq, dtype=np.uint8)
        if nq:
            with self.lock:
                _native.check(_native.lib().lshx_index_query_vectors(
                    self._handle, hasher._ensure_handle(), x.ctypes.data, nq, cap, ids.ctypes.data, coll.ctypes.data,
                    counts.ctypes.data, zero.ctypes.data))
        return ids, coll, counts, zero

    SMALL_RERANK_CAPACITY = 1024

    def query_rerank_vectors(self, hasher, reranker, vectors: np.ndarray, corpus, *, k: int = 0, p: float = 0.0,
                             stride: int):
        """Latency path with the rerank fused in (``lshx_index_query_rerank_vectors``): hash, join, cosine rerank
        against the CUDA tensor ``corpus`` (candidate id = row) and id gather in four launches and one
        synchronisation.  Returns ``(ids int64[nq, stride], scores float32[nq, stride], counts int32[nq],
        zero int32[nq], candidates int32[nq], zero_flag uint8[nq])``; ``candidates[i]`` is -1 when query i matches
        more than 1024 bucket entries (nothing was ranked: take ``query`` + ``rerank``)."""
        x = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, hasher.dim)
        nq, stride = x.shape[0], int(stride)
        ids = np.full((nq, stride), -1, dtype=np.int64)
        scores = np.zeros((nq, stride), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        zero = np.zeros(nq, dtype=np.int32)
        cands = np.zeros(nq, dtype=np.int32)
        flag = np.zeros(nq, dtype=np.uint8)
        if nq:
            if corpus.dim() != 2 or corpus.shape[1] != hasher.dim or not corpus.is_contiguous():
                raise ValueError(f"corpus must be a contiguous (n, {hasher.dim}) float32 CUDA tensor")
            with self.lock:
                _native.check(_native.lib().lshx_index_query_rerank_vectors(
                    self._handle, hasher._ensure_handle(), reranker._handle, x.ctypes.data, nq, int(corpus.data_ptr()),
                    int(corpus.shape[0]), int(k), float(p), stride, ids.ctypes.data, scores.ctypes.data,
                    counts.ctypes.data, zero.ctypes.data, cands.ctypes.data, flag.ctypes.data))
        return ids, scores, counts, zero, cands, flag

    def query_one(self, signature: np.ndarray):
        """``(ids int64[c], collisions int32[c])`` of ONE query, ordered by (-collisions, id)."""
        with self.lock:
            total, _ = self.query(signature)
            if not total:
                return np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int32)
            _, counts, ids, coll = self.fetch(collisions=True)
        c = int(counts[0])
        return ids[:c], coll[:c]

    def query_host_vectors(self, hasher, vectors: np.ndarray):
        """:meth:`query` for host VECTORS in one pass over PCIe (``lshx_index_query_host_vectors``): uploaded once,
        hashed on the device, joined; the vectors stay in the handle for :meth:`rerank` with ``queries=None``.
        Returns ``(candidate slots, largest list bound, zero_flag uint8[nq])``."""
        x = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, hasher.dim)
        nq = x.shape[0]
        flag = np.zeros(nq, dtype=np.uint8)
        total, maxc = ctypes.c_int64(0), ctypes.c_int64(0)
        with self.lock:
            _native.check(_native.lib().lshx_index_query_host_vectors(
                self._handle, hasher._ensure_handle(), x.ctypes.data, nq, flag.ctypes.data, ctypes.byref(total),
                ctypes.byref(maxc)))
            self._nq, self._total = nq, int(total.value)
        return self._total, int(maxc.value), flag

    def fetch(self, *, collisions: bool = False):
        """``(offsets int64[nq+1], counts int32[nq], ids int64[slots] [, collisions int32[slots]])`` of the last
        query: list i is ``ids[offsets[i] : offsets[i] + counts[i]]``, ordered by (-collisions, id)."""
        nq, total = self._nq, self._total
        offs = np.zeros(nq + 1, dtype=np.int64)
        counts = np.zeros(nq, dtype=np.int32)
        ids = np.empty(total, dtype=np.int64)
        coll = np.empty(total, dtype=np.int32) if collisions else None
        _native.check(_native.lib().lshx_index_fetch(self._handle, offs.ctypes.data, counts.ctypes.data,
                                                     ids.ctypes.data, coll.ctypes.data if collisions else None))
        return (offs, counts, ids, coll) if collisions else (offs, counts, ids)

    def topk(self, top_k: int):
        """``(ids int64[nq, top_k] (-1 padded), counts int32[nq])``: get_top_k of the last query."""
        ids = np.empty((self._nq, int(top_k)), dtype=np.int64)
        counts = np.zeros(self._nq, dtype=np.int32)
        if self._nq:
            _native.check(_native.lib().lshx_index_topk(self._handle, int(top_k), ids.ctypes.data, counts.ctypes.data))
        return ids, counts

    def rerank(self, reranker, queries: np.ndarray, corpus, *, k: int = 0, p: float = 0.0, stride: int):
        """Cosine rerank of the last query's lists against a CUDA corpus tensor (candidate id = row).

        Returns ``(ids int64[nq, stride] (-1 padded), scores float32[nq, stride], counts int32[nq], zero int32[nq])``."""
        q = None if queries is None else np.ascontiguousarray(queries, dtype=np.float32)   # None: the handle has them
        nq = self._nq
        ids = np.empty((nq, stride), dtype=np.int64)
        scores = np.empty((nq, stride), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        zero = np.zeros(nq, dtype=np.int32)
        if nq:
            if corpus.dim() != 2 or corpus.shape[1] != reranker.dim or not corpus.is_contiguous():
                raise ValueError(f"corpus must be a contiguous (n, {reranker.dim}) float32 CUDA tensor")
            _native.check(_native.lib().lshx_index_rerank(
                self._handle, reranker._handle, None if q is None else q.ctypes.data, 0, int(corpus.data_ptr()),
                int(corpus.shape[0]),
                int(k), float(p), int(stride), ids.ctypes.data, scores.ctypes.data, counts.ctypes.data,
                zero.ctypes.data))
        return ids, scores, counts, zero

    def close(self) -> None:
        if self._handle is not None:
            _native.lib().lshx_index_destroy(self._handle)
            self._handle = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class DeviceBucketStorage:
    """The bucket store itself in HBM: the reference's storage protocol (``batch_add`` / ``get_bucket`` /
    ``remove_indices`` / ``clear`` / ``close``, reference lshrs/storage/redis.py) on a :class:`DeviceIndex`.

    ``LSHRS(storage=DeviceBucketStorage())`` needs no Redis: ``index()`` hands the packed signatures of a whole
    batch to :meth:`add_packed` (no ``(band, bytes, id)`` tuples are built -- at tens of millions of signatures
    per second they are what costs), ``query`` / ``get_top_k`` / ``get_above_p`` join on the device, and
    ``query_batch`` takes the device path by default.  The protocol methods are all there, so anything that talks
    to a ``RedisStorage`` can talk to this: ``batch_add`` accepts arbitrary operations (whole-vector runs are
    detected and take the packed path, stray ones become single entries), ``get_bucket`` returns a ``set``.
    Bucket contents equal a dict of sets fed the same operations (tests/test_device_storage*.py).

    The shape (``num_bands``, ``bytes_per_band``, device) is bound by ``LSHRS`` when not given here.  Unlike a
    Redis server the store lives and dies with the process: :meth:`save` / :meth:`load` write and read an ``.npz``.
    """

    def __init__(self, num_bands: int | None = None, bytes_per_band: int | None = None, device: int | None = None,
                 prefix: str = "lsh") -> None:
        self.prefix = prefix
        self._device = device
        self.index: DeviceIndex | None = None
        if num_bands is not None and bytes_per_band is not None:
            self.bind(num_bands, bytes_per_band, device)

    # ---- shape
    def bind(self, num_bands: int, bytes_per_band: int, device: int | None = None) -> "DeviceBucketStorage":
        """Fix the index shape (idempotent; a different shape on a bound store is an error)."""
        if self.index is not None:
            if (self.index.num_bands, self.index.bytes_per_band) != (int(num_bands), int(bytes_per_band)):
                raise ValueError(
                    f"store is bound to {self.index.num_bands} bands x {self.index.bytes_per_band} bytes, "
                    f"cannot serve {num_bands} x {bytes_per_band}")
            return self
        self.index = DeviceIndex(num_bands, bytes_per_band, device=self._device if device is None else device)
        return self

    def _ix(self) -> DeviceIndex:
        if self.index is None:
            raise RuntimeError("DeviceBucketStorage is not bound to an index shape yet (pass it to LSHRS(storage=...) "
                               "or give num_bands / bytes_per_band)")
        return self.index

    def bucket_key(self, band_id: int, hash_val: bytes) -> str:
        from lshrs_b200.storage.memory import bucket_key

        return bucket_key(self.prefix, band_id, hash_val)

    # ---- the packed path LSHRS.index() takes
    def add_packed(self, signatures: np.ndarray, ids) -> None:
        """``n`` whole vectors: uint8 ``(n, num_bands, bytes_per_band)`` + ``n`` ids -- ``batch_add`` of their
        ``n * num_bands`` operations without building them."""
        self._ix().add(signatures, ids)

    # ---- reference protocol
    def batch_add(self, operations) -> None:
        ix = self._ix()
        ops = operations if isinstance(operations, list) else list(operations)
        m = len(ops)
        if not m:
            return
        nb, bpb = ix.num_bands, ix.bytes_per_band
        bands = np.fromiter((op[0] for op in ops), dtype=np.int64, count=m)
        ids = np.fromiter((op[2] for op in ops), dtype=np.int64, count=m)
        if bands.min() < 0 or bands.max() >= nb:
            raise ValueError(f"band id outside [0, {nb})")
        blob = b"".join(op[1] for op in ops)
        if len(blob) != m * bpb:
            raise ValueError(f"band keys must be {bpb} bytes each")
        keys = np.frombuffer(blob, dtype=np.uint8).reshape(m, bpb)
        # whole-vector runs (what LSHRS enqueues: bands 0 .. nb-1 of one id, in order) take the packed path
        if m % nb == 0 and np.array_equal(bands.reshape(-1, nb), np.broadcast_to(np.arange(nb), (m // nb, nb))) \
                and (ids.reshape(-1, nb) == ids.reshape(-1, nb)[:, :1]).all():
            ix.add(keys.reshape(m // nb, nb, bpb), ids.reshape(-1, nb)[:, 0].copy())
            return
        sig = np.zeros((m, nb, bpb), dtype=np.uint8)
        per_band = np.full((m, nb), -1, dtype=np.int64)
        rows = np.arange(m)
        sig[rows, bands] = keys
        per_band[rows, bands] = ids
        if (ids < 0).any():
            raise ValueError("vector ids must lie in [0, 2^56) for the device index")
        ix.add_entries(sig, per_band)

    def add_to_bucket(self, band_id: int, hash_val: bytes, index: int) -> None:
        self.batch_add([(int(band_id), bytes(hash_val), int(index))])

    def get_bucket(self, band_id: int, hash_val: bytes) -> set[int]:
        return self.get_buckets([(band_id, hash_val)])[0]

    def get_buckets(self, keys) -> list[set[int]]:
        ix = self._ix()
        keys = keys if isinstance(keys, list) else list(keys)
        m = len(keys)
        if not m:
            return []
        blob = b"".join(bytes(k) for _, k in keys)
        if len(blob) != m * ix.bytes_per_band:   # a key of another length names no bucket of this index
            out = []
            for b, k in keys:
                out.extend(self.get_buckets([(b, k)]) if len(bytes(k)) == ix.bytes_per_band else [set()])
            return out
        offs, ids = ix.get_buckets(np.fromiter((b for b, _ in keys), dtype=np.int32, count=m),
                                   np.frombuffer(blob, dtype=np.uint8))
        flat, o = ids.tolist(), offs.tolist()
        return [set(flat[o[t]:o[t + 1]]) for t in range(m)]

    def remove_indices(self, indices) -> None:
        self._ix().remove([int(i) for i in indices])

    def clear(self) -> None:
        if self.index is not None:
            self.index.clear()

    def close(self) -> None:
        pass   # like InMemoryStorage: the data must outlive LSHRS.close(); the handle goes with the object

    def __len__(self) -> int:
        return len(self.index) if self.index is not None else 0

    # ---- persistence (the store dies with the process; Redis has RDB/AOF for this)
    def _state(self) -> dict:
        ix = self._ix()
        keys, ids = ix.export()
        return {"num_bands": ix.num_bands, "bytes_per_band": ix.bytes_per_band, "prefix": self.prefix, "keys": keys, "ids": ids}

    def _restore(self, state: dict, device: int | None = None) -> None:
        self.prefix = state.get("prefix", "lsh")
        self.bind(int(state["num_bands"]), int(state["bytes_per_band"]), device)
        keys, ids = np.asarray(state["keys"], dtype=np.uint8), np.asarray(state["ids"], dtype=np.int64)
        nb, n = ids.shape
        if n:
            # entry e of every band travels as one row; the per-band ids need not belong to one vector
            self.index.add_entries(np.ascontiguousarray(keys.transpose(1, 0, 2)), np.ascontiguousarray(ids.T))

    def save(self, path) -> None:
        st = self._state()
        np.savez_compressed(path, num_bands=st["num_bands"], bytes_per_band=st["bytes_per_band"],
                            prefix=np.array(st["prefix"]), keys=st["keys"], ids=st["ids"])

    @classmethod
    def load(cls, path, device: int | None = None) -> "DeviceBucketStorage":
        with np.load(path) as data:
            st = {"num_bands": int(data["num_bands"]), "bytes_per_band": int(data["bytes_per_band"]),
                  "prefix": str(data["prefix"]), "keys": data["keys"], "ids": data["ids"]}
        out = cls(device=device)
        out._restore(st, device)
        return out

    def __getstate__(self) -> dict:
        return {"device": self._device, "state": self._state() if self.index is not None else None, "prefix": self.prefix}

    def __setstate__(self, state: dict) -> None:
        self.prefix, self._device, self.index = state["prefix"], state["device"], None
        if state["state"] is not None:
            self._restore(state["state"], self._device)
