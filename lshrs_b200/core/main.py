"""``LSHRS`` -- the reference orchestrator's hot-path call sites on top of the GPU kernels.

Mirrors the parts of reference lshrs/core/main.py that call the hasher and the
reranker: ``ingest`` (main.py:386-411), ``index`` (442-518), ``flush``
(413-440), ``query`` (524-658), ``get_top_k`` (660-693), ``get_above_p``
(695-738), ``delete`` / ``clear`` / ``stats`` and the buffer helpers
(1050-1143) -- same signatures, validation order, error types and messages,
same ``(band_id, bytes, index)`` operations with the same flush boundaries.
Bucket storage is NOT re-implemented: pass the reference's ``RedisStorage`` (or
anything with its five methods) as ``storage=``.  ``save_to_disk`` /
``load_from_disk`` / pickling keep the reference's formats, ``create_signatures``
takes the Arrow-buffer Parquet feed or the binary-COPY PostgreSQL feed
(``lshrs_b200/io``; INTEGRATION.md shows the two-import patch that puts the
reference's own ``LSHRS`` on these kernels instead).

What is new is batching: ``index()`` hashes the whole batch in ONE kernel call
with the zero-vector test fused in, and ``query_batch`` hashes all queries at
once and reranks all of them in one launch.
"""

from __future__ import annotations

import logging
import math
from collections.abc import Callable, Sequence
from threading import Lock
from typing import Any, Optional, Union

import numpy as np

from lshrs_b200._config.config import HashSignatures
from lshrs_b200.hash.lsh import LSHHasher
from lshrs_b200.storage.memory import BucketOperation, BucketStorage, InMemoryStorage
from lshrs_b200.utils.br import get_optimal_config
from lshrs_b200.utils.similarity import _get_reranker

logger = logging.getLogger(__name__)

VectorFetchFn = Callable[[Sequence[int]], np.ndarray]
CandidateScores = list[tuple[int, float]]

_ZERO_VECTOR_MSG = "Cannot index zero vector - norm undefined. Check embeddings for corruption."

try:  # Redis bucket storage is the reference's, unchanged (out of scope here).  The name
    # lives at module level because the reference's tests patch it there
    # (reference tests/test_redis_pooling.py:34: patch("lshrs.core.main.RedisStorage")).
    from lshrs.storage.redis import RedisStorage  # type: ignore
except ImportError:  # the reference package is not installed next to us: storage= must be given
    RedisStorage = None


def _auto_config(num_perm: int, threshold: float) -> tuple[int, int]:
    """``(num_bands, rows_per_band)`` when the caller gives only ``num_perm`` (reference main.py:253-255)."""
    b, r = get_optimal_config(num_perm, threshold)
    return int(b), int(r)


class LSHRS:
    """LSH index front-end whose hashing and reranking run on a B200."""

    def __init__(
        self,
        *,
        dim: int,
        num_perm: int = 128,
        num_bands: Optional[int] = None,
        rows_per_band: Optional[int] = None,
        similarity_threshold: float = 0.5,
        buffer_size: int = 10_000,
        vector_fetch_fn: Optional[VectorFetchFn] = None,
        storage: Optional[BucketStorage] = None,
        redis_host: str = "localhost",
        redis_port: int = 6379,
        redis_db: int = 0,
        redis_password: Optional[str] = None,
        redis_prefix: str = "lsh",
        redis_max_connections: int = 50,
        decode_responses: bool = False,
        seed: int = 42,
        device: Optional[int] = None,
        device_index: bool = False,
        corpus=None,
    ) -> None:
        if dim <= 0:
            raise ValueError("Vector dimensionality must be greater than zero")
        if num_perm <= 0:
            raise ValueError("num_perm must be greater than zero")
        if buffer_size <= 0:
            raise ValueError("buffer_size must be greater than zero")
        if num_bands is None or rows_per_band is None:
            num_bands, rows_per_band = _auto_config(num_perm, similarity_threshold)
        if num_bands * rows_per_band != num_perm:
            raise ValueError(
                f"num_bands * rows_per_band must equal num_perm (received {num_bands} * {rows_per_band} != {num_perm})"
            )
        self._dim = dim
        self._buffer_size = buffer_size
        self._vector_fetch_fn = vector_fetch_fn
        # optional: the indexed vectors resident in HBM (CUDA float32 tensor (N, dim), candidate id = row) -- the
        # device-side stand-in for vector_fetch_fn: reranks gather from it instead of fetching to the host
        self._corpus = None
        if corpus is not None:
            self.set_corpus(corpus)
        self._hasher = LSHHasher(num_bands=num_bands, rows_per_band=rows_per_band, dim=dim, seed=seed, device=device)
        if storage is None:
            storage = self._make_redis_storage(
                host=redis_host, port=redis_port, db=redis_db, password=redis_password,
                decode_responses=decode_responses, prefix=redis_prefix, max_connections=redis_max_connections,
            )
        self._storage = storage
        self._buffer: list[BucketOperation] = []
        self._buffer_lock = Lock()
        # optional mirror of the bucket store in HBM (lshrs_b200/storage/device.py): what reaches the store in a
        # flush is added to it right after, so batched queries can generate their candidates on the GPU
        self._dindex = None
        self._mirror_pending: list[tuple[np.ndarray, np.ndarray]] = []   # (signatures, ids) of buffered operations
        # storage=DeviceBucketStorage(): the store itself lives in HBM -- no mirror to keep, index() hands it packed
        # signatures, queries join on the device
        self._store_on_device = False
        if hasattr(storage, "add_packed") and hasattr(storage, "bind"):
            storage.bind(num_bands, self._hasher.bytes_per_band, self._hasher.device)
            self._dindex = storage.index
            self._store_on_device = True
        elif device_index:
            from lshrs_b200.storage.device import DeviceIndex

            self._dindex = DeviceIndex(num_bands, self._hasher.bytes_per_band, device=self._hasher.device)
        self._config: dict[str, Any] = {
            "dim": dim, "num_perm": num_perm, "num_bands": num_bands, "rows_per_band": rows_per_band,
            "similarity_threshold": similarity_threshold, "buffer_size": buffer_size, "seed": seed,
        }
        self._redis_config: dict[str, Any] = {
            "host": redis_host, "port": redis_port, "db": redis_db, "password": redis_password,
            "prefix": redis_prefix, "decode_responses": decode_responses, "max_connections": redis_max_connections,
        }

    @staticmethod
    def _make_redis_storage(**kwargs: Any) -> BucketStorage:
        cls = RedisStorage      # module global: patched by the reference's tests, None when not importable at load
        if cls is None:
            try:
                from lshrs.storage.redis import RedisStorage as cls  # type: ignore
            except ImportError as exc:
                raise RuntimeError(
                    "no storage given: pass storage=<RedisStorage or compatible> (Redis bucket storage stays in the "
                    "reference package lshrs.storage.redis, which is not importable here)"
                ) from exc
        return cls(**kwargs)

    def set_corpus(self, corpus) -> None:
        """Rerank against ``corpus`` (contiguous CUDA float32 tensor ``(N, dim)``, candidate id = row) instead of
        calling ``vector_fetch_fn``; ``None`` goes back to the callback."""
        if corpus is not None:
            if not getattr(corpus, "is_cuda", False) or corpus.dim() != 2 or corpus.shape[1] != self._dim \
                    or not corpus.is_contiguous() or str(corpus.dtype) != "torch.float32":
                raise ValueError(f"corpus must be a contiguous float32 CUDA tensor of shape (n, {self._dim})")
        self._corpus = corpus

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        self.flush()
        self._storage.close()

    def __enter__(self) -> "LSHRS":
        return self

    def __exit__(self, exc_type, exc_value, traceback) -> None:
        self.close()

    def __repr__(self) -> str:  # pragma: no cover
        c = self._config
        return (f"LSHRS(dim={self._dim}, num_perm={c['num_perm']}, num_bands={c['num_bands']}, "
                f"rows_per_band={c['rows_per_band']}, redis_prefix='{self._redis_config['prefix']}')")

    # ------------------------------------------------------------------ ingestion
    def create_signatures(self, *, format: str = "postgres", **loader_kwargs: Any) -> None:
        """Stream ``(indices, vectors)`` batches from a loader into :meth:`index` (reference main.py:315-384)."""
        loader = self._resolve_loader(format)
        from lshrs_b200.io.parquet import prefetched

        # decode the next row group / read the next COPY block while this batch is hashed
        batches = prefetched(loader(**loader_kwargs))
        for indices, vectors in batches:
            self.index(indices, vectors)

    @staticmethod
    def _resolve_loader(format: str):
        normalized = format.lower()
        if normalized in {"parquet", "pq"}:
            from lshrs_b200.io.parquet import iter_parquet_vectors

            return iter_parquet_vectors
        if normalized in {"postgres", "pg"}:
            from lshrs_b200.io.postgres import iter_postgres_vectors

            return iter_postgres_vectors
        raise ValueError(f"Unsupported signature creation format '{format}'")

    def ingest(self, index: int, vector: np.ndarray) -> None:
        """Hash one vector and buffer its ``num_bands`` bucket operations."""
        if index < 0:
            raise ValueError("index must be non-negative")
        vec = self._prepare_vector(vector)
        signatures = self._hasher.hash_vector(vec)
        self._enqueue_operations(index, signatures)
        if self._dindex is not None and not self._store_on_device:
            with self._buffer_lock:   # (bytes, int): turned into arrays when the flush hands them to the mirror
                self._mirror_pending.append((b"".join(signatures), int(index)))
        self._flush_buffer_if_needed()

    def flush(self) -> None:
        """Send buffered operations to storage; on failure put them back at the front and re-raise."""
        with self._buffer_lock:
            if not self._buffer:
                return
            pending = list(self._buffer)
            self._buffer.clear()
            mirror, self._mirror_pending = self._mirror_pending, []
        try:
            self._storage.batch_add(pending)
        except Exception as exc:
            logger.error(f"Failed to flush buffer to Redis: {exc}")
            with self._buffer_lock:
                self._buffer[0:0] = pending
                self._mirror_pending[0:0] = mirror
            raise
        # the store has them: now the device mirror may (single ingests are coalesced into one add)
        singles: list[tuple[bytes, int]] = []
        for sig, ids in [*mirror, (None, None)]:
            if isinstance(sig, bytes):
                singles.append((sig, ids))
                continue
            if singles:
                self._dindex.add(np.frombuffer(b"".join(s for s, _ in singles), dtype=np.uint8),
                                 np.fromiter((i for _, i in singles), dtype=np.int64, count=len(singles)))
                singles = []
            if sig is not None:
                self._dindex.add(sig, ids)

    def index(self, indices: Sequence[int], vectors: Optional[np.ndarray] = None) -> None:
        """Ingest a batch: ONE kernel call hashes every row and flags zero vectors.

        Observable behaviour equals the reference's per-row ``ingest`` loop: the
        same operations in the same order, a flush whenever the buffer reaches
        ``buffer_size`` after a whole vector, rows before the first invalid one
        are ingested before the error is raised, and a final flush.
        """
        if not len(indices):
            return
        if vectors is None:
            vectors = self._require_vector_fetch_fn()(indices)
        if self._store_on_device and getattr(vectors, "is_cuda", False):
            return self._index_resident(indices, vectors)
        if isinstance(vectors, np.ndarray) and vectors.dtype in self._hasher._TYPED and vectors.ndim == 2:
            arr = vectors   # float16 / int8 / uint8 batches are cast on the device (exact), not here
        else:
            arr = np.asarray(vectors, dtype=np.float32)
        if arr.ndim != 2 or arr.shape[1] != self._dim:
            raise ValueError(f"Vectors must have shape (n, {self._dim}); received {arr.shape}")
        if arr.shape[0] != len(indices):
            raise ValueError(
                "Number of vectors does not match number of indices "
                f"(received {arr.shape[0]} vectors for {len(indices)} indices)"
            )
        packed, zero_flag = self._hasher.hash_batch_packed(arr, return_zero_flag=True)
        nb, bpb = self._hasher.num_bands, self._hasher.bytes_per_band
        if self._store_on_device:
            return self._index_packed(indices, packed, zero_flag)
        ids = self._ids_array(indices)
        if ids is None:
            return self._index_rows(indices, packed, zero_flag)      # odd index types: row by row, like the reference
        # The reference's loop (main.py:504-518) flushes after the row at which the buffer reaches buffer_size and
        # once at the end.  Same operations, same batches -- but built one flush segment at a time (one list
        # comprehension, one lock round trip per segment instead of per row: the Python loop was 3x the store's time).
        n = len(indices)
        negative = ids < 0
        invalid = negative | (zero_flag != 0)
        stop = int(np.argmax(invalid)) if invalid.any() else n
        # every band key as an exact-length bytes object, created in C (void-dtype tolist)
        keys = np.ascontiguousarray(packed[:stop]).reshape(stop, nb * bpb).view(f"V{bpb}").tolist()
        id_list = ids[:stop].tolist()
        buffer, lock, limit = self._buffer, self._buffer_lock, self._buffer_size
        with lock:
            pre = len(buffer)
        per_flush = max(1, -(-limit // nb))              # rows between two flushes of the loop
        first = max(1, -(-(limit - pre) // nb))          # rows until its first flush
        mirrored = 0                                     # rows [0, mirrored) are already queued for the device mirror

        def queue_mirror(upto: int) -> None:
            nonlocal mirrored
            if self._dindex is not None and upto > mirrored:
                with lock:
                    self._mirror_pending.append((packed[mirrored:upto], ids[mirrored:upto].copy()))
                mirrored = upto

        row, seg_end = 0, min(stop, first)
        while row < stop:
            ops = [(b, key, idx) for krow, idx in zip(keys[row:seg_end], id_list[row:seg_end])
                   for b, key in enumerate(krow)]
            with lock:
                buffer.extend(ops)
                due = len(buffer) >= limit
            if due:
                queue_mirror(seg_end)
                self.flush()
            row, seg_end = seg_end, min(stop, seg_end + per_flush)
        queue_mirror(stop)
        if stop < n:        # the rows before the invalid one stay buffered, as in the reference's loop
            raise ValueError("index must be non-negative" if negative[stop] else _ZERO_VECTOR_MSG)
        self.flush()

    @staticmethod
    def _ids_array(indices) -> Optional[np.ndarray]:
        """``indices`` as an int64 array, or None when they are not plain integers (the caller then walks them)."""
        try:
            ids = np.asarray(indices)
            if ids.dtype.kind not in "iu" or ids.ndim != 1:
                return None
            return ids.astype(np.int64, copy=False)
        except (TypeError, ValueError, OverflowError):
            return None

    def _index_rows(self, indices: Sequence[int], packed: np.ndarray, zero_flag: np.ndarray) -> None:
        """``index()`` row by row (index sequences that are not plain integers: ``int()`` is applied to each in
        turn, so a bad one is met exactly where the reference's loop meets it)."""
        nb, bpb = self._hasher.num_bands, self._hasher.bytes_per_band
        keys = np.ascontiguousarray(packed).reshape(len(indices), nb * bpb).view(f"V{bpb}").tolist()
        flags = zero_flag.tolist()
        band_ids = range(nb)
        buffer, lock, limit = self._buffer, self._buffer_lock, self._buffer_size
        mirrored = 0                # rows [0, mirrored) are already queued for the device mirror
        store_packed = self._store_on_device

        def queue_mirror(upto: int) -> None:
            nonlocal mirrored
            if self._dindex is not None and not store_packed and upto > mirrored:
                ids_arr = np.fromiter((int(i) for i in indices[mirrored:upto]), dtype=np.int64, count=upto - mirrored)
                with lock:
                    self._mirror_pending.append((packed[mirrored:upto], ids_arr))
                mirrored = upto

        row = 0
        try:
            for row, idx in enumerate(indices):
                idx = int(idx)
                if idx < 0:
                    raise ValueError("index must be non-negative")
                if flags[row]:
                    raise ValueError(_ZERO_VECTOR_MSG)
                ops = [(b, key, idx) for b, key in zip(band_ids, keys[row])]
                with lock:
                    buffer.extend(ops)
                    due = len(buffer) >= limit
                if due:
                    queue_mirror(row + 1)
                    self.flush()
        except ValueError:
            queue_mirror(row)       # the rows before the invalid one stay buffered, as in the reference's loop
            raise
        queue_mirror(len(indices))
        self.flush()

    def _index_resident(self, indices, vectors) -> None:
        """``index()`` of vectors that are already in HBM (a CUDA float32 tensor) into the store in HBM: hash and
        append without anything crossing PCIe but one validity flag.  ``indices``: a sequence or an int64 tensor."""
        import torch

        if vectors.dim() != 2 or vectors.shape[1] != self._dim:
            raise ValueError(f"Vectors must have shape (n, {self._dim}); received {tuple(vectors.shape)}")
        n = int(vectors.shape[0])
        if n != len(indices):
            raise ValueError(
                "Number of vectors does not match number of indices "
                f"(received {n} vectors for {len(indices)} indices)"
            )
        dev = vectors.device
        if isinstance(indices, torch.Tensor):
            ids_dev = indices.to(device=dev, dtype=torch.int64).contiguous()
        else:
            ids_dev = torch.from_numpy(np.ascontiguousarray(np.asarray(indices, dtype=np.int64))).to(dev)
        flag = torch.zeros(n, dtype=torch.uint8, device=dev)
        packed = self._hasher.hash_device(vectors.to(torch.float32), zero_flag=flag)
        if bool(((ids_dev < 0) | (flag != 0)).any()):
            # an invalid row: the per-row semantics (what is flushed, what stays buffered) live on the host path
            return self._index_packed(ids_dev.cpu().numpy(), packed.cpu().numpy(), flag.cpu().numpy())
        self.flush()
        self._dindex.add_device(packed, ids_dev, torch.cuda.current_stream(dev).cuda_stream)

    def _index_packed(self, indices: Sequence[int], packed: np.ndarray, zero_flag: np.ndarray) -> None:
        """``index()`` on a store that takes packed signatures (``DeviceBucketStorage.add_packed``).

        No ``(band, bytes, id)`` tuples are built.  What the store holds afterwards, what stays buffered when a row
        is invalid and the order of operations equal the reference's per-row loop (main.py:504-518): that loop
        flushes after the row at which the buffer reaches ``buffer_size`` and once at the end, so on an invalid
        row ``stop`` the rows since the last such flush point stay in the buffer (as tuples, the rare path)."""
        n, nb, bpb = len(indices), self._hasher.num_bands, self._hasher.bytes_per_band
        ids = self._ids_array(indices)
        if ids is None:
            return self._index_rows(indices, packed, zero_flag)      # through batch_add, row by row
        negative = ids < 0
        invalid = negative | (zero_flag != 0)
        stop = int(np.argmax(invalid)) if invalid.any() else n
        with self._buffer_lock:
            pre = len(self._buffer)
        per_flush = max(1, -(-self._buffer_size // nb))            # rows between two flushes of the loop
        first = max(1, -(-(self._buffer_size - pre) // nb))        # rows until its first flush
        if stop == n:
            sent = n
        else:
            sent = 0 if stop < first else first + (stop - first) // per_flush * per_flush

        def as_ops(lo: int, hi: int) -> list[BucketOperation]:
            keys = np.ascontiguousarray(packed[lo:hi]).reshape(hi - lo, nb * bpb).view(f"V{bpb}").tolist()
            return [(b, key, idx) for row, idx in zip(keys, ids[lo:hi].tolist()) for b, key in enumerate(row)]

        if sent:
            self.flush()                    # operations buffered by earlier ingest() calls go first, in order
            try:
                self._storage.add_packed(packed[:sent], ids[:sent])
            except Exception as exc:
                logger.error(f"Failed to flush buffer to Redis: {exc}")
                with self._buffer_lock:     # like flush(): nothing is lost, the operations are buffered again
                    self._buffer[0:0] = as_ops(0, stop)
                raise
        if stop < n:
            if stop > sent:
                with self._buffer_lock:
                    self._buffer.extend(as_ops(sent, stop))
            raise ValueError("index must be non-negative" if negative[stop] else _ZERO_VECTOR_MSG)

    # ------------------------------------------------------------------ querying
    def query(self, vector: np.ndarray, *, top_k: Optional[int] = 10,
              top_p: Optional[float] = None) -> Union[list[int], CandidateScores]:
        """Candidates by band collisions; with ``top_p`` reranked by cosine on the GPU."""
        q = self._prepare_vector(vector)
        if self._store_on_device and top_p is None and top_k is not None and top_k > 0:
            return self._device_candidates(q, top_k)[0].tolist()     # only the slice crosses PCIe
        if (self._store_on_device and self._corpus is not None and top_p is not None and 0 < top_p <= 1
                and (top_k is None or top_k > 0)):
            ranked = self._device_rerank(q, top_k, float(top_p))
            if ranked is not None:
                return ranked
        counts = self._candidate_counts(q)
        if not counts:
            return []
        ordered = sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))
        if top_p is None:
            if top_k is None:
                top_k = len(ordered)
            if top_k <= 0:
                raise ValueError("top_k must be greater than zero when provided")
            return [idx for idx, _ in ordered[:top_k]]
        if not 0 < top_p <= 1:
            raise ValueError("top_p must be within the range (0, 1]")
        candidate_indices = [idx for idx, _ in ordered]
        if top_k is not None and top_k <= 0:
            raise ValueError("top_k must be greater than zero when provided")
        n = len(candidate_indices)
        rer = _get_reranker(self._dim, self._hasher.device)
        if self._corpus is not None:
            # the candidates' vectors are gathered from HBM by id; nothing but ids and the kept scores move
            ids_arr = np.asarray(candidate_indices, dtype=np.int64)
            if ids_arr.max() >= self._corpus.shape[0]:
                raise ValueError(f"candidate id {int(ids_arr.max())} is not a row of the corpus "
                                 f"({self._corpus.shape[0]} rows)")
            pos, score, count, zero = rer.topk(q, self._corpus, np.array([0, n], dtype=np.int64), ids_arr,
                                               k=int(top_k) if top_k is not None else 0, p=float(top_p),
                                               vectors_on_device=True)
        else:
            fetched = self._require_vector_fetch_fn()(candidate_indices)
            arr = np.ascontiguousarray(np.asarray(fetched, dtype=np.float32))
            if arr.ndim != 2 or arr.shape[1] != self._dim:
                raise ValueError(f"Fetched vectors must have shape (n, {self._dim}); received {arr.shape}")
            if arr.shape[0] != len(candidate_indices):
                raise ValueError(
                    "vector_fetch_fn returned mismatched batch size "
                    f"(expected {len(candidate_indices)}, received {arr.shape[0]})"
                )
            # The reference sorts ALL n candidates (top_k_cosine(k=n)) and slices
            # max(1, ceil(n * top_p)) [then min(., top_k)]; the kernel applies the same
            # cut and only the kept rows come back over PCIe.
            pos, score, count, zero = rer.topk(q, arr, np.array([0, n], dtype=np.int64),
                                               k=int(top_k) if top_k is not None else 0, p=float(top_p))
        if zero.any():
            raise ValueError("Cannot normalize zero vector")
        return [(candidate_indices[int(pos[0, i])], float(score[0, i])) for i in range(int(count[0]))]

    def get_top_k(self, vector: np.ndarray, topk: int = 10) -> list[int]:
        return list(self.query(vector, top_k=topk, top_p=None))  # type: ignore[arg-type]

    def get_above_p(self, vector: np.ndarray, p: float = 0.95) -> CandidateScores:
        return list(self.query(vector, top_k=None, top_p=p))  # type: ignore[arg-type]

    def query_batch(self, vectors: np.ndarray, *, top_k: Optional[int] = 10, top_p: Optional[float] = None,
                    corpus=None, device_index: Optional[bool] = None, as_arrays: bool = False):
        """``query`` for many vectors: one hash launch, one rerank launch.

        Returns one result list per row, identical to calling :meth:`query` row
        by row.  With ``top_p``, candidate vectors come from ``corpus`` when given
        (a CUDA torch tensor ``(N, dim)`` resident in HBM, candidate id = row --
        the device-side stand-in for ``vector_fetch_fn``), else from ``vector_fetch_fn``.

        ``device_index=True`` (needs ``LSHRS(device_index=True)`` or ``storage=DeviceBucketStorage()``, with which
        it is the default): the candidates come from the band index in HBM instead of ``num_bands`` bucket reads
        per query -- same lists, same order, same results (``lshx_index_query``); with ``corpus`` the lists never
        leave the GPU before the rerank.
        ``as_arrays=True`` returns numpy arrays instead of Python lists (-1 padded ids ``(nq, k)``, then scores
        for ``top_p``, then counts): at several hundred thousand queries per second the lists are what costs.
        """
        arr = np.asarray(vectors, dtype=np.float32)
        if arr.ndim != 2 or arr.shape[1] != self._dim:
            raise ValueError(f"Vectors must have shape (n, {self._dim}); received {arr.shape}")
        nq = arr.shape[0]
        if corpus is None:
            corpus = self._corpus
        if device_index is None:
            device_index = self._store_on_device
        if as_arrays and not device_index:
            raise ValueError("as_arrays=True is only available with device_index=True")
        if device_index and self._dindex is None:
            raise RuntimeError("device_index=True needs an index built with LSHRS(..., device_index=True)")
        if nq == 0:
            return []
        if top_p is None and top_k is not None and top_k <= 0:
            raise ValueError("top_k must be greater than zero when provided")
        if top_p is not None and not 0 < top_p <= 1:
            raise ValueError("top_p must be within the range (0, 1]")
        if top_p is not None and top_k is not None and top_k <= 0:
            raise ValueError("top_k must be greater than zero when provided")
        nb, bpb = self._hasher.num_bands, self._hasher.bytes_per_band
        # one pass over PCIe (vectors up once, hashed and joined on the device) when the batch takes the batch
        # kernel anyway; 32 rows and fewer keep the FP32 latency kernel every other path uses for them
        one_pass = device_index and nq > 32 and self._hasher._kernel == 0
        if one_pass:
            packed = None
        else:
            packed, zero_flag = self._hasher.hash_batch_packed(arr, return_zero_flag=True)
            if zero_flag.any():
                raise ValueError(_ZERO_VECTOR_MSG)
        if device_index:
            with self._dindex.lock:     # the result of query() lives in the handle until it is consumed
                done, ordered_all = self._query_batch_on_device(arr, packed, top_k, top_p, corpus, as_arrays)
            if done:
                return ordered_all
        else:
            # every (query, band) bucket in ONE storage round trip when the backend allows it
            keys = np.ascontiguousarray(packed).reshape(nq, nb * bpb).view(f"V{bpb}").tolist()
            buckets = self._fetch_buckets([(b, key) for row in keys for b, key in enumerate(row)])
            ordered_all = []
            for row in range(nq):
                counts: dict[int, int] = {}
                for members in buckets[row * nb:(row + 1) * nb]:
                    for cand in members:
                        counts[cand] = counts.get(cand, 0) + 1
                ordered = sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))
                ordered_all.append([idx for idx, _ in ordered])
        if top_p is None:
            return [ids if top_k is None else ids[:top_k] for ids in ordered_all]
        lens = np.array([len(x) for x in ordered_all], dtype=np.int64)
        offsets = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        if offsets[-1] == 0:
            return [[] for _ in range(nq)]
        flat_ids = np.fromiter((i for ids in ordered_all for i in ids), dtype=np.int64, count=int(offsets[-1]))
        rer = _get_reranker(self._dim, self._hasher.device)
        if corpus is not None:
            pos, score, count, zero = rer.topk(arr, corpus, offsets, flat_ids, k=int(top_k or 0), p=float(top_p),
                                               vectors_on_device=True)
        else:
            fetched = np.ascontiguousarray(np.asarray(self._require_vector_fetch_fn()(flat_ids.tolist()), dtype=np.float32))
            if fetched.ndim != 2 or fetched.shape[1] != self._dim:
                raise ValueError(f"Fetched vectors must have shape (n, {self._dim}); received {fetched.shape}")
            if fetched.shape[0] != flat_ids.shape[0]:
                raise ValueError(
                    "vector_fetch_fn returned mismatched batch size "
                    f"(expected {flat_ids.shape[0]}, received {fetched.shape[0]})"
                )
            pos, score, count, zero = rer.topk(arr, fetched, offsets, None, k=int(top_k or 0), p=float(top_p))
        nonempty = lens > 0
        if zero[nonempty].any():
            raise ValueError("Cannot normalize zero vector")
        out: list = []
        for row in range(nq):
            ids = ordered_all[row]
            out.append([(ids[int(pos[row, i])], float(score[row, i])) for i in range(int(count[row]))] if ids else [])
        return out

    def _query_batch_on_device(self, arr, packed, top_k, top_p, corpus, as_arrays):
        """The device-index branch of :meth:`query_batch`; ``(True, result)`` or ``(False, candidate lists)``."""
        nq = arr.shape[0]
        queries = arr
        if packed is None:
            _, maxc, zero_flag = self._dindex.query_host_vectors(self._hasher, arr)
            if zero_flag.any():
                raise ValueError(_ZERO_VECTOR_MSG)
            queries = None              # the handle holds the uploaded vectors for the rerank
        else:
            _, maxc = self._dindex.query(packed)
        if top_p is None and top_k is not None:
            ids, counts = self._dindex.topk(top_k)
            if as_arrays:
                return True, (ids, counts)
            rows, cl = ids.tolist(), counts.tolist()
            return True, [rows[i][:cl[i]] for i in range(nq)]
        if top_p is not None and corpus is not None:
            stride = max(1, math.ceil(maxc * top_p))
            if top_k is not None:
                stride = min(stride, int(top_k))
            rer = _get_reranker(self._dim, self._hasher.device)
            ids, scores, counts, zero = self._dindex.rerank(rer, queries, corpus, k=int(top_k or 0), p=float(top_p),
                                                            stride=stride)
            if zero[counts > 0].any():
                raise ValueError("Cannot normalize zero vector")
            if as_arrays:
                return True, (ids, scores, counts)
            rows, sc, cl = ids.tolist(), scores.tolist(), counts.tolist()
            return True, [list(zip(rows[i][:cl[i]], sc[i][:cl[i]])) for i in range(nq)]
        offs, counts, flat = self._dindex.fetch()
        if as_arrays and top_p is None:
            return True, (offs, counts, flat)
        if as_arrays:
            raise ValueError("as_arrays=True with top_p needs corpus= (the rerank on the device)")
        return False, [flat[o:o + c].tolist() for o, c in zip(offs[:-1].tolist(), counts.tolist())]

    # ------------------------------------------------------------------ maintenance
    def delete(self, indices: Union[int, Sequence[int]]) -> None:
        to_remove = [indices] if isinstance(indices, int) else [int(i) for i in indices]
        self._storage.remove_indices(to_remove)
        if self._dindex is not None and not self._store_on_device:
            self._dindex.remove(to_remove)

    def clear(self) -> None:
        self.flush()
        self._storage.clear()
        if self._dindex is not None and not self._store_on_device:
            self._dindex.clear()

    def stats(self) -> dict[str, Any]:
        c = self._config
        return {
            "dimension": self._dim, "num_perm": c["num_perm"], "num_bands": c["num_bands"],
            "rows_per_band": c["rows_per_band"], "buffer_size": self._buffer_size,
            "similarity_threshold": c["similarity_threshold"], "redis_prefix": self._redis_config["prefix"],
        }

    # ------------------------------------------------------------------ persistence
    # Same on-disk / pickle formats as the reference (main.py:846-1044): metadata.json + projections.npz
    # (arr_0 .. arr_{b-1}); bucket data lives in the storage backend and is not saved.
    def save_to_disk(self, path) -> None:
        import json
        from pathlib import Path

        self.flush()
        out = Path(path)
        out.mkdir(parents=True, exist_ok=True)
        redis_cfg = dict(self._redis_config)
        if "password" in redis_cfg:
            redis_cfg["password"] = "<REDACTED>"
        meta = {"version": "0.1.1a4", "config": self._config, "redis_config": redis_cfg}
        (out / "metadata.json").write_text(json.dumps(meta, indent=2))
        np.savez_compressed(out / "projections.npz", *self._hasher.projections)

    @classmethod
    def load_from_disk(cls, path, *, redis_config: Optional[dict[str, Any]] = None,
                       vector_fetch_fn: Optional[VectorFetchFn] = None,
                       storage: Optional[BucketStorage] = None) -> "LSHRS":
        import json
        from pathlib import Path

        src = Path(path)
        if not src.exists():
            raise FileNotFoundError(f"Directory not found: {src}")
        meta = json.loads((src / "metadata.json").read_text())   # FileNotFoundError when absent, like the reference
        if not (src / "projections.npz").exists():   # before any storage / device handle is created
            raise FileNotFoundError(f"No such file or directory: '{src / 'projections.npz'}'")
        cfg = meta["config"]
        red = dict(meta["redis_config"])
        if redis_config:
            red.update(redis_config)
        inst = cls(dim=cfg["dim"], num_perm=cfg["num_perm"], num_bands=cfg["num_bands"],
                   rows_per_band=cfg["rows_per_band"], similarity_threshold=cfg["similarity_threshold"],
                   buffer_size=cfg["buffer_size"], vector_fetch_fn=vector_fetch_fn, storage=storage,
                   redis_host=red["host"], redis_port=red["port"], redis_db=red["db"], redis_password=red["password"],
                   redis_prefix=red["prefix"], decode_responses=red["decode_responses"], seed=cfg["seed"])
        with np.load(src / "projections.npz") as data:
            inst._hasher.projections = [data[f"arr_{i}"].astype(np.float32) for i in range(len(data.files))]
        return inst

    def __getstate__(self) -> dict[str, Any]:
        self.flush()
        return {
            "config": dict(self._config), "redis_config": dict(self._redis_config),
            "projections": [np.asarray(m, dtype=np.float32) for m in self._hasher.projections],
            # the reference does not persist its storage (a Redis connection); the in-memory double travels
            "storage": self._storage if isinstance(self._storage, InMemoryStorage) or self._store_on_device else None,
        }

    def __setstate__(self, state: dict[str, Any]) -> None:
        cfg, red = state["config"], state["redis_config"]
        restored = self.__class__(
            dim=cfg["dim"], num_perm=cfg["num_perm"], num_bands=cfg["num_bands"], rows_per_band=cfg["rows_per_band"],
            similarity_threshold=cfg["similarity_threshold"], buffer_size=cfg["buffer_size"], vector_fetch_fn=None,
            storage=state.get("storage"), redis_host=red["host"], redis_port=red["port"], redis_db=red["db"],
            redis_password=red["password"], redis_prefix=red["prefix"], decode_responses=red["decode_responses"],
            seed=cfg["seed"])
        self.__dict__ = restored.__dict__
        self._hasher.projections = [np.asarray(m, dtype=np.float32) for m in state["projections"]]

    # ------------------------------------------------------------------ helpers
    def _prepare_vector(self, vector: np.ndarray) -> np.ndarray:
        arr = np.asarray(vector, dtype=np.float32).reshape(-1)
        if arr.shape[0] != self._dim:
            raise ValueError(f"Vector must have dimension {self._dim}; received {arr.shape[0]}")
        # np.allclose(arr, 0, atol=1e-8) without its temporaries: NaN makes it False, like numpy
        if bool((np.abs(arr) <= 1e-8).all()):
            raise ValueError(_ZERO_VECTOR_MSG)
        return arr

    def _candidate_counts(self, query_vector: np.ndarray) -> dict[int, int]:
        if self._store_on_device:
            ids, coll = self._device_candidates(query_vector, None)
            return dict(zip(ids.tolist(), coll.tolist()))
        signatures = self._hasher.hash_vector(query_vector)
        counts: dict[int, int] = {}
        for band_id, hash_val in enumerate(signatures):
            for candidate in self._storage.get_bucket(band_id, hash_val):
                counts[candidate] = counts.get(candidate, 0) + 1
        return counts

    def _device_candidates(self, query_vector: np.ndarray, limit: Optional[int]):
        """Candidates of one query from the store in HBM, ordered by (-collisions, id): ``(ids, collisions)``,
        the first ``limit`` of them when given.  One join on the device instead of ``num_bands`` bucket reads
        (reference main.py:1088-1111); two launches and one synchronisation on the latency path."""
        from lshrs_b200 import _native

        ix, hasher = self._dindex, self._hasher
        vec = np.asarray(query_vector, dtype=np.float32).reshape(1, self._dim)
        if hasher._kernel == _native.KERNEL_AUTO and self._dim * 4 <= 65536:
            cap = ix.SMALL_MAX_CAPACITY if limit is None else max(1, min(int(limit), ix.SMALL_MAX_CAPACITY))
            ids, coll, counts, _ = ix.query_vectors(hasher, vec, cap)
            c = int(counts[0])
            if 0 <= c <= cap or (c > cap and limit is not None and limit <= cap):
                take = min(c, cap)
                return ids[0, :take], coll[0, :take]
        ids, coll = ix.query_one(hasher.hash_batch_packed(vec))   # lists beyond the latency path's sizes
        return (ids, coll) if limit is None else (ids[:limit], coll[:limit])

    def _device_rerank(self, q: np.ndarray, top_k: Optional[int], top_p: float):
        """``query(top_p=...)`` with store AND vectors in HBM: hash, join, rerank and id gather in four launches and
        one synchronisation; ``None`` when the query is outside the latency path's sizes or met something the
        general path reports precisely (a candidate id that is not a corpus row, a zero-norm vector)."""
        from lshrs_b200 import _native

        ix, hasher = self._dindex, self._hasher
        if hasher._kernel != _native.KERNEL_AUTO or self._dim * 4 > 65536:
            return None
        cap = ix.SMALL_RERANK_CAPACITY
        stride = max(1, math.ceil(cap * top_p))
        if top_k is not None:
            stride = min(stride, int(top_k))
        rer = _get_reranker(self._dim, hasher.device)
        ids, scores, counts, zero, cands, _ = ix.query_rerank_vectors(
            hasher, rer, q.reshape(1, self._dim), self._corpus, k=int(top_k or 0), p=top_p, stride=stride)
        if cands[0] < 0 or zero[0]:
            return None
        c = int(counts[0])
        return list(zip(ids[0, :c].tolist(), scores[0, :c].tolist()))

    def _fetch_buckets(self, keys: list) -> list:
        """Members of many ``(band_id, band_bytes)`` buckets, in order.

        The reference asks Redis once per band per query (``get_bucket``: one SMEMBERS round trip each,
        reference lshrs/core/main.py:1105-1109).  A batch of queries needs ``nq * num_bands`` buckets, so:
        a backend with ``get_buckets`` gets one call; a reference ``RedisStorage`` gets ONE pipelined
        round trip through its redis-py client; anything else falls back to per-bucket calls.
        """
        storage = self._storage
        if hasattr(storage, "get_buckets"):
            return storage.get_buckets(keys)
        client = getattr(storage, "_client", None)
        if client is not None and hasattr(client, "pipeline") and hasattr(storage, "bucket_key"):
            out: list = []
            for start in range(0, len(keys), 10_000):  # redis-py buffers the whole pipeline client-side
                pipe = client.pipeline()
                try:
                    for band_id, hash_val in keys[start:start + 10_000]:
                        pipe.smembers(storage.bucket_key(band_id, hash_val))
                    replies = pipe.execute()
                finally:
                    pipe.reset()
                out.extend({int(m) for m in reply} for reply in replies)
            return out
        return [storage.get_bucket(band_id, hash_val) for band_id, hash_val in keys]

    def _enqueue_operations(self, index: int, signatures: Union[HashSignatures, Sequence[bytes]]) -> None:
        ops = [(band_id, hash_val, int(index)) for band_id, hash_val in enumerate(signatures)]
        with self._buffer_lock:
            self._buffer.extend(ops)

    def _flush_buffer_if_needed(self) -> None:
        with self._buffer_lock:
            due = len(self._buffer) >= self._buffer_size
        if due:
            self.flush()

    def _require_vector_fetch_fn(self) -> VectorFetchFn:
        if self._vector_fetch_fn is None:
            raise RuntimeError("vector_fetch_fn must be supplied for operations requiring reranking")
        return self._vector_fetch_fn


lshrs = LSHRS

__all__ = ["LSHRS", "lshrs", "VectorFetchFn", "CandidateScores"]
