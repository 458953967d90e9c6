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

import numpy as np

from lshrs_b200 import _native

__all__ = ["DeviceIndex"]


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

    # ---- mirror of the storage operations
    def add(self, signatures: np.ndarray, ids) -> None:
        """``signatures``: uint8 ``(n, num_bands, bytes_per_band)`` as ``hash_batch_packed`` returns them."""
        sig = np.ascontiguousarray(signatures, dtype=np.uint8)
        idx = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        n = idx.shape[0]
        if sig.size != n * self.num_bands * self.bytes_per_band:
            raise ValueError(f"signatures of {sig.shape} do not match {n} ids x {self.num_bands} x {self.bytes_per_band}")
        if n:
            _native.check(_native.lib().lshx_index_add(self._handle, sig.ctypes.data, idx.ctypes.data, n, 0, None))

    def add_device(self, signatures, ids, stream=None) -> None:
        """Same with CUDA tensors (uint8 signatures / int64 ids resident in HBM): nothing crosses PCIe."""
        n = int(ids.numel())
        if n:
            _native.check(_native.lib().lshx_index_add(
                self._handle, int(signatures.data_ptr()), int(ids.data_ptr()), n, 1,
                ctypes.c_void_p(stream) if stream else None))

    def remove(self, ids) -> None:
        idx = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        if idx.shape[0]:
            _native.check(_native.lib().lshx_index_remove(self._handle, idx.ctypes.data, idx.shape[0]))

    def clear(self) -> None:
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
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = self._nq
        ids = np.empty((nq, stride), dtype=np.int64)
        scores = np.empty((nq, stride), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        zero = np.zeros(nq, dtype=np.int32)
        if nq:
            if corpus.dim() != 2 or corpus.shape[1] != reranker.dim or not corpus.is_contiguous():
                raise ValueError(f"corpus must be a contiguous (n, {reranker.dim}) float32 CUDA tensor")
            _native.check(_native.lib().lshx_index_rerank(
                self._handle, reranker._handle, q.ctypes.data, 0, int(corpus.data_ptr()), int(corpus.shape[0]),
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
