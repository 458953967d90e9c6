"""Cosine rerank on the GPU: ``cosine_similarity`` / ``top_k_cosine`` and batched forms.

Drop-in for reference lshrs/utils/similarity.py:26-183.  The reference
normalises every candidate in a Python loop, stacks, and does one sgemv; here
one CUDA kernel (``csrc/rerank.cu``) gathers the candidate rows with 128-bit
loads, accumulates dot(c, q) and dot(c, c) in fp32, forms the cosine and
selects the top results per query.  Scores agree with the reference within
1e-5 (BASELINE.json north_star); ordering is descending score with ties broken
by ascending position (the reference's tie order is unspecified).
"""

from __future__ import annotations

import ctypes
import math
import threading
from collections.abc import Sequence

import numpy as np

from lshrs_b200 import _native

__all__ = ["cosine_similarity", "top_k_cosine", "top_k_cosine_batch", "Reranker"]


class Reranker:
    """Per-(device, dim) handle around ``lshx_rerank_*`` with its staging buffers."""

    def __init__(self, dim: int, device: int | None = None) -> None:
        if dim <= 0:
            raise ValueError("dim must be > 0")
        self.dim = int(dim)
        self.device = _native.default_device() if device is None else int(device)
        handle = ctypes.c_void_p()
        _native.check(_native.lib().lshx_rerank_create(self.device, self.dim, ctypes.byref(handle)))
        self._handle = handle

    def close(self) -> None:
        if self._handle is not None:
            _native.lib().lshx_rerank_destroy(self._handle)
            self._handle = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- scores only -------------------------------------------------------------
    def scores(self, queries: np.ndarray, vectors, offsets: np.ndarray, ids: np.ndarray | None = None,
               *, vectors_on_device: bool = False, n_vectors: int | None = None):
        """Cosine of every candidate slot; returns (scores float32[total], zero_counts int32[nq])."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        total = int(offs[-1])
        out = np.empty(max(total, 0), dtype=np.float32)
        zero = np.zeros(nq, dtype=np.int32)
        ids_arr = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        vptr, nvec = _vectors_ptr(vectors, vectors_on_device, n_vectors, self.dim)
        _native.check(
            _native.lib().lshx_rerank_scores(
                self._handle, q.ctypes.data, nq, vptr, nvec, offs.ctypes.data,
                0 if ids_arr is None else ids_arr.ctypes.data, total, out.ctypes.data, zero.ctypes.data,
                2 if vectors_on_device else 0, None)
        )
        return out, zero

    # -- selection ---------------------------------------------------------------
    def topk(self, queries: np.ndarray, vectors, offsets: np.ndarray, ids: np.ndarray | None = None, *,
             k: int = 0, p: float = 0.0, vectors_on_device: bool = False, n_vectors: int | None = None,
             out: tuple | None = None):
        """Best candidates per query.

        Returns ``(pos int32[nq, stride], score float32[nq, stride], count int32[nq], zero int32[nq])``;
        row i holds ``count[i]`` valid entries, best first.  ``k`` > 0 keeps the k
        best; ``p`` in (0, 1] keeps ``max(1, ceil(n_i * p))`` (and at most ``k`` when both are given).
        ``out`` may supply the four result arrays (e.g. views of pinned memory, so the D2H is a DMA).
        """
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        counts = np.diff(offs)
        maxc = int(counts.max()) if nq else 0
        if p and p > 0:
            stride = max(1, math.ceil(maxc * p))
            if k and k > 0:
                stride = min(stride, int(k))
        else:
            stride = int(k)
        stride = max(1, min(stride, max(maxc, 1)))
        if out is not None:
            pos, score, count, zero = out
            if (pos.shape != (nq, stride) or score.shape != (nq, stride) or pos.dtype != np.int32
                    or score.dtype != np.float32 or count.shape != (nq,) or zero.shape != (nq,)
                    or count.dtype != np.int32 or zero.dtype != np.int32
                    or not (pos.flags.c_contiguous and score.flags.c_contiguous)):
                raise ValueError(f"out arrays must be C-contiguous int32/float32 ({nq}, {stride}) and int32 ({nq},)")
        else:
            pos = np.empty((nq, stride), dtype=np.int32)
            score = np.empty((nq, stride), dtype=np.float32)
            count = np.zeros(nq, dtype=np.int32)
            zero = np.zeros(nq, dtype=np.int32)
        ids_arr = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        vptr, nvec = _vectors_ptr(vectors, vectors_on_device, n_vectors, self.dim)
        _native.check(
            _native.lib().lshx_rerank_topk(
                self._handle, q.ctypes.data, nq, vptr, nvec, offs.ctypes.data,
                0 if ids_arr is None else ids_arr.ctypes.data, maxc, int(k), float(p or 0.0), stride,
                pos.ctypes.data, score.ctypes.data, count.ctypes.data, zero.ctypes.data,
                2 if vectors_on_device else 0, None)
        )
        return pos, score, count, zero


def _vectors_ptr(vectors, on_device: bool, n_vectors: int | None, dim: int):
    if on_device:
        if hasattr(vectors, "data_ptr"):
            if vectors.dim() != 2 or vectors.shape[1] != dim or not vectors.is_contiguous():
                raise ValueError(f"device vectors must be a contiguous (n, {dim}) float32 tensor")
            return int(vectors.data_ptr()), int(vectors.shape[0])
        if n_vectors is None:
            raise ValueError("n_vectors is required with a raw device pointer")
        return int(vectors), int(n_vectors)
    arr = vectors  # caller passes a contiguous float32 (n, dim) ndarray
    return arr.ctypes.data, int(arr.shape[0])


_rerankers: dict[tuple[int, int], Reranker] = {}
_rerankers_lock = threading.Lock()


def _get_reranker(dim: int, device: int | None = None) -> Reranker:
    dev = _native.default_device() if device is None else int(device)
    key = (dev, int(dim))
    r = _rerankers.get(key)
    if r is None:
        with _rerankers_lock:
            r = _rerankers.get(key)
            if r is None:
                r = Reranker(dim, dev)
                _rerankers[key] = r
    return r


def _stack_candidates(candidates) -> np.ndarray:
    """Candidates as one contiguous float32 (n, dim) array; each candidate flattened like l2_norm does."""
    if isinstance(candidates, np.ndarray) and candidates.ndim == 2:
        return np.ascontiguousarray(candidates, dtype=np.float32)
    rows = [np.asarray(vec, dtype=np.float32).reshape(-1) for vec in candidates]
    if not rows:
        # the reference reaches np.stack([]) here (similarity.py:85)
        raise ValueError("need at least one array to stack")
    return np.ascontiguousarray(np.stack(rows))


def _check_zero(zero: np.ndarray) -> None:
    if zero.any():
        raise ValueError("Cannot normalize zero vector")


def cosine_similarity(query: np.ndarray, candidates: Sequence[np.ndarray]) -> np.ndarray:
    """Cosine between one query and every candidate (reference similarity.py:26-90).

    Raises ``ValueError("Cannot normalize zero vector")`` when the query or any
    candidate has zero norm, as the reference's ``l2_norm`` does.
    """
    q = np.asarray(query, dtype=np.float32).reshape(-1)
    cands = _stack_candidates(candidates)
    if cands.shape[1] != q.shape[0]:
        raise ValueError(f"shapes {cands.shape} and {q.shape} not aligned")
    n = cands.shape[0]
    rer = _get_reranker(q.shape[0])
    scores, zero = rer.scores(q, cands, np.array([0, n], dtype=np.int64))
    _check_zero(zero)
    return scores


def top_k_cosine(query: np.ndarray, candidates: Sequence[np.ndarray], *, k: int) -> list[tuple[int, float]]:
    """The ``k`` most similar candidates as ``(position, score)``, best first (reference similarity.py:93-183)."""
    if k <= 0:
        raise ValueError("k must be > 0")
    q = np.asarray(query, dtype=np.float32).reshape(-1)
    cands = _stack_candidates(candidates)
    if cands.shape[1] != q.shape[0]:
        raise ValueError(f"shapes {cands.shape} and {q.shape} not aligned")
    n = cands.shape[0]
    rer = _get_reranker(q.shape[0])
    pos, score, count, zero = rer.topk(q, cands, np.array([0, n], dtype=np.int64), k=min(int(k), n))
    _check_zero(zero)
    c = int(count[0])
    return [(int(pos[0, i]), float(score[0, i])) for i in range(c)]


def top_k_cosine_batch(queries: np.ndarray, vectors, offsets, ids=None, *, k: int = 0, p: float = 0.0,
                       vectors_on_device: bool = False, device: int | None = None):
    """Batched rerank: many queries, CSR candidate lists, optional gather ids into a corpus.

    ``vectors`` is a host ndarray ``(N, dim)`` or (``vectors_on_device``) a CUDA
    torch tensor resident in HBM.  Returns ``(pos, score, count)`` arrays (see
    :meth:`Reranker.topk`); raises ``ValueError`` on zero-norm vectors.
    """
    if (k is None or k <= 0) and not (p and p > 0):
        raise ValueError("k must be > 0")
    if p and not 0 < p <= 1:
        raise ValueError("top_p must be within the range (0, 1]")
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim != 2:
        raise ValueError("queries must be a 2D array")
    if not vectors_on_device:
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        if vectors.ndim != 2 or vectors.shape[1] != q.shape[1]:
            raise ValueError(f"vectors must have shape (n, {q.shape[1]})")
    rer = _get_reranker(q.shape[1], device)
    pos, score, count, zero = rer.topk(q, vectors, offsets, ids, k=int(k or 0), p=float(p or 0.0),
                                       vectors_on_device=vectors_on_device)
    _check_zero(zero)
    return pos, score, count
