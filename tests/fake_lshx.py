"""A CPU stand-in for ``liblshx.so`` built on the oracle -- TEST DOUBLE, never shipped, never timed.

``install()`` puts an object with the C ABI's entry points (``include/lshx.h``) in place of the ctypes
library inside ``lshrs_b200._native``.  Everything ABOVE the C ABI -- ``LSHHasher``, ``LSHRS``, the rerank
wrappers, the ``lshrs`` compat package -- then runs unchanged on a machine without a GPU, which is how the
reference's own test-suite exercises the host logic in the ``-m "not gpu"`` tier
(tests/test_reference_suite.py).  The arithmetic is the oracle's (the reference's numpy path), so this
double says nothing about the kernels; the ``-m gpu`` tier runs the same suite on the real library.
"""

from __future__ import annotations

import ctypes
import math

import numpy as np

from oracle import lshrs_oracle as oracle


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=dtype)
    ptr = ptr.value if isinstance(ptr, ctypes.c_void_p) else int(ptr)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _null(ptr) -> bool:
    if ptr is None:
        return True
    if isinstance(ptr, ctypes.c_void_p):
        return not ptr.value
    return int(ptr) == 0


class _Hasher:
    def __init__(self, dim, nb, r, R):
        self.dim, self.nb, self.r = dim, nb, r
        self.bpb = (r + 7) // 8
        self.set(R)

    def set(self, R):
        self.projs = [np.array(R[b * self.r:(b + 1) * self.r], dtype=np.float32) for b in range(self.nb)]


class FakeLshx:
    """The subset of ctypes.CDLL behaviour lshrs_b200 relies on: attribute = callable taking raw addresses."""

    def __init__(self):
        self._handles: dict[int, object] = {}
        self._next = 0x1000
        self._err = b""
        self.launches = 0

    # ---- library
    def lshx_abi_version(self):
        from lshrs_b200 import _native

        return _native.ABI_VERSION

    def lshx_last_error(self):
        return self._err

    def lshx_device_count(self):
        return 1

    def lshx_launch_count(self):
        return self.launches

    def lshx_env_overrides(self):
        return 0

    def _new(self, obj, out_ref):
        self._next += 16
        self._handles[self._next] = obj
        out_ref._obj.value = self._next
        return 0

    def _get(self, h):
        return self._handles[h.value if isinstance(h, ctypes.c_void_p) else int(h)]

    # ---- hasher
    def lshx_hasher_create(self, device, dim, nb, r, R_ptr, out_ref):
        if dim <= 0 or nb <= 0 or r <= 0:
            self._err = b"bad size"
            return -1
        return self._new(_Hasher(dim, nb, r, _arr(R_ptr, (nb * r, dim), np.float32)), out_ref)

    def lshx_hasher_set_projections(self, h, R_ptr):
        hs = self._get(h)
        hs.set(_arr(R_ptr, (hs.nb * hs.r, hs.dim), np.float32))
        return 0

    def lshx_hasher_set_kernel(self, h, kernel):
        return 0

    def lshx_hasher_last_kernel(self, h):
        return 1

    def lshx_hasher_signature_bytes(self, h):
        hs = self._get(h)
        return hs.nb * hs.bpb

    def _hash(self, hs, X, out_ptr, flag_ptr):
        n = X.shape[0]
        out = _arr(out_ptr, (n, hs.nb, hs.bpb), np.uint8)
        out[...] = oracle.hash_batch_packed(hs.projs, X) if n <= 64 else oracle.hash_batch_vectorized(hs.projs, X)
        if not _null(flag_ptr):
            _arr(flag_ptr, (n,), np.uint8)[...] = [1 if oracle.is_zero_vector(x) else 0 for x in X]
        self.launches += 1
        return 0

    def lshx_hash_batch(self, h, X_ptr, n, x_dev, out_ptr, out_dev, flag_ptr, stream):
        hs = self._get(h)
        if n == 0:
            return 0
        return self._hash(hs, _arr(X_ptr, (n, hs.dim), np.float32), out_ptr, flag_ptr)

    def lshx_hash_batch_typed(self, h, X_ptr, dtype, n, out_ptr, flag_ptr):
        hs = self._get(h)
        np_dtype = {0: np.float32, 1: np.float16, 2: np.uint8, 3: np.int8}[dtype]
        if n == 0:
            return 0
        return self._hash(hs, _arr(X_ptr, (n, hs.dim), np_dtype).astype(np.float32), out_ptr, flag_ptr)

    def lshx_signatures_to_hex(self, sig_ptr, n, sig_bytes, hex_ptr):
        sig = _arr(sig_ptr, (n * sig_bytes,), np.uint8)
        _arr(hex_ptr, (2 * n * sig_bytes,), np.uint8)[...] = np.frombuffer(sig.tobytes().hex().encode(), np.uint8)
        return 0

    def lshx_hasher_destroy(self, h):
        self._handles.pop(h.value if isinstance(h, ctypes.c_void_p) else int(h), None)
        return 0

    # ---- reranker
    def lshx_rerank_create(self, device, dim, out_ref):
        return self._new({"dim": dim}, out_ref)

    def lshx_rerank_destroy(self, h):
        return self.lshx_hasher_destroy(h)

    def _scores(self, dim, Q, V, offs, ids, zero):
        """float32 cosine per slot; NaN for slots whose vector (or query) has zero norm."""
        out = np.empty(int(offs[-1]), dtype=np.float32)
        for i in range(Q.shape[0]):
            s0, s1 = int(offs[i]), int(offs[i + 1])
            rows = V[ids[s0:s1]] if ids is not None else V[s0:s1]
            qn = np.linalg.norm(Q[i])
            nz = 1 if qn == 0 else 0
            for j, c in enumerate(rows):
                cn = np.linalg.norm(c)
                if cn == 0 or qn == 0:
                    out[s0 + j] = np.nan
                    nz += 1 if cn == 0 else 0
                else:
                    out[s0 + j] = np.float32(np.dot(c / cn, Q[i] / qn))
            zero[i] = nz
        return out

    def lshx_rerank_scores(self, h, Q_ptr, nq, V_ptr, nvec, offs_ptr, ids_ptr, total, out_ptr, zero_ptr, on_dev, stream):
        dim = self._get(h)["dim"]
        offs = _arr(offs_ptr, (nq + 1,), np.int64)
        ids = None if _null(ids_ptr) else _arr(ids_ptr, (int(offs[-1]),), np.int64)
        zero = np.zeros(nq, np.int32)
        sc = self._scores(dim, _arr(Q_ptr, (nq, dim), np.float32), _arr(V_ptr, (nvec, dim), np.float32), offs, ids, zero)
        _arr(out_ptr, (int(offs[-1]),), np.float32)[...] = sc
        if not _null(zero_ptr):
            _arr(zero_ptr, (nq,), np.int32)[...] = zero
        self.launches += 1
        return 0

    def lshx_rerank_topk(self, h, Q_ptr, nq, V_ptr, nvec, offs_ptr, ids_ptr, maxc, k, p, stride, pos_ptr, score_ptr,
                         count_ptr, zero_ptr, on_dev, stream):
        dim = self._get(h)["dim"]
        offs = _arr(offs_ptr, (nq + 1,), np.int64)
        ids = None if _null(ids_ptr) else _arr(ids_ptr, (int(offs[-1]),), np.int64)
        zero = np.zeros(nq, np.int32)
        sc = self._scores(dim, _arr(Q_ptr, (nq, dim), np.float32), _arr(V_ptr, (nvec, dim), np.float32), offs, ids, zero)
        pos = _arr(pos_ptr, (nq, stride), np.int32)
        score = _arr(score_ptr, (nq, stride), np.float32)
        count = _arr(count_ptr, (nq,), np.int32)
        for i in range(nq):
            s = sc[int(offs[i]):int(offs[i + 1])]
            n = s.shape[0]
            limit = n
            if p > 0:
                limit = max(1, math.ceil(n * p))
                if k > 0:
                    limit = min(limit, k)
            elif k > 0:
                limit = min(k, n)
            limit = min(limit, n, stride)
            order = np.lexsort((np.arange(n), -np.nan_to_num(s, nan=-np.inf)))[:limit]
            pos[i, :limit] = order
            score[i, :limit] = s[order]
            count[i] = limit
        if not _null(zero_ptr):
            _arr(zero_ptr, (nq,), np.int32)[...] = zero
        self.launches += 1
        return 0

    def lshx_l2_normalize(self, h, X_ptr, n, out_ptr, zero_ptr, on_dev, stream):
        dim = self._get(h)["dim"]
        X = _arr(X_ptr, (n, dim), np.float32)
        out = _arr(out_ptr, (n, dim), np.float32)
        zero = np.zeros(n, np.int32)
        for i in range(n):
            nrm = np.linalg.norm(X[i])
            if nrm == 0:
                zero[i] = 1
                out[i] = 0
            else:
                out[i] = X[i] / nrm
        if not _null(zero_ptr):
            _arr(zero_ptr, (n,), np.int32)[...] = zero
        self.launches += 1
        return 0


def install() -> FakeLshx:
    """Replace the loaded library inside lshrs_b200._native; returns the double."""
    from lshrs_b200 import _native

    fake = FakeLshx()
    _native._lib = fake
    return fake
