"""``LSHHasher`` -- banded random-projection signatures on a B200.

Drop-in for the reference class (reference lshrs/hash/lsh.py:18-247): same
constructor, attributes, methods, validation and error behaviour; the
arithmetic (projection, sign test, little-endian bit packing) runs in the
hand-written sm_100a kernels of ``liblshx.so`` instead of numpy.  The
projection matrices themselves are still drawn on the host with numpy's PCG64
stream exactly as the reference does (lsh.py:93-94), so a hasher built with the
same ``(num_bands, rows_per_band, dim, seed)`` has bit-identical planes.

Beyond the reference surface there is a packed-array fast path
(:meth:`hash_batch_packed`, :meth:`hash_device`) -- the reference's
``list[HashSignatures]`` of Python ``bytes`` cannot be produced at GPU speed, so
throughput is only meaningful at the ``uint8[n, num_bands, ceil(r/8)]``
boundary.
"""

from __future__ import annotations

import ctypes
import threading
from typing import Any

import numpy as np

from lshrs_b200 import _native
from lshrs_b200._config.config import HashSignatures, signatures_from_packed

__all__ = ["LSHHasher"]


def _as_ptr(obj: Any) -> int:
    """Address of a numpy array / torch tensor / raw int pointer."""
    if obj is None:
        return 0
    if isinstance(obj, int):
        return obj
    if isinstance(obj, np.ndarray):
        return obj.ctypes.data
    if hasattr(obj, "data_ptr"):
        return int(obj.data_ptr())
    raise TypeError(f"cannot take the address of {type(obj)!r}")


class LSHHasher:
    """Random-projection LSH hasher whose hot path runs on the GPU.

    Attributes (as in the reference): ``num_bands``, ``rows_per_band``, ``dim``
    and ``projections`` -- a list of ``num_bands`` float32 arrays of shape
    ``(rows_per_band, dim)``.  ``projections`` may be re-bound from outside
    (``LSHRS.load_from_disk`` / ``__setstate__`` do, reference
    lshrs/core/main.py:981, 1044); the device copy is refreshed on the next
    hash.  After mutating an array IN PLACE call :meth:`sync_projections`.

    Arithmetic and the call path.  In the reference ``hash_batch`` IS ``hash_vector`` per row, so a vector
    always gets the same bytes.  Here a call of up to 32 host rows runs the FP32 latency kernel, larger
    batches the tensor-core kernel (scaled FP16x3 split, FP32 accumulation) or the FFMA kernel (``dim % 4 != 0``,
    unaligned input).  They agree -- and agree with the reference -- on every bit whose projection satisfies
    ``|x.r| > 1e-5 |x||r|`` (measured: the largest margin at which any arm ever differed from a plain FP32
    computation is 2.6e-7, tests/test_config1_gpu.py); a bit closer to zero than that is decided by rounding in ANY float32
    implementation and may differ between ``index()`` (batch) and ``query()`` (one vector) of the same vector,
    about one band key in 10^5 on Gaussian data.  ``set_kernel("ffma")`` pins one FP32 arithmetic for all batch
    sizes above 32 rows when that matters more than throughput.
    """

    def __init__(self, num_bands: int, rows_per_band: int, dim: int, seed: int = 42, *, device: int | None = None) -> None:
        # same checks and messages as reference lsh.py:78-83
        if num_bands <= 0:
            raise ValueError("num_bands must be > 0")
        if rows_per_band <= 0:
            raise ValueError("rows_per_band must be > 0")
        if dim <= 0:
            raise ValueError("dim must be > 0")
        self.num_bands = int(num_bands)
        self.rows_per_band = int(rows_per_band)
        self.dim = int(dim)
        self.seed = seed
        self.bytes_per_band = (self.rows_per_band + 7) // 8
        self.signature_bytes = self.num_bands * self.bytes_per_band
        self._device = _native.default_device() if device is None else int(device)
        self._lock = threading.Lock()
        self._handle: ctypes.c_void_p | None = None
        # the arrays whose contents are on the device: held by reference so that CPython cannot recycle their
        # id() for a new array while we still compare against it
        self._uploaded: tuple[np.ndarray, ...] | None = None
        self._kernel = _native.KERNEL_AUTO
        # host draw, identical to the reference (float64 draws cast to float32, one rng, band order)
        rng = np.random.default_rng(seed)
        self._projections = [
            rng.standard_normal((self.rows_per_band, self.dim)).astype(np.float32) for _ in range(self.num_bands)
        ]

    # ------------------------------------------------------------------ projections
    @property
    def projections(self) -> list[np.ndarray]:
        return self._projections

    @projections.setter
    def projections(self, value) -> None:
        self._projections = list(value)
        self._uploaded = None  # re-upload lazily

    def sync_projections(self) -> None:
        """Force a re-upload of ``projections`` (after in-place edits of the arrays)."""
        self._uploaded = None

    def _stacked_projections(self) -> np.ndarray:
        mats = []
        for m in self._projections:
            a = np.asarray(m, dtype=np.float32)
            if a.shape != (self.rows_per_band, self.dim):
                raise ValueError(
                    f"projection matrix of shape {a.shape} does not match ({self.rows_per_band}, {self.dim})"
                )
            mats.append(a)
        if len(mats) != self.num_bands:
            raise ValueError(f"expected {self.num_bands} projection matrices, found {len(mats)}")
        return np.ascontiguousarray(np.concatenate(mats, axis=0))

    def _is_current(self) -> bool:
        up, cur = self._uploaded, self._projections
        return up is not None and len(up) == len(cur) and all(a is b for a, b in zip(up, cur))

    def _ensure_handle(self) -> ctypes.c_void_p:
        if self._handle is not None and self._is_current():
            return self._handle
        with self._lock:
            if self._handle is not None and self._is_current():
                return self._handle
            lib = _native.lib()
            current = tuple(self._projections)
            stacked = self._stacked_projections()
            if self._handle is None:
                handle = ctypes.c_void_p()
                _native.check(
                    lib.lshx_hasher_create(self._device, self.dim, self.num_bands, self.rows_per_band,
                                           stacked.ctypes.data, ctypes.byref(handle))
                )
                self._handle = handle
                if self._kernel != _native.KERNEL_AUTO:
                    _native.check(lib.lshx_hasher_set_kernel(self._handle, self._kernel))
            else:
                _native.check(lib.lshx_hasher_set_projections(self._handle, stacked.ctypes.data))
            self._uploaded = current
            return self._handle

    # ------------------------------------------------------------------ kernel choice
    def set_kernel(self, kernel: str | int) -> None:
        """Choose the projection kernel: "auto", "ffma", "tcgen05" (scaled FP16x3 operand split), or the
        same tensor-core kernel with another split: "tcgen05_3xtf32", "tcgen05_tf32bf16"."""
        table = {"auto": _native.KERNEL_AUTO, "ffma": _native.KERNEL_FFMA, "tcgen05": _native.KERNEL_TCGEN05,
                 "tcgen05_3xtf32": _native.KERNEL_TCGEN05_3XTF32,
                 "tcgen05_tf32bf16": _native.KERNEL_TCGEN05_TF32BF16}
        code = table[kernel] if isinstance(kernel, str) else int(kernel)
        self._kernel = code
        if self._handle is not None:
            _native.check(_native.lib().lshx_hasher_set_kernel(self._handle, code))

    @property
    def last_kernel(self) -> str:
        if self._handle is None:
            return "none"
        code = _native.lib().lshx_hasher_last_kernel(self._handle)
        return {0: "none", 1: "ffma", 2: "tcgen05", 3: "small", 4: "tcgen05_3xtf32", 5: "tcgen05_tf32bf16"}.get(code, str(code))

    @property
    def device(self) -> int:
        return self._device

    # ------------------------------------------------------------------ reference API
    def hash_vector(self, vector: np.ndarray) -> HashSignatures:
        """Hash one vector (reference lsh.py:96-134)."""
        vec = self._validate_vector(vector)
        packed = self._hash_host(vec.reshape(1, self.dim))
        return HashSignatures.from_packed(packed[0], self.bytes_per_band)

    def hash_batch(self, vectors: np.ndarray) -> list[HashSignatures]:
        """Hash a 2-D batch, one ``HashSignatures`` per row (reference lsh.py:136-169)."""
        arr = self._validate_batch(vectors)
        if arr.shape[0] == 0:
            return []
        return signatures_from_packed(self._hash_host(arr), self.bytes_per_band)

    def _project_and_pack(self, projection: np.ndarray, vector: np.ndarray) -> bytes:
        """One band of one vector with an explicit matrix (reference lsh.py:171-211).

        Kept for API compatibility; it runs the same kernel through a
        one-band hasher bound to ``projection``.
        """
        proj = np.ascontiguousarray(projection, dtype=np.float32)
        if proj.ndim != 2:
            raise ValueError("projection must be a 2D array")
        vec = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        if vec.shape[0] != proj.shape[1]:
            raise ValueError(f"Expected vector of dimension {proj.shape[1]}, received {vec.shape}")
        tmp = LSHHasher.__new__(LSHHasher)
        LSHHasher.__init__(tmp, 1, proj.shape[0], proj.shape[1], seed=0, device=self._device)
        tmp.projections = [proj]
        try:
            return tmp.hash_vector(vec).bands[0]
        finally:
            tmp.close()

    def _validate_vector(self, vector: np.ndarray) -> np.ndarray:
        """float32, flattened, length ``dim`` (reference lsh.py:213-247)."""
        vec = np.asarray(vector, dtype=np.float32).reshape(-1)
        if vec.ndim != 1 or vec.shape[0] != self.dim:
            raise ValueError(f"Expected vector of dimension {self.dim}, received {vec.shape}")
        return vec

    # host dtypes whose cast to float32 is exact and that the library casts ON THE DEVICE for large
    # batches (lshx_hash_batch_typed): half / a quarter of the float32 bytes over PCIe, same signatures
    _TYPED = {np.dtype(np.float16): _native.DTYPE_F16, np.dtype(np.uint8): _native.DTYPE_U8,
              np.dtype(np.int8): _native.DTYPE_I8}
    _TYPED_MIN_ROWS = 4096

    def _validate_batch(self, vectors) -> np.ndarray:
        if (isinstance(vectors, np.ndarray) and vectors.dtype in self._TYPED and vectors.ndim == 2
                and vectors.shape[1] == self.dim and vectors.shape[0] >= self._TYPED_MIN_ROWS):
            return np.ascontiguousarray(vectors)   # stays typed: _hash_host takes the typed entry point
        arr = np.asarray(vectors, dtype=np.float32)
        if arr.ndim != 2:
            raise ValueError("Batch input must be a 2D array")
        if arr.shape[1] != self.dim:
            raise ValueError(f"Expected vectors of dimension {self.dim}, received {arr.shape[1]}")
        return arr

    # ------------------------------------------------------------------ packed fast paths
    def _hash_host(self, arr: np.ndarray, zero_flag: np.ndarray | None = None) -> np.ndarray:
        handle = self._ensure_handle()
        n = arr.shape[0]
        out = np.empty((n, self.signature_bytes), dtype=np.uint8)
        if arr.dtype in self._TYPED and n >= self._TYPED_MIN_ROWS:
            x = np.ascontiguousarray(arr)
            _native.check(
                _native.lib().lshx_hash_batch_typed(handle, x.ctypes.data, self._TYPED[arr.dtype], n, out.ctypes.data,
                                                    0 if zero_flag is None else zero_flag.ctypes.data)
            )
            return out
        x = np.ascontiguousarray(arr, dtype=np.float32)
        _native.check(
            _native.lib().lshx_hash_batch(handle, x.ctypes.data, n, 0, out.ctypes.data, 0,
                                          0 if zero_flag is None else zero_flag.ctypes.data, None)
        )
        return out

    def hash_batch_packed(self, vectors, *, return_zero_flag: bool = False):
        """Hash a host batch into ``uint8[n, num_bands, bytes_per_band]``.

        Same bytes as ``hash_batch`` (row i, band b == ``hash_batch(v)[i][b]``)
        without building Python objects.  With ``return_zero_flag`` also returns
        ``uint8[n]`` marking rows that ``LSHRS._prepare_vector`` would reject
        as zero vectors (reference lshrs/core/main.py:1083).
        """
        arr = self._validate_batch(vectors)
        n = arr.shape[0]
        flag = np.zeros(n, dtype=np.uint8) if return_zero_flag else None
        if n == 0:
            out = np.empty((0, self.signature_bytes), dtype=np.uint8)
        else:
            out = self._hash_host(arr, flag)
        out = out.reshape(n, self.num_bands, self.bytes_per_band)
        return (out, flag) if return_zero_flag else out

    def hash_into(self, x, n: int, out, *, x_on_device: bool, out_on_device: bool, zero_flag=None, stream=None) -> None:
        """Raw C-ABI call: ``x`` / ``out`` / ``zero_flag`` are numpy arrays, torch tensors or addresses.

        ``x`` is ``n x dim`` float32 contiguous; ``out`` is ``n x signature_bytes``
        uint8; device pointers must live on this hasher's device.  With both on
        the device the launch is asynchronous on ``stream`` (a ``cudaStream_t``
        address, e.g. ``torch.cuda.current_stream().cuda_stream``).
        """
        handle = self._ensure_handle()
        _native.check(
            _native.lib().lshx_hash_batch(handle, _as_ptr(x), int(n), 1 if x_on_device else 0, _as_ptr(out),
                                          1 if out_on_device else 0, _as_ptr(zero_flag),
                                          ctypes.c_void_p(stream) if stream else None)
        )

    def hash_device(self, x, *, out=None, zero_flag=None, stream=None):
        """Hash a CUDA torch tensor ``(n, dim)`` float32 already resident in HBM.

        Returns a CUDA uint8 tensor ``(n, num_bands, bytes_per_band)``; the
        launch is asynchronous on the current torch stream.
        """
        import torch

        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise TypeError("hash_device expects a CUDA torch tensor")
        if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.dim:
            raise ValueError(f"Expected a float32 tensor of shape (n, {self.dim}); received {tuple(x.shape)} {x.dtype}")
        if x.device.index != self._device:
            raise ValueError(f"tensor is on cuda:{x.device.index}, hasher on cuda:{self._device}")
        x = x.contiguous()
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, self.num_bands, self.bytes_per_band), dtype=torch.uint8, device=x.device)
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        if n:
            self.hash_into(x, n, out, x_on_device=True, out_on_device=True, zero_flag=zero_flag, stream=stream)
        return out

    def debug_accumulators(self, vectors) -> np.ndarray:
        """DIAGNOSTICS: the tcgen05 kernel's raw fp32 accumulators for the first tile of ``vectors``
        (``lshx_hasher_debug_accumulators``): ``float32[rows, cols]``, column c = signature bit c for
        byte-aligned bands.  Used by the tests to measure the split arithmetic on the hardware."""
        arr = np.ascontiguousarray(self._validate_batch(np.asarray(vectors, dtype=np.float32)), dtype=np.float32)
        handle = self._ensure_handle()
        out = np.zeros((256, 256), dtype=np.float32)
        rows, cols = ctypes.c_int(0), ctypes.c_int(0)
        _native.check(_native.lib().lshx_hasher_debug_accumulators(
            handle, arr.ctypes.data, arr.shape[0], out.ctypes.data, out.size, ctypes.byref(rows), ctypes.byref(cols)))
        return out.reshape(-1)[: rows.value * cols.value].reshape(rows.value, cols.value).copy()

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        """Free the device copy of the projections and the staging buffers."""
        with self._lock:
            if self._handle is not None:
                _native.lib().lshx_hasher_destroy(self._handle)
                self._handle = None
                self._uploaded = None

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __getstate__(self) -> dict:
        return {
            "num_bands": self.num_bands, "rows_per_band": self.rows_per_band, "dim": self.dim,
            "seed": self.seed, "device": self._device,
            "projections": [np.asarray(m, dtype=np.float32) for m in self._projections],
        }

    def __setstate__(self, state: dict) -> None:
        LSHHasher.__init__(self, state["num_bands"], state["rows_per_band"], state["dim"], state["seed"],
                           device=state.get("device"))
        self.projections = state["projections"]
