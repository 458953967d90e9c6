"""ctypes binding of ``liblshx.so`` -- the C-ABI declared in ``include/lshx.h``.

This is the only place the package touches native code.  There is no CPU
fallback: if the shared library is missing (and cannot be built) or no sm_100
device is usable, the first call raises :class:`LshxError` /
:class:`LshxUnavailable` loudly.
"""

from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p
from pathlib import Path

__all__ = [
    "LshxError",
    "LshxUnavailable",
    "lib",
    "lib_path",
    "check",
    "default_device",
    "launch_count",
    "device_count",
    "KERNEL_AUTO",
    "KERNEL_FFMA",
    "KERNEL_TCGEN05",
    "KERNEL_TCGEN05_3XTF32",
    "KERNEL_TCGEN05_TF32BF16",
    "EXPORTED_SYMBOLS",
]

ABI_VERSION = 3   # LSHX_ABI_VERSION of include/lshx.h this binding was written against
KERNEL_AUTO, KERNEL_FFMA, KERNEL_TCGEN05 = 0, 1, 2
KERNEL_TCGEN05_3XTF32 = 4
KERNEL_TCGEN05_TF32BF16 = 5
DTYPE_F32, DTYPE_F16, DTYPE_U8, DTYPE_I8 = 0, 1, 2, 3
ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_ZERO_VECTOR = -1, -2, -3, -4, -5

# every symbol include/lshx.h declares (tests/test_cabi.py checks them against the header)
EXPORTED_SYMBOLS = (
    "lshx_abi_version",
    "lshx_last_error",
    "lshx_device_count",
    "lshx_launch_count",
    "lshx_hasher_create",
    "lshx_hasher_set_projections",
    "lshx_hasher_set_kernel",
    "lshx_hasher_last_kernel",
    "lshx_hasher_signature_bytes",
    "lshx_hash_batch",
    "lshx_hash_batch_typed",
    "lshx_signatures_to_hex",
    "lshx_hasher_destroy",
    "lshx_env_overrides",
    "lshx_hasher_debug_accumulators",
    "lshx_rerank_create",
    "lshx_rerank_topk",
    "lshx_rerank_scores",
    "lshx_l2_normalize",
    "lshx_rerank_destroy",
    "lshx_index_create",
    "lshx_index_destroy",
    "lshx_index_size",
    "lshx_index_add",
    "lshx_index_add_entries",
    "lshx_index_remove",
    "lshx_index_get_buckets",
    "lshx_index_query_host_vectors",
    "lshx_index_query_vectors",
    "lshx_index_query_rerank_vectors",
    "lshx_index_debug_timeline",
    "lshx_index_export",
    "lshx_index_clear",
    "lshx_index_query",
    "lshx_index_fetch",
    "lshx_index_topk",
    "lshx_index_rerank",
    "lshx_ipc_mem_alloc",
    "lshx_ipc_mem_open",
    "lshx_ipc_mem_close",
    "lshx_ipc_mem_free",
    "lshx_ipc_event_create",
    "lshx_ipc_event_open",
    "lshx_ipc_event_record",
    "lshx_ipc_event_wait",
    "lshx_ipc_event_destroy",
    "lshx_memcpy_async",
)


class LshxError(RuntimeError):
    """A liblshx call failed (CUDA error, out of memory, unsupported shape)."""

    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"liblshx error {code}: {message}")
        self.code = code
        self.message = message


class LshxUnavailable(LshxError):
    """The CUDA extension cannot run here (no library, or no sm_100 device)."""


_lock = threading.Lock()
_lib: ctypes.CDLL | None = None


def lib_path() -> Path:
    override = os.environ.get("LSHX_LIBRARY")
    if override:
        return Path(override)
    return Path(__file__).resolve().parent / "_lib" / "liblshx.so"


def _declare(cdll: ctypes.CDLL) -> None:
    vp = c_void_p
    cdll.lshx_abi_version.restype = c_int
    cdll.lshx_abi_version.argtypes = []
    cdll.lshx_last_error.restype = c_char_p
    cdll.lshx_last_error.argtypes = []
    cdll.lshx_device_count.restype = c_int
    cdll.lshx_device_count.argtypes = []
    cdll.lshx_launch_count.restype = c_uint64
    cdll.lshx_launch_count.argtypes = []

    cdll.lshx_hasher_create.restype = c_int
    cdll.lshx_hasher_create.argtypes = [c_int, c_int, c_int, c_int, vp, POINTER(vp)]
    cdll.lshx_hasher_set_projections.restype = c_int
    cdll.lshx_hasher_set_projections.argtypes = [vp, vp]
    cdll.lshx_hasher_set_kernel.restype = c_int
    cdll.lshx_hasher_set_kernel.argtypes = [vp, c_int]
    cdll.lshx_hasher_last_kernel.restype = c_int
    cdll.lshx_hasher_last_kernel.argtypes = [vp]
    cdll.lshx_hasher_signature_bytes.restype = c_int
    cdll.lshx_hasher_signature_bytes.argtypes = [vp]
    cdll.lshx_hash_batch.restype = c_int
    cdll.lshx_hash_batch.argtypes = [vp, vp, c_int64, c_int, vp, c_int, vp, vp]
    cdll.lshx_hash_batch_typed.restype = c_int
    cdll.lshx_hash_batch_typed.argtypes = [vp, vp, c_int, c_int64, vp, vp]
    cdll.lshx_signatures_to_hex.restype = c_int
    cdll.lshx_signatures_to_hex.argtypes = [vp, c_int64, c_int, vp]
    cdll.lshx_hasher_destroy.restype = c_int
    cdll.lshx_hasher_destroy.argtypes = [vp]
    cdll.lshx_env_overrides.restype = c_int
    cdll.lshx_env_overrides.argtypes = []
    cdll.lshx_hasher_debug_accumulators.restype = c_int
    cdll.lshx_hasher_debug_accumulators.argtypes = [vp, vp, c_int64, vp, c_int64, POINTER(c_int), POINTER(c_int)]

    cdll.lshx_rerank_create.restype = c_int
    cdll.lshx_rerank_create.argtypes = [c_int, c_int, POINTER(vp)]
    cdll.lshx_rerank_topk.restype = c_int
    cdll.lshx_rerank_topk.argtypes = [vp, vp, c_int64, vp, c_int64, vp, vp, c_int64, c_int, c_double,
                                      c_int, vp, vp, vp, vp, c_int, vp]
    cdll.lshx_rerank_scores.restype = c_int
    cdll.lshx_rerank_scores.argtypes = [vp, vp, c_int64, vp, c_int64, vp, vp, c_int64, vp, vp, c_int, vp]
    cdll.lshx_l2_normalize.restype = c_int
    cdll.lshx_l2_normalize.argtypes = [vp, vp, c_int64, vp, vp, c_int, vp]
    cdll.lshx_rerank_destroy.restype = c_int
    cdll.lshx_rerank_destroy.argtypes = [vp]

    cdll.lshx_index_create.restype = c_int
    cdll.lshx_index_create.argtypes = [c_int, c_int, c_int, POINTER(vp)]
    cdll.lshx_index_destroy.restype = c_int
    cdll.lshx_index_destroy.argtypes = [vp]
    cdll.lshx_index_size.restype = c_int64
    cdll.lshx_index_size.argtypes = [vp]
    cdll.lshx_index_add.restype = c_int
    cdll.lshx_index_add.argtypes = [vp, vp, vp, c_int64, c_int, vp]
    cdll.lshx_index_add_entries.restype = c_int
    cdll.lshx_index_add_entries.argtypes = [vp, vp, vp, c_int64]
    cdll.lshx_index_get_buckets.restype = c_int
    cdll.lshx_index_get_buckets.argtypes = [vp, vp, vp, c_int64, vp, vp, c_int64, POINTER(c_int64)]
    cdll.lshx_index_query_host_vectors.restype = c_int
    cdll.lshx_index_query_host_vectors.argtypes = [vp, vp, vp, c_int64, vp, POINTER(c_int64), POINTER(c_int64)]
    cdll.lshx_index_query_vectors.restype = c_int
    cdll.lshx_index_query_vectors.argtypes = [vp, vp, vp, c_int, c_int, vp, vp, vp, vp]
    cdll.lshx_index_query_rerank_vectors.restype = c_int
    cdll.lshx_index_query_rerank_vectors.argtypes = [vp, vp, vp, vp, c_int, vp, c_int64, c_int, c_double, c_int,
                                                     vp, vp, vp, vp, vp, vp]
    cdll.lshx_index_debug_timeline.restype = c_int
    cdll.lshx_index_debug_timeline.argtypes = [vp, c_int, vp]
    cdll.lshx_index_export.restype = c_int
    cdll.lshx_index_export.argtypes = [vp, vp, vp, c_int64, POINTER(c_int64)]
    cdll.lshx_index_remove.restype = c_int
    cdll.lshx_index_remove.argtypes = [vp, vp, c_int64]
    cdll.lshx_index_clear.restype = c_int
    cdll.lshx_index_clear.argtypes = [vp]
    cdll.lshx_index_query.restype = c_int
    cdll.lshx_index_query.argtypes = [vp, vp, c_int64, c_int, vp, POINTER(c_int64), POINTER(c_int64)]
    cdll.lshx_index_fetch.restype = c_int
    cdll.lshx_index_fetch.argtypes = [vp, vp, vp, vp, vp]
    cdll.lshx_index_topk.restype = c_int
    cdll.lshx_index_topk.argtypes = [vp, c_int, vp, vp]
    cdll.lshx_index_rerank.restype = c_int
    cdll.lshx_index_rerank.argtypes = [vp, vp, vp, c_int, vp, c_int64, c_int, c_double, c_int, vp, vp, vp, vp]

    from ctypes import c_size_t

    cdll.lshx_ipc_mem_alloc.restype = c_int
    cdll.lshx_ipc_mem_alloc.argtypes = [c_int, c_size_t, POINTER(vp), vp]
    cdll.lshx_ipc_mem_open.restype = c_int
    cdll.lshx_ipc_mem_open.argtypes = [c_int, vp, POINTER(vp)]
    cdll.lshx_ipc_mem_close.restype = c_int
    cdll.lshx_ipc_mem_close.argtypes = [vp]
    cdll.lshx_ipc_mem_free.restype = c_int
    cdll.lshx_ipc_mem_free.argtypes = [c_int, vp]
    cdll.lshx_ipc_event_create.restype = c_int
    cdll.lshx_ipc_event_create.argtypes = [c_int, POINTER(vp), vp]
    cdll.lshx_ipc_event_open.restype = c_int
    cdll.lshx_ipc_event_open.argtypes = [c_int, vp, POINTER(vp)]
    cdll.lshx_ipc_event_record.restype = c_int
    cdll.lshx_ipc_event_record.argtypes = [c_int, vp, vp]
    cdll.lshx_ipc_event_wait.restype = c_int
    cdll.lshx_ipc_event_wait.argtypes = [c_int, vp, vp]
    cdll.lshx_ipc_event_destroy.restype = c_int
    cdll.lshx_ipc_event_destroy.argtypes = [vp]
    cdll.lshx_memcpy_async.restype = c_int
    cdll.lshx_memcpy_async.argtypes = [c_int, vp, vp, c_size_t, vp]


def lib() -> ctypes.CDLL:
    """Load (once) and return liblshx.so; builds it with nvcc if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.environ.get("LSHX_LIBRARY"):
            # (re)build when the library is missing or was built from other sources than the ones in the tree
            # (content hash, lshrs_b200/_lib/liblshx.srchash); a stale library without nvcc is an error, not
            # something to run silently
            from lshrs_b200 import _build

            if _build._stale():
                try:
                    _build.build()
                except Exception as exc:  # noqa: BLE001 - re-raised as the loud failure below
                    what = "is missing" if not path.exists() else "was built from different sources"
                    raise LshxUnavailable(
                        ERR_NO_DEVICE,
                        f"{path} {what} and could not be (re)built ({exc}); lshrs_b200 has no CPU fallback",
                    ) from exc
        try:
            cdll = ctypes.CDLL(str(path))
        except OSError as exc:
            raise LshxUnavailable(ERR_NO_DEVICE, f"cannot load {path}: {exc}; lshrs_b200 has no CPU fallback") from exc
        _declare(cdll)
        if cdll.lshx_abi_version() != ABI_VERSION:
            raise LshxUnavailable(ERR_INVALID_ARG,
                                  f"{path} has ABI version {cdll.lshx_abi_version()}, expected {ABI_VERSION}")
        _lib = cdll
        return cdll


def check(code: int) -> None:
    """Raise for a negative liblshx status."""
    if code >= 0:
        return
    msg = lib().lshx_last_error().decode("utf-8", "replace")
    if code == ERR_NO_DEVICE:
        raise LshxUnavailable(code, msg)
    if code == ERR_OOM:
        raise MemoryError(f"liblshx: {msg}")
    raise LshxError(code, msg)


def default_device() -> int:
    """Device a new hasher / reranker binds to: $LSHRS_B200_DEVICE, else $LOCAL_RANK, else 0."""
    for var in ("LSHRS_B200_DEVICE", "LOCAL_RANK"):
        val = os.environ.get(var)
        if val is not None and val.strip() != "":
            return int(val)
    return 0


def env_overrides() -> list[str]:
    """Names of the LSHX_* tuning overrides set in this process (empty in a product run)."""
    mask = int(lib().lshx_env_overrides())
    names = ("LSHX_TC_FLAGS", "LSHX_TC_SPLIT", "LSHX_COPY_THREADS", "LSHX_BOUNCE_MB", "LSHX_TRACE_PAGEABLE")
    return [n for i, n in enumerate(names) if mask & (1 << i)]


def launch_count() -> int:
    """Kernels launched by liblshx in this process so far."""
    return int(lib().lshx_launch_count())


def device_count() -> int:
    """Usable sm_100 devices (0 on a CPU-only machine)."""
    return int(lib().lshx_device_count())

