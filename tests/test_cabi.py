"""The C-ABI library loads on a CPU-only machine and exports what include/lshx.h declares."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from lshrs_b200 import _build, _native

REPO = Path(__file__).resolve().parents[1]


def _declared_symbols() -> list[str]:
    text = (REPO / "include" / "lshx.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # drop comments
    return sorted(set(re.findall(r"\b(lshx_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = _build.build()
    assert path.exists()
    lib = _native.lib()
    assert lib.lshx_abi_version() == _native.ABI_VERSION == 3


def test_every_declared_symbol_is_exported():
    declared = _declared_symbols()
    assert declared == sorted(_native.EXPORTED_SYMBOLS)
    cdll = ctypes.CDLL(str(_native.lib_path()))
    for name in declared:
        assert getattr(cdll, name) is not None, name


def test_library_has_no_hard_driver_dependency():
    # cudart is static and the driver API is resolved at run time, so the .so loads without libcuda
    import subprocess

    out = subprocess.run(["ldd", str(_native.lib_path())], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out
    assert "not found" not in out


def test_hex_helper_is_bucket_key_hex():
    sig = np.array([[0xAB, 0xCD, 0x00, 0x0F]], dtype=np.uint8)
    out = ctypes.create_string_buffer(8)
    _native.check(_native.lib().lshx_signatures_to_hex(sig.ctypes.data, 1, 4, out))
    assert out.raw == b"abcd000f" == sig.tobytes().hex().encode()


def test_no_cpu_fallback_without_a_gpu():
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    import lshrs_b200

    hasher = lshrs_b200.LSHHasher(2, 3, 4)
    with pytest.raises(lshrs_b200.LshxUnavailable):
        hasher.hash_vector(np.ones(4, dtype=np.float32))
    with pytest.raises(lshrs_b200.LshxUnavailable):
        lshrs_b200.top_k_cosine(np.ones(4), [np.ones(4)], k=1)


def test_create_validates_arguments_like_the_reference():
    # lsh.py:78-83 -> LSHX_ERR_INVALID_ARG before any device work
    lib = _native.lib()
    handle = ctypes.c_void_p()
    R = np.ones((1, 1), dtype=np.float32)
    for nb, r, dim in ((0, 1, 1), (1, 0, 1), (1, 1, 0)):
        rc = lib.lshx_hasher_create(0, dim, nb, r, R.ctypes.data, ctypes.byref(handle))
        assert rc == -1 and b"must be > 0" in lib.lshx_last_error()


def test_product_never_imports_the_oracle():
    for path in (REPO / "lshrs_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in text.replace("no CPU fallback", ""), f"{path} mentions the oracle"
