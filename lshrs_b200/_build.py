"""Build liblshx.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m lshrs_b200._build [--force]

The shared object lands in ``lshrs_b200/_lib/liblshx.so`` so that it travels
with a snapshot of the repo to the GPU box (it is git-ignored, not
gpurun-ignored).  cudart is linked statically and the driver API is resolved at
run time through ``cudaGetDriverEntryPoint``, so the library loads (and exports
every symbol of ``include/lshx.h``) on a machine without a GPU driver.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "_lib"
LIB_PATH = LIB_DIR / "liblshx.so"
SOURCES = ["lshx_api.cu", "hash_ffma.cu", "hash_tc.cu", "rerank.cu", "index_join.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; liblshx.so cannot be built")


HASH_PATH = LIB_DIR / "liblshx.srchash"   # sha256 of the sources the shipped library was built from


def source_hash() -> str:
    """Content hash of everything the library is compiled from (mtimes do not survive a snapshot copy)."""
    import hashlib

    h = hashlib.sha256()
    deps = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [REPO / "include" / "lshx.h"]
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(ARCH_FLAGS).encode())
    return h.hexdigest()


def _stale() -> bool:
    if not LIB_PATH.exists() or not HASH_PATH.exists():
        return True
    return HASH_PATH.read_text().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link liblshx.so; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    LIB_DIR.mkdir(exist_ok=True)
    import fcntl

    with open(LIB_DIR / ".build.lock", "w") as lock:      # several ranks may find the library stale at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB_PATH
        return _build_locked(nvcc, verbose)


def _build_locked(nvcc: str, verbose: bool) -> Path:
    want_hash = source_hash()
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    common = [nvcc, "-O3", "-std=c++17", *ARCH_FLAGS, "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v" if verbose else "-warn-spills", f"-I{REPO / 'include'}", f"-I{CSRC}"]

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src + ".o")
        cmd = [*common, "-c", str(CSRC / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose or "warning" in res.stderr.lower():
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    link = [nvcc, "-shared", *ARCH_FLAGS, "-cudart", "static", "-o", str(tmp), *map(str, objs)]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    HASH_PATH.write_text(want_hash + "\n")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
