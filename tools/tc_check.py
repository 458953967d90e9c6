#!/usr/bin/env python
"""Bring-up check of the tcgen05 hash kernel against the oracle, smallest shapes first.

    python tools/tc_check.py            # run every case, each in its own process with a timeout
    python tools/tc_check.py CASE_INDEX # run one case in this process

Not a test (tests/test_hash_gpu.py covers the kernel); a debugging aid that reports, per case,
how many bits differ from the oracle and where, and survives a hung kernel.
"""

from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

CASES = [
    # nb, r, dim, n
    (16, 8, 32, 128),      # one tile, one K chunk, 128 columns
    (16, 8, 64, 128),      # two K chunks
    (16, 8, 128, 300),     # three tiles (ragged)
    (32, 8, 128, 1000),    # 256 columns: nt = 2
    (16, 16, 768, 5000),   # BASELINE shape
    (16, 4, 128, 5000),    # nibble bands
    (16, 32, 1536, 2000),  # 512 columns: two passes
    (5, 20, 100, 2000),    # ragged
    (16, 16, 768, 200_000),
    # compact-column (repack) shapes: rows_per_band % 8 != 0
    (20, 5, 128, 3000),    # 100 bits -> N = 112
    (7, 3, 64, 1000),      # 21 bits  -> N = 32
    (2, 70, 64, 1000),     # 140 bits -> N = 144, 9 bytes per band
    (40, 7, 32, 2000),     # 280 bits -> two passes of 36 + 4 bands
    (200, 3, 16, 700),     # 600 bits -> three passes
    (3, 100, 256, 1500),   # 300 bits, 13 bytes per band, two bands per pass
    (16, 4, 128, 1_000_000),
    # element magnitudes spread over 2^-40 .. 2^40 inside every row: the FP16x3 split must send the rows
    # that leave its scaled range to the FP32 recomputation
    (16, 16, 768, 30_000, "wild"),
    (16, 32, 1536, 1_000, "wild"),
    (16, 16, 64, 3_000, "tiny"),   # whole rows around 1e-38 .. 1e-42 (scale not representable / denormals)
]


def run_case(i: int) -> int:
    from lshrs_b200 import LSHHasher
    from oracle import lshrs_oracle as oracle

    nb, r, dim, n = CASES[i][:4]
    dist = CASES[i][4] if len(CASES[i]) > 4 else "gauss"
    rng = np.random.default_rng(i)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    if dist == "wild":
        X = (X * np.exp2(rng.integers(-40, 41, size=X.shape))).astype(np.float32)
        X[::7] = rng.standard_normal((len(X[::7]), dim)).astype(np.float32)   # ordinary rows in between
    elif dist == "tiny":
        X = (X * np.float32(1e-38) * np.exp2(rng.integers(-14, 1, size=(n, 1)))).astype(np.float32)
    h = LSHHasher(nb, r, dim, seed=42)
    h._ensure_handle()
    h.set_kernel(os.environ.get("TC_KERNEL", "tcgen05"))
    t0 = time.perf_counter()
    got, flag = h.hash_batch_packed(X, return_zero_flag=True)
    dt = time.perf_counter() - t0
    want = oracle.hash_batch_vectorized(h.projections, X)
    rep = oracle.compare_packed(got, want, oracle.projection_margins(h.projections, X), 1e-5)
    want_flag = (np.abs(X) <= 1e-8).all(axis=1)   # LSHRS._prepare_vector's np.allclose(arr, 0, atol=1e-8)
    bad = rep["flips_outside_margin"] or rep["nonzero_pad_bits"] or (flag.astype(bool) != want_flag).any()
    print(f"case {i} {CASES[i]} kernel={h.last_kernel} {dt * 1e3:.1f} ms -> {'FAIL' if bad else 'ok'} {rep}", flush=True)
    if bad:
        diff = np.unpackbits(got ^ want, axis=2, bitorder="little")
        rows = np.nonzero(diff.any(axis=(1, 2)))[0]
        cols = np.nonzero(diff.reshape(n, -1).any(axis=0))[0]
        print(f"   differing rows: {len(rows)} first {rows[:8].tolist()} last {rows[-4:].tolist()}")
        print(f"   differing bit columns: {len(cols)} first {cols[:16].tolist()}")
        print(f"   got[0]  = {got[0].reshape(-1)[:16].tolist()}\n   want[0] = {want[0].reshape(-1)[:16].tolist()}")
    return 1 if bad else 0


def main() -> None:
    if len(sys.argv) > 1:
        sys.exit(run_case(int(sys.argv[1])))
    failed = 0
    for i in range(len(CASES)):
        try:
            res = subprocess.run([sys.executable, __file__, str(i)], timeout=120)
            failed += res.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"case {i} {CASES[i]} TIMED OUT (kernel hang?)", flush=True)
            failed += 1
            break
    print(f"tc_check: {failed} failing case(s)")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
