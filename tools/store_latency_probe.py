"""Per-call latency of the single-query path on a DeviceBucketStorage (768 / 256 bits, 50 000 indexed vectors)."""
import ctypes
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lshrs_b200 import LSHRS, DeviceBucketStorage, _native  # noqa: E402

dim, n = 768, 50_000
rng = np.random.default_rng(11)
centers = rng.standard_normal((n // 8, dim)).astype(np.float32)
X = (np.repeat(centers, 8, axis=0) + 0.15 * rng.standard_normal((n, dim))).astype(np.float32)
lsh = LSHRS(dim=dim, num_perm=256, storage=DeviceBucketStorage(), device=0)
lsh.index(np.arange(n), X)
Q = (X[rng.integers(0, n, 2000)] + 0.05 * rng.standard_normal((2000, dim))).astype(np.float32)


def per_call(fn, reps=2000):
    fn(0)
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    return 1e6 * (time.perf_counter() - t0) / reps


ix, h = lsh._dindex, lsh._hasher
print(f"LSHRS.get_top_k            {per_call(lambda i: lsh.get_top_k(Q[i], topk=10)):7.1f} us")
print(f"LSHRS._prepare_vector      {per_call(lambda i: lsh._prepare_vector(Q[i])):7.1f} us")
print(f"DeviceIndex.query_vectors  {per_call(lambda i: ix.query_vectors(h, Q[i:i + 1], 10)):7.1f} us")
print(f"LSHHasher.hash_vector      {per_call(lambda i: h.hash_vector(Q[i])):7.1f} us")
# the bare C call with preallocated outputs
lib = _native.lib()
ids = np.empty((1, 10), np.int64); coll = np.empty((1, 10), np.int32); cnt = np.zeros(1, np.int32); z = np.zeros(1, np.uint8)
hh, ih = h._ensure_handle(), ix._handle
args = (ids.ctypes.data, coll.ctypes.data, cnt.ctypes.data, z.ctypes.data)
print(f"lshx_index_query_vectors   {per_call(lambda i: lib.lshx_index_query_vectors(ih, hh, Q[i].ctypes.data, 1, 10, *args)):7.1f} us")
out = np.empty((1, 32), np.uint8)
print(f"lshx_hash_batch (1 row)    {per_call(lambda i: lib.lshx_hash_batch(hh, Q[i].ctypes.data, 1, 0, out.ctypes.data, 0, None, None)):7.1f} us")
print(f"ctypes no-op (abi_version) {per_call(lambda i: lib.lshx_abi_version()):7.1f} us")
# phase stamps of the fused latency kernel (lshx_index_debug_timeline), median over 200 calls
stamps = np.zeros(9, np.uint64)
lib.lshx_index_debug_timeline(ih, 1, None)
rows, wall = [], []
for i in range(300):
    t0 = time.perf_counter_ns()
    lib.lshx_index_query_vectors(ih, hh, Q[i].ctypes.data, 1, 10, *args)
    wall.append(time.perf_counter_ns() - t0)
    lib.lshx_index_debug_timeline(ih, 1, stamps.ctypes.data)
    rows.append(stamps.astype(np.int64).copy())
rows = np.array(rows[100:])
names = ["hash (last CTA)", "ticket", "band searches", "gather", "first sort", "count + second sort", "store"]
d = np.diff(rows[:, :8], axis=1)
print("fused kernel phases of the last CTA, median ns:", {n: int(np.median(d[:, i])) for i, n in enumerate(names)},
      "total", int(np.median(rows[:, 7] - rows[:, 0])), "raw entries", int(np.median(rows[:, 8])),
      "| C call wall ns (stamps on)", int(np.median(wall[100:])))
lib.lshx_index_debug_timeline(ih, 0, None)
