"""hash_batch_packed throughput for mid-size HOST batches (what query_batch / small index() calls send), dim 768."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lshrs_b200 import LSHHasher  # noqa: E402

h = LSHHasher(16, 16, 768, seed=42, device=0)
rng = np.random.default_rng(0)
X = rng.standard_normal((131072, 768)).astype(np.float32)
h.hash_batch_packed(X)
for n in (33, 64, 128, 256, 512, 1024, 2048, 4095, 4096, 8192, 16384, 32768, 65536, 131072):
    reps = max(3, min(200, 2_000_000 // n))
    h.hash_batch_packed(X[:n])
    t0 = time.perf_counter()
    for _ in range(reps):
        h.hash_batch_packed(X[:n])
    dt = (time.perf_counter() - t0) / reps
    print(f"n={n:7d}: {1e3 * dt:8.3f} ms  {n / dt / 1e6:7.2f} M vectors/s  {n * 3072 / dt / 1e9:6.1f} GB/s", flush=True)
