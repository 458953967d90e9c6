#!/usr/bin/env python
"""Tiny workload for `compute-sanitizer --tool memcheck python tools/sanitize_small.py` (one tool per call)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lshrs_b200 import LSHHasher, cosine_similarity, top_k_cosine, top_k_cosine_batch  # noqa: E402
from oracle import lshrs_oracle as oracle  # noqa: E402

rng = np.random.default_rng(0)
for nb, r, dim, n in [(16, 16, 64, 300), (16, 4, 128, 257), (5, 20, 100, 130), (3, 5, 7, 100), (32, 16, 32, 129)]:
    X = rng.standard_normal((n, dim)).astype(np.float32)
    for kernel in ("ffma", "tcgen05"):
        h = LSHHasher(nb, r, dim)
        h._ensure_handle()
        try:
            h.set_kernel(kernel)
        except Exception:
            continue
        got, flag = h.hash_batch_packed(X, return_zero_flag=True)
        rep = oracle.compare_packed(got, oracle.hash_batch_vectorized(h.projections, X),
                                    oracle.projection_margins(h.projections, X))
        assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0, rep
        h.close()
        print("hash ok", kernel, nb, r, dim, n, flush=True)
q = rng.standard_normal(64).astype(np.float32)
C = rng.standard_normal((301, 64)).astype(np.float32)
assert np.allclose(cosine_similarity(q, C), oracle.cosine_similarity(q, C), atol=1e-5)
assert [i for i, _ in top_k_cosine(q, C, k=7)] == [i for i, _ in oracle.top_k_cosine(q, C, k=7)]
offs = np.array([0, 5, 5, 301], dtype=np.int64)
pos, score, count = top_k_cosine_batch(rng.standard_normal((3, 64)).astype(np.float32), C, offs, None, k=4)
assert count.tolist() == [4, 0, 4]
q33 = rng.standard_normal(33).astype(np.float32)
C33 = rng.standard_normal((20000, 33)).astype(np.float32)   # chunked running top-k path
assert len(top_k_cosine(q33, C33, k=50)) == 50
print("rerank ok")
