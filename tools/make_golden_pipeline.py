#!/usr/bin/env python
"""Generate tests/golden/pipeline_lshrs.npz by running the UNMODIFIED reference ``LSHRS`` end to end.

    python tools/make_golden_pipeline.py [--reference /root/reference]

(build container only, like tools/make_golden.py, whose hash / rerank fixtures pin the kernels; this one pins the
PIPELINE -- SURVEY section 8 rows a11 / a12: ``index`` -> ``get_top_k`` / ``get_above_p`` / ``query`` / ``delete`` --
to outputs of the reference itself.)  The reference keeps buckets in Redis; the script gives it the same dict-of-sets
double its own test-suite uses (reference tests/conftest.py:15-78 ``MockStorage``).

So that an implementation with a different float32 summation order must reproduce the ids EXACTLY, the data set is
filtered: no indexed or query vector has a projection closer to zero than 10x the parity margin (1e-4 relative), and
no query has two candidate scores closer than 1e-4 (rank order and the top-p cut are then decided by the data, not by
rounding).
"""

from __future__ import annotations

import argparse
import sys
import threading
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import _install_redis_stub  # noqa: E402

DIM, NUM_PERM, SEED = 48, 128, 7


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=str(Path(__file__).resolve().parents[1] / "tests" / "golden" / "pipeline_lshrs.npz"))
    args = ap.parse_args()
    _install_redis_stub()
    sys.path.insert(0, args.reference)
    from lshrs import LSHRS  # noqa: E402
    from lshrs.storage.redis import RedisStorage  # noqa: E402

    class DictStorage(RedisStorage):
        def __init__(self):
            self.data, self._lock = {}, threading.Lock()

        def batch_add(self, operations):
            with self._lock:
                for band_id, hash_val, index in operations:
                    self.data.setdefault((band_id, hash_val.hex()), set()).add(index)

        def get_bucket(self, band_id, hash_val):
            with self._lock:
                return set(self.data.get((band_id, hash_val.hex()), set()))

        def remove_indices(self, indices):
            with self._lock:
                gone = set(indices)
                for key in self.data:
                    self.data[key] -= gone

        def clear(self):
            self.data.clear()

        def close(self):
            pass

    rng = np.random.default_rng(2024)
    n0, nq0 = 3000, 120
    centers = rng.standard_normal((n0 // 6, DIM)).astype(np.float32)
    X = (np.repeat(centers, 6, axis=0) + 0.35 * rng.standard_normal((n0, DIM))).astype(np.float32)
    Q = (X[rng.integers(0, n0, nq0)] + 0.2 * rng.standard_normal((nq0, DIM))).astype(np.float32)

    store = {"X": X}
    lsh = LSHRS(dim=DIM, num_perm=NUM_PERM, storage=DictStorage(), seed=SEED,
                vector_fetch_fn=lambda ids: store["X"][np.asarray(ids, dtype=np.int64)])
    R = np.concatenate([np.asarray(p, dtype=np.float64) for p in lsh._hasher.projections])

    def clear_of_the_margin(V):
        V64 = V.astype(np.float64)
        cos = np.abs(V64 @ R.T) / (np.linalg.norm(V64, axis=1, keepdims=True) * np.linalg.norm(R, axis=1)[None, :])
        return (cos > 1e-4).all(axis=1)

    X = np.ascontiguousarray(X[clear_of_the_margin(X)])
    Q = np.ascontiguousarray(Q[clear_of_the_margin(Q)])
    store["X"] = X
    n = X.shape[0]
    lsh.index(list(range(n)), X)

    keep = []
    for i, q in enumerate(Q):                      # well-separated scores only (see the module docstring)
        full = lsh.query(q, top_k=None, top_p=1.0)
        sc = np.array([s for _, s in full], dtype=np.float64)
        if len(full) >= 3 and (np.abs(np.diff(sc)) > 1e-4).all():
            keep.append(i)
    Q = np.ascontiguousarray(Q[keep[:64]])

    def ragged(lists, dtype):
        offs = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in lists], out=offs[1:])
        flat = np.array([v for x in lists for v in x], dtype=dtype) if offs[-1] else np.empty(0, dtype=dtype)
        return offs, flat

    out = {"X": X, "Q": Q, "dim": DIM, "num_perm": NUM_PERM, "seed": SEED,
           "num_bands": lsh.stats()["num_bands"], "rows_per_band": lsh.stats()["rows_per_band"]}
    topk = [lsh.get_top_k(q, topk=10) for q in Q]
    out["topk_offs"], out["topk_ids"] = ragged(topk, np.int64)
    allc = [lsh.query(q, top_k=None) for q in Q]
    out["all_offs"], out["all_ids"] = ragged(allc, np.int64)
    above = [lsh.get_above_p(q, p=0.3) for q in Q]
    out["above_offs"], out["above_ids"] = ragged([[i for i, _ in r] for r in above], np.int64)
    _, out["above_scores"] = ragged([[s for _, s in r] for r in above], np.float32)
    both = [lsh.query(q, top_k=5, top_p=0.5) for q in Q]
    out["both_offs"], out["both_ids"] = ragged([[i for i, _ in r] for r in both], np.int64)
    _, out["both_scores"] = ragged([[s for _, s in r] for r in both], np.float32)
    gone = sorted({int(r[0]) for r in topk if r})[:40]
    lsh.delete(gone)
    out["deleted"] = np.array(gone, dtype=np.int64)
    after = [lsh.get_top_k(q, topk=10) for q in Q]
    out["after_delete_offs"], out["after_delete_ids"] = ragged(after, np.int64)
    np.savez_compressed(args.out, **out)
    print(f"{args.out}: {n} vectors, {len(Q)} queries, {out['num_bands']} x {out['rows_per_band']}, "
          f"{int(out['all_offs'][-1])} candidates, {int(out['above_offs'][-1])} reranked results")


if __name__ == "__main__":
    main()
