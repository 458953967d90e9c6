#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (mxngjxa/lshrs).

Run in the build container only (``/root/reference`` does not exist on the GPU
box):

    python tools/make_golden.py [--reference /root/reference]

The reference imports ``redis`` at package-import time
(lshrs/storage/redis.py:33) and the image has no redis-py, so a ten-line stub
module is installed in ``sys.modules`` first; nothing on the hash / rerank path
touches it.  Every array written here is an output of reference code
(``LSHHasher`` / ``cosine_similarity`` / ``top_k_cosine`` / ``l2_norm`` /
``RedisStorage.bucket_key`` / ``get_optimal_config``), together with the inputs
that produced it, so the fixtures are self-contained.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np


def _install_redis_stub() -> None:
    redis = types.ModuleType("redis")

    class ConnectionPool:
        def __init__(self, **kw):
            self.kw = kw

        def disconnect(self):
            pass

    class Redis:
        def __init__(self, connection_pool=None, **kw):
            self.connection_pool = connection_pool

    redis.ConnectionPool = ConnectionPool
    redis.Redis = Redis
    sys.modules["redis"] = redis


def _packed(sigs, nb, bpb):
    out = np.zeros((len(sigs), nb, bpb), dtype=np.uint8)
    for i, s in enumerate(sigs):
        for b, raw in enumerate(s):
            assert isinstance(raw, bytes) and len(raw) == bpb
            out[i, b] = np.frombuffer(raw, dtype=np.uint8)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=str(Path(__file__).resolve().parents[1] / "tests" / "golden"))
    args = ap.parse_args()

    _install_redis_stub()
    sys.path.insert(0, args.reference)
    from lshrs.hash.lsh import LSHHasher  # noqa: E402
    from lshrs.storage.redis import RedisStorage  # noqa: E402
    from lshrs.utils.br import get_optimal_config  # noqa: E402
    from lshrs.utils.norm import l2_norm  # noqa: E402
    from lshrs.utils.similarity import cosine_similarity, top_k_cosine  # noqa: E402

    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    manifest: dict = {"numpy": np.__version__, "cases": {}}

    # ------------------------------------------------------------------ hashing
    def sift_like(rng, n, dim):
        return np.minimum(255.0, np.floor(np.abs(rng.standard_normal((n, dim))) * 40.0)).astype(np.float32)

    hash_cases = [
        # name, bands, rows, dim, seed, input builder
        ("kat1_3x5x4", 3, 5, 4, 123, lambda: np.arange(4, dtype=np.float32)[None, :]),
        ("kat2_4x6x3", 4, 6, 3, 42,
         lambda: np.array([[1, 0, -1], [-1, 1, 0], [0.5, 0.5, 0.5]], dtype=np.float32)),
        ("cfg_768_16x16", 16, 16, 768, 42,
         lambda: np.random.default_rng(0).standard_normal((64, 768)).astype(np.float32)),
        ("cfg_1536_16x32", 16, 32, 1536, 42,
         lambda: np.random.default_rng(0).standard_normal((24, 1536)).astype(np.float32)),
        ("cfg_128_16x4_gauss", 16, 4, 128, 42,
         lambda: np.random.default_rng(0).standard_normal((96, 128)).astype(np.float32)),
        ("cfg_128_16x4_sift", 16, 4, 128, 42, lambda: sift_like(np.random.default_rng(5), 96, 128)),
        ("cfg_128_8x16", 8, 16, 128, 7,
         lambda: np.random.default_rng(11).standard_normal((96, 128)).astype(np.float32)),
        ("ragged_5x20x100", 5, 20, 100, 3,
         lambda: np.random.default_rng(12).standard_normal((40, 100)).astype(np.float32)),
        ("ragged_7x3x33", 7, 3, 33, 9,
         lambda: np.random.default_rng(13).standard_normal((50, 33)).astype(np.float32)),
        ("ragged_2x70x17", 2, 70, 17, 1,
         lambda: np.random.default_rng(14).standard_normal((30, 17)).astype(np.float32)),
        ("tiny_1x1x1", 1, 1, 1, 42, lambda: np.array([[1.0], [-1.0], [0.0]], dtype=np.float32)),
        ("edge_zero_nan_32_4x4", 4, 4, 32, 42,
         lambda: np.stack([np.zeros(32, np.float32), np.full(32, np.nan, np.float32),
                           np.full(32, 1e-30, np.float32), -np.ones(32, np.float32),
                           np.random.default_rng(15).standard_normal(32).astype(np.float32) * 1e20])),
        # typed inputs: the reference casts them with np.asarray(vectors, float32) (lsh.py:162); 4096 rows so
        # that lshrs_b200 takes lshx_hash_batch_typed (raw rows over PCIe, cast on the device)
        ("typed_u8_4096x32_4x4", 4, 4, 32, 42,
         lambda: sift_like(np.random.default_rng(21), 4096, 32).astype(np.uint8)),
        ("typed_f16_4096x32_4x4", 4, 4, 32, 42,
         lambda: np.random.default_rng(22).standard_normal((4096, 32)).astype(np.float16)),
        ("typed_i8_4096x32_4x4", 4, 4, 32, 42,
         lambda: np.random.default_rng(23).integers(-128, 128, size=(4096, 32), dtype=np.int8)),
    ]
    for name, nb, r, dim, seed, build in hash_cases:
        h = LSHHasher(num_bands=nb, rows_per_band=r, dim=dim, seed=seed)
        X = build()
        sigs_batch = h.hash_batch(X)
        sigs_single = [h.hash_vector(x) for x in X]
        assert [s.as_tuple() for s in sigs_batch] == [s.as_tuple() for s in sigs_single]
        bpb = (r + 7) // 8
        packed = _packed(sigs_batch, nb, bpb)
        R = np.concatenate(h.projections, axis=0)
        assert R.dtype == np.float32 and R.shape == (nb * r, dim)
        store_R = R if R.size <= 4096 else np.zeros((0, dim), np.float32)
        np.savez_compressed(
            out / f"hash_{name}.npz",
            num_bands=nb, rows_per_band=r, dim=dim, seed=seed,
            X=X, signatures=packed, R_small=store_R,
            R_sha256=np.frombuffer(hashlib.sha256(R.tobytes()).digest(), dtype=np.uint8),
            R_corner=np.array([R[0, 0], R[0, -1], R[-1, 0], R[-1, -1]], dtype=np.float32),
        )
        manifest["cases"][f"hash_{name}"] = {
            "n": int(X.shape[0]), "first_hex": [b.hex() for b in sigs_batch[0].as_tuple()],
        }

    # bucket keys (lshrs/storage/redis.py:225) for the first vector of the 768 case
    storage = RedisStorage.__new__(RedisStorage)
    storage.prefix = "lsh"
    h = LSHHasher(16, 16, 768, seed=42)
    X = np.random.default_rng(0).standard_normal((2, 768)).astype(np.float32)
    keys = [[storage.bucket_key(b, hv) for b, hv in enumerate(h.hash_vector(x))] for x in X]
    manifest["bucket_keys_768"] = keys
    manifest["bucket_key_example"] = storage.bucket_key(5, b"\xab\xcd")

    # auto-config shapes the kernel must support (lshrs/utils/br.py:325)
    manifest["optimal_config"] = {
        str(n): list(map(int, get_optimal_config(n, 0.5))) for n in (64, 128, 256, 512, 1024)
    }
    manifest["optimal_config_4096_0.9"] = list(map(int, get_optimal_config(4096, 0.9)))
    manifest["optimal_config_grid"] = {
        f"{n}@{t}": list(map(int, get_optimal_config(n, t)))
        for n in (1, 2, 3, 7, 16, 17, 32, 48, 64, 96, 100, 120, 128, 200, 256, 384, 512, 768, 1000, 1024, 2048,
                  4096, 8192, 16384, 65536)
        for t in (0.1, 0.3, 0.42, 0.5, 0.55, 0.6, 0.7, 0.8, 0.87, 0.9, 0.95, 0.99)
    }

    # ------------------------------------------------------------------ rerank
    rer_cases = [
        ("ref_test_cos", np.array([1.0, 0.0, 0.0], np.float32),
         np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [1, 1, 0]], np.float32), [1, 2, 4, 10]),
        ("ref_test_topk", np.array([1.0, 0.0], np.float32),
         np.array([[1.0, 0.1], [0.0, 1.0], [1.0, 0.0], [-1.0, 0.0], [0.9, 0.2]], np.float32), [3, 5, 10]),
        ("gauss_300x768", np.random.default_rng(2).standard_normal(768).astype(np.float32),
         np.random.default_rng(1).standard_normal((300, 768)).astype(np.float32), [1, 10, 60, 300, 1000]),
        ("gauss_2000x128", np.random.default_rng(21).standard_normal(128).astype(np.float32),
         np.random.default_rng(22).standard_normal((2000, 128)).astype(np.float32), [10, 400, 2000]),
        ("scaled_257x33", (np.random.default_rng(23).standard_normal(33) * 1e3).astype(np.float32),
         (np.random.default_rng(24).standard_normal((257, 33))
          * np.logspace(-6, 6, 257)[:, None]).astype(np.float32), [7, 257]),
        ("single_1x5", np.arange(1, 6, dtype=np.float32),
         np.array([[5, 4, 3, 2, 1]], np.float32), [1, 3]),
    ]
    for name, q, C, ks in rer_cases:
        sims = cosine_similarity(q, C)
        assert sims.dtype == np.float32
        payload = {"query": q, "candidates": C, "similarities": sims,
                   "normalized_query": l2_norm(q), "ks": np.array(ks, np.int64)}
        for k in ks:
            res = top_k_cosine(q, C, k=k)
            payload[f"top{k}_idx"] = np.array([i for i, _ in res], np.int64)
            payload[f"top{k}_score"] = np.array([s for _, s in res], np.float64)
        np.savez_compressed(out / f"rerank_{name}.npz", **payload)
        manifest["cases"][f"rerank_{name}"] = {"n": int(C.shape[0]), "ks": ks}

    (out / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print(f"wrote {len(manifest['cases'])} cases to {out}")


if __name__ == "__main__":
    main()
