"""CPU oracle for the lshrs hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a plain-numpy restatement of the two hot-path algorithms of the
reference (mxngjxa/lshrs 0.1.1b2, pure Python):

  * banded random-projection signatures   reference lshrs/hash/lsh.py
  * cosine candidate reranking            reference lshrs/utils/similarity.py,
                                          lshrs/utils/norm.py

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product package (``lshrs_b200``) never imports it and
has no CPU fallback.

Pinning: the oracle is pinned against outputs of the reference itself.
``tools/make_golden.py`` imports the unmodified reference from
``/root/reference`` (with a stub ``redis`` module) and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` asserts that this
restatement reproduces every one of those vectors byte for byte (signatures)
or to <= 1 ulp-level tolerance (cosine scores), plus the reference's own
known-answer tests (tests/test_lshrs.py:115-153 of the reference).

The arithmetic lives in numpy (reference pyproject.toml:35 ``numpy>=1.24``;
uv.lock pins 2.2.6 / 2.3.4; this image has 2.3.x): PCG64 + ziggurat
``standard_normal`` for the projection matrices, OpenBLAS sgemv for the
projections, ``np.packbits(bitorder="little")`` for the band bytes.
"""

from __future__ import annotations

import math
from collections.abc import Sequence

import numpy as np

__all__ = [
    "make_projections",
    "project_and_pack",
    "hash_vector",
    "hash_batch",
    "hash_batch_packed",
    "hash_batch_vectorized",
    "band_bytes",
    "projection_margins",
    "compare_packed",
    "bucket_key",
    "l2_norm",
    "cosine_similarity",
    "top_k_cosine",
    "top_p_limit",
    "is_zero_vector",
]


# --------------------------------------------------------------------------
# Hashing (reference lshrs/hash/lsh.py)
# --------------------------------------------------------------------------

def make_projections(num_bands: int, rows_per_band: int, dim: int, seed: int = 42) -> list[np.ndarray]:
    """Projection matrices, one (rows_per_band, dim) float32 array per band.

    Follows reference lshrs/hash/lsh.py:78-94: validation of the three sizes,
    then ``num_bands`` sequential float64 ``standard_normal`` draws from one
    ``default_rng(seed)`` each cast to float32.
    """
    if num_bands <= 0:
        raise ValueError("num_bands must be > 0")
    if rows_per_band <= 0:
        raise ValueError("rows_per_band must be > 0")
    if dim <= 0:
        raise ValueError("dim must be > 0")
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((rows_per_band, dim)).astype(np.float32) for _ in range(num_bands)]


def band_bytes(rows_per_band: int) -> int:
    """Length of one band's packed signature: ceil(rows_per_band / 8) (np.packbits)."""
    return (rows_per_band + 7) // 8


def project_and_pack(projection: np.ndarray, vector: np.ndarray) -> bytes:
    """One band of one vector.  Reference lshrs/hash/lsh.py:200-211.

    fp32 mat-vec, strict ``> 0`` (zero and NaN give bit 0), little-endian bit
    packing (bit i of byte j is row 8j+i, high bits of the last byte zero).
    """
    projected = projection @ vector
    binary = projected > 0
    return np.packbits(binary.astype(np.uint8), bitorder="little").tobytes()


def _validate_vector(vector, dim: int) -> np.ndarray:
    """Reference lshrs/hash/lsh.py:241-247."""
    vec = np.asarray(vector, dtype=np.float32).reshape(-1)
    if vec.ndim != 1 or vec.shape[0] != dim:
        raise ValueError(f"Expected vector of dimension {dim}, received {vec.shape}")
    return vec


def hash_vector(projections: Sequence[np.ndarray], vector) -> tuple[bytes, ...]:
    """All bands of one vector.  Reference lshrs/hash/lsh.py:129-134."""
    dim = projections[0].shape[1]
    vec = _validate_vector(vector, dim)
    return tuple(project_and_pack(p, vec) for p in projections)


def hash_batch(projections: Sequence[np.ndarray], vectors) -> list[tuple[bytes, ...]]:
    """Reference lshrs/hash/lsh.py:161-169: a Python loop of hash_vector."""
    dim = projections[0].shape[1]
    arr = np.asarray(vectors, dtype=np.float32)
    if arr.ndim != 2:
        raise ValueError("Batch input must be a 2D array")
    if arr.shape[1] != dim:
        raise ValueError(f"Expected vectors of dimension {dim}, received {arr.shape[1]}")
    return [hash_vector(projections, vec) for vec in arr]


def hash_batch_packed(projections: Sequence[np.ndarray], vectors) -> np.ndarray:
    """hash_batch laid out as uint8[n, num_bands, ceil(r/8)] (the C-ABI layout).

    Same arithmetic as the reference (one sgemv per vector per band); only the
    container differs.
    """
    sigs = hash_batch(projections, vectors)
    nb = len(projections)
    bpb = band_bytes(projections[0].shape[0])
    out = np.zeros((len(sigs), nb, bpb), dtype=np.uint8)
    for i, bands in enumerate(sigs):
        for b, raw in enumerate(bands):
            out[i, b] = np.frombuffer(raw, dtype=np.uint8)
    return out


def hash_batch_vectorized(projections: Sequence[np.ndarray], vectors) -> np.ndarray:
    """NOT reference code: one fp32 sgemm + packbits over all BLAS threads.

    The honest "well-vectorised numpy" figure that BASELINE.md section 2 row 7 quotes;
    same layout as :func:`hash_batch_packed`.  fp32 with another summation
    order, so it can differ from the reference only inside the near-zero margin.
    """
    arr = np.ascontiguousarray(vectors, dtype=np.float32)
    nb = len(projections)
    r = projections[0].shape[0]
    R = np.concatenate(list(projections), axis=0)  # (num_perm, dim)
    bits = (arr @ R.T) > 0  # (n, num_perm)
    bits = bits.reshape(arr.shape[0], nb, r)
    return np.packbits(bits, axis=2, bitorder="little")


def projection_margins(projections: Sequence[np.ndarray], vectors) -> np.ndarray:
    """fp64 relative margins |x.r| / (||x|| ||r||), shape (n, num_bands, rows_per_band).

    BASELINE.json north_star: band keys must be bit-exact wherever this value
    exceeds 1e-5; projections below it are exempt (counted and reported).
    Rows with ||x|| == 0 get margin 0 (every bit exempt; both sides give 0).
    """
    X = np.asarray(vectors, dtype=np.float64)
    R = np.concatenate([np.asarray(p, dtype=np.float64) for p in projections], axis=0)
    dots = np.abs(X @ R.T)
    xn = np.linalg.norm(X, axis=1)[:, None]
    rn = np.linalg.norm(R, axis=1)[None, :]
    denom = xn * rn
    with np.errstate(divide="ignore", invalid="ignore"):
        m = np.where(denom > 0, dots / denom, 0.0)
    return m.reshape(X.shape[0], len(projections), projections[0].shape[0])


def compare_packed(got: np.ndarray, want: np.ndarray, margins: np.ndarray, rel_margin: float = 1e-5) -> dict:
    """Bit-level comparison of two packed signature arrays.

    Returns counts of differing bits outside / inside the exempt margin, the
    number of differing band keys, and ``max_flipped_margin``: the largest relative
    margin |x.r| / (||x|| ||r||) at which a bit actually differs (0.0 when none
    does) -- the empirical headroom of an arithmetic against the 1e-5 parity band.
    """
    assert got.shape == want.shape, (got.shape, want.shape)
    n, nb, bpb = want.shape
    r = margins.shape[2]
    gbits = np.unpackbits(got, axis=2, bitorder="little")[:, :, :r]
    wbits = np.unpackbits(want, axis=2, bitorder="little")[:, :, :r]
    pad_g = np.unpackbits(got, axis=2, bitorder="little")[:, :, r:]
    diff = gbits != wbits
    exempt = margins <= rel_margin
    return {
        "bits": int(diff.size),
        "flips_outside_margin": int(np.count_nonzero(diff & ~exempt)),
        "flips_inside_margin": int(np.count_nonzero(diff & exempt)),
        "bits_inside_margin": int(np.count_nonzero(exempt)),
        "band_keys": int(n * nb),
        "band_keys_differing": int(np.count_nonzero(diff.any(axis=2))),
        "nonzero_pad_bits": int(np.count_nonzero(pad_g)),
        "max_flipped_margin": float(margins[diff].max()) if diff.any() else 0.0,
    }


def bucket_key(prefix: str, band_id: int, hash_val: bytes) -> str:
    """Reference lshrs/storage/redis.py:225 -- the only consumer of the band bytes."""
    return f"{prefix}:{band_id}:bucket:{hash_val.hex()}"


def is_zero_vector(vector) -> bool:
    """Reference lshrs/core/main.py:1083 (np.allclose(arr, 0.0, atol=1e-8); NaN passes)."""
    arr = np.asarray(vector, dtype=np.float32).reshape(-1)
    return bool(np.allclose(arr, 0.0, atol=1e-8))


# --------------------------------------------------------------------------
# Rerank (reference lshrs/utils/norm.py, lshrs/utils/similarity.py)
# --------------------------------------------------------------------------

def l2_norm(vector) -> np.ndarray:
    """Reference lshrs/utils/norm.py:48-61."""
    vec = np.asarray(vector, dtype=np.float32).reshape(-1)
    norm = np.linalg.norm(vec)
    if norm == 0:
        raise ValueError("Cannot normalize zero vector")
    return vec / norm


def cosine_similarity(query, candidates) -> np.ndarray:
    """Reference lshrs/utils/similarity.py:80-90 (normalise each candidate, stack, sgemv)."""
    normalized_query = l2_norm(query)
    normalized_candidates = np.stack([l2_norm(vec) for vec in candidates])
    return normalized_candidates @ normalized_query


def top_k_cosine(query, candidates, *, k: int) -> list[tuple[int, float]]:
    """Reference lshrs/utils/similarity.py:157-183."""
    if k <= 0:
        raise ValueError("k must be > 0")
    similarities = cosine_similarity(query, candidates)
    if len(similarities) == 0:
        return []
    top_indices = np.argpartition(-similarities, kth=min(k, len(similarities) - 1))[:k]
    sorted_indices = top_indices[np.argsort(-similarities[top_indices])]
    return [(int(idx), float(similarities[idx])) for idx in sorted_indices]


def top_p_limit(n: int, top_p: float, top_k: int | None = None) -> int:
    """Reference lshrs/core/main.py:650-656: rank-fraction cut used by get_above_p."""
    limit = max(1, math.ceil(n * top_p))
    if top_k is not None:
        limit = min(limit, top_k)
    return limit
