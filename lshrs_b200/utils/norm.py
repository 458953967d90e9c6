"""``l2_norm`` -- unit-length copy of a vector (drop-in for reference lshrs/utils/norm.py:4-61).

Runs on the GPU like everything else on the path (``lshx_l2_normalize``: one warp per row, fp32
sum of squares, sqrt, per-element divide -- the reference's order of operations).  On the rerank
hot path the per-candidate normalisation the reference does by calling this function once per
candidate (reference lshrs/utils/similarity.py:85) is fused into the rerank kernel instead; nothing
there calls this helper.
"""

from __future__ import annotations

import numpy as np

from lshrs_b200 import _native

__all__ = ["l2_norm", "l2_norm_batch"]


def l2_norm_batch(vectors) -> np.ndarray:
    """Normalise every row of a 2-D array; raises ``ValueError`` if any row has zero norm."""
    arr = np.ascontiguousarray(vectors, dtype=np.float32)
    if arr.ndim != 2:
        raise ValueError("Batch input must be a 2D array")
    n, dim = arr.shape
    if n == 0 or dim == 0:
        return arr.copy()
    from lshrs_b200.utils.similarity import _get_reranker

    rer = _get_reranker(dim)
    out = np.empty_like(arr)
    zero = np.zeros(n, dtype=np.int32)
    _native.check(_native.lib().lshx_l2_normalize(rer._handle, arr.ctypes.data, n, out.ctypes.data,
                                                  zero.ctypes.data, 0, None))
    if zero.any():
        raise ValueError("Cannot normalize zero vector")
    return out


def l2_norm(vector) -> np.ndarray:
    """Flatten to float32 and divide by the Euclidean norm; zero vectors raise ``ValueError``."""
    flat = np.asarray(vector, dtype=np.float32).reshape(-1)
    if flat.shape[0] == 0:
        raise ValueError("Cannot normalize zero vector")
    return l2_norm_batch(flat.reshape(1, -1))[0]
