"""``l2_norm`` -- unit-length copy of one vector (mirrors reference lshrs/utils/norm.py:4-61).

This helper is kept for API compatibility only.  On the rerank hot path the
per-candidate normalisation the reference does by calling this function once
per candidate (reference lshrs/utils/similarity.py:85) is fused into the CUDA
rerank kernel (``csrc/rerank.cu``); nothing on that path calls this function.
"""

from __future__ import annotations

import numpy as np

__all__ = ["l2_norm"]


def l2_norm(vector) -> np.ndarray:
    """Flatten to float32 and divide by the Euclidean norm; zero vectors raise ``ValueError``."""
    flat = np.asarray(vector, dtype=np.float32).reshape(-1)
    length = np.linalg.norm(flat)
    if length == 0:
        raise ValueError("Cannot normalize zero vector")
    return flat / length
