"""``lshrs.utils.similarity`` -> the B200 rerank (replaces reference lshrs/utils/similarity.py:26-183)."""

from lshrs_b200.utils.similarity import cosine_similarity, top_k_cosine

__all__ = ["cosine_similarity", "top_k_cosine"]
