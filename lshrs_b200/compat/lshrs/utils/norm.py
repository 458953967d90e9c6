"""``lshrs.utils.norm`` -> the B200 ``l2_norm`` (replaces reference lshrs/utils/norm.py:4-61)."""

from lshrs_b200.utils.norm import l2_norm

__all__ = ["l2_norm"]
