from .main import LSHRS, lshrs

__all__ = ["LSHRS", "lshrs"]
