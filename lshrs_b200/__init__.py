"""lshrs_b200 -- the compute hot path of lshrs (mxngjxa/lshrs) on NVIDIA B200.

Same Python surface as the reference for the path it replaces:

    from lshrs_b200 import LSHRS, LSHHasher, HashSignatures
    from lshrs_b200 import cosine_similarity, top_k_cosine, l2_norm

``LSHHasher`` (banded random-projection signatures) and ``top_k_cosine`` /
``cosine_similarity`` (candidate rerank) run in hand-written sm_100a CUDA
kernels behind a C ABI (``include/lshx.h``, ``liblshx.so``).  There is no CPU
fallback: without the library or an sm_100 GPU the first hash / rerank call
raises ``LshxUnavailable``.  Importing the package itself needs neither.
"""

from lshrs_b200._config.config import HashSignatures
from lshrs_b200._native import LshxError, LshxUnavailable
from lshrs_b200.core.main import LSHRS, lshrs
from lshrs_b200.hash.lsh import LSHHasher
from lshrs_b200.storage.device import DeviceBucketStorage, DeviceIndex
from lshrs_b200.storage.memory import BucketOperation, InMemoryStorage, bucket_key
from lshrs_b200.utils.norm import l2_norm, l2_norm_batch
from lshrs_b200.utils.similarity import Reranker, cosine_similarity, top_k_cosine, top_k_cosine_batch

__version__ = "0.1.0"

__all__ = [
    "LSHRS",
    "lshrs",
    "LSHHasher",
    "HashSignatures",
    "cosine_similarity",
    "top_k_cosine",
    "top_k_cosine_batch",
    "Reranker",
    "l2_norm",
    "l2_norm_batch",
    "InMemoryStorage",
    "DeviceBucketStorage",
    "DeviceIndex",
    "BucketOperation",
    "bucket_key",
    "LshxError",
    "LshxUnavailable",
    "__version__",
]
