"""``lshrs.core.main`` IS ``lshrs_b200.core.main``: the same module object under the reference's name, so
that attribute patches made through this path (reference tests/test_redis_pooling.py:34 patches
``lshrs.core.main.RedisStorage``) reach the code that runs."""

import sys

import lshrs_b200.core.main as _impl

sys.modules[__name__] = _impl
