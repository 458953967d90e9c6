"""``lshrs`` -- import-path shim over ``lshrs_b200`` (see ``lshrs_b200/compat/__init__.py``).

Hot-path modules (``hash.lsh``, ``utils.similarity``, ``utils.norm``, ``utils.br``, ``_config.config``,
``core.main``) are the B200 implementation; ``storage`` and ``io`` are looked up in the reference package
found further down ``sys.path`` (they stay host code, unchanged -- reference lshrs/storage/redis.py,
lshrs/io/*.py).
"""

import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_here]
for _entry in _sys.path:
    _cand = _os.path.join(_entry or ".", "lshrs")
    if (_os.path.isfile(_os.path.join(_cand, "storage", "redis.py"))
            and _os.path.realpath(_cand) != _os.path.realpath(_here)):
        __path__.append(_cand)      # the reference's lshrs/: serves lshrs.storage and lshrs.io only
        break

from lshrs.core.main import LSHRS, lshrs  # noqa: E402

__version__ = "0.1.1b2+b200"
__all__ = ["LSHRS", "lshrs"]
