"""GPU feature caches with the reference's API (gnnflow/cache/__init__.py:1-4).  LRU and FIFO are the policies on
the north-star path; LFU / GNNLab-static share the same gather kernel and are listed as next rows in DESIGN.md."""
from .cache import Cache  # noqa: F401
from .fifo_cache import FIFOCache  # noqa: F401
from .lru_cache import LRUCache  # noqa: F401
