"""GPU feature caches with the reference's API (gnnflow/cache/__init__.py:1-4): LRU / FIFO / LFU policies and the
GNNLab static cache over one gather kernel (gf_cache_gather)."""
from .cache import Cache  # noqa: F401
from .fifo_cache import FIFOCache  # noqa: F401
from .gnnlab_static_cache import GNNLabStaticCache  # noqa: F401
from .lfu_cache import LFUCache  # noqa: F401
from .lru_cache import LRUCache  # noqa: F401
