"""First-in-first-out cache (reference gnnflow/cache/fifo_cache.py:9-161).  The ring pointers live on the device
so that an update needs no host synchronisation; `cache_node_pointer` / `cache_edge_pointer` read them back."""
import torch

from .._lib import check
from .cache import Cache


class FIFOCache(Cache):
    _policy = 1  # GF_CACHE_FIFO (include/gnnflow_b200.h)

    def _fifo_ptr(self, kind: str):
        return self._node_ptr if kind == "node" else self._edge_ptr

    def __init__(self, *args, **kwargs):
        super(FIFOCache, self).__init__(*args, **kwargs)
        self.name = 'fifo'
        self._node_ptr = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._edge_ptr = torch.zeros(1, dtype=torch.int64, device=self.device)

    @property
    def cache_node_pointer(self) -> int:
        return int(self._node_ptr.item())

    @cache_node_pointer.setter
    def cache_node_pointer(self, v: int):
        self._node_ptr.fill_(int(v))

    @property
    def cache_edge_pointer(self) -> int:
        return int(self._edge_ptr.item())

    @cache_edge_pointer.setter
    def cache_edge_pointer(self, v: int):
        self._edge_ptr.fill_(int(v))

    def init_cache(self, *args, **kwargs):
        """fifo_cache.py:57-68"""
        super(FIFOCache, self).init_cache(*args, **kwargs)
        if self.node_feats is not None:
            self.cache_node_pointer = self.node_capacity - 1
        if self.edge_feats is not None:
            self.cache_edge_pointer = self.edge_capacity - 1

    def reset(self):
        """fifo_cache.py:70-75"""
        if self.dim_edge_feat != 0:
            self.cache_edge_pointer = self.edge_capacity - 1

    def _update(self, kind, ids, hit_mask, ptr):
        feats = getattr(self, "%s_feats" % kind)
        st = self._state(kind)
        n = ids.shape[0]
        scratch = self._get_scratch(n, st.capacity, st.num_items)
        check(self._L.gf_cache_update_fifo(st, ids.data_ptr(), hit_mask.data_ptr(), n, feats.data_ptr(), ptr.data_ptr(),
                                           scratch.data_ptr(), scratch.numel(), self._stream()))

    def update_node_cache(self, ids, hit_mask):
        """fifo_cache.py:77-118"""
        self._update("node", ids, hit_mask, self._node_ptr)

    def update_edge_cache(self, ids, hit_mask):
        """fifo_cache.py:120-161"""
        self._update("edge", ids, hit_mask, self._edge_ptr)
