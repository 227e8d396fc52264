"""Least-frequently-used cache (reference gnnflow/cache/lfu_cache.py:9-210)."""
import torch

from .._lib import check
from .cache import Cache


class LFUCache(Cache):
    _policy = 2  # GF_CACHE_LFU (include/gnnflow_b200.h)

    def __init__(self, *args, **kwargs):
        super(LFUCache, self).__init__(*args, **kwargs)
        self.name = 'lfu'
        if self.dim_node_feat != 0:
            self.cache_node_count = torch.zeros(self.node_capacity, dtype=torch.int32, device=self.device)
        if self.dim_edge_feat != 0:
            self.cache_edge_count = torch.zeros(self.edge_capacity, dtype=torch.int32, device=self.device)

    def get_mem_size(self) -> int:
        mem_size = super(LFUCache, self).get_mem_size()
        if self.dim_node_feat != 0:
            mem_size += self.cache_node_count.element_size() * self.cache_node_count.nelement()
        if self.dim_edge_feat != 0:
            mem_size += self.cache_edge_count.element_size() * self.cache_edge_count.nelement()
        return mem_size

    def init_cache(self, *args, **kwargs):
        """lfu_cache.py:72-84: the rows loaded at init start with a use count of 1"""
        super(LFUCache, self).init_cache(*args, **kwargs)
        if self.dim_node_feat != 0:
            self.cache_node_count[self.cache_index_to_node_id] += 1
        if self.dim_edge_feat != 0:
            self.cache_edge_count[self.cache_index_to_edge_id] += 1

    def reset(self):
        """NB: only the edge cache is reset (lfu_cache.py:86-118)"""
        if self.edge_feats is not None and self.dim_edge_feat != 0:
            ids = torch.arange(self.edge_capacity, dtype=torch.int64, device=self.device)
            self.cache_edge_buffer[ids] = self.edge_feats[:self.edge_capacity].to(self.device, non_blocking=True)
            self.cache_edge_flag[ids] = True
            self.cache_index_to_edge_id = ids
            self.cache_edge_map[ids] = ids
            self.cache_edge_count.zero_()

    def resize(self, new_num_nodes: int, new_num_edges: int):
        """lfu_cache.py:120-131; the new (empty) slots start at use count 0: the first victims"""
        super(LFUCache, self).resize(new_num_nodes, new_num_edges)
        for kind in ("node", "edge"):
            if getattr(self, "dim_%s_feat" % kind) != 0:
                setattr(self, "cache_%s_count" % kind,
                        self._grown(getattr(self, "cache_%s_count" % kind), getattr(self, "%s_capacity" % kind), 0))

    def _update(self, kind, ids, hit_mask):
        feats = getattr(self, "%s_feats" % kind)
        st = self._state(kind)
        n = ids.shape[0]
        scratch = self._get_scratch(n, st.capacity, st.num_items)
        check(self._L.gf_cache_update_lfu(st, ids.data_ptr(), hit_mask.data_ptr(), n, feats.data_ptr(),
                                          self._count_bound(kind), scratch.data_ptr(), scratch.numel(), self._stream()))

    def update_node_cache(self, ids, hit_mask):
        """lfu_cache.py:133-171: use count of every hit slot += 1 (once per fetch), admit the (unique, ascending)
        misses over the slots with the smallest use count; admitted slots start at 1."""
        self._update("node", ids, hit_mask)

    def update_edge_cache(self, ids, hit_mask):
        """lfu_cache.py:173-210"""
        self._update("edge", ids, hit_mask)
