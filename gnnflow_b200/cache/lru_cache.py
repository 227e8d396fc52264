"""Least-recently-used cache (reference gnnflow/cache/lru_cache.py:9-201)."""
import torch

from .._lib import check
from .cache import Cache


class LRUCache(Cache):
    _policy = 0  # GF_CACHE_LRU (include/gnnflow_b200.h)

    def __init__(self, *args, **kwargs):
        super(LRUCache, self).__init__(*args, **kwargs)
        self.name = 'lru'
        if self.dim_node_feat != 0:
            self.cache_node_count = torch.zeros(self.node_capacity, dtype=torch.int32, device=self.device)
        if self.dim_edge_feat != 0:
            self.cache_edge_count = torch.zeros(self.edge_capacity, dtype=torch.int32, device=self.device)
        # lower bound of the water levels, kept on the device by gf_cache_fetch: the victim sort orders [floor, 0] only
        self._floor = {k: torch.zeros(1, dtype=torch.int32, device=self.device) for k in ("node", "edge")}
        self._track = {}

    def _count_floor(self, kind: str):
        return self._floor[kind]

    _FLOOR_READ_EVERY = 32

    def _count_bound(self, kind: str) -> int:
        """Bound of the counts the library sizes the victim sort's launches with.  Without knowledge of the device's
        state that is the number of updates so far; every few updates the floor is copied back asynchronously (no
        synchronisation: the value is used once its event has completed), after which the bound is -floor + the
        updates issued since -- in steady state a single sort pass instead of two or three, one of which would return
        at once."""
        static = super(LRUCache, self)._count_bound(kind)
        n = static - 1  # updates issued so far, this one included
        t = self._track.setdefault(kind, dict(known=None, known_at=0, issued_at=0, evt=None,
                                              pin=torch.zeros(1, dtype=torch.int32).pin_memory()))
        if t["evt"] is not None and t["evt"].query():
            t["known"], t["known_at"], t["evt"] = int(t["pin"][0]), t["issued_at"], None
        if t["evt"] is None and n - t["issued_at"] >= self._FLOOR_READ_EVERY:
            t["pin"].copy_(self._floor[kind], non_blocking=True)  # the floor after update n - 1 (same stream)
            t["evt"] = torch.cuda.Event()
            t["evt"].record()
            t["issued_at"] = n
        if t["known"] is None:
            return static
        return min(static, -t["known"] + (n - t["known_at"]) + 1)

    def _sync_count_floor(self, kind: str):
        """after cache_<kind>_count was changed outside gf_cache_fetch"""
        self._track.pop(kind, None)  # what the host knew about the floor is void
        cnt = getattr(self, "cache_%s_count" % kind, None)
        if cnt is not None and cnt.numel():
            self._floor[kind].copy_(torch.clamp(cnt.min(), max=0).reshape(1))
        else:
            self._floor[kind].zero_()

    def get_mem_size(self) -> int:
        mem_size = super(LRUCache, self).get_mem_size()
        if self.dim_node_feat != 0:
            mem_size += self.cache_node_count.element_size() * self.cache_node_count.nelement()
        if self.dim_edge_feat != 0:
            mem_size += self.cache_edge_count.element_size() * self.cache_edge_count.nelement()
        return mem_size

    def reset(self):
        """NB: only the edge cache is reset (lru_cache.py:73-106)"""
        if self.edge_feats is not None and self.dim_edge_feat != 0:
            ids = torch.arange(self.edge_capacity, dtype=torch.int64, device=self.device)
            self.cache_edge_buffer[ids] = self.edge_feats[:self.edge_capacity].to(self.device, non_blocking=True)
            self.cache_edge_flag[ids] = True
            self.cache_index_to_edge_id = ids
            self.cache_edge_map[ids] = ids
            self.cache_edge_count.zero_()
            self._sync_count_floor("edge")

    def resize(self, new_num_nodes: int, new_num_edges: int):
        """lru_cache.py:107-119; the new (empty) slots get a water level below every existing one, so they are the
        first victims"""
        super(LRUCache, self).resize(new_num_nodes, new_num_edges)
        for kind in ("node", "edge"):
            if getattr(self, "dim_%s_feat" % kind) != 0:
                cnt = getattr(self, "cache_%s_count" % kind)
                low = (cnt.min() - 1) if cnt.numel() else 0
                setattr(self, "cache_%s_count" % kind, self._grown(cnt, getattr(self, "%s_capacity" % kind), low))
                self._sync_count_floor(kind)

    def _update(self, kind, ids, hit_mask):
        feats = getattr(self, "%s_feats" % kind)
        st = self._state(kind)
        n = ids.shape[0]
        scratch = self._get_scratch(n, st.capacity, st.num_items)
        check(self._L.gf_cache_update_lru(st, ids.data_ptr(), hit_mask.data_ptr(), n, feats.data_ptr(),
                                          self._count_bound(kind), scratch.data_ptr(), scratch.numel(), self._stream()))
        self._floor[kind] -= 1  # an update lowers every water level by at most one

    def update_node_cache(self, ids, hit_mask):
        """lru_cache.py:121-160: age every slot by one, refresh the hit slots, admit the (unique, ascending) misses
        over the slots with the smallest water level."""
        self._update("node", ids, hit_mask)

    def update_edge_cache(self, ids, hit_mask):
        """lru_cache.py:162-201"""
        self._update("edge", ids, hit_mask)
