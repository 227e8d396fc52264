"""
TEST INFRASTRUCTURE -- numpy restatement of the reference feature caches (parity unpinned: the reference has no
cache tests; the contract checked is (1) fetched values == feats[ids] and (2) the policy
restated here from the reference source).

  fetch            <- gnnflow/cache/cache.py:255-413
  LRU update       <- gnnflow/cache/lru_cache.py:121-160  (torch.topk tie-break is unspecified there; here: the
                      slot with the lowest index wins, which is what the CUDA path implements)
  FIFO update      <- gnnflow/cache/fifo_cache.py:77-118
  LFU update       <- gnnflow/cache/lfu_cache.py:133-171  (`count[cached_index] += 1` is a non-accumulating
                      index_put: a slot hit several times in one fetch gains 1; same tie-break as LRU)
  GNNLab static    <- gnnflow/cache/gnnlab_static_cache.py:87-168 (StaticCacheOracle; top-k ties -> lowest id)
  init_cache       <- gnnflow/cache/cache.py:175-195, fifo_cache.py:57-68
"""
import numpy as np


class CacheOracle:
    def __init__(self, policy, ratio, feats):
        assert policy in ("lru", "fifo", "lfu")
        self.policy = policy
        self.feats = np.asarray(feats, dtype=np.float32)
        self.n, self.dim = self.feats.shape
        self.capacity = int(ratio * self.n)
        self.buffer = np.zeros((self.capacity, self.dim), np.float32)
        self.flag = np.zeros(self.n, bool)
        self.map = np.full(self.n, -1, np.int64)
        self.index_to_id = np.full(self.capacity, -1, np.int64)
        self.count = np.zeros(self.capacity, np.int32)
        self.pointer = 0

    def init_cache(self):
        ids = np.arange(self.capacity)
        self.buffer[ids] = self.feats[:self.capacity]
        self.flag[ids] = True
        self.index_to_id = ids.astype(np.int64)
        self.map[ids] = ids
        if self.policy == "fifo":
            self.pointer = self.capacity - 1
        if self.policy == "lfu":  # lfu_cache.py:80-84
            self.count[self.index_to_id] += 1

    def fetch(self, ids, update_cache=True):
        ids = np.asarray(ids, dtype=np.int64)
        mask = self.flag[ids]
        out = np.zeros((len(ids), self.dim), np.float32)
        cached_index = self.map[ids[mask]]
        out[mask] = self.buffer[cached_index]
        uncached = ids[~mask]
        hit_ratio = mask.sum() / max(1, len(ids))
        if len(uncached) > 0:
            uniq, inv = np.unique(uncached, return_inverse=True)
            ufeat = self.feats[uniq]
            out[~mask] = ufeat[inv]
            if update_cache and self.capacity > 0:
                self._update(cached_index, uniq, ufeat)
        return out, mask, hit_ratio

    def _update(self, cached_index, uncached_id, uncached_feature):
        k = min(len(uncached_id), self.capacity)
        ids_to_cache, feat_to_cache = uncached_id[:k], uncached_feature[:k]
        if self.policy == "lru":
            self.count -= 1
            self.count[cached_index] = 0
            removing = np.argsort(self.count, kind="stable")[:k]
        elif self.policy == "lfu":
            self.count[np.unique(cached_index)] += 1
            removing = np.argsort(self.count, kind="stable")[:k]
        else:
            if self.pointer + k < self.capacity:
                removing = np.arange(self.pointer + 1, self.pointer + k + 1)
                self.pointer = self.pointer + k
            else:
                r = k - (self.capacity - 1 - self.pointer)
                removing = np.concatenate([np.arange(r), np.arange(self.pointer + 1, self.capacity)])
                self.pointer = r - 1
        removing_id = self.index_to_id[removing]
        self.buffer[removing] = feat_to_cache
        if self.policy == "lru":
            self.count[removing] = 0
        if self.policy == "lfu":
            self.count[removing] = 1
        live = removing_id >= 0  # the reference indexes flag[-1] for empty slots (a quirk not reproduced)
        self.flag[removing_id[live]] = False
        self.flag[ids_to_cache] = True
        self.map[removing_id[live]] = -1
        self.map[ids_to_cache] = removing
        self.index_to_id[removing] = ids_to_cache


class StaticCacheOracle:
    """GNNLab static cache: sampling statistics -> top-k rows, never updated afterwards."""

    def __init__(self, ratio, feats):
        self.feats = np.asarray(feats, dtype=np.float32)
        self.n, self.dim = self.feats.shape
        self.capacity = int(ratio * self.n)
        self.sampled_count = np.zeros(self.n, np.int32)
        self.buffer = np.zeros((self.capacity, self.dim), np.float32)
        self.flag = np.zeros(self.n, bool)
        self.map = np.full(self.n, -1, np.int64)

    def presample(self, ids):
        self.sampled_count[np.unique(np.asarray(ids, dtype=np.int64))] += 1  # gnnlab_static_cache.py:104-111

    def fill(self):
        order = np.argsort(-self.sampled_count.astype(np.int64), kind="stable")[:self.capacity]  # :130-141
        self.buffer[:] = self.feats[order]
        self.flag[:] = False
        self.map[:] = -1
        self.flag[order] = True
        self.map[order] = np.arange(self.capacity)

    def fetch(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        mask = self.flag[ids]
        out = self.feats[ids].copy()
        out[mask] = self.buffer[self.map[ids[mask]]]
        return out, mask, mask.sum() / max(1, len(ids))
