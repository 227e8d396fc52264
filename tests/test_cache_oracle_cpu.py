"""CPU checks of the cache restatements (oracle/cache_oracle.py): fetched values are feats[ids] whatever the policy
state, the flag / map / index_to_id tables stay a consistent bijection, and the LFU / static policies evict what
their definitions say on hand-checked cases (gnnflow/cache/lfu_cache.py:133-171, gnnlab_static_cache.py:87-168)."""
import numpy as np
import pytest

from oracle.cache_oracle import CacheOracle, StaticCacheOracle


@pytest.mark.parametrize("policy", ["lru", "fifo", "lfu"])
def test_values_and_tables_consistent(policy):
    rng = np.random.default_rng(1)
    feats = rng.standard_normal((400, 6)).astype(np.float32)
    o = CacheOracle(policy, 0.1, feats)
    o.init_cache()
    for _ in range(30):
        ids = np.minimum((rng.pareto(1.0, int(rng.integers(1, 200))) * 15).astype(np.int64), 399)
        out, mask, _ = o.fetch(ids)
        assert np.array_equal(out, feats[ids])
        cached = np.flatnonzero(o.flag)
        assert len(cached) == o.capacity
        assert np.array_equal(np.sort(o.index_to_id), cached)
        assert np.array_equal(o.index_to_id[o.map[cached]], cached)
        assert np.all(o.map[~o.flag] == -1)
        assert np.array_equal(o.buffer, feats[o.index_to_id])


def test_lfu_hand_case():
    feats = np.arange(40, dtype=np.float32).reshape(10, 4)
    o = CacheOracle("lfu", 0.3, feats)          # 3 slots
    o.init_cache()                              # slots 0,1,2 hold ids 0,1,2 with count 1
    assert list(o.count) == [1, 1, 1]
    o.fetch([0, 0, 0, 1, 7])                    # hits 0 (x3 -> +1 once), 1; miss 7 evicts the least used slot (2)
    assert list(o.count) == [2, 2, 1] and list(o.index_to_id) == [0, 1, 7]
    o.fetch([1, 8, 9])                          # hit 1 -> 3; two misses evict slot 2 (count 1) then slot 0 (count 2)
    assert list(o.index_to_id) == [9, 1, 8] and list(o.count) == [1, 3, 1]
    o.fetch([1, 1])                             # no miss: the reference does not update at all (cache.py:317)
    assert list(o.count) == [1, 3, 1]


def test_static_hand_case():
    feats = np.arange(24, dtype=np.float32).reshape(8, 3)
    o = StaticCacheOracle(0.5, feats)           # 4 slots
    o.presample([5, 5, 5, 2])                   # duplicates inside one block count once
    o.presample([5, 3])
    o.presample([7, 3, 2])
    o.fill()
    # counts: id5=2, id3=2, id2=2, id7=1 -> ties by lowest id: 2, 3, 5, then 7
    assert list(np.flatnonzero(o.flag)) == [2, 3, 5, 7]
    assert [int(o.map[i]) for i in (2, 3, 5, 7)] == [0, 1, 2, 3]
    out, mask, ratio = o.fetch([0, 2, 7, 7])
    assert np.array_equal(out, feats[[0, 2, 7, 7]]) and list(mask) == [False, True, True, True] and ratio == 0.75
