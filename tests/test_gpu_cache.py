"""Feature-cache gather + LRU / FIFO / LFU policy kernels and the GNNLab static cache against the numpy restatement and the identity
out == feats[ids] (bit-exact)."""
import numpy as np
import pytest
import torch

from helpers import assert_same
from oracle.cache_oracle import CacheOracle, StaticCacheOracle

pytestmark = pytest.mark.gpu


class FakeBlock:
    def __init__(self, node_ids, edge_ids):
        self.srcdata = {"ID": node_ids}
        self.edata = {"ID": edge_ids}


def _mk(policy, ratio, nfeat, efeat, where):
    from gnnflow_b200.cache import FIFOCache, LFUCache, LRUCache
    cls = {"lru": LRUCache, "fifo": FIFOCache, "lfu": LFUCache}[policy]
    nf, ef = torch.from_numpy(nfeat), torch.from_numpy(efeat)
    if where == "cuda":
        nf, ef = nf.cuda(), ef.cuda()
    elif where == "pinned":
        nf, ef = nf.pin_memory(), ef.pin_memory()
    c = cls(ratio, ratio, nfeat.shape[0], efeat.shape[0], "cuda", nf, ef, nfeat.shape[1], efeat.shape[1])
    return c


def _check_state(tag, c, kind, o, policy):
    assert_same(tag + ".flag", getattr(c, "cache_%s_flag" % kind).cpu().numpy(), o.flag)
    assert_same(tag + ".map", getattr(c, "cache_%s_map" % kind).cpu().numpy(), o.map)
    assert_same(tag + ".index_to_id", getattr(c, "cache_index_to_%s_id" % kind).cpu().numpy(), o.index_to_id)
    assert_same(tag + ".buffer", getattr(c, "cache_%s_buffer" % kind).cpu().numpy().ravel(), o.buffer.ravel())
    if policy in ("lru", "lfu"):
        assert_same(tag + ".count", getattr(c, "cache_%s_count" % kind).cpu().numpy(), o.count)
    else:
        assert getattr(c, "cache_%s_pointer" % kind) == o.pointer, tag


@pytest.mark.parametrize("where", ["cuda", "pinned", "pageable"])
@pytest.mark.parametrize("policy,ratio,dn,de,init", [
    ("lru", 0.2, 172, 172, True), ("fifo", 0.2, 172, 172, True), ("lru", 0.05, 413, 186, False),
    ("fifo", 0.03, 7, 3, False), ("lru", 1.0, 16, 8, True), ("fifo", 0.5, 130, 2, False),
    ("lfu", 0.2, 172, 172, True), ("lfu", 0.04, 33, 5, False), ("lfu", 1.0, 16, 8, True)])
def test_cache_parity(policy, ratio, dn, de, init, where):
    rng = np.random.default_rng(17)
    N, E = 700, 3000
    nfeat = rng.standard_normal((N, dn)).astype(np.float32)
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    c = _mk(policy, ratio, nfeat, efeat, where)
    on, oe = CacheOracle(policy, ratio, nfeat), CacheOracle(policy, ratio, efeat)
    if init:
        c.init_cache(); on.init_cache(); oe.init_cache()
    for step in range(12):
        # skewed ids with duplicates; sizes change every step, one step is larger than the node capacity
        n_ids = int(rng.integers(1, 900))
        nid = np.minimum((rng.pareto(1.2, n_ids) * 20).astype(np.int64), N - 1)
        eid = [np.minimum((rng.pareto(1.0, int(rng.integers(1, 1500))) * 50).astype(np.int64), E - 1) for _ in range(2)]
        blocks = [[FakeBlock(torch.from_numpy(nid).cuda(), torch.from_numpy(eid[0]).cuda())],
                  [FakeBlock(torch.from_numpy(nid[:5]).cuda(), torch.from_numpy(eid[1]).cuda())]]
        tgt = rng.integers(0, E, 50)
        c.fetch_feature(blocks, eid=tgt)
        h, _, hr = on.fetch(nid)
        assert_same("step%d.h" % step, blocks[0][0].srcdata['h'].cpu().numpy().ravel(), nfeat[nid].ravel())
        assert_same("step%d.h.oracle" % step, h.ravel(), nfeat[nid].ravel())
        assert float(c.cache_node_ratio) == pytest.approx(hr, abs=1e-6)
        ratios = []
        for k in range(2):
            f, _, r = oe.fetch(eid[k])
            ratios.append(r)
            got = blocks[k][0].edata['f'].cpu().numpy()
            assert_same("step%d.f%d" % (step, k), got.ravel(), efeat[eid[k]].ravel())
        assert float(c.cache_edge_ratio) == pytest.approx(np.mean(ratios), abs=1e-6)
        assert_same("step%d.target" % step, c.target_edge_features.cpu().numpy().ravel(), efeat[tgt].ravel())
        _check_state("step%d.node" % step, c, "node", on, policy)
        _check_state("step%d.edge" % step, c, "edge", oe, policy)


def test_cache_no_update_and_empty_blocks():
    rng = np.random.default_rng(3)
    nfeat = rng.standard_normal((100, 12)).astype(np.float32)
    efeat = rng.standard_normal((400, 20)).astype(np.float32)
    c = _mk("lru", 0.1, nfeat, efeat, "cuda")
    c.init_cache()
    before = c.cache_edge_map.clone()
    nid = torch.from_numpy(rng.integers(0, 100, 64)).cuda()
    blocks = [[FakeBlock(nid, torch.empty(0, dtype=torch.int64, device="cuda"))],
              [FakeBlock(nid, torch.from_numpy(rng.integers(0, 400, 300)).cuda())]]
    c.fetch_feature(blocks, update_cache=False, target_edge_features=False)
    assert 'f' not in blocks[0][0].edata  # empty block skipped (cache.py:342-343)
    assert torch.equal(before, c.cache_edge_map)
    assert_same("h", blocks[0][0].srcdata['h'].cpu().numpy().ravel(), nfeat[nid.cpu().numpy()].ravel())
    c.reset()
    assert c.get_mem_size() > 0


@pytest.mark.parametrize("ratio", [0.0, 0.15, 1.0])
def test_gnnlab_static_cache(ratio):
    """pre-sampling statistics + top-k fill through the real sampler, against the numpy restatement"""
    import pandas as pd
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.cache import GNNLabStaticCache
    from gnnflow_b200.cache.gnnlab_static_cache import get_batch_no_neg
    rng = np.random.default_rng(5)
    N, E, dn, de = 300, 6000, 24, 20
    src = np.minimum((rng.pareto(1.1, E) * 8).astype(np.int64), 199)
    dst = rng.integers(200, N, E).astype(np.int64)
    ts = np.sort(rng.uniform(0, 1000, E)).astype(np.float32)
    eid = np.arange(E, dtype=np.int64)
    g = DynamicGraph(initial_pool_size=8 << 20, maximum_pool_size=1 << 28, mem_resource_type="cuda", minimum_block_size=8,
                     blocks_to_preallocate=64, insertion_policy="insert")
    g.add_edges(src, dst, ts, eid, add_reverse=True)
    smp = TemporalSampler(g, [5, 5], "recent")
    df = pd.DataFrame({"src": src, "dst": dst, "time": ts, "eid": eid})
    nfeat = rng.standard_normal((N, dn)).astype(np.float32)
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    c = GNNLabStaticCache(ratio, N, E, "cuda", torch.from_numpy(nfeat).cuda(), torch.from_numpy(efeat).cuda(), dn, de)
    c.init_cache(sampler=smp, train_df=df, pre_sampling_rounds=2, batch_size=500)
    on, oe = StaticCacheOracle(ratio, nfeat), StaticCacheOracle(ratio, efeat)
    for _ in range(2):
        for roots, rts, _e in get_batch_no_neg(df, 500):
            mfgs = smp.sample(roots, rts)
            for b in mfgs[0]:
                on.presample(b.srcdata['ID'].cpu().numpy())
            for mfg in mfgs:
                for b in mfg:
                    if b.num_src_nodes() > b.num_dst_nodes():
                        oe.presample(b.edata['ID'].cpu().numpy())
    on.fill(); oe.fill()
    assert_same("node.count", c.node_sampled_count.cpu().numpy(), on.sampled_count)
    assert_same("edge.count", c.edge_sampled_count.cpu().numpy(), oe.sampled_count)
    for kind, o in (("node", on), ("edge", oe)):
        assert_same(kind + ".flag", getattr(c, "cache_%s_flag" % kind).cpu().numpy(), o.flag)
        assert_same(kind + ".map", getattr(c, "cache_%s_map" % kind).cpu().numpy(), o.map)
        assert_same(kind + ".buffer", getattr(c, "cache_%s_buffer" % kind).cpu().numpy().ravel(), o.buffer.ravel())
    # fetch: values are feats[ids]; the hit ratio is the oracle's; the cache never changes
    before = c.cache_edge_map.clone()
    roots, rts, e0 = next(iter(get_batch_no_neg(df.iloc[3000:3600].reset_index(drop=True), 600)))
    mfgs = c.fetch_feature(smp.sample(roots, rts), eid=e0)
    b = mfgs[0][0]
    ids = b.srcdata['ID'].cpu().numpy()
    assert_same("h", b.srcdata['h'].cpu().numpy().ravel(), nfeat[ids].ravel())
    assert float(c.cache_node_ratio) == pytest.approx(on.fetch(ids)[2], abs=1e-6)
    assert_same("f", b.edata['f'].cpu().numpy().ravel(), efeat[b.edata['ID'].cpu().numpy()].ravel())
    assert_same("target", c.target_edge_features.cpu().numpy().ravel(), efeat[e0].ravel())
    assert torch.equal(before, c.cache_edge_map)
    c.reset()
    assert c.get_mem_size() >= 0


@pytest.mark.parametrize("policy", ["lru", "fifo", "lfu"])
def test_cache_resize_and_bad_ids(policy):
    """Cache.resize (cache.py:197-221) after the dataset grew: what was cached stays cached, new ids start uncached,
    new slots empty (never uninitialised), values stay == feats[ids] through further fetches and policy updates; ids
    beyond the table are zero-filled and reported (the reference's torch indexing raises IndexError)."""
    rng = np.random.default_rng(23)
    N0, E0, N1, E1, dn, de = 500, 2000, 900, 5000, 20, 12
    nfeat = rng.standard_normal((N1, dn)).astype(np.float32)
    efeat = rng.standard_normal((E1, de)).astype(np.float32)
    c = _mk(policy, 0.2, nfeat[:N0].copy(), efeat[:E0].copy(), "cuda")
    c.init_cache()

    def fetch(nn, ne):
        nid = rng.integers(0, nn, 700).astype(np.int64)
        eid = rng.integers(0, ne, 1500).astype(np.int64)
        b = FakeBlock(torch.from_numpy(nid).cuda(), torch.from_numpy(eid).cuda())
        c.fetch_feature([[b]])
        assert_same("h", b.srcdata['h'].cpu().numpy().ravel(), nfeat[nid].ravel())
        assert_same("f", b.edata['f'].cpu().numpy().ravel(), efeat[eid].ravel())

    for _ in range(4):
        fetch(N0, E0)
    cached_before = c.cache_node_flag.clone()
    # an id beyond the table before the resize: zero row + IndexError at check_ids, nothing dereferenced
    b = FakeBlock(torch.tensor([3, N0 + 5, -1], dtype=torch.int64).cuda(), torch.tensor([1, E0], dtype=torch.int64).cuda())
    c.fetch_feature([[b]])
    h = b.srcdata['h'].cpu().numpy()
    assert_same("h[0]", h[0], nfeat[3])
    assert not h[1].any() and not h[2].any() and not b.edata['f'][1].cpu().numpy().any()
    with pytest.raises(IndexError):
        c.check_ids()
    c.check_ids()  # the counter was reset
    c.resize(N1, E1)
    c.set_feats(torch.from_numpy(nfeat).cuda(), torch.from_numpy(efeat).cuda())
    assert c.num_nodes == N1 and c.node_capacity == int(0.2 * N1) and c.edge_capacity == int(0.2 * E1)
    assert c.cache_node_flag.shape[0] == N1 and c.cache_node_map.shape[0] == N1
    assert c.cache_node_buffer.shape == (c.node_capacity, dn) and c.cache_index_to_node_id.shape[0] == c.node_capacity
    assert torch.equal(c.cache_node_flag[:N0], cached_before) and not c.cache_node_flag[N0:].any()
    assert (c.cache_node_map[N0:] == -1).all() and (c.cache_index_to_node_id[int(0.2 * N0):] == -1).all()
    for _ in range(8):
        fetch(N1, E1)
    c.check_ids()
    # consistency of the grown state: flag <=> map, map <-> index_to_id inverse of each other
    flag, mp, i2id = c.cache_node_flag.cpu().numpy(), c.cache_node_map.cpu().numpy(), c.cache_index_to_node_id.cpu().numpy()
    assert np.array_equal(flag, mp >= 0)
    assert np.array_equal(i2id[mp[flag]], np.nonzero(flag)[0])
    assert flag[N0:].any()  # new ids did get admitted into the new slots


@pytest.mark.parametrize("policy", ["lru", "lfu", "fifo"])
def test_cache_many_small_updates(policy):
    """Hundreds of updates that admit only a few ids each: the static bound of the counts passes 2^8 (the host launches
    two passes of the victim sort) while the water levels on the device first span less than a pass (the second
    returns at once) and later more than one (slots that are never touched grow old): every stage of
    gf_cache_fetch's count-floor logic, state compared with the numpy restatement as it goes."""
    rng = np.random.default_rng(29)
    E, de = 4000, 12
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    nfeat = rng.standard_normal((10, 4)).astype(np.float32)
    c = _mk(policy, 0.5, nfeat, efeat, "cuda")
    oe = CacheOracle(policy, 0.5, efeat)
    c.init_cache(); oe.init_cache()
    nid = torch.zeros(1, dtype=torch.int64, device="cuda")
    for step in range(420):
        # a hot set that keeps hitting + a slow sweep over cold ids (1-4 admissions per update)
        hot = rng.integers(0, 300, int(rng.integers(4, 40)))
        cold = (2000 + (3 * step + rng.integers(0, 4, int(rng.integers(1, 5)))) % 2000)
        eid = np.concatenate([hot, cold]).astype(np.int64)
        rng.shuffle(eid)
        b = FakeBlock(nid, torch.from_numpy(eid).cuda())
        c.fetch_feature([[b]])
        f, _, r = oe.fetch(eid)
        if step % 20 == 19 or step > 400:
            assert_same("step%d.f" % step, b.edata['f'].cpu().numpy().ravel(), efeat[eid].ravel())
            assert float(c.cache_edge_ratio) == pytest.approx(r, abs=1e-6)
            _check_state("step%d.edge" % step, c, "edge", oe, policy)
    if policy == "lru":  # the scenario did reach both regimes
        assert int(c.cache_edge_count.min()) < -256
        assert int(c._floor["edge"].item()) <= int(c.cache_edge_count.min())


def test_cache_standalone_update_between_fetches():
    """update_edge_cache (the stand-alone call: gather elsewhere, update here) interleaved with fetch_feature: the
    count floor kept by the fused call stays a lower bound"""
    rng = np.random.default_rng(31)
    E, de = 3000, 8
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    nfeat = rng.standard_normal((10, 4)).astype(np.float32)
    c = _mk("lru", 0.3, nfeat, efeat, "cuda")
    oe = CacheOracle("lru", 0.3, efeat)
    c.init_cache(); oe.init_cache()
    nid = torch.zeros(1, dtype=torch.int64, device="cuda")
    for step in range(60):
        eid = np.minimum((rng.pareto(1.0, int(rng.integers(1, 400))) * 40).astype(np.int64), E - 1)
        if step % 3 == 1:
            ids, feat, hit, _ = c._gather("edge", torch.from_numpy(eid).cuda())
            c.update_edge_cache(ids, hit)
            got = feat
        else:
            b = FakeBlock(nid, torch.from_numpy(eid).cuda())
            c.fetch_feature([[b]])
            got = b.edata['f']
        oe.fetch(eid)
        assert_same("step%d.f" % step, got.cpu().numpy().ravel(), efeat[eid].ravel())
        _check_state("step%d.edge" % step, c, "edge", oe, "lru")
        assert int(c._floor["edge"].item()) <= int(c.cache_edge_count.min())


@pytest.mark.parametrize("policy", ["lru", "fifo"])
def test_cache_large_id_space(policy):
    """3 M ids: the bitmap has more 256-bit chunks than one CTA ranks in shared memory, so the chunk prefixes come from the
    look-back scan in its own launch (the path a GDELT-sized edge-feature table takes)"""
    rng = np.random.default_rng(37)
    E, de = 3_000_000, 4
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    nfeat = rng.standard_normal((10, 4)).astype(np.float32)
    c = _mk(policy, 0.01, nfeat, efeat, "cuda")
    oe = CacheOracle(policy, 0.01, efeat)
    c.init_cache(); oe.init_cache()
    nid = torch.zeros(1, dtype=torch.int64, device="cuda")
    for step in range(6):
        eid = np.concatenate([rng.integers(0, E, 20000), rng.integers(0, 40000, 20000)]).astype(np.int64)
        b = FakeBlock(nid, torch.from_numpy(eid).cuda())
        c.fetch_feature([[b]])
        _, _, r = oe.fetch(eid)
        assert_same("step%d.f" % step, b.edata['f'].cpu().numpy().ravel(), efeat[eid].ravel())
        assert float(c.cache_edge_ratio) == pytest.approx(r, abs=1e-6)
        _check_state("step%d.edge" % step, c, "edge", oe, policy)


def test_cache_big_fetch_small_id_space():
    """a fetch of 100 000 ids over 6 000 ids: the rank pass has too many CTAs to let each scan the bitmap itself"""
    rng = np.random.default_rng(41)
    E, de = 6000, 8
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    nfeat = rng.standard_normal((10, 4)).astype(np.float32)
    c = _mk("lru", 0.3, nfeat, efeat, "cuda")
    oe = CacheOracle("lru", 0.3, efeat)
    c.init_cache(); oe.init_cache()
    nid = torch.zeros(1, dtype=torch.int64, device="cuda")
    for step in range(4):
        eid = np.minimum((rng.pareto(0.6, 100_000) * 300).astype(np.int64), E - 1)
        b = FakeBlock(nid, torch.from_numpy(eid).cuda())
        c.fetch_feature([[b]])
        oe.fetch(eid)
        assert_same("step%d.f" % step, b.edata['f'].cpu().numpy().ravel(), efeat[eid].ravel())
        _check_state("step%d.edge" % step, c, "edge", oe, "lru")
