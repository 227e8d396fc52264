"""Feature-cache gather + LRU / FIFO policy kernels against the numpy restatement and the identity
out == feats[ids] (bit-exact)."""
import numpy as np
import pytest
import torch

from helpers import assert_same
from oracle.cache_oracle import CacheOracle

pytestmark = pytest.mark.gpu


class FakeBlock:
    def __init__(self, node_ids, edge_ids):
        self.srcdata = {"ID": node_ids}
        self.edata = {"ID": edge_ids}


def _mk(policy, ratio, nfeat, efeat, where):
    from gnnflow_b200.cache import FIFOCache, LRUCache
    cls = LRUCache if policy == "lru" else FIFOCache
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
    if policy == "lru":
        assert_same(tag + ".count", getattr(c, "cache_%s_count" % kind).cpu().numpy(), o.count)
    else:
        assert getattr(c, "cache_%s_pointer" % kind) == o.pointer, tag


@pytest.mark.parametrize("where", ["cuda", "pinned", "pageable"])
@pytest.mark.parametrize("policy,ratio,dn,de,init", [
    ("lru", 0.2, 172, 172, True), ("fifo", 0.2, 172, 172, True), ("lru", 0.05, 413, 186, False),
    ("fifo", 0.03, 7, 3, False), ("lru", 1.0, 16, 8, True), ("fifo", 0.5, 130, 2, False)])
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
