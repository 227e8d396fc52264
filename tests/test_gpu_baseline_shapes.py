"""Parity on the BASELINE.json shapes: the CUDA path against the CPU oracle at the sizes, block
sizes and pipelines the five configs name -- not on toy graphs.

  config 1  WIKI-shaped, full size (9,227 nodes, 157,474 edges, undirected -> 314,948 stored, min block 18), TGN [10] recent
  config 2  REDDIT-shaped, full size (10,984 nodes, 672,447 edges, min block 62), ingest in 100,000-edge batches,
            TGN [10] recent over all 1,121 batches of 600 (2,017,341 targets)
  config 3  REDDIT-shaped TGAT [10,10] uniform -> LRUCache(0.2).fetch_feature == feats[ID]
  config 4  GDELT-16.7K-shaped (16,682 vertices, blocks of >= 10^4 edges, min block 123), DySAT [10,10] uniform,
            3 snapshots, window 25, prop_time
  config 5  GDELT-16.7M-shaped ids: add_edges(100,000) interleaved with 2-layer sampling of the edges just added
"""
import numpy as np
import pytest
import torch

from helpers import assert_same, compare_block, compare_graphs
from oracle.oracle import OracleGraph, OracleSampler

pytestmark = pytest.mark.gpu

GB = 1 << 30


def _cfg(stream, **over):
    c = dict(initial_pool_size=20 << 20, maximum_pool_size=16 * GB, mem_resource_type="cuda",
             minimum_block_size=stream["minimum_block_size"], blocks_to_preallocate=1024, insertion_policy="insert")
    c.update(over)
    return c


def _build(stream, batch, add_reverse=False, device_arrays=False):
    from gnnflow_b200 import DynamicGraph
    g, og = DynamicGraph(**_cfg(stream)), OracleGraph(**_cfg(stream))
    n = len(stream["src"])
    for lo in range(0, n, batch):
        sl = slice(lo, lo + batch)
        a = [stream[k][sl] for k in ("src", "dst", "ts", "eid")]
        og.add_edges(*a, add_reverse=add_reverse)
        if device_arrays:
            a = [torch.from_numpy(x).cuda() for x in a]
        g.add_edges(*a, add_reverse=add_reverse)
    return g, og


def _check_batched(tag, g, og, nodes, rts, offs, fanout, strategy, every=1):
    """one multi-batch launch over the whole replay == the oracle batch by batch"""
    from gnnflow_b200 import TemporalSampler
    s, os_ = TemporalSampler(g, [fanout], strategy), OracleSampler(og, [fanout], strategy)
    out = s.sample_layer_batched(torch.from_numpy(nodes).cuda(), torch.from_numpy(rts).cuda(), torch.from_numpy(offs).cuda())
    eo = out["edge_offsets"].cpu().numpy()
    res = {k: out[k].cpu().numpy() for k in ("nbr", "ts", "dt", "eid", "row")}
    nb = len(offs) - 1
    total = 0
    for b in range(nb):
        if b % every and strategy == "recent":
            continue  # (uniform: every batch advances the oracle's RNG launch index, none can be skipped)
        o = os_.sample_layer(nodes[offs[b]:offs[b + 1]], rts[offs[b]:offs[b + 1]], 0, 0)
        T = offs[b + 1] - offs[b]
        sl = slice(eo[b], eo[b + 1])
        assert eo[b + 1] - eo[b] == len(o["eids"]), "%s batch %d: %d vs %d neighbours" % (tag, b, eo[b + 1] - eo[b], len(o["eids"]))
        assert_same("%s.b%d.nbr" % (tag, b), res["nbr"][sl], o["all_nodes"][T:])
        assert_same("%s.b%d.ts" % (tag, b), res["ts"][sl], o["all_timestamps"][T:])
        assert_same("%s.b%d.dt" % (tag, b), res["dt"][sl], o["delta_timestamps"])
        assert_same("%s.b%d.eid" % (tag, b), res["eid"][sl], o["eids"])
        assert_same("%s.b%d.row" % (tag, b), res["row"][sl], o["row"])
        total += len(o["eids"])
    return total


def test_config1_wiki_full_shape():
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth("WIKI", seed=42)
    assert stream["undirected"] and stream["minimum_block_size"] == 18 and len(stream["src"]) == 157474
    g, og = _build(stream, 100000, add_reverse=True)
    assert g.num_edges() == og.num_edges() == 157474  # both directions share the eid
    rng = np.random.default_rng(0)
    compare_graphs(g, og, np.unique(rng.integers(0, stream["num_nodes"], 300)))
    nodes, rts, offs = tgn_batches(stream, 600, seed=7)
    S = _check_batched("wiki.recent", g, og, nodes, rts, offs, 10, "recent")
    assert S > 1_000_000
    _check_batched("wiki.uniform", g, og, nodes, rts, offs, 10, "uniform")


def test_config2_reddit_full_shape_ingest_and_recent():
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth("REDDIT", seed=42)
    assert stream["minimum_block_size"] == 62 and len(stream["src"]) == 672447
    g, og = _build(stream, 100000, device_arrays=True)
    assert g.num_edges() == 672447 and g.num_vertices() == og.num_vertices()
    assert g.get_graph_memory_usage() == og.get_graph_memory_usage()
    assert abs(g.avg_linked_list_length() - og.avg_linked_list_length()) < 1e-6
    rng = np.random.default_rng(1)
    compare_graphs(g, og, np.unique(np.concatenate([np.arange(40), rng.integers(0, stream["num_nodes"], 200)])))
    nodes, rts, offs = tgn_batches(stream, 600, seed=7)
    assert len(nodes) == 2017341
    S = _check_batched("reddit.recent", g, og, nodes, rts, offs, 10, "recent")
    assert S > 10_000_000
    # the reference's other ingest batch size (scripts/offline_edge_prediction.py:54): 1,000-edge batches, same store
    sub = {k: (v[:150000] if isinstance(v, np.ndarray) else v) for k, v in stream.items()}
    g2, og2 = _build(sub, 1000)
    compare_graphs(g2, og2, np.unique(np.concatenate([np.arange(20), rng.integers(0, stream["num_nodes"], 100)])))


@pytest.mark.parametrize("policy", ["lru", "fifo"])
def test_config3_reddit_tgat_uniform_with_cache(policy):
    """TGAT [10,10] uniform through TemporalSampler.sample -> Cache.fetch_feature, as the training loop chains them
    (scripts/offline_edge_prediction.py:385-459): MFGs == oracle, b.srcdata['h'] / b.edata['f'] == feats[ID]"""
    from gnnflow_b200 import TemporalSampler
    from gnnflow_b200.cache import FIFOCache, LRUCache
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth("REDDIT", seed=42)
    g, og = _build(stream, 100000, device_arrays=True)
    nodes, rts, offs = tgn_batches(stream, 600, seed=7)
    s, os_ = TemporalSampler(g, [10, 10], "uniform"), OracleSampler(og, [10, 10], "uniform")
    N, E, dn, de = stream["num_nodes"], len(stream["src"]), 16, 172  # REDDIT: 172-dim edge features
    rng = np.random.default_rng(3)
    nfeat = rng.standard_normal((N, dn)).astype(np.float32)
    efeat = rng.standard_normal((E, de)).astype(np.float32)
    cls = LRUCache if policy == "lru" else FIFOCache
    cache = cls(0.2, 0.2, N, E, "cuda", torch.from_numpy(nfeat).cuda(), torch.from_numpy(efeat).pin_memory(), dn, de)
    cache.init_cache()
    nb = len(offs) - 1
    for b in list(range(0, 12)) + list(range(400, 1121, 60)):
        r, t = nodes[offs[b]:offs[b + 1]], rts[offs[b]:offs[b + 1]]
        mfgs, om = s.sample(r, t), os_.sample(r, t)
        for l in range(2):
            compare_block("tgat.b%d.l%d" % (b, l), mfgs[l][0], om[l][0])
        eid = stream["eid"][b * 600:(b + 1) * 600]
        cache.fetch_feature(mfgs, eid)
        blk = mfgs[0][0]
        assert_same("h", blk.srcdata['h'].cpu().numpy().ravel(), nfeat[om[0][0]["all_nodes"]].ravel())
        for l in range(2):
            if len(om[l][0]["eids"]):
                assert_same("f.l%d" % l, mfgs[l][0].edata['f'].cpu().numpy().ravel(), efeat[om[l][0]["eids"]].ravel())
        assert_same("target_edge_features", cache.target_edge_features.cpu().numpy().ravel(), efeat[eid].ravel())
    assert 0.0 < float(cache.cache_edge_ratio) <= 1.0
    cache.check_ids()


def _gdelt_like(num_nodes, n, t_max, seed):
    from gnnflow_b200.synth import synth_stream
    src, dst, ts, eid = synth_stream(num_nodes, 0, n, seed=seed, t_max=t_max)
    return dict(src=src, dst=dst, ts=ts, eid=eid, num_nodes=num_nodes, minimum_block_size=123, undirected=False,
                name="GDELT-like")


def test_config4_gdelt16k_dysat_three_snapshots():
    """GDELT-16.7K shape: 16,682 vertices; Zipf(0.8) sources put > 10^4 edges on the hot vertices, so their blocks reach
    10^4 edges (every pivot level of blk_lower_bound).  DySAT: [10,10] uniform, 3 snapshots, window 25, prop_time
    (config.py:61-76, scripts/offline_edge_prediction.py:171-172).  The time axis is scaled so that a window of 25
    holds tens to hundreds of edges of a hot vertex, as on the real dataset."""
    from gnnflow_b200 import TemporalSampler
    stream = _gdelt_like(16682, 4_000_000, 4000.0, seed=5)
    g, og = _build(stream, 500000, device_arrays=True)
    sizes = g.block_shapes(0)[0]
    assert sizes.max() >= 10_000, sizes.max()
    compare_graphs(g, og, np.array([0, 1, 2, 17, 500, 16000]))
    kw = dict(sample_strategy="uniform", num_snapshots=3, snapshot_time_window=25.0, prop_time=True)
    s, os_ = TemporalSampler(g, [10, 10], **kw), OracleSampler(og, [10, 10], **kw)
    rng = np.random.default_rng(9)
    n = len(stream["src"])
    for lo in (3_999_400, 2_000_000, 40_000):
        roots = np.concatenate([stream["src"][lo:lo + 600], stream["dst"][lo:lo + 600],
                                rng.integers(0, 16682, 600)]).astype(np.int64)
        rts = np.concatenate([stream["ts"][lo:lo + 600]] * 3).astype(np.float32)
        mfgs, om = s.sample(roots, rts), os_.sample(roots, rts)
        assert len(mfgs) == 2 and len(mfgs[0]) == 3
        tot = 0
        for l in range(2):
            for k in range(3):
                compare_block("dysat.lo%d.l%d.s%d" % (lo, l, k), mfgs[l][k], om[l][k])
                tot += mfgs[l][k].num_edges()
        assert tot > 0
    # recent policy on the same store (window arithmetic over blocks of 10^4 edges)
    kw["sample_strategy"] = "recent"
    s, os_ = TemporalSampler(g, [10, 10], **kw), OracleSampler(og, [10, 10], **kw)
    mfgs, om = s.sample(roots, rts), os_.sample(roots, rts)
    for l in range(2):
        for k in range(3):
            compare_block("dysat.recent.l%d.s%d" % (l, k), mfgs[l][k], om[l][k])
    assert n == g.num_edges()


@pytest.mark.parametrize("strategy", ["recent", "uniform"])
def test_config5_online_ingest_interleaved_with_sampling(strategy):
    """add_edges(100,000) alternating with 2-layer sampling of the batches those edges form
    (scripts/online_edge_prediction.py), on GDELT-16.7M-shaped ids (16.7 M-entry vertex table, almost every touched
    vertex gets its own minimum-size block), sliding-window offload every other round."""
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    stream = _gdelt_like(16_700_000, 600_000, 2.6e6, seed=6)
    g, og = DynamicGraph(**_cfg(stream)), OracleGraph(**_cfg(stream))
    s, os_ = TemporalSampler(g, [10, 10], strategy), OracleSampler(og, [10, 10], strategy)
    rng = np.random.default_rng(2)
    for it, lo in enumerate(range(0, 600_000, 100_000)):
        sl = slice(lo, lo + 100_000)
        a = [stream[k][sl] for k in ("src", "dst", "ts", "eid")]
        g.add_edges(*[torch.from_numpy(x).cuda() for x in a])
        og.add_edges(*a)
        for b0 in (lo, lo + 50_000, lo + 99_400):
            roots = np.concatenate([stream["src"][b0:b0 + 600], stream["dst"][b0:b0 + 600],
                                    rng.integers(0, 16_700_000, 600)]).astype(np.int64)
            rts = np.concatenate([stream["ts"][b0:b0 + 600]] * 3).astype(np.float32)
            mfgs, om = s.sample(roots, rts), os_.sample(roots, rts)
            for l in range(2):
                compare_block("online.%s.it%d.b%d.l%d" % (strategy, it, b0, l), mfgs[l][0], om[l][0])
        if it % 2 == 1:
            t_old = float(stream["ts"][lo]) - 3e5
            assert g.offload_old_blocks(t_old) == og.offload_old_blocks(t_old)
            assert g.num_edges() == og.num_edges()
    assert g.num_vertices() == og.num_vertices() and g.max_vertex_id() == og.max_vertex_id()
    hot = np.array([0, 1, 2, 3, 10, 100], dtype=np.int64)
    compare_graphs(g, og, hot)
