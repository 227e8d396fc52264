"""gf_graph_add_edges_async / gf_graph_flush: batches queued without a host synchronisation build the same store as the
synchronous calls -- checked against the CPU oracle (which follows DynamicGraph::AddEdges, dynamic_graph.cu:77-287) --
including the cases where a queued batch is rejected on the device (table / edge-id / arena growth, unsorted batch,
out-of-order batch) and replayed at the flush."""
import numpy as np
import pytest
import torch

from helpers import assert_same, compare_block, compare_graphs, synth_stream
from oracle.oracle import OracleGraph, OracleSampler

pytestmark = pytest.mark.gpu
GB = 1 << 30
CFG = dict(initial_pool_size=1 << 20, maximum_pool_size=8 * GB, mem_resource_type="cuda", minimum_block_size=6,
           blocks_to_preallocate=64, insertion_policy="insert")


def _dev(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


@pytest.mark.parametrize("policy,batch", [("insert", 700), ("replace", 1900), ("insert", 20000)])
def test_async_ingest_equals_sync(policy, batch):
    from gnnflow_b200 import DynamicGraph
    # ids grow with time so that the vertex table, the edge-id table and the 1 MB arena all have to grow mid-queue
    src, dst, ts, eid = synth_stream(3000, 400, 60000, seed=21, t_max=9000.0)
    ts = np.floor(ts).astype(np.float32)
    grow = (np.arange(len(src)) * 4 // len(src)) * 5000
    src, dst = src + grow, dst + grow
    cfg = {**CFG, "insertion_policy": policy}
    g, og = DynamicGraph(**cfg), OracleGraph(**cfg)
    for i in range(0, len(src), batch):
        sl = slice(i, i + batch)
        og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        g.add_edges_async(*_dev(src[sl], dst[sl], ts[sl], eid[sl]))
    g.flush()
    compare_graphs(g, og, np.unique(np.concatenate([src, dst]))[::7])
    assert_same("edges", g.edges(), og.edges())


def test_async_unsorted_batches_and_implicit_flush():
    """batches shuffled in time (the timestamp sort pass is scheduled by a replay) and no explicit flush: the sampler
    settles the queue itself"""
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    rng = np.random.default_rng(5)
    src, dst, ts, eid = synth_stream(200, 40, 12000, seed=9, t_max=900.0)
    ts = np.floor(ts).astype(np.float32)
    g, og = DynamicGraph(**CFG), OracleGraph(**CFG)
    for i in range(0, len(src), 1000):
        sl = slice(i, i + 1000)
        p = rng.permutation(len(src[sl])) if (i // 1000) % 3 == 1 else np.arange(len(src[sl]))
        s, d, t, e = src[sl][p], dst[sl][p], ts[sl][p], eid[sl][p]
        og.add_edges(s, d, t, e)
        g.add_edges_async(*_dev(s, d, t, e))
    roots = np.concatenate([src[-300:], dst[-300:]]).astype(np.int64)
    rts = np.concatenate([ts[-300:]] * 2).astype(np.float32)
    m = TemporalSampler(g, [7, 3], "recent").sample(roots, rts)  # no flush() before
    om = OracleSampler(og, [7, 3], "recent").sample(roots, rts)
    for l in range(2):
        compare_block("async.l%d" % l, m[l][0], om[l][0])
    compare_graphs(g, og, np.arange(0, 241, 3))


def test_async_error_surfaces_at_flush_and_later_batches_are_dropped():
    from gnnflow_b200 import DynamicGraph
    src, dst, ts, eid = synth_stream(100, 30, 4000, seed=2, t_max=400.0)
    g, og = DynamicGraph(**CFG), OracleGraph(**CFG)
    keep = []
    for k, i in enumerate(range(0, 4000, 500)):
        sl = slice(i, i + 500)
        t = ts[sl] if k != 4 else ts[sl] - 1000.0  # batch 4 is older than what its vertices already hold
        keep.append(_dev(src[sl], dst[sl], t.astype(np.float32), eid[sl]))
        g.add_edges_async(*keep[-1])
        if k < 4:
            og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    with pytest.raises(ValueError):
        g.flush()
    compare_graphs(g, og, np.arange(0, 131))      # batches 0..3 applied, 4 rejected, 5..7 dropped
    g.add_edges(src[2000:2500], dst[2000:2500], ts[2000:2500], eid[2000:2500])  # the graph is usable afterwards
    og.add_edges(src[2000:2500], dst[2000:2500], ts[2000:2500], eid[2000:2500])
    compare_graphs(g, og, np.arange(0, 131))
    with pytest.raises(ValueError):
        g.add_edges_async(src[:10], dst[:10], ts[:10], eid[:10])  # host arrays: not supported
