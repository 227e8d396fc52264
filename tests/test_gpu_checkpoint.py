"""Whole-graph checkpoint (the reference has none): DynamicGraph.save / DynamicGraph.load.  The loaded
graph must be indistinguishable from the saved one: getters, block shapes, sampling (recent and uniform, bit for bit
against the oracle), and the result of every LATER add_edges / offload_old_blocks -- the allocator state travels too."""
import numpy as np
import pytest

from helpers import compare_block, compare_graphs, synth_stream
from oracle.oracle import OracleGraph, OracleSampler

pytestmark = pytest.mark.gpu
MB = 1 << 20


def _roots(src, dst, ts, lo, n, rng, num_nodes):
    roots = np.concatenate([src[lo:lo + n], dst[lo:lo + n], rng.integers(0, num_nodes, n)]).astype(np.int64)
    return roots, np.concatenate([ts[lo:lo + n]] * 3).astype(np.float32)


def _check_sampling(g, og, roots, rts, tag):
    from gnnflow_b200 import TemporalSampler
    for strat, kw in (("recent", {}), ("uniform", {}), ("recent", dict(num_snapshots=2, snapshot_time_window=300.0))):
        mf = TemporalSampler(g, [5, 4], strat, **kw).sample(roots, rts)
        om = OracleSampler(og, [5, 4], strat, **kw).sample(roots, rts)
        for l in range(2):
            for k in range(kw.get("num_snapshots", 1)):
                compare_block("%s.%s.l%d.s%d" % (tag, strat, l, k), mf[l][k], om[l][k])


@pytest.mark.parametrize("policy", ["insert", "replace"])
def test_save_load_continue(tmp_path, policy):
    from gnnflow_b200 import DynamicGraph
    src, dst, ts, eid = synth_stream(400, 100, 90000, seed=17, t_max=9000.0)
    cfg = dict(initial_pool_size=2 * MB, maximum_pool_size=512 * MB, mem_resource_type="cuda", minimum_block_size=9,
               blocks_to_preallocate=1024, insertion_policy=policy)
    g, og = DynamicGraph(**cfg), OracleGraph(**cfg)
    B = 4000
    for lo in range(0, 60000, B):
        sl = slice(lo, lo + B)
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl]); og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    assert g.offload_old_blocks(1500.0) == og.offload_old_blocks(1500.0)  # a dropped prefix and free blocks travel too
    path = tmp_path / "graph.ckpt"
    g.save(path)
    before = g.get_memory_breakdown()
    g2 = DynamicGraph.load(path)
    verts = np.array([0, 1, 2, 3, 50, 399, 400, 450, 499])
    rng = np.random.default_rng(5)
    roots, rts = _roots(src, dst, ts, 59000, 300, rng, 500)
    for gg, tag in ((g, "saved"), (g2, "loaded")):
        compare_graphs(gg, og, verts)
        _check_sampling(gg, og, roots, rts, tag)
    after = g2.get_memory_breakdown()
    for k in ("pool", "bump_used", "free", "free_blocks", "vertex_table", "eid_refcounts"):
        assert before[k] == after[k], (k, before, after)
    del g  # the loaded graph does not lean on the saved one's memory
    # both streams go on: the loaded graph takes the same batches as the oracle that never stopped
    for lo in range(60000, 90000, B):
        sl = slice(lo, lo + B)
        g2.add_edges(src[sl], dst[sl], ts[sl], eid[sl]); og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    assert g2.offload_old_blocks(4000.0) == og.offload_old_blocks(4000.0)
    compare_graphs(g2, og, verts)
    roots, rts = _roots(src, dst, ts, 89000, 300, rng, 500)
    _check_sampling(g2, og, roots, rts, "continued")
    # a second generation: save the continued graph, load it, same answers
    g2.save(path)
    g3 = DynamicGraph.load(path)
    compare_graphs(g3, og, verts)
    _check_sampling(g3, og, roots, rts, "second")


def test_load_rejects_foreign_files(tmp_path):
    from gnnflow_b200 import DynamicGraph
    p = tmp_path / "junk.ckpt"
    p.write_bytes(b"not a checkpoint" * 100)
    with pytest.raises((ValueError, RuntimeError)):
        DynamicGraph.load(p)
    with pytest.raises((ValueError, RuntimeError)):
        DynamicGraph.load(tmp_path / "missing.ckpt")
    g = DynamicGraph(initial_pool_size=1 * MB, maximum_pool_size=64 * MB, mem_resource_type="cuda", minimum_block_size=4,
                     blocks_to_preallocate=16, insertion_policy="insert")
    g.add_edges(np.array([0, 1]), np.array([2, 3]), np.array([1.0, 2.0], dtype=np.float32))
    q = tmp_path / "small.ckpt"
    g.save(q)
    raw = q.read_bytes()
    (tmp_path / "cut.ckpt").write_bytes(raw[:len(raw) // 2])
    with pytest.raises((ValueError, RuntimeError)):
        DynamicGraph.load(tmp_path / "cut.ckpt")
    g2 = DynamicGraph.load(q)  # and an empty-ish graph round-trips
    assert g2.num_edges() == 2 and g2.num_vertices() == 4
