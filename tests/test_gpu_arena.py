"""The reclaiming arena (TemporalBlockAllocator::Allocate / Deallocate / Reallocate, temporal_block_allocator.cu:83-180):
memory freed by offload_old_blocks, by replace-policy reallocation, by directory growth and at chunk boundaries is
allocated again -- with the store staying bit-identical to the oracle's throughout."""
import numpy as np
import pytest
import torch

from helpers import compare_graphs
from oracle.oracle import OracleGraph, OracleSampler

pytestmark = pytest.mark.gpu

MB, GB = 1 << 20, 1 << 30


def _stream(num_nodes, n, seed, t_max):
    from gnnflow_b200.synth import synth_stream
    return synth_stream(num_nodes, 0, n, seed=seed, t_max=t_max)


def test_online_window_keeps_device_memory_flat():
    """scripts/online_edge_prediction.py:349-353: add a batch, drop what is older than the window, 200 times.  The
    footprint reached once the window is full must not grow afterwards (the reference frees offloaded blocks through
    rmm; round 1 of this repo only ever bumped a pointer)."""
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    n_iter, batch = 200, 20000
    src, dst, ts, eid = _stream(50000, n_iter * batch, seed=3, t_max=float(n_iter))  # one time unit per batch
    cfg = dict(initial_pool_size=8 * MB, maximum_pool_size=4 * GB, mem_resource_type="cuda", minimum_block_size=16,
               blocks_to_preallocate=1024, insertion_policy="insert")
    g, og = DynamicGraph(**cfg), OracleGraph(**cfg)
    window = 12.0
    sizes, books = [], []
    rng = np.random.default_rng(0)
    for it in range(n_iter):
        sl = slice(it * batch, (it + 1) * batch)
        a = (src[sl], dst[sl], ts[sl], eid[sl])
        g.add_edges(*[torch.from_numpy(x).cuda() for x in a])
        check = it % 40 == 39 or it == n_iter - 1
        if check:
            og.add_edges(*a)
        t_old = float(ts[sl][0]) - window
        nb = g.offload_old_blocks(t_old)
        if check:
            assert nb >= 0
        sizes.append(g.get_device_memory_usage())
        if it % 20 == 19:
            books.append(g.get_memory_breakdown())
        if it % 40 != 39 and it != n_iter - 1:
            og.add_edges(*a)
        og_nb = og.offload_old_blocks(t_old)
        assert nb == og_nb, (it, nb, og_nb)
        if check:
            assert g.num_edges() == og.num_edges()
            roots = rng.integers(0, 50000, 3000).astype(np.int64)
            rts = np.full(3000, float(ts[sl][-1]), dtype=np.float32)
            for strat in ("recent", "uniform"):
                m = TemporalSampler(g, [10], strat).sample(roots, rts)[0][0]
                o = OracleSampler(og, [10], strat).sample(roots, rts)[0][0]
                assert np.array_equal(m.edata['ID'].cpu().numpy(), o["eids"]), (it, strat)
                assert np.array_equal(m.srcdata['ID'].cpu().numpy(), o["all_nodes"]), (it, strat)
    compare_graphs(g, og, np.array([0, 1, 2, 5, 77, 1000, 49999]))
    # everything but the edge-id reference counts is flat once the window is full (iteration ~ 20); the reference counts
    # are a dense table over the LIVE id range, 4 B per id between the oldest id still stored and the newest -- and a
    # vertex that gets an edge now and then keeps its first, partly filled block (hence an early id) alive for the
    # whole run
    flat = [b["pool"] + b["vertex_table"] + b["scratch"] + b["allocator_books"] for b in books]
    assert max(flat) == flat[0], flat
    assert books[-1]["bump_used"] <= books[1]["bump_used"] * 1.03, [b["bump_used"] for b in books]
    assert books[-1]["free_blocks"] < 4 * books[1]["free_blocks"], [b["free_blocks"] for b in books]
    assert books[-1]["eid_refcounts"] <= 2 * 4 * n_iter * batch + (1 << 20)
    assert sizes[-1] == sum(books[-1][k] for k in ("pool", "vertex_table", "eid_refcounts", "scratch", "allocator_books"))
    payload = g.get_graph_memory_usage()
    assert books[-1]["bump_used"] < 3 * payload, (books[-1], payload)


def test_replace_policy_reuses_superseded_payloads():
    """insertion_policy='replace': every overflow re-allocates the vertex's single block and frees the old payload
    (Reallocate, temporal_block_allocator.cu:122-132).  Small batches into hot vertices used to leak O(n^2) slots.  Worst
    case for size-class free lists: the hot vertices grow in lock-step, so nobody ever asks for the sizes they leave
    behind -- the pool stays within 1.5 x the live payload because the host coalesces free neighbours into new bump
    regions (arena_defrag) before it adds a chunk."""
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    n, batch = 1000000, 1000
    rng = np.random.default_rng(1)
    src = rng.integers(0, 40, n).astype(np.int64)  # 40 hot vertices, 25 000 edges each, 1 000 batches
    dst = rng.integers(40, 4000, n).astype(np.int64)
    ts = np.sort(rng.uniform(0, 1e5, n)).astype(np.float32)
    eid = np.arange(n, dtype=np.int64)
    cfg = dict(initial_pool_size=4 * MB, maximum_pool_size=2 * GB, mem_resource_type="cuda", minimum_block_size=8,
               blocks_to_preallocate=1024, insertion_policy="replace")
    g, og = DynamicGraph(**cfg), OracleGraph(**cfg)
    for lo in range(0, n, batch):
        sl = slice(lo, lo + batch)
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    compare_graphs(g, og, np.arange(0, 60))
    live = g.get_graph_memory_usage()  # sum of capacities * 20 B == what is stored (replace: capacity == size)
    assert live == n * 20
    books = g.get_memory_breakdown()
    # without reuse: sum over batches of the running sizes ~ 40 * 20 B * 25 * (1000^2 / 2) = 10 GB; with 1.25 x growth
    # and no coalescing: 5 x the final payloads
    assert books["pool"] < 1.5 * live * 1.03 + 8 * MB, (books, live)
    roots, rts = np.arange(0, 60).astype(np.int64), np.full(60, 2e5, np.float32)
    for strat in ("recent", "uniform"):  # the survivors of all that moving are intact
        m = TemporalSampler(g, [10], strat).sample(roots, rts)[0][0]
        o = OracleSampler(og, [10], strat).sample(roots, rts)[0][0]
        assert np.array_equal(m.edata['ID'].cpu().numpy(), o["eids"]), strat
        assert np.array_equal(m.srcdata['ID'].cpu().numpy(), o["all_nodes"]), strat


def test_gdelt_16m_shape_footprint():
    """GDELT-16.7M-shaped ids: almost every touched vertex owns one minimum-size block (123 slots).  What the store holds
    must stay within 1.3 x sum(capacity) * 20.6 B + the vertex table (round 1: 2.3 x, doubling chunks with stranded
    tails).  Scaled to 2 M edges over 16.7 M ids so that the test stays small."""
    from gnnflow_b200 import DynamicGraph
    src, dst, ts, eid = _stream(16_700_000, 2_000_000, seed=8, t_max=2.6e6)
    cfg = dict(initial_pool_size=64 * MB, maximum_pool_size=64 * GB, mem_resource_type="cuda", minimum_block_size=123,
               blocks_to_preallocate=1024, insertion_policy="insert")
    g = DynamicGraph(**cfg)
    for lo in range(0, len(src), 250000):
        sl = slice(lo, lo + 250000)
        g.add_edges(*[torch.from_numpy(x[sl]).cuda() for x in (src, dst, ts, eid)])
    slots = g.get_graph_memory_usage() / 20.0
    table = (int(g.max_vertex_id()) + 1) * 66
    held = g.get_device_memory_usage()
    nblocks = g.avg_linked_list_length() * g.num_vertices()
    arena_need = slots * 20.6 + nblocks * 32 + g.num_source_vertices() * 128  # payloads + descriptors + directories
    assert held - table * 2 < 1.3 * arena_need + 300 * MB, (held, table, arena_need)


def test_clear_reuses_every_chunk():
    from gnnflow_b200 import DynamicGraph
    src, dst, ts, eid = _stream(3000, 600000, seed=9, t_max=1e4)
    cfg = dict(initial_pool_size=1 * MB, maximum_pool_size=1 * GB, mem_resource_type="cuda", minimum_block_size=32,
               blocks_to_preallocate=1024, insertion_policy="insert")
    g, og = DynamicGraph(**cfg), OracleGraph(**cfg)
    sizes = []
    for rep in range(4):
        g.clear()
        for lo in range(0, len(src), 100000):
            sl = slice(lo, lo + 100000)
            g.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
            if rep == 0:
                og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        sizes.append(g.get_device_memory_usage())
        compare_graphs(g, og, np.array([0, 1, 2, 3, 100, 2999]))
    assert sizes[-1] == sizes[1], sizes  # replays after the first allocate nothing new


def test_edge_ids_int64_window():
    """EIDType is int64 (csrc/common.h:14).  The reference counts live in a dense table over the LIVE id range: ids far
    above 2^31 work, the base follows the ids down (a batch with smaller ids) and up (after offloading), and only a span
    of live ids beyond 2^31 is refused."""
    from gnnflow_b200 import DynamicGraph
    cfg = dict(initial_pool_size=4 * MB, maximum_pool_size=1 * GB, mem_resource_type="cuda", minimum_block_size=8,
               blocks_to_preallocate=1024, insertion_policy="insert")
    g, og = DynamicGraph(**cfg), OracleGraph(**cfg)
    rng = np.random.default_rng(4)
    base = (1 << 40) + 12345
    n = 300000
    src = rng.integers(0, 500, n).astype(np.int64)
    dst = rng.integers(500, 900, n).astype(np.int64)
    ts = np.sort(rng.uniform(0, 3000, n)).astype(np.float32)
    eid = base + np.arange(n, dtype=np.int64)
    for lo in range(100000, 300000, 50000):  # starts in the middle of the id range ...
        sl = slice(lo, lo + 50000)
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl]); og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    # ... then a batch with SMALLER ids (its own vertices, later timestamps: the time order per vertex is kept)
    s2, d2 = src[:100000] + 1000, dst[:100000] + 1000
    t2 = (ts[:100000] + 4000).astype(np.float32)
    g.add_edges(s2, d2, t2, eid[:100000]); og.add_edges(s2, d2, t2, eid[:100000])
    assert g.num_edges() == og.num_edges() == n
    e = g.edges()
    assert e.min() == base and e.max() == base + n - 1 and len(e) == n
    compare_graphs(g, og, np.array([0, 1, 2, 499, 1000, 1499]))
    before = g.get_memory_breakdown()
    assert g.offload_old_blocks(2500.0) == og.offload_old_blocks(2500.0)  # drops most of the first 200 000 ids
    assert g.num_edges() == og.num_edges()
    assert np.array_equal(g.edges(), np.sort(og.edges()))
    after = g.get_memory_breakdown()
    assert after["eid_refcounts"] <= before["eid_refcounts"] and after["pool"] == before["pool"]
    with pytest.raises(ValueError):  # a live span beyond 2^31 ids
        g.add_edges(np.array([3]), np.array([4]), np.array([9000.0], dtype=np.float32), np.array([base + (1 << 32)]))
    with pytest.raises(ValueError):
        g.add_edges(np.array([3]), np.array([4]), np.array([9000.0], dtype=np.float32), np.array([-5]))
    g.add_edges(np.array([3]), np.array([4]), np.array([9000.0], dtype=np.float32), np.array([base + n]))
    assert g.num_edges() == og.num_edges() + 1
