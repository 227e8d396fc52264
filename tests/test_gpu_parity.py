"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import assert_same, block_to_dict, compare_block, compare_graphs, synth_stream
from oracle.oracle import OracleGraph, OracleSampler

pytestmark = pytest.mark.gpu

GB = 1 << 30
CFG = dict(initial_pool_size=64 << 20, maximum_pool_size=8 * GB, mem_resource_type="cuda",
           blocks_to_preallocate=1024)


def make_graph(**cfg):
    from gnnflow_b200 import DynamicGraph
    return DynamicGraph(**cfg)


def make_sampler(g, fanouts, **kw):
    from gnnflow_b200 import TemporalSampler
    return TemporalSampler(g, fanouts, **kw)


# ------------------------------------------------------------------ the reference's own golden vectors
@pytest.mark.parametrize("name,fn", gc.ALL_STORE, ids=[n for n, _ in gc.ALL_STORE])
def test_store_golden(name, fn):
    fn(make_graph)


@pytest.mark.parametrize("name,fn", gc.ALL_SAMPLER, ids=[n for n, _ in gc.ALL_SAMPLER])
def test_sampler_golden(name, fn):
    fn(make_graph, make_sampler, block_to_dict)


@pytest.mark.parametrize("mem", ["cuda", "unified", "pinned", "shared"])
def test_golden_all_mem_resource_types(mem):  # the reference parameterises its tests over these (:24-25)
    gc.store_sorted(make_graph, mem_resource_type=mem)
    gc.store_add_reverse(make_graph, mem_resource_type=mem)


def test_block_sizing_and_stats():
    g = gc.store_multiple_times(make_graph, "insert")
    s, c, a, b = g.block_shapes(0)
    assert s.tolist() == [4, 2] and c.tolist() == [4, 4]
    assert a.tolist() == [0.0, 4.0] and b.tolist() == [3.0, 5.0]
    assert g.avg_linked_list_length() == pytest.approx(6 / 4)
    assert g.get_graph_memory_usage() == 6 * 4 * 20
    assert g.get_metadata_memory_usage() == 64 * 6 + 8 * 4
    g = gc.store_multiple_times(make_graph, "replace")
    s, c, _, _ = g.block_shapes(0)
    assert s.tolist() == [6] and c.tolist() == [6]


def test_errors():
    g = make_graph(**{**gc.graph_config, "minimum_block_size": 4})
    g.add_edges(np.array([0, 1, 2]), np.array([1, 2, 3]), np.array([0, 1, 2]))
    with pytest.raises(ValueError):  # documented by the reference (dynamic_graph.py:99-101), its test is skipped
        g.add_edges(np.array([2]), np.array([1]), np.array([0]))
    assert g.num_edges() == 3 and g.max_vertex_id() == 3
    with pytest.raises(ValueError):
        g.add_edges(np.array([-1]), np.array([1]), np.array([5]))
    with pytest.raises(ValueError):
        make_graph(**{**gc.graph_config, "insertion_policy": "bogus"})
    with pytest.raises(ValueError):
        make_graph(**{**gc.graph_config, "mem_resource_type": "bogus"})
    with pytest.raises(ValueError):
        make_sampler(g, [2], sample_strategy="bogus")
    with pytest.raises(MemoryError):  # pool exhaustion: reference LOG(FATAL)s (temporal_block_allocator.cu:95-101)
        small = make_graph(**{**gc.graph_config, "initial_pool_size": 1 << 20, "maximum_pool_size": 1 << 20,
                              "minimum_block_size": 1 << 16})
        small.add_edges(np.arange(64), np.arange(64), np.arange(64))
    # still usable after the failures
    g.add_edges(np.array([2]), np.array([1]), np.array([7]))
    assert g.num_edges() == 4


# ------------------------------------------------------------------------- randomized store parity
def _ingest_both(src, dst, ts, eid, batch, add_reverse=False, **cfg):
    g = make_graph(**{**CFG, **cfg})
    og = OracleGraph(**{**CFG, **cfg})
    for i in range(0, len(src), batch):
        sl = slice(i, i + batch)
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl], add_reverse=add_reverse)
        og.add_edges(src[sl], dst[sl], ts[sl], eid[sl], add_reverse=add_reverse)
    return g, og


@pytest.mark.parametrize("policy,adaptive,minblk,batch,rev", [
    ("insert", True, 4, 257, False), ("insert", False, 3, 1000, False), ("replace", True, 8, 333, False),
    ("insert", True, 18, 5000, True), ("insert", True, 1, 64, False), ("insert", True, 64, 20000, False)])
def test_store_random_parity(policy, adaptive, minblk, batch, rev):
    src, dst, ts, eid = synth_stream(300, 50, 20000, seed=3, t_max=5000.0)
    ts = np.floor(ts).astype(np.float32)  # many ties
    g, og = _ingest_both(src, dst, ts, eid, batch, add_reverse=rev, insertion_policy=policy,
                         adaptive_block_size=adaptive, minimum_block_size=minblk)
    compare_graphs(g, og, np.arange(0, 352))
    assert_same("edges", g.edges(), og.edges())


def test_store_unsorted_batches_and_device_input():
    rng = np.random.default_rng(5)
    src, dst, ts, eid = synth_stream(100, 30, 6000, seed=9, t_max=600.0)
    ts = np.floor(ts).astype(np.float32)
    g = make_graph(**{**CFG, "insertion_policy": "insert", "minimum_block_size": 5})
    og = OracleGraph(**{**CFG, "insertion_policy": "insert", "minimum_block_size": 5})
    for i in range(0, len(src), 1500):
        sl = slice(i, i + 1500)
        p = rng.permutation(len(src[sl]))  # shuffled inside the batch: add_edges must sort (stable)
        s, d, t, e = src[sl][p], dst[sl][p], ts[sl][p], eid[sl][p]
        og.add_edges(s, d, t, e)
        dev = torch.device("cuda")
        g.add_edges(torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(t).to(dev),
                    torch.from_numpy(e).to(dev))
    compare_graphs(g, og, np.arange(0, 131))


@pytest.mark.parametrize("shuffle", [False, True], ids=["time_ordered", "shuffled"])
def test_store_million_edge_batches(shuffle):
    """batches above 2^20 edges take the 4096-pair sort tiles, many tiles per digit look-back and multi-CTA stats"""
    rng = np.random.default_rng(11)
    n = 2_600_000
    src, dst, ts, eid = synth_stream(70_000, 9_000, n, seed=21, t_max=40_000.0)
    ts = np.floor(ts).astype(np.float32)  # heavy ties: stability of every pass matters
    cfg = {**CFG, "insertion_policy": "insert", "minimum_block_size": 8, "initial_pool_size": 256 << 20}
    g, og = make_graph(**cfg), OracleGraph(**cfg)
    for lo, hi in ((0, 1_300_000), (1_300_000, n)):
        s, d, t, e = src[lo:hi], dst[lo:hi], ts[lo:hi], eid[lo:hi]
        if shuffle:
            p = rng.permutation(hi - lo)
            s, d, t, e = s[p], d[p], t[p], e[p]
        og.add_edges(s, d, t, e)
        g.add_edges(s, d, t, e)
    verts = np.concatenate([np.arange(0, 64), rng.integers(0, 79_000, 400)])
    compare_graphs(g, og, verts)


@pytest.mark.parametrize("case", ["contiguous_ids", "spaced_ids", "repeated_ids", "wide_table", "queued"])
def test_store_large_batch_bookkeeping(case):
    """batches of 2^20 edges and more keep the vertex flags / edge-id reference counts in a pass of its own
    (ingest_bookkeep_kernel) when the vertex table has at most n / 64 entries: per-CTA bitmaps + merge for the flags, the
    streaming read-modify-write for ids that are one contiguous increasing run, atomics otherwise; wider tables
    (wide_table) leave the upkeep to the apply pass.  nodes(), num_vertices() and num_edges() (distinct edge ids) come from
    exactly that state."""
    n = 1_200_000
    num_src, num_dst = (300_000, 200_000) if case == "wide_table" else (3_000, 2_000)
    src, dst, ts, eid = synth_stream(num_src, num_dst, 2 * n, seed=33, t_max=50_000.0)
    if case == "spaced_ids":
        eid = eid * 3 + 7  # increasing, but not one run: every id needs its own atomic
    elif case == "repeated_ids":
        eid = eid // 2  # every id twice (what add_reverse does): num_edges counts it once
    cfg = {**CFG, "insertion_policy": "insert", "minimum_block_size": 16, "initial_pool_size": 256 << 20}
    g, og = make_graph(**cfg), OracleGraph(**cfg)
    keep = []
    for lo in (0, n):  # the second batch finds most flags set and a table that has its final size
        s, d, t, e = src[lo:lo + n], dst[lo:lo + n], ts[lo:lo + n], eid[lo:lo + n]
        og.add_edges(s, d, t, e)
        if case == "queued":  # add_edges_async: the first batch is rejected on the device (table too small) and replayed
            keep.append([torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (s, d, t, e)])
            g.add_edges_async(*keep[-1])
            continue
        g.add_edges(s, d, t, e)
        assert g.num_edges() == og.num_edges()
        assert_same("nodes", g.nodes(), og.nodes())
    if case == "queued":
        g.flush()
    compare_graphs(g, og, np.arange(0, 48))
    # the same ids again (a stream replayed with later timestamps): no id is new, no vertex is new
    s, d, t, e = src[:n], dst[:n], ts[:n] + np.float32(60_000.0), eid[:n]
    og.add_edges(s, d, t, e)
    g.add_edges(s, d, t, e)
    compare_graphs(g, og, np.arange(0, 48))


def test_default_eids_and_offload():
    src, dst, ts, _ = synth_stream(50, 20, 4000, seed=11, t_max=400.0)
    g, og = None, None
    g = make_graph(**{**CFG, "insertion_policy": "insert", "minimum_block_size": 4})
    og = OracleGraph(**{**CFG, "insertion_policy": "insert", "minimum_block_size": 4})
    for i in range(0, 4000, 500):
        sl = slice(i, i + 500)
        g.add_edges(src[sl], dst[sl], ts[sl], add_reverse=True)
        og.add_edges(src[sl], dst[sl], ts[sl], add_reverse=True)
    assert g.offload_old_blocks(150.0) == og.offload_old_blocks(150.0)
    compare_graphs(g, og, np.arange(0, 71))
    # keep ingesting after the offload
    s2, d2, t2, _ = synth_stream(50, 20, 1000, seed=12, t_max=100.0)
    t2 = t2 + np.float32(400.0)
    g.add_edges(s2, d2, t2)
    og.add_edges(s2, d2, t2)
    assert g.offload_old_blocks(1e9) == og.offload_old_blocks(1e9)  # drops everything
    compare_graphs(g, og, np.arange(0, 71))
    g.add_edges(s2, d2, t2 + np.float32(1000))
    og.add_edges(s2, d2, t2 + np.float32(1000))
    compare_graphs(g, og, np.arange(0, 71))


# ----------------------------------------------------------------------- randomized sampler parity
def _roots(src, dst, ts, lo, hi, num_nodes, rng):
    neg = rng.integers(0, num_nodes, hi - lo)
    return (np.concatenate([src[lo:hi], dst[lo:hi], neg]).astype(np.int64),
            np.concatenate([ts[lo:hi], ts[lo:hi], ts[lo:hi]]).astype(np.float32))


SAMPLER_CASES = [
    dict(fanouts=[10], sample_strategy="recent"),
    dict(fanouts=[5, 3], sample_strategy="recent"),
    dict(fanouts=[10], sample_strategy="uniform"),
    dict(fanouts=[4, 4], sample_strategy="uniform", seed=99),
    dict(fanouts=[3, 2], sample_strategy="recent", num_snapshots=3, snapshot_time_window=40.0),
    dict(fanouts=[3, 2], sample_strategy="uniform", num_snapshots=3, snapshot_time_window=40.0, prop_time=True),
    dict(fanouts=[7], sample_strategy="recent", snapshot_time_window=25.0),
    dict(fanouts=[32], sample_strategy="recent"),
    dict(fanouts=[2, 2, 2], sample_strategy="uniform", prop_time=True),
]


@pytest.mark.parametrize("variant", [0, 1, 3], ids=["warp", "thread", "persistent"])
@pytest.mark.parametrize("case", SAMPLER_CASES, ids=[str(i) for i in range(len(SAMPLER_CASES))])
def test_sampler_random_parity(case, variant):
    src, dst, ts, eid = synth_stream(200, 40, 30000, seed=21, t_max=3000.0)
    ts = np.floor(ts * 4).astype(np.float32) / 4
    g, og = _ingest_both(src, dst, ts, eid, 4000, insertion_policy="insert", minimum_block_size=6)
    s = make_sampler(g, **case)
    s.set_variant(variant)
    os_ = OracleSampler(og, **case)
    rng = np.random.default_rng(1)
    for lo in (0, 3000, 15000, 29400):
        roots, rts = _roots(src, dst, ts, lo, lo + 600, 245, rng)
        mfgs = s.sample(roots, rts)
        omfgs = os_.sample(roots, rts)
        assert len(mfgs) == len(omfgs)
        for l in range(len(mfgs)):
            for k in range(len(mfgs[l])):
                compare_block("case%s.lo%d.l%d.s%d" % (case, lo, l, k), mfgs[l][k], omfgs[l][k])
    # sample_layer on its own + device-resident inputs
    roots, rts = _roots(src, dst, ts, 10000, 10600, 245, rng)
    b = s.sample_layer(torch.from_numpy(roots).cuda(), torch.from_numpy(rts).cuda(), 0, 0)
    compare_block("sample_layer", b, os_.sample_layer(roots, rts, 0, 0))
    assert s.launch_index() == os_._L.og_sampler_launch_index(os_._h)


def test_sampler_deep_history_uniform():
    # one hub vertex with hundreds of blocks: exercises the directory search and the cross-block walks
    n = 40000
    src = np.zeros(n, dtype=np.int64)
    src[::7] = 1
    dst = (np.arange(n) % 97 + 2).astype(np.int64)
    ts = (np.arange(n) // 3).astype(np.float32)
    eid = np.arange(n, dtype=np.int64)
    g, og = _ingest_both(src, dst, ts, eid, 100, insertion_policy="insert", minimum_block_size=4,
                         adaptive_block_size=False)
    assert g.block_shapes(0)[0].shape[0] > 300
    rng = np.random.default_rng(2)
    roots = rng.integers(0, 3, 4000).astype(np.int64)
    rts = rng.uniform(0, n / 3 + 10, 4000).astype(np.float32)
    for case in (dict(fanouts=[10], sample_strategy="uniform"), dict(fanouts=[25], sample_strategy="recent"),
                 dict(fanouts=[8], sample_strategy="uniform", snapshot_time_window=900.0),
                 dict(fanouts=[8], sample_strategy="recent", num_snapshots=2, snapshot_time_window=50.0)):
        for variant in (0, 1, 3):
            s = make_sampler(g, **case)
            s.set_variant(variant)
            os_ = OracleSampler(og, **case)
            m, om = s.sample(roots, rts), os_.sample(roots, rts)
            for k in range(len(m[0])):
                compare_block("deep%s.v%d.s%d" % (case, variant, k), m[0][k], om[0][k])


@pytest.mark.parametrize("policy,minblk,batch", [("replace", 4, 3000), ("insert", 9000, 2500), ("insert", 700, 40000),
                                                 ("insert", 17, 999), ("replace", 130, 1111)])
def test_sampler_big_blocks_pivot_levels(policy, minblk, batch):
    """Few vertices with tens of thousands of edges each: blocks of 10^2 .. 10^4.5 edges, i.e. every depth of the
    pivot hierarchy (gf_common.cuh: blk_lower_bound), partially filled top levels, and the realloc path that rebuilds
    the pivots for a new capacity.  Many duplicate timestamps (ties) and window starts inside blocks."""
    n = 60000
    rng = np.random.default_rng(11)
    src = rng.integers(0, 3, n).astype(np.int64)
    dst = rng.integers(3, 50, n).astype(np.int64)
    ts = np.sort(np.floor(rng.uniform(0, 9000, n))).astype(np.float32)
    eid = np.arange(n, dtype=np.int64)
    g, og = _ingest_both(src, dst, ts, eid, batch, insertion_policy=policy, minimum_block_size=minblk)
    compare_graphs(g, og, np.arange(0, 50))
    roots = rng.integers(0, 4, 3000).astype(np.int64)
    rts = np.concatenate([rng.uniform(-5, 9100, 2000), rng.integers(0, 9001, 1000)]).astype(np.float32)
    for case in (dict(fanouts=[10], sample_strategy="recent"), dict(fanouts=[6], sample_strategy="uniform"),
                 dict(fanouts=[5], sample_strategy="recent", snapshot_time_window=333.0),
                 dict(fanouts=[4], sample_strategy="uniform", num_snapshots=3, snapshot_time_window=1000.5)):
        s, os_ = make_sampler(g, **case), OracleSampler(og, **case)
        m, om = s.sample(roots, rts), os_.sample(roots, rts)
        for k in range(len(m[0])):
            compare_block("big%s.s%d" % (case, k), m[0][k], om[0][k])


def test_sampler_edge_cases():
    src, dst, ts, eid = synth_stream(20, 5, 500, seed=4, t_max=50.0)
    g, og = _ingest_both(src, dst, ts, eid, 100, insertion_policy="insert", minimum_block_size=4)
    s, os_ = make_sampler(g, [3, 3]), OracleSampler(og, [3, 3])
    # empty input (temporal_sampler.cu:107-114)
    m = s.sample(np.array([], dtype=np.int64), np.array([], dtype=np.float32))
    assert m[0][0].num_src_nodes() == 0 and m[1][0].num_dst_nodes() == 0
    # ids beyond the vertex table, vertices without edges, timestamps before any edge, is_static
    roots = np.array([0, 24, 25, 1000000, 3, 3], dtype=np.int64)
    rts = np.array([10, 10, 10, 10, -5, 1e30], dtype=np.float32)
    m, om = s.sample(roots, rts), os_.sample(roots, rts)
    for l in range(2):
        compare_block("edge.l%d" % l, m[l][0], om[l][0])
    st, ost = make_sampler(g, [4], is_static=True), OracleSampler(og, [4], is_static=True)
    compare_block("static", st.sample(roots, rts)[0][0], ost.sample(roots, rts)[0][0])


@pytest.fixture(params=["0", "1"], ids=["all_targets", "listed_targets"])
def compact(request, monkeypatch):
    """multi-batch launches over all targets / over the list of targets whose vertex has out-edges (the library reads
    GNNFLOW_B200_COMPACT_MIN at a sampler's first multi-batch call: 0 = never list, 1 = always)"""
    monkeypatch.setenv("GNNFLOW_B200_COMPACT_MIN", request.param)
    return request.param


@pytest.mark.parametrize("variant", [3, 1], ids=["persistent", "thread"])
def test_sampler_batched_equals_per_batch(variant, compact):
    src, dst, ts, eid = synth_stream(200, 40, 30000, seed=21, t_max=3000.0)
    g, og = _ingest_both(src, dst, ts, eid, 5000, insertion_policy="insert", minimum_block_size=6)
    rng = np.random.default_rng(8)
    for strat in ("recent", "uniform"):
        s = make_sampler(g, [10], sample_strategy=strat)
        s.set_variant(variant)
        os_ = OracleSampler(og, [10], sample_strategy=strat)
        batches = [_roots(src, dst, ts, lo, lo + 600, 245, rng) for lo in range(0, 30000, 600)]
        nodes = torch.from_numpy(np.concatenate([b[0] for b in batches])).cuda()
        tss = torch.from_numpy(np.concatenate([b[1] for b in batches])).cuda()
        offs = torch.from_numpy(np.cumsum([0] + [len(b[0]) for b in batches])).cuda()
        out = s.sample_layer_batched(nodes, tss, offs)
        eo = out["edge_offsets"].cpu().numpy()
        for i, (r, t) in enumerate(batches):
            o = os_.sample_layer(r, t, 0, 0)
            sl = slice(eo[i], eo[i + 1])
            assert_same("b%d.nbr" % i, out["nbr"][sl].cpu().numpy(), o["all_nodes"][len(r):])
            assert_same("b%d.ts" % i, out["ts"][sl].cpu().numpy(), o["all_timestamps"][len(r):])
            assert_same("b%d.dt" % i, out["dt"][sl].cpu().numpy(), o["delta_timestamps"])
            assert_same("b%d.eid" % i, out["eid"][sl].cpu().numpy(), o["eids"])
            assert_same("b%d.row" % i, out["row"][sl].cpu().numpy(), o["row"])
        assert s.launch_index() == len(batches)


@pytest.mark.parametrize("pinned,mode", [(True, 2), (True, 1), (True, 0), (False, 0)],
                         ids=["pinned-inplace", "pinned-mirror", "pinned-auto", "pageable"])
def test_sampler_batched_host_arrays(pinned, mode, compact):
    # one C-ABI call with HOST arrays for a whole replay == the device-array call, element for element
    src, dst, ts, eid = synth_stream(200, 40, 30000, seed=22, t_max=3000.0)
    g, og = _ingest_both(src, dst, ts, eid, 5000, insertion_policy="insert", minimum_block_size=6)
    rng = np.random.default_rng(9)
    batches = [_roots(src, dst, ts, lo, lo + 600, 245, rng) for lo in range(0, 30000, 600)]
    nodes = np.concatenate([b[0] for b in batches])
    tss = np.concatenate([b[1] for b in batches])
    offs = np.cumsum([0] + [len(b[0]) for b in batches])
    for strat in ("recent", "uniform"):
        s = make_sampler(g, [10], sample_strategy=strat)
        s.set_host_output_mode(mode)
        host_out = s.alloc_batched_host_out(len(nodes), len(batches), pinned=pinned)
        h = s.sample_layer_batched_numpy(nodes, tss, offs, out=host_out)
        s.set_launch_index(0)
        d = s.sample_layer_batched(torch.from_numpy(nodes).cuda(), torch.from_numpy(tss).cuda(), torch.from_numpy(offs).cuda())
        eo = d["edge_offsets"].cpu().numpy()
        assert_same("edge_offsets", h["edge_offsets"].astype(np.int64), eo)
        S = int(eo[-1])
        assert S > 0
        for k in ("nbr", "ts", "dt", "eid", "row"):
            assert_same(k, h[k], d[k][:S].cpu().numpy())
        os_ = OracleSampler(og, [10], sample_strategy=strat)
        for i in (0, len(batches) // 2, len(batches) - 1):
            if strat == "uniform":
                os_.set_launch_index(i)
            o = os_.sample_layer(batches[i][0], batches[i][1], 0, 0)
            sl = slice(int(eo[i]), int(eo[i + 1]))
            assert_same("b%d.nbr" % i, h["nbr"][sl], o["all_nodes"][len(batches[i][0]):])
            assert_same("b%d.eid" % i, h["eid"][sl], o["eids"])
            assert_same("b%d.dt" % i, h["dt"][sl], o["delta_timestamps"])
        # 32-bit neighbour ids / rows (gf_sampler_sample_layer_batched_ids32): the same values in 24 B per neighbour
        s.set_launch_index(0)
        out32 = s.alloc_batched_host_out(len(nodes), len(batches), pinned=pinned, ids32=True)
        h32 = s.sample_layer_batched_numpy(nodes, tss, offs, out=out32, ids32=True)
        assert h32["nbr"].dtype == np.uint32 and h32["row"].dtype == np.uint32
        assert_same("ids32.edge_offsets", h32["edge_offsets"].astype(np.int64), eo)
        for k in ("nbr", "row"):
            assert_same("ids32." + k, h32[k].astype(np.int64), h[k])
        for k in ("ts", "dt", "eid"):
            assert_same("ids32." + k, h32[k], h[k])
    # empty replay
    e = s.sample_layer_batched_numpy(np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(3, np.uint64))
    assert len(e["nbr"]) == 0 and list(e["edge_offsets"]) == [0, 0, 0]
    e = s.sample_layer_batched_numpy(np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(3, np.uint64), ids32=True)
    assert len(e["nbr"]) == 0 and e["nbr"].dtype == np.uint32 and list(e["edge_offsets"]) == [0, 0, 0]


def test_uniform_distribution():
    # membership / causality / chi-square on the draw positions (SURVEY 8c acceptance test), independent of the oracle
    n = 600
    src = np.zeros(n, dtype=np.int64)
    dst = np.arange(1, n + 1, dtype=np.int64)
    ts = np.arange(n, dtype=np.float32)
    g = make_graph(**{**CFG, "insertion_policy": "insert", "minimum_block_size": 16})
    for i in range(0, n, 50):
        g.add_edges(src[i:i + 50], dst[i:i + 50], ts[i:i + 50])
    s = make_sampler(g, [20], sample_strategy="uniform", seed=7)
    T = 10000
    b = s.sample(np.zeros(T, dtype=np.int64), np.full(T, 400.5, dtype=np.float32))[0][0]
    nbr = b.srcdata['ID'][T:].cpu().numpy()
    nts = b.srcdata['ts'][T:].cpu().numpy()
    assert len(nbr) == T * 20
    assert nts.max() < 400.5 and nts.min() >= 0
    assert np.array_equal(nbr, nts.astype(np.int64) + 1)
    counts = np.bincount(nts.astype(np.int64), minlength=401)
    assert len(counts) == 401
    exp = T * 20 / 401
    chi2 = ((counts - exp) ** 2 / exp).sum()
    assert chi2 < 400 + 6 * np.sqrt(2 * 400), chi2  # dof = 400; > 6 sigma would be a broken sampler


def test_sample_numpy_host_io():
    """host in / host out through pinned buffers written in place by the kernel, and the pageable-mirror path"""
    import ctypes as C
    from gnnflow_b200 import _lib
    src, dst, ts, eid = synth_stream(200, 40, 30000, seed=21, t_max=3000.0)
    g, og = _ingest_both(src, dst, ts, eid, 5000, insertion_policy="insert", minimum_block_size=6)
    rng = np.random.default_rng(4)
    for case in (dict(fanouts=[10], sample_strategy="recent"), dict(fanouts=[3, 3], sample_strategy="uniform"),
                 dict(fanouts=[2, 2], sample_strategy="recent", num_snapshots=2, snapshot_time_window=50.0)):
        for variant in (3, 1):
            s, os_ = make_sampler(g, **case), OracleSampler(og, **case)
            s.set_variant(variant)
            for lo in (5000, 29000, 100):
                roots, rts = _roots(src, dst, ts, lo, lo + 400, 245, rng)
                res = s.sample_numpy(roots, rts)
                ores = os_.sample(roots, rts)[::-1]
                for l in range(len(res)):
                    for k in range(len(res[l])):
                        r, o = res[l][k], ores[l][k]
                        for key in ("all_nodes", "all_timestamps", "delta_timestamps", "eids", "row", "col"):
                            assert_same("numpy.%s.l%d.s%d" % (key, l, k), r[key], o[key])
    # pageable host outputs (mirror + D2H) through the raw C ABI
    L = _lib.lib()
    s, os_ = make_sampler(g, [5]), OracleSampler(og, [5])
    roots, rts = _roots(src, dst, ts, 20000, 20300, 245, rng)
    T = len(roots)
    an, at = np.zeros(T * 6, np.int64), np.zeros(T * 6, np.float32)
    dt, ei, ro, co = np.zeros(T * 5, np.float32), np.zeros(T * 5, np.int64), np.zeros(T * 5, np.int64), np.zeros(T * 5, np.int64)
    r = _lib.SamplingResultC(an.ctypes.data, at.ctypes.data, dt.ctypes.data, ei.ctypes.data, ro.ctypes.data,
                             co.ctypes.data, T, 0, 0)
    _lib.check(L.gf_sampler_sample_layer(s._h, roots.ctypes.data, rts.ctypes.data, T, 0, 0, C.byref(r), 0, 0, None))
    o = os_.sample_layer(roots, rts, 0, 0)
    S = int(r.num_edges)
    assert S == len(o["eids"]) and int(r.num_dst) == T
    assert_same("pageable.all_nodes", an[:T + S], o["all_nodes"])
    assert_same("pageable.dt", dt[:S], o["delta_timestamps"])
    assert_same("pageable.row", ro[:S], o["row"])
    assert_same("pageable.col", co[:S], o["col"])


def test_sampler_batched_two_layers_chained_on_device(compact):
    """gf_sampler_chain_batched + the batched launch: both layers of a whole replay == the oracle batch by batch
    (layer 1 samples every batch's [roots || neighbours], temporal_sampler.cu:242-262, 279-305)"""
    src, dst, ts, eid = synth_stream(200, 40, 30000, seed=23, t_max=3000.0)
    src2, dst2 = np.concatenate([src, dst]), np.concatenate([dst, src])  # undirected: the second layer finds neighbours
    ts2, eid2 = np.concatenate([ts, ts]), np.concatenate([eid, eid])
    o = np.argsort(ts2, kind="stable")
    g, og = _ingest_both(src2[o], dst2[o], ts2[o], eid2[o], 5000, insertion_policy="insert", minimum_block_size=6)
    rng = np.random.default_rng(12)
    all_batches = [_roots(src, dst, ts, lo, lo + 600, 245, rng) for lo in range(0, 30000, 1500)]
    for strat in ("recent", "uniform"):
        batches = list(all_batches)
        if strat == "recent":  # an empty batch in the middle (uniform: batch b draws from RNG launch index base + b,
            batches.insert(3, (np.zeros(0, np.int64), np.zeros(0, np.float32)))  # the oracle skips empty launches)
        nodes = torch.from_numpy(np.concatenate([b[0] for b in batches])).cuda()
        tss = torch.from_numpy(np.concatenate([b[1] for b in batches])).cuda()
        offs = torch.from_numpy(np.cumsum([0] + [len(b[0]) for b in batches])).cuda()
        s = make_sampler(g, [5, 4], sample_strategy=strat)
        os_ = OracleSampler(og, [5, 4], sample_strategy=strat)
        layers = s.sample_batched(nodes, tss, offs)
        assert len(layers) == 2
        o0 = [os_.sample_layer(r, t, 0, 0) for r, t in batches]                      # launch indices 0 .. nb-1
        o1 = [os_.sample_layer(a["all_nodes"], a["all_timestamps"], 1, 0) for a in o0]  # nb .. 2 nb - 1
        for l, oo in ((0, o0), (1, o1)):
            L = layers[l]
            bo, eo = L["batch_offsets"].cpu().numpy(), L["edge_offsets"].cpu().numpy()
            for i, ob in enumerate(oo):
                nt = len(ob["all_nodes"]) - len(ob["eids"])  # targets of this batch in this layer
                assert bo[i + 1] - bo[i] == nt
                assert_same("%s.l%d.b%d.targets" % (strat, l, i), L["nodes"][bo[i]:bo[i + 1]].cpu().numpy(), ob["all_nodes"][:nt])
                assert_same("%s.l%d.b%d.tts" % (strat, l, i), L["timestamps"][bo[i]:bo[i + 1]].cpu().numpy(),
                            ob["all_timestamps"][:nt])
                sl = slice(eo[i], eo[i + 1])
                assert_same("%s.l%d.b%d.nbr" % (strat, l, i), L["nbr"][sl].cpu().numpy(), ob["all_nodes"][nt:])
                assert_same("%s.l%d.b%d.ts" % (strat, l, i), L["ts"][sl].cpu().numpy(), ob["all_timestamps"][nt:])
                assert_same("%s.l%d.b%d.dt" % (strat, l, i), L["dt"][sl].cpu().numpy(), ob["delta_timestamps"])
                assert_same("%s.l%d.b%d.eid" % (strat, l, i), L["eid"][sl].cpu().numpy(), ob["eids"])
                assert_same("%s.l%d.b%d.row" % (strat, l, i), L["row"][sl].cpu().numpy(), ob["row"])


@pytest.mark.parametrize("policy", ["insert", "replace"])
def test_sample_after_offload_and_directory_growth(policy):
    """Regression (round-1 advisor finding): blocks dropped by offload_old_blocks must not count as candidates once the
    vertex's directory has been re-allocated.  Ingest -> partial offload -> ingest until the directories grow ->
    sample recent / uniform with the default window of 0 and with windows that start inside the dropped range; the
    sliding-window loop of scripts/online_edge_prediction.py:349-353."""
    n = 60000
    rng = np.random.default_rng(5)
    src = rng.integers(0, 6, n).astype(np.int64)
    dst = rng.integers(6, 60, n).astype(np.int64)
    ts = np.sort(np.floor(rng.uniform(0, 6000, n))).astype(np.float32)
    eid = np.arange(n, dtype=np.int64)
    cfg = dict(insertion_policy=policy, minimum_block_size=8, adaptive_block_size=False)
    g, og = make_graph(**{**CFG, **cfg}), OracleGraph(**{**CFG, **cfg})
    roots = rng.integers(0, 8, 2000).astype(np.int64)
    cases = (dict(fanouts=[10], sample_strategy="recent"), dict(fanouts=[10], sample_strategy="uniform"),
             dict(fanouts=[4, 3], sample_strategy="uniform", snapshot_time_window=700.0),
             dict(fanouts=[5], sample_strategy="recent", num_snapshots=2, snapshot_time_window=450.0))
    step = 200
    for it, lo in enumerate(range(0, n, step)):
        sl = slice(lo, lo + step)
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        if it % 25 == 24:  # slide the window: drop what is older than 800 time units, then keep ingesting
            t_now = float(ts[lo + step - 1])
            assert g.offload_old_blocks(t_now - 800.0) == og.offload_old_blocks(t_now - 800.0)
        if it % 50 == 49 or lo + step >= n:
            t_now = float(ts[lo + step - 1])
            rts = rng.uniform(t_now - 1200, t_now + 5, len(roots)).astype(np.float32)
            for case in cases:
                for variant in (3, 1):
                    s, os_ = make_sampler(g, **case), OracleSampler(og, **case)
                    s.set_variant(variant)
                    m, om = s.sample(roots, rts), os_.sample(roots, rts)
                    for l in range(len(m)):
                        for k in range(len(m[l])):
                            compare_block("offgrow.%s.it%d.v%d.l%d.s%d" % (case, it, variant, l, k), m[l][k], om[l][k])
    compare_graphs(g, og, np.arange(0, 60))
    assert g.block_shapes(0)[0].shape[0] < 400  # old blocks really were dropped


@pytest.mark.parametrize("shape", ["bursty", "geometric_blocks"])
def test_sampler_directory_search_skewed(shape):
    """The directory searches interpolate (by time for the window bounds, by position for uniform draws) and fall back
    to bisection: histories where interpolation guesses badly -- edges bunched in bursts with long silences, and block
    sizes growing geometrically -- must give the oracle's answer all the same."""
    rng = np.random.default_rng(17)
    n = 50000
    if shape == "bursty":
        # 50 bursts of 1000 edges, burst k at time 2^(k/3): almost all blocks cover a vanishing part of the time axis
        ts = np.repeat(2.0 ** (np.arange(50) / 3.0), 1000) + np.tile(np.arange(1000) * 1e-4, 50)
        batch = 250
    else:
        ts = np.sort(rng.uniform(0, 1e4, n))
        batch = None
    ts = np.sort(ts).astype(np.float32)
    src = (rng.random(n) < 0.1).astype(np.int64)  # vertex 0: 90 % of the edges
    dst = rng.integers(2, 90, n).astype(np.int64)
    eid = np.arange(n, dtype=np.int64)
    cfg = dict(insertion_policy="insert", minimum_block_size=4, adaptive_block_size=(shape != "bursty"))
    g, og = make_graph(**{**CFG, **cfg}), OracleGraph(**{**CFG, **cfg})
    lo = 0
    k = 0
    while lo < n:
        b = batch if batch else int(4 * 1.25 ** k) + 1  # geometric: batch sizes (hence block sizes) grow by 25 %
        sl = slice(lo, min(n, lo + b))
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        og.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        lo += b
        k += 1
    assert g.block_shapes(0)[0].shape[0] > 25
    roots = rng.integers(0, 3, 6000).astype(np.int64)
    tmax = float(ts[-1])
    rts = np.concatenate([rng.uniform(0, tmax * 1.01, 3000), rng.choice(ts, 3000)]).astype(np.float32)
    w = tmax / 7
    for case in (dict(fanouts=[10], sample_strategy="uniform"), dict(fanouts=[12], sample_strategy="recent"),
                 dict(fanouts=[6], sample_strategy="uniform", snapshot_time_window=w),
                 dict(fanouts=[5], sample_strategy="recent", num_snapshots=3, snapshot_time_window=w / 2),
                 dict(fanouts=[4, 3], sample_strategy="uniform", num_snapshots=2, snapshot_time_window=w)):
        s, os_ = make_sampler(g, **case), OracleSampler(og, **case)
        m, om = s.sample(roots, rts), os_.sample(roots, rts)
        for l in range(len(m)):
            for k in range(len(m[l])):
                compare_block("skew.%s.%s.l%d.s%d" % (shape, case, l, k), m[l][k], om[l][k])
